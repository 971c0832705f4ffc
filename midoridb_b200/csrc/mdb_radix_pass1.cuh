// mdb_radix_pass1.cuh - pass 1 of the radix join: stream the keys once, append their 2-byte remainders to the
// per-partition streams (mdb_radix_types.cuh).  Included by mdb_radix.cu.
//
// One persistent 1024-thread CTA per SM.  Per round every thread takes 8 keys:
//   insert   partition = (key - kmin) >> shift; ONE shared-memory atomic hands out a slot of the partition's
//            40-byte staging row, ONE 2-byte shared store writes the remainder.  The lane that fills slot 16 puts
//            the partition on ITS WARP's worklist (position from a ballot: no atomics, no CTA-wide queue);
//   barrier  (all remainders of the queued rows are in shared memory)
//   flush    every warp flushes its own worklist, one lane per row: one global atomic on the partition's cursor
//            gives the position, the first 16 remainders leave as ONE 256-bit store = one aligned 32-byte sector;
//   barrier
// Measurements behind this shape (profiles/microbench/p1_lab*.cu and profiles/p1_lab_*.txt, B200, 2^28 keys):
//   * 256-bit key loads stream at 6.4 TB/s where 128-bit loads reach 4.7;
//   * a CTA that needs more than 195 KiB of shared memory pushes the SM into its largest carve-out, the L1 that
//     is left cannot hold the key loads in flight and the whole kernel loses 25%;
//   * slot atomics + 2-byte stores cost 0.05 ms on top of the loads; a CTA-wide worklist fed by same-address
//     atomics costs 0.15 ms more than ballots; stores that complete whole 128-byte lines are 0.13 ms cheaper than
//     scattered 32-byte sectors (hence shared per-partition streams instead of per-CTA chunks).

#define RJ_P1_WARPS (RJ_P1_THREADS / 32)
#define RJ_P1_KEYS 8               // keys per thread per round
#define RJ_WL_CAP 128              // rows one warp may complete per round (16 expected, four times that while all rows fill in step at start-up)
#define RJ_TAIL_ROUNDS 64          // rounds spent on parked keys after the last tile before giving up (skew)

#define RJ_HINT_PREFETCH 1u        // prefetch.global.L2 two tiles ahead
#define RJ_HINT_LOAD_EVICT_FIRST 2u
#define RJ_HINT_STORE_EVICT_LAST 4u
#define RJ_HINT_DEFAULT 0u         // (MDBCU_P1_HINTS overrides; none of them pays once the L1 is large enough)

struct RJP1Smem {
	uint16_t stage[RJ_MAX_PART * RJ_CAP];     // 160 KiB: 20 two-byte slots per partition
	uint32_t fill[RJ_MAX_PART];               // slots handed out since the last flush (may overshoot RJ_CAP)
	uint16_t worklist[RJ_P1_WARPS][RJ_WL_CAP]; // per warp: partitions whose 16th slot it filled this round
	uint32_t ovf[2][RJ_OVF_CAP];              // (partition << 16 | remainder) waiting for the next round
	uint32_t ovf_count[2];
};

static_assert(sizeof(RJP1Smem) <= 195 * 1024, "pass-1 shared memory must stay inside the 196 KiB carve-out");

// plain shared-memory atomic: kept in PTX so the compiler does not expand it into warp-aggregation code
__device__ __forceinline__ uint32_t rj_smem_add(uint32_t *p, uint32_t v)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
	return old;
}

// hand out the next slot of partition p's staging row
__device__ __forceinline__ uint32_t rj_fill_claim(RJP1Smem *sm, uint32_t p)
{
	return rj_smem_add(&sm->fill[p], 1u);
}

__device__ __forceinline__ uint32_t rj_lanemask_lt()
{
	uint32_t m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

// one 32-byte sector in one store instruction (sm_100: 256-bit global accesses)
__device__ __forceinline__ void rj_store_sector(void *dst, uint2 a, uint2 b, uint2 c, uint2 d, bool evict_last)
{
	if (evict_last) {
		uint64_t pol;
		asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
		asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(b.x),
				"r"(b.y), "r"(c.x), "r"(c.y), "r"(d.x), "r"(d.y), "l"(pol) : "memory");
	} else {
		asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(b.x), "r"(b.y), "r"(c.x),
				"r"(c.y), "r"(d.x), "r"(d.y) : "memory");
	}
}

// low words of the four keys in 32 bytes (the caller guarantees key - kmin < 2^32)
__device__ __forceinline__ void rj_load_keys256(const void *p, uint32_t *lo, bool evict_first)
{
	uint32_t t[8];
	if (evict_first) {
		uint64_t pol;
		asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
		asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
				: "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "l"(p), "l"(pol));
	} else {
		asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				: "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "l"(p));
	}
	lo[0] = t[0];
	lo[1] = t[2];
	lo[2] = t[4];
	lo[3] = t[6];
}

__device__ static inline void rj_park(RJP1Smem *sm, const RJParams &pr, uint32_t item, int par)
{
	// staging row full until this round's flush: the key waits one round
	const uint32_t o = rj_smem_add(&sm->ovf_count[par], 1u);
	if (o < RJ_OVF_CAP)
		sm->ovf[par][o] = item;
	else
		atomicOr(pr.error_flag, RJ_ERR_SKEW);
}

// Warp-synchronous insert of one item per lane (lanes without a key pass valid = false): slot, store, and - for
// the lane that completed a sector - a place on the warp's worklist.
__device__ __forceinline__ void rj_insert_ws(RJP1Smem *sm, const RJParams &pr, uint32_t item, bool valid, int par, uint32_t &wl_n)
{
	const uint32_t p = item >> 16;
	uint32_t pos = RJ_NONE;
	if (valid)
		pos = rj_fill_claim(sm, p);
	if (pos < RJ_CAP)
		sm->stage[p * RJ_CAP + pos] = (uint16_t)item;
	const bool done = pos == RJ_FLUSH - 1;
	const uint32_t bal = __ballot_sync(0xffffffffu, done);
	if (done) {
		const uint32_t idx = wl_n + __popc(bal & rj_lanemask_lt());
		if (idx < RJ_WL_CAP)
			sm->worklist[threadIdx.x >> 5][idx] = (uint16_t)p;
	}
	wl_n += __popc(bal);
	if (valid && pos >= RJ_CAP)
		rj_park(sm, pr, item, par);
}

// keys parked by the previous round are inserted again (their rows were flushed since)
__device__ static inline void rj_reinsert_parked(const RJParams &pr, RJP1Smem *sm, int par, uint32_t &wl_n)
{
	const uint32_t tid = threadIdx.x;
	const uint32_t novf = min(sm->ovf_count[par ^ 1], (uint32_t)RJ_OVF_CAP);
	for (uint32_t i0 = tid & ~31u; i0 < novf; i0 += RJ_P1_THREADS) { // warp-uniform trip count
		const uint32_t i = i0 + (tid & 31u);
		const bool valid = i < novf;
		rj_insert_ws(sm, pr, valid ? sm->ovf[par ^ 1][i] : 0u, valid, par, wl_n);
	}
}

// 8 keys per thread.  PACKED: item = (partition << 16 | remainder), RJ_NONE = no key (generic kernel);
// otherwise item = key - kmin and every item is a key (lean kernel).
// All slot requests of a thread are issued back to back (independent shared-memory atomics), then consumed.
template <bool PACKED>
__device__ __forceinline__ void rj_insert_items(RJP1Smem *sm, const RJParams &pr, const uint32_t *item, int par, uint32_t &wl_n)
{
	constexpr bool ALL_VALID = !PACKED;
	const int pshift = PACKED ? 16 : pr.shift;
	uint32_t pos[RJ_P1_KEYS];
#pragma unroll
	for (int k = 0; k < RJ_P1_KEYS; k++)
		pos[k] = (ALL_VALID || item[k] != RJ_NONE) ? rj_fill_claim(sm, item[k] >> pshift) : RJ_NONE;
	const uint32_t lt = rj_lanemask_lt();
	uint16_t *wl = sm->worklist[threadIdx.x >> 5];
	uint32_t worst = 0; // largest slot handed to this thread (+1 with holes, so that RJ_NONE counts as 0)
#pragma unroll
	for (int k = 0; k < RJ_P1_KEYS; k++) {
		const uint32_t p = item[k] >> pshift;
		if (pos[k] < RJ_CAP)
			sm->stage[p * RJ_CAP + pos[k]] = (uint16_t)(PACKED ? item[k] : item[k] & pr.mask);
		const bool done = pos[k] == RJ_FLUSH - 1;
		const uint32_t bal = __ballot_sync(0xffffffffu, done);
		if (done) {
			const uint32_t idx = wl_n + __popc(bal & lt);
			if (idx < RJ_WL_CAP)
				wl[idx] = (uint16_t)p;
		}
		wl_n += __popc(bal);
		worst = max(worst, ALL_VALID ? pos[k] : pos[k] + 1u);
	}
	if (worst >= (ALL_VALID ? RJ_CAP : RJ_CAP + 1u)) { // rare: some row was full
#pragma unroll
		for (int k = 0; k < RJ_P1_KEYS; k++)
			if (pos[k] >= RJ_CAP && (ALL_VALID || item[k] != RJ_NONE))
				rj_park(sm, pr, PACKED ? item[k] : (((item[k] >> pshift) << 16) | (item[k] & pr.mask)), par);
	}
}

// barrier, every warp flushes the rows on its own worklist, barrier
__device__ static inline void rj_round_end(const RJSide &s, const RJParams &pr, RJP1Smem *sm, int par, uint32_t wl_n)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	const uint16_t *wl = sm->worklist[tid >> 5];
	__syncthreads();
	if (tid == 0)
		sm->ovf_count[par ^ 1] = 0; // the other parity's parked keys were re-inserted during this round
	if (wl_n > RJ_WL_CAP) {
		if (lane == 0)
			atomicOr(pr.error_flag, RJ_ERR_SKEW);
		wl_n = RJ_WL_CAP;
	}
	const bool evict_last = (s.hints & RJ_HINT_STORE_EVICT_LAST) != 0;
	for (uint32_t w = lane; w < wl_n; w += 32) {
		const uint32_t p = wl[w];
		const uint32_t at = atomicAdd(&s.cursor[p], (uint32_t)RJ_FLUSH);
		uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * RJ_CAP]); // 40-byte rows are 8-byte aligned
		const uint32_t have = sm->fill[p];
		const uint2 a = row[0], b = row[1], c = row[2], d = row[3], e = row[4];
		row[0] = e; // keep the (at most 4) remainders behind the flushed sector
		sm->fill[p] = min(have, (uint32_t)RJ_CAP) - RJ_FLUSH;
		if (at + RJ_FLUSH <= s.cap)
			rj_store_sector(s.stream + (size_t)p * s.cap + at, a, b, c, d, evict_last);
		else
			atomicOr(pr.error_flag, RJ_ERR_STREAM);
	}
	__syncthreads();
}

// every partition's partial sector goes to the partition's tail stream
__device__ static inline void rj_drain(const RJSide &s, const RJParams &pr, RJP1Smem *sm)
{
	for (int p = threadIdx.x; p < pr.nparts; p += RJ_P1_THREADS) {
		const uint32_t f = min(sm->fill[p], (uint32_t)RJ_CAP);
		if (f == 0)
			continue;
		const uint32_t at = atomicAdd(&s.tail_cursor[p], f);
		if (at + f <= s.tail_cap) {
			uint16_t *dst = s.tail + (size_t)p * s.tail_cap + at;
			for (uint32_t i = 0; i < f; i++)
				dst[i] = sm->stage[p * RJ_CAP + i];
		} else {
			atomicOr(pr.error_flag, RJ_ERR_STREAM);
		}
	}
}

__device__ static inline void rj_smem_init(RJP1Smem *sm)
{
	const int tid = threadIdx.x;
	for (int p = tid; p < RJ_MAX_PART; p += RJ_P1_THREADS)
		sm->fill[p] = 0;
	if (tid < 2)
		sm->ovf_count[tid] = 0;
	__syncthreads();
}

// rounds for the keys still parked after the last tile (bounded: a row that never drains means extreme skew)
__device__ static inline void rj_finish(const RJSide &s, const RJParams &pr, RJP1Smem *sm, int par)
{
	for (int guard = 0; sm->ovf_count[par ^ 1] != 0; guard++) { // block-uniform: written before the last barrier
		if (guard == RJ_TAIL_ROUNDS) {
			if (threadIdx.x == 0)
				atomicOr(pr.error_flag, RJ_ERR_SKEW);
			break;
		}
		uint32_t wl_n = 0;
		rj_reinsert_parked(pr, sm, par, wl_n);
		rj_round_end(s, pr, sm, par, wl_n);
		par ^= 1;
	}
	rj_drain(s, pr, sm);
}

// Pass 1, lean variant: column without NULLs/tombstones whose [min, max] lies inside the partitioned range and
// key - kmin < 2^32: no per-key validity test, 32-bit arithmetic on the low words, 256-bit key loads
// (double-buffered in registers: 8 registers per tile in flight).  The ragged tail (< one tile) goes through CTA 0.
__global__ void __launch_bounds__(RJ_P1_THREADS, 1) k_radix_partition_fast(RJSide s, RJParams pr)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RJP1Smem *sm = reinterpret_cast<RJP1Smem*>(smem_raw);
	rj_smem_init(sm);

	constexpr uint32_t TILE = RJ_P1_THREADS * RJ_P1_KEYS;
	const uint32_t tid = threadIdx.x;
	const uint64_t nfull = s.n / TILE;
	const uint32_t kmin_lo = (uint32_t)(unsigned long long)pr.kmin;
	const bool pf = (s.hints & RJ_HINT_PREFETCH) != 0, evict_first = (s.hints & RJ_HINT_LOAD_EVICT_FIRST) != 0;
	uint32_t buf_a[RJ_P1_KEYS], buf_b[RJ_P1_KEYS];
	int par = 0;
	auto load = [&](uint64_t tile, uint32_t *dst) {
		if (pf) {
			// pull the tile this CTA will load two rounds from now into L2 (one 128-byte line per thread)
			const uint64_t pf_first = (tile + 2ull * gridDim.x) * TILE + (uint64_t)tid * 16;
			if (tid < TILE / 16 && pf_first + 16 <= s.n)
				asm volatile("prefetch.global.L2 [%0];" ::"l"(s.keys + pf_first));
		}
		const char *t = reinterpret_cast<const char*>(s.keys + tile * TILE) + tid * 32u;
		rj_load_keys256(t, dst, evict_first);
		rj_load_keys256(t + RJ_P1_THREADS * 32u, dst + 4, evict_first);
	};
	// a round = one tile (8 keys per thread) and one flush.  (Two tiles per flush halve the barriers but park 25x
	// more keys: measured 4% slower.)
	uint32_t wl_n = 0;
	auto insert = [&](const uint32_t *buf) {
		uint32_t item[RJ_P1_KEYS];
#pragma unroll
		for (int k = 0; k < RJ_P1_KEYS; k++)
			item[k] = buf[k] - kmin_lo;
		rj_insert_items<false>(sm, pr, item, par, wl_n);
	};
	auto end_round = [&]() {
		rj_reinsert_parked(pr, sm, par, wl_n);
		rj_round_end(s, pr, sm, par, wl_n);
		par ^= 1;
		wl_n = 0;
	};
	uint64_t tile = blockIdx.x;
	if (tile < nfull)
		load(tile, buf_a);
	while (tile < nfull) {
		uint64_t next = tile + gridDim.x;
		if (next < nfull)
			load(next, buf_b);
		insert(buf_a);
		end_round();
		tile = next;
		if (tile >= nfull)
			break;
		next = tile + gridDim.x;
		if (next < nfull)
			load(next, buf_a);
		insert(buf_b);
		end_round();
		tile = next;
	}
	if (blockIdx.x == 0 && nfull * TILE != s.n) {
		for (uint64_t r0 = nfull * TILE + (tid & ~31u); r0 < s.n; r0 += RJ_P1_THREADS) { // warp-uniform trip count
			const uint64_t r = r0 + (tid & 31u);
			const bool valid = r < s.n;
			const uint32_t d = valid ? (uint32_t)(unsigned long long)s.keys[r] - kmin_lo : 0u;
			rj_insert_ws(sm, pr, ((d >> pr.shift) << 16) | (d & pr.mask), valid, par, wl_n);
		}
		rj_reinsert_parked(pr, sm, par, wl_n);
		rj_round_end(s, pr, sm, par, wl_n);
		par ^= 1;
	}
	rj_finish(s, pr, sm, par);
}

// generic tile load: 128-bit loads of whole 64-bit keys (range test needs the high words)
__device__ static inline void rj_load_tile(const RJSide &s, uint64_t tile, int4 *dst)
{
	constexpr uint32_t TILE = RJ_P1_THREADS * RJ_P1_KEYS;
	const int4 *src = reinterpret_cast<const int4*>(s.keys);
	const uint64_t npairs = s.n / 2;
	const uint64_t base_pair = tile * (TILE / 2);
	if (base_pair + TILE / 2 <= npairs) {
#pragma unroll
		for (int j = 0; j < RJ_P1_KEYS / 2; j++)
			dst[j] = mdb_ldg_stream(src + base_pair + (uint64_t)j * RJ_P1_THREADS + threadIdx.x);
	} else {
#pragma unroll
		for (int j = 0; j < RJ_P1_KEYS / 2; j++) {
			const uint64_t pi = base_pair + (uint64_t)j * RJ_P1_THREADS + threadIdx.x;
			if (pi < npairs) {
				dst[j] = mdb_ldg_stream(src + pi);
			} else if (pi == npairs && (s.n & 1)) {
				const unsigned long long last = (unsigned long long)s.keys[s.n - 1];
				dst[j] = make_int4((int)(unsigned)last, (int)(unsigned)(last >> 32), 0, 0);
			} else {
				dst[j] = make_int4(0, 0, 0, 0);
			}
		}
	}
}

// generic insert phase: range test, NULL/tombstone bitmap, ragged last tile.  FULL: every row of the tile exists
template <bool HAS_PRESENT, bool FULL>
__device__ static inline void rj_insert_tile(const RJSide &s, const RJParams &pr, RJP1Smem *sm, const int4 *buf, uint64_t tile,
		int par, uint32_t &wl_n)
{
	constexpr uint32_t TILE = RJ_P1_THREADS * RJ_P1_KEYS;
	const uint64_t base_pair = tile * (TILE / 2);
	uint32_t item[RJ_P1_KEYS];
#pragma unroll
	for (int j = 0; j < RJ_P1_KEYS / 2; j++) {
		const uint64_t pi = base_pair + (uint64_t)j * RJ_P1_THREADS + threadIdx.x;
		const unsigned long long k0 = ((unsigned long long)(unsigned)buf[j].y << 32) | (unsigned)buf[j].x;
		const unsigned long long k1 = ((unsigned long long)(unsigned)buf[j].w << 32) | (unsigned)buf[j].z;
		const unsigned long long d0 = k0 - (unsigned long long)pr.kmin, d1 = k1 - (unsigned long long)pr.kmin;
		bool ok0 = d0 < pr.range, ok1 = d1 < pr.range;
		if (!FULL) {
			ok0 = ok0 && pi * 2 < s.n;
			ok1 = ok1 && pi * 2 + 1 < s.n;
		}
		if (HAS_PRESENT) {
			const uint32_t pw = (FULL || pi * 2 < s.n) ? (s.present[pi >> 4] >> ((pi & 15) * 2)) : 0u;
			ok0 = ok0 && (pw & 1u);
			ok1 = ok1 && (pw & 2u);
		}
		item[2 * j] = ok0 ? ((((uint32_t)d0 >> pr.shift) << 16) | ((uint32_t)d0 & pr.mask)) : RJ_NONE;
		item[2 * j + 1] = ok1 ? ((((uint32_t)d1 >> pr.shift) << 16) | ((uint32_t)d1 & pr.mask)) : RJ_NONE;
	}
	rj_insert_items<true>(sm, pr, item, par, wl_n);
}

// Pass 1, generic variant (NULLs / tombstones / keys outside the partitioned range): same rounds as the lean
// kernel, whole keys double-buffered in registers (ping-pong, no copies).
template <bool HAS_PRESENT>
__global__ void __launch_bounds__(RJ_P1_THREADS, 1) k_radix_partition(RJSide s, RJParams pr)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RJP1Smem *sm = reinterpret_cast<RJP1Smem*>(smem_raw);
	rj_smem_init(sm);

	constexpr uint32_t TILE = RJ_P1_THREADS * RJ_P1_KEYS;
	const uint64_t ntiles = (s.n + TILE - 1) / TILE;
	const uint64_t nfull = s.n / TILE; // tiles [0, nfull) are complete
	int4 buf_a[RJ_P1_KEYS / 2], buf_b[RJ_P1_KEYS / 2];
	uint64_t tile = blockIdx.x;
	int par = 0;
	if (tile < ntiles)
		rj_load_tile(s, tile, buf_a);
	auto round = [&](const int4 *buf, uint64_t t) {
		uint32_t wl_n = 0;
		if (t < nfull)
			rj_insert_tile<HAS_PRESENT, true>(s, pr, sm, buf, t, par, wl_n);
		else
			rj_insert_tile<HAS_PRESENT, false>(s, pr, sm, buf, t, par, wl_n);
		rj_reinsert_parked(pr, sm, par, wl_n);
		rj_round_end(s, pr, sm, par, wl_n);
		par ^= 1;
	};
	while (tile < ntiles) {
		uint64_t next = tile + gridDim.x;
		if (next < ntiles)
			rj_load_tile(s, next, buf_b);
		round(buf_a, tile);
		tile = next;
		if (tile >= ntiles)
			break;
		next = tile + gridDim.x;
		if (next < ntiles)
			rj_load_tile(s, next, buf_a);
		round(buf_b, tile);
		tile = next;
	}
	rj_finish(s, pr, sm, par);
}
