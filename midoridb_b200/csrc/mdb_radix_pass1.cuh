// mdb_radix_pass1.cuh - pass 1 of the radix join: stream the keys once, scatter 2-byte remainders into
// per-partition 512-byte chunks.  Included by mdb_radix.cu (structs RJSide/RJParams/RJTarget are defined there).
//
// One persistent 1024-thread CTA per SM.  Per round every thread takes 8 keys:
//   insert   partition = (key - kmin) >> shift; ONE shared-memory atomic hands out a slot of the partition's
//            40-byte staging row, ONE 2-byte shared store writes the remainder.  The lane that fills slot 16 puts
//            the partition on ITS WARP's worklist (position from a ballot: no atomics, no CTA-wide queue);
//   barrier  (all remainders of the queued rows are in shared memory)
//   flush    every warp flushes its own worklist, one lane per row: the first 16 remainders leave as ONE
//            256-bit store = one aligned 32-byte sector of the CTA's current chunk of that partition; fresh chunk
//            ids are handed out warp-wide (one shared-memory atomic per warp, ids reserved in bulk by thread 0);
//   barrier
// What the measurements behind this shape say (profiles/microbench/p1_lab*.cu, B200, 2^28 keys): 256-bit key loads
// stream at 6.4 TB/s where 128-bit loads reach 4.7; slot atomics + stores cost 0.05 ms on top of the loads;
// a CTA-wide worklist fed by same-address atomics costs 0.15 ms more than ballots; scattered sector stores 0.2 ms.

#ifndef RJ_LAB
#define RJ_LAB 0                   // profiles/microbench/p1_lab3.cu sets bits to knock out parts of the kernel (timing only)
#endif

#define RJ_P1_WARPS (RJ_P1_THREADS / 32)
#define RJ_P1_KEYS 8               // keys per thread per round
#define RJ_WL_CAP 64               // rows one warp may complete per round (16 expected; more = skew -> general operators)
#define RJ_TAIL_ROUNDS 64          // rounds spent on parked keys after the last tile before giving up (skew)

#define RJ_HINT_PREFETCH 1u        // prefetch.global.L2 two tiles ahead
#define RJ_HINT_LOAD_EVICT_FIRST 2u
#define RJ_HINT_STORE_EVICT_LAST 4u
#define RJ_HINT_DEFAULT 0u         // (MDBCU_P1_HINTS overrides; see profiles/ for the sweep)

struct RJP1Smem {
	uint16_t stage[RJ_MAX_PART * RJ_CAP];     // 160 KiB: 20 two-byte slots per partition
	uint32_t fill[RJ_MAX_PART / 2];           // slots handed out since the last flush (may overshoot RJ_CAP): 16 bits per partition
	uint32_t chunk[RJ_MAX_PART];              // current chunk of this CTA: chunk id * 32 + sectors used, or RJ_NONE
	uint16_t worklist[RJ_P1_WARPS][RJ_WL_CAP]; // per warp: partitions whose 16th slot it filled this round
	uint32_t ovf[2][RJ_OVF_CAP];              // (partition << 16 | remainder) waiting for the next round
	uint32_t ovf_count[2];
	uint32_t local_next[RJ_MAX_RANKS];        // chunk ids reserved by this CTA in each owner's pool
	uint32_t local_end[RJ_MAX_RANKS];         // (refilled in bulk by thread 0)
};

// Measured (profiles/microbench/p1_lab4.cu, PAD sweep): a CTA that needs more than 195 KiB pushes the SM into its
// largest shared-memory carve-out, the L1 that remains cannot hold the key loads in flight and the kernel loses 25%.
static_assert(sizeof(RJP1Smem) <= 195 * 1024, "pass-1 shared memory must stay inside the 196 KiB carve-out");

// slot counters: two partitions share a 32-bit word
__device__ __forceinline__ uint32_t rj_fill_get(const RJP1Smem *sm, uint32_t p)
{
	return (sm->fill[p >> 1] >> ((p & 1u) << 4)) & 0xffffu;
}

// plain shared-memory atomic: kept in PTX so the compiler does not expand it into warp-aggregation code
__device__ __forceinline__ uint32_t rj_smem_add(uint32_t *p, uint32_t v)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
	return old;
}

__device__ __forceinline__ uint32_t rj_smem_inc(uint32_t *p)
{
	return rj_smem_add(p, 1u);
}

// hand out the next slot of partition p's staging row
__device__ __forceinline__ uint32_t rj_fill_claim(RJP1Smem *sm, uint32_t p)
{
	const uint32_t sh = (p & 1u) << 4;
	return (rj_smem_add(&sm->fill[p >> 1], 1u << sh) >> sh) & 0xffffu;
}

// after a flush: the count drops from `have` to `keep` (atomic: the neighbour's half of the word may change concurrently)
__device__ __forceinline__ void rj_fill_drop(RJP1Smem *sm, uint32_t p, uint32_t have, uint32_t keep)
{
	rj_smem_add(&sm->fill[p >> 1], (keep - have) << ((p & 1u) << 4));
}

__device__ __forceinline__ void rj_global_red_inc(uint32_t *p)
{
	asm volatile("red.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
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

// rank that owns partition p: ranks own the contiguous blocks [r*P/W, (r+1)*P/W)
__device__ __forceinline__ int rj_owner(const RJSide &s, const RJParams &pr, uint32_t p)
{
	return s.world == 1 ? 0 : (int)(((p + 1) * (uint32_t)s.world - 1) / (uint32_t)pr.nparts);
}

// book-keeping of a freshly allocated chunk `cid` that replaces `old` (chunk id * 32 + 16, or RJ_NONE) for partition p
__device__ __forceinline__ bool rj_open_chunk(const RJSide &s, const RJParams &pr, const RJTarget &t, uint32_t p, uint32_t old, uint32_t cid)
{
	if (cid >= s.pool_chunks) {
		atomicOr(pr.error_flag, RJ_ERR_POOL);
		return false;
	}
	if (RJ_LAB & 4)
		return true;
	if (old != RJ_NONE)
		t.chunk_entries[old >> 5] = RJ_CHUNK; // a chunk is only replaced when all its sectors are written
	t.chunk_part[cid] = (uint16_t)p;
	rj_global_red_inc(&t.dir_cnt[p]);
	return true;
}

// single-thread allocation (drain): ids come from the CTA's reserved range in the owner's pool
__device__ static inline void rj_new_chunk(const RJSide &s, const RJParams &pr, RJP1Smem *sm, uint32_t p)
{
	const int o = rj_owner(s, pr, p);
	const RJTarget &t = s.dst[o];
	uint32_t cid = rj_smem_inc(&sm->local_next[o]);
	if (cid >= sm->local_end[o])
		cid = atomicAdd(t.pool_next, 1u);
	if (rj_open_chunk(s, pr, t, p, sm->chunk[p], cid))
		sm->chunk[p] = cid << 5;
}

// thread 0 tops the CTA's id ranges up with ONE global (for a remote owner: NVLink) atomic per id_batch chunks
// while the other threads insert keys, so that no flush waits on L2 or on the link
__device__ static inline void rj_refill_ids(const RJSide &s, RJP1Smem *sm)
{
	for (int o = 0; o < s.world; o++) {
		if (sm->local_end[o] - min(sm->local_next[o], sm->local_end[o]) < s.id_low) {
			const uint32_t base = atomicAdd(s.dst[o].pool_next, s.id_batch);
			sm->local_next[o] = base;
			sm->local_end[o] = base + s.id_batch; // ids left in the old range stay unused (chunk_part 0xffff)
		}
	}
}

__device__ static inline void rj_park(RJP1Smem *sm, const RJParams &pr, uint32_t item, int par)
{
	// staging row full until this round's flush: the key waits one round
	if (RJ_LAB & 1)
		return;
	const uint32_t o = rj_smem_inc(&sm->ovf_count[par]);
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

// start of a round: id ranges are topped up, keys parked by the previous round go first (their rows were flushed since)
__device__ static inline void rj_round_begin(const RJSide &s, const RJParams &pr, RJP1Smem *sm, int par, uint32_t &wl_n)
{
	const uint32_t tid = threadIdx.x;
	wl_n = 0;
	if (tid == 0)
		rj_refill_ids(s, sm);
	const uint32_t novf = (RJ_LAB & 1) ? 0u : min(sm->ovf_count[par ^ 1], (uint32_t)RJ_OVF_CAP);
	for (uint32_t i0 = tid & ~31u; i0 < novf; i0 += RJ_P1_THREADS) { // warp-uniform trip count
		const uint32_t i = i0 + (tid & 31u);
		const bool valid = i < novf;
		rj_insert_ws(sm, pr, valid ? sm->ovf[par ^ 1][i] : 0u, valid, par, wl_n);
	}
}

// 8 keys per thread whose partition/remainder are already packed as (partition << 16 | remainder); RJ_NONE = no key.
// All slot requests of a thread are issued back to back (independent shared-memory atomics), then consumed.
template <bool ALL_VALID>
__device__ __forceinline__ void rj_insert_items(RJP1Smem *sm, const RJParams &pr, const uint32_t *item, int par, uint32_t &wl_n)
{
	uint32_t pos[RJ_P1_KEYS];
#pragma unroll
	for (int k = 0; k < RJ_P1_KEYS; k++)
		pos[k] = (ALL_VALID || item[k] != RJ_NONE) ? rj_fill_claim(sm, item[k] >> 16) : RJ_NONE;
	const uint32_t lt = rj_lanemask_lt();
	uint16_t *wl = sm->worklist[threadIdx.x >> 5];
	uint32_t park_mask = 0;
#pragma unroll
	for (int k = 0; k < RJ_P1_KEYS; k++) {
		const uint32_t p = item[k] >> 16;
		if (pos[k] < RJ_CAP)
			sm->stage[p * RJ_CAP + pos[k]] = (uint16_t)item[k];
		const bool done = pos[k] == RJ_FLUSH - 1;
		const uint32_t bal = __ballot_sync(0xffffffffu, done);
		if (done) {
			const uint32_t idx = wl_n + __popc(bal & lt);
			if (idx < RJ_WL_CAP)
				wl[idx] = (uint16_t)p;
		}
		wl_n += __popc(bal);
		park_mask |= (pos[k] >= RJ_CAP && (ALL_VALID || item[k] != RJ_NONE)) ? (1u << k) : 0u;
	}
	if (park_mask) {
#pragma unroll
		for (int k = 0; k < RJ_P1_KEYS; k++)
			if (park_mask & (1u << k))
				rj_park(sm, pr, item[k], par);
	}
}

// barrier, every warp flushes the rows on its own worklist, barrier
__device__ static inline void rj_round_end(const RJSide &s, const RJParams &pr, RJP1Smem *sm, int par, uint32_t wl_n)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	const uint16_t *wl = sm->worklist[tid >> 5];
	__syncthreads();
	if (tid == 0)
		sm->ovf_count[par ^ 1] = 0; // the other parity's parked keys were re-inserted at the start of this round
	if (wl_n > RJ_WL_CAP) {
		if (lane == 0)
			atomicOr(pr.error_flag, RJ_ERR_SKEW);
		wl_n = RJ_WL_CAP;
	}
	const uint32_t lt = rj_lanemask_lt();
	const bool evict_last = (s.hints & RJ_HINT_STORE_EVICT_LAST) != 0;
	for (uint32_t base = 0; base < wl_n; base += 32) { // warp-uniform
		const bool act = base + lane < wl_n;
		const uint32_t p = act ? wl[base + lane] : 0u;
		uint32_t ch = act ? sm->chunk[p] : 0u;
		if (RJ_LAB & 2) { // no chunk logic: sectors go to a private, hashed position
			if (act) {
				uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * RJ_CAP]);
				const uint32_t have = rj_fill_get(sm, p);
				const uint2 a = row[0], b = row[1], c = row[2], d = row[3], e = row[4];
				const uint32_t sec = (((tid >> 5) << 12) + ((par * 977u + base + lane) & 4095u)) * 40503u & 131071u;
				if (!(RJ_LAB & 8))
					rj_store_sector(s.dst[0].pool + ((size_t)blockIdx.x * 131072u + sec) * RJ_FLUSH, a, b, c, d, evict_last);
				row[0] = e;
				rj_fill_drop(sm, p, have, min(have, (uint32_t)RJ_CAP) - RJ_FLUSH);
			}
			continue;
		}
		const bool need = act && (ch == RJ_NONE || (ch & 31u) == RJ_BLOCKS_PER_CHUNK);
		const int o = rj_owner(s, pr, p);
		bool ok = act;
		if (__any_sync(0xffffffffu, need)) {
			// fresh chunks for the whole warp: one shared-memory atomic per owner
			uint32_t cid = 0;
			for (int oo = 0; oo < s.world; oo++) {
				const uint32_t m = __ballot_sync(0xffffffffu, need && o == oo);
				if (m == 0)
					continue;
				uint32_t first = 0;
				if (lane == 0)
					first = rj_smem_add(&sm->local_next[oo], (uint32_t)__popc(m));
				first = __shfl_sync(0xffffffffu, first, 0);
				if (need && o == oo)
					cid = first + __popc(m & lt);
			}
			if (need) {
				if (cid >= sm->local_end[o])
					cid = atomicAdd(s.dst[o].pool_next, 1u); // the reserve ran dry inside one round (start-up, skew)
				ok = rj_open_chunk(s, pr, s.dst[o], p, ch, cid);
				ch = cid << 5;
			}
		}
		if (ok) {
			uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * RJ_CAP]); // 40-byte rows are 8-byte aligned
			const uint32_t have = rj_fill_get(sm, p);
			const uint2 a = row[0], b = row[1], c = row[2], d = row[3], e = row[4];
			if (!(RJ_LAB & 8))
				rj_store_sector(s.dst[o].pool + (size_t)(ch >> 5) * RJ_CHUNK + (ch & 31u) * RJ_FLUSH, a, b, c, d, evict_last);
			sm->chunk[p] = ch + 1;
			row[0] = e; // keep the (at most 4) remainders behind the flushed sector
			rj_fill_drop(sm, p, have, min(have, (uint32_t)RJ_CAP) - RJ_FLUSH);
		}
	}
	__syncthreads();
}

// every partition's partial sector goes out, chunk entry counts are finalised
__device__ static inline void rj_drain(const RJSide &s, const RJParams &pr, RJP1Smem *sm)
{
	for (int p = threadIdx.x; p < pr.nparts; p += RJ_P1_THREADS) {
		const uint32_t f = min(rj_fill_get(sm, p), (uint32_t)RJ_CAP);
		const RJTarget &t = s.dst[rj_owner(s, pr, p)];
		uint32_t ch = sm->chunk[p];
		if (f > 0) {
			if (ch == RJ_NONE || (ch & 31u) == RJ_BLOCKS_PER_CHUNK) {
				rj_new_chunk(s, pr, sm, p);
				ch = sm->chunk[p];
			}
			if (ch != RJ_NONE && (ch & 31u) < RJ_BLOCKS_PER_CHUNK) {
				uint16_t *dst = t.pool + (size_t)(ch >> 5) * RJ_CHUNK + (ch & 31u) * RJ_FLUSH;
				for (uint32_t i = 0; i < f; i++)
					dst[i] = sm->stage[p * RJ_CAP + i];
				t.chunk_entries[ch >> 5] = (uint16_t)((ch & 31u) * RJ_FLUSH + f);
			}
		} else if (ch != RJ_NONE) {
			t.chunk_entries[ch >> 5] = (uint16_t)((ch & 31u) * RJ_FLUSH);
		}
	}
	__threadfence_system(); // remote owners read these chunks after the next cross-rank barrier
}

__device__ static inline void rj_smem_init(RJP1Smem *sm)
{
	const int tid = threadIdx.x;
	for (int p = tid; p < RJ_MAX_PART; p += RJ_P1_THREADS) {
		if (p < RJ_MAX_PART / 2)
			sm->fill[p] = 0;
		sm->chunk[p] = RJ_NONE;
	}
	if (tid < 2)
		sm->ovf_count[tid] = 0;
	if (tid < RJ_MAX_RANKS)
		sm->local_next[tid] = sm->local_end[tid] = 0;
	__syncthreads();
}

// rounds for the keys still parked after the last tile (bounded: a row that never drains means extreme skew)
__device__ static inline void rj_finish(const RJSide &s, const RJParams &pr, RJP1Smem *sm, int par)
{
	uint32_t wl_n;
	for (int guard = 0; sm->ovf_count[par ^ 1] != 0; guard++) { // block-uniform: written before the last barrier
		if (guard == RJ_TAIL_ROUNDS) {
			if (threadIdx.x == 0)
				atomicOr(pr.error_flag, RJ_ERR_SKEW);
			break;
		}
		rj_round_begin(s, pr, sm, par, wl_n);
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
	auto round = [&](const uint32_t *buf) {
		uint32_t wl_n, item[RJ_P1_KEYS];
		rj_round_begin(s, pr, sm, par, wl_n);
#pragma unroll
		for (int k = 0; k < RJ_P1_KEYS; k++) {
			const uint32_t d = buf[k] - kmin_lo;
			item[k] = ((d >> pr.shift) << 16) | (d & pr.mask);
		}
		rj_insert_items<true>(sm, pr, item, par, wl_n);
		rj_round_end(s, pr, sm, par, wl_n);
		par ^= 1;
	};
	uint64_t tile = blockIdx.x;
	if (tile < nfull)
		load(tile, buf_a);
	while (tile < nfull) {
		uint64_t next = tile + gridDim.x;
		if (next < nfull)
			load(next, buf_b);
		round(buf_a);
		tile = next;
		if (tile >= nfull)
			break;
		next = tile + gridDim.x;
		if (next < nfull)
			load(next, buf_a);
		round(buf_b);
		tile = next;
	}
	if (blockIdx.x == 0 && nfull * TILE != s.n) {
		uint32_t wl_n;
		rj_round_begin(s, pr, sm, par, wl_n);
		for (uint64_t r0 = nfull * TILE + (tid & ~31u); r0 < s.n; r0 += RJ_P1_THREADS) { // warp-uniform trip count
			const uint64_t r = r0 + (tid & 31u);
			const bool valid = r < s.n;
			const uint32_t d = valid ? (uint32_t)(unsigned long long)s.keys[r] - kmin_lo : 0u;
			rj_insert_ws(sm, pr, ((d >> pr.shift) << 16) | (d & pr.mask), valid, par, wl_n);
		}
		rj_round_end(s, pr, sm, par, wl_n);
		par ^= 1;
	}
	rj_finish(s, pr, sm, par);
}

// generic tile load: 128-bit loads of whole 64-bit keys (range test needs the high words)
template <bool HAS_PRESENT>
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
	rj_insert_items<false>(sm, pr, item, par, wl_n);
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
		rj_load_tile<HAS_PRESENT>(s, tile, buf_a);
	auto round = [&](const int4 *buf, uint64_t t) {
		uint32_t wl_n;
		rj_round_begin(s, pr, sm, par, wl_n);
		if (t < nfull)
			rj_insert_tile<HAS_PRESENT, true>(s, pr, sm, buf, t, par, wl_n);
		else
			rj_insert_tile<HAS_PRESENT, false>(s, pr, sm, buf, t, par, wl_n);
		rj_round_end(s, pr, sm, par, wl_n);
		par ^= 1;
	};
	while (tile < ntiles) {
		uint64_t next = tile + gridDim.x;
		if (next < ntiles)
			rj_load_tile<HAS_PRESENT>(s, next, buf_b);
		round(buf_a, tile);
		tile = next;
		if (tile >= ntiles)
			break;
		next = tile + gridDim.x;
		if (next < ntiles)
			rj_load_tile<HAS_PRESENT>(s, next, buf_a);
		round(buf_b, tile);
		tile = next;
	}
	rj_finish(s, pr, sm, par);
}
