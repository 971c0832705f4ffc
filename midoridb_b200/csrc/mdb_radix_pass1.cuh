// mdb_radix_pass1.cuh - pass 1 of the radix join: stream the keys once, append their 2-byte remainders to the
// per-partition streams (mdb_radix_types.cuh).  Included by mdb_radix.cu.
//
// One persistent 1024-thread CTA per SM.  Per round every thread takes 8 keys:
//   insert   partition = (key - kmin) / width; ONE shared-memory atomic hands out a slot of the partition's
//            40-byte staging row, ONE 2-byte shared store writes the remainder - nothing else per key.  A key that
//            finds its row full (about 7 of 8192 per round) stays in its thread's register until the flush;
//   barrier  (all remainders handed out this round are in shared memory)
//   flush    every warp owns 128 partitions: one 128-bit read of four slot counters per lane finds the rows that hold
//            a whole sector, ballots compact them so that every lane gets one row; one global atomic on the
//            partition's cursor gives the position, the first 16 remainders leave as ONE 256-bit store = one aligned
//            32-byte sector; spilled keys become tail sectors of their own in the lanes the flush leaves idle;
//   barrier
// Measurements behind this shape (profiles/microbench/p1_lab*.cu and profiles/p1_lab_*.txt, B200, 2^28 keys):
//   * 256-bit key loads stream at 6.4 TB/s where 128-bit loads reach 4.7;
//   * a CTA that needs more than 195 KiB of shared memory pushes the SM into its largest carve-out, the L1 that
//     is left cannot hold the key loads in flight and the whole kernel loses 25%;
//   * slot atomics + 2-byte stores cost 0.05 ms on top of the loads; a CTA-wide worklist fed by same-address
//     atomics costs 0.15 ms more than ballots; stores that complete whole 128-byte lines are 0.13 ms cheaper than
//     scattered 32-byte sectors (hence shared per-partition streams instead of per-CTA chunks);
//   * the kernel is bound by instruction issue and shared-memory latency, not by DRAM: noticing "row complete" per
//     key (compare, ballot, rank, queue) was a third of all executed instructions - hence the scan after the barrier;
//   * re-inserting overflowing keys in the next round cost 0.07 ms (the re-insert sits on one warp's critical path
//     before the barrier) and appending them to the tail stream right away cost 0.08 ms (a global atomic's latency
//     in the insert phase): they wait on a list until the flush, whose atomics are in flight anyway;
//   * two tiles per flush halve the barriers but overflow 25x more often: slower.

#ifndef RJ_LAB
#define RJ_LAB 0                   // profiles/microbench/p1_lab3.cu knocks parts of the kernel out (timing experiments only):
#endif                             // 1 drop overflowing keys, 2 no cursor atomic (fake positions), 4 no sector store

#if RJ_LAB & 8
// timeline of one warp of CTA 0 (p1_lab3): cycles spent in [insert | wait barrier 1 | flush | wait barrier 2], summed over rounds
__device__ unsigned long long rj_timeline[4][8];
#define RJ_STAMP(i) do { if ((RJ_LAB & 8) && blockIdx.x == 0 && (threadIdx.x & 31u) == 0 && (threadIdx.x >> 5) % (RJ_P1_THREADS / 96) == 0) { \
	const long long now_ = clock64(); atomicAdd(&rj_timeline[(threadIdx.x >> 5) / (RJ_P1_THREADS / 96)][i], (unsigned long long)(now_ - rj_t_)); rj_t_ = now_; } } while (0)
#else
#define RJ_STAMP(i) do { } while (0)
#endif

#define RJ_P1_WARPS (RJ_P1_THREADS / 32)
#define RJ_P1_KEYS 8               // keys per thread per round
#define RJ_WARP_PARTS (RJ_ROWS / RJ_P1_WARPS) // staging rows one warp flushes

#ifndef RJ_LOAD_AT
#define RJ_LOAD_AT 1               // where in the round the next tile's keys are requested (k_radix_partition_fast)
#endif
#ifndef RJ_PF_DIST
#define RJ_PF_DIST 2               // tiles between a CTA's load and its L2 prefetch
#endif
#define RJ_HINT_PREFETCH 1u        // prefetch.global.L2 RJ_PF_DIST tiles ahead
#define RJ_HINT_LOAD_EVICT_FIRST 2u
#define RJ_HINT_STORE_EVICT_LAST 4u
#define RJ_HINT_DEFAULT 0u         // (MDBCU_P1_HINTS overrides; none of them pays once the L1 is large enough)

struct RJP1Smem {
	uint16_t stage[RJ_ROWS * RJ_CAP];         // 20 two-byte slots per staged partition (row r = partition r * RJ_SPLIT + this CTA's residue)
	uint32_t fill[RJ_ROWS];                   // slots handed out since the last flush (may overshoot RJ_CAP)
	uint16_t worklist[RJ_P1_WARPS][RJ_WARP_PARTS]; // per warp: those of its partitions that hold a whole sector this round
};

static_assert((sizeof(RJP1Smem) + 1024) * RJ_SPLIT <= 196 * 1024, "pass-1 shared memory of all CTAs of an SM must stay inside the 196 KiB carve-out");
static_assert(RJ_WARP_PARTS == 128, "the flush scan reads four slot counters per lane");

// plain shared-memory atomic: kept in PTX so the compiler does not expand it into warp-aggregation code
__device__ __forceinline__ uint32_t rj_smem_add(uint32_t *p, uint32_t v)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
	return old;
}

// hand out the next slot of staging row r
__device__ __forceinline__ uint32_t rj_fill_claim(RJP1Smem *sm, uint32_t r)
{
	return rj_smem_add(&sm->fill[r], 1u);
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

// d = key - kmin (< range <= 2^28) -> (partition << 16 | remainder).  q = __umulhi(d, floor(2^32 / width)) is d / width or
// one less (d * magic / 2^32 > d / width - 1 because d < 2^32); one compare fixes it.  Exact for every width, and for a
// power of two the correction never fires.  The remainder is the low half: STS.U16 stores it without masking.
__device__ __forceinline__ uint32_t rj_pack(const RJParams &pr, uint32_t d)
{
	uint32_t q = __umulhi(d, pr.magic);
	q += (d - q * pr.width >= pr.width) ? 1u : 0u;
	return d + q * (65536u - pr.width); // = (q << 16) + (d - q * width)
}

// a single remainder as a tail sector of its own: [remainder, 14 unused, count = 1] (mdb_radix_types.cuh)
// (out of line: cold, and the unrolled insert loop that calls it exists twice per kernel)
__device__ __forceinline__ void rj_tail_single(uint32_t *tail_cursor, uint16_t *tail, uint32_t tail_cap, uint32_t *error_flag, uint32_t p, uint32_t rem)
{
	const uint32_t at = atomicAdd(&tail_cursor[p * RJ_CUR_STRIDE], (uint32_t)RJ_FLUSH);
	if (at + RJ_FLUSH <= tail_cap)
		rj_store_sector(tail + (size_t)p * tail_cap + at, make_uint2(rem, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u),
				make_uint2(0u, 1u << 16), false);
	else
		atomicOr(error_flag, RJ_ERR_STREAM);
}

// insert of one item = (partition << 16 | remainder) if its partition is staged by this CTA: slot, store.
// (Ragged end of the column, CTA 0 only, lanes may be missing: a key that finds its row full goes to the tail directly.)
__device__ __forceinline__ void rj_insert_one(const RJSide &s, RJP1Smem *sm, const RJParams &pr, uint32_t item, uint32_t half)
{
	const uint32_t p = item >> 16;
	if (RJ_SPLIT > 1 && p % RJ_SPLIT != half)
		return;
	const uint32_t r = p / RJ_SPLIT;
	const uint32_t pos = rj_fill_claim(sm, r);
	if (pos < RJ_CAP)
		sm->stage[r * RJ_CAP + pos] = (uint16_t)item;
	else if (!(RJ_LAB & 1))
		rj_tail_single(s.tail_cursor, s.tail, s.tail_cap, pr.error_flag, p, item & 0xffffu);
}

// 8 keys per thread, item = (partition << 16 | remainder) (rj_pack); ALL_VALID: every item is a key (lean kernel),
// otherwise RJ_NONE = no key (generic kernel).
// All slot requests of a thread are issued back to back (independent shared-memory atomics), then consumed.
// Returns RJ_NONE, or (partition << 16 | remainder) of a key of THIS THREAD that found its row full (7 keys in 8192 do):
// it stays in a register until the flush, which turns it into a tail sector.  Nothing warp-wide and no shared memory
// here: a CTA-wide list behind a shared counter, per-warp lists filled by ballots - every cooperative scheme made the
// warp that took it several hundred cycles late at the barrier, round after round (0.07-0.1 ms per launch,
// profiles/r02_lab3_*.txt).  A thread's second overflow in one round (about once per launch) goes to the tail directly.
// kept_at: the key's position in its partition's tail stream - the global atomic is ISSUED here and its result is first
// looked at after the flush, a thousand cycles later, so nobody ever waits for it.
template <bool ALL_VALID>
__device__ __forceinline__ uint32_t rj_insert_items(const RJSide &s, RJP1Smem *sm, const RJParams &pr, const uint32_t *item, uint32_t half,
		uint32_t &kept_at)
{
	constexpr bool ALL_MINE = ALL_VALID && RJ_SPLIT == 1; // every item is a key of a partition this CTA stages
	uint32_t pos[RJ_P1_KEYS];
#pragma unroll
	for (int k = 0; k < RJ_P1_KEYS; k++) {
		const uint32_t p = item[k] >> 16;
		const bool mine = (ALL_VALID || item[k] != RJ_NONE) && (RJ_SPLIT == 1 || p % RJ_SPLIT == half);
		pos[k] = mine ? rj_fill_claim(sm, p / RJ_SPLIT) : RJ_NONE;
	}
	uint32_t worst = 0; // largest slot handed to this thread (+1 unless ALL_MINE, so that RJ_NONE counts as 0)
#pragma unroll
	for (int k = 0; k < RJ_P1_KEYS; k++) {
		const uint32_t r = (item[k] >> 16) / RJ_SPLIT;
		if (pos[k] < RJ_CAP)
			sm->stage[r * RJ_CAP + pos[k]] = (uint16_t)item[k];
		worst = max(worst, ALL_MINE ? pos[k] : pos[k] + 1u);
	}
	// rare (7 keys in 8192): some row was full
	uint32_t kept = RJ_NONE;
	if (__any_sync(0xffffffffu, worst >= (ALL_MINE ? RJ_CAP : RJ_CAP + 1u)) && !(RJ_LAB & 1)) {
		// the whole warp, branch-free: the thread's LAST key that found its row full is kept
		uint32_t n_over = 0;
#pragma unroll
		for (int k = 0; k < RJ_P1_KEYS; k++) {
			const bool over = pos[k] >= RJ_CAP && pos[k] != RJ_NONE;
			kept = over ? item[k] : kept;
			n_over += over ? 1u : 0u;
		}
		if (kept != RJ_NONE)
			kept_at = atomicAdd(&s.tail_cursor[(kept >> 16) * RJ_CUR_STRIDE], (uint32_t)RJ_FLUSH);
		if (n_over > 4) {
			// clustered keys (a sorted column: all keys of a tile fall into one or two rows): this launch's result is
			// abandoned, do not write millions of single-key tail sectors on the way out
			atomicOr(pr.error_flag, RJ_ERR_SKEW);
		} else if (n_over > 1) { // the others (about 700 keys per launch of 2^28) go to the tail right away
			bool last = true;
#pragma unroll
			for (int k = RJ_P1_KEYS - 1; k >= 0; k--) {
				if (pos[k] < RJ_CAP || pos[k] == RJ_NONE)
					continue;
				if (!last)
					rj_tail_single(s.tail_cursor, s.tail, s.tail_cap, pr.error_flag, item[k] >> 16, item[k] & 0xffffu);
				last = false;
			}
		}
	}
	return kept;
}

// barrier, every warp flushes the full rows among ITS 128 partitions, kept keys go to the tail streams, barrier.
// (The 7 keys in 8192 that find their row full once cost 0.15 ms of a 0.85 ms launch: handled ahead of the flush by the
// first threads of the CTA they put one more global-atomic latency on warp 0's path to the second barrier, and every
// cooperative way of listing them made one warp late at the first: profiles/r02_lab3_*.txt.)
// kept: this thread's key that found its row full, or RJ_NONE; kept_at: its tail position (rj_insert_items)
struct RJNoHook {
	__device__ __forceinline__ void operator()() const {}
};

// after_scan / after_flush: called by every thread once the warp's worklist is built / once its sectors are stored (the lean
// kernel requests the next tile's keys at one of these points)
template <class AfterScan = RJNoHook, class AfterFlush = RJNoHook>
__device__ __forceinline__ void rj_round_end(const RJSide &s, const RJParams &pr, RJP1Smem *sm, uint32_t kept, uint32_t kept_at, uint32_t half,
		long long &rj_t_, AfterScan after_scan = AfterScan(), AfterFlush after_flush = AfterFlush())
{
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	uint16_t *wl = sm->worklist[warp];
	RJ_STAMP(0);
	__syncthreads();
	RJ_STAMP(1);
	// which of this warp's rows hold a whole sector?  (the insert loop itself never looks at "row complete")
	uint32_t wl_n = 0;
	{
		const uint32_t first = warp * RJ_WARP_PARTS + lane * 4u;
		const uint4 f = *reinterpret_cast<const uint4*>(&sm->fill[first]);
		const uint32_t cnt[4] = {f.x, f.y, f.z, f.w};
		const uint32_t lt = rj_lanemask_lt();
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const bool full = cnt[j] >= RJ_FLUSH;
			const uint32_t bal = __ballot_sync(0xffffffffu, full);
			if (full)
				wl[wl_n + __popc(bal & lt)] = (uint16_t)(first + j);
			wl_n += __popc(bal);
		}
		__syncwarp();
	}
	RJ_STAMP(4);
	after_scan();
	const bool evict_last = (s.hints & RJ_HINT_STORE_EVICT_LAST) != 0;
	for (uint32_t w = lane; w < wl_n; w += 32) {
		const uint32_t r = wl[w], p = r * RJ_SPLIT + half;
		const uint32_t at = (RJ_LAB & 2) ? ((warp * 997u + w * 16u) & 0xfff0u) : atomicAdd(&s.cursor[p * RJ_CUR_STRIDE], (uint32_t)RJ_FLUSH);
		uint2 *row = reinterpret_cast<uint2*>(&sm->stage[r * RJ_CAP]); // 40-byte rows are 8-byte aligned
		const uint32_t have = sm->fill[r];
		const uint2 a = row[0], b = row[1], c = row[2], d = row[3], e = row[4];
		row[0] = e; // keep the (at most 4) remainders behind the flushed sector
		sm->fill[r] = min(have, (uint32_t)RJ_CAP) - RJ_FLUSH;
		if (RJ_LAB & 4)
			continue;
		if (at + RJ_FLUSH <= s.cap)
			rj_store_sector(s.stream + (size_t)p * s.cap + ((RJ_LAB & 32) ? (at & 0x3f0u) : at), a, b, c, d, evict_last);
		else
			atomicOr(pr.error_flag, RJ_ERR_STREAM);
	}
	// a kept key becomes a tail sector of its own, [remainder, 14 unused, count = 1]; its position was requested during
	// the insert phase
	if (kept != RJ_NONE && !(RJ_LAB & 256)) {
		if (kept_at + RJ_FLUSH <= s.tail_cap)
			rj_store_sector(s.tail + (size_t)(kept >> 16) * s.tail_cap + kept_at, make_uint2(kept & 0xffffu, 0u), make_uint2(0u, 0u),
					make_uint2(0u, 0u), make_uint2(0u, 1u << 16), false);
		else
			atomicOr(pr.error_flag, RJ_ERR_STREAM);
	}
	after_flush();
	RJ_STAMP(2);
	__syncthreads();
	RJ_STAMP(3);
}

// every staged partition's partial sector (at most 15 remainders after the last flush) becomes ONE tail sector:
// [remainders..., count in entry 15] - one global atomic and one 32-byte store per row
__device__ static inline void rj_drain(const RJSide &s, const RJParams &pr, RJP1Smem *sm, uint32_t half)
{
	if (RJ_LAB & 16)
		return;
	for (uint32_t r = threadIdx.x; r < RJ_ROWS; r += RJ_P1_THREADS) {
		const uint32_t p = r * RJ_SPLIT + half;
		const uint32_t f = min(sm->fill[r], (uint32_t)RJ_CAP);
		if (f == 0 || p >= (uint32_t)pr.nparts)
			continue;
		if (f >= RJ_FLUSH) { // cannot happen: the last round's flush leaves fewer than 16 behind
			atomicOr(pr.error_flag, RJ_ERR_STREAM);
			continue;
		}
		const uint32_t at = atomicAdd(&s.tail_cursor[p * RJ_CUR_STRIDE], (uint32_t)RJ_FLUSH);
		if (at + RJ_FLUSH <= s.tail_cap) {
			const uint2 *row = reinterpret_cast<const uint2*>(&sm->stage[r * RJ_CAP]);
			uint2 d = row[3];
			d.y = (d.y & 0xffffu) | (f << 16);
			rj_store_sector(s.tail + (size_t)p * s.tail_cap + at, row[0], row[1], row[2], d, false);
		} else {
			atomicOr(pr.error_flag, RJ_ERR_STREAM);
		}
	}
}

__device__ static inline void rj_smem_init(RJP1Smem *sm)
{
	for (int r = threadIdx.x; r < RJ_ROWS; r += RJ_P1_THREADS)
		sm->fill[r] = 0;
	__syncthreads();
}

// Pass 1, lean variant: column without NULLs/tombstones whose [min, max] lies inside the partitioned range and
// key - kmin < 2^32: no per-key validity test, 32-bit arithmetic on the low words, 256-bit key loads.
// The ragged tail (< one tile) goes through CTA 0.
//
// WHEN the next tile is requested decides a seventh of this kernel's time (profiles/r02_lab3_r.txt, r02_lab3_s.txt).  A tile
// is 64 KiB per SM; the L1 accepts such a burst over about a thousand cycles, and a warp's shared-memory instructions
// queue behind its own pending loads.  Requested at the top of the round (double-buffered, the obvious software
// pipeline) the loads sit in front of every warp's slot atomics: insert phase 3300 cycles.  Requested by each warp when
// ITS keys have been handed to shared memory - into the same registers, one buffer - they are accepted while the warp
// waits at the barrier and arrive during the flush: insert phase 2100 cycles, 0.756 -> 0.649 ms per launch.  Behind the
// barrier (after the counter scan) they delay the flush instead: 0.727 ms.  RJ_LOAD_AT selects the point (lab only).
// W16: the partitions are exactly 2^16 key values wide (key ranges of 2^28 - 65535 .. 2^28, the headline workload's):
// key - kmin IS (partition << 16 | remainder), the seven instructions of rj_pack per key disappear.
template <bool W16>
__global__ void __launch_bounds__(RJ_P1_THREADS, RJ_SPLIT) k_radix_partition_fast(RJSide s, RJParams pr)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RJP1Smem *sm = reinterpret_cast<RJP1Smem*>(smem_raw);
	rj_smem_init(sm);

	constexpr uint32_t TILE = RJ_P1_THREADS * RJ_P1_KEYS;
	const uint32_t tid = threadIdx.x;
	const uint32_t half = blockIdx.x % RJ_SPLIT, slot = blockIdx.x / RJ_SPLIT, nslots = gridDim.x / RJ_SPLIT; // CTA group = one SM's worth
	const uint64_t nfull = s.n / TILE;
	const uint32_t kmin_lo = (uint32_t)(unsigned long long)pr.kmin;
	const bool pf = (s.hints & RJ_HINT_PREFETCH) != 0, evict_first = (s.hints & RJ_HINT_LOAD_EVICT_FIRST) != 0;
	uint32_t buf[RJ_P1_KEYS];
	long long rj_t_ = (RJ_LAB & 8) ? clock64() : 0; // (timeline experiments only)
	uint64_t tile = slot;
	const char *src = reinterpret_cast<const char*>(s.keys + tile * TILE) + tid * 32u;
	if (tile < nfull) {
		rj_load_keys256(src, buf, evict_first);
		rj_load_keys256(src + RJ_P1_THREADS * 32u, buf + 4, evict_first);
	}
	while (tile < nfull) {
		uint32_t item[RJ_P1_KEYS];
#pragma unroll
		for (int k = 0; k < RJ_P1_KEYS; k++)
			item[k] = W16 ? buf[k] - kmin_lo : rj_pack(pr, buf[k] - kmin_lo);
		uint32_t kept_at = 0;
		const uint32_t kept = rj_insert_items<true>(s, sm, pr, item, half, kept_at);
		tile += nslots;
		src += (size_t)nslots * TILE * sizeof(int64_t);
		const bool more = tile < nfull;
		if (pf) {
			// pull the tile this CTA will load RJ_PF_DIST rounds from now into L2 (one 128-byte line per thread)
			const uint64_t pf_first = (tile + (uint64_t)RJ_PF_DIST * nslots) * TILE + (uint64_t)tid * 16;
			if (tid < TILE / 16 && pf_first + 16 <= s.n)
				asm volatile("prefetch.global.L2 [%0];" ::"l"(s.keys + pf_first));
		}
		auto load_lo = [&]() {
			if (more)
				rj_load_keys256(src, buf, evict_first);
		};
		auto load_hi = [&]() {
			if (more)
				rj_load_keys256(src + RJ_P1_THREADS * 32u, buf + 4, evict_first);
		};
		if (RJ_LOAD_AT == 1) { // both 256-bit loads when this warp's keys are in shared memory, ahead of the first barrier
			load_lo();
			load_hi();
			rj_round_end(s, pr, sm, kept, kept_at, half, rj_t_);
		} else if (RJ_LOAD_AT == 2) { // second load after the flush, ahead of the second barrier
			load_lo();
			rj_round_end(s, pr, sm, kept, kept_at, half, rj_t_, RJNoHook(), load_hi);
		} else if (RJ_LOAD_AT == 3) { // both behind the first barrier, after the scan of the slot counters
			rj_round_end(s, pr, sm, kept, kept_at, half, rj_t_, [&]() { load_lo(); load_hi(); });
		} else { // second load after the scan
			load_lo();
			rj_round_end(s, pr, sm, kept, kept_at, half, rj_t_, load_hi);
		}
	}
	if (slot == 0 && nfull * TILE != s.n) {
		for (uint64_t r = nfull * TILE + tid; r < s.n; r += RJ_P1_THREADS) {
			const uint32_t d = (uint32_t)(unsigned long long)s.keys[r] - kmin_lo;
			rj_insert_one(s, sm, pr, rj_pack(pr, d), half);
		}
		rj_round_end(s, pr, sm, RJ_NONE, 0u, half, rj_t_);
	}
	rj_drain(s, pr, sm, half);
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

// generic decode: range test, NULL/tombstone bitmap, ragged last tile -> item = (partition << 16 | remainder) or RJ_NONE.
// FULL: every row of the tile exists
template <bool HAS_PRESENT, bool FULL>
__device__ static inline void rj_decode_tile(const RJSide &s, const RJParams &pr, const int4 *buf, uint64_t tile, uint32_t *item)
{
	constexpr uint32_t TILE = RJ_P1_THREADS * RJ_P1_KEYS;
	const uint64_t base_pair = tile * (TILE / 2);
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
		item[2 * j] = ok0 ? rj_pack(pr, (uint32_t)d0) : RJ_NONE;
		item[2 * j + 1] = ok1 ? rj_pack(pr, (uint32_t)d1) : RJ_NONE;
	}
}

// Pass 1, generic variant (NULLs / tombstones / keys outside the partitioned range): same rounds as the lean kernel,
// whole keys; the next tile is requested into the same registers once this tile's keys are in shared memory.
template <bool HAS_PRESENT>
__global__ void __launch_bounds__(RJ_P1_THREADS, RJ_SPLIT) k_radix_partition(RJSide s, RJParams pr)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RJP1Smem *sm = reinterpret_cast<RJP1Smem*>(smem_raw);
	rj_smem_init(sm);

	constexpr uint32_t TILE = RJ_P1_THREADS * RJ_P1_KEYS;
	const uint32_t half = blockIdx.x % RJ_SPLIT, slot = blockIdx.x / RJ_SPLIT, nslots = gridDim.x / RJ_SPLIT;
	const uint64_t ntiles = (s.n + TILE - 1) / TILE;
	const uint64_t nfull = s.n / TILE; // tiles [0, nfull) are complete
	int4 buf[RJ_P1_KEYS / 2];
	uint64_t tile = slot;
	long long rj_t_ = (RJ_LAB & 8) ? clock64() : 0; // (timeline experiments only)
	if (tile < ntiles)
		rj_load_tile(s, tile, buf);
	while (tile < ntiles) {
		uint32_t item[RJ_P1_KEYS];
		if (tile < nfull)
			rj_decode_tile<HAS_PRESENT, true>(s, pr, buf, tile, item);
		else
			rj_decode_tile<HAS_PRESENT, false>(s, pr, buf, tile, item);
		uint32_t kept_at = 0;
		const uint32_t kept = rj_insert_items<false>(s, sm, pr, item, half, kept_at);
		tile += nslots;
		if (tile < ntiles)
			rj_load_tile(s, tile, buf);
		rj_round_end(s, pr, sm, kept, kept_at, half, rj_t_);
	}
	rj_drain(s, pr, sm, half);
}
