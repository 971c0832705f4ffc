// mdb_radix_pass2.cuh - pass 2 of the radix join+count (included by mdb_radix.cu).
//
// Per partition: count both sides' 2-byte remainders into packed BITS-wide counters in shared memory, check
// the field sums against the number of remainders read (a wrapped counter lowers the sum), then emit every
// key present on both sides with count cntA * cntB.
//
// What bounds it: the remainders are read once (2 B/key) and the groups written once (16 B/group); everything
// else is shared memory.
//   * a partition's remainders are contiguous per source rank (main stream + tail sectors, mdb_radix_types.cuh):
//     256-bit loads, a warp covers 1 KiB per instruction, two loads in flight per lane;
//   * groups are written with warp-ballot compaction: for one counter position at a time the matching lanes
//     write to consecutive output rows, so every store instruction covers a contiguous run;
//   * with 4-bit counters two CTAs share an SM: one emits while the other loads.
#pragma once
#ifndef RJ_P2_MLP
#define RJ_P2_MLP 4 // 32-byte vectors a thread has in flight while it counts (measured: 2 -> 0.597 ms, 3 -> 0.587, 4 -> 0.585)
#endif

struct RJOut {
	int nout;
	int is_count[4];
	int64_t *cells[4];
	unsigned long long *cursor;
	uint64_t cap;
};

// counters are packed BITS wide into 32-bit words: KPW keys per word
template <int BITS>
__device__ __forceinline__ void rj_count(uint32_t *cnt, uint32_t rem)
{
	constexpr int KPW = 32 / BITS;
	atomicAdd(&cnt[rem / KPW], 1u << ((rem % KPW) * BITS));
}

template <int BITS>
__device__ __forceinline__ uint32_t rj_nonzero_mask(uint32_t x)
{
	// top bit of every BITS-wide field set iff the field is non-zero
	constexpr uint32_t LOW = BITS == 4 ? 0x77777777u : 0x7f7f7f7fu;
	constexpr uint32_t TOP = BITS == 4 ? 0x88888888u : 0x80808080u;
	return (((x & LOW) + LOW) | x) & TOP;
}

template <int BITS>
__device__ __forceinline__ uint32_t rj_field_sum(uint32_t x)
{
	if (BITS == 4)
		x = (x & 0x0f0f0f0fu) + ((x >> 4) & 0x0f0f0f0fu);
	return __dp4a(x, 0x01010101u, 0u);
}

__device__ __forceinline__ void rj_load256(const void *p, uint32_t *w)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			: "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

// count the `ne` remainders starting at `run` (32-byte aligned; reading up to the next 32-byte boundary is safe).
// TAIL: `run` is a tail stream - every 32-byte sector holds up to 15 remainders and their number in entry 15
// (mdb_radix_types.cuh); returns how many remainders this thread counted (0 for main streams, whose length is known)
template <int BITS, int THREADS, bool TAIL>
__device__ __forceinline__ uint32_t rj_histogram_run(const uint16_t *__restrict__ run, uint32_t ne, uint32_t *cnt)
{
	uint32_t counted = 0;
	constexpr int MLP = RJ_P2_MLP; // 32-byte vectors in flight per thread
	const uint32_t nvec = (ne + 15u) / 16u;
	for (uint32_t v0 = threadIdx.x; v0 < nvec; v0 += THREADS * MLP) {
		uint32_t w[MLP][8];
#pragma unroll
		for (int u = 0; u < MLP; u++) // unconditional (address clamped): keeps the vectors in registers
			rj_load256(run + (size_t)min(v0 + u * THREADS, nvec - 1u) * 16u, w[u]);
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			const uint32_t v = v0 + u * THREADS;
			if (v >= nvec)
				continue;
			const uint32_t valid = TAIL ? min(15u, w[u][7] >> 16) : min(16u, ne - v * 16u);
			if (TAIL)
				counted += valid;
			if (valid == 16u) {
#pragma unroll
				for (int j = 0; j < 8; j++) {
					rj_count<BITS>(cnt, w[u][j] & 0xffffu);
					rj_count<BITS>(cnt, w[u][j] >> 16);
				}
			} else {
				// ragged last vector of a stream
#pragma unroll
				for (int j = 0; j < 16; j++)
					if ((uint32_t)j < valid)
						rj_count<BITS>(cnt, (w[u][j >> 1] >> ((j & 1) * 16)) & 0xffffu);
			}
		}
	}
	return counted;
}

// count the keys of rows [lo, hi) of a sorted column (all of them belong to the partition being counted)
template <int BITS, int THREADS>
// (base_lo = low word of kmin + p * width: the remainder of a key of partition p is key - base)
__device__ __forceinline__ void rj_histogram_sorted(const int64_t *__restrict__ keys, uint64_t lo, uint64_t hi, uint32_t base_lo,
		uint32_t *cnt)
{
	const uint64_t a = min((uint64_t)((lo + 3) & ~3ull), hi), b = max(a, (uint64_t)(hi & ~3ull)); // [a, b): whole 32-byte quads
	for (uint64_t i = lo + threadIdx.x; i < a; i += THREADS)
		rj_count<BITS>(cnt, (uint32_t)(unsigned long long)keys[i] - base_lo);
	for (uint64_t i = b + threadIdx.x; i < hi; i += THREADS)
		rj_count<BITS>(cnt, (uint32_t)(unsigned long long)keys[i] - base_lo);
	constexpr int MLP = 2;
	const uint64_t q0 = a / 4, nq = (b - a) / 4;
	for (uint64_t v0 = threadIdx.x; v0 < nq; v0 += THREADS * MLP) {
		uint32_t w[MLP][8];
#pragma unroll
		for (int u = 0; u < MLP; u++)
			rj_load256(keys + (q0 + min((uint64_t)(v0 + u * THREADS), nq - 1)) * 4, w[u]);
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			if (v0 + u * THREADS >= nq)
				continue;
#pragma unroll
			for (int j = 0; j < 4; j++)
				rj_count<BITS>(cnt, w[u][2 * j] - base_lo);
		}
	}
}

// all streams of partition p on one side; returns the number of remainders of the main streams (identical in every
// thread) and adds the remainders found in tail sectors to *tail_total (shared memory).
// counts[s] = {main, tail} entries of source s, fetched for all sources at once by rj_fetch_counts.
template <int BITS, int THREADS, bool MULTI>
__device__ __forceinline__ uint32_t rj_histogram_side(const RJRuns &r, const RJParams &pr, uint32_t p, uint32_t *cnt,
		const uint32_t (*counts)[2], uint32_t *tail_total)
{
	if (r.nsrc == 0) {
		const uint64_t lo = r.sorted_bnd[p], hi = r.sorted_bnd[p + 1];
		rj_histogram_sorted<BITS, THREADS>(r.sorted_keys, lo, hi, (uint32_t)(unsigned long long)pr.kmin + p * pr.width, cnt);
		return (uint32_t)(hi - lo);
	}
	uint32_t total = 0, in_tails = 0;
	for (int s = 0; s < r.nsrc; s++) {
		const uint32_t q = p - r.first[s];
		uint32_t n_main, n_tail;
		if (MULTI) {
			n_main = counts[s][0];
			n_tail = counts[s][1];
		} else {
			n_main = min(r.cursor[s][q * r.cur_stride[s]], r.cap);
			n_tail = min(r.tail_cursor[s][q * r.cur_stride[s]], r.tail_cap);
		}
		rj_histogram_run<BITS, THREADS, false>(r.stream[s] + (size_t)q * r.cap, n_main, cnt);
		in_tails += rj_histogram_run<BITS, THREADS, true>(r.tail[s] + (size_t)q * r.tail_cap, n_tail, cnt);
		total += n_main;
	}
	if (in_tails)
		atomicAdd(tail_total, in_tails);
	return total;
}

// Multi-GPU plans: a partition's remainders sit in 2 W short runs per side (main + tail stream of every source rank; with 8
// GPUs about 16 KiB each).  Walking them one after the other makes every run a round trip of its own - 32 dependent
// latencies per side and partition, which is what pass 2 cost at 8 GPUs (0.146 ms where 0.08 was its share).  Here the
// runs of a side are ONE index space: vector g belongs to the run whose prefix interval contains it, so all loads of a
// thread are independent and MLP of them are in flight whatever run they fall into.
struct RJRunList {
	const uint16_t *ptr[2 * RJ_MAX_RANKS];
	uint32_t end[2 * RJ_MAX_RANKS]; // exclusive prefix end, in 32-byte vectors
	uint32_t is_tail;               // bit r: run r is a tail stream (count in entry 15 of every sector)
	uint32_t n;
};

// (one thread) the runs of partition p on one side, from the counts fetched by rj_fetch_counts; returns the main-stream entries
__device__ __forceinline__ uint32_t rj_build_run_list(const RJRuns &r, uint32_t p, const uint32_t (*counts)[2], RJRunList *l)
{
	uint32_t n = 0, acc = 0, total_main = 0, tails = 0, remote = 0;
	for (int s = 0; s < r.nsrc; s++) {
		const uint32_t q = p - r.first[s];
		const uint32_t n_main = counts[s][0], n_tail = counts[s][1];
		if (s != r.self)
			remote += n_main + n_tail;
		if (n_main) {
			l->ptr[n] = r.stream[s] + (size_t)q * r.cap;
			acc += (n_main + 15u) / 16u;
			l->end[n++] = acc;
			total_main += n_main;
		}
		if (n_tail) {
			l->ptr[n] = r.tail[s] + (size_t)q * r.tail_cap;
			acc += (n_tail + 15u) / 16u;
			tails |= 1u << n;
			l->end[n++] = acc;
		}
	}
	l->n = n;
	l->is_tail = tails;
	if (remote && r.pulled_bytes)
		atomicAdd(r.pulled_bytes, 2ull * remote);
	return total_main;
}

template <int BITS, int THREADS>
__device__ __forceinline__ void rj_histogram_multi(const RJRunList &l, uint32_t *cnt, uint32_t *tail_total)
{
	constexpr int MLP = 4;
	const uint32_t total = l.n ? l.end[l.n - 1] : 0u;
	uint32_t in_tails = 0;
	for (uint32_t g0 = threadIdx.x; g0 < total; g0 += THREADS * MLP) {
		uint32_t w[MLP][8];
		uint32_t tail_of[MLP];
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			const uint32_t g = min(g0 + u * THREADS, total - 1u); // clamped: keeps the vectors in registers
			uint32_t run = 0;
			while (g >= l.end[run])
				run++;
			const uint32_t first = run ? l.end[run - 1] : 0u;
			tail_of[u] = (l.is_tail >> run) & 1u;
			rj_load256(l.ptr[run] + (size_t)(g - first) * 16u, w[u]);
		}
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			if (g0 + u * THREADS >= total)
				continue;
			if (!tail_of[u]) {
#pragma unroll
				for (int j = 0; j < 8; j++) {
					rj_count<BITS>(cnt, w[u][j] & 0xffffu);
					rj_count<BITS>(cnt, w[u][j] >> 16);
				}
			} else {
				const uint32_t valid = min(15u, w[u][7] >> 16);
				in_tails += valid;
#pragma unroll
				for (int j = 0; j < 15; j++)
					if ((uint32_t)j < valid)
						rj_count<BITS>(cnt, (w[u][j >> 1] >> ((j & 1) * 16)) & 0xffffu);
			}
		}
	}
	if (in_tails)
		atomicAdd(tail_total, in_tails);
}

// one thread per (source, main | tail): with 8 GPUs a partition has 16 streams per side, and 16 dependent cursor loads
// in a row would cost more than counting the streams
__device__ __forceinline__ void rj_fetch_counts(const RJRuns &r, uint32_t p, uint32_t (*counts)[2], uint32_t t)
{
	const uint32_t s = t >> 1, which = t & 1u;
	if ((int)s < r.nsrc) {
		const uint32_t q = (p - r.first[s]) * r.cur_stride[s];
		counts[s][which] = which ? min(r.tail_cursor[s][q], r.tail_cap) : min(r.cursor[s][q], r.cap);
	}
}

// ---- emit: one counter position at a time, matching lanes write consecutive rows (coalesced runs).
// The emit loop is 64 % of pass 2's instruction stream and pass 2 is bound by instruction issue (ncu: issue slots 66 % busy
// at 50 % occupancy), so the per-position body is written out: ONE predicate drives the ballot and both stores (no
// branch, no reconvergence point), a row address is one 32-bit x 8 multiply-add onto the warp's first row of the column,
// and only the low word of the key is computed (the caller guarantees that the partition does not straddle a multiple
// of 2^32).  17 instructions per position where the compiler's version of the same loop had 31.
template <int BITS, int THREADS>
__device__ __forceinline__ void rj_emit_lean(const uint32_t *cntA, const uint32_t *cntB, int words, long long key_base, int64_t *key_col,
		int64_t *cnt_col)
{
	constexpr int KPW = 32 / BITS;
	constexpr uint32_t FIELD = (1u << BITS) - 1u;
	const int tid = threadIdx.x;
	const uint32_t lt_mask = (1u << (tid & 31)) - 1u;
	const uint32_t key_base_lo = (uint32_t)(unsigned long long)key_base, key_hi = (uint32_t)((unsigned long long)key_base >> 32);
	// (the caller also guarantees that the warp's rows of either column lie inside one 4 GiB-aligned window: 32-bit row addresses)
	const uint32_t key_col_lo = (uint32_t)(uintptr_t)key_col, key_col_hi = (uint32_t)((uintptr_t)key_col >> 32);
	const uint32_t cnt_col_lo = (uint32_t)(uintptr_t)cnt_col, cnt_col_hi = (uint32_t)((uintptr_t)cnt_col >> 32);
	uint32_t pos = 0; // rows this warp has written for this partition
	for (int w0 = 0; w0 < words; w0 += THREADS) {
		const int w = w0 + tid;
		uint32_t x = 0, y = 0;
		if (w < words) {
			x = cntA[w];
			y = cntB[w];
		}
		const uint32_t m = rj_nonzero_mask<BITS>(x) & rj_nonzero_mask<BITS>(y);
		if (__ballot_sync(0xffffffffu, m != 0) == 0)
			continue;
		const uint32_t key_w_lo = key_base_lo + (uint32_t)w * KPW;
#pragma unroll
		for (int f = 0; f < KPW; f++) {
			const uint32_t cnt = ((x >> (f * BITS)) & FIELD) * ((y >> (f * BITS)) & FIELD);
			uint32_t bal;
			asm volatile("{\n\t"
					".reg .pred p;\n\t"
					".reg .b32 t, r, alo, blo;\n\t"
					".reg .b64 a, b;\n\t"
					"setp.ne.u32 p, %1, 0;\n\t"
					"vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
					"and.b32 t, %0, %2;\n\t"
					"popc.b32 t, t;\n\t"
					"add.u32 r, %3, t;\n\t"
					"mad.lo.u32 alo, r, 8, %4;\n\t"
					"mad.lo.u32 blo, r, 8, %6;\n\t"
					"mov.b64 a, {alo, %5};\n\t"
					"mov.b64 b, {blo, %7};\n\t"
					"@p st.global.v2.b32 [a], {%8, %9};\n\t"
					"@p st.global.v2.b32 [b], {%10, %11};\n\t"
					"}"
					: "=&r"(bal)
					: "r"(m & (1u << (f * BITS + BITS - 1))), "r"(lt_mask), "r"(pos), "r"(key_col_lo), "r"(key_col_hi), "r"(cnt_col_lo),
					  "r"(cnt_col_hi), "r"(key_w_lo + f), "r"(key_hi), "r"(cnt), "r"(0u)
					: "memory");
			pos += __popc(bal);
		}
	}
}

// any result layout, any partition: the plain version of the same loop
template <int BITS, int THREADS>
__device__ __forceinline__ void rj_emit_any(const uint32_t *cntA, const uint32_t *cntB, int words, long long key_base, const RJOut &out,
		unsigned long long row0)
{
	constexpr int KPW = 32 / BITS;
	constexpr uint32_t FIELD = (1u << BITS) - 1u;
	const int tid = threadIdx.x;
	const uint32_t lt_mask = (1u << (tid & 31)) - 1u;
	uint32_t pos = 0;
	for (int w0 = 0; w0 < words; w0 += THREADS) {
		const int w = w0 + tid;
		uint32_t x = 0, y = 0;
		if (w < words) {
			x = cntA[w];
			y = cntB[w];
		}
		const uint32_t m = rj_nonzero_mask<BITS>(x) & rj_nonzero_mask<BITS>(y);
		if (__ballot_sync(0xffffffffu, m != 0) == 0)
			continue;
		const long long key_w = key_base + (long long)w * KPW;
#pragma unroll
		for (int f = 0; f < KPW; f++) {
			const bool has = (m & (1u << (f * BITS + BITS - 1))) != 0;
			const uint32_t bal = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t r = pos + __popc(bal & lt_mask);
				const long long key = key_w + f;
				const long long cnt = (long long)(((x >> (f * BITS)) & FIELD) * ((y >> (f * BITS)) & FIELD));
#pragma unroll
				for (int o = 0; o < 4; o++)
					if (o < out.nout)
						out.cells[o][row0 + r] = out.is_count[o] ? cnt : key;
			}
			pos += __popc(bal);
		}
	}
}

// LAYOUT fixes the result columns at compile time (the emit loop is the largest part of this kernel's instruction
// stream): 0 = read RJOut at run time, 1 = [key, count], 2 = [count, key]
// MULTI: multi-GPU plan (several source streams per partition; their counts are fetched together)
template <int BITS, int THREADS, int LAYOUT, bool MULTI>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_radix_joincount(RJRuns a_param, RJRuns b_param, RJParams pr, RJOut out, uint32_t *__restrict__ part_counter)
{
	// the per-source arrays are indexed at run time: keep them in shared memory, not in a local copy of the parameters
	__shared__ RJRuns s_runs[2];
	for (uint32_t i = threadIdx.x; i < sizeof(RJRuns) / 4; i += THREADS) {
		reinterpret_cast<uint32_t*>(&s_runs[0])[i] = reinterpret_cast<const uint32_t*>(&a_param)[i];
		reinterpret_cast<uint32_t*>(&s_runs[1])[i] = reinterpret_cast<const uint32_t*>(&b_param)[i];
	}
	const RJRuns &a = s_runs[0], &b = s_runs[1];
	for (int r = 0; r < pr.n_peer_flags; r++)
		if (pr.peer_flags[r] != 0)
			return; // pass 1 failed on some rank: every rank skips pass 2 and reports it
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int KPW = 32 / BITS;
	constexpr int NWARPS = THREADS / 32;
	const int words = (int)((pr.width + KPW - 1) / KPW); // counters of one side (the fields beyond `width` stay zero)
	uint32_t *cntA = reinterpret_cast<uint32_t*>(smem_raw);
	uint32_t *cntB = cntA + words;
	__shared__ uint32_t s_part, s_sumA, s_sumB, s_tailA, s_tailB;
	__shared__ uint32_t s_counts[2][RJ_MAX_RANKS][2]; // [side][source][main | tail] entries of the current partition
	__shared__ RJRunList s_list[2];                   // multi-GPU plans: the runs of the current partition, per side
	__shared__ uint32_t s_main[2];
	__shared__ uint32_t s_warp[NWARPS + 1];
	__shared__ unsigned long long s_base;
	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;

	while (true) {
		if (tid == 0) {
			s_part = (uint32_t)pr.part_first + atomicAdd(part_counter, 1u);
			s_sumA = s_sumB = s_tailA = s_tailB = 0;
		}
		for (int w = tid; w < words * 2; w += THREADS)
			cntA[w] = 0; // cntB follows cntA
		__syncthreads();
		const uint32_t p = s_part;
		if (p >= (uint32_t)pr.part_end) // this rank owns partitions [part_first, part_end)
			break;
		if (MULTI) {
			if (tid < 2 * RJ_MAX_RANKS)
				rj_fetch_counts(a, p, s_counts[0], tid);
			else if (tid < 4 * RJ_MAX_RANKS)
				rj_fetch_counts(b, p, s_counts[1], tid - 2 * RJ_MAX_RANKS);
			__syncthreads();
			if (tid == 0)
				s_main[0] = rj_build_run_list(a, p, s_counts[0], &s_list[0]);
			else if (tid == 32)
				s_main[1] = rj_build_run_list(b, p, s_counts[1], &s_list[1]);
			__syncthreads();
		}

		// ---- count both sides
		uint32_t totA, totB;
		if (MULTI) {
			rj_histogram_multi<BITS, THREADS>(s_list[0], cntA, &s_tailA);
			rj_histogram_multi<BITS, THREADS>(s_list[1], cntB, &s_tailB);
			totA = s_main[0];
			totB = s_main[1];
		} else {
			totA = rj_histogram_side<BITS, THREADS, MULTI>(a, pr, p, cntA, s_counts[0], &s_tailA);
			totB = rj_histogram_side<BITS, THREADS, MULTI>(b, pr, p, cntB, s_counts[1], &s_tailB);
		}
		__syncthreads();

		// ---- checksum + number of groups of this partition
		uint32_t sumA = 0, sumB = 0, matches = 0;
		for (int w = tid; w < words; w += THREADS) {
			const uint32_t x = cntA[w], y = cntB[w];
			sumA += rj_field_sum<BITS>(x);
			sumB += rj_field_sum<BITS>(y);
			matches += __popc(rj_nonzero_mask<BITS>(x) & rj_nonzero_mask<BITS>(y));
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			sumA += __shfl_xor_sync(0xffffffffu, sumA, o);
			sumB += __shfl_xor_sync(0xffffffffu, sumB, o);
			matches += __shfl_xor_sync(0xffffffffu, matches, o);
		}
		if (lane == 0) {
			atomicAdd(&s_sumA, sumA);
			atomicAdd(&s_sumB, sumB);
			s_warp[warp] = matches; // groups found by this warp
		}
		__syncthreads();
		if (warp == 0) {
			uint32_t v = lane < NWARPS ? s_warp[lane] : 0u, incl = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o)
					incl += n;
			}
			if (lane < NWARPS)
				s_warp[lane] = incl - v; // exclusive offset of each warp
			if (lane == 31) {
				s_base = incl ? atomicAdd(out.cursor, (unsigned long long)incl) : 0ull;
				s_warp[NWARPS] = incl;
				if (s_sumA != totA + s_tailA || s_sumB != totB + s_tailB)
					atomicOr(pr.error_flag, RJ_ERR_COUNTER);
			}
		}
		__syncthreads();

		// ---- emit: one counter position at a time, matching lanes write consecutive rows (coalesced runs)
		const uint32_t total = s_warp[NWARPS];
		if (total && s_base + total <= out.cap) {
			const long long key_base = pr.kmin + (long long)((unsigned long long)p * pr.width);
			const unsigned long long row0 = s_base + s_warp[warp];
			// the keys of one partition differ in their low words only, unless the partition straddles a multiple of 2^32
			const uint32_t key_base_lo = (uint32_t)(unsigned long long)key_base;
			const bool one_hi = key_base_lo + (pr.width - 1u) >= key_base_lo; // (warp-uniform)
			int64_t *const key_col = out.cells[LAYOUT == 2 ? 1 : 0] + row0, *const cnt_col = out.cells[LAYOUT == 2 ? 0 : 1] + row0;
			// (the warp writes at most `total` rows per column: their byte addresses must not cross a multiple of 2^32)
			const bool near = (uint32_t)(uintptr_t)key_col + 8u * total >= (uint32_t)(uintptr_t)key_col &&
					(uint32_t)(uintptr_t)cnt_col + 8u * total >= (uint32_t)(uintptr_t)cnt_col;
			if (LAYOUT != 0 && one_hi && near && !pr.plain_emit)
				rj_emit_lean<BITS, THREADS>(cntA, cntB, words, key_base, key_col, cnt_col);
			else
				rj_emit_any<BITS, THREADS>(cntA, cntB, words, key_base, out, row0);
		}
		__syncthreads();
	}
}
