// mdb_radix_pass2.cuh - pass 2 of the radix join+count (included by mdb_radix.cu).
//
// Per partition: count both sides' 2-byte remainders into packed BITS-wide counters in shared memory, check
// the field sums against the number of remainders read (a wrapped counter lowers the sum), then emit every
// key present on both sides with count cntA * cntB.
//
// What bounds it (ncu, profiles/): the remainders are read once (2 B/key) and the groups written once
// (16 B/group); everything else is shared memory.  Three things keep it off the latency floor:
//   * the chunk descriptors (chunk id, entries) of a partition are fetched into shared memory up front, so the
//     only dependent global access per chunk is the 512-byte chunk itself, several chunks in flight per warp;
//   * groups are written with warp-ballot compaction: for one counter position at a time the matching lanes
//     write to consecutive output rows, so every store instruction covers a contiguous run;
//   * with 4-bit counters two CTAs share an SM: one emits while the other loads.
#pragma once

#define RJ_DESC_CAP 1024 // chunk descriptors staged per side and batch

__global__ void k_radix_dir_fill(RJSide s)
{
	const RJTarget &t = s.dst[s.self];
	uint32_t nchunks = min(*t.pool_next, s.pool_chunks);
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += gridDim.x * blockDim.x) {
		uint32_t p = t.chunk_part[c];
		if (p == 0xffffu)
			continue; // id reserved by a CTA but never used
		uint32_t pos = atomicAdd(&s.dir_fill[p], 1u);
		RJDesc d;
		d.off16 = c * (RJ_CHUNK / 8);
		d.ne = t.chunk_entries[c];
		s.dir[s.dir_off[p] + pos] = d;
	}
}

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

// count the remainders of chunks [0, n) described in shared memory into cnt; returns (per warp) entries seen
template <int BITS, int THREADS>
__device__ __forceinline__ uint32_t rj_histogram(const uint16_t *__restrict__ pool, const RJDesc *descs, uint32_t n, uint32_t *cnt)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	constexpr int NWARPS = THREADS / 32;
	constexpr int MLP = 4; // chunks in flight per warp
	uint32_t seen = 0;
	for (uint32_t c = warp * MLP; c < n; c += NWARPS * MLP) {
		int4 v[MLP];
		uint32_t ne[MLP];
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			ne[u] = 0;
			v[u] = make_int4(0, 0, 0, 0);
			if (c + u < n) {
				const RJDesc d = descs[c + u];
				ne[u] = d.ne;
				if ((uint32_t)lane * 8 < d.ne) // partially filled chunks: only touch the sectors that hold data
					v[u] = mdb_ldg_stream(reinterpret_cast<const int4*>(pool) + (size_t)d.off16 + lane);
			}
		}
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			const uint32_t w[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z, (uint32_t)v[u].w};
			const uint32_t first = (uint32_t)lane * 8;
			if (first + 8 <= ne[u]) {
#pragma unroll
				for (int j = 0; j < 4; j++) {
					rj_count<BITS>(cnt, w[j] & 0xffffu);
					rj_count<BITS>(cnt, w[j] >> 16);
				}
			} else if (first < ne[u]) {
				// the one lane holding the ragged tail of a drained chunk
				for (uint32_t j = 0; first + j < ne[u]; j++)
					rj_count<BITS>(cnt, (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
			}
			seen += ne[u];
		}
	}
	return seen;
}

template <int BITS, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_radix_joincount(RJSide a, RJSide b, RJParams pr, RJOut out, uint32_t *__restrict__ part_counter)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int KPW = 32 / BITS;
	constexpr int NWARPS = THREADS / 32;
	constexpr uint32_t FIELD = (1u << BITS) - 1u;
	const int D = 1 << pr.shift;
	const int words = D >= KPW ? D / KPW : 1;
	uint32_t *cntA = reinterpret_cast<uint32_t*>(smem_raw);
	uint32_t *cntB = cntA + words;
	RJDesc *descA = reinterpret_cast<RJDesc*>(cntB + words);
	RJDesc *descB = descA + RJ_DESC_CAP;
	__shared__ uint32_t s_part, s_totA, s_totB, s_sumA, s_sumB;
	__shared__ uint32_t s_warp[NWARPS + 1];
	__shared__ unsigned long long s_base;
	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;

	while (true) {
		if (tid == 0) {
			s_part = (uint32_t)pr.part_first + atomicAdd(part_counter, 1u);
			s_totA = s_totB = s_sumA = s_sumB = 0;
		}
		__syncthreads();
		const uint32_t p = s_part;
		if (p >= (uint32_t)pr.part_end) // this rank owns partitions [part_first, part_end)
			break;

		// ---- count both sides (descriptor batches of RJ_DESC_CAP chunks per side; one batch is the common case)
		const uint64_t a0 = a.dir_off[p], a1 = a.dir_off[p + 1], b0 = b.dir_off[p], b1 = b.dir_off[p + 1];
		for (int w = tid; w < words * 2; w += THREADS)
			cntA[w] = 0; // cntB follows cntA
		uint32_t seenA = 0, seenB = 0;
		for (uint64_t off = 0; a0 + off < a1 || b0 + off < b1; off += RJ_DESC_CAP) {
			const uint32_t na = a0 + off < a1 ? (uint32_t)min((uint64_t)RJ_DESC_CAP, a1 - a0 - off) : 0u;
			const uint32_t nb = b0 + off < b1 ? (uint32_t)min((uint64_t)RJ_DESC_CAP, b1 - b0 - off) : 0u;
			for (uint32_t i = tid; i < na; i += THREADS)
				descA[i] = a.dir[a0 + off + i];
			for (uint32_t i = tid; i < nb; i += THREADS)
				descB[i] = b.dir[b0 + off + i];
			__syncthreads();
			seenA += rj_histogram<BITS, THREADS>(a.pool, descA, na, cntA);
			seenB += rj_histogram<BITS, THREADS>(b.pool, descB, nb, cntB);
			__syncthreads();
		}
		if (lane == 0) {
			if (seenA)
				atomicAdd(&s_totA, seenA);
			if (seenB)
				atomicAdd(&s_totB, seenB);
		}

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
				if (s_sumA != s_totA || s_sumB != s_totB)
					atomicOr(pr.error_flag, RJ_ERR_COUNTER);
			}
		}
		__syncthreads();

		// ---- emit: one counter position at a time, matching lanes write consecutive rows (coalesced runs)
		const uint32_t total = s_warp[NWARPS];
		unsigned long long pos = s_base + s_warp[warp];
		if (total && s_base + total <= out.cap) {
			const long long key_base = pr.kmin + (long long)((unsigned long long)p << pr.shift);
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
#pragma unroll
				for (int f = 0; f < KPW; f++) {
					const bool has = (m >> (f * BITS + BITS - 1)) & 1u;
					const uint32_t bal = __ballot_sync(0xffffffffu, has);
					if (has) {
						const unsigned long long row = pos + __popc(bal & lt_mask);
						const long long key = key_base + (long long)w * KPW + f;
						const long long cnt = (long long)((x >> (f * BITS)) & FIELD) * (long long)((y >> (f * BITS)) & FIELD);
#pragma unroll
						for (int o = 0; o < 4; o++)
							if (o < out.nout)
								out.cells[o][row] = out.is_count[o] ? cnt : key;
					}
					pos += __popc(bal);
				}
			}
		}
		__syncthreads();
	}
}
