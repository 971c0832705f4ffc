// Pass-1 laboratory, part 6 (NOT YET RUN - written at the end of round 1 when the GPU budget was spent).
//
// Question: does it pay to take the flush off the insert warps' critical path?  In the production kernel a round is
// insert (52 %) -> barrier -> flush (35 %) -> barrier; the LSU work of both phases together is about a quarter of the round,
// the rest is latency and barrier skew.  Here the CTA is warp-specialised:
//   * 28 INSERT warps never wait for a flush.  A row is a RING of 20 slots.  One 32-bit word per row holds
//     [flushed entries : 16 | claimed entries : 16]; `old = atom.add(word, 1)` returns the claim index AND how far the
//     flush has got, so the overflow guard (claim - flushed < 20) costs no extra shared-memory access.  The lane whose claim
//     is the 16th of a sector puts (row, claim) on its warp's worklist (ballot compaction, no atomics).  At the end of a
//     round the warp ARRIVES on a named barrier (bar.arrive: it does not wait).
//   * 4 FLUSH warps wait on that barrier (bar.sync), so every slot of every listed sector has been written (all 16 claims
//     of a sector precede the completing claim, and all stores of the round precede the arrive).  One lane per sector:
//     stream position from the partition's global cursor, 32-byte sector store, then `atom.add(word, 16 << 16)` publishes
//     the 16 freed slots.  Meanwhile the insert warps are already in the next round.
//   * worklists are double-buffered by round parity; a second pair of named barriers (flush warps arrive, insert warps sync)
//     keeps a worklist from being overwritten before it has been consumed (normally already satisfied: two rounds later).
// What the lab does NOT do (count only): keys that find their ring full are dropped and counted ("overflow") - production
// would send them to the tail stream and flush the sector with a hole through the slow path; partial rows at the end are
// not drained.  Compare with p1_lab4 VAR 22 (same streams, synchronous flush) in the same gpurun call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o p1_lab6 p1_lab6.cu && ./p1_lab6 [log2_rows]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define NPART 4096
#define RING 20
#define SECT 16
#define THREADS 1024
#define A_WARPS 28
#define B_WARPS 4
#define A_THREADS (A_WARPS * 32)
#define NK 8
#define TILE (A_THREADS * NK)
#define WLCAP 64
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void ldg_stream256(const void *p, uint32_t *a)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			: "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, const uint32_t *r)
{
	asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
			"r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint32_t smem_add(uint32_t *p, uint32_t v)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
	return old;
}
__device__ __forceinline__ void bar_arrive(int id, int count)
{
	asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int count)
{
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__global__ void k_gen(int64_t *k, uint64_t n, uint64_t domain)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x1234567;
		x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
		k[i] = (int64_t)(x % domain);
	}
}

struct Smem {
	uint16_t stage[NPART * RING];        // 160 KiB: ring of 20 slots per partition (40-byte rows, 4-byte aligned)
	uint32_t word[NPART];                // 16 KiB: [flushed : 16 | claimed : 16]
	uint32_t wl[2][A_WARPS][WLCAP];      // 14 KiB: (claim << 12 | row) of every sector completed in a round
	uint32_t wl_n[2][A_WARPS];
};
static_assert(sizeof(Smem) <= 195 * 1024, "more shared memory than this costs L1 (p1_lab_4c)");

// named barriers: 1 + parity = "inserts of the round are done", 3 + parity = "worklist of the round has been consumed"
template <int VAR>
__global__ void __launch_bounds__(THREADS, 1) k_p1(const int64_t *keys, uint64_t n, int shift, uint16_t *streams, uint32_t cap,
		uint32_t *cursor, unsigned long long *stats)
{
	extern __shared__ __align__(16) unsigned char raw[];
	Smem *sm = reinterpret_cast<Smem*>(raw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t lt = (1u << lane) - 1u;
	for (int p = tid; p < NPART; p += THREADS)
		sm->word[p] = 0;
	__syncthreads();
	const uint64_t nfull = n / TILE;
	const uint32_t rounds = blockIdx.x < nfull ? (uint32_t)((nfull - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
	const uint32_t mask = (1u << shift) - 1u;

	if (warp < A_WARPS) {
		// ------------------------------------------------------------------ insert warps
		uint32_t a[NK], b[NK], overflow = 0, lost = 0;
		auto load = [&](uint64_t tile, uint32_t *dst) {
			uint32_t t[8];
			const char *base = reinterpret_cast<const char*>(keys + tile * TILE);
			ldg_stream256(base + (size_t)tid * 32, t);
			dst[0] = t[0]; dst[1] = t[2]; dst[2] = t[4]; dst[3] = t[6];
			ldg_stream256(base + (size_t)(A_THREADS + tid) * 32, t);
			dst[4] = t[0]; dst[5] = t[2]; dst[6] = t[4]; dst[7] = t[6];
		};
		auto round = [&](uint32_t r, const uint32_t *d) {
			const int q = r & 1;
			if (r >= 2)
				bar_sync(3 + q, THREADS); // the flush warps are done with this worklist buffer (round r - 2)
			uint32_t old[NK];
#pragma unroll
			for (int k = 0; k < NK; k++)
				old[k] = smem_add(&sm->word[d[k] >> shift], 1u);
			uint32_t cnt = 0;
#pragma unroll
			for (int k = 0; k < NK; k++) {
				const uint32_t p = d[k] >> shift;
				const uint32_t c = old[k] & 0xffffu, f = old[k] >> 16;
				const bool fits = ((c - f) & 0xffffu) < RING;
				if (fits)
					sm->stage[p * RING + c % RING] = (uint16_t)(d[k] & mask);
				else
					overflow++;
				const bool done = (c & (SECT - 1)) == SECT - 1; // my claim is the last of its sector
				const uint32_t bal = __ballot_sync(0xffffffffu, done);
				const uint32_t at = cnt + __popc(bal & lt);
				if (done) {
					if (at < WLCAP)
						sm->wl[q][warp][at] = (c << 12) | p;
					else
						lost++;
				}
				cnt += __popc(bal);
			}
			if (lane == 0)
				sm->wl_n[q][warp] = min(cnt, (uint32_t)WLCAP);
			bar_arrive(1 + q, THREADS); // does not wait
		};
		uint64_t tile = blockIdx.x;
		if (rounds)
			load(tile, a);
		for (uint32_t r = 0; r < rounds; r += 2) {
			if (r + 1 < rounds)
				load(tile + gridDim.x, b);
			round(r, a);
			if (r + 1 >= rounds)
				break;
			if (r + 2 < rounds)
				load(tile + 2ull * gridDim.x, a);
			round(r + 1, b);
			tile += 2ull * gridDim.x;
		}
		if (overflow)
			atomicAdd(&stats[0], (unsigned long long)overflow);
		if (lost)
			atomicAdd(&stats[1], (unsigned long long)lost);
	} else {
		// ------------------------------------------------------------------ flush warps
		const int bw = warp - A_WARPS;
		uint32_t flushed = 0;
		for (uint32_t r = 0; r < rounds; r++) {
			const int q = r & 1;
			bar_sync(1 + q, THREADS); // every insert warp has finished round r: the listed sectors are complete
			if (VAR != 41) {
				for (int w = bw; w < A_WARPS; w += B_WARPS) {
					const uint32_t cnt = sm->wl_n[q][w];
					for (uint32_t i = lane; i < cnt; i += 32) {
						const uint32_t e = sm->wl[q][w][i];
						const uint32_t p = e & 0xfffu, c = e >> 12;
						const uint32_t at = VAR == 42 ? 0u : atomicAdd(&cursor[p], (uint32_t)SECT);
						// the sector's 16 slots start at claim c - 15; ring positions are even-aligned pairs (RING and SECT are even)
						const uint32_t first = ((c - (SECT - 1)) & 0xffffu) % RING;
						const uint32_t *row = reinterpret_cast<const uint32_t*>(&sm->stage[p * RING]);
						uint32_t v[8];
#pragma unroll
						for (int j = 0; j < 8; j++) {
							uint32_t pos = first / 2 + j;
							pos = pos >= RING / 2 ? pos - RING / 2 : pos;
							v[j] = row[pos];
						}
						smem_add(&sm->word[p], (uint32_t)SECT << 16); // the 16 slots are free again
						if (at + SECT <= cap)
							stg256(streams + (size_t)p * cap + at, v);
						flushed++;
					}
				}
			}
			if (r + 2 < rounds)
				bar_arrive(3 + q, THREADS);
		}
		if (flushed)
			atomicAdd(&stats[2], (unsigned long long)flushed);
	}
}

template <int VAR>
static void run(const char *name, const int64_t *keys, uint64_t n, int shift, uint16_t *streams, uint32_t cap, uint32_t *cursor,
		unsigned long long *stats, int sms)
{
	CK(cudaFuncSetAttribute(k_p1<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float total = 0;
	const int reps = 5;
	for (int i = 0; i < reps + 2; i++) {
		CK(cudaMemsetAsync(cursor, 0, NPART * 4));
		CK(cudaMemsetAsync(stats, 0, 4 * 8));
		CK(cudaEventRecord(e0));
		k_p1<VAR><<<sms, THREADS, sizeof(Smem)>>>(keys, n, shift, streams, cap, cursor, stats);
		CK(cudaEventRecord(e1));
		CK(cudaDeviceSynchronize());
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (i >= 2)
			total += ms;
	}
	static uint32_t h[NPART];
	unsigned long long st[4];
	CK(cudaMemcpy(h, cursor, sizeof(h), cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(st, stats, sizeof(st), cudaMemcpyDeviceToHost));
	uint64_t sum = 0;
	uint32_t mx = 0;
	for (int p = 0; p < NPART; p++) {
		sum += h[p];
		mx = h[p] > mx ? h[p] : mx;
	}
	const uint64_t used = n / TILE * TILE;
	printf("%-64s %8.3f ms  %7.1f GB/s of keys  appended %llu of %llu (%.2f %%), ring overflow %llu, worklist lost %llu, sectors %llu, max stream %u of %u\n",
			name, total / reps, 8.0 * used / (total / reps) / 1e6, (unsigned long long)sum, (unsigned long long)used, 100.0 * sum / used,
			st[0], st[1], st[2], mx, cap);
}

int main(int argc, char **argv)
{
	const int lg = argc > 1 ? atoi(argv[1]) : 28;
	const uint64_t n = 1ull << lg;
	int sms;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	int64_t *keys;
	uint16_t *streams;
	uint32_t *cursor;
	unsigned long long *stats;
	const uint32_t cap = (uint32_t)(2 * n / NPART);
	CK(cudaMalloc(&keys, n * 8));
	CK(cudaMalloc(&streams, (size_t)NPART * cap * 2));
	CK(cudaMalloc(&stats, 4 * 8));
	CK(cudaMalloc(&cursor, NPART * 4));
	k_gen<<<sms * 8, 256>>>(keys, n, n);
	CK(cudaDeviceSynchronize());
	const int shift = lg - 12;
	printf("n = 2^%d keys, %d SMs, shift %d, stream capacity %u entries, tile %d keys, %zu bytes of shared memory\n", lg, sms, shift, cap,
			TILE, sizeof(Smem));
	run<40>("40 warp-specialised: 28 insert warps, 4 flush warps, ring rows", keys, n, shift, streams, cap, cursor, stats, sms);
	run<41>("41 same, flush warps idle (insert cost alone; rings overflow)", keys, n, shift, streams, cap, cursor, stats, sms);
	run<42>("42 same, no cursor atomics (all sectors of a row overwrite)", keys, n, shift, streams, cap, cursor, stats, sms);
	return 0;
}
