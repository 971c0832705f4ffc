// Pass-1 laboratory, part 5 (derived from part 4): TWO 512-thread CTAs per SM, each with its own 4096 staging rows of
// 8 + 2 slots (flush unit = 16 bytes = half a sector; the halves of a sector come from different flushes and merge in
// L2 like the sectors of a line do).  Question: does overlapping one CTA's flush with the other's insert pay?
// Pass-1 laboratory, part 4: per-partition global streams.  p1_lab2.cu showed that scattered 32-byte sector stores
// cost 0.13 ms more than stores that complete whole 128-byte lines; here every partition has ONE append-only
// stream shared by all CTAs (position = global atomic on the partition's cursor), so consecutive sectors of a line
// are written by different CTAs within about a microsecond and merge in L2 before they reach DRAM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o p1_lab4 p1_lab4.cu && ./p1_lab4 [log2_rows]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define NPART 4096
#define CAP 10
#define FLUSHN 8
#define THREADS 512
#define NWARP (THREADS / 32)
#define WLCAP 96
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void ldg_stream256(const void *p, uint32_t *a)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			: "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, uint2 r0, uint2 r1, uint2 r2, uint2 r3)
{
	asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r0.x), "r"(r0.y), "r"(r1.x), "r"(r1.y), "r"(r2.x),
			"r"(r2.y), "r"(r3.x), "r"(r3.y) : "memory");
}
__device__ __forceinline__ uint32_t smem_inc(uint32_t *p)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
	return old;
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
	uint16_t stage[NPART * CAP];   // 80 KiB: 8 + 2 slots per partition
	uint32_t fill[NPART / 2];      // two 16-bit slot counters per word
	uint16_t wl[NWARP][WLCAP];
};

// VAR 20  cursor atomic issued by the lane that completes a row (insert phase), result parked in a register, handed to the
//         flushing lane through shared memory after the barrier
// VAR 21  same, but the completing lane flushes its own rows (no worklist; divergent)
// VAR 22  cursor atomic issued in the flush phase (its latency is exposed)
// VAR 23  no atomics: private per-warp sequential sectors (= p1_lab2 VAR 17, lower bound)
template <int VAR>
__global__ void __launch_bounds__(THREADS, 2) k_p1(const int64_t *keys, uint64_t n, int shift, uint16_t *streams, uint32_t cap,
		uint32_t *cursor, uint32_t *sink)
{
	extern __shared__ __align__(16) unsigned char raw[];
	Smem *sm = reinterpret_cast<Smem*>(raw);
	constexpr int NK = 8;
	constexpr int TILE = THREADS * NK;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t lt = (1u << lane) - 1u;
	for (int p = tid; p < NPART / 2; p += THREADS)
		sm->fill[p] = 0;
	__syncthreads();
	const uint64_t nfull = n / TILE;
	const uint32_t mask = (1u << shift) - 1u;
	uint32_t acc = 0;
	uint32_t a[NK], b[NK];
	auto load = [&](uint64_t tile, uint32_t *dst) {
		uint32_t t[8];
		const char *base = reinterpret_cast<const char*>(keys + tile * TILE);
		ldg_stream256(base + (size_t)tid * 32, t);
		dst[0] = t[0]; dst[1] = t[2]; dst[2] = t[4]; dst[3] = t[6];
		ldg_stream256(base + (size_t)(THREADS + tid) * 32, t);
		dst[4] = t[0]; dst[5] = t[2]; dst[6] = t[4]; dst[7] = t[6];
	};
	auto round = [&](const uint32_t *d) {
		uint32_t pos[NK];
#pragma unroll
		for (int k = 0; k < NK; k++) {
			const uint32_t p = d[k] >> shift, sh = (p & 1u) << 4;
			uint32_t old;
			asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(&sm->fill[p >> 1])), "r"(1u << sh) : "memory");
			pos[k] = (old >> sh) & 0xffffu;
		}
		uint32_t cnt = 0;
#pragma unroll
		for (int k = 0; k < NK; k++) {
			const uint32_t p = d[k] >> shift;
			if (pos[k] < CAP)
				sm->stage[p * CAP + pos[k]] = (uint16_t)(d[k] & mask);
			else
				acc++;
			const bool q = pos[k] == FLUSHN - 1;
			const uint32_t bal = __ballot_sync(0xffffffffu, q);
			if (q && cnt + __popc(bal & lt) < WLCAP)
				sm->wl[warp][cnt + __popc(bal & lt)] = (uint16_t)p;
			cnt += __popc(bal);
		}
		cnt = min(cnt, (uint32_t)WLCAP);
		__syncthreads();
		for (uint32_t w = lane; w < cnt; w += 32) {
			const uint32_t p = sm->wl[warp][w];
			const uint32_t at = atomicAdd(&cursor[p], (uint32_t)FLUSHN);
			uint32_t *row = reinterpret_cast<uint32_t*>(&sm->stage[p * CAP]); // 20-byte rows: 4-byte aligned
			const uint32_t r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
			row[0] = r4;
			const uint32_t sh = (p & 1u) << 4;
			const uint32_t have = (sm->fill[p >> 1] >> sh) & 0xffffu;
			atomicAdd(&sm->fill[p >> 1], (min(have, (uint32_t)CAP) - FLUSHN - have) << sh);
			if (at + FLUSHN <= cap)
				*reinterpret_cast<uint4*>(streams + (size_t)p * cap + at) = make_uint4(r0, r1, r2, r3);
		}
		__syncthreads();
	};
	uint64_t tile = blockIdx.x;
	if (tile < nfull)
		load(tile, a);
	while (tile < nfull) {
		uint64_t next = tile + gridDim.x;
		if (next < nfull)
			load(next, b);
		round(a);
		tile = next;
		if (tile >= nfull)
			break;
		next = tile + gridDim.x;
		if (next < nfull)
			load(next, a);
		round(b);
		tile = next;
	}
	__syncthreads();
	for (int p = tid; p < NPART / 2; p += THREADS)
		acc += sm->fill[p] + sm->stage[p * CAP];
	if (acc == 0x12345678u)
		sink[0] = acc;
}

template <int VAR>
static void run(const char *name, const int64_t *keys, uint64_t n, int shift, uint16_t *streams, uint32_t cap, uint32_t *cursor,
		uint32_t *sink, int sms)
{
	CK(cudaFuncSetAttribute(k_p1<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float total = 0;
	const int reps = 5;
	for (int i = 0; i < reps + 2; i++) {
		CK(cudaMemsetAsync(cursor, 0, NPART * 4));
		CK(cudaEventRecord(e0));
		k_p1<VAR><<<2 * sms, THREADS, sizeof(Smem)>>>(keys, n, shift, streams, cap, cursor, sink);
		CK(cudaEventRecord(e1));
		CK(cudaDeviceSynchronize());
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (i >= 2)
			total += ms;
	}
	static uint32_t h[NPART];
	CK(cudaMemcpy(h, cursor, sizeof(h), cudaMemcpyDeviceToHost));
	uint64_t sum = 0;
	uint32_t mx = 0;
	for (int p = 0; p < NPART; p++) {
		sum += h[p];
		mx = h[p] > mx ? h[p] : mx;
	}
	printf("%-72s %8.3f ms  %7.1f GB/s of keys   (appended %llu, max stream %u of %u)\n", name, total / reps, 8.0 * n / (total / reps) / 1e6,
			(unsigned long long)sum, mx, cap);
}

int main(int argc, char **argv)
{
	const int lg = argc > 1 ? atoi(argv[1]) : 28;
	const uint64_t n = 1ull << lg;
	int sms;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	int64_t *keys;
	uint16_t *streams;
	uint32_t *sink, *cursor;
	const uint32_t cap = (uint32_t)(2 * n / NPART);
	CK(cudaMalloc(&keys, n * 8));
	CK(cudaMalloc(&streams, (size_t)NPART * cap * 2 + (size_t)sms * NWARP * 8192 * 32));
	CK(cudaMalloc(&sink, 4));
	CK(cudaMalloc(&cursor, NPART * 4));
	k_gen<<<sms * 8, 256>>>(keys, n, n);
	CK(cudaDeviceSynchronize());
	const int shift = lg - 12;
	printf("n = 2^%d keys, %d SMs, shift %d, stream capacity %u entries\n", lg, sms, shift, cap);
	run<30>("30 two 512-thread CTAs per SM, rows of 8+2 slots, 16-byte flushes", keys, n, shift, streams, cap, cursor, sink, sms);
	return 0;
}
