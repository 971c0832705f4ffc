// Microbenchmark behind the design of mdb_radix.cu: per-SM throughput of the primitives a partition / histogram
// kernel can be built from (random 16-bit bins, 1024 threads per SM, one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define THREADS 1024
#define ITERS 256
#define KEYS_PER_ITER 8

__device__ __forceinline__ uint32_t rnd(uint32_t &s)
{
	s ^= s << 13; s ^= s >> 17; s ^= s << 5;
	return s;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) bench(uint32_t *gscratch, uint32_t *sink, int bins_log2)
{
	extern __shared__ uint32_t sm[];
	const uint32_t mask = (1u << bins_log2) - 1u;
	uint32_t *g = gscratch + (size_t)blockIdx.x * (1u << bins_log2);
	for (int i = threadIdx.x; i < (1 << bins_log2); i += THREADS)
		sm[i] = 0;
	__syncthreads();
	uint32_t s = 0x9e3779b9u * (blockIdx.x * THREADS + threadIdx.x + 1);
	uint32_t acc = 0;
	for (int it = 0; it < ITERS; it++) {
		uint32_t k[KEYS_PER_ITER];
#pragma unroll
		for (int j = 0; j < KEYS_PER_ITER; j++)
			k[j] = rnd(s) & mask;
#pragma unroll
		for (int j = 0; j < KEYS_PER_ITER; j++) {
			if (MODE == 0) {            // shared atomic, result unused
				atomicAdd(&sm[k[j]], 1u);
			} else if (MODE == 1) {     // shared atomic, result used
				acc += atomicAdd(&sm[k[j]], 1u);
			} else if (MODE == 2) {     // global reduction into an L2-resident per-CTA array
				atomicAdd(&g[k[j]], 1u);
			} else if (MODE == 3) {     // half shared, half global
				if (j & 1) atomicAdd(&g[k[j]], 1u); else atomicAdd(&sm[k[j]], 1u);
			} else if (MODE == 4) {     // plain 2-byte shared store
				reinterpret_cast<uint16_t*>(sm)[k[j]] = (uint16_t)it;
			} else if (MODE == 5) {     // non-atomic read-modify-write
				sm[k[j]] = sm[k[j]] + 1;
			} else if (MODE == 6) {     // hardware match
				acc += __popc(__match_any_sync(0xffffffffu, k[j] & 0xfffu));
			} else if (MODE == 7) {     // 12 ballots (software match on a 12-bit digit)
				uint32_t m = 0xffffffffu;
#pragma unroll
				for (int b = 0; b < 12; b++) {
					uint32_t v = __ballot_sync(0xffffffffu, (k[j] >> b) & 1u);
					m &= ((k[j] >> b) & 1u) ? v : ~v;
				}
				acc += __popc(m);
			} else if (MODE == 8) {     // global atomic with result used (slot assignment through L2)
				acc += atomicAdd(&g[k[j]], 1u);
			} else if (MODE == 9) {     // claim-by-store + read back (2 plain shared ops)
				reinterpret_cast<uint16_t*>(sm)[k[j]] = (uint16_t)threadIdx.x;
				acc += reinterpret_cast<volatile uint16_t*>(sm)[k[j] ^ 1];
			}
		}
	}
	if (acc == 0xdeadbeefu)
		sink[0] = acc;
}

template <int MODE>
static void run(const char *name, uint32_t *g, uint32_t *sink, int sms, int bins_log2)
{
	size_t smem = sizeof(uint32_t) << bins_log2;
	cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	bench<MODE><<<sms, THREADS, smem>>>(g, sink, bins_log2);
	cudaEventRecord(a);
	bench<MODE><<<sms, THREADS, smem>>>(g, sink, bins_log2);
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	double ops = (double)sms * THREADS * ITERS * KEYS_PER_ITER;
	int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	printf("%-46s bins=2^%-2d %8.3f ms  %7.2f Gops/s  %6.3f cycles/op/SM (at %d MHz)  err=%s\n", name, bins_log2, ms, ops / ms / 1e6,
	       ms * 1e-3 * clk * 1e3 / (ops / sms), clk / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
	int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	uint32_t *g, *sink;
	cudaMalloc(&g, (size_t)sms * (1 << 16) * sizeof(uint32_t));
	cudaMemset(g, 0, (size_t)sms * (1 << 16) * sizeof(uint32_t));
	cudaMalloc(&sink, 64);
	for (int bl : {12, 14}) {
		run<0>("ATOMS add, result unused", g, sink, sms, bl);
		run<1>("ATOMS add, result used", g, sink, sms, bl);
		run<2>("RED.global add (L2-resident per-CTA array)", g, sink, sms, bl);
		run<3>("half ATOMS + half RED.global", g, sink, sms, bl);
		run<4>("STS.U16 random", g, sink, sms, bl);
		run<5>("LDS + STS random (non-atomic RMW)", g, sink, sms, bl);
		run<6>("match.any.sync on 12-bit digit", g, sink, sms, bl);
		run<7>("12 x ballot software match", g, sink, sms, bl);
		run<8>("ATOM.global add, result used", g, sink, sms, bl);
		run<9>("STS.U16 claim + LDS.U16 read back", g, sink, sms, bl);
	}
	return 0;
}
