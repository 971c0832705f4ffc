// Pass-1 laboratory, part 3: the PRODUCTION kernel (mdb_radix_pass1.cuh) run in isolation on the bench workload,
// so variants can be timed without the rest of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../../include -I../../midoridb_b200/csrc -o p1_lab3 p1_lab3.cu
#include "mdb_common.cuh"
#include "mdb_radix_types.cuh"
#include "mdb_radix_pass1.cuh"

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
#ifndef P1_ALL_HINTS
#define P1_ALL_HINTS 0
#endif
#ifndef P1_W16
#define P1_W16 false   // -DP1_W16=true: the variant for partitions of exactly 2^16 key values (the lab uses shift 16)
#endif

__global__ void k_gen(int64_t *k, uint64_t n, uint64_t domain)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x1234567;
		x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
		k[i] = (int64_t)(x % domain);
	}
}

int main(int argc, char **argv)
{
	const int lg = argc > 1 ? atoi(argv[1]) : 28;
	const uint64_t n = 1ull << lg;
	int sms;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	int64_t *keys;
	CK(cudaMalloc(&keys, n * 8));
	k_gen<<<sms * 8, 256>>>(keys, n, n);
	RJSide s;
	memset(&s, 0, sizeof(s));
	s.keys = keys;
	s.n = n;
	s.all_in_range = 1;
	s.cap = (uint32_t)(2 * (n / 4096) + 2048);
	s.tail_cap = (sms * RJ_FLUSH + s.cap / 16 + 15u) & ~15u;
	CK(cudaMalloc(&s.stream, (size_t)4096 * s.cap * 2));
	CK(cudaMalloc(&s.tail, (size_t)4096 * s.tail_cap * 2));
	CK(cudaMalloc(&s.cursor, (size_t)RJ_MAX_PART * RJ_CUR_STRIDE * 4));
	s.tail_cursor = s.cursor + 1;
	uint32_t *flag;
	CK(cudaMalloc(&flag, 8));
	CK(cudaMemset(flag, 0, 8));
	RJParams pr;
	memset(&pr, 0, sizeof(pr));
	pr.kmin = 0;
	pr.range = n;
	pr.width = 1u << (lg - 12);
	pr.magic = (uint32_t)((1ull << 32) / pr.width);
	pr.nparts = 4096;
	pr.part_end = 4096;
	pr.error_flag = flag;
	CK(cudaFuncSetAttribute(k_radix_partition_fast<P1_W16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
	CK(cudaFuncSetAttribute(k_radix_partition<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	printf("n = 2^%d keys, %d SMs, smem %zu\n", lg, sms, sizeof(RJP1Smem));
	for (int variant = 0; variant < ((RJ_LAB & 8) ? 1 : 9); variant += ((variant == 1 && !P1_ALL_HINTS) ? 7 : 1)) {
		s.hints = variant < 8 ? (uint32_t)variant : 0;
		float total = 0;
		const int reps = 5;
		for (int i = 0; i < reps + 2; i++) {
			CK(cudaMemsetAsync(s.cursor, 0, (size_t)RJ_MAX_PART * RJ_CUR_STRIDE * 4));
			CK(cudaEventRecord(e0));
			if (variant < 8)
				k_radix_partition_fast<P1_W16><<<sms * RJ_SPLIT, RJ_P1_THREADS, sizeof(RJP1Smem)>>>(s, pr);
			else
				k_radix_partition<false><<<sms * RJ_SPLIT, RJ_P1_THREADS, sizeof(RJP1Smem)>>>(s, pr);
			CK(cudaEventRecord(e1));
			CK(cudaDeviceSynchronize());
			float ms;
			CK(cudaEventElapsedTime(&ms, e0, e1));
			if (i >= 2)
				total += ms;
		}
		uint32_t h[2];
		static uint32_t cur[RJ_MAX_PART * RJ_CUR_STRIDE];
		CK(cudaMemcpy(h, flag, 8, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(cur, s.cursor, sizeof(cur), cudaMemcpyDeviceToHost));
		unsigned long long total_entries = 0;
		for (int i = 0; i < RJ_MAX_PART * RJ_CUR_STRIDE; i++)
			total_entries += cur[i];
		// every key must be in exactly one place: main streams hold whole sectors, tail sectors carry their count in entry 15
		unsigned long long keys_found = 0, tail_sectors = 0;
		{
			static uint16_t *h_tail = nullptr;
			if (!h_tail)
				h_tail = (uint16_t*)malloc((size_t)4096 * s.tail_cap * 2);
			CK(cudaMemcpy(h_tail, s.tail, (size_t)4096 * s.tail_cap * 2, cudaMemcpyDeviceToHost));
			for (int p = 0; p < 4096; p++) {
				keys_found += cur[p * RJ_CUR_STRIDE];
				const uint32_t nt = cur[p * RJ_CUR_STRIDE + 1] < s.tail_cap ? cur[p * RJ_CUR_STRIDE + 1] : s.tail_cap;
				for (uint32_t e = 0; e + 16 <= nt; e += 16) {
					const uint32_t c = h_tail[(size_t)p * s.tail_cap + e + 15];
					keys_found += c < 15 ? c : 15;
					tail_sectors++;
				}
			}
		}
		printf("   keys accounted for: %llu of %llu (%lld missing), %llu tail sectors\n", keys_found, (unsigned long long)n,
				(long long)n - (long long)keys_found, tail_sectors);
		printf("%s hints %d: %8.3f ms  %7.1f GB/s of keys   (error flags %u, entries in streams %llu of %llu)\n",
				variant < 8 ? "fast   " : "generic", variant < 8 ? variant : 0, total / reps, 8.0 * n / (total / reps) / 1e6, h[0], total_entries,
				(unsigned long long)n);
	}
#if RJ_LAB & 8
	unsigned long long tl[4][8];
	CK(cudaMemcpyFromSymbol(tl, rj_timeline, sizeof(tl)));
	const char *names[5] = {"insert", "wait barrier 1", "flush rows (atomic, store)", "wait barrier 2", "spill + scan counters"};
	const int order[5] = {0, 1, 4, 2, 3};
	for (int w = 0; w < 4; w++) {
		unsigned long long sum = 0;
		for (int i = 0; i < 5; i++)
			sum += tl[w][i];
		printf("warp %2d of CTA 0:", w * (RJ_P1_THREADS / 96));
		for (int k = 0; k < 5; k++)
			printf("  %s %4.1f%%", names[order[k]], 100.0 * tl[w][order[k]] / (double)sum);
		printf("   (cycles per launch %.0f)\n", sum / 7.0);
	}
#endif
	return 0;
}
