// Pass-1 laboratory: where does the time of k_radix_partition go?  Decomposes the kernel into its ingredients on
// the bench workload (2^28 uniform int64 keys, 4096 partitions, 16-bit remainders, one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o p1_lab p1_lab.cu && ./p1_lab [log2_rows]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define NPART 4096
#define CAP 20
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ int4 ldg_stream(const int4 *p)
{
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
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
	uint16_t stage[NPART * CAP];
	uint32_t fill[NPART];
	uint32_t chunk[NPART];
	uint16_t worklist[2][NPART];
	uint32_t wl_count[2];
};

// MODE 0 loads only; 1 + slot atomics + 2-byte stores; 2 = 1 + two barriers per round; 3 = 1 + one barrier per round;
// 4 = 2 + worklist + flush of queued rows (sector writes to a private scratch region, simplified chunk bookkeeping)
// 5 = 4 but the flush of round r runs after the inserts of round r+1 were issued (one barrier per round)
template <int MODE, int THREADS, int LOADS>
__global__ void __launch_bounds__(THREADS, 1) k_p1(const int64_t *keys, uint64_t n, int shift, uint16_t *scratch, size_t scratch_per_cta,
		uint32_t *sink)
{
	extern __shared__ __align__(16) unsigned char raw[];
	Smem *sm = reinterpret_cast<Smem*>(raw);
	constexpr int TILE = THREADS * LOADS * 2;
	constexpr int NK = LOADS * 2;
	const int tid = threadIdx.x;
	for (int p = tid; p < NPART; p += THREADS) {
		sm->fill[p] = 0;
		sm->chunk[p] = 0;
	}
	if (tid < 2)
		sm->wl_count[tid] = 0;
	__syncthreads();
	const uint64_t nfull = n / TILE;
	const int4 *src = reinterpret_cast<const int4*>(keys);
	const uint32_t mask = (1u << shift) - 1u;
	uint16_t *my = scratch + (size_t)blockIdx.x * scratch_per_cta;
	uint32_t wr = 0; // sectors written by this CTA (MODE 4)
	__shared__ uint32_t wr_sh;
	if (tid == 0)
		wr_sh = 0;
	uint32_t acc = 0;
	int4 a[LOADS], b[LOADS];
	int par = 0;
	auto load = [&](uint64_t tile, int4 *dst) {
		const int4 *t = src + tile * (TILE / 2) + tid;
#pragma unroll
		for (int j = 0; j < LOADS; j++)
			dst[j] = ldg_stream(t + j * THREADS);
	};
	auto flush = [&](int fp) {
		const uint32_t nwl = sm->wl_count[fp];
		for (uint32_t w = tid; w < nwl; w += THREADS) {
			const uint32_t p = sm->worklist[fp][w];
			const uint32_t f = min(sm->fill[p], (uint32_t)CAP);
			uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * CAP]);
			const uint2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
			const uint32_t sec = smem_inc(&wr_sh);
			int4 *dst = reinterpret_cast<int4*>(my + ((size_t)(sec * 2654435761u) % (scratch_per_cta / 16)) * 16); // scattered sectors
			dst[0] = make_int4((int)r0.x, (int)r0.y, (int)r1.x, (int)r1.y);
			dst[1] = make_int4((int)r2.x, (int)r2.y, (int)r3.x, (int)r3.y);
			row[0] = r4;
			sm->fill[p] = f - 16;
		}
	};
	auto round = [&](const int4 *buf) {
		if (MODE == 0) {
#pragma unroll
			for (int j = 0; j < LOADS; j++)
				acc ^= buf[j].x ^ buf[j].z;
			return;
		}
		uint32_t d[NK], pos[NK];
#pragma unroll
		for (int j = 0; j < LOADS; j++) {
			d[2 * j] = (uint32_t)buf[j].x;
			d[2 * j + 1] = (uint32_t)buf[j].z;
		}
#pragma unroll
		for (int k = 0; k < NK; k++)
			pos[k] = smem_inc(&sm->fill[d[k] >> shift]);
		if (MODE <= 3) {
#pragma unroll
			for (int k = 0; k < NK; k++)
				sm->stage[(d[k] >> shift) * CAP + (pos[k] & 15u)] = (uint16_t)(d[k] & mask);
		} else {
			uint32_t widx[NK];
#pragma unroll
			for (int k = 0; k < NK; k++) {
				widx[k] = 0;
				if (pos[k] == 15)
					widx[k] = smem_inc(&sm->wl_count[par]);
			}
#pragma unroll
			for (int k = 0; k < NK; k++) {
				const uint32_t p = d[k] >> shift;
				if (pos[k] < CAP)
					sm->stage[p * CAP + pos[k]] = (uint16_t)(d[k] & mask);
				else
					acc++; // parked (dropped here)
				if (pos[k] == 15)
					sm->worklist[par][widx[k]] = (uint16_t)p;
			}
		}
		if (MODE == 2 || MODE == 4)
			__syncthreads();
		if (MODE == 4) {
			if (tid == 0)
				sm->wl_count[par ^ 1] = 0;
			flush(par);
		}
		if (MODE == 2 || MODE == 3 || MODE == 4)
			__syncthreads();
		par ^= 1;
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
	if (MODE >= 1) {
		for (int p = tid; p < NPART; p += THREADS)
			acc += sm->fill[p] + sm->stage[p * CAP];
	}
	acc += wr;
	if (acc == 0x12345678u)
		sink[0] = acc;
}

template <int MODE, int THREADS, int LOADS>
static void run(const char *name, const int64_t *keys, uint64_t n, int shift, uint16_t *scratch, size_t per_cta, uint32_t *sink, int sms)
{
	CK(cudaFuncSetAttribute(k_p1<MODE, THREADS, LOADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	for (int i = 0; i < 2; i++)
		k_p1<MODE, THREADS, LOADS><<<sms, THREADS, sizeof(Smem)>>>(keys, n, shift, scratch, per_cta, sink);
	CK(cudaDeviceSynchronize());
	const int reps = 5;
	CK(cudaEventRecord(e0));
	for (int i = 0; i < reps; i++)
		k_p1<MODE, THREADS, LOADS><<<sms, THREADS, sizeof(Smem)>>>(keys, n, shift, scratch, per_cta, sink);
	CK(cudaEventRecord(e1));
	CK(cudaDeviceSynchronize());
	float ms;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	ms /= reps;
	printf("%-64s %8.3f ms  %7.1f GB/s of keys\n", name, ms, 8.0 * n / ms / 1e6);
}

int main(int argc, char **argv)
{
	const int lg = argc > 1 ? atoi(argv[1]) : 28;
	const uint64_t n = 1ull << lg;
	int sms;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	int64_t *keys;
	uint16_t *scratch;
	uint32_t *sink;
	CK(cudaMalloc(&keys, n * 8));
	const size_t per_cta = (n / sms + 65536) & ~(size_t)15;
	CK(cudaMalloc(&scratch, per_cta * sms * 2));
	CK(cudaMalloc(&sink, 4));
	k_gen<<<sms * 8, 256>>>(keys, n, n);
	CK(cudaDeviceSynchronize());
	const int shift = lg - 12;
	printf("n = 2^%d keys, %d SMs, shift %d\n", lg, sms, shift);
	run<0, 1024, 4>("0 loads only                              1024 thr x 8 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<1, 1024, 4>("1 + slot atomics + 2-byte stores          1024 thr x 8 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<3, 1024, 4>("3 + one barrier per round                 1024 thr x 8 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<2, 1024, 4>("2 + two barriers per round                1024 thr x 8 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<4, 1024, 4>("4 + worklist + flush (sector writes)      1024 thr x 8 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<0, 512, 8>("0 loads only                               512 thr x 16 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<1, 512, 8>("1 + slot atomics + 2-byte stores           512 thr x 16 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<2, 512, 8>("2 + two barriers per round                 512 thr x 16 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<4, 512, 8>("4 + worklist + flush (sector writes)       512 thr x 16 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<1, 1024, 2>("1 + slot atomics + 2-byte stores          1024 thr x 4 keys", keys, n, shift, scratch, per_cta, sink, sms);
	run<4, 1024, 2>("4 + worklist + flush (sector writes)      1024 thr x 4 keys", keys, n, shift, scratch, per_cta, sink, sms);
	return 0;
}
