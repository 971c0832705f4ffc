// Pass-1 laboratory, part 2: flush variants (p1_lab.cu showed that queueing + flushing doubles the kernel time).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o p1_lab2 p1_lab2.cu && ./p1_lab2 [log2_rows]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define NPART 4096
#define CAP 20
#define THREADS 1024
#define NWARP (THREADS / 32)
#define WLCAP 96
#define SEC_LOG2 17
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ int4 ldg_stream(const int4 *p)
{
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
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
	uint16_t stage[NPART * CAP];
	uint32_t fill[NPART];
	uint32_t chunk[NPART];
	uint16_t wl[NWARP][WLCAP];   // per-warp worklists (VAR >= 11); VAR 10 uses it as one flat list
	uint32_t wl_count[2];
	uint32_t wr_sh;
};

// VAR 10  one CTA-wide worklist (same-address atomics), sector index from a same-address atomic, 2 x STG.128   [= today's kernel]
// VAR 11  per-warp worklists built with ballots (no atomics), per-warp sector counters, 2 x STG.128
// VAR 12  VAR 11 with one STG.256 per sector
// VAR 13  VAR 12 + 256-bit key loads
// VAR 14  loads only, 256-bit
// VAR 15  VAR 12 without the barrier after the flush (timing only: unsafe as written)
// VAR 16  VAR 12, flush without the global store (shared-memory part only)
template <int VAR>
__global__ void __launch_bounds__(THREADS, 1) k_p1(const int64_t *keys, uint64_t n, int shift, uint16_t *scratch, uint32_t *sink)
{
	extern __shared__ __align__(16) unsigned char raw[];
	Smem *sm = reinterpret_cast<Smem*>(raw);
	constexpr int NK = 8;
	constexpr int TILE = THREADS * NK;
	constexpr bool L256 = VAR == 13 || VAR == 14 || VAR >= 17;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int p = tid; p < NPART; p += THREADS) {
		sm->fill[p] = 0;
		sm->chunk[p] = 0;
	}
	if (tid < 2)
		sm->wl_count[tid] = 0;
	if (tid == 0)
		sm->wr_sh = 0;
	__syncthreads();
	const uint64_t nfull = n / TILE;
	const uint32_t mask = (1u << shift) - 1u;
	uint16_t *my = scratch + ((size_t)blockIdx.x << SEC_LOG2) * 16;
	uint32_t warp_sec = warp << (SEC_LOG2 - 5); // each warp owns 1/32 of the CTA's sectors
	uint32_t acc = 0;
	uint32_t a[NK], b[NK]; // low words of 8 keys
	int par = 0;
	auto load = [&](uint64_t tile, uint32_t *dst) {
		if (L256) {
			uint32_t t[8];
			const char *base = reinterpret_cast<const char*>(keys + tile * TILE);
			ldg_stream256(base + (size_t)tid * 32, t);
			dst[0] = t[0]; dst[1] = t[2]; dst[2] = t[4]; dst[3] = t[6];
			ldg_stream256(base + (size_t)(THREADS + tid) * 32, t);
			dst[4] = t[0]; dst[5] = t[2]; dst[6] = t[4]; dst[7] = t[6];
		} else {
			const int4 *src = reinterpret_cast<const int4*>(keys) + tile * (TILE / 2) + tid;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int4 v = ldg_stream(src + j * THREADS);
				dst[2 * j] = (uint32_t)v.x;
				dst[2 * j + 1] = (uint32_t)v.z;
			}
		}
	};
	auto flush_row = [&](uint32_t p, uint32_t sec) {
		const uint32_t f = min(sm->fill[p], (uint32_t)CAP);
		uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * CAP]);
		const uint2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
	uint16_t *dst = my + (size_t)(VAR == 17 ? sec : VAR == 18 ? ((sec & ~3u) * 40503u & ((1u << SEC_LOG2) - 1u) & ~3u) | (sec & 3u) : (sec * 40503u) & ((1u << SEC_LOG2) - 1u)) * 16; // 17 sequential per warp; 18 scattered 128-byte lines; else scattered sectors
		if (VAR == 10 || VAR == 11) {
			reinterpret_cast<int4*>(dst)[0] = make_int4((int)r0.x, (int)r0.y, (int)r1.x, (int)r1.y);
			reinterpret_cast<int4*>(dst)[1] = make_int4((int)r2.x, (int)r2.y, (int)r3.x, (int)r3.y);
		} else if (VAR != 16) {
			stg256(dst, r0, r1, r2, r3);
		} else {
			acc += r0.x + r1.y + r2.x + r3.y;
		}
		row[0] = r4;
		sm->fill[p] = f - 16;
	};
	auto round = [&](const uint32_t *d) {
		if (VAR == 14) {
#pragma unroll
			for (int k = 0; k < NK; k++)
				acc ^= d[k];
			return;
		}
		uint32_t pos[NK];
#pragma unroll
		for (int k = 0; k < NK; k++)
			pos[k] = smem_inc(&sm->fill[d[k] >> shift]);
		if (VAR == 10) {
			uint32_t widx[NK];
#pragma unroll
			for (int k = 0; k < NK; k++) {
				widx[k] = 0;
				if (pos[k] == 15)
					widx[k] = smem_inc(&sm->wl_count[par]);
			}
			uint16_t *wl = &sm->wl[0][0] + par * (NWARP * WLCAP / 2);
#pragma unroll
			for (int k = 0; k < NK; k++) {
				const uint32_t p = d[k] >> shift;
				if (pos[k] < CAP)
					sm->stage[p * CAP + pos[k]] = (uint16_t)(d[k] & mask);
				else
					acc++;
				if (pos[k] == 15)
					wl[widx[k]] = (uint16_t)p;
			}
			__syncthreads();
			if (tid == 0)
				sm->wl_count[par ^ 1] = 0;
			const uint32_t nwl = sm->wl_count[par];
			for (uint32_t w = tid; w < nwl; w += THREADS)
				flush_row(wl[w], smem_inc(&sm->wr_sh));
			__syncthreads();
		} else {
			uint32_t cnt = 0;
#pragma unroll
			for (int k = 0; k < NK; k++) {
				const uint32_t p = d[k] >> shift;
				if (pos[k] < CAP)
					sm->stage[p * CAP + pos[k]] = (uint16_t)(d[k] & mask);
				else
					acc++;
				const bool q = pos[k] == 15;
				const uint32_t bal = __ballot_sync(0xffffffffu, q);
				if (q)
					sm->wl[warp][cnt + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)p;
				cnt += __popc(bal);
			}
			__syncthreads();
			for (uint32_t w = lane; w < cnt; w += 32)
				flush_row(sm->wl[warp][w], warp_sec + w);
			warp_sec += cnt;
			if (VAR != 15)
				__syncthreads();
		}
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
	for (int p = tid; p < NPART; p += THREADS)
		acc += sm->fill[p] + sm->stage[p * CAP];
	if (acc == 0x12345678u)
		sink[0] = acc;
}

template <int VAR>
static void run(const char *name, const int64_t *keys, uint64_t n, int shift, uint16_t *scratch, uint32_t *sink, int sms)
{
	CK(cudaFuncSetAttribute(k_p1<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	for (int i = 0; i < 2; i++)
		k_p1<VAR><<<sms, THREADS, sizeof(Smem)>>>(keys, n, shift, scratch, sink);
	CK(cudaDeviceSynchronize());
	const int reps = 5;
	CK(cudaEventRecord(e0));
	for (int i = 0; i < reps; i++)
		k_p1<VAR><<<sms, THREADS, sizeof(Smem)>>>(keys, n, shift, scratch, sink);
	CK(cudaEventRecord(e1));
	CK(cudaDeviceSynchronize());
	float ms;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	ms /= reps;
	printf("%-72s %8.3f ms  %7.1f GB/s of keys\n", name, ms, 8.0 * n / ms / 1e6);
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
	CK(cudaMalloc(&scratch, ((size_t)sms << SEC_LOG2) * 32));
	CK(cudaMalloc(&sink, 4));
	k_gen<<<sms * 8, 256>>>(keys, n, n);
	CK(cudaDeviceSynchronize());
	const int shift = lg - 12;
	printf("n = 2^%d keys, %d SMs, shift %d\n", lg, sms, shift);
	run<14>("14 loads only, 256-bit", keys, n, shift, scratch, sink, sms);
	run<10>("10 CTA-wide worklist + sector index by same-address atomics, 2xSTG.128", keys, n, shift, scratch, sink, sms);
	run<11>("11 per-warp worklists (ballot), per-warp sector counters, 2xSTG.128", keys, n, shift, scratch, sink, sms);
	run<12>("12 = 11 with STG.256", keys, n, shift, scratch, sink, sms);
	run<13>("13 = 12 with 256-bit key loads", keys, n, shift, scratch, sink, sms);
	run<15>("15 = 12 without the barrier after the flush (timing only)", keys, n, shift, scratch, sink, sms);
	run<16>("16 = 12 without the global store", keys, n, shift, scratch, sink, sms);
	run<17>("17 = 13 with sequential sectors per warp", keys, n, shift, scratch, sink, sms);
	run<18>("18 = 13 with scattered 128-byte lines (4 consecutive sectors)", keys, n, shift, scratch, sink, sms);
	run<19>("19 = 13 again (scattered sectors)", keys, n, shift, scratch, sink, sms);
	return 0;
}
