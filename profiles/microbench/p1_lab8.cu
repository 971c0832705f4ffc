// Pass-1 laboratory, part 8 (round 2): p1_lab8 plus RESERVED stream positions (VAR 25/26): every CTA holds one reserved
// sector position per partition in shared memory, the flush stores without waiting for a global atomic and reserves the
// partition's next position for later (the atomic's result is parked in a register until the next round).
// Derived from part 4: per-partition global streams.  p1_lab2.cu showed that scattered 32-byte sector stores
// cost 0.13 ms more than stores that complete whole 128-byte lines; here every partition has ONE append-only
// stream shared by all CTAs (position = global atomic on the partition's cursor), so consecutive sectors of a line
// are written by different CTAs within about a microsecond and merge in L2 before they reach DRAM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o p1_lab8 p1_lab8.cu && ./p1_lab8 [log2_rows]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define NPART 4096
#ifndef CAP
#define CAP 20
#endif
#define THREADS 1024
#define NWARP (THREADS / 32)
#define WLCAP 64
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
__device__ __forceinline__ void stg256_keep(void *p, uint2 r0, uint2 r1, uint2 r2, uint2 r3)
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;" ::"l"(p), "r"(r0.x), "r"(r0.y), "r"(r1.x), "r"(r1.y),
			"r"(r2.x), "r"(r2.y), "r"(r3.x), "r"(r3.y), "l"(pol) : "memory");
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
	uint16_t wl[2][NWARP][WLCAP];
	uint16_t base[NPART]; // VAR 25/26: reserved position (in sectors) of the partition's next flush, 0xffff = none
#ifdef PAD
	unsigned char pad[PAD];
#endif
};

// VAR 20  cursor atomic issued by the lane that completes a row (insert phase), result parked in a register, handed to the
//         flushing lane through shared memory after the barrier
// VAR 21  same, but the completing lane flushes its own rows (no worklist; divergent)
// VAR 22  cursor atomic issued in the flush phase (its latency is exposed)
// VAR 23  no atomics: private per-warp sequential sectors (= p1_lab2 VAR 17, lower bound)
template <int VAR>
__global__ void __launch_bounds__(THREADS, 1) k_p1(const int64_t *keys, uint64_t n, int shift, uint16_t *streams, uint32_t cap,
		uint32_t *cursor, uint32_t *sink)
{
	extern __shared__ __align__(16) unsigned char raw[];
	Smem *sm = reinterpret_cast<Smem*>(raw);
	constexpr int NK = 8;
	constexpr int TILE = THREADS * NK;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t lt = (1u << lane) - 1u;
	for (int p = tid; p < NPART; p += THREADS)
		sm->fill[p] = 0;
	if (VAR == 25 || VAR == 26)
		for (int p = tid; p < NPART; p += THREADS)
			sm->base[p] = (uint16_t)(atomicAdd(&cursor[p], 16u) >> 4);
	__syncthreads();
	const uint64_t nfull = n / TILE;
	const uint32_t mask = (1u << shift) - 1u;
	uint32_t warp_sec = 0;
	int par = 0;
	uint32_t held_pos = 0, held_cnt = 0; // VAR 24: cursor positions requested one round ago
	uint32_t held_p[2] = {0xffffffffu, 0xffffffffu}, held_at[2] = {0, 0}; // VAR 25/26
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
	auto flush_row = [&](uint32_t p, uint32_t at) { // at = entry index inside the partition's stream
		const uint32_t f = min(sm->fill[p], (uint32_t)CAP);
		uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * CAP]);
		const uint2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
		if (at + 16 <= cap) {
			if (VAR == 26)
				stg256_keep(streams + (size_t)p * cap + at, r0, r1, r2, r3);
			else
				stg256(streams + (size_t)p * cap + at, r0, r1, r2, r3);
		}
		row[0] = r4;
		sm->fill[p] = f - 16;
	};
	auto round = [&](const uint32_t *d) {
		uint32_t pos[NK], wpos[NK];
#pragma unroll
		for (int k = 0; k < NK; k++)
			pos[k] = smem_inc(&sm->fill[d[k] >> shift]);
		uint32_t cnt = 0;
#pragma unroll
		for (int k = 0; k < NK; k++) {
			const uint32_t p = d[k] >> shift;
			if (pos[k] < CAP)
				sm->stage[p * CAP + pos[k]] = (uint16_t)(d[k] & mask);
			else
				acc++;
			const bool q = pos[k] == 15;
			wpos[k] = 0;
			if (q && (VAR == 20 || VAR == 21))
				wpos[k] = atomicAdd(&cursor[p], 16u);
			if (VAR != 21) {
				const uint32_t bal = __ballot_sync(0xffffffffu, q);
				if (q)
					sm->wl[par][warp][cnt + __popc(bal & lt)] = (uint16_t)p;
				cnt += __popc(bal);
			}
		}
		__syncthreads();
		if (VAR == 21) {
#pragma unroll
			for (int k = 0; k < NK; k++)
				if (pos[k] == 15)
					flush_row(d[k] >> shift, wpos[k]);
		} else if (VAR == 24) {
			// lazy flush: rows completed LAST round go out now (their positions were requested a round ago),
			// rows completed this round only request their position
			if ((uint32_t)lane < held_cnt) {
				const uint32_t p = sm->wl[par ^ 1][warp][lane];
				const uint32_t f = min(sm->fill[p], (uint32_t)CAP);
				uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * CAP]);
				const uint2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
				if (held_pos + 16 <= cap)
					stg256(streams + (size_t)p * cap + held_pos, r0, r1, r2, r3);
				row[0] = r4;
				sm->fill[p] = f - 16;
			}
			held_cnt = min(cnt, 32u);
			if ((uint32_t)lane < held_cnt)
				held_pos = atomicAdd(&cursor[sm->wl[par][warp][lane]], 16u);
			for (uint32_t w = 32 + lane; w < cnt; w += 32) { // more than 32 rows in one warp and round: eager
				const uint32_t p = sm->wl[par][warp][w];
				flush_row(p, atomicAdd(&cursor[p], 16u));
			}
		} else if (VAR == 25 || VAR == 26) {
			// retire the reservations requested in the previous round (their atomics have long returned)
#pragma unroll
			for (int j = 0; j < 2; j++)
				if (held_p[j] != 0xffffffffu)
					sm->base[held_p[j]] = (uint16_t)(held_at[j] >> 4);
			__syncwarp();
#pragma unroll
			for (int j = 0; j < 2; j++) {
				const uint32_t w = lane + 32 * j;
				held_p[j] = 0xffffffffu;
				if (w < cnt) {
					const uint32_t p = sm->wl[par][warp][w];
					const uint32_t b = sm->base[p];
					const uint32_t at = b == 0xffffu ? atomicAdd(&cursor[p], 16u) : b * 16u;
					sm->base[p] = 0xffffu;
					flush_row(p, at);
					held_at[j] = atomicAdd(&cursor[p], 16u); // next reservation: nobody waits for it here
					held_p[j] = p;
				}
			}
		} else if (VAR == 22) {
			for (uint32_t w = lane; w < cnt; w += 32) {
				const uint32_t p = sm->wl[par][warp][w];
				flush_row(p, atomicAdd(&cursor[p], 16u));
			}
		} else {
			for (uint32_t w = lane; w < cnt; w += 32) {
				// private sequential sectors: warp-major layout inside the CTA's slice of the stream area
				const uint32_t p = sm->wl[par][warp][w];
				const uint32_t f = min(sm->fill[p], (uint32_t)CAP);
				uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * CAP]);
				const uint2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
				stg256(streams + ((size_t)(blockIdx.x * NWARP + warp) * 8192 + ((warp_sec + w) & 8191u)) * 16, r0, r1, r2, r3);
				row[0] = r4;
				sm->fill[p] = f - 16;
			}
			warp_sec += cnt;
		}
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
	for (int p = tid; p < NPART; p += THREADS)
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
		k_p1<VAR><<<sms, THREADS, sizeof(Smem)>>>(keys, n, shift, streams, cap, cursor, sink);
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
	const uint32_t cap = (uint32_t)(2 * n / NPART) + 148 * 32;
	CK(cudaMalloc(&keys, n * 8));
	CK(cudaMalloc(&streams, (size_t)NPART * cap * 2 + (size_t)sms * NWARP * 8192 * 32));
	CK(cudaMalloc(&sink, 4));
	CK(cudaMalloc(&cursor, NPART * 4));
	k_gen<<<sms * 8, 256>>>(keys, n, n);
	CK(cudaDeviceSynchronize());
	const int shift = lg - 12;
	printf("n = 2^%d keys, %d SMs, shift %d, stream capacity %u entries\n", lg, sms, shift, cap);
	run<23>("23 private sequential sectors per warp (no atomics; lower bound)", keys, n, shift, streams, cap, cursor, sink, sms);
	run<22>("22 per-partition streams, cursor atomic inside the flush", keys, n, shift, streams, cap, cursor, sink, sms);
	run<25>("25 per-partition streams, reserved positions (no atomic on the flush path)", keys, n, shift, streams, cap, cursor, sink, sms);
	run<26>("26 = 25 with evict-last stream stores", keys, n, shift, streams, cap, cursor, sink, sms);
	return 0;
}
