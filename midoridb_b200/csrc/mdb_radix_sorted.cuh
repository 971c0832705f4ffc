// mdb_radix_sorted.cuh - a join side whose key column is already sorted (auto-increment ids, the reference README's own
// example) needs no pass 1: the keys of partition p are the contiguous rows [bnd[p], bnd[p+1]) of the column, and pass 2
// counts them straight from there.  (It also could not go through pass 1: a sorted tile hits one or two staging rows.)
// Included by mdb_radix.cu.
#pragma once

// any descent keys[i] > keys[i+1] in [0, n)?  Four keys per 256-bit load plus the first key of the next quad.
__global__ void k_is_sorted(const int64_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ descents)
{
	const uint64_t nquads = n / 4;
	bool bad = false;
	for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < nquads; q += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t w[8];
		asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				: "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(keys + q * 4));
		long long k[5];
#pragma unroll
		for (int j = 0; j < 4; j++)
			k[j] = (long long)(((unsigned long long)w[2 * j + 1] << 32) | w[2 * j]);
		k[4] = q * 4 + 4 < n ? keys[q * 4 + 4] : k[3];
		bad |= k[0] > k[1] || k[1] > k[2] || k[2] > k[3] || k[3] > k[4];
	}
	// (rows beyond the last whole quad were compared as k[4] of the last quad or are the ragged tail:)
	if (blockIdx.x == 0 && threadIdx.x == 0)
		for (uint64_t i = nquads * 4; i + 1 < n; i++)
			bad |= keys[i] > keys[i + 1];
	if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0)
		atomicOr(descents, 1u);
}

// bnd[p] = first row whose key is >= kmin + p * width, p = 0 .. nparts (bnd[nparts]: first row beyond the range)
__global__ void k_sorted_bounds(const int64_t *__restrict__ keys, uint64_t n, RJParams pr, uint64_t *__restrict__ bnd)
{
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p > pr.nparts)
		return;
	const unsigned long long off = p == pr.nparts ? pr.range : (unsigned long long)p * pr.width;
	uint64_t lo = 0, hi = n;
	while (lo < hi) {
		const uint64_t mid = lo + (hi - lo) / 2;
		// compare in the unsigned "distance from kmin" domain, with keys below kmin sorting first
		const long long k = keys[mid];
		const bool below = k < pr.kmin || (unsigned long long)k - (unsigned long long)pr.kmin < off;
		if (below)
			lo = mid + 1;
		else
			hi = mid;
	}
	bnd[p] = lo;
}
