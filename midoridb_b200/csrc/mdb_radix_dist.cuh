// mdb_radix_dist.cuh - multi-GPU exchange of the radix join+count (included by mdb_radix.cu).
//
// One process per GPU.  Every rank partitions ITS shard of both join sides with pass 1 over the same global key
// range (mdbcu_table_sync_stats), so partition p means the same key interval everywhere.  Rank r owns the
// contiguous partition block [r*P/W, (r+1)*P/W): key ranges are disjoint, so after the exchange every rank
// runs pass 2 on its own partitions and the per-rank results simply concatenate - no merge step.
//
// The exchange is a PUSH over NVLink / NVSwitch: every rank's arena (one cudaMalloc block, mapped into all peers
// with CUDA IPC, mdb_comm.cu) has one slot per source rank; k_radix_ship copies the streams of the partitions a
// peer owns straight into that peer's slot with 256-bit stores (a warp writes 1 KiB of contiguous remote memory per
// instruction) together with their entry counts.  What crosses the link is the 2-byte remainders, never the 8-byte keys.
// The cross-rank barrier (pushes landed + every rank's error flags) is a flag word per source rank in the arena header
// (k_arena_barrier, mdb_comm.cu); alternate queries use alternate halves of the arena, so one barrier per exchange suffices.
//
// Measured on 2 B200s (profiles/): writing every flushed 32-byte sector directly into the owner's memory from inside
// pass 1 (the first version of this exchange) ran pass 1 at half speed - remote 32-byte stores are bound by the
// number of stores in flight, not by the link - and a staged NCCL send/recv of gathered chunks took 2-3 ms;
// the bulk push below moves the same bytes at link speed after a pass 1 that runs at its single-GPU speed.
#pragma once

// Partition ownership, the one piece of arithmetic every rank (host and device) must agree on: rank r owns the contiguous
// block [r * nparts / W, (r + 1) * nparts / W).  Exported for the CPU tests as mdbcu_dist_describe (mdb_radix.cu).
__host__ __device__ static inline uint32_t rj_part_first(uint32_t rank, uint32_t nparts, uint32_t world)
{
	return (uint32_t)((uint64_t)rank * nparts / world);
}
__host__ __device__ static inline uint32_t rj_owner_of(uint32_t p, uint32_t nparts, uint32_t world)
{
	return (uint32_t)((((uint64_t)p + 1) * world - 1) / nparts);
}

// byte layout of one side inside an arena slot (identical on every rank)
struct RJSlotLayout {
	size_t main, tail, cursor, tail_cursor, bytes;
};

static RJSlotLayout rj_slot_layout(uint32_t pown, uint32_t cap, uint32_t tail_cap)
{
	auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
	RJSlotLayout l;
	l.main = 0;
	l.tail = up(l.main + (size_t)pown * cap * sizeof(uint16_t));
	l.cursor = up(l.tail + (size_t)pown * tail_cap * sizeof(uint16_t));
	l.tail_cursor = up(l.cursor + (size_t)pown * sizeof(uint32_t));
	l.bytes = up(l.tail_cursor + (size_t)pown * sizeof(uint32_t));
	return l;
}

// where this rank's streams go: for every destination rank the slot [self] of that rank's arena
struct RJShip {
	int world, self, nparts;
	uint16_t *main[RJ_MAX_RANKS];
	uint16_t *tail[RJ_MAX_RANKS];
	uint32_t *cursor[RJ_MAX_RANKS];
	uint32_t *tail_cursor[RJ_MAX_RANKS];
	unsigned long long *shipped_bytes; // statistics
};

// copy nvec 32-byte vectors with one warp, four loads in flight per lane before the four stores
__device__ __forceinline__ void rj_copy_vectors(char *dst, const char *src, uint32_t nvec)
{
	constexpr int U = 4;
	const uint32_t lane = threadIdx.x & 31u;
	for (uint32_t v0 = lane; v0 < nvec; v0 += 32 * U) {
		uint32_t w[U][8];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t v = min(v0 + u * 32u, nvec - 1u); // clamped: keeps the vectors in registers
			asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
					: "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7])
					: "l"(src + (size_t)v * 32u));
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t v = v0 + u * 32u;
			if (v < nvec)
				asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + (size_t)v * 32u), "r"(w[u][0]), "r"(w[u][1]),
						"r"(w[u][2]), "r"(w[u][3]), "r"(w[u][4]), "r"(w[u][5]), "r"(w[u][6]), "r"(w[u][7]) : "memory");
		}
	}
}

// one WARP per partition owned by a peer (grid-stride over warps: with 8 GPUs a partition's stream is 16 KiB, a CTA
// per partition would be latency-bound): push its main and tail stream and their counts.  Launched either over the
// whole GPU or - with enough dynamic shared memory to own an SM - on a handful of SMs next to pass 1 of the other side.
__global__ void __launch_bounds__(1024) k_radix_ship(RJSide s, RJShip sh)
{
	const uint32_t warps = blockDim.x >> 5, gw = blockIdx.x * warps + (threadIdx.x >> 5), nw = gridDim.x * warps;
	for (uint32_t p = gw; p < (uint32_t)sh.nparts; p += nw) {
		const int o = (int)rj_owner_of(p, (uint32_t)sh.nparts, (uint32_t)sh.world);
		if (o == sh.self)
			continue;
		const uint32_t q = p - rj_part_first((uint32_t)o, (uint32_t)sh.nparts, (uint32_t)sh.world);
		const uint32_t n_main = min(s.cursor[p * RJ_CUR_STRIDE], s.cap), n_tail = min(s.tail_cursor[p * RJ_CUR_STRIDE], s.tail_cap);
		rj_copy_vectors(reinterpret_cast<char*>(sh.main[o] + (size_t)q * s.cap), reinterpret_cast<const char*>(s.stream + (size_t)p * s.cap),
				n_main / 16u);
		rj_copy_vectors(reinterpret_cast<char*>(sh.tail[o] + (size_t)q * s.tail_cap),
				reinterpret_cast<const char*>(s.tail + (size_t)p * s.tail_cap), (n_tail + 15u) / 16u);
		if ((threadIdx.x & 31u) == 0) {
			sh.cursor[o][q] = n_main;
			sh.tail_cursor[o][q] = n_tail;
			atomicAdd(sh.shipped_bytes, 2ull * (n_main + n_tail) + 8ull);
		}
	}
	__threadfence_system(); // the owners read these slots after the next cross-rank barrier
}
