// mdb_radix_dist.cuh - multi-GPU exchange of the radix join+count (included by mdb_radix.cu).
//
// One process (or one context) per GPU.  Every rank partitions ITS shard of both join sides with pass 1 over the same
// global key range (mdbcu_table_sync_stats), so partition p means the same key interval everywhere.  Rank r owns the
// contiguous partition block [r*P/W, (r+1)*P/W): key ranges are disjoint, so every rank runs pass 2 on its own
// partitions and the per-rank results simply concatenate - no merge step.
//
// The exchange is a PULL over NVLink / NVSwitch: pass 1 writes its streams into the rank's own arena (one cudaMalloc
// block, mapped into all peers: mdb_comm.cu); after one cross-rank barrier (a flag word per source rank in the arena
// header, k_arena_barrier) the pass 2 of the rank that owns partition p reads p's runs from all W arenas with the same
// 256-bit loads it uses locally, all runs of a partition in one index space so that the remote latency is paid once per
// partition, not once per run (rj_histogram_multi).  What crosses the link is the 2-byte remainders, never the 8-byte keys.
//
// History, all measured on B200s (profiles/): (1) pass 1 storing every flushed 32-byte sector directly into the owner's
// memory ran at half speed (remote 32-byte stores are bound by the number of stores in flight); (2) a staged NCCL grouped
// send/recv of gathered chunks took 2-3 ms; (3) a push kernel after pass 1 (one warp per stream, 1 KiB of contiguous
// remote memory per store instruction) moved the bytes at 450-650 GB/s but stayed exposed on the critical path - 0.135 ms
// of a 0.70 ms step at 8 GPUs - and cost pass 1 an eighth of the SMs while it overlapped (profiles/r02_scale_push_design_n*.json);
// (4) everything pulled by pass 2 itself: no copy kernel at all, but pass 2 becomes NVLink-bound (0.57 ms for 281 MB at 2 GPUs)
// and nothing overlaps pass 1 any more (profiles/r02_pull_n2.json); (5) side A FETCHED by a copy kernel on 16 SMs while pass 1
// of side B runs: remote loads wait for their round trip, 155 GB/s, the step got slower (profiles/r02_hybrid_fetch_n2.json);
// (6) this version: side A is PUSHED into the owners' slots by k_radix_ship on 16 SMs while pass 1 of side B runs (stores do
// not wait: 650 GB/s), side B is PULLED by pass 2 itself - every byte crosses the link under cover of some computation, and
// one cross-rank barrier per query (after pass 1 of side B) is enough.
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

// byte layout of ONE join side's streams inside a rank's arena half (identical on every rank: stream capacities come from
// the global row counts): main streams of all partitions, tail streams, then the [main, tail] cursor pairs
struct RJRegionLayout {
	size_t main, tail, cursor, bytes;
};

static RJRegionLayout rj_region_layout(uint32_t nparts, uint32_t cap, uint32_t tail_cap)
{
	auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
	RJRegionLayout l;
	l.main = 0;
	l.tail = up(l.main + (size_t)nparts * cap * sizeof(uint16_t));
	l.cursor = up(l.tail + (size_t)nparts * tail_cap * sizeof(uint16_t));
	l.bytes = up(l.cursor + (size_t)RJ_MAX_PART * RJ_CUR_STRIDE * sizeof(uint32_t));
	return l;
}

// Side A's runs of the partitions a rank owns, pushed into its arena by the peers while pass 1 of side B runs: slot o of a
// rank's arena holds rank o's streams of the partitions the arena's owner owns (partition q = p - first), counts included.
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

// where this rank's side-A streams go: for every destination rank the slot [self] of that rank's arena
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

// one WARP per partition owned by a peer (grid-stride over warps: with 8 GPUs a partition's stream is 16 KiB, a CTA per
// partition would be latency-bound): push its main and tail stream and their counts into the owner's slot - 256-bit stores,
// a warp writes 1 KiB of contiguous remote memory per instruction, and stores (unlike loads) do not wait for the round trip:
// 16 SMs move 650 GB/s this way where the same kernel READING from the peers reached 155 GB/s (profiles/r02_hybrid_fetch_n2.json).
// Launched with enough dynamic shared memory to own its SMs, next to pass 1 of side B on the others.
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
