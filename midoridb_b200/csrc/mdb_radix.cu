// mdb_radix.cu - radix-partitioned join + GROUP BY join key + COUNT(*)   (the README query; BASELINE configs 1 and 3)
//
//   SELECT k, COUNT(*) FROM A INNER JOIN B ON A.k = B.k GROUP BY k
//
// replaces _join_nested_loop_tbl2tbl (src/engine/executor_select.c:1076) + proc_groupby_clause (:1526):
// result(k) = cntA[k] * cntB[k]; no candidate pair and no joined row is ever materialised.
//
//   pass 1  k_radix_partition   streams the 8-byte keys ONCE, writes 2-byte remainders into per-partition
//                               512-byte chunks (partition = high bits of key - kmin, <= 4096 partitions);
//   (dir)   k_radix_dir_*       counting sort of chunk ids by partition (a few microseconds);
//   pass 2  k_radix_joincount   per partition: both sides' remainders -> packed 4- or 8-bit counters in shared
//                               memory, checksum against the number of remainders, multiply, emit groups.
//
// Algorithmic bytes: 8|A| + 8|B| in, 16 G out.  Extra traffic of this design: 2 bytes per key
// written + read back (the remainders).  HBM-bound integer work; no tensor cores (nothing here is a contraction).
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>

#define RJ_MAX_PART 4096           // partitions (12 radix bits)
#define RJ_MAX_SHIFT 16            // remainder bits: 2-byte remainders
#define RJ_CAP 20                  // staging slots per partition in shared memory
#define RJ_FLUSH 16                // a partition is flushed when 16 remainders (= one 32-byte sector) are staged
#define RJ_CHUNK 256               // remainders per chunk (512 bytes = one warp-wide 128-bit load)
#define RJ_BLOCKS_PER_CHUNK (RJ_CHUNK / RJ_FLUSH)
#define RJ_P1_THREADS 1024
#define RJ_P1_LOADS 4              // 128-bit loads (2 keys each) per thread per round
#define RJ_P1_TILE (RJ_P1_THREADS * RJ_P1_LOADS * 2)
#define RJ_OVF_CAP 1536            // keys per round that may find their staging row full and wait one round
#define RJ_NONE 0xffffffffu

#define RJ_ERR_POOL 1u             // chunk pool exhausted
#define RJ_ERR_COUNTER 2u          // a packed counter wrapped (too many equal keys for the counter width)
#define RJ_ERR_SKEW 4u             // more than RJ_OVF_CAP keys per round hit full staging rows

static bool col_all_present(const mdbcu_table *t, int col)
{
	return t->all_live && !t->cols[col].has_nulls;
}

// one run of <= RJ_CHUNK remainders: 16-byte aligned offset into a remainder buffer, valid entries
struct RJDesc {
	uint32_t off16; // in units of 16 bytes (8 remainders)
	uint32_t ne;
};

#define RJ_MAX_RANKS 8

// where the chunks of the partitions owned by one rank are written: this GPU's own arrays, or - in a
// multi-GPU plan - the owner's arena mapped over NVLink (CUDA IPC), so pass 1 IS the exchange
struct RJTarget {
	uint16_t *pool;            // pool_chunks * RJ_CHUNK remainders
	uint32_t *pool_next;       // allocation cursor
	uint16_t *chunk_part;      // partition of each chunk
	uint16_t *chunk_entries;   // valid remainders in each chunk
	uint32_t *dir_cnt;         // chunks per partition
};

struct RJSide {
	const int64_t *keys;
	const uint32_t *present;
	uint64_t n;
	int all_in_range;          // every key of the column lies in [kmin, kmin + range): no per-key range test
	int world, self;           // owner ranks; index of this GPU in dst[]
	uint32_t pool_chunks;      // capacity of every target's pool
	uint32_t id_batch, id_low; // chunk ids a CTA reserves per owner at a time / refill threshold
	RJTarget dst[RJ_MAX_RANKS];
	uint16_t *pool;            // pass 2 reads remainders from here (dst[self].pool unless an NCCL exchange staged them)
	RJDesc *dir;               // (offset, entries) of this GPU's chunks grouped by partition
	uint64_t *dir_off;         // exclusive offsets into dir
	uint32_t *dir_fill;
};

struct RJParams {
	long long kmin;
	unsigned long long range;  // keys in [kmin, kmin + range) can match
	int shift;                 // remainder bits
	uint32_t mask;             // (1 << shift) - 1
	int nparts;
	int part_first, part_end;  // pass 2 handles partitions [part_first, part_end) (all of them on one GPU)
	uint32_t *error_flag;
};

struct RJP1Smem {
	uint16_t stage[RJ_MAX_PART * RJ_CAP];     // 160 KiB: 20 two-byte slots per partition
	uint32_t fill[RJ_MAX_PART];               // slots handed out this round (may overshoot RJ_CAP)
	uint32_t chunk[RJ_MAX_PART];              // current chunk of this CTA: chunk id * 32 + sectors used, or RJ_NONE
	uint16_t worklist[2][RJ_MAX_PART];        // partitions whose 16th slot filled this round
	uint32_t ovf[2][RJ_OVF_CAP];              // (partition << 16 | remainder) waiting for the next round
	uint32_t wl_count[2];
	uint32_t ovf_count[2];
	uint32_t local_next[RJ_MAX_RANKS];        // chunk ids reserved by this CTA in each owner's pool
	uint32_t local_end[RJ_MAX_RANKS];         // (refilled in bulk by thread 0)
};

static_assert(sizeof(RJP1Smem) <= 227 * 1024, "pass-1 shared memory exceeds the 227 KiB a CTA can opt into");

// plain shared-memory atomic: kept in PTX so the compiler does not expand it into warp-aggregation code
__device__ __forceinline__ uint32_t rj_smem_inc(uint32_t *p)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
	return old;
}

__device__ __forceinline__ void rj_global_red_inc(uint32_t *p)
{
	asm volatile("red.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// rank that owns partition p: ranks own the contiguous blocks [r*P/W, (r+1)*P/W)
__device__ __forceinline__ int rj_owner(const RJSide &s, const RJParams &pr, uint32_t p)
{
	return s.world == 1 ? 0 : (int)(((p + 1) * (uint32_t)s.world - 1) / (uint32_t)pr.nparts);
}

// chunk ids come from the CTA's reserved range in the owner's pool (one shared-memory atomic); thread 0 tops a
// range up with ONE global (for a remote owner: NVLink) atomic per id_batch chunks while the other threads
// insert keys, so no flush ever waits on L2 or on the link
__device__ static inline void rj_new_chunk(const RJSide &s, const RJParams &pr, RJP1Smem *sm, uint32_t p)
{
	const int o = rj_owner(s, pr, p);
	const RJTarget &t = s.dst[o];
	const uint32_t old = sm->chunk[p];
	uint32_t cid = rj_smem_inc(&sm->local_next[o]);
	if (cid >= sm->local_end[o])
		cid = atomicAdd(t.pool_next, 1u); // reserve ran dry inside one round (extreme skew)
	if (cid >= s.pool_chunks) {
		atomicOr(pr.error_flag, RJ_ERR_POOL);
		return;
	}
	if (old != RJ_NONE)
		t.chunk_entries[old >> 5] = RJ_CHUNK; // a chunk is only replaced when all its sectors are written
	t.chunk_part[cid] = (uint16_t)p;
	rj_global_red_inc(&t.dir_cnt[p]);
	sm->chunk[p] = cid << 5;
}

__device__ static inline void rj_refill_ids(const RJSide &s, RJP1Smem *sm)
{
	// thread 0 only, during the insert phase (no flush lane is allocating then)
	for (int o = 0; o < s.world; o++) {
		if (sm->local_end[o] - min(sm->local_next[o], sm->local_end[o]) < s.id_low) {
			const uint32_t base = atomicAdd(s.dst[o].pool_next, s.id_batch);
			sm->local_next[o] = base;
			sm->local_end[o] = base + s.id_batch; // ids left in the old range stay unused (chunk_part 0xffff)
		}
	}
}

__device__ static inline void rj_park(RJP1Smem *sm, const RJParams &pr, uint32_t item, int par)
{
	// staging row full until this round's flush: the key waits one round
	const uint32_t o = rj_smem_inc(&sm->ovf_count[par]);
	if (o < RJ_OVF_CAP)
		sm->ovf[par][o] = item;
	else
		atomicOr(pr.error_flag, RJ_ERR_SKEW);
}

__device__ static inline void rj_insert(RJP1Smem *sm, const RJParams &pr, uint32_t item, int par)
{
	const uint32_t p = item >> 16;
	const uint32_t pos = rj_smem_inc(&sm->fill[p]);
	if (pos < RJ_CAP) {
		sm->stage[p * RJ_CAP + pos] = (uint16_t)item;
		if (pos == RJ_FLUSH - 1)
			sm->worklist[par][rj_smem_inc(&sm->wl_count[par])] = (uint16_t)p;
	} else {
		rj_park(sm, pr, item, par);
	}
}

__device__ static inline void rj_round_begin(const RJSide &s, const RJParams &pr, RJP1Smem *sm, int par)
{
	const int tid = threadIdx.x;
	if (tid == 0)
		rj_refill_ids(s, sm);
	// keys parked by the previous round go first (their rows were flushed since)
	const uint32_t novf = min(sm->ovf_count[par ^ 1], (uint32_t)RJ_OVF_CAP);
	for (uint32_t i = tid; i < novf; i += RJ_P1_THREADS)
		rj_insert(sm, pr, sm->ovf[par ^ 1][i], par);
}

// Hot loop of pass 1 for the common case: a complete tile of a column without NULLs/tombstones whose
// [min, max] lies inside the partitioned key range, so no per-key validity test is needed and all arithmetic
// is 32-bit.  Per key: subtract, shift, one shared-memory atomic (slot), one 2-byte shared store; the rare
// events (row completed a sector -> queue it; row full -> park the key) are predicated, never branched on.
__device__ static inline void rj_insert_tile_fast(const RJParams &pr, RJP1Smem *sm, const int4 *buf, int par)
{
	constexpr int NK = RJ_P1_LOADS * 2;
	uint32_t d[NK], pos[NK], widx[NK];
	const uint32_t kmin_lo = (uint32_t)(unsigned long long)pr.kmin;
#pragma unroll
	for (int j = 0; j < RJ_P1_LOADS; j++) {
		d[2 * j] = (uint32_t)buf[j].x - kmin_lo;     // low words: key - kmin < 2^32 is guaranteed by the caller
		d[2 * j + 1] = (uint32_t)buf[j].z - kmin_lo;
	}
#pragma unroll
	for (int k = 0; k < NK; k++)
		pos[k] = rj_smem_inc(&sm->fill[d[k] >> pr.shift]);
#pragma unroll
	for (int k = 0; k < NK; k++) {
		widx[k] = 0;
		if (pos[k] == RJ_FLUSH - 1)
			widx[k] = rj_smem_inc(&sm->wl_count[par]);
	}
	uint32_t park_mask = 0;
#pragma unroll
	for (int k = 0; k < NK; k++) {
		const uint32_t p = d[k] >> pr.shift;
		if (pos[k] < RJ_CAP)
			sm->stage[p * RJ_CAP + pos[k]] = (uint16_t)(d[k] & pr.mask);
		if (pos[k] == RJ_FLUSH - 1)
			sm->worklist[par][widx[k]] = (uint16_t)p;
		park_mask |= pos[k] >= RJ_CAP ? (1u << k) : 0u;
	}
	if (park_mask) {
#pragma unroll
		for (int k = 0; k < NK; k++)
			if (park_mask & (1u << k))
				rj_park(sm, pr, ((d[k] >> pr.shift) << 16) | (d[k] & pr.mask), par);
	}
}

// generic insert phase.  FULL: every row of the tile exists (no bounds checks)
template <bool HAS_PRESENT, bool FULL>
__device__ static inline void rj_insert_tile(const RJSide &s, const RJParams &pr, RJP1Smem *sm, const int4 *buf, uint64_t tile,
		int par)
{
	const int tid = threadIdx.x;
	{
		const uint64_t base_pair = tile * (RJ_P1_TILE / 2);
		uint32_t item[RJ_P1_LOADS * 2], pos[RJ_P1_LOADS * 2];
#pragma unroll
		for (int j = 0; j < RJ_P1_LOADS; j++) {
			const uint64_t pi = base_pair + (uint64_t)j * RJ_P1_THREADS + tid;
			const unsigned long long k0 = ((unsigned long long)(unsigned)buf[j].y << 32) | (unsigned)buf[j].x;
			const unsigned long long k1 = ((unsigned long long)(unsigned)buf[j].w << 32) | (unsigned)buf[j].z;
			const unsigned long long d0 = k0 - (unsigned long long)pr.kmin, d1 = k1 - (unsigned long long)pr.kmin;
			bool ok0 = d0 < pr.range, ok1 = d1 < pr.range;
			if (!FULL) {
				ok0 = ok0 && pi * 2 < s.n;
				ok1 = ok1 && pi * 2 + 1 < s.n;
			}
			if (HAS_PRESENT) {
				const uint32_t pw = (FULL || pi * 2 < s.n) ? (s.present[pi >> 4] >> ((pi & 15) * 2)) : 0u;
				ok0 = ok0 && (pw & 1u);
				ok1 = ok1 && (pw & 2u);
			}
			item[2 * j] = ok0 ? ((((uint32_t)d0 >> pr.shift) << 16) | ((uint32_t)d0 & pr.mask)) : RJ_NONE;
			item[2 * j + 1] = ok1 ? ((((uint32_t)d1 >> pr.shift) << 16) | ((uint32_t)d1 & pr.mask)) : RJ_NONE;
		}
		// all slot requests of this thread are issued back to back (independent shared-memory atomics) ...
#pragma unroll
		for (int k = 0; k < RJ_P1_LOADS * 2; k++)
			pos[k] = item[k] != RJ_NONE ? rj_smem_inc(&sm->fill[item[k] >> 16]) : RJ_NONE;
		// ... then consumed; the two rare events (row just completed a sector / row full) are only recorded here
		uint32_t queue_mask = 0, park_mask = 0;
#pragma unroll
		for (int k = 0; k < RJ_P1_LOADS * 2; k++) {
			if (pos[k] < RJ_CAP)
				sm->stage[(item[k] >> 16) * RJ_CAP + pos[k]] = (uint16_t)item[k];
			queue_mask |= (pos[k] == RJ_FLUSH - 1) ? (1u << k) : 0u;
			park_mask |= (pos[k] >= RJ_CAP && item[k] != RJ_NONE) ? (1u << k) : 0u;
		}
		if (queue_mask) {
#pragma unroll
			for (int k = 0; k < RJ_P1_LOADS * 2; k++)
				if (queue_mask & (1u << k))
					sm->worklist[par][rj_smem_inc(&sm->wl_count[par])] = (uint16_t)(item[k] >> 16);
		}
		if (park_mask) {
#pragma unroll
			for (int k = 0; k < RJ_P1_LOADS * 2; k++)
				if (park_mask & (1u << k))
					rj_park(sm, pr, item[k], par);
		}
	}
}

// barrier, flush the queued rows, barrier
__device__ static inline void rj_round_end(const RJSide &s, const RJParams &pr, RJP1Smem *sm, int par)
{
	const int tid = threadIdx.x;
	__syncthreads();

	const uint32_t nwl = sm->wl_count[par];
	if (tid == 0) {
		// the other parity's lists were consumed (worklist: last round's flush; parked keys: above)
		sm->wl_count[par ^ 1] = 0;
		sm->ovf_count[par ^ 1] = 0;
	}
	// flush: one lane per queued partition writes its first 16 remainders as ONE aligned 32-byte sector
	for (uint32_t w = tid; w < nwl; w += RJ_P1_THREADS) {
		const uint32_t p = sm->worklist[par][w];
		const uint32_t f = min(sm->fill[p], (uint32_t)RJ_CAP);
		uint32_t ch = sm->chunk[p];
		if (ch == RJ_NONE || (ch & 31u) == RJ_BLOCKS_PER_CHUNK) {
			rj_new_chunk(s, pr, sm, p);
			ch = sm->chunk[p];
		}
		uint2 *row = reinterpret_cast<uint2*>(&sm->stage[p * RJ_CAP]); // 40-byte rows are 8-byte aligned
		const uint2 a = row[0], b = row[1], c = row[2], d = row[3], e = row[4];
		if (ch != RJ_NONE && (ch & 31u) < RJ_BLOCKS_PER_CHUNK) {
			int4 *dst = reinterpret_cast<int4*>(s.dst[rj_owner(s, pr, p)].pool + (size_t)(ch >> 5) * RJ_CHUNK + (ch & 31u) * RJ_FLUSH);
			dst[0] = make_int4((int)a.x, (int)a.y, (int)b.x, (int)b.y);
			dst[1] = make_int4((int)c.x, (int)c.y, (int)d.x, (int)d.y);
			sm->chunk[p] = ch + 1;
		}
		row[0] = e; // keep the (at most 4) remainders behind the flushed sector
		sm->fill[p] = f - RJ_FLUSH;
	}
	__syncthreads();
}

template <bool HAS_PRESENT>
__device__ static inline void rj_load_tile(const RJSide &s, uint64_t tile, int4 *dst)
{
	const int4 *src = reinterpret_cast<const int4*>(s.keys);
	const uint64_t npairs = s.n / 2;
	const uint64_t base_pair = tile * (RJ_P1_TILE / 2);
	// pull the tile this CTA will load two rounds from now into L2 (one 128-byte line per thread)
	const uint64_t pf_first = (tile + 2ull * gridDim.x) * RJ_P1_TILE + (uint64_t)threadIdx.x * 16;
	if (threadIdx.x < RJ_P1_TILE / 16 && pf_first + 16 <= s.n)
		asm volatile("prefetch.global.L2 [%0];" ::"l"(s.keys + pf_first));
	if (base_pair + RJ_P1_TILE / 2 <= npairs) {
#pragma unroll
		for (int j = 0; j < RJ_P1_LOADS; j++)
			dst[j] = mdb_ldg_stream(src + base_pair + (uint64_t)j * RJ_P1_THREADS + threadIdx.x);
	} else {
#pragma unroll
		for (int j = 0; j < RJ_P1_LOADS; j++) {
			const uint64_t pi = base_pair + (uint64_t)j * RJ_P1_THREADS + threadIdx.x;
			if (pi < npairs) {
				dst[j] = mdb_ldg_stream(src + pi);
			} else if (pi == npairs && (s.n & 1)) {
				const unsigned long long last = (unsigned long long)s.keys[s.n - 1];
				dst[j] = make_int4((int)(unsigned)last, (int)(unsigned)(last >> 32), 0, 0);
			} else {
				dst[j] = make_int4(0, 0, 0, 0);
			}
		}
	}
}

// every partition's partial sector goes out, chunk entry counts are finalised
__device__ static inline void rj_drain(const RJSide &s, const RJParams &pr, RJP1Smem *sm)
{
	for (int p = threadIdx.x; p < pr.nparts; p += RJ_P1_THREADS) {
		const uint32_t f = min(sm->fill[p], (uint32_t)RJ_CAP);
		const RJTarget &t = s.dst[rj_owner(s, pr, p)];
		uint32_t ch = sm->chunk[p];
		if (f > 0) {
			if (ch == RJ_NONE || (ch & 31u) == RJ_BLOCKS_PER_CHUNK) {
				rj_new_chunk(s, pr, sm, p);
				ch = sm->chunk[p];
			}
			if (ch != RJ_NONE && (ch & 31u) < RJ_BLOCKS_PER_CHUNK) {
				uint16_t *dst = t.pool + (size_t)(ch >> 5) * RJ_CHUNK + (ch & 31u) * RJ_FLUSH;
				for (uint32_t i = 0; i < f; i++)
					dst[i] = sm->stage[p * RJ_CAP + i];
				t.chunk_entries[ch >> 5] = (uint16_t)((ch & 31u) * RJ_FLUSH + f);
			}
		} else if (ch != RJ_NONE) {
			t.chunk_entries[ch >> 5] = (uint16_t)((ch & 31u) * RJ_FLUSH);
		}
	}
	__threadfence_system(); // remote owners read these chunks after the next cross-rank barrier
}

__device__ static inline void rj_smem_init(RJP1Smem *sm)
{
	const int tid = threadIdx.x;
	for (int p = tid; p < RJ_MAX_PART; p += RJ_P1_THREADS) {
		sm->fill[p] = 0;
		sm->chunk[p] = RJ_NONE;
	}
	if (tid < 2) {
		sm->ovf_count[tid] = 0;
		sm->wl_count[tid] = 0;
	}
	if (tid < RJ_MAX_RANKS)
		sm->local_next[tid] = sm->local_end[tid] = 0;
	__syncthreads();
}

// Pass 1, lean variant: column without NULLs/tombstones whose [min, max] lies inside the partitioned range.
// Complete tiles run the branch-free 32-bit hot loop; the ragged tail (< one tile) is inserted key by key by CTA 0.
__global__ void __launch_bounds__(RJ_P1_THREADS, 1) k_radix_partition_fast(RJSide s, RJParams pr)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RJP1Smem *sm = reinterpret_cast<RJP1Smem*>(smem_raw);
	rj_smem_init(sm);

	const uint64_t nfull = s.n / RJ_P1_TILE;
	const int4 *src = reinterpret_cast<const int4*>(s.keys);
	int4 buf_a[RJ_P1_LOADS], buf_b[RJ_P1_LOADS];
	int par = 0;
	auto load = [&](uint64_t tile, int4 *dst) {
		const uint64_t pf_first = (tile + 2ull * gridDim.x) * RJ_P1_TILE + (uint64_t)threadIdx.x * 16;
		if (threadIdx.x < RJ_P1_TILE / 16 && pf_first + 16 <= s.n)
			asm volatile("prefetch.global.L2 [%0];" ::"l"(s.keys + pf_first));
		const int4 *t = src + tile * (RJ_P1_TILE / 2) + threadIdx.x;
#pragma unroll
		for (int j = 0; j < RJ_P1_LOADS; j++)
			dst[j] = mdb_ldg_stream(t + j * RJ_P1_THREADS);
	};
	auto round = [&](const int4 *buf) {
		rj_round_begin(s, pr, sm, par);
		rj_insert_tile_fast(pr, sm, buf, par);
		rj_round_end(s, pr, sm, par);
		par ^= 1;
	};
	uint64_t tile = blockIdx.x;
	if (tile < nfull)
		load(tile, buf_a);
	while (tile < nfull) {
		uint64_t next = tile + gridDim.x;
		if (next < nfull)
			load(next, buf_b);
		round(buf_a);
		tile = next;
		if (tile >= nfull)
			break;
		next = tile + gridDim.x;
		if (next < nfull)
			load(next, buf_a);
		round(buf_b);
		tile = next;
	}
	// ragged tail and parked keys
	bool tail_done = blockIdx.x != 0 || nfull * RJ_P1_TILE == s.n;
	while (!tail_done || sm->ovf_count[par ^ 1] != 0) {
		rj_round_begin(s, pr, sm, par);
		if (!tail_done) {
			for (uint64_t r = nfull * RJ_P1_TILE + threadIdx.x; r < s.n; r += RJ_P1_THREADS) {
				const uint32_t d = (uint32_t)(unsigned long long)s.keys[r] - (uint32_t)(unsigned long long)pr.kmin;
				rj_insert(sm, pr, ((d >> pr.shift) << 16) | (d & pr.mask), par);
			}
			tail_done = true;
		}
		rj_round_end(s, pr, sm, par);
		par ^= 1;
	}
	rj_drain(s, pr, sm);
}

// Pass 1.  One persistent 1024-thread CTA per SM.  Keys are streamed with 128-bit loads, double-buffered in
// registers (ping-pong, no copies).  Each key costs one shared-memory atomic (slot in its partition's staging
// row) and one 2-byte shared store.  The thread that fills slot 16 of a row queues the partition; after the
// round's barrier one lane per queued partition writes 16 remainders as one aligned 32-byte sector into the
// CTA's current 512-byte chunk of that partition: DRAM only sees full-sector writes, 2 bytes per key.
template <bool HAS_PRESENT>
__global__ void __launch_bounds__(RJ_P1_THREADS, 1) k_radix_partition(RJSide s, RJParams pr)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RJP1Smem *sm = reinterpret_cast<RJP1Smem*>(smem_raw);
	const int tid = threadIdx.x;

	rj_smem_init(sm);

	const uint64_t ntiles = (s.n + RJ_P1_TILE - 1) / RJ_P1_TILE;
	const uint64_t nfull = s.n / RJ_P1_TILE; // tiles [0, nfull) are complete
	int4 buf_a[RJ_P1_LOADS], buf_b[RJ_P1_LOADS];
	uint64_t tile = blockIdx.x;
	int par = 0;
	if (tile < ntiles)
		rj_load_tile<HAS_PRESENT>(s, tile, buf_a);
	auto round = [&](const int4 *buf, uint64_t t) {
		rj_round_begin(s, pr, sm, par);
		if (t < nfull) {
			rj_insert_tile<HAS_PRESENT, true>(s, pr, sm, buf, t, par);
		} else if (t < ntiles) {
			rj_insert_tile<HAS_PRESENT, false>(s, pr, sm, buf, t, par);
		}
		rj_round_end(s, pr, sm, par);
		par ^= 1;
	};
	while (tile < ntiles) {
		uint64_t next = tile + gridDim.x;
		if (next < ntiles)
			rj_load_tile<HAS_PRESENT>(s, next, buf_b);
		round(buf_a, tile);
		tile = next;
		if (tile >= ntiles)
			break;
		next = tile + gridDim.x;
		if (next < ntiles)
			rj_load_tile<HAS_PRESENT>(s, next, buf_a);
		round(buf_b, tile);
		tile = next;
	}
	// keys still parked by the last round(s)
	while (sm->ovf_count[par ^ 1] != 0) // block-uniform: written before the last barrier
		round(buf_a, ntiles);

	rj_drain(s, pr, sm);
}

// exclusive scan of the per-partition chunk counts (single block, nparts <= 4096)
__global__ void k_radix_dir_scan(RJSide s, int nparts)
{
	__shared__ uint64_t warp_tot[33];
	uint64_t carry = 0;
	for (int base = 0; base < nparts + 1; base += blockDim.x) {
		int i = base + threadIdx.x;
		uint64_t v = i < nparts ? s.dst[s.self].dir_cnt[i] : 0;
		uint64_t incl = v;
		int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		for (int o = 1; o < 32; o <<= 1) {
			uint64_t n = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o)
				incl += n;
		}
		if (lane == 31)
			warp_tot[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint64_t w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0, wi = w;
			for (int o = 1; o < 32; o <<= 1) {
				uint64_t n = __shfl_up_sync(0xffffffffu, wi, o);
				if (lane >= o)
					wi += n;
			}
			warp_tot[lane] = wi - w;
			if (lane == 31)
				warp_tot[32] = wi;
		}
		__syncthreads();
		if (i < nparts + 1)
			s.dir_off[i] = carry + warp_tot[warp] + incl - v;
		carry += warp_tot[32];
		__syncthreads();
	}
}

#include "mdb_radix_pass2.cuh"
#include "mdb_radix_dist.cuh"

// chunk-id reservation sizes: a CTA needs at most one fresh chunk per owned partition at start-up
static void rj_id_policy(int world, uint32_t *batch, uint32_t *low)
{
	*batch = world == 1 ? 8192u : std::max(1024u, 2u * RJ_MAX_PART / (uint32_t)world);
	*low = *batch / 4;
}

// chunks one owner's pool must hold: data chunks (with slack for imbalance) + one ragged chunk per
// (source CTA, owned partition) + ids abandoned at refills + one reserve per (source CTA, owner)
static uint64_t rj_pool_chunks(uint64_t rows_for_owner, int world, int grid)
{
	uint32_t batch, low;
	rj_id_policy(world, &batch, &low);
	uint64_t chunks = rows_for_owner / RJ_CHUNK + (uint64_t)grid * RJ_MAX_PART;
	chunks += chunks / 2 + (uint64_t)world * grid * batch + 1024;
	return chunks;
}

// byte layout of one side inside an exchange arena (identical on every rank)
struct RJArenaLayout {
	size_t pool, chunk_part, chunk_entries, dir_cnt, pool_next, bytes;
};

static RJArenaLayout rj_arena_layout(uint64_t chunks)
{
	auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
	RJArenaLayout l;
	l.pool = 0;
	l.chunk_part = up(l.pool + chunks * RJ_CHUNK * sizeof(uint16_t));
	l.chunk_entries = up(l.chunk_part + chunks * sizeof(uint16_t));
	l.dir_cnt = up(l.chunk_entries + chunks * sizeof(uint16_t));
	l.pool_next = up(l.dir_cnt + (RJ_MAX_PART + 1) * sizeof(uint32_t));
	l.bytes = up(l.pool_next + 256);
	return l;
}

static void rj_target_at(RJTarget *t, void *base, const RJArenaLayout &l)
{
	char *b = (char*)base;
	t->pool = (uint16_t*)(b + l.pool);
	t->chunk_part = (uint16_t*)(b + l.chunk_part);
	t->chunk_entries = (uint16_t*)(b + l.chunk_entries);
	t->dir_cnt = (uint32_t*)(b + l.dir_cnt);
	t->pool_next = (uint32_t*)(b + l.pool_next);
}

// arena_bases == nullptr: all chunks stay on this GPU (single-GPU plan, or NCCL exchange afterwards);
// otherwise dst[r] points into rank r's arena (at byte offset arena_off) and pass 1 writes over NVLink
static int rj_side_setup(mdbcu_ctx *ctx, DevTemp &tmp, RJSide *s, const mdbcu_table *t, int col, int grid, uint64_t chunks,
		void *const *arena_bases, size_t arena_off)
{
	memset(s, 0, sizeof(*s));
	s->keys = t->cols[col].data;
	s->present = col_all_present(t, col) ? nullptr : t->cols[col].present;
	s->n = t->n_slots;
	if (chunks >= (1ull << 27))
		return MDBCU_EUNSUPPORTED;
	s->pool_chunks = (uint32_t)chunks;
	s->world = arena_bases ? ctx->world : 1;
	s->self = arena_bases ? ctx->rank : 0;
	rj_id_policy(s->world, &s->id_batch, &s->id_low);
	RJTarget &own = s->dst[s->self];
	if (arena_bases) {
		const RJArenaLayout l = rj_arena_layout(chunks);
		for (int r = 0; r < ctx->world; r++)
			rj_target_at(&s->dst[r], (char*)arena_bases[r] + arena_off, l);
	} else {
		MDB_TRY(tmp.alloc(&own.pool, chunks * RJ_CHUNK));
		MDB_TRY(tmp.alloc(&own.pool_next, 1));
		MDB_TRY(tmp.alloc(&own.chunk_part, chunks));
		MDB_TRY(tmp.alloc(&own.chunk_entries, chunks));
		MDB_TRY(tmp.alloc(&own.dir_cnt, RJ_MAX_PART + 1));
	}
	s->pool = own.pool;
	MDB_TRY(tmp.alloc(&s->dir_fill, RJ_MAX_PART + 1));
	MDB_TRY(tmp.alloc(&s->dir_off, RJ_MAX_PART + 2));
	MDB_TRY(tmp.alloc(&s->dir, chunks));
	CUDA_TRY(ctx, cudaMemsetAsync(own.pool_next, 0, sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(own.chunk_part, 0xff, chunks * sizeof(uint16_t), ctx->stream)); // 0xffff = never allocated
	CUDA_TRY(ctx, cudaMemsetAsync(own.dir_cnt, 0, (RJ_MAX_PART + 1) * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(s->dir_fill, 0, (RJ_MAX_PART + 1) * sizeof(uint32_t), ctx->stream));
	return MDBCU_OK;
}

static void launch_partition(mdbcu_ctx *ctx, int grid, const RJSide &s, const RJParams &pr)
{
	if (s.present)
		MDB_LAUNCH(ctx, k_radix_partition<true>, grid, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
	else if (s.all_in_range && pr.range <= 0xffffffffull)
		MDB_LAUNCH(ctx, k_radix_partition_fast, grid, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
	else
		MDB_LAUNCH(ctx, k_radix_partition<false>, grid, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
}

int mdb_select_radix_joincount(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (plan->n_tables != 2 || plan->n_joins != 1 || plan->joins[0].cross || plan->n_pred != 0 || plan->n_group != 1 ||
			plan->n_out < 1 || plan->n_out > 4)
		return MDBCU_EUNSUPPORTED;
	const bool dist = (plan->flags & MDBCU_PLAN_DISTRIBUTED) != 0;
	if (dist && !ctx->nccl_comm)
		return mdb_fail(ctx, MDBCU_EERROR, "MDBCU_PLAN_DISTRIBUTED needs mdbcu_comm_init first");
	const mdbcu_join &jn = plan->joins[0];
	if (jn.left.tbl != 0 || jn.right.tbl != 1)
		return MDBCU_EUNSUPPORTED;
	const mdbcu_table *ta = plan->tables[0], *tb = plan->tables[1];
	if (jn.left.col < 0 || jn.left.col >= ta->ncols || jn.right.col < 0 || jn.right.col >= tb->ncols)
		return MDBCU_EUNSUPPORTED;
	DevColumn ca = ta->cols[jn.left.col], cb = tb->cols[jn.right.col]; // copies: distributed plans overwrite the bounds
	auto intlike = [](int type) { return type == MDBCU_CT_INTEGER || type == MDBCU_CT_DATE || type == MDBCU_CT_DATETIME; };
	if (!intlike(ca.type) || !intlike(cb.type) || !ca.stats_ok || !cb.stats_ok)
		return MDBCU_EUNSUPPORTED;
	if (dist) {
		// every rank must take the same decisions: use the bounds over ALL shards
		if (!ca.gstats_ok || !cb.gstats_ok)
			return mdb_fail(ctx, MDBCU_EERROR, "distributed plan: call mdbcu_table_sync_stats on every sharded table first");
		ca.imin = ca.gmin;
		ca.imax = ca.gmax;
		cb.imin = cb.gmin;
		cb.imax = cb.gmax;
	}
	auto is_key = [&](const mdbcu_colref &r) {
		return (r.tbl == 0 && r.col == jn.left.col) || (r.tbl == 1 && r.col == jn.right.col);
	};
	if (!is_key(plan->group[0]))
		return MDBCU_EUNSUPPORTED;
	for (int o = 0; o < plan->n_out; o++) {
		if (plan->out[o].kind == MDBCU_OUT_COUNT_STAR)
			continue;
		if (plan->out[o].kind != MDBCU_OUT_COLUMN || !is_key(plan->out[o].ref))
			return MDBCU_EUNSUPPORTED;
	}
	if (!dist && ta->n_slots + tb->n_slots < (1ull << 20))
		return MDBCU_EUNSUPPORTED; // small inputs: general operators (they also return the reference's row order)

	// only keys inside both columns' [min, max] (zone-map statistics kept by the mirror) can ever match
	long long kmin = std::max(ca.imin, cb.imin), kmax = std::min(ca.imax, cb.imax);
	if (ca.imin > ca.imax || cb.imin > cb.imax || kmin > kmax) {
		ctx->stats.path = MDBCU_PATH_RADIX_JOINCOUNT;
		return mdb_result_alloc(ctx, plan, res, 0, false);
	}
	unsigned long long range = (unsigned long long)kmax - (unsigned long long)kmin + 1ull;
	if (range == 0 || range > ((unsigned long long)RJ_MAX_PART << RJ_MAX_SHIFT) || range < 4096)
		return MDBCU_EUNSUPPORTED;
	{
		// when the two columns cover almost the same interval, partition over the UNION of the intervals instead:
		// every key of both sides is then in range and the hot loop needs no per-key range test
		long long umin = std::min(ca.imin, cb.imin), umax = std::max(ca.imax, cb.imax);
		unsigned long long urange = (unsigned long long)umax - (unsigned long long)umin + 1ull;
		if (urange != 0 && urange <= ((unsigned long long)RJ_MAX_PART << RJ_MAX_SHIFT) && urange <= range + range / 4) {
			kmin = umin;
			kmax = umax;
			range = urange;
		}
	}
	int bits = 0;
	while ((1ull << bits) < range)
		bits++;
	int shift = std::max(0, bits - 12);
	int nparts = (int)((range + (1ull << shift) - 1) >> shift);

	ctx->stats.path = MDBCU_PATH_RADIX_JOINCOUNT;
	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	const int grid1 = ctx->num_sms;
	RJSide sa, sb;
	RJParams pr;
	// multi-GPU plans: pass 1 writes every chunk straight into the owner GPU's arena over NVLink (CUDA IPC);
	// MDBCU_EXCHANGE=nccl selects the staged variant instead (partition locally, then a grouped send/recv)
	static const bool use_nccl_exchange = getenv("MDBCU_EXCHANGE") && strcmp(getenv("MDBCU_EXCHANGE"), "nccl") == 0;
	const bool p2p = dist && !use_nccl_exchange;
	if (p2p) {
		if (!ta->global_slots || !tb->global_slots)
			return mdb_fail(ctx, MDBCU_EERROR, "distributed plan: call mdbcu_table_sync_stats on every sharded table first");
		const uint64_t ca_chunks = rj_pool_chunks(ta->global_slots / ctx->world + 1, ctx->world, grid1);
		const uint64_t cb_chunks = rj_pool_chunks(tb->global_slots / ctx->world + 1, ctx->world, grid1);
		const RJArenaLayout la = rj_arena_layout(ca_chunks), lb = rj_arena_layout(cb_chunks);
		void *bases[MDB_MAX_RANKS];
		MDB_TRY(mdb_comm_arena(ctx, la.bytes + lb.bytes, bases));
		MDB_TRY(rj_side_setup(ctx, tmp, &sa, ta, jn.left.col, grid1, ca_chunks, bases, 0));
		MDB_TRY(rj_side_setup(ctx, tmp, &sb, tb, jn.right.col, grid1, cb_chunks, bases, la.bytes));
	} else {
		MDB_TRY(rj_side_setup(ctx, tmp, &sa, ta, jn.left.col, grid1, rj_pool_chunks(ta->n_slots, 1, grid1), nullptr, 0));
		MDB_TRY(rj_side_setup(ctx, tmp, &sb, tb, jn.right.col, grid1, rj_pool_chunks(tb->n_slots, 1, grid1), nullptr, 0));
	}
	sa.all_in_range = ca.imin >= kmin && ca.imax <= kmax;
	sb.all_in_range = cb.imin >= kmin && cb.imax <= kmax;
	pr.kmin = kmin;
	pr.range = range;
	pr.shift = shift;
	pr.mask = (1u << shift) - 1u;
	pr.nparts = nparts;
	pr.part_first = 0;
	pr.part_end = nparts;
	if (dist) {
		pr.part_first = (int)((uint64_t)ctx->rank * nparts / ctx->world);
		pr.part_end = (int)((uint64_t)(ctx->rank + 1) * nparts / ctx->world);
	}
	uint32_t *d_flags; // [0] error flags, [1] partition counter
	unsigned long long *d_cursor;
	MDB_TRY(tmp.alloc(&d_flags, 2));
	MDB_TRY(tmp.alloc(&d_cursor, 1));
	CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
	pr.error_flag = d_flags;

	// upper bound of groups this rank can emit: one per key of the partitions it owns
	uint64_t cap_groups = std::min<uint64_t>((uint64_t)(pr.part_end - pr.part_first) << shift, range);
	if (!dist)
		cap_groups = std::min<uint64_t>(cap_groups, std::min<uint64_t>(ta->n_slots, tb->n_slots));
	MDB_TRY(mdb_result_alloc(ctx, plan, res, 0, false));
	RJOut out;
	memset(&out, 0, sizeof(out));
	out.nout = plan->n_out;
	out.cursor = d_cursor;
	out.cap = cap_groups;
	for (int o = 0; o < plan->n_out; o++) {
		mdb_free(ctx, res->cols[o].cells);
		mdb_free(ctx, res->cols[o].nulls);
		res->cols[o].nulls = nullptr; // NULL keys never join (executor_select.c:716-738): no NULL cells in this result
		res->cols[o].cells = nullptr;
		MDB_TRY(mdb_alloc(ctx, &res->cols[o].cells, cap_groups));
		out.cells[o] = res->cols[o].cells;
		out.is_count[o] = plan->out[o].kind == MDBCU_OUT_COUNT_STAR;
	}

	static bool attr_done = false;
	if (!attr_done) {
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2 * RJ_DESC_CAP * (int)sizeof(RJDesc)));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 2 * RJ_DESC_CAP * (int)sizeof(RJDesc)));
		attr_done = true;
	}

	if (p2p) {
		// every rank's arena must be reset before any peer starts writing into it
		clock.begin(6);
		MDB_TRY(mdb_comm_barrier_or(ctx, d_flags, nullptr));
	}
	clock.begin(1);
	launch_partition(ctx, grid1, sa, pr);
	launch_partition(ctx, grid1, sb, pr);
	if (p2p) {
		// all remote stores have landed once every rank's pass 1 has completed; error flags are shared so that
		// every rank takes the same exit
		clock.begin(6);
		uint32_t any = 0;
		MDB_TRY(mdb_comm_barrier_or(ctx, d_flags, &any));
		if (any)
			return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "distributed radix join: pass 1 failed on some rank (flags %u: "
					"1 = chunk pool exhausted, 4 = extreme skew)", any);
	}
	clock.begin(7);
	MDB_LAUNCH(ctx, k_radix_dir_scan, 1, 1024, 0, sa, nparts);
	MDB_LAUNCH(ctx, k_radix_dir_scan, 1, 1024, 0, sb, nparts);
	MDB_LAUNCH(ctx, k_radix_dir_fill, ctx->num_sms * 4, 256, 0, sa);
	MDB_LAUNCH(ctx, k_radix_dir_fill, ctx->num_sms * 4, 256, 0, sb);
	uint64_t exchanged = 0;
	if (dist && !p2p) {
		clock.begin(6);
		RJSide *both[2] = {&sa, &sb};
		MDB_TRY(rj_exchange(ctx, tmp, &pr, both, &exchanged));
	}
	clock.begin(2);

	// 4-bit counters (two CTAs per SM) when keys are mostly unique per side, 8-bit otherwise; a wrapped
	// counter is detected by the checksum and the pass is repeated one width up before giving up
	const uint64_t D = 1ull << shift;
	bool try4 = std::max(ta->n_slots, tb->n_slots) <= 2 * range;
	uint64_t ngroups = 0;
	uint32_t flags = 0;
	for (int attempt = try4 ? 0 : 1; attempt < 2; attempt++) {
		const int bitsw = attempt == 0 ? 4 : 8;
		const size_t smem2 = 2 * (size_t)std::max<uint64_t>(1, D * bitsw / 32) * sizeof(uint32_t) + 2 * RJ_DESC_CAP * sizeof(RJDesc);
		const int grid2 = std::max(1, std::min(pr.part_end - pr.part_first, ctx->num_sms * (bitsw == 4 ? 2 : 1)));
		if (bitsw == 4)
			MDB_LAUNCH(ctx, (k_radix_joincount<4, 512>), grid2, 512, smem2, sa, sb, pr, out, d_flags + 1);
		else
			MDB_LAUNCH(ctx, (k_radix_joincount<8, 1024>), grid2, 1024, smem2, sa, sb, pr, out, d_flags + 1);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			return mdb_fail(ctx, MDBCU_ECUDA, "radix join launch failed: %s", cudaGetErrorString(e));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, d_cursor, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar + 1, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		ngroups = ctx->h_scalar[0];
		flags = (uint32_t)(ctx->h_scalar[1] & 0xffffffffu);
		if (flags != RJ_ERR_COUNTER || attempt == 1)
			break;
		// a 4-bit counter wrapped: repeat pass 2 with 8-bit counters (the partitioned remainders are still valid)
		CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), ctx->stream));
		CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
	}
	clock.finish();
	if (flags || ngroups > cap_groups) {
		if (dist)
			return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "distributed radix join: key multiplicity or skew beyond the counter width "
					"(flags %u); the general operators are single-GPU only", flags);
		return MDBCU_EUNSUPPORTED; // heavy duplicates / skew / pool exhaustion: the general operators redo the query
	}
	res->nrows = ngroups;
	ctx->stats.exchange_bytes = exchanged;

	ctx->stats.algorithmic_bytes = 8ull * (ta->n_slots + tb->n_slots) + 8ull * plan->n_out * ngroups;
	ctx->stats.dominant_ms = ctx->stats.phase_ms[1] + ctx->stats.phase_ms[2];
	ctx->stats.dominant_bytes = ctx->stats.algorithmic_bytes;
	return MDBCU_OK;
}
