// mdb_common.cuh - shared declarations of libmidoridb_cuda.so (sm_100a only).
//
// Device mirror layout (SURVEY.md 8a "a1"): the reference keeps rows as an array-of-structs inside
// 4 KiB pages (include/primitive/row.h:22-28, datablock.h:9-13).  The mirror is columnar: one
// 8-byte-cell array per column plus a "present" bitmap (bit = row is live AND cell is not NULL) and one
// "live" bitmap per table.  Device row id = page * rows_per_page + slot with
// rows_per_page = floor(4095 / row_size) (row.c:38), so (page, slot) <-> row id is stable across DML.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <string>
#include <vector>

#include "midoridb_cuda.h"

#define MDB_WARP 32
#define MDB_MAX_RANKS 8 // one node: up to 8 GPUs behind NVSwitch
// error bits that travel with the cross-rank barrier (4 bits per rank; 1, 2, 4 are the radix join's RJ_ERR_* flags)
#define MDB_PEER_ABORT 8u   // the rank gave up on the query on the host (allocation failure ...)
#define MDB_PEER_TIMEOUT 8u // the rank never reached the barrier

struct mdbcu_ctx {
	int device = 0;
	int num_sms = 0;
	cudaStream_t stream = nullptr;
	std::string err;
	mdbcu_stats stats{};
	uint64_t total_launches = 0;
	// multi-GPU
	int rank = 0, world = 1;
	void *nccl_comm = nullptr;
	struct mdb_local_group *local_group = nullptr; // in-process communicator (mdbcu_comm_init_local), see mdb_comm.cu
	// exchange arena of the multi-GPU radix join (pass 1 writes its streams here, the peers' pass 2 reads them): own block +
	// every peer's block (CUDA IPC between processes, plain pointers inside one process)
	void *arena_local = nullptr;
	size_t arena_bytes = 0;
	void *arena_peer[MDB_MAX_RANKS] = {};
	uint32_t arena_epoch = 0;  // barriers passed on the arena's flag words (mdb_comm_arena_barrier)
	cudaStream_t side_stream = nullptr; // multi-GPU: side A's streams are fetched from the peers here while pass 1 of side B runs
	cudaEvent_t side_ev[2] = {};
	uint32_t arena_queries = 0; // distributed queries so far: alternate queries use alternate halves of the arena
	bool radix_attr_done = false; // dynamic shared-memory limits of the radix kernels raised on this context's device
	bool radix_gave_up = false; // the radix join handed the running query back because of its DATA (duplicates, skew)
	// reusable scratch: pinned host word for small D2H reads
	uint64_t *h_scalar = nullptr; // pinned, 64 entries
	uint64_t *d_scalar = nullptr; // device, 64 entries
	cudaEvent_t user_events[MDBCU_EVENT_SLOTS] = {};
	// scratch arena of the general operators: their temporaries are bump-allocated here in stack order (DevTemp) because
	// the stream-ordered pool takes 1-5 ms of host time for every block of hundreds of MB (MDBCU_TRACE=2 shows them).
	// scratch_top may run past scratch_cap: the overflow comes from the pool and the arena is regrown to scratch_peak
	// when the query has finished (mdb_scratch_settle).
	char *scratch = nullptr;
	size_t scratch_cap = 0, scratch_top = 0, scratch_peak = 0;
};

struct DevColumn {
	int type = 0;
	int width = 8;             // bytes in the reference row (column.c:255)
	int row_off = 0;           // byte offset inside row->data
	int64_t *data = nullptr;   // n_slots cells (double kept as raw bits)
	uint32_t *present = nullptr; // bitmap, bit i = live && not NULL
	bool has_nulls = false;
	bool stats_ok = false;     // imin/imax valid (conservative bounds over present cells)
	int64_t imin = 0, imax = 0;
	uint64_t sorted_version = 0; // table version at which `sorted` was computed (0 = never)
	bool sorted = false;       // cells [0, n_slots) are non-decreasing (checked on demand by the radix join, cached)
	bool gstats_ok = false;    // gmin/gmax = bounds over ALL ranks' shards (mdbcu_table_sync_stats)
	int64_t gmin = 0, gmax = 0;
};

struct mdbcu_table {
	mdbcu_ctx *ctx = nullptr;
	std::string name;
	int ncols = 0;
	size_t row_size = 0;       // 24 + sum(width)  (row.c:21)
	size_t rows_per_page = 0;  // floor(4095/row_size)
	uint64_t n_slots = 0;      // device rows in use (index space)
	uint64_t cap = 0;          // allocated rows
	uint64_t n_pages = 0;
	bool paged = false;
	bool all_live = true;      // every slot in [0, n_slots) is live
	uint64_t global_slots = 0; // sum of n_slots over all ranks' shards (mdbcu_table_sync_stats)
	uint64_t version = 1;      // bumped by every call that changes the table's contents
	uint32_t *live = nullptr;  // bitmap
	std::vector<DevColumn> cols;
};

struct ResultColumn {
	int type = MDBCU_CT_INTEGER;
	int64_t *cells = nullptr; // device
	uint8_t *nulls = nullptr; // device byte flags (nullptr = no NULLs)
};

struct mdbcu_result {
	mdbcu_ctx *ctx = nullptr;
	uint64_t nrows = 0;
	std::vector<ResultColumn> cols;
	uint64_t *order_key = nullptr; // device; optional canonical-order key (reference row order)
	bool owns = true;
};

// ---------------------------------------------------------------- error handling

int mdb_fail(mdbcu_ctx *ctx, int code, const char *fmt, ...);
void mdb_set_global_error(const char *msg);

#define CUDA_TRY(ctx, call)                                                                  \
	do {                                                                                 \
		cudaError_t _e = (call);                                                     \
		if (_e != cudaSuccess)                                                       \
			return mdb_fail((ctx), MDBCU_ECUDA, "%s failed at %s:%d: %s", #call, \
					__FILE__, __LINE__, cudaGetErrorString(_e));         \
	} while (0)

#define MDB_TRY(call)              \
	do {                       \
		int _rc = (call);  \
		if (_rc != MDBCU_OK) \
			return _rc; \
	} while (0)

// every kernel launch goes through this so `gpu_launches` is counted, not guessed
#define MDB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
	do {                                                                             \
		kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);         \
		(ctx)->stats.kernel_launches++;                                          \
		(ctx)->total_launches++;                                                 \
	} while (0)

#define CUDA_CHECK_LAUNCH(ctx) CUDA_TRY(ctx, cudaGetLastError())

// MDBCU_TRACE=1: path decisions; 2: also host calls that stall a query (pool growth)
static inline int mdb_trace_level()
{
	static int level = -1;
	if (level < 0) {
		const char *e = getenv("MDBCU_TRACE");
		level = e ? atoi(e) : 0;
	}
	return level;
}

// host wall clock between two points of a query, printed under MDBCU_TRACE=2
struct HostLap {
	std::chrono::steady_clock::time_point prev = std::chrono::steady_clock::now();
	void operator()(const char *where, unsigned long long n = 0)
	{
		if (mdb_trace_level() < 2)
			return;
		auto t = std::chrono::steady_clock::now();
		fprintf(stderr, "[mdbcu] %-28s %8.2f ms host  (%llu)\n", where,
				std::chrono::duration<double, std::milli>(t - prev).count(), n);
		prev = t;
	}
};

// ---------------------------------------------------------------- memory (stream-ordered pool)

template <typename T>
static inline int mdb_alloc(mdbcu_ctx *ctx, T **out, size_t count)
{
	void *p = nullptr;
	size_t bytes = (count ? count : 1) * sizeof(T);
	auto t0 = std::chrono::steady_clock::now();
	cudaError_t e = cudaMallocAsync(&p, bytes, ctx->stream);
	if (e == cudaErrorMemoryAllocation && ctx->scratch && ctx->scratch_top == 0) {
		// out of memory while the scratch arena of the general operators sits idle: give it back and try once more
		cudaGetLastError();
		cudaStreamSynchronize(ctx->stream);
		cudaFree(ctx->scratch);
		ctx->scratch = nullptr;
		ctx->scratch_cap = ctx->scratch_peak = 0;
		e = cudaMallocAsync(&p, bytes, ctx->stream);
	}
	if (mdb_trace_level() >= 2) {
		double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		if (ms > 0.2)
			fprintf(stderr, "[mdbcu] slow allocation: %zu bytes took %.2f ms\n", bytes, ms);
	}
	if (e != cudaSuccess) {
		*out = nullptr;
		return mdb_fail(ctx, e == cudaErrorMemoryAllocation ? MDBCU_ENOMEM : MDBCU_ECUDA,
				"device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
	}
	*out = (T*)p;
	return MDBCU_OK;
}

static inline void mdb_free(mdbcu_ctx *ctx, void *p)
{
	if (p)
		cudaFreeAsync(p, ctx->stream);
}

// RAII holder for query temporaries. use_scratch: take them from the context's scratch arena (operators of the
// general path: one stream, scopes strictly nested, so releasing = moving the top back).
struct DevTemp {
	mdbcu_ctx *ctx;
	std::vector<void*> ptrs;
	bool use_scratch;
	size_t mark;
	explicit DevTemp(mdbcu_ctx *c, bool scratch = false) : ctx(c), use_scratch(scratch), mark(c->scratch_top) {}
	~DevTemp()
	{
		for (void *p : ptrs)
			mdb_free(ctx, p);
		if (use_scratch)
			ctx->scratch_top = mark;
	}
	template <typename T>
	int alloc(T **out, size_t count)
	{
		if (use_scratch) {
			size_t bytes = ((count ? count : 1) * sizeof(T) + 511) & ~(size_t)511;
			size_t off = ctx->scratch_top;
			ctx->scratch_top += bytes;
			if (ctx->scratch_top > ctx->scratch_peak)
				ctx->scratch_peak = ctx->scratch_top;
			if (off + bytes <= ctx->scratch_cap) {
				*out = (T*)(ctx->scratch + off);
				return MDBCU_OK;
			}
		}
		int rc = mdb_alloc(ctx, out, count);
		if (rc == MDBCU_OK)
			ptrs.push_back(*out);
		return rc;
	}
};

// after a query: grow the arena to the largest footprint seen (at most a quarter of the device memory)
static inline void mdb_scratch_settle(mdbcu_ctx *ctx)
{
	ctx->scratch_top = 0;
	if (ctx->scratch_peak <= ctx->scratch_cap)
		return;
	size_t free_b = 0, total_b = 0;
	cudaStreamSynchronize(ctx->stream);
	if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
		cudaGetLastError();
		return;
	}
	size_t want = (ctx->scratch_peak + (ctx->scratch_peak >> 3) + ((size_t)2 << 20)) & ~(((size_t)2 << 20) - 1);
	if (want > total_b / 4 || want > (free_b + ctx->scratch_cap) / 2) {
		ctx->scratch_peak = ctx->scratch_cap; // does not fit: stay with the pool for the overflow
		return;
	}
	if (ctx->scratch)
		cudaFree(ctx->scratch);
	ctx->scratch = nullptr;
	ctx->scratch_cap = 0;
	void *p = nullptr;
	if (cudaMalloc(&p, want) != cudaSuccess) {
		cudaGetLastError();
		ctx->scratch_peak = 0;
		return;
	}
	ctx->scratch = (char*)p;
	ctx->scratch_cap = want;
	if (mdb_trace_level() >= 1)
		fprintf(stderr, "[mdbcu] scratch arena grown to %zu MiB\n", want >> 20);
}

// ---------------------------------------------------------------- device helpers

__host__ __device__ static inline uint64_t mdb_mix64(uint64_t x)
{
	x ^= x >> 33;
	x *= 0xff51afd7ed558ccdULL;
	x ^= x >> 33;
	x *= 0xc4ceb9fe1a85ec53ULL;
	x ^= x >> 33;
	return x;
}

__device__ static inline bool mdb_bit(const uint32_t *bm, uint64_t i)
{
	return (bm[i >> 5] >> (i & 31)) & 1u;
}

// streaming 128-bit load: data read once, keep it out of L1
__device__ static inline int4 mdb_ldg_stream(const int4 *p)
{
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
			: "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
			: "l"(p));
	return r;
}

__device__ static inline uint32_t mdb_ldg_stream_u32(const uint32_t *p)
{
	uint32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}

// order-preserving map double <-> int64 so MIN/MAX on doubles can use integer atomics
__host__ __device__ static inline int64_t mdb_dbl_to_ordered(int64_t bits)
{
	return bits < 0 ? (int64_t)(~(uint64_t)bits ^ 0x8000000000000000ULL) : bits;
}
__host__ __device__ static inline int64_t mdb_ordered_to_dbl(int64_t o)
{
	return o < 0 ? (int64_t)(~((uint64_t)o ^ 0x8000000000000000ULL)) : o;
}

static inline size_t mdb_div_up(size_t a, size_t b)
{
	return (a + b - 1) / b;
}

static inline int mdb_col_width(int type)
{
	return type == MDBCU_CT_TINYINT ? 1 : 8; // table_calc_column_space, column.c:255-293
}

// ---------------------------------------------------------------- cross-file entry points

// exclusive prefix sums (mdb_ops.cu); *_total are device pointers receiving the grand total
int mdb_scan_u32_u64(mdbcu_ctx *ctx, const uint32_t *in, uint64_t *out, size_t n, uint64_t *d_total);
int mdb_scan_u64_u64(mdbcu_ctx *ctx, const uint64_t *in, uint64_t *out, size_t n, uint64_t *d_total);
// read one 64-bit device word (synchronises the stream)
int mdb_read_u64(mdbcu_ctx *ctx, const uint64_t *d_ptr, uint64_t *out);

int mdb_table_reserve(mdbcu_table *t, uint64_t rows);
int mdb_result_alloc(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res, uint64_t nrows, bool with_order);
int mdb_table_refresh_stats(mdbcu_table *t, int col);

// physical paths (return MDBCU_EUNSUPPORTED when the plan does not fit so the caller can fall through)
int mdb_select_general(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res);
int mdb_select_scan_agg(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res);
int mdb_select_general_dist(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res);
int out_result_type(const mdbcu_plan *plan, int o); // MDBCU_CT_INTEGER or MDBCU_CT_DOUBLE of result column o
int mdb_select_radix_joincount(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res);
int mdb_select_direct_star(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res);

// multi-GPU exchange used by the radix join path (mdb_comm.cu)
void mdb_comm_destroy(mdbcu_ctx *ctx);
int mdb_comm_allgather_bytes(mdbcu_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank);
int mdb_comm_arena(mdbcu_ctx *ctx, size_t bytes, void **bases);
int mdb_comm_arena_barrier(mdbcu_ctx *ctx, const uint32_t *d_err, uint32_t *d_all);
int mdb_comm_arena_barrier_on(mdbcu_ctx *ctx, cudaStream_t stream, const uint32_t *d_err, uint32_t *d_all);
void mdb_comm_arena_destroy(mdbcu_ctx *ctx);
void mdb_comm_arena_abort(mdbcu_ctx *ctx);
int mdb_comm_allgather_u64(mdbcu_ctx *ctx, const uint64_t *send, uint64_t *recv, size_t count);
bool mdb_comm_ready(const mdbcu_ctx *ctx);
int mdb_comm_reduce_owned_u32(mdbcu_ctx *ctx, uint32_t *buf, uint64_t n, int nsides, uint64_t *bytes_sent);
int mdb_select_direct_count(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res, bool forced);
// tail operators on the finished result (mdb_tail.cu): HAVING, DISTINCT, ORDER BY, LIMIT
bool mdb_plan_has_tail(const mdbcu_plan *plan);
int mdb_validate_tail(mdbcu_ctx *ctx, const mdbcu_plan *plan);
int mdb_apply_tail(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res);

// two events around one kernel (stats.dominant_ms); destroyed on every return path
struct KernelTimer {
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	KernelTimer()
	{
		cudaEventCreate(&e0);
		cudaEventCreate(&e1);
	}
	KernelTimer(const KernelTimer&) = delete;
	KernelTimer &operator=(const KernelTimer&) = delete;
	void start(cudaStream_t s) { cudaEventRecord(e0, s); }
	void stop(cudaStream_t s) { cudaEventRecord(e1, s); }
	float ms() const // (the stream has been synchronised by the caller)
	{
		float v = 0.f;
		cudaEventElapsedTime(&v, e0, e1);
		return v;
	}
	~KernelTimer()
	{
		if (e0)
			cudaEventDestroy(e0);
		if (e1)
			cudaEventDestroy(e1);
	}
};

// phase clock: events are only recorded while the query runs; one synchronise at the end
struct PhaseClock {
	mdbcu_ctx *ctx;
	std::vector<cudaEvent_t> ev;
	std::vector<int> ph;
	explicit PhaseClock(mdbcu_ctx *c) : ctx(c) {}
	void begin(int phase)
	{
		cudaEvent_t e;
		cudaEventCreate(&e);
		cudaEventRecord(e, ctx->stream);
		ev.push_back(e);
		ph.push_back(phase);
	}
	// closes the last phase, waits for the stream and fills stats.phase_ms / total_ms
	void finish()
	{
		if (ev.empty())
			return;
		begin(-1);
		cudaEventSynchronize(ev.back());
		for (size_t i = 0; i + 1 < ev.size(); i++) {
			float ms = 0.f;
			cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
			if (ph[i] >= 0 && ph[i] < 8)
				ctx->stats.phase_ms[ph[i]] += ms;
		}
		float total = 0.f;
		cudaEventElapsedTime(&total, ev.front(), ev.back());
		ctx->stats.total_ms = total;
		for (cudaEvent_t e : ev)
			cudaEventDestroy(e);
		ev.clear();
		ph.clear();
	}
	~PhaseClock()
	{
		for (cudaEvent_t e : ev)
			cudaEventDestroy(e);
	}
};
