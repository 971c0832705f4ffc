// mdb_direct.cu - direct-count join + GROUP BY join key + COUNT(*): the README query when keys repeat without bound
//
//   SELECT k, COUNT(*) FROM A INNER JOIN B ON A.k = B.k GROUP BY k        result(k) = cntA[k] * cntB[k]
//
// replaces _join_nested_loop_tbl2tbl (src/engine/executor_select.c:1076) + proc_groupby_clause (:1526) for the inputs the
// radix path (mdb_radix.cu) hands back: its packed 4- / 8-bit counters wrap when one key occurs more than 255 times on a
// side, its per-partition streams overflow when one partition receives more than twice its share, and the general
// operators would have to materialise every joined pair (a Zipf(1.1) key on both sides of 2^20 x 2^20 rows joins to
// 3*10^10 pairs).  Here both sides are counted into one 32-bit counter per key value of the common range:
//
//   k_dc_count   streams the 8-byte keys once (256-bit loads); equal keys inside a warp are combined first
//                (match.any), one global reduction per distinct key and warp - a hot key costs one RED per warp,
//                not one per row;
//   (reduce)     multi-GPU plans: every rank counts ITS shard of both sides, the counters of the key range a rank owns
//                are summed over the ranks (mdb_comm_reduce_owned_u32) - skew cannot unbalance the exchange, the
//                bytes moved depend on the key range only;
//   k_dc_emit    one pass over the (owned) counters, groups written with block-level compaction.
//
// HBM-bound integer work: 8 B per key in, 8 B of counters per key VALUE read once, 16 B per group out.  The random
// counter updates stay in L2 while the range fits it (2^24 key values per side); beyond that they cost a DRAM sector
// each, which is why this is the fall-back and not the headline path.
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>

#define DC_MAX_RANGE (1ull << 30) // key values per side and window (4 GiB of counters each)
#define DC_MAX_WINDOWS 16         // single-GPU plans: wider key ranges are counted window by window (every window reads the keys again)
#define DC_THREADS 256

struct DCSide {
	const int64_t *keys;
	const uint32_t *present; // nullptr: every slot is live and not NULL
	uint64_t n;
};

__device__ __forceinline__ void dc_load256(const void *p, uint32_t *w)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			: "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

// one key of every lane (ok = the lane has one): lanes with equal keys elect a leader that adds their number
__device__ __forceinline__ void dc_add(uint32_t *__restrict__ cnt, unsigned long long d, bool ok)
{
	const uint32_t active = __ballot_sync(0xffffffffu, ok);
	if (!ok)
		return;
	const uint32_t same = __match_any_sync(active, d);
	if ((threadIdx.x & 31u) == (uint32_t)(__ffs(same) - 1))
		atomicAdd(&cnt[d], (uint32_t)__popc(same));
}

__global__ void __launch_bounds__(DC_THREADS) k_dc_count(DCSide s, long long kmin, unsigned long long range, uint32_t *__restrict__ cnt)
{
	const uint64_t tid = blockIdx.x * (uint64_t)DC_THREADS + threadIdx.x, nthreads = (uint64_t)gridDim.x * DC_THREADS;
	const bool aligned = ((uintptr_t)s.keys & 31u) == 0;
	const uint64_t nquads = aligned ? s.n / 4 : 0;
	// whole warps stay in the loop together (dc_add votes): the trip count is rounded up to the warp
	const uint64_t quad_iters = (nquads + nthreads - 1) / nthreads;
	for (uint64_t it = 0; it < quad_iters; it++) {
		const uint64_t q = it * nthreads + tid;
		const bool have = q < nquads;
		uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		uint32_t pres = 0xfu;
		if (have) {
			dc_load256(s.keys + q * 4, w);
			if (s.present)
				pres = (s.present[q >> 3] >> ((q & 7u) * 4u)) & 0xfu;
		}
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const unsigned long long key = ((unsigned long long)w[2 * j + 1] << 32) | w[2 * j];
			const unsigned long long d = key - (unsigned long long)kmin;
			dc_add(cnt, d, have && ((pres >> j) & 1u) && d < range);
		}
	}
	const uint64_t first = nquads * 4, rest = s.n - first;
	const uint64_t rest_iters = (rest + nthreads - 1) / nthreads;
	for (uint64_t it = 0; it < rest_iters; it++) {
		const uint64_t r = first + it * nthreads + tid;
		const bool have = r < s.n;
		unsigned long long d = 0;
		bool ok = false;
		if (have) {
			d = (unsigned long long)s.keys[r] - (unsigned long long)kmin;
			ok = d < range && (!s.present || mdb_bit(s.present, r));
		}
		dc_add(cnt, d, ok);
	}
}

struct DCOut {
	int nout;
	int is_count[4];
	int64_t *cells[4];
	unsigned long long *cursor; // groups emitted so far
	uint64_t cap;
};

// counters [first, end) of both sides -> one group per key value present on both; every thread takes 4 consecutive values
__global__ void __launch_bounds__(DC_THREADS) k_dc_emit(const uint32_t *__restrict__ ca, const uint32_t *__restrict__ cb, uint64_t first,
		uint64_t end, long long kmin, DCOut out)
{
	__shared__ uint32_t s_warp[DC_THREADS / 32];
	__shared__ unsigned long long s_base;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint64_t per_block = (uint64_t)DC_THREADS * 4;
	for (uint64_t b0 = first + blockIdx.x * per_block; b0 < end; b0 += (uint64_t)gridDim.x * per_block) {
		const uint64_t v0 = b0 + threadIdx.x * 4ull;
		uint32_t a[4], b[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			a[j] = v0 + j < end ? ca[v0 + j] : 0u;
			b[j] = v0 + j < end ? cb[v0 + j] : 0u;
		}
		uint32_t mine = 0;
#pragma unroll
		for (int j = 0; j < 4; j++)
			mine += (a[j] != 0 && b[j] != 0) ? 1u : 0u;
		uint32_t incl = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o)
				incl += v;
		}
		if (lane == 31)
			s_warp[warp] = incl;
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t total = 0;
			for (int w = 0; w < DC_THREADS / 32; w++) {
				const uint32_t t = s_warp[w];
				s_warp[w] = total;
				total += t;
			}
			s_base = total ? atomicAdd(out.cursor, (unsigned long long)total) : 0ull;
		}
		__syncthreads();
		unsigned long long row = s_base + s_warp[warp] + (incl - mine);
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (a[j] == 0 || b[j] == 0)
				continue;
			if (row < out.cap) {
				const long long key = kmin + (long long)(v0 + j);
				const long long c = (long long)((unsigned long long)a[j] * (unsigned long long)b[j]);
#pragma unroll
				for (int o = 0; o < 4; o++)
					if (o < out.nout)
						out.cells[o][row] = out.is_count[o] ? c : key;
			}
			row++;
		}
		__syncthreads();
	}
}

static bool dc_col_all_present(const mdbcu_table *t, int col)
{
	return t->all_live && !t->cols[col].has_nulls;
}

// Same plan shape as mdb_select_radix_joincount.  `forced`: the radix path gave up on this query's DATA (counter wrapped,
// stream overflow, skew) - without it only distributed plans come here on their own (small single-GPU inputs belong to the
// general operators, which also return the reference's row order).
int mdb_select_direct_count(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res, bool forced)
{
	if (plan->n_tables != 2 || plan->n_joins != 1 || plan->joins[0].cross || plan->n_pred != 0 || plan->n_group != 1 ||
			plan->n_out < 1 || plan->n_out > 4)
		return MDBCU_EUNSUPPORTED;
	const bool dist = (plan->flags & MDBCU_PLAN_DISTRIBUTED) != 0;
	if (!forced && !dist)
		return MDBCU_EUNSUPPORTED;
	if (dist && !mdb_comm_ready(ctx))
		return mdb_fail(ctx, MDBCU_EERROR, "MDBCU_PLAN_DISTRIBUTED needs mdbcu_comm_init first");
	const mdbcu_join &jn = plan->joins[0];
	if (jn.left.tbl != 0 || jn.right.tbl != 1)
		return MDBCU_EUNSUPPORTED;
	const mdbcu_table *ta = plan->tables[0], *tb = plan->tables[1];
	if (jn.left.col < 0 || jn.left.col >= ta->ncols || jn.right.col < 0 || jn.right.col >= tb->ncols)
		return MDBCU_EUNSUPPORTED;
	DevColumn ca = ta->cols[jn.left.col], cb = tb->cols[jn.right.col];
	auto intlike = [](int type) { return type == MDBCU_CT_INTEGER || type == MDBCU_CT_DATE || type == MDBCU_CT_DATETIME; };
	if (!intlike(ca.type) || !intlike(cb.type) || !ca.stats_ok || !cb.stats_ok)
		return MDBCU_EUNSUPPORTED;
	if (dist) {
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
	if (ta->n_slots >= (1ull << 32) || tb->n_slots >= (1ull << 32))
		return MDBCU_EUNSUPPORTED; // 32-bit counters
	const long long kmin = std::max(ca.imin, cb.imin), kmax = std::min(ca.imax, cb.imax);
	ctx->stats.path = MDBCU_PATH_DIRECT_COUNT;
	if (ca.imin > ca.imax || cb.imin > cb.imax || kmin > kmax)
		return mdb_result_alloc(ctx, plan, res, 0, false);
	const unsigned long long full_range = (unsigned long long)kmax - (unsigned long long)kmin + 1ull;
	if (full_range == 0 || full_range > DC_MAX_RANGE * (dist ? 1ull : (unsigned long long)DC_MAX_WINDOWS))
		return MDBCU_EUNSUPPORTED;
	// Single-GPU plans count key ranges beyond 2^30 values in WINDOWS of 2^30: per window the keys are streamed again (keys
	// outside the window are skipped by the range test of k_dc_count), the counters are reused and the groups appended to
	// the same result.  32-bit keys spread over [0, 2^32) - four windows - stay off the general operators, which would
	// build a hash table over every row.
	const unsigned long long range = std::min<unsigned long long>(full_range, DC_MAX_RANGE);
	const int n_windows = (int)((full_range + DC_MAX_RANGE - 1) / DC_MAX_RANGE);

	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	const int W = dist ? ctx->world : 1, me = dist ? ctx->rank : 0;
	if (mdb_trace_level() >= 1)
		fprintf(stderr, "[mdbcu] rank %d: direct count over %llu key values in %d window(s) (%s)\n", me, full_range, n_windows,
				forced ? "handed over by the radix join" : "distributed plan");
	uint32_t *cnt; // [side A: range][side B: range]
	unsigned long long *d_cursor;
	MDB_TRY(tmp.alloc(&cnt, 2 * (size_t)range));
	MDB_TRY(tmp.alloc(&d_cursor, 1));
	CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, 2 * (size_t)range * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));

	clock.begin(2);
	const mdbcu_table *tabs[2] = {ta, tb};
	const int cols[2] = {jn.left.col, jn.right.col};
	auto count_window = [&](long long wmin, unsigned long long wrange) -> int {
		for (int side = 0; side < 2; side++) {
			DCSide s;
			s.keys = tabs[side]->cols[cols[side]].data;
			s.present = dc_col_all_present(tabs[side], cols[side]) ? nullptr : tabs[side]->cols[cols[side]].present;
			s.n = tabs[side]->n_slots;
			if (s.n == 0)
				continue;
			const int grid = (int)std::min<uint64_t>(mdb_div_up(s.n, (size_t)DC_THREADS * 4), (uint64_t)ctx->num_sms * 8);
			MDB_LAUNCH(ctx, k_dc_count, grid, DC_THREADS, 0, s, wmin, wrange, cnt + (size_t)side * range);
		}
		CUDA_CHECK_LAUNCH(ctx);
		return MDBCU_OK;
	};
	MDB_TRY(count_window(kmin, range));

	// this rank's slice of the key range (all of it on one GPU)
	uint64_t own_first = 0, own_end = range;
	if (W > 1) {
		own_first = (uint64_t)((unsigned __int128)range * me / W);
		own_end = (uint64_t)((unsigned __int128)range * (me + 1) / W);
		clock.begin(6);
		// both sides in one call: slice r of side A and of side B end up summed on rank r
		MDB_TRY(mdb_comm_reduce_owned_u32(ctx, cnt, range, 2, &ctx->stats.exchange_bytes));
	}

	clock.begin(3);
	uint64_t cap_groups = n_windows > 1 ? full_range : own_end - own_first;
	if (!dist)
		cap_groups = std::min<uint64_t>(cap_groups, std::min<uint64_t>(ta->n_slots, tb->n_slots));
	MDB_TRY(mdb_result_alloc(ctx, plan, res, 0, false));
	DCOut out;
	memset(&out, 0, sizeof(out));
	out.nout = plan->n_out;
	out.cursor = d_cursor;
	out.cap = cap_groups;
	for (int o = 0; o < plan->n_out; o++) {
		mdb_free(ctx, res->cols[o].cells);
		mdb_free(ctx, res->cols[o].nulls);
		res->cols[o].nulls = nullptr; // NULL keys never join (executor_select.c:716-738)
		res->cols[o].cells = nullptr;
		MDB_TRY(mdb_alloc(ctx, &res->cols[o].cells, cap_groups));
		out.cells[o] = res->cols[o].cells;
		out.is_count[o] = plan->out[o].kind == MDBCU_OUT_COUNT_STAR;
	}
	if (own_end > own_first) {
		const int grid = (int)std::min<uint64_t>(mdb_div_up(own_end - own_first, (size_t)DC_THREADS * 4), (uint64_t)ctx->num_sms * 8);
		MDB_LAUNCH(ctx, k_dc_emit, grid, DC_THREADS, 0, cnt, cnt + range, own_first, own_end, kmin, out);
		CUDA_CHECK_LAUNCH(ctx);
	}
	// the further windows of a wide key range (single-GPU plans): same counters, same result columns
	for (int w = 1; w < n_windows; w++) {
		const long long wmin = (long long)((unsigned long long)kmin + (unsigned long long)w * DC_MAX_RANGE);
		const unsigned long long wrange = std::min<unsigned long long>(DC_MAX_RANGE, full_range - (unsigned long long)w * DC_MAX_RANGE);
		clock.begin(2);
		CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, 2 * (size_t)range * sizeof(uint32_t), ctx->stream));
		MDB_TRY(count_window(wmin, wrange));
		clock.begin(3);
		const int grid = (int)std::min<uint64_t>(mdb_div_up(wrange, (size_t)DC_THREADS * 4), (uint64_t)ctx->num_sms * 8);
		MDB_LAUNCH(ctx, k_dc_emit, grid, DC_THREADS, 0, cnt, cnt + range, (uint64_t)0, (uint64_t)wrange, wmin, out);
		CUDA_CHECK_LAUNCH(ctx);
	}
	uint64_t ngroups = 0;
	MDB_TRY(mdb_read_u64(ctx, (const uint64_t*)d_cursor, &ngroups));
	clock.finish();
	if (ngroups > cap_groups)
		return mdb_fail(ctx, MDBCU_EINTERNAL, "direct count emitted %llu groups into %llu rows", (unsigned long long)ngroups,
				(unsigned long long)cap_groups);
	res->nrows = ngroups;
	ctx->stats.algorithmic_bytes = 8ull * (ta->n_slots + tb->n_slots) + 8ull * plan->n_out * ngroups;
	ctx->stats.dominant_ms = ctx->stats.phase_ms[2];
	ctx->stats.dominant_bytes = 8ull * (ta->n_slots + tb->n_slots);
	return MDBCU_OK;
}
