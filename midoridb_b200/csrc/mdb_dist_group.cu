// mdb_dist_group.cu - GROUP BY / aggregates over ONE sharded table in a distributed plan (MDBCU_PLAN_DISTRIBUTED):
//
//   SELECT g.., COUNT(*), COUNT(x), SUM(x), MIN(x), MAX(x), AVG(x) FROM T [WHERE ..] [GROUP BY g..]
//
// replaces proc_groupby_clause (src/engine/executor_select.c:1526) + handle_countonly_case (:1590) for a table whose rows
// are spread over the ranks.  Any group key: the partial-merge scheme BASELINE.json's north_star names for GROUP BY.
//
//   1. every rank aggregates ITS shard with the general operators (mdb_ops.cu) under a rewritten plan whose aggregates
//      are all decomposable: AVG(x) becomes SUM(x), COUNT(x); everything else is its own partial;
//   2. the partial rows (one per group the shard has seen: key cells, partial cells, NULL flags) are all-gathered - their
//      number depends on the number of GROUPS, not of rows, so this is the cheap end of the exchange;
//   3. rank 0 merges them by key with the same general operators: the gathered partials become a temporary device
//      table, the merge is GROUP BY key with SUM of sums and counts, MIN of mins, MAX of maxs (NULL partials - a group
//      without a non-NULL input on some shard - are skipped by the aggregates exactly as NULL cells are);
//   4. AVG = merged sum / merged count (NULL when the count is zero).
//
// Result placement follows the other aggregate plans (mdb_fast.cu, mdb_star.cu): rank 0 returns the groups, the other
// ranks an empty result.  The reference's rules survive the split: a group exists iff some shard has a qualifying row of
// it, no qualifying row anywhere -> no result row (:1590), NULL group keys collate equal (:1476-1482: the NULL group of every
// shard is one NULL-keyed partial row, and the merge groups NULL keys together again).
// Plain (non-key) columns under GROUP BY carry "the group's first row" in the reference's row order, which has no
// meaning across shards: such plans and joins are not distributed here (MDBCU_EUNSUPPORTED).  HAVING / DISTINCT / ORDER BY /
// LIMIT are applied by mdbcu_select afterwards, on rank 0's complete result (mdb_tail.cu).
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>
#include <vector>

// avg = sum / count; sums arrive as int64 (INT input) or as double bits
__global__ void k_dg_avg(const int64_t *__restrict__ sum, const uint8_t *__restrict__ sum_null, int sum_is_dbl, const int64_t *__restrict__ cnt,
		const uint8_t *__restrict__ cnt_null, uint64_t n, int64_t *__restrict__ out, uint8_t *__restrict__ out_null)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const bool isnull = (sum_null && sum_null[i]) || (cnt_null && cnt_null[i]) || cnt[i] == 0;
		const double s = sum_is_dbl ? __longlong_as_double(sum[i]) : (double)sum[i];
		out[i] = isnull ? 0 : __double_as_longlong(s / (double)cnt[i]);
		out_null[i] = isnull ? 1 : 0;
	}
}

static bool dg_is_key(const mdbcu_plan *plan, const mdbcu_colref &r, int *which)
{
	for (int g = 0; g < plan->n_group; g++)
		if (plan->group[g].tbl == r.tbl && plan->group[g].col == r.col) {
			*which = g;
			return true;
		}
	return false;
}

int mdb_select_general_dist(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (!(plan->flags & MDBCU_PLAN_DISTRIBUTED))
		return MDBCU_EUNSUPPORTED;
	if (plan->n_tables != 1 || plan->n_joins != 0 || plan->n_out < 1)
		return MDBCU_EUNSUPPORTED;
	if (!mdb_comm_ready(ctx))
		return mdb_fail(ctx, MDBCU_EERROR, "MDBCU_PLAN_DISTRIBUTED needs mdbcu_comm_init first");

	// ---- the local plan: decomposable partials only.  where[o] = first partial column of output o
	mdbcu_plan local = *plan;
	local.flags &= ~(uint32_t)MDBCU_PLAN_DISTRIBUTED;
	local.n_out = 0;
	int where[MDBCU_MAX_OUT];
	bool any_agg = false;
	for (int o = 0; o < plan->n_out; o++) {
		const mdbcu_out &out = plan->out[o];
		int g;
		where[o] = local.n_out;
		const int need = out.kind == MDBCU_OUT_AVG ? 2 : 1;
		if (local.n_out + need > MDBCU_MAX_OUT)
			return MDBCU_EUNSUPPORTED;
		switch (out.kind) {
		case MDBCU_OUT_COLUMN:
			if (!dg_is_key(plan, out.ref, &g))
				return MDBCU_EUNSUPPORTED; // "first row of the group" does not survive sharding
			local.out[local.n_out++] = out;
			break;
		case MDBCU_OUT_AVG:
			local.out[local.n_out] = out;
			local.out[local.n_out++].kind = MDBCU_OUT_SUM;
			local.out[local.n_out] = out;
			local.out[local.n_out++].kind = MDBCU_OUT_COUNT_COL;
			any_agg = true;
			break;
		default:
			local.out[local.n_out++] = out;
			any_agg = true;
			break;
		}
	}
	if (!any_agg && plan->n_group == 0)
		return MDBCU_EUNSUPPORTED; // a plain projection: nothing to merge
	// every group key must be in the select list: the partial rows are merged ON those columns (a decision of the plan's
	// shape, taken before any collective so that all ranks take it together)
	for (int g = 0; g < plan->n_group; g++) {
		bool listed = false;
		for (int o = 0; o < plan->n_out; o++) {
			int which;
			listed = listed || (plan->out[o].kind == MDBCU_OUT_COLUMN && dg_is_key(plan, plan->out[o].ref, &which) && which == g);
		}
		if (!listed)
			return MDBCU_EUNSUPPORTED;
	}
	const int W = ctx->world, me = ctx->rank;
	const int nc = local.n_out;

	// ---- 1. this rank's partial groups.  A rank that fails must still take part in the exchange of the row counts, or its
	// peers would wait in the collective: the failure travels as a count of ~0
	mdbcu_result part;
	part.ctx = ctx;
	int local_rc = mdb_select_general(ctx, &local, &part);
	struct PartGuard {
		mdbcu_ctx *ctx;
		mdbcu_result *r;
		~PartGuard()
		{
			for (auto &c : r->cols) {
				mdb_free(ctx, c.cells);
				mdb_free(ctx, c.nulls);
			}
			mdb_free(ctx, r->order_key);
		}
	} part_guard = {ctx, &part};
	ctx->stats.path = MDBCU_PATH_GENERAL;

	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	clock.begin(6);
	uint64_t *d_counts;
	MDB_TRY(tmp.alloc(&d_counts, (size_t)W + 1));
	const uint64_t mine = local_rc == MDBCU_OK ? part.nrows : ~0ull;
	CUDA_TRY(ctx, cudaMemcpyAsync(d_counts + W, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // (`mine` lives on this stack frame)
	MDB_TRY(mdb_comm_allgather_u64(ctx, d_counts + W, d_counts, 1));
	std::vector<uint64_t> counts(W);
	CUDA_TRY(ctx, cudaMemcpyAsync(counts.data(), d_counts, W * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	uint64_t nmax = 0, total = 0;
	for (int r = 0; r < W; r++) {
		if (counts[r] == ~0ull) {
			if (local_rc != MDBCU_OK)
				return local_rc; // (its own message is already in place)
			return mdb_fail(ctx, MDBCU_EERROR, "distributed GROUP BY: rank %d could not aggregate its shard", r);
		}
		nmax = std::max(nmax, counts[r]);
		total += counts[r];
	}
	if (total == 0)
		return mdb_result_alloc(ctx, plan, res, 0, false); // no qualifying row on any shard: no result row (:1590)

	// ---- 2. all-gather the partial rows: per rank one block [column][nmax cells] [column][nmax NULL flags]
	const size_t cells_bytes = (size_t)nmax * sizeof(int64_t), block_bytes = (size_t)nc * (cells_bytes + nmax);
	unsigned char *send, *recv;
	MDB_TRY(tmp.alloc(&send, block_bytes));
	MDB_TRY(tmp.alloc(&recv, block_bytes * W));
	CUDA_TRY(ctx, cudaMemsetAsync(send, 0, block_bytes, ctx->stream));
	for (int c = 0; c < nc && part.nrows; c++) {
		CUDA_TRY(ctx, cudaMemcpyAsync(send + (size_t)c * cells_bytes, part.cols[c].cells, part.nrows * sizeof(int64_t),
				cudaMemcpyDeviceToDevice, ctx->stream));
		if (part.cols[c].nulls)
			CUDA_TRY(ctx, cudaMemcpyAsync(send + (size_t)nc * cells_bytes + (size_t)c * nmax, part.cols[c].nulls, part.nrows,
					cudaMemcpyDeviceToDevice, ctx->stream));
	}
	MDB_TRY(mdb_comm_allgather_bytes(ctx, send, recv, block_bytes));
	ctx->stats.exchange_bytes += (uint64_t)(W - 1) * block_bytes;
	if (me != 0) {
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		clock.finish();
		return mdb_result_alloc(ctx, plan, res, 0, false); // rank 0 returns the groups
	}

	// ---- 3. the partials of all ranks as one temporary table (dense: every rank's rows behind the previous rank's)
	int64_t *cells;
	uint8_t *nulls;
	MDB_TRY(tmp.alloc(&cells, (size_t)nc * total));
	MDB_TRY(tmp.alloc(&nulls, (size_t)nc * total));
	uint64_t at = 0;
	for (int r = 0; r < W; r++) {
		const unsigned char *blk = recv + (size_t)r * block_bytes;
		for (int c = 0; c < nc && counts[r]; c++) {
			CUDA_TRY(ctx, cudaMemcpyAsync(cells + (size_t)c * total + at, blk + (size_t)c * cells_bytes, counts[r] * sizeof(int64_t),
					cudaMemcpyDeviceToDevice, ctx->stream));
			CUDA_TRY(ctx, cudaMemcpyAsync(nulls + (size_t)c * total + at, blk + (size_t)nc * cells_bytes + (size_t)c * nmax, counts[r],
					cudaMemcpyDeviceToDevice, ctx->stream));
		}
		at += counts[r];
	}
	int32_t types[MDBCU_MAX_OUT];
	const void *col_ptrs[MDBCU_MAX_OUT];
	const uint8_t *null_ptrs[MDBCU_MAX_OUT];
	for (int c = 0; c < nc; c++) {
		types[c] = part.cols[c].type == MDBCU_CT_DOUBLE ? MDBCU_CT_DOUBLE : MDBCU_CT_INTEGER;
		col_ptrs[c] = cells + (size_t)c * total;
		null_ptrs[c] = nulls + (size_t)c * total;
	}
	mdbcu_table *pt = nullptr;
	MDB_TRY(mdbcu_table_create(ctx, "partials", nc, types, &pt));
	struct TableGuard {
		mdbcu_table *t;
		~TableGuard() { mdbcu_table_drop(t); }
	} table_guard = {pt};
	MDB_TRY(mdbcu_table_append_columns(pt, total, col_ptrs, null_ptrs)); // (device pointers: the copies are cudaMemcpyDefault)

	// the merge plan: GROUP BY the key columns of the partial table; every partial is merged by its own kind
	clock.begin(4);
	mdbcu_plan merge;
	memset(&merge, 0, sizeof(merge));
	merge.n_tables = 1;
	merge.tables[0] = pt;
	merge.n_group = plan->n_group;
	int key_col[MDBCU_MAX_GROUP];
	for (int g = 0; g < plan->n_group; g++)
		key_col[g] = -1;
	for (int c = 0; c < nc; c++) {
		int g;
		if (local.out[c].kind == MDBCU_OUT_COLUMN && dg_is_key(plan, local.out[c].ref, &g) && key_col[g] < 0)
			key_col[g] = c;
	}
	for (int g = 0; g < plan->n_group; g++) {
		if (key_col[g] < 0)
			return mdb_fail(ctx, MDBCU_EINTERNAL, "distributed GROUP BY: group key %d is missing from the partial rows", g);
		merge.group[g].tbl = 0;
		merge.group[g].col = key_col[g];
	}
	merge.n_out = nc;
	for (int c = 0; c < nc; c++) {
		merge.out[c].ref.tbl = 0;
		merge.out[c].ref.col = c;
		switch (local.out[c].kind) {
		case MDBCU_OUT_COLUMN: {
			int g = 0;
			dg_is_key(plan, local.out[c].ref, &g);
			merge.out[c].kind = MDBCU_OUT_COLUMN;
			merge.out[c].ref.col = key_col[g];
			break;
		}
		case MDBCU_OUT_COUNT_STAR: case MDBCU_OUT_COUNT_COL: case MDBCU_OUT_SUM:
			merge.out[c].kind = MDBCU_OUT_SUM;
			break;
		case MDBCU_OUT_MIN:
			merge.out[c].kind = MDBCU_OUT_MIN;
			break;
		default:
			merge.out[c].kind = MDBCU_OUT_MAX;
			break;
		}
	}
	mdbcu_result merged;
	merged.ctx = ctx;
	PartGuard merged_guard = {ctx, &merged};
	MDB_TRY(mdb_select_general(ctx, &merge, &merged));
	ctx->stats.path = MDBCU_PATH_GENERAL;

	// ---- 4. the result in the plan's own column order: merged columns change hands, AVG is computed
	const uint64_t n = merged.nrows;
	res->nrows = n;
	res->cols.resize(plan->n_out);
	for (int o = 0; o < plan->n_out; o++) {
		ResultColumn &rc = res->cols[o];
		rc.type = out_result_type(plan, o);
		const int c = where[o];
		if (plan->out[o].kind == MDBCU_OUT_AVG) {
			MDB_TRY(mdb_alloc(ctx, &rc.cells, n));
			MDB_TRY(mdb_alloc(ctx, &rc.nulls, n));
			if (n)
				MDB_LAUNCH(ctx, k_dg_avg, (int)std::min<uint64_t>(mdb_div_up(n, 256), (uint64_t)ctx->num_sms * 8), 256, 0,
						(const int64_t*)merged.cols[c].cells, (const uint8_t*)merged.cols[c].nulls, merged.cols[c].type == MDBCU_CT_DOUBLE ? 1 : 0,
						(const int64_t*)merged.cols[c + 1].cells, (const uint8_t*)merged.cols[c + 1].nulls, n, rc.cells, rc.nulls);
		} else {
			rc.cells = merged.cols[c].cells; // (every output has partial columns of its own: each is handed over once)
			rc.nulls = merged.cols[c].nulls;
			merged.cols[c].cells = nullptr;
			merged.cols[c].nulls = nullptr;
		}
	}
	CUDA_CHECK_LAUNCH(ctx);
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // the temporaries below this frame are released on return
	clock.finish();
	return MDBCU_OK;
}
