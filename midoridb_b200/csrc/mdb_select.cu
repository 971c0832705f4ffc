// mdb_select.cu - mdbcu_select dispatch, result materialisation (K8) and statistics.
//
// K8 `k_materialise_pages` replaces table_insert_row on the result table plus proc_select_clause's
// column compaction and table_vacuum (src/primitive/row.c:26, src/engine/executor_select.c:1369,
// src/primitive/vacuum.c:12): columnar result -> page images in the reference's row format, so the
// reference's query_cur_step / query_column_int64 (src/engine/query.c:108,148) read them unchanged.
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>
#include <new>
#include <numeric>

static int validate_plan(mdbcu_ctx *ctx, const mdbcu_plan *plan)
{
	if (!plan)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: plan is NULL");
	if (plan->n_tables < 1 || plan->n_tables > MDBCU_MAX_TABLES)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: %d tables (1..%d supported)", plan->n_tables, MDBCU_MAX_TABLES);
	if (plan->n_joins != plan->n_tables - 1)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: %d tables need %d joins", plan->n_tables, plan->n_tables - 1);
	if (plan->n_out < 1 || plan->n_out > MDBCU_MAX_OUT)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: %d output columns (1..%d supported)", plan->n_out, MDBCU_MAX_OUT);
	for (int t = 0; t < plan->n_tables; t++) {
		if (!plan->tables[t])
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: table %d is NULL", t);
		if (plan->tables[t]->ctx != ctx)
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: table %d belongs to another context", t);
	}
	return MDBCU_OK;
}

static void release_result_buffers(mdbcu_result *r)
{
	for (auto &c : r->cols) {
		mdb_free(r->ctx, c.cells);
		mdb_free(r->ctx, c.nulls);
		c.cells = nullptr;
		c.nulls = nullptr;
	}
	r->cols.clear();
	mdb_free(r->ctx, r->order_key);
	r->order_key = nullptr;
	r->nrows = 0;
}

extern "C" int mdbcu_select(mdbcu_ctx *ctx, const struct mdbcu_plan *plan, mdbcu_result **out)
{
	if (!ctx)
		return MDBCU_EERROR;
	if (!out)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: out is NULL");
	*out = nullptr;
	cudaSetDevice(ctx->device);
	MDB_TRY(validate_plan(ctx, plan));
	MDB_TRY(mdb_validate_tail(ctx, plan));

	mdbcu_result *res = new (std::nothrow) mdbcu_result();
	if (!res)
		return mdb_fail(ctx, MDBCU_ENOMEM, "out of host memory");
	res->ctx = ctx;

	uint64_t total_before = ctx->total_launches;
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	for (int t = 0; t < plan->n_tables; t++)
		ctx->stats.input_rows += plan->tables[t]->n_slots;

	int rc = MDBCU_EUNSUPPORTED;
	if (!(plan->flags & MDBCU_PLAN_NO_FASTPATH)) {
		rc = mdb_select_scan_agg(ctx, plan, res);
		if (rc == MDBCU_EUNSUPPORTED) {
			release_result_buffers(res);
			ctx->radix_gave_up = false;
			rc = mdb_select_radix_joincount(ctx, plan, res);
			if (rc == MDBCU_EUNSUPPORTED) {
				// same plan shape, any key multiplicity / skew / order: one counter per key value (also the distributed answer)
				release_result_buffers(res);
				const bool forced = ctx->radix_gave_up;
				uint64_t keep_rows = ctx->stats.input_rows;
				memset(&ctx->stats, 0, sizeof(ctx->stats));
				ctx->stats.input_rows = keep_rows;
				rc = mdb_select_direct_count(ctx, plan, res, forced);
			}
		}
		if (rc == MDBCU_EUNSUPPORTED) {
			release_result_buffers(res);
			rc = mdb_select_direct_star(ctx, plan, res);
		}
	}
	if (rc == MDBCU_EUNSUPPORTED) {
		if (plan->flags & MDBCU_PLAN_DISTRIBUTED) {
			// GROUP BY / aggregates over one sharded table: local aggregate, partials all-gathered, merged by key on rank 0
			release_result_buffers(res);
			uint64_t keep_rows = ctx->stats.input_rows;
			memset(&ctx->stats, 0, sizeof(ctx->stats));
			ctx->stats.input_rows = keep_rows;
			rc = mdb_select_general_dist(ctx, plan, res);
			if (rc == MDBCU_EUNSUPPORTED)
				rc = mdb_fail(ctx, MDBCU_EUNSUPPORTED, "this plan shape has no distributed implementation (join + GROUP BY join key + "
						"COUNT(*), filter + aggregate scans, small-dimension star joins and GROUP BY over one sharded table have)");
		} else {
			release_result_buffers(res);
			uint64_t keep_rows = ctx->stats.input_rows;
			memset(&ctx->stats, 0, sizeof(ctx->stats));
			ctx->stats.input_rows = keep_rows;
			rc = mdb_select_general(ctx, plan, res);
		}
	}
	if (rc == MDBCU_OK)
		rc = mdb_apply_tail(ctx, plan, res); // HAVING, DISTINCT, ORDER BY, LIMIT on the device-resident result
	ctx->stats.total_kernel_launches = ctx->total_launches;
	ctx->stats.kernel_launches = ctx->total_launches - total_before;
	mdb_scratch_settle(ctx);
	if (rc != MDBCU_OK) {
		cudaStreamSynchronize(ctx->stream);
		cudaGetLastError();
		release_result_buffers(res);
		delete res;
		return rc;
	}
	ctx->stats.result_rows = res->nrows;
	if (!ctx->stats.algorithmic_bytes) {
		// compulsory bytes: every referenced input cell once + every result cell once
		ctx->stats.algorithmic_bytes = 8ull * res->nrows * res->cols.size();
	}
	*out = res;
	return MDBCU_OK;
}

extern "C" int mdbcu_get_stats(mdbcu_ctx *ctx, struct mdbcu_stats *out)
{
	if (!ctx || !out)
		return MDBCU_EERROR;
	*out = ctx->stats;
	out->total_kernel_launches = ctx->total_launches;
	return MDBCU_OK;
}

extern "C" void *mdbcu_host_alloc(mdbcu_ctx *ctx, size_t bytes)
{
	void *p = nullptr;
	if (!ctx)
		return nullptr;
	cudaSetDevice(ctx->device);
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
		cudaGetLastError();
		mdb_fail(ctx, MDBCU_ENOMEM, "cannot page-lock %zu bytes of host memory", bytes);
		return nullptr;
	}
	return p;
}

extern "C" void mdbcu_host_free(mdbcu_ctx *ctx, void *p)
{
	if (ctx && p) {
		cudaSetDevice(ctx->device);
		cudaFreeHost(p);
	}
}

extern "C" int mdbcu_event_record(mdbcu_ctx *ctx, int slot)
{
	if (!ctx || slot < 0 || slot >= MDBCU_EVENT_SLOTS)
		return MDBCU_EERROR;
	cudaSetDevice(ctx->device);
	if (!ctx->user_events[slot])
		CUDA_TRY(ctx, cudaEventCreate(&ctx->user_events[slot]));
	CUDA_TRY(ctx, cudaEventRecord(ctx->user_events[slot], ctx->stream));
	return MDBCU_OK;
}

extern "C" int mdbcu_event_elapsed_ms(mdbcu_ctx *ctx, int a, int b, double *ms)
{
	if (!ctx || !ms || a < 0 || b < 0 || a >= MDBCU_EVENT_SLOTS || b >= MDBCU_EVENT_SLOTS || !ctx->user_events[a] ||
			!ctx->user_events[b])
		return MDBCU_EERROR;
	cudaSetDevice(ctx->device);
	float f = 0.f;
	CUDA_TRY(ctx, cudaEventSynchronize(ctx->user_events[b]));
	CUDA_TRY(ctx, cudaEventElapsedTime(&f, ctx->user_events[a], ctx->user_events[b]));
	*ms = f;
	return MDBCU_OK;
}

// ------------------------------------------------------------------------------------------ results

extern "C" uint64_t mdbcu_result_rows(const mdbcu_result *r)
{
	return r ? r->nrows : 0;
}

extern "C" int mdbcu_result_cols(const mdbcu_result *r)
{
	return r ? (int)r->cols.size() : 0;
}

extern "C" int mdbcu_result_col_type(const mdbcu_result *r, int col)
{
	if (!r || col < 0 || col >= (int)r->cols.size())
		return -1;
	return r->cols[col].type;
}

extern "C" int mdbcu_result_column_device_ptr(mdbcu_result *r, int col, const void **cells, const uint8_t **nulls)
{
	if (!r || !cells || col < 0 || col >= (int)r->cols.size())
		return MDBCU_EERROR;
	cudaSetDevice(r->ctx->device);
	cudaStreamSynchronize(r->ctx->stream);
	*cells = r->cols[col].cells;
	if (nulls)
		*nulls = r->cols[col].nulls;
	return MDBCU_OK;
}

extern "C" void mdbcu_result_free(mdbcu_result *r)
{
	if (!r)
		return;
	cudaSetDevice(r->ctx->device);
	release_result_buffers(r);
	delete r;
}

// results up to this many rows are returned in the reference's row order (left-major nested-loop order,
// groups by first occurrence): executor_select.c:1096-1142, :1542-1582
#define ORDER_LIMIT (1ull << 22)

static int result_permutation(mdbcu_result *r, std::vector<uint32_t> *perm)
{
	mdbcu_ctx *ctx = r->ctx;
	perm->clear();
	if (!r->order_key || r->nrows < 2 || r->nrows > ORDER_LIMIT)
		return MDBCU_OK;
	std::vector<uint64_t> keys(r->nrows);
	CUDA_TRY(ctx, cudaMemcpyAsync(keys.data(), r->order_key, r->nrows * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	bool sorted = true;
	for (uint64_t i = 1; i < r->nrows && sorted; i++)
		sorted = keys[i - 1] <= keys[i];
	if (sorted)
		return MDBCU_OK;
	perm->resize(r->nrows);
	std::iota(perm->begin(), perm->end(), 0u);
	std::stable_sort(perm->begin(), perm->end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
	return MDBCU_OK;
}

extern "C" int mdbcu_result_fetch_columns(mdbcu_result *r, void *const *cells, uint8_t *const *nulls)
{
	if (!r)
		return MDBCU_EERROR;
	mdbcu_ctx *ctx = r->ctx;
	cudaSetDevice(ctx->device);
	if (r->nrows == 0)
		return MDBCU_OK;
	std::vector<uint32_t> perm;
	MDB_TRY(result_permutation(r, &perm));
	std::vector<int64_t> tmp_cells;
	std::vector<uint8_t> tmp_nulls;
	for (size_t c = 0; c < r->cols.size(); c++) {
		if (cells && cells[c]) {
			if (perm.empty()) {
				CUDA_TRY(ctx, cudaMemcpyAsync(cells[c], r->cols[c].cells, r->nrows * 8, cudaMemcpyDeviceToHost, ctx->stream));
			} else {
				tmp_cells.resize(r->nrows);
				CUDA_TRY(ctx, cudaMemcpyAsync(tmp_cells.data(), r->cols[c].cells, r->nrows * 8, cudaMemcpyDeviceToHost, ctx->stream));
				CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
				int64_t *dst = (int64_t*)cells[c];
				for (uint64_t i = 0; i < r->nrows; i++)
					dst[i] = tmp_cells[perm[i]];
			}
		}
		if (nulls && nulls[c]) {
			if (!r->cols[c].nulls) {
				memset(nulls[c], 0, r->nrows);
			} else if (perm.empty()) {
				CUDA_TRY(ctx, cudaMemcpyAsync(nulls[c], r->cols[c].nulls, r->nrows, cudaMemcpyDeviceToHost, ctx->stream));
			} else {
				tmp_nulls.resize(r->nrows);
				CUDA_TRY(ctx, cudaMemcpyAsync(tmp_nulls.data(), r->cols[c].nulls, r->nrows, cudaMemcpyDeviceToHost, ctx->stream));
				CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
				for (uint64_t i = 0; i < r->nrows; i++)
					nulls[c][i] = tmp_nulls[perm[i]];
			}
		}
	}
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return MDBCU_OK;
}

extern "C" size_t mdbcu_result_row_size(const mdbcu_result *r)
{
	return r ? MDBCU_ROW_HEADER + 8 * r->cols.size() : 0; // table_calc_row_size, row.c:21
}

extern "C" size_t mdbcu_result_page_count(const mdbcu_result *r)
{
	if (!r)
		return 0;
	size_t rpp = (MDBCU_PAGE_SIZE - 1) / mdbcu_result_row_size(r);
	size_t n = (size_t)((r->nrows + rpp - 1) / rpp);
	return n ? n : 1; // an empty result is one page of empty slots, never zero pages
}

struct PageCols {
	int ncols;
	const int64_t *cells[MDBCU_MAX_OUT];
	const uint8_t *nulls[MDBCU_MAX_OUT];
};

#define MAT_THREADS 128 // >= floor(4095/32) rows per page

__global__ void __launch_bounds__(MAT_THREADS)
k_materialise_pages(PageCols cols, uint64_t nrows, const uint32_t *__restrict__ perm, int row_size, int rows_per_page,
		uint64_t n_pages, unsigned char *__restrict__ pages)
{
	__shared__ __align__(16) unsigned char page[MDBCU_PAGE_SIZE];
	const int slots = MDBCU_PAGE_SIZE / row_size; // table_datablock_init covers every slot, table.c:124-132

	for (uint64_t p = blockIdx.x; p < n_pages; p += gridDim.x) {
		for (int i = threadIdx.x; i < MDBCU_PAGE_SIZE / 16; i += blockDim.x)
			reinterpret_cast<int4*>(page)[i] = make_int4(0, 0, 0, 0);
		__syncthreads();
		for (int s = threadIdx.x; s < slots; s += blockDim.x) {
			unsigned char *row = page + (size_t)s * row_size;
			uint64_t i = p * (uint64_t)rows_per_page + s;
			if (s < rows_per_page && i < nrows) {
				uint64_t src = perm ? perm[i] : i;
				// flags.empty = flags.deleted = false (row.c:62-63); bit c set <=> column c is NULL (row.c:64)
				for (int c = 0; c < cols.ncols; c++) {
					bool isnull = cols.nulls[c] && cols.nulls[c][src];
					if (isnull)
						row[MDBCU_NULL_BITMAP_OFF + (c >> 3)] |= (unsigned char)(1u << (c & 7));
					else
						*reinterpret_cast<int64_t*>(row + MDBCU_ROW_HEADER + 8 * c) = cols.cells[c][src];
				}
			} else {
				row[0] = 1; // flags.empty
			}
		}
		__syncthreads();
		int4 *dst = reinterpret_cast<int4*>(pages + p * MDBCU_PAGE_SIZE);
		for (int i = threadIdx.x; i < MDBCU_PAGE_SIZE / 16; i += blockDim.x)
			dst[i] = reinterpret_cast<int4*>(page)[i];
		__syncthreads();
	}
}

extern "C" int mdbcu_result_fetch_pages(mdbcu_result *r, void *pages, size_t n_pages)
{
	if (!r)
		return MDBCU_EERROR;
	mdbcu_ctx *ctx = r->ctx;
	cudaSetDevice(ctx->device);
	size_t need = mdbcu_result_page_count(r);
	if (!pages || n_pages < need)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_result_fetch_pages: need room for %zu pages", need);
	if (r->cols.size() > MDBCU_MAX_OUT)
		return mdb_fail(ctx, MDBCU_EINTERNAL, "result has too many columns");

	std::vector<uint32_t> perm;
	MDB_TRY(result_permutation(r, &perm));

	DevTemp tmp(ctx);
	unsigned char *d_pages;
	uint32_t *d_perm = nullptr;
	MDB_TRY(tmp.alloc(&d_pages, need * MDBCU_PAGE_SIZE));
	if (!perm.empty()) {
		MDB_TRY(tmp.alloc(&d_perm, perm.size()));
		CUDA_TRY(ctx, cudaMemcpyAsync(d_perm, perm.data(), perm.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
	}
	PageCols pc;
	memset(&pc, 0, sizeof(pc));
	pc.ncols = (int)r->cols.size();
	for (int c = 0; c < pc.ncols; c++) {
		pc.cells[c] = r->cols[c].cells;
		pc.nulls[c] = r->cols[c].nulls;
	}
	int row_size = (int)mdbcu_result_row_size(r);
	int rpp = (MDBCU_PAGE_SIZE - 1) / row_size;
	int grid = (int)std::min<size_t>(need, (size_t)ctx->num_sms * 16);
	MDB_LAUNCH(ctx, k_materialise_pages, grid, MAT_THREADS, 0, pc, r->nrows, (const uint32_t*)d_perm, row_size, rpp,
			(uint64_t)need, d_pages);
	CUDA_CHECK_LAUNCH(ctx);
	CUDA_TRY(ctx, cudaMemcpyAsync(pages, d_pages, need * MDBCU_PAGE_SIZE, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return MDBCU_OK;
}
