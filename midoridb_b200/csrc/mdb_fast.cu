// mdb_fast.cu - fused kernels for the benchmark shapes of the hot path (BASELINE.json configs).
//
//  K2 scan_filter_aggregate  : SELECT <aggregates> FROM T WHERE <range conjunction>   (config 2)
//       replaces proc_from_clause_table :1282 + proc_where_clause :1435 + handle_countonly_case :1590
//       of src/engine/executor_select.c with ONE pass over the referenced columns.
//  K7/K3/K4 radix join+count : SELECT k, COUNT(*) FROM A INNER JOIN B ON A.k = B.k GROUP BY k   (README query,
//       configs 1 and 3) replaces _join_nested_loop_tbl2tbl :1076 + proc_groupby_clause :1526.
//       result(k) = cntA[k] * cntB[k]; no pair is ever materialised.
//       pass 1 (k_radix_partition): stream the 8-byte keys once, write 2-byte remainders into
//               per-partition chunk lists (partition = high bits of key - kmin);
//       pass 2 (k_radix_joincount): per partition, two byte-counter histograms in shared memory,
//               multiply, emit (key, count) groups.
//  HBM-bound integer work throughout: no tensor cores (nothing here is a dense contraction).
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>

static bool col_all_present(const mdbcu_table *t, int col)
{
	return t->all_live && !t->cols[col].has_nulls;
}

// ===================================================================================== K2 scan + aggregate

#define SA_MAX_COLS 3
#define SA_MAX_AGGS 8
#define SA_THREADS 256

struct SACol {
	const int64_t *data;
	const uint32_t *present; // nullptr: all present
	int32_t is_dbl;
	int32_t has_range;
	long long ilo, ihi;      // inclusive integer range
	double dlo, dhi;
	int32_t dlo_incl, dhi_incl;
};

struct SAAgg {
	int32_t kind;
	int32_t col; // index into cols, -1 for COUNT(*)
};

struct SASpec {
	int32_t ncols, naggs;
	SACol cols[SA_MAX_COLS];
	SAAgg aggs[SA_MAX_AGGS];
};

// accumulators are 8-byte words: integer sum / ordered min / ordered max, or the bits of a double sum
struct SAPartial {
	unsigned long long rows; // qualifying rows
	unsigned long long nn[SA_MAX_AGGS];
	long long acc[SA_MAX_AGGS];
};

template <int NA>
struct SAState {
	unsigned long long rows;
	unsigned long long nn[NA];
	long long acc[NA];
};

__device__ static inline bool sa_is_dsum(const SASpec &sp, int a)
{
	int kind = sp.aggs[a].kind;
	return (kind == MDBCU_OUT_SUM || kind == MDBCU_OUT_AVG) && sp.aggs[a].col >= 0 && sp.cols[sp.aggs[a].col].is_dbl;
}

template <int NA>
__device__ static inline void sa_init(const SASpec &sp, SAState<NA> &st)
{
	st.rows = 0;
#pragma unroll
	for (int a = 0; a < NA; a++) {
		st.nn[a] = 0;
		st.acc[a] = 0; // also +0.0
		if (sp.aggs[a].kind == MDBCU_OUT_MIN)
			st.acc[a] = INT64_MAX;
		else if (sp.aggs[a].kind == MDBCU_OUT_MAX)
			st.acc[a] = INT64_MIN;
	}
}

template <int NA>
__device__ static inline void sa_merge(const SASpec &sp, SAState<NA> &st, unsigned long long rows, const unsigned long long *nn,
		const long long *acc)
{
	st.rows += rows;
#pragma unroll
	for (int a = 0; a < NA; a++) {
		st.nn[a] += nn[a];
		int kind = sp.aggs[a].kind;
		if (kind == MDBCU_OUT_MIN)
			st.acc[a] = min(st.acc[a], acc[a]);
		else if (kind == MDBCU_OUT_MAX)
			st.acc[a] = max(st.acc[a], acc[a]);
		else if (sa_is_dsum(sp, a))
			st.acc[a] = __double_as_longlong(__longlong_as_double(st.acc[a]) + __longlong_as_double(acc[a]));
		else
			st.acc[a] = (long long)((unsigned long long)st.acc[a] + (unsigned long long)acc[a]);
	}
}

template <int NC, int NA>
__device__ static inline void sa_row(const SASpec &sp, SAState<NA> &st, const long long *v, const bool *pres)
{
	bool ok = true;
#pragma unroll
	for (int c = 0; c < NC; c++) {
		const SACol &col = sp.cols[c];
		if (!col.has_range)
			continue;
		if (col.is_dbl) {
			double d = __longlong_as_double(v[c]);
			bool lo = col.dlo_incl ? d >= col.dlo : d > col.dlo;
			bool hi = col.dhi_incl ? d <= col.dhi : d < col.dhi;
			ok = ok && pres[c] && lo && hi;
		} else {
			ok = ok && pres[c] && v[c] >= col.ilo && v[c] <= col.ihi;
		}
	}
	if (!ok)
		return;
	st.rows++;
#pragma unroll
	for (int a = 0; a < NA; a++) {
		int kind = sp.aggs[a].kind, c = sp.aggs[a].col;
		if (kind == MDBCU_OUT_COUNT_STAR) {
			st.nn[a]++;
			continue;
		}
		long long x = 0;
		bool p = false, dbl = false;
#pragma unroll
		for (int cc = 0; cc < NC; cc++) {
			if (cc == c) {
				x = v[cc];
				p = pres[cc];
				dbl = sp.cols[cc].is_dbl;
			}
		}
		if (!p)
			continue;
		st.nn[a]++;
		switch (kind) {
		case MDBCU_OUT_SUM: case MDBCU_OUT_AVG:
			if (dbl)
				st.acc[a] = __double_as_longlong(__longlong_as_double(st.acc[a]) + __longlong_as_double(x));
			else
				st.acc[a] = (long long)((unsigned long long)st.acc[a] + (unsigned long long)x);
			break;
		case MDBCU_OUT_MIN:
			st.acc[a] = min(st.acc[a], dbl ? mdb_dbl_to_ordered(x) : x);
			break;
		case MDBCU_OUT_MAX:
			st.acc[a] = max(st.acc[a], dbl ? mdb_dbl_to_ordered(x) : x);
			break;
		}
	}
}

// Each thread streams pairs of rows with one 128-bit load per referenced column (2 x 8-byte cells),
// four pairs in flight; per-thread partials -> warp shuffles -> block -> one partial per block.
// A second, single-block kernel folds the block partials in a fixed order (deterministic doubles).
template <int NC, int NA>
__global__ void __launch_bounds__(SA_THREADS, 2) k_scan_filter_aggregate(SASpec sp, uint64_t n, SAPartial *__restrict__ partials)
{
	SAState<NA> st;
	sa_init<NA>(sp, st);

	const uint64_t npairs = n / 2;
	const uint64_t stride = (uint64_t)gridDim.x * SA_THREADS;
	uint64_t pair = (uint64_t)blockIdx.x * SA_THREADS + threadIdx.x;

	constexpr int UNROLL = 4;
	for (; pair + (UNROLL - 1) * stride < npairs; pair += UNROLL * stride) {
		int4 raw[UNROLL][NC];
		uint32_t pw[UNROLL][NC];
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			uint64_t pi = pair + u * stride;
#pragma unroll
			for (int c = 0; c < NC; c++) {
				raw[u][c] = mdb_ldg_stream(reinterpret_cast<const int4*>(sp.cols[c].data) + pi);
				pw[u][c] = sp.cols[c].present ? sp.cols[c].present[pi >> 4] : 0xffffffffu;
			}
		}
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			uint64_t pi = pair + u * stride;
			int bit = (int)((pi & 15) * 2);
			long long v0[NC], v1[NC];
			bool p0[NC], p1[NC];
#pragma unroll
			for (int c = 0; c < NC; c++) {
				v0[c] = (long long)(((unsigned long long)(unsigned)raw[u][c].y << 32) | (unsigned)raw[u][c].x);
				v1[c] = (long long)(((unsigned long long)(unsigned)raw[u][c].w << 32) | (unsigned)raw[u][c].z);
				p0[c] = (pw[u][c] >> bit) & 1;
				p1[c] = (pw[u][c] >> (bit + 1)) & 1;
			}
			sa_row<NC, NA>(sp, st, v0, p0);
			sa_row<NC, NA>(sp, st, v1, p1);
		}
	}
	// remainder: single rows
	for (uint64_t r = pair * 2; r < n; r += 2 * stride) {
		for (int k = 0; k < 2 && r + k < n; k++) {
			long long v[NC];
			bool p[NC];
#pragma unroll
			for (int c = 0; c < NC; c++) {
				v[c] = sp.cols[c].data[r + k];
				p[c] = sp.cols[c].present ? mdb_bit(sp.cols[c].present, r + k) : true;
			}
			sa_row<NC, NA>(sp, st, v, p);
		}
	}

	// warp reduce (fixed butterfly order)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		unsigned long long rows = __shfl_xor_sync(0xffffffffu, st.rows, o);
		unsigned long long nn[NA];
		long long ac[NA];
#pragma unroll
		for (int a = 0; a < NA; a++) {
			nn[a] = __shfl_xor_sync(0xffffffffu, st.nn[a], o);
			ac[a] = __shfl_xor_sync(0xffffffffu, st.acc[a], o);
		}
		sa_merge<NA>(sp, st, rows, nn, ac);
	}
	__shared__ SAPartial warp_part[SA_THREADS / 32];
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) {
		warp_part[warp].rows = st.rows;
#pragma unroll
		for (int a = 0; a < NA; a++) {
			warp_part[warp].nn[a] = st.nn[a];
			warp_part[warp].acc[a] = st.acc[a];
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		SAState<NA> tot;
		sa_init<NA>(sp, tot);
		for (int w = 0; w < SA_THREADS / 32; w++)
			sa_merge<NA>(sp, tot, warp_part[w].rows, warp_part[w].nn, warp_part[w].acc);
		SAPartial &out = partials[blockIdx.x];
		out.rows = tot.rows;
#pragma unroll
		for (int a = 0; a < NA; a++) {
			out.nn[a] = tot.nn[a];
			out.acc[a] = tot.acc[a];
		}
	}
}

struct SAOut {
	int64_t *cells[SA_MAX_AGGS];
	uint8_t *nulls[SA_MAX_AGGS];
};

__global__ void k_scan_aggregate_final(SASpec sp, const SAPartial *__restrict__ partials, int nparts, SAOut out,
		unsigned long long *__restrict__ d_rows)
{
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	SAState<SA_MAX_AGGS> tot;
	// only the first naggs entries are meaningful; the others fold zeros
	SASpec full = sp;
	for (int a = sp.naggs; a < SA_MAX_AGGS; a++) {
		full.aggs[a].kind = MDBCU_OUT_COUNT_STAR;
		full.aggs[a].col = -1;
	}
	sa_init<SA_MAX_AGGS>(full, tot);
	for (int b = 0; b < nparts; b++)
		sa_merge<SA_MAX_AGGS>(full, tot, partials[b].rows, partials[b].nn, partials[b].acc);
	*d_rows = tot.rows;
	for (int a = 0; a < sp.naggs; a++) {
		int kind = sp.aggs[a].kind;
		bool dbl = sp.aggs[a].col >= 0 && sp.cols[sp.aggs[a].col].is_dbl;
		long long cell = 0;
		bool isnull = false;
		switch (kind) {
		case MDBCU_OUT_COUNT_STAR: case MDBCU_OUT_COUNT_COL:
			cell = (long long)tot.nn[a];
			break;
		case MDBCU_OUT_SUM:
			isnull = tot.nn[a] == 0;
			cell = tot.acc[a];
			break;
		case MDBCU_OUT_MIN: case MDBCU_OUT_MAX:
			isnull = tot.nn[a] == 0;
			cell = dbl ? mdb_ordered_to_dbl(tot.acc[a]) : tot.acc[a];
			break;
		case MDBCU_OUT_AVG:
			isnull = tot.nn[a] == 0;
			if (!isnull) {
				double sum = dbl ? __longlong_as_double(tot.acc[a]) : (double)tot.acc[a];
				cell = __double_as_longlong(sum / (double)tot.nn[a]);
			}
			break;
		}
		out.cells[a][0] = isnull ? 0 : cell;
		out.nulls[a][0] = isnull;
	}
}

template <int NC>
static void launch_scan_agg(mdbcu_ctx *ctx, int grid, const SASpec &sp, uint64_t n, SAPartial *partials)
{
	// aggregate slots beyond naggs are padded with COUNT(*) by the caller so NA can be rounded up
	if (sp.naggs <= 1)
		MDB_LAUNCH(ctx, (k_scan_filter_aggregate<NC, 1>), grid, SA_THREADS, 0, sp, n, partials);
	else if (sp.naggs <= 2)
		MDB_LAUNCH(ctx, (k_scan_filter_aggregate<NC, 2>), grid, SA_THREADS, 0, sp, n, partials);
	else if (sp.naggs <= 4)
		MDB_LAUNCH(ctx, (k_scan_filter_aggregate<NC, 4>), grid, SA_THREADS, 0, sp, n, partials);
	else
		MDB_LAUNCH(ctx, (k_scan_filter_aggregate<NC, 8>), grid, SA_THREADS, 0, sp, n, partials);
}

// turn the WHERE program into per-column inclusive ranges; false when the shape is anything but
// a conjunction of <column> <cmp> <literal> (either operand order, the "yoda" form of test_select_7)
static bool pred_to_ranges(const mdbcu_plan *plan, SASpec *sp, int *col_of /* table col -> spec col */)
{
	struct Item {
		int kind; // 0 column, 1 int literal, 2 double literal, 3 boolean (comparison result)
		int col;
		long long i;
		double d;
	};
	Item st[MDBCU_MAX_PRED];
	int depth = 0;
	const mdbcu_table *t = plan->tables[0];

	for (int k = 0; k < plan->n_pred; k++) {
		const mdbcu_pred_op &op = plan->pred[k];
		Item it{};
		switch (op.op) {
		case MDBCU_P_COL:
			if (op.tbl != 0 || op.col < 0 || op.col >= t->ncols || t->cols[op.col].type == MDBCU_CT_VARCHAR)
				return false;
			it.kind = 0;
			it.col = op.col;
			st[depth++] = it;
			break;
		case MDBCU_P_INT:
			it.kind = 1;
			it.i = op.ival;
			st[depth++] = it;
			break;
		case MDBCU_P_DBL:
			it.kind = 2;
			it.d = op.dval;
			st[depth++] = it;
			break;
		case MDBCU_P_CMP: {
			if (depth < 2)
				return false;
			Item b = st[--depth], a = st[--depth];
			int cmp = op.arg;
			if (a.kind != 0 && b.kind == 0) {
				// literal <cmp> column  ==  column <mirrored cmp> literal
				std::swap(a, b);
				cmp = cmp == 1 ? 2 : cmp == 2 ? 1 : cmp == 5 ? 6 : cmp == 6 ? 5 : cmp;
			}
			if (a.kind != 0 || b.kind == 0 || b.kind == 3 || cmp == 3)
				return false;
			bool dbl = t->cols[a.col].type == MDBCU_CT_DOUBLE;
			if (!dbl && b.kind == 2)
				return false; // INT column vs FLOAT literal: leave to the general operators
			int sc = col_of[a.col];
			if (sc < 0) {
				if (sp->ncols == SA_MAX_COLS)
					return false;
				sc = sp->ncols++;
				col_of[a.col] = sc;
				SACol &c = sp->cols[sc];
				memset(&c, 0, sizeof(c));
				c.data = t->cols[a.col].data;
				c.present = col_all_present(t, a.col) ? nullptr : t->cols[a.col].present;
				c.is_dbl = dbl;
			}
			SACol &c = sp->cols[sc];
			if (!c.has_range) {
				c.has_range = 1;
				c.ilo = INT64_MIN;
				c.ihi = INT64_MAX;
				c.dlo = -__builtin_inf();
				c.dhi = __builtin_inf();
				c.dlo_incl = c.dhi_incl = 1;
			}
			if (dbl) {
				double lit = b.kind == 2 ? b.d : (double)b.i;
				if (lit != lit)
					return false;
				auto raise_lo = [&](double v, int incl) {
					if (v > c.dlo || (v == c.dlo && !incl)) { c.dlo = v; c.dlo_incl = incl; }
				};
				auto lower_hi = [&](double v, int incl) {
					if (v < c.dhi || (v == c.dhi && !incl)) { c.dhi = v; c.dhi_incl = incl; }
				};
				switch (cmp) {
				case 1: lower_hi(lit, 0); break;
				case 2: raise_lo(lit, 0); break;
				case 4: raise_lo(lit, 1); lower_hi(lit, 1); break;
				case 5: lower_hi(lit, 1); break;
				case 6: raise_lo(lit, 1); break;
				default: return false;
				}
			} else {
				long long lit = b.i;
				switch (cmp) {
				case 1: if (lit == INT64_MIN) { c.ihi = INT64_MIN; c.ilo = INT64_MAX; } else c.ihi = std::min(c.ihi, lit - 1); break;
				case 2: if (lit == INT64_MAX) { c.ihi = INT64_MIN; c.ilo = INT64_MAX; } else c.ilo = std::max(c.ilo, lit + 1); break;
				case 4: c.ilo = std::max(c.ilo, lit); c.ihi = std::min(c.ihi, lit); break;
				case 5: c.ihi = std::min(c.ihi, lit); break;
				case 6: c.ilo = std::max(c.ilo, lit); break;
				default: return false;
				}
			}
			it.kind = 3;
			st[depth++] = it;
			break;
		}
		case MDBCU_P_AND: {
			if (depth < 2)
				return false;
			Item b = st[--depth], a = st[--depth];
			if (a.kind != 3 || b.kind != 3)
				return false;
			it.kind = 3;
			st[depth++] = it;
			break;
		}
		default:
			return false;
		}
	}
	return plan->n_pred == 0 || (depth == 1 && st[0].kind == 3);
}

int mdb_select_scan_agg(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (plan->n_tables != 1 || plan->n_group != 0 || plan->n_out > SA_MAX_AGGS || (plan->flags & MDBCU_PLAN_DISTRIBUTED))
		return MDBCU_EUNSUPPORTED;
	const mdbcu_table *t = plan->tables[0];
	for (int o = 0; o < plan->n_out; o++)
		if (plan->out[o].kind == MDBCU_OUT_COLUMN)
			return MDBCU_EUNSUPPORTED;
	if (t->n_slots < (1u << 16))
		return MDBCU_EUNSUPPORTED; // tiny tables: the general operators are just as good and keep row order logic in one place

	SASpec sp;
	memset(&sp, 0, sizeof(sp));
	int col_of[MDBCU_MAX_COLUMNS];
	for (int c = 0; c < MDBCU_MAX_COLUMNS; c++)
		col_of[c] = -1;
	if (!pred_to_ranges(plan, &sp, col_of))
		return MDBCU_EUNSUPPORTED;

	bool live_checked = sp.ncols > 0;
	for (int o = 0; o < plan->n_out; o++) {
		const mdbcu_out &po = plan->out[o];
		sp.aggs[o].kind = po.kind;
		sp.aggs[o].col = -1;
		if (po.kind == MDBCU_OUT_COUNT_STAR)
			continue;
		if (po.ref.tbl != 0 || po.ref.col < 0 || po.ref.col >= t->ncols || t->cols[po.ref.col].type == MDBCU_CT_VARCHAR)
			return MDBCU_EUNSUPPORTED;
		int sc = col_of[po.ref.col];
		if (sc < 0) {
			if (sp.ncols == SA_MAX_COLS)
				return MDBCU_EUNSUPPORTED;
			sc = sp.ncols++;
			col_of[po.ref.col] = sc;
			SACol &c = sp.cols[sc];
			memset(&c, 0, sizeof(c));
			c.data = t->cols[po.ref.col].data;
			c.present = col_all_present(t, po.ref.col) ? nullptr : t->cols[po.ref.col].present;
			c.is_dbl = t->cols[po.ref.col].type == MDBCU_CT_DOUBLE;
		}
		sp.aggs[o].col = sc;
	}
	sp.naggs = plan->n_out;
	// COUNT(*) with tombstones but no column constraint would count dead rows: let the general path do it
	if (!t->all_live && !live_checked)
		return MDBCU_EUNSUPPORTED;
	if (sp.ncols == 0) {
		// SELECT COUNT(*) FROM T without WHERE: nothing to scan
		return MDBCU_EUNSUPPORTED;
	}

	ctx->stats.path = MDBCU_PATH_SCAN_AGG;
	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	int grid = ctx->num_sms * 2;
	SAPartial *partials;
	unsigned long long *d_rows;
	MDB_TRY(tmp.alloc(&partials, grid));
	MDB_TRY(tmp.alloc(&d_rows, 1));
	MDB_TRY(mdb_result_alloc(ctx, plan, res, 1, false));
	SAOut out;
	memset(&out, 0, sizeof(out));
	for (int o = 0; o < plan->n_out; o++) {
		out.cells[o] = res->cols[o].cells;
		out.nulls[o] = res->cols[o].nulls;
	}

	clock.begin(0);
	cudaEvent_t k0, k1;
	cudaEventCreate(&k0);
	cudaEventCreate(&k1);
	cudaEventRecord(k0, ctx->stream);
	for (int a2 = sp.naggs; a2 < SA_MAX_AGGS; a2++) {
		sp.aggs[a2].kind = MDBCU_OUT_COUNT_STAR; // padding slots (ignored by the final kernel)
		sp.aggs[a2].col = -1;
	}
	CUDA_TRY(ctx, cudaMemsetAsync(partials, 0, sizeof(SAPartial) * grid, ctx->stream));
	if (sp.ncols == 1)
		launch_scan_agg<1>(ctx, grid, sp, t->n_slots, partials);
	else if (sp.ncols == 2)
		launch_scan_agg<2>(ctx, grid, sp, t->n_slots, partials);
	else
		launch_scan_agg<3>(ctx, grid, sp, t->n_slots, partials);
	cudaEventRecord(k1, ctx->stream);
	clock.begin(4);
	MDB_LAUNCH(ctx, k_scan_aggregate_final, 1, 32, 0, sp, (const SAPartial*)partials, grid, out, d_rows);
	cudaError_t e = cudaGetLastError();
	clock.finish();
	float kms = 0.f;
	cudaEventElapsedTime(&kms, k0, k1);
	cudaEventDestroy(k0);
	cudaEventDestroy(k1);
	if (e != cudaSuccess)
		return mdb_fail(ctx, MDBCU_ECUDA, "scan+aggregate launch failed: %s", cudaGetErrorString(e));

	uint64_t rows = 0;
	MDB_TRY(mdb_read_u64(ctx, (const uint64_t*)d_rows, &rows));
	if (rows == 0)
		res->nrows = 0; // nothing qualified: the reference returns no row (handle_countonly_case keeps zero rows)

	ctx->stats.algorithmic_bytes = 8ull * t->n_slots * sp.ncols + 8ull * plan->n_out;
	ctx->stats.dominant_ms = kms;
	ctx->stats.dominant_bytes = 8ull * t->n_slots * sp.ncols;
	return MDBCU_OK;
}

// ===================================================================================== radix join + count

#define RJ_MAX_PART 4096           // partitions per pass (12 radix bits)
#define RJ_MAX_SHIFT 16            // remainder bits: 2-byte remainders, 64 KiB byte-counter histogram per side
#define RJ_CAP 20                  // staging slots per partition (flush at 16 = one 32-byte sector)
#define RJ_FLUSH 16
#define RJ_CHUNK 256               // remainders per chunk (512 bytes = one warp-wide 128-bit load)
#define RJ_BLOCKS_PER_CHUNK (RJ_CHUNK / RJ_FLUSH)
#define RJ_P1_THREADS 512
#define RJ_P1_LOADS 8              // 128-bit loads (2 keys each) per thread per round
#define RJ_P1_TILE (RJ_P1_THREADS * RJ_P1_LOADS * 2)
#define RJ_P2_THREADS 1024
#define RJ_NONE 0xffffffffu

struct RJSide {
	const int64_t *keys;
	const uint32_t *present;
	uint64_t n;
	// chunk pool
	uint16_t *pool;            // pool_chunks * RJ_CHUNK remainders
	uint32_t pool_chunks;
	uint32_t *pool_next;       // allocation cursor
	uint16_t *chunk_part;      // partition of each chunk
	uint16_t *chunk_entries;   // valid remainders in each chunk
	uint32_t *dir_cnt;         // chunks per partition (RJ_MAX_PART + 1)
	uint32_t *dir;             // chunk ids grouped by partition
	uint64_t *dir_off;         // exclusive offsets into dir (RJ_MAX_PART + 1)
	uint32_t *dir_fill;
};

struct RJParams {
	long long kmin;
	unsigned long long range;  // keys in [kmin, kmin + range) can match
	int shift;                 // remainder bits
	int nparts;
	uint32_t *error_flag;      // bit 0: chunk pool exhausted, bit 1: byte counter overflow
};

// shared memory layout of pass 1 (dynamic)
struct RJP1Smem {
	uint16_t stage[RJ_MAX_PART * RJ_CAP]; // 160 KiB
	uint32_t fill[RJ_MAX_PART];
	uint32_t chunk_id[RJ_MAX_PART];
	uint16_t worklist[RJ_MAX_PART];
	uint8_t chunk_blocks[RJ_MAX_PART];
	uint32_t wl_count;
};

__device__ static inline uint32_t rj_new_chunk(const RJSide &s, const RJParams &pr, RJP1Smem *sm, uint32_t p)
{
	uint32_t old = sm->chunk_id[p];
	uint32_t cid = atomicAdd(s.pool_next, 1u);
	if (cid >= s.pool_chunks) {
		atomicOr(pr.error_flag, 1u);
		return RJ_NONE;
	}
	if (old != RJ_NONE)
		s.chunk_entries[old] = RJ_CHUNK; // retired chunks are always full
	s.chunk_part[cid] = (uint16_t)p;
	atomicAdd(&s.dir_cnt[p], 1u);
	sm->chunk_id[p] = cid;
	sm->chunk_blocks[p] = 0;
	return cid;
}

// Pass 1.  One persistent CTA per SM.  Keys are streamed with 128-bit loads (double-buffered in
// registers); each key is reduced to (partition, 16-bit remainder) and appended to the partition's
// staging slots in shared memory with one shared-memory atomic.  A partition whose 16th slot fills
// is queued and flushed as ONE aligned 32-byte sector into the CTA's current 512-byte chunk of that
// partition, so DRAM only ever sees full-sector writes.
__global__ void __launch_bounds__(RJ_P1_THREADS, 1) k_radix_partition(RJSide s, RJParams pr)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RJP1Smem *sm = reinterpret_cast<RJP1Smem*>(smem_raw);
	const int tid = threadIdx.x;

	for (int p = tid; p < RJ_MAX_PART; p += RJ_P1_THREADS) {
		sm->fill[p] = 0;
		sm->chunk_id[p] = RJ_NONE;
		sm->chunk_blocks[p] = 0;
	}
	if (tid == 0)
		sm->wl_count = 0;
	__syncthreads();

	const uint64_t ntiles = (s.n + RJ_P1_TILE - 1) / RJ_P1_TILE;
	const int4 *src = reinterpret_cast<const int4*>(s.keys);
	const uint64_t npairs = s.n / 2;

	int4 cur[RJ_P1_LOADS], nxt[RJ_P1_LOADS];
	auto load_tile = [&](uint64_t tile, int4 *dst) {
		uint64_t base_pair = tile * (RJ_P1_TILE / 2);
#pragma unroll
		for (int j = 0; j < RJ_P1_LOADS; j++) {
			uint64_t pi = base_pair + (uint64_t)j * RJ_P1_THREADS + tid;
			if (pi < npairs) {
				dst[j] = mdb_ldg_stream(src + pi);
			} else if (pi == npairs && (s.n & 1)) {
				long long last = s.keys[s.n - 1];
				dst[j] = make_int4((int)(unsigned)(unsigned long long)last, (int)(unsigned)((unsigned long long)last >> 32), 0, 0);
			} else {
				dst[j] = make_int4(0, 0, 0, 0);
			}
		}
	};

	uint64_t tile = blockIdx.x;
	if (tile < ntiles)
		load_tile(tile, cur);

	for (; tile < ntiles; tile += gridDim.x) {
		uint64_t next_tile = tile + gridDim.x;
		if (next_tile < ntiles)
			load_tile(next_tile, nxt);

		// decode this round's keys into (partition << 16 | remainder), RJ_NONE = nothing to insert
		uint32_t item[RJ_P1_LOADS * 2];
		uint64_t base_pair = tile * (RJ_P1_TILE / 2);
#pragma unroll
		for (int j = 0; j < RJ_P1_LOADS; j++) {
			uint64_t pi = base_pair + (uint64_t)j * RJ_P1_THREADS + tid;
			uint32_t pw = 0xffffffffu;
			if (s.present && pi * 2 < s.n)
				pw = s.present[pi >> 4];
			int bit = (int)((pi & 15) * 2);
#pragma unroll
			for (int h = 0; h < 2; h++) {
				uint64_t row = pi * 2 + h;
				long long key = h == 0 ? (long long)(((unsigned long long)(unsigned)cur[j].y << 32) | (unsigned)cur[j].x)
						       : (long long)(((unsigned long long)(unsigned)cur[j].w << 32) | (unsigned)cur[j].z);
				unsigned long long d = (unsigned long long)key - (unsigned long long)pr.kmin;
				bool ok = row < s.n && ((pw >> (bit + h)) & 1) && d < pr.range;
				item[j * 2 + h] = ok ? (uint32_t)(((d >> pr.shift) << 16) | (d & ((1ull << pr.shift) - 1))) : RJ_NONE;
			}
		}

		bool pending;
		do {
			pending = false;
#pragma unroll
			for (int k = 0; k < RJ_P1_LOADS * 2; k++) {
				uint32_t it = item[k];
				if (it == RJ_NONE)
					continue;
				uint32_t p = it >> 16;
				uint32_t pos = atomicAdd(&sm->fill[p], 1u);
				if (pos < RJ_CAP) {
					sm->stage[p * RJ_CAP + pos] = (uint16_t)it;
					item[k] = RJ_NONE;
					if (pos == RJ_FLUSH - 1)
						sm->worklist[atomicAdd(&sm->wl_count, 1u)] = (uint16_t)p;
				} else {
					pending = true; // staging full: retry after this round's flush
				}
			}
			pending = __syncthreads_or(pending);
			uint32_t nwl = sm->wl_count;
			__syncthreads();
			if (tid == 0)
				sm->wl_count = 0;
			// flush: one lane per queued partition moves its first 16 remainders (one 32-byte sector)
			for (uint32_t w = tid; w < nwl; w += RJ_P1_THREADS) {
				uint32_t p = sm->worklist[w];
				uint32_t f = min(sm->fill[p], (uint32_t)RJ_CAP);
				uint32_t cid = sm->chunk_id[p];
				if (cid == RJ_NONE || sm->chunk_blocks[p] == RJ_BLOCKS_PER_CHUNK)
					cid = rj_new_chunk(s, pr, sm, p);
				const uint2 *st = reinterpret_cast<const uint2*>(&sm->stage[p * RJ_CAP]); // 40-byte rows: 8-byte aligned
				uint2 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4];
				if (cid != RJ_NONE) {
					uint32_t blk = sm->chunk_blocks[p];
					int4 *dst = reinterpret_cast<int4*>(s.pool + (size_t)cid * RJ_CHUNK + blk * RJ_FLUSH);
					dst[0] = make_int4((int)a.x, (int)a.y, (int)b.x, (int)b.y);
					dst[1] = make_int4((int)c.x, (int)c.y, (int)d.x, (int)d.y);
					sm->chunk_blocks[p] = (uint8_t)(blk + 1);
				}
				// keep the (at most 4) remainders behind the flushed sector
				reinterpret_cast<uint2*>(&sm->stage[p * RJ_CAP])[0] = e;
				sm->fill[p] = f - RJ_FLUSH;
			}
			__syncthreads();
		} while (pending);

#pragma unroll
		for (int j = 0; j < RJ_P1_LOADS; j++)
			cur[j] = nxt[j];
	}

	// drain: every partition's partial sector goes out, chunk entry counts are finalised
	__syncthreads();
	for (int p = tid; p < pr.nparts; p += RJ_P1_THREADS) {
		uint32_t f = sm->fill[p];
		uint32_t cid = sm->chunk_id[p];
		if (f > 0) {
			if (cid == RJ_NONE || sm->chunk_blocks[p] == RJ_BLOCKS_PER_CHUNK)
				cid = rj_new_chunk(s, pr, sm, p);
			if (cid != RJ_NONE) {
				uint32_t blk = sm->chunk_blocks[p];
				uint16_t *dst = s.pool + (size_t)cid * RJ_CHUNK + blk * RJ_FLUSH;
				for (uint32_t i = 0; i < f; i++)
					dst[i] = sm->stage[p * RJ_CAP + i];
				s.chunk_entries[cid] = (uint16_t)(blk * RJ_FLUSH + f);
			}
		} else if (cid != RJ_NONE) {
			s.chunk_entries[cid] = (uint16_t)(sm->chunk_blocks[p] * RJ_FLUSH);
		}
	}
}

// group chunk ids by partition (counting sort; counts were accumulated in pass 1)
__global__ void k_radix_dir_scan(RJSide s, int nparts)
{
	// single block, nparts <= 4096
	__shared__ uint64_t warp_tot[33];
	uint64_t carry = 0;
	for (int base = 0; base < nparts + 1; base += blockDim.x) {
		int i = base + threadIdx.x;
		uint64_t v = i < nparts ? s.dir_cnt[i] : 0;
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

__global__ void k_radix_dir_fill(RJSide s)
{
	uint32_t nchunks = min(*s.pool_next, s.pool_chunks);
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += gridDim.x * blockDim.x) {
		uint32_t p = s.chunk_part[c];
		uint32_t pos = atomicAdd(&s.dir_fill[p], 1u);
		s.dir[s.dir_off[p] + pos] = c;
	}
}

struct RJOut {
	int nout;
	int is_count[4];
	int64_t *cells[4];
	unsigned long long *cursor;
	uint64_t cap;
};

__device__ static inline void rj_histogram(const RJSide &s, uint32_t p, uint32_t *cnt, uint32_t *total_smem)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = RJ_P2_THREADS / 32;
	const uint64_t c0 = s.dir_off[p], c1 = s.dir_off[p + 1];
	uint32_t seen = 0;
	constexpr int MLP = 4; // chunks in flight per warp
	for (uint64_t c = c0 + (uint64_t)warp * MLP; c < c1; c += (uint64_t)nwarps * MLP) {
		int4 v[MLP];
		uint32_t ne[MLP];
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			ne[u] = 0;
			if (c + u < c1) {
				uint32_t cid = s.dir[c + u];
				ne[u] = s.chunk_entries[cid];
				v[u] = mdb_ldg_stream(reinterpret_cast<const int4*>(s.pool + (size_t)cid * RJ_CHUNK) + lane);
			}
		}
#pragma unroll
		for (int u = 0; u < MLP; u++) {
			if (!ne[u])
				continue;
			uint32_t w[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z, (uint32_t)v[u].w};
#pragma unroll
			for (int j = 0; j < 8; j++) {
				uint32_t idx = lane * 8 + j;
				if (idx < ne[u]) {
					uint32_t rem = (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu;
					atomicAdd(&cnt[rem >> 2], 1u << ((rem & 3) * 8));
				}
			}
			if (lane == 0)
				seen += ne[u];
		}
	}
	if (lane == 0 && seen)
		atomicAdd(total_smem, seen);
}

// Pass 2.  Persistent CTAs take partitions from an atomic counter.  Both sides of a partition are
// histogrammed into byte counters in shared memory (4 keys per 32-bit word, shared-memory atomics),
// the byte sums are checked against the number of remainders (a wrapped byte counter changes the sum),
// and every key present on both sides is emitted with count cntA * cntB.
__global__ void __launch_bounds__(RJ_P2_THREADS, 1)
k_radix_joincount(RJSide a, RJSide b, RJParams pr, RJOut out, uint32_t *__restrict__ part_counter)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int D = 1 << pr.shift;
	const int words = D >= 4 ? D / 4 : 1;
	uint32_t *cntA = reinterpret_cast<uint32_t*>(smem_raw);
	uint32_t *cntB = cntA + words;
	__shared__ uint32_t s_part, s_totA, s_totB, s_sumA, s_sumB;
	__shared__ uint32_t s_scan[33];
	__shared__ unsigned long long s_base;
	const int tid = threadIdx.x;

	while (true) {
		if (tid == 0) {
			s_part = atomicAdd(part_counter, 1u);
			s_totA = s_totB = s_sumA = s_sumB = 0;
		}
		__syncthreads();
		const uint32_t p = s_part;
		if (p >= (uint32_t)pr.nparts)
			break;

		for (int w = tid; w < words * 2; w += RJ_P2_THREADS)
			cntA[w] = 0; // cntB follows cntA
		__syncthreads();
		rj_histogram(a, p, cntA, &s_totA);
		rj_histogram(b, p, cntB, &s_totB);
		__syncthreads();

		// byte-sum check + count matches
		uint32_t sumA = 0, sumB = 0, matches = 0;
		for (int w = tid; w < words; w += RJ_P2_THREADS) {
			uint32_t x = cntA[w], y = cntB[w];
			sumA = __dp4a(x, 0x01010101u, sumA);
			sumB = __dp4a(y, 0x01010101u, sumB);
			uint32_t m = __vcmpne4(x, 0) & __vcmpne4(y, 0);
			matches += __popc(m) >> 3;
		}
		for (int o = 16; o > 0; o >>= 1) {
			sumA += __shfl_xor_sync(0xffffffffu, sumA, o);
			sumB += __shfl_xor_sync(0xffffffffu, sumB, o);
		}
		if ((tid & 31) == 0) {
			atomicAdd(&s_sumA, sumA);
			atomicAdd(&s_sumB, sumB);
		}
		// block exclusive scan of match counts
		uint32_t incl = matches;
		int lane = tid & 31, warp = tid >> 5;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o)
				incl += n;
		}
		if (lane == 31)
			s_scan[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = s_scan[lane], wi = w;
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
				if (lane >= o)
					wi += n;
			}
			s_scan[lane] = wi - w;
			if (lane == 31) {
				s_scan[32] = wi;
				s_base = wi ? atomicAdd(out.cursor, (unsigned long long)wi) : 0ull;
				if (s_sumA != s_totA || s_sumB != s_totB)
					atomicOr(pr.error_flag, 2u); // a byte counter wrapped (>255 equal keys): host falls back
			}
		}
		__syncthreads();
		unsigned long long pos = s_base + s_scan[warp] + incl - matches;
		if (matches && pos + matches <= out.cap) {
			long long key_base = pr.kmin + (long long)((unsigned long long)p << pr.shift);
			for (int w = tid; w < words; w += RJ_P2_THREADS) {
				uint32_t x = cntA[w], y = cntB[w];
				uint32_t m = __vcmpne4(x, 0) & __vcmpne4(y, 0);
				while (m) {
					int byte = (__ffs(m) - 1) >> 3;
					m &= ~(0xffu << (byte * 8));
					long long key = key_base + (long long)w * 4 + byte;
					long long cnt = (long long)((x >> (byte * 8)) & 0xff) * (long long)((y >> (byte * 8)) & 0xff);
#pragma unroll
					for (int o = 0; o < 4; o++)
						if (o < out.nout)
							out.cells[o][pos] = out.is_count[o] ? cnt : key;
					pos++;
				}
			}
		}
		__syncthreads();
	}
}

static int rj_side_setup(mdbcu_ctx *ctx, DevTemp &tmp, RJSide *s, const mdbcu_table *t, int col, int grid)
{
	memset(s, 0, sizeof(*s));
	s->keys = t->cols[col].data;
	s->present = col_all_present(t, col) ? nullptr : t->cols[col].present;
	s->n = t->n_slots;
	uint64_t chunks = t->n_slots / RJ_CHUNK + (uint64_t)grid * RJ_MAX_PART + 1024;
	if (chunks >= 0xfffffff0ull)
		return MDBCU_EUNSUPPORTED;
	s->pool_chunks = (uint32_t)chunks;
	MDB_TRY(tmp.alloc(&s->pool, chunks * RJ_CHUNK));
	MDB_TRY(tmp.alloc(&s->pool_next, 1));
	MDB_TRY(tmp.alloc(&s->chunk_part, chunks));
	MDB_TRY(tmp.alloc(&s->chunk_entries, chunks));
	MDB_TRY(tmp.alloc(&s->dir_cnt, RJ_MAX_PART + 1));
	MDB_TRY(tmp.alloc(&s->dir_fill, RJ_MAX_PART + 1));
	MDB_TRY(tmp.alloc(&s->dir_off, RJ_MAX_PART + 2));
	MDB_TRY(tmp.alloc(&s->dir, chunks));
	CUDA_TRY(ctx, cudaMemsetAsync(s->pool_next, 0, sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(s->dir_cnt, 0, (RJ_MAX_PART + 1) * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(s->dir_fill, 0, (RJ_MAX_PART + 1) * sizeof(uint32_t), ctx->stream));
	return MDBCU_OK;
}

int mdb_select_radix_joincount(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (plan->n_tables != 2 || plan->n_joins != 1 || plan->joins[0].cross || plan->n_pred != 0 || plan->n_group != 1 ||
			plan->n_out < 1 || plan->n_out > 4)
		return MDBCU_EUNSUPPORTED;
	if (plan->flags & MDBCU_PLAN_DISTRIBUTED)
		return MDBCU_EUNSUPPORTED; // the exchange variant lives in mdb_comm.cu (round 2)
	const mdbcu_join &jn = plan->joins[0];
	if (jn.left.tbl != 0 || jn.right.tbl != 1)
		return MDBCU_EUNSUPPORTED;
	const mdbcu_table *ta = plan->tables[0], *tb = plan->tables[1];
	if (jn.left.col < 0 || jn.left.col >= ta->ncols || jn.right.col < 0 || jn.right.col >= tb->ncols)
		return MDBCU_EUNSUPPORTED;
	const DevColumn &ca = ta->cols[jn.left.col], &cb = tb->cols[jn.right.col];
	auto intlike = [](int type) { return type == MDBCU_CT_INTEGER || type == MDBCU_CT_DATE || type == MDBCU_CT_DATETIME; };
	if (!intlike(ca.type) || !intlike(cb.type) || !ca.stats_ok || !cb.stats_ok)
		return MDBCU_EUNSUPPORTED;
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
	if (ta->n_slots + tb->n_slots < (1ull << 20))
		return MDBCU_EUNSUPPORTED; // small inputs: general operators (they also return the reference's row order)

	// only keys inside both columns' [min, max] can ever match
	long long kmin = std::max(ca.imin, cb.imin), kmax = std::min(ca.imax, cb.imax);
	if (ca.imin > ca.imax || cb.imin > cb.imax || kmin > kmax) {
		ctx->stats.path = MDBCU_PATH_RADIX_JOINCOUNT;
		return mdb_result_alloc(ctx, plan, res, 0, false);
	}
	unsigned long long range = (unsigned long long)kmax - (unsigned long long)kmin + 1ull;
	if (range == 0 || range > ((unsigned long long)RJ_MAX_PART << RJ_MAX_SHIFT) || range < 4096)
		return MDBCU_EUNSUPPORTED;
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
	MDB_TRY(rj_side_setup(ctx, tmp, &sa, ta, jn.left.col, grid1));
	MDB_TRY(rj_side_setup(ctx, tmp, &sb, tb, jn.right.col, grid1));
	pr.kmin = kmin;
	pr.range = range;
	pr.shift = shift;
	pr.nparts = nparts;
	uint32_t *d_flags; // [0] error flag, [1] partition counter
	unsigned long long *d_cursor;
	MDB_TRY(tmp.alloc(&d_flags, 2));
	MDB_TRY(tmp.alloc(&d_cursor, 1));
	CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
	pr.error_flag = d_flags;

	uint64_t cap_groups = std::min<uint64_t>(std::min<uint64_t>(ta->n_slots, tb->n_slots), range);
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
	size_t smem1 = sizeof(RJP1Smem);
	size_t smem2 = 2 * (size_t)std::max(1, (1 << shift) / 4) * sizeof(uint32_t);
	if (!attr_done) {
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_joincount, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
		attr_done = true;
	}

	cudaEvent_t k0, k1;
	cudaEventCreate(&k0);
	cudaEventCreate(&k1);
	clock.begin(1);
	cudaEventRecord(k0, ctx->stream);
	MDB_LAUNCH(ctx, k_radix_partition, grid1, RJ_P1_THREADS, smem1, sa, pr);
	MDB_LAUNCH(ctx, k_radix_partition, grid1, RJ_P1_THREADS, smem1, sb, pr);
	cudaEventRecord(k1, ctx->stream);
	clock.begin(7);
	MDB_LAUNCH(ctx, k_radix_dir_scan, 1, 1024, 0, sa, nparts);
	MDB_LAUNCH(ctx, k_radix_dir_scan, 1, 1024, 0, sb, nparts);
	MDB_LAUNCH(ctx, k_radix_dir_fill, ctx->num_sms * 4, 256, 0, sa);
	MDB_LAUNCH(ctx, k_radix_dir_fill, ctx->num_sms * 4, 256, 0, sb);
	clock.begin(2);
	cudaEvent_t k2, k3;
	cudaEventCreate(&k2);
	cudaEventCreate(&k3);
	cudaEventRecord(k2, ctx->stream);
	MDB_LAUNCH(ctx, k_radix_joincount, std::min(nparts, ctx->num_sms), RJ_P2_THREADS, smem2, sa, sb, pr, out, d_flags + 1);
	cudaEventRecord(k3, ctx->stream);
	cudaError_t e = cudaGetLastError();
	clock.finish();
	float ms1 = 0.f, ms2 = 0.f;
	cudaEventElapsedTime(&ms1, k0, k1);
	cudaEventElapsedTime(&ms2, k2, k3);
	cudaEventDestroy(k0);
	cudaEventDestroy(k1);
	cudaEventDestroy(k2);
	cudaEventDestroy(k3);
	if (e != cudaSuccess)
		return mdb_fail(ctx, MDBCU_ECUDA, "radix join launch failed: %s", cudaGetErrorString(e));

	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, d_cursor, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar + 1, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	uint64_t ngroups = ctx->h_scalar[0];
	uint32_t flags = (uint32_t)(ctx->h_scalar[1] & 0xffffffffu);
	if (flags || ngroups > cap_groups) {
		// more than 255 equal keys in one partition (or pool exhaustion): redo with the general operators
		return MDBCU_EUNSUPPORTED;
	}
	res->nrows = ngroups;

	// compulsory traffic: every key once in, every (key, count) group once out
	ctx->stats.algorithmic_bytes = 8ull * (ta->n_slots + tb->n_slots) + 8ull * plan->n_out * ngroups;
	ctx->stats.dominant_ms = ms1 + ms2;
	ctx->stats.dominant_bytes = ctx->stats.algorithmic_bytes;
	return MDBCU_OK;
}

// ===================================================================================== small-build star join

int mdb_select_direct_star(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	(void)ctx;
	(void)plan;
	(void)res;
	return MDBCU_EUNSUPPORTED; // config 5 currently runs on the general operators
}
