// mdb_fast.cu - the fused filter + aggregate scan (BASELINE.json config 2), single-GPU and distributed.
//
//  K2 scan_filter_aggregate  : SELECT <aggregates> FROM T WHERE <range conjunction>   (config 2)
//       replaces proc_from_clause_table :1282 + proc_where_clause :1435 + handle_countonly_case :1590
//       of src/engine/executor_select.c with ONE pass over the referenced columns.
//  (the radix join+count of the README query lives in mdb_radix.cu, the small-build star join in mdb_star.cu)
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

// One row.  Branch-free in everything that depends on the data: the row's verdict is a predicate that is folded into
// every accumulator update (the only branches left test the plan and are uniform across the grid), so a warp never
// diverges on selectivity.
template <int NC, int NA>
__device__ __forceinline__ void sa_row(const SASpec &sp, SAState<NA> &st, const long long *v, const bool *pres)
{
	bool ok = true;
#pragma unroll
	for (int c = 0; c < NC; c++) {
		const SACol &col = sp.cols[c];
		if (!col.has_range)
			continue;
		bool in;
		if (col.is_dbl) {
			const double d = __longlong_as_double(v[c]);
			const bool lo = col.dlo_incl ? d >= col.dlo : d > col.dlo;
			const bool hi = col.dhi_incl ? d <= col.dhi : d < col.dhi;
			in = lo & hi;
		} else {
			in = (v[c] >= col.ilo) & (v[c] <= col.ihi);
		}
		ok = ok & pres[c] & in;
	}
	st.rows += ok ? 1u : 0u;
#pragma unroll
	for (int a = 0; a < NA; a++) {
		const int kind = sp.aggs[a].kind, c = sp.aggs[a].col;
		if (kind == MDBCU_OUT_COUNT_STAR) {
			st.nn[a] += ok ? 1u : 0u;
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
		const bool take = ok & p;
		st.nn[a] += take ? 1u : 0u;
		switch (kind) {
		case MDBCU_OUT_SUM: case MDBCU_OUT_AVG:
			if (dbl)
				st.acc[a] = __double_as_longlong(__longlong_as_double(st.acc[a]) + (take ? __longlong_as_double(x) : 0.0));
			else
				st.acc[a] = (long long)((unsigned long long)st.acc[a] + (take ? (unsigned long long)x : 0ull));
			break;
		case MDBCU_OUT_MIN:
			st.acc[a] = min(st.acc[a], take ? (dbl ? mdb_dbl_to_ordered(x) : x) : INT64_MAX);
			break;
		case MDBCU_OUT_MAX:
			st.acc[a] = max(st.acc[a], take ? (dbl ? mdb_dbl_to_ordered(x) : x) : INT64_MIN);
			break;
		}
	}
}

// Each thread streams quads of rows with one 256-bit load per referenced column (4 x 8-byte cells),
// two quads in flight; per-thread partials -> warp shuffles -> block -> one partial per block.
// A second, single-block kernel folds the block partials in a fixed order (deterministic doubles).
template <int NC, int NA>
__global__ void __launch_bounds__(SA_THREADS, (NA >= 4 ? 2 : 3)) k_scan_filter_aggregate(SASpec sp, uint64_t n, SAPartial *__restrict__ partials)
{
	SAState<NA> st;
	sa_init<NA>(sp, st);

	// 256-bit loads: four rows of one column per load, two loads per column in flight per thread (a warp covers
	// 1 KiB of a column per instruction; measured on B200: 256-bit loads stream at 6.4 TB/s, 128-bit loads at 4.7)
	const uint64_t nquads = n / 4;
	const uint64_t stride = (uint64_t)gridDim.x * SA_THREADS;
	uint64_t quad = (uint64_t)blockIdx.x * SA_THREADS + threadIdx.x;

	// Eight rows per thread and iteration (two 256-bit loads per referenced column).  The plan is decoded once per
	// batch, not once per row: first the rows' verdict as an 8-bit mask (column by column), then aggregate by aggregate.
	constexpr int UNROLL = 2, ROWS = 4 * UNROLL;
	for (; quad + (UNROLL - 1) * stride < nquads; quad += UNROLL * stride) {
		uint32_t raw[NC][UNROLL][8], pbits[NC];
#pragma unroll
		for (int c = 0; c < NC; c++) {
			pbits[c] = 0;
#pragma unroll
			for (int u = 0; u < UNROLL; u++) {
				const uint64_t qi = quad + u * stride;
				const char *src = reinterpret_cast<const char*>(sp.cols[c].data) + qi * 32;
				asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
						: "=r"(raw[c][u][0]), "=r"(raw[c][u][1]), "=r"(raw[c][u][2]), "=r"(raw[c][u][3]), "=r"(raw[c][u][4]),
						  "=r"(raw[c][u][5]), "=r"(raw[c][u][6]), "=r"(raw[c][u][7]) : "l"(src));
				const uint32_t pw = sp.cols[c].present ? sp.cols[c].present[qi >> 3] : 0xffffffffu;
				pbits[c] |= ((pw >> ((qi & 7) * 4)) & 0xfu) << (4 * u);
			}
		}
		uint32_t ok = (1u << ROWS) - 1u;
#pragma unroll
		for (int c = 0; c < NC; c++) {
			const SACol &col = sp.cols[c];
			if (!col.has_range) // uniform
				continue;
			uint32_t in = 0;
			if (col.is_dbl) {
#pragma unroll
				for (int r = 0; r < ROWS; r++) {
					const double d = __hiloint2double((int)raw[c][r / 4][2 * (r % 4) + 1], (int)raw[c][r / 4][2 * (r % 4)]);
					const bool lo = col.dlo_incl ? d >= col.dlo : d > col.dlo;
					const bool hi = col.dhi_incl ? d <= col.dhi : d < col.dhi;
					in |= (lo && hi) ? (1u << r) : 0u;
				}
			} else {
#pragma unroll
				for (int r = 0; r < ROWS; r++) {
					const long long v = (long long)(((unsigned long long)raw[c][r / 4][2 * (r % 4) + 1] << 32) | raw[c][r / 4][2 * (r % 4)]);
					in |= (v >= col.ilo && v <= col.ihi) ? (1u << r) : 0u;
				}
			}
			ok &= in & pbits[c];
		}
		st.rows += __popc(ok);
#pragma unroll
		for (int a = 0; a < NA; a++) {
			const int kind = sp.aggs[a].kind, ac = sp.aggs[a].col; // uniform
			if (kind == MDBCU_OUT_COUNT_STAR) {
				st.nn[a] += __popc(ok);
				continue;
			}
#pragma unroll
			for (int c = 0; c < NC; c++) {
				if (ac != c)
					continue;
				const uint32_t take = ok & pbits[c];
				st.nn[a] += __popc(take);
				const bool dbl = sp.cols[c].is_dbl;
				if (kind == MDBCU_OUT_SUM || kind == MDBCU_OUT_AVG) {
					if (dbl) {
						double sum = __longlong_as_double(st.acc[a]);
#pragma unroll
						for (int r = 0; r < ROWS; r++)
							sum += ((take >> r) & 1u) ? __hiloint2double((int)raw[c][r / 4][2 * (r % 4) + 1], (int)raw[c][r / 4][2 * (r % 4)]) : 0.0;
						st.acc[a] = __double_as_longlong(sum);
					} else {
						unsigned long long sum = (unsigned long long)st.acc[a];
#pragma unroll
						for (int r = 0; r < ROWS; r++)
							sum += ((take >> r) & 1u) ? (((unsigned long long)raw[c][r / 4][2 * (r % 4) + 1] << 32) | raw[c][r / 4][2 * (r % 4)]) : 0ull;
						st.acc[a] = (long long)sum;
					}
				} else if (kind == MDBCU_OUT_MIN || kind == MDBCU_OUT_MAX) {
					long long best = st.acc[a];
#pragma unroll
					for (int r = 0; r < ROWS; r++) {
						const long long x = (long long)(((unsigned long long)raw[c][r / 4][2 * (r % 4) + 1] << 32) | raw[c][r / 4][2 * (r % 4)]);
						const long long y = dbl ? mdb_dbl_to_ordered(x) : x;
						if ((take >> r) & 1u)
							best = kind == MDBCU_OUT_MIN ? min(best, y) : max(best, y);
					}
					st.acc[a] = best;
				}
			}
		}
	}
	// remainder: single rows
	for (uint64_t r = quad * 4; r < n; r += 4 * stride) {
		for (int k = 0; k < 4 && r + k < n; k++) {
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

// one rank's block partials folded (in block order) into a single partial: what a distributed plan exchanges
__global__ void k_scan_aggregate_fold(SASpec sp, const SAPartial *__restrict__ partials, int nparts, SAPartial *__restrict__ dst)
{
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	SAState<SA_MAX_AGGS> tot;
	SASpec full = sp;
	for (int a = sp.naggs; a < SA_MAX_AGGS; a++) {
		full.aggs[a].kind = MDBCU_OUT_COUNT_STAR;
		full.aggs[a].col = -1;
	}
	sa_init<SA_MAX_AGGS>(full, tot);
	for (int b = 0; b < nparts; b++)
		sa_merge<SA_MAX_AGGS>(full, tot, partials[b].rows, partials[b].nn, partials[b].acc);
	dst->rows = tot.rows;
	for (int a = 0; a < SA_MAX_AGGS; a++) {
		dst->nn[a] = tot.nn[a];
		dst->acc[a] = tot.acc[a];
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
	if (plan->n_tables != 1 || plan->n_group != 0 || plan->n_out > SA_MAX_AGGS)
		return MDBCU_EUNSUPPORTED;
	// Distributed plan ("C2: embarrassingly parallel scan + a small reduce"): every rank scans ITS shard, the
	// ranks' partials are all-gathered and folded in rank order (deterministic DOUBLE sums for a given number of ranks),
	// rank 0 returns the row and the other ranks return no row, so that the ranks' results concatenate to the answer.
	// Every decision below depends on the plan and the schema only: all ranks take this path or none does.
	const bool dist = (plan->flags & MDBCU_PLAN_DISTRIBUTED) != 0;
	if (dist && !mdb_comm_ready(ctx))
		return mdb_fail(ctx, MDBCU_EERROR, "MDBCU_PLAN_DISTRIBUTED needs mdbcu_comm_init first");
	const mdbcu_table *t = plan->tables[0];
	for (int o = 0; o < plan->n_out; o++)
		if (plan->out[o].kind == MDBCU_OUT_COLUMN)
			return MDBCU_EUNSUPPORTED;
	if (!dist && t->n_slots < (1u << 16))
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
	// (live_checked is false only when no column is referenced at all, which the next test rejects on every rank alike)
	if (!t->all_live && !live_checked)
		return MDBCU_EUNSUPPORTED;
	if (sp.ncols == 0) {
		// SELECT COUNT(*) FROM T without WHERE: nothing to scan
		return MDBCU_EUNSUPPORTED;
	}

	ctx->stats.path = MDBCU_PATH_SCAN_AGG;
	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	int grid = ctx->num_sms * 3;
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
	KernelTimer ktimer;
	ktimer.start(ctx->stream);
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
	ktimer.stop(ctx->stream);
	clock.begin(4);
	const int W = dist ? ctx->world : 1;
	if (W > 1) {
		// this rank's block partials -> one partial; all ranks' partials -> every rank; rank 0 folds them in rank order
		SAPartial *mine, *all;
		MDB_TRY(tmp.alloc(&mine, 1));
		MDB_TRY(tmp.alloc(&all, W));
		MDB_LAUNCH(ctx, k_scan_aggregate_fold, 1, 32, 0, sp, (const SAPartial*)partials, grid, mine);
		clock.begin(6);
		MDB_TRY(mdb_comm_allgather_bytes(ctx, mine, all, sizeof(SAPartial)));
		ctx->stats.exchange_bytes = (uint64_t)(W - 1) * sizeof(SAPartial);
		clock.begin(4);
		MDB_LAUNCH(ctx, k_scan_aggregate_final, 1, 32, 0, sp, (const SAPartial*)all, W, out, d_rows);
	} else {
		MDB_LAUNCH(ctx, k_scan_aggregate_final, 1, 32, 0, sp, (const SAPartial*)partials, grid, out, d_rows);
	}
	cudaError_t e = cudaGetLastError();
	clock.finish();
	const float kms = ktimer.ms();
	if (e != cudaSuccess)
		return mdb_fail(ctx, MDBCU_ECUDA, "scan+aggregate launch failed: %s", cudaGetErrorString(e));

	uint64_t rows = 0;
	MDB_TRY(mdb_read_u64(ctx, (const uint64_t*)d_rows, &rows));
	if (rows == 0 || (W > 1 && ctx->rank != 0))
		res->nrows = 0; // nothing qualified: the reference returns no row (handle_countonly_case keeps zero rows); rank > 0: rank 0 has it

	ctx->stats.algorithmic_bytes = 8ull * t->n_slots * sp.ncols + 8ull * plan->n_out;
	ctx->stats.dominant_ms = kms;
	ctx->stats.dominant_bytes = 8ull * t->n_slots * sp.ncols;
	return MDBCU_OK;
}
