/*
 * mdb_oracle.c - TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's SELECT hot path, used only as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under midoridb_b200/ links,
 * loads or calls it.
 *
 * Parity pinning: this oracle is checked (tests/test_oracle.py) against
 *   (1) the reference's own known-answer vectors, tests/engine/executor_select.c:47-401
 *       (test_select_1..12), committed as tests/golden/reference_select.json, and
 *   (2) outputs of the UNMODIFIED reference executor compiled in place (oracle/_ref, oracle/Makefile)
 *       on randomised inputs inside the reference's correct domain.
 * SUM / MIN / MAX / AVG, BETWEEN, composite GROUP BY and >1 JOIN cannot be run on the reference at all
 * (midorisql.y:286-287 has COUNT only; defect D3): for those parity is UNPINNED by the reference and
 * the oracle follows ISO SQL with the reference's NULL rules, cross-checked against sqlite3.
 *
 * It follows the reference's pipeline stage by stage (src/engine/executor_select.c):
 *   proc_from_clause_table :1282   live rows in storage order (stop at flags.empty, skip flags.deleted)
 *   _join_nested_loop_tbl2tbl :1076 left-major pair order; ON equality; NULL never matches (:557-579)
 *   proc_where_clause :1435        2-valued predicate, comparison with NULL is false (:629-631)
 *   proc_groupby_clause :1526      survivor = first row of a group, NULL keys collate equal (:1476-1482)
 *   inc_count_cols :1501           COUNT(*) of the survivor
 *   proc_select_clause :1369       projection
 *   handle_countonly_case :1590    COUNT(*)-only collapses to one row (no row when nothing qualifies)
 * but replaces the O(|A||B|) pair loop and the O(n^2) dedupe by hashing, and fixes the defects the
 * reference has outside its correct domain (D2-D5): full 64-bit compares, IN = any-of, SQL joins.
 *
 * Plans are the same `struct mdbcu_plan` the CUDA library takes (include/midoridb_cuda.h) with
 * plan->tables[] holding `struct orc_table *` instead of device handles.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "midoridb_cuda.h"

struct orc_table {
	int ncols;
	int types[MDBCU_MAX_COLUMNS];
	size_t nrows, cap;
	int64_t *cells[MDBCU_MAX_COLUMNS]; /* raw 8-byte cells (double stored as bits) */
	uint8_t *nulls[MDBCU_MAX_COLUMNS];
};

struct orc_result {
	int ncols;
	int types[MDBCU_MAX_OUT];
	size_t nrows;
	int64_t *cells[MDBCU_MAX_OUT];
	uint8_t *nulls[MDBCU_MAX_OUT];
};

static char orc_err[256];

const char *orc_last_error(void)
{
	return orc_err;
}

static size_t col_width(int type)
{
	/* table_calc_column_space, src/primitive/column.c:255-293 */
	return type == MDBCU_CT_TINYINT ? 1 : 8;
}

struct orc_table *orc_table_create(int ncols, const int32_t *types)
{
	struct orc_table *t;

	if (ncols <= 0 || ncols > MDBCU_MAX_COLUMNS)
		return NULL;
	t = calloc(1, sizeof(*t));
	if (!t)
		return NULL;
	t->ncols = ncols;
	for (int i = 0; i < ncols; i++)
		t->types[i] = types[i];
	return t;
}

void orc_table_free(struct orc_table *t)
{
	if (!t)
		return;
	for (int i = 0; i < t->ncols; i++) {
		free(t->cells[i]);
		free(t->nulls[i]);
	}
	free(t);
}

static int orc_reserve(struct orc_table *t, size_t extra)
{
	size_t need = t->nrows + extra;

	if (need <= t->cap)
		return 0;
	size_t ncap = t->cap ? t->cap : 1024;
	while (ncap < need)
		ncap *= 2;
	for (int i = 0; i < t->ncols; i++) {
		int64_t *c = realloc(t->cells[i], ncap * sizeof(int64_t));
		if (!c)
			return -1;
		t->cells[i] = c;
		uint8_t *n = realloc(t->nulls[i], ncap);
		if (!n)
			return -1;
		t->nulls[i] = n;
	}
	t->cap = ncap;
	return 0;
}

size_t orc_table_rows(const struct orc_table *t)
{
	return t->nrows;
}

/* decode page images exactly like the executor's scan loops (executor_select.c:1299-1306) */
int orc_table_append_pages(struct orc_table *t, const void *pages, size_t n_pages, size_t stride)
{
	size_t row_size = MDBCU_ROW_HEADER;

	for (int c = 0; c < t->ncols; c++)
		row_size += col_width(t->types[c]);

	size_t slots = MDBCU_PAGE_SIZE / row_size;

	for (size_t p = 0; p < n_pages; p++) {
		const uint8_t *page = (const uint8_t*)pages + p * stride;

		if (orc_reserve(t, slots))
			return MDBCU_ENOMEM;
		for (size_t s = 0; s < slots; s++) {
			const uint8_t *row = page + s * row_size;

			if (row[0]) /* flags.empty */
				break;
			if (row[1]) /* flags.deleted */
				continue;
			size_t off = MDBCU_ROW_HEADER;
			for (int c = 0; c < t->ncols; c++) {
				size_t w = col_width(t->types[c]);
				int64_t v = 0;
				int isnull = (row[MDBCU_NULL_BITMAP_OFF + c / 8] >> (c % 8)) & 1; /* bit_test, src/lib/bit.c:3 */

				memcpy(&v, row + off, w);
				t->cells[c][t->nrows] = isnull ? 0 : v;
				t->nulls[c][t->nrows] = (uint8_t)isnull;
				off += w;
			}
			t->nrows++;
		}
	}
	return MDBCU_OK;
}

int orc_table_append_columns(struct orc_table *t, size_t n_rows, const void *const *col_data,
		const uint8_t *const *col_nulls)
{
	if (orc_reserve(t, n_rows))
		return MDBCU_ENOMEM;
	for (int c = 0; c < t->ncols; c++) {
		memcpy(t->cells[c] + t->nrows, col_data[c], n_rows * 8);
		if (col_nulls && col_nulls[c]) {
			memcpy(t->nulls[c] + t->nrows, col_nulls[c], n_rows);
			for (size_t r = 0; r < n_rows; r++)
				if (t->nulls[c][t->nrows + r])
					t->cells[c][t->nrows + r] = 0;
		} else {
			memset(t->nulls[c] + t->nrows, 0, n_rows);
		}
	}
	t->nrows += n_rows;
	return MDBCU_OK;
}

/* ------------------------------------------------------------------ tuples */

struct tuples {
	int ntab;
	size_t n, cap;
	uint32_t *rid[MDBCU_MAX_TABLES];
};

static int tuples_push(struct tuples *ts, const uint32_t *rids)
{
	if (ts->n == ts->cap) {
		size_t ncap = ts->cap ? ts->cap * 2 : 1024;
		for (int i = 0; i < ts->ntab; i++) {
			uint32_t *r = realloc(ts->rid[i], ncap * sizeof(uint32_t));
			if (!r)
				return -1;
			ts->rid[i] = r;
		}
		ts->cap = ncap;
	}
	for (int i = 0; i < ts->ntab; i++)
		ts->rid[i][ts->n] = rids[i];
	ts->n++;
	return 0;
}

static void tuples_free(struct tuples *ts)
{
	for (int i = 0; i < MDBCU_MAX_TABLES; i++)
		free(ts->rid[i]);
	memset(ts, 0, sizeof(*ts));
}

static uint64_t mix64(uint64_t x)
{
	x ^= x >> 33;
	x *= 0xff51afd7ed558ccdULL;
	x ^= x >> 33;
	x *= 0xc4ceb9fe1a85ec53ULL;
	x ^= x >> 33;
	return x;
}

/* typed scalar on the predicate stack */
struct val {
	int kind; /* 0 int, 1 double, 2 null, 3 bool */
	int64_t i;
	double d;
};

static int is_dbl_type(int type)
{
	return type == MDBCU_CT_DOUBLE;
}

static int cmp_vals(int cmp, struct val a, struct val b)
{
	if (a.kind == 2 || b.kind == 2)
		return 0; /* NULL operand: never true, executor_select.c:629-631 */
	if (a.kind == 1 || b.kind == 1) {
		double x = a.kind == 1 ? a.d : (double)a.i;
		double y = b.kind == 1 ? b.d : (double)b.i;
		switch (cmp) { /* cmp_double_value_to_value, executor_select.c:440 */
		case 1: return x < y;
		case 2: return x > y;
		case 3: return x != y;
		case 4: return x == y;
		case 5: return x <= y;
		case 6: return x >= y;
		}
		return 0;
	}
	switch (cmp) { /* cmp_int_value_to_value, executor_select.c:462 (full 64-bit here, see D5) */
	case 1: return a.i < b.i;
	case 2: return a.i > b.i;
	case 3: return a.i != b.i;
	case 4: return a.i == b.i;
	case 5: return a.i <= b.i;
	case 6: return a.i >= b.i;
	}
	return 0;
}

static struct val load_val(const struct mdbcu_plan *plan, const struct tuples *ts, size_t i, int tbl, int col)
{
	const struct orc_table *t = (const struct orc_table*)plan->tables[tbl];
	uint32_t r = ts->rid[tbl][i];
	struct val v = {0};

	if (t->nulls[col][r]) {
		v.kind = 2;
	} else if (is_dbl_type(t->types[col])) {
		v.kind = 1;
		memcpy(&v.d, &t->cells[col][r], 8);
	} else {
		v.kind = 0;
		v.i = t->cells[col][r];
	}
	return v;
}

/* eval_row_cond, executor_select.c:1027: postfix evaluation, boolean results are 2-valued */
static int eval_pred(const struct mdbcu_plan *plan, const struct tuples *ts, size_t i)
{
	struct val st[MDBCU_MAX_PRED];
	int sp = 0;

	for (int k = 0; k < plan->n_pred; k++) {
		const struct mdbcu_pred_op *op = &plan->pred[k];
		struct val v = {0};

		switch (op->op) {
		case MDBCU_P_COL:
			st[sp++] = load_val(plan, ts, i, op->tbl, op->col);
			break;
		case MDBCU_P_INT:
			v.kind = 0; v.i = op->ival; st[sp++] = v;
			break;
		case MDBCU_P_DBL:
			v.kind = 1; v.d = op->dval; st[sp++] = v;
			break;
		case MDBCU_P_NULL:
			v.kind = 2; st[sp++] = v;
			break;
		case MDBCU_P_BOOL:
			v.kind = 3; v.i = op->ival != 0; st[sp++] = v;
			break;
		case MDBCU_P_CMP: {
			struct val b = st[--sp], a = st[--sp];
			v.kind = 3; v.i = cmp_vals(op->arg, a, b); st[sp++] = v;
			break;
		}
		case MDBCU_P_AND: case MDBCU_P_OR: case MDBCU_P_XOR: {
			struct val b = st[--sp], a = st[--sp];
			int x = a.i != 0, y = b.i != 0;
			v.kind = 3;
			v.i = op->op == MDBCU_P_AND ? (x && y) : (op->op == MDBCU_P_OR ? (x || y) : (x ^ y));
			st[sp++] = v;
			break;
		}
		case MDBCU_P_ISNULL: case MDBCU_P_ISNOTNULL: {
			struct val a = st[--sp]; /* eval_isxnull, executor_select.c:923 */
			v.kind = 3; v.i = (a.kind == 2) ^ (op->op == MDBCU_P_ISNOTNULL); st[sp++] = v;
			break;
		}
		case MDBCU_P_IN: case MDBCU_P_NOTIN: {
			int n = op->arg, any = 0, all_diff = 1;
			struct val probe = st[sp - n - 1];
			for (int j = 0; j < n; j++) {
				struct val e = st[sp - n + j];
				if (cmp_vals(4, probe, e))
					any = 1;
				if (!cmp_vals(3, probe, e))
					all_diff = 0;
			}
			sp -= n + 1;
			v.kind = 3; v.i = op->op == MDBCU_P_IN ? any : all_diff; st[sp++] = v;
			break;
		}
		default:
			return 0;
		}
	}
	return sp == 1 && st[0].i != 0;
}

/* ------------------------------------------------------------------ join */

struct hidx { /* chained hash index on one column, chains keep storage order */
	size_t nb;
	int64_t *head, *next;
};

static int hidx_build(struct hidx *h, const struct orc_table *t, int col)
{
	size_t nb = 16;

	while (nb < t->nrows * 2)
		nb *= 2;
	h->nb = nb;
	h->head = malloc(nb * sizeof(int64_t));
	h->next = malloc((t->nrows + 1) * sizeof(int64_t));
	if (!h->head || !h->next)
		return -1;
	for (size_t i = 0; i < nb; i++)
		h->head[i] = -1;
	for (size_t r = t->nrows; r-- > 0;) { /* reverse insert at head => forward chains */
		if (t->nulls[col][r])
			continue;
		size_t b = mix64((uint64_t)t->cells[col][r]) & (nb - 1);
		h->next[r] = h->head[b];
		h->head[b] = (int64_t)r;
	}
	return 0;
}

static int join_step(const struct mdbcu_plan *plan, int j, struct tuples *in, struct tuples *out)
{
	const struct mdbcu_join *jn = &plan->joins[j];
	const struct orc_table *rt = (const struct orc_table*)plan->tables[j + 1];
	uint32_t rids[MDBCU_MAX_TABLES];
	int rc = 0;

	out->ntab = in->ntab + 1;

	if (jn->cross) {
		for (size_t i = 0; i < in->n && !rc; i++) {
			for (int k = 0; k < in->ntab; k++)
				rids[k] = in->rid[k][i];
			for (size_t r = 0; r < rt->nrows && !rc; r++) {
				rids[in->ntab] = (uint32_t)r;
				rc = tuples_push(out, rids);
			}
		}
		return rc;
	}

	const struct orc_table *lt = (const struct orc_table*)plan->tables[jn->left.tbl];
	int ldbl = is_dbl_type(lt->types[jn->left.col]), rdbl = is_dbl_type(rt->types[jn->right.col]);
	struct hidx h = {0};

	if (ldbl != rdbl) {
		snprintf(orc_err, sizeof(orc_err), "join key types differ");
		return -1;
	}
	if (hidx_build(&h, rt, jn->right.col))
		return -1;

	for (size_t i = 0; i < in->n && !rc; i++) {
		uint32_t lr = in->rid[jn->left.tbl][i];

		if (lt->nulls[jn->left.col][lr])
			continue; /* NULL never equals anything, executor_select.c:716-738 */
		int64_t key = lt->cells[jn->left.col][lr];
		for (int k = 0; k < in->ntab; k++)
			rids[k] = in->rid[k][i];
		for (int64_t r = h.head[mix64((uint64_t)key) & (h.nb - 1)]; r >= 0 && !rc; r = h.next[r]) {
			int eq;
			if (ldbl) {
				double a, b;
				memcpy(&a, &key, 8);
				memcpy(&b, &rt->cells[jn->right.col][r], 8);
				eq = a == b;
			} else {
				eq = rt->cells[jn->right.col][r] == key;
			}
			if (eq) {
				rids[in->ntab] = (uint32_t)r;
				rc = tuples_push(out, rids);
			}
		}
	}
	free(h.head);
	free(h.next);
	return rc;
}

/* ------------------------------------------------------------------ group by / aggregate */

struct agg_state {
	int64_t count;    /* COUNT(*) */
	int64_t nn;       /* non-NULL inputs */
	int64_t isum;
	double dsum;
	int64_t imin, imax;
	double dmin, dmax;
};

struct group {
	size_t first; /* tuple index of the survivor row */
	struct agg_state *st; /* one per output column */
};

static void agg_update(const struct mdbcu_plan *plan, const struct tuples *ts, size_t i, struct agg_state *st)
{
	for (int o = 0; o < plan->n_out; o++) {
		const struct mdbcu_out *out = &plan->out[o];
		struct agg_state *s = &st[o];

		s->count++;
		if (out->kind == MDBCU_OUT_COLUMN || out->kind == MDBCU_OUT_COUNT_STAR)
			continue;
		struct val v = load_val(plan, ts, i, out->ref.tbl, out->ref.col);
		if (v.kind == 2)
			continue;
		if (v.kind == 1) {
			if (!s->nn || v.d < s->dmin) s->dmin = v.d;
			if (!s->nn || v.d > s->dmax) s->dmax = v.d;
			s->dsum += v.d;
		} else {
			if (!s->nn || v.i < s->imin) s->imin = v.i;
			if (!s->nn || v.i > s->imax) s->imax = v.i;
			s->isum = (int64_t)((uint64_t)s->isum + (uint64_t)v.i);
		}
		s->nn++;
	}
}

static int out_type(const struct mdbcu_plan *plan, int o)
{
	const struct mdbcu_out *out = &plan->out[o];
	const struct orc_table *t;

	if (out->kind == MDBCU_OUT_COUNT_STAR || out->kind == MDBCU_OUT_COUNT_COL)
		return MDBCU_CT_INTEGER;
	if (out->kind == MDBCU_OUT_AVG)
		return MDBCU_CT_DOUBLE;
	t = (const struct orc_table*)plan->tables[out->ref.tbl];
	if (out->kind == MDBCU_OUT_COLUMN)
		return t->types[out->ref.col];
	return is_dbl_type(t->types[out->ref.col]) ? MDBCU_CT_DOUBLE : MDBCU_CT_INTEGER;
}

static void emit_row(const struct mdbcu_plan *plan, const struct tuples *ts, size_t first,
		const struct agg_state *st, struct orc_result *res, size_t row)
{
	for (int o = 0; o < plan->n_out; o++) {
		const struct mdbcu_out *out = &plan->out[o];
		int64_t cell = 0;
		uint8_t isnull = 0;
		int dbl = 0;

		if (out->kind != MDBCU_OUT_COUNT_STAR) {
			const struct orc_table *t = (const struct orc_table*)plan->tables[out->ref.tbl];
			dbl = is_dbl_type(t->types[out->ref.col]);
		}

		switch (out->kind) {
		case MDBCU_OUT_COLUMN: {
			const struct orc_table *t = (const struct orc_table*)plan->tables[out->ref.tbl];
			uint32_t r = ts->rid[out->ref.tbl][first];
			cell = t->cells[out->ref.col][r];
			isnull = t->nulls[out->ref.col][r];
			break;
		}
		case MDBCU_OUT_COUNT_STAR:
			cell = st[o].count;
			break;
		case MDBCU_OUT_COUNT_COL:
			cell = st[o].nn;
			break;
		case MDBCU_OUT_SUM:
			if (!st[o].nn) isnull = 1;
			else if (dbl) memcpy(&cell, &st[o].dsum, 8);
			else cell = st[o].isum;
			break;
		case MDBCU_OUT_MIN:
			if (!st[o].nn) isnull = 1;
			else if (dbl) memcpy(&cell, &st[o].dmin, 8);
			else cell = st[o].imin;
			break;
		case MDBCU_OUT_MAX:
			if (!st[o].nn) isnull = 1;
			else if (dbl) memcpy(&cell, &st[o].dmax, 8);
			else cell = st[o].imax;
			break;
		case MDBCU_OUT_AVG:
			if (!st[o].nn) {
				isnull = 1;
			} else {
				double avg = (dbl ? st[o].dsum : (double)st[o].isum) / (double)st[o].nn;
				memcpy(&cell, &avg, 8);
			}
			break;
		}
		res->cells[o][row] = isnull ? 0 : cell;
		res->nulls[o][row] = isnull;
	}
}

static int res_alloc(struct orc_result *res, const struct mdbcu_plan *plan, size_t nrows)
{
	res->ncols = plan->n_out;
	res->nrows = nrows;
	for (int o = 0; o < plan->n_out; o++) {
		res->types[o] = out_type(plan, o);
		res->cells[o] = calloc(nrows ? nrows : 1, 8);
		res->nulls[o] = calloc(nrows ? nrows : 1, 1);
		if (!res->cells[o] || !res->nulls[o])
			return -1;
	}
	return 0;
}

static int has_aggregate(const struct mdbcu_plan *plan)
{
	for (int o = 0; o < plan->n_out; o++)
		if (plan->out[o].kind != MDBCU_OUT_COLUMN)
			return 1;
	return 0;
}

static int group_key(const struct mdbcu_plan *plan, const struct tuples *ts, size_t i, int64_t *key, uint8_t *knull)
{
	for (int g = 0; g < plan->n_group; g++) {
		const struct orc_table *t = (const struct orc_table*)plan->tables[plan->group[g].tbl];
		uint32_t r = ts->rid[plan->group[g].tbl][i];
		knull[g] = t->nulls[plan->group[g].col][r];
		key[g] = knull[g] ? 0 : t->cells[plan->group[g].col][r];
		/* -0.0 and +0.0 compare equal in the reference's double difference (executor_select.c:1484) */
		if (!knull[g] && is_dbl_type(t->types[plan->group[g].col]) && key[g] == (int64_t)0x8000000000000000LL)
			key[g] = 0;
	}
	return 0;
}


/* ------------------------------------------------------------------ tail operators: HAVING, DISTINCT, ORDER BY, LIMIT
 * The reference parses and validates them (src/parser/midorisql.y:180-196,203; semantic_select.c:1895,2004) and never
 * executes them (executor_select.c:1723): PARITY UNPINNED BY THE REFERENCE.  The semantics restated here
 * are SQL's as sqlite3 implements them - tests/test_oracle.py checks this code against sqlite3 itself:
 *   HAVING    filters result rows; a comparison with a NULL operand is not true;
 *   DISTINCT  one row per combination of ALL result columns, NULL equal to NULL, -0.0 equal to 0.0;
 *   ORDER BY  NULLs first ascending / last descending; ties in no particular order (this code: stable);
 *   LIMIT     [offset,] count, applied last. */
static struct val res_val(const struct orc_result *r, int c, size_t row)
{
	struct val v;
	memset(&v, 0, sizeof(v));
	if (r->nulls[c][row]) {
		v.kind = 2;
	} else if (is_dbl_type(r->types[c])) {
		v.kind = 1;
		memcpy(&v.d, &r->cells[c][row], 8);
	} else {
		v.kind = 0;
		v.i = r->cells[c][row];
	}
	return v;
}

static int eval_having(const struct mdbcu_plan *plan, const struct orc_result *r, size_t row)
{
	struct val st[MDBCU_MAX_HAVING + 1];
	int sp = 0;
	for (int k = 0; k < plan->n_having; k++) {
		const struct mdbcu_pred_op *op = &plan->having[k];
		struct val v;
		memset(&v, 0, sizeof(v));
		v.kind = 3;
		switch (op->op) {
		case MDBCU_P_OUT: st[sp++] = res_val(r, op->col, row); break;
		case MDBCU_P_INT: v.kind = 0; v.i = op->ival; st[sp++] = v; break;
		case MDBCU_P_DBL: v.kind = 1; v.d = op->dval; st[sp++] = v; break;
		case MDBCU_P_NULL: v.kind = 2; st[sp++] = v; break;
		case MDBCU_P_BOOL: v.i = op->ival != 0; st[sp++] = v; break;
		case MDBCU_P_CMP: {
			struct val b = st[--sp], a = st[--sp];
			v.i = cmp_vals(op->arg, a, b);
			st[sp++] = v;
			break;
		}
		case MDBCU_P_AND: case MDBCU_P_OR: case MDBCU_P_XOR: {
			struct val b = st[--sp], a = st[--sp];
			int x = a.i != 0, y = b.i != 0;
			v.i = op->op == MDBCU_P_AND ? (x && y) : (op->op == MDBCU_P_OR ? (x || y) : (x != y));
			st[sp++] = v;
			break;
		}
		case MDBCU_P_ISNULL: case MDBCU_P_ISNOTNULL: {
			struct val a = st[--sp];
			v.i = (a.kind == 2) != (op->op == MDBCU_P_ISNOTNULL);
			st[sp++] = v;
			break;
		}
		case MDBCU_P_IN: case MDBCU_P_NOTIN: {
			int n = op->arg, any = 0, all_diff = 1;
			struct val probe = st[sp - n - 1];
			for (int j = 0; j < n; j++) {
				any = any || cmp_vals(4, probe, st[sp - n + j]);
				all_diff = all_diff && cmp_vals(3, probe, st[sp - n + j]);
			}
			sp -= n + 1;
			v.i = op->op == MDBCU_P_IN ? any : all_diff;
			st[sp++] = v;
			break;
		}
		}
	}
	return sp == 1 && st[0].i != 0;
}

/* three-way comparison of two result cells of column c: NULL sorts before every value */
static int cmp_cells(const struct orc_result *r, int c, size_t a, size_t b)
{
	int na = r->nulls[c][a], nb = r->nulls[c][b];
	if (na || nb)
		return nb - na; /* NULL < value; NULL = NULL */
	if (is_dbl_type(r->types[c])) {
		double x, y;
		memcpy(&x, &r->cells[c][a], 8);
		memcpy(&y, &r->cells[c][b], 8);
		return x < y ? -1 : (x > y ? 1 : 0);
	}
	return r->cells[c][a] < r->cells[c][b] ? -1 : (r->cells[c][a] > r->cells[c][b] ? 1 : 0);
}

struct sort_spec {
	const struct orc_result *r;
	int ncols;
	int col[MDBCU_MAX_OUT];
	int desc[MDBCU_MAX_OUT];
};

static int cmp_rows(const struct sort_spec *sp, size_t a, size_t b)
{
	for (int k = 0; k < sp->ncols; k++) {
		int c = cmp_cells(sp->r, sp->col[k], a, b);
		if (c)
			return sp->desc[k] ? -c : c; /* descending: values reversed, NULLs last */
	}
	return 0;
}

/* stable merge sort of row numbers */
static int sort_rows(const struct sort_spec *sp, size_t *idx, size_t n)
{
	size_t *tmp = malloc((n ? n : 1) * sizeof(*tmp));
	if (!tmp)
		return -1;
	for (size_t w = 1; w < n; w *= 2) {
		for (size_t lo = 0; lo < n; lo += 2 * w) {
			size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, i = lo, j = mid, k = lo;
			while (i < mid && j < hi)
				tmp[k++] = cmp_rows(sp, idx[j], idx[i]) < 0 ? idx[j++] : idx[i++];
			while (i < mid)
				tmp[k++] = idx[i++];
			while (j < hi)
				tmp[k++] = idx[j++];
		}
		memcpy(idx, tmp, n * sizeof(*idx));
	}
	free(tmp);
	return 0;
}

static int apply_tail(const struct mdbcu_plan *plan, struct orc_result *r)
{
	size_t n = r->nrows, m = 0;
	size_t *idx = malloc((n ? n : 1) * sizeof(*idx));
	struct sort_spec sp;
	if (!idx)
		return -1;
	for (size_t i = 0; i < n; i++)
		if (plan->n_having == 0 || eval_having(plan, r, i))
			idx[m++] = i;
	n = m;
	memset(&sp, 0, sizeof(sp));
	sp.r = r;
	if (plan->distinct && n > 1) {
		sp.ncols = r->ncols;
		for (int c = 0; c < r->ncols; c++)
			sp.col[c] = c;
		if (sort_rows(&sp, idx, n)) {
			free(idx);
			return -1;
		}
		m = 1;
		for (size_t i = 1; i < n; i++)
			if (cmp_rows(&sp, idx[i], idx[m - 1]) != 0)
				idx[m++] = idx[i];
		n = m;
	}
	if (plan->n_order > 0) {
		memset(&sp, 0, sizeof(sp));
		sp.r = r;
		sp.ncols = plan->n_order;
		for (int k = 0; k < plan->n_order; k++) {
			sp.col[k] = plan->order[k].out_col;
			sp.desc[k] = plan->order[k].desc != 0;
		}
		if (sort_rows(&sp, idx, n)) {
			free(idx);
			return -1;
		}
	}
	size_t first = 0, count = n;
	if (plan->has_limit) {
		first = (size_t)plan->offset < n ? (size_t)plan->offset : n;
		count = (size_t)plan->limit < n - first ? (size_t)plan->limit : n - first;
	}
	for (int c = 0; c < r->ncols; c++) {
		int64_t *nc = malloc((count ? count : 1) * 8);
		uint8_t *nn = malloc(count ? count : 1);
		if (!nc || !nn) {
			free(nc);
			free(nn);
			free(idx);
			return -1;
		}
		for (size_t i = 0; i < count; i++) {
			nc[i] = r->cells[c][idx[first + i]];
			nn[i] = r->nulls[c][idx[first + i]];
		}
		free(r->cells[c]);
		free(r->nulls[c]);
		r->cells[c] = nc;
		r->nulls[c] = nn;
	}
	r->nrows = count;
	free(idx);
	return 0;
}

void orc_result_free(struct orc_result *res);

int orc_select(const struct mdbcu_plan *plan, struct orc_result **out_res)
{
	struct tuples cur = {0}, nxt = {0};
	struct orc_result *res = NULL;
	int rc = MDBCU_EINTERNAL;

	orc_err[0] = 0;
	*out_res = NULL;
	if (plan->n_tables < 1 || plan->n_tables > MDBCU_MAX_TABLES || plan->n_joins != plan->n_tables - 1) {
		snprintf(orc_err, sizeof(orc_err), "bad plan");
		return MDBCU_EERROR;
	}

	/* FROM: first table, storage order */
	{
		const struct orc_table *t0 = (const struct orc_table*)plan->tables[0];
		uint32_t rid;

		cur.ntab = 1;
		for (size_t r = 0; r < t0->nrows; r++) {
			rid = (uint32_t)r;
			if (tuples_push(&cur, &rid))
				goto nomem;
		}
	}
	for (int j = 0; j < plan->n_joins; j++) {
		if (join_step(plan, j, &cur, &nxt)) {
			if (orc_err[0]) {
				rc = MDBCU_EERROR;
				goto out;
			}
			goto nomem;
		}
		tuples_free(&cur);
		cur = nxt;
		memset(&nxt, 0, sizeof(nxt));
	}

	/* WHERE */
	if (plan->n_pred) {
		size_t w = 0;
		for (size_t i = 0; i < cur.n; i++) {
			if (eval_pred(plan, &cur, i)) {
				for (int k = 0; k < cur.ntab; k++)
					cur.rid[k][w] = cur.rid[k][i];
				w++;
			}
		}
		cur.n = w;
	}

	res = calloc(1, sizeof(*res));
	if (!res)
		goto nomem;

	if (plan->n_group > 0) {
		/* GROUP BY: open-addressing table of groups, output in first-occurrence order */
		size_t nb = 16, ngroups = 0, gcap = 1024;
		int64_t *slots;
		struct group *groups;

		while (nb < cur.n * 2)
			nb *= 2;
		slots = malloc(nb * sizeof(int64_t));
		groups = malloc(gcap * sizeof(*groups));
		if (!slots || !groups)
			goto nomem;
		for (size_t i = 0; i < nb; i++)
			slots[i] = -1;

		for (size_t i = 0; i < cur.n; i++) {
			int64_t key[MDBCU_MAX_GROUP] = {0}, k2[MDBCU_MAX_GROUP] = {0};
			uint8_t kn[MDBCU_MAX_GROUP] = {0}, kn2[MDBCU_MAX_GROUP] = {0};
			uint64_t h = 0x9e3779b97f4a7c15ULL;

			group_key(plan, &cur, i, key, kn);
			for (int g = 0; g < plan->n_group; g++)
				h = mix64(h ^ (uint64_t)key[g] ^ ((uint64_t)kn[g] << 63 >> (g + 1)));
			size_t b = h & (nb - 1);
			int64_t gi;
			for (;; b = (b + 1) & (nb - 1)) {
				gi = slots[b];
				if (gi < 0)
					break;
				group_key(plan, &cur, groups[gi].first, k2, kn2);
				if (memcmp(key, k2, sizeof(int64_t) * plan->n_group) == 0 &&
						memcmp(kn, kn2, plan->n_group) == 0)
					break;
			}
			if (gi < 0) {
				if (ngroups == gcap) {
					struct group *ng = realloc(groups, gcap * 2 * sizeof(*groups));
					if (!ng)
						goto nomem;
					groups = ng;
					gcap *= 2;
				}
				gi = (int64_t)ngroups++;
				slots[b] = gi;
				groups[gi].first = i;
				groups[gi].st = calloc(plan->n_out, sizeof(struct agg_state));
				if (!groups[gi].st)
					goto nomem;
			}
			agg_update(plan, &cur, i, groups[gi].st);
		}
		if (res_alloc(res, plan, ngroups))
			goto nomem;
		for (size_t g = 0; g < ngroups; g++) {
			emit_row(plan, &cur, groups[g].first, groups[g].st, res, g);
			free(groups[g].st);
		}
		free(groups);
		free(slots);
	} else if (has_aggregate(plan)) {
		/* handle_countonly_case (executor_select.c:1590) generalised to all aggregates:
		 * one row, but NO row when nothing qualifies (the reference keeps zero rows) */
		struct agg_state st[MDBCU_MAX_OUT];

		memset(st, 0, sizeof(st));
		for (size_t i = 0; i < cur.n; i++)
			agg_update(plan, &cur, i, st);
		if (res_alloc(res, plan, cur.n ? 1 : 0))
			goto nomem;
		if (cur.n)
			emit_row(plan, &cur, 0, st, res, 0);
	} else {
		struct agg_state st[MDBCU_MAX_OUT];

		memset(st, 0, sizeof(st));
		if (res_alloc(res, plan, cur.n))
			goto nomem;
		for (size_t i = 0; i < cur.n; i++)
			emit_row(plan, &cur, i, st, res, i);
	}

	if (plan->distinct || plan->n_having > 0 || plan->n_order > 0 || plan->has_limit)
		if (apply_tail(plan, res))
			goto nomem;
	*out_res = res;
	res = NULL;
	rc = MDBCU_OK;
	goto out;
nomem:
	snprintf(orc_err, sizeof(orc_err), "out of memory");
	rc = MDBCU_ENOMEM;
out:
	tuples_free(&cur);
	tuples_free(&nxt);
	if (res)
		orc_result_free(res);
	return rc;
}

size_t orc_result_rows(const struct orc_result *r)
{
	return r->nrows;
}

int orc_result_cols(const struct orc_result *r)
{
	return r->ncols;
}

int orc_result_col_type(const struct orc_result *r, int c)
{
	return r->types[c];
}

int orc_result_fetch_columns(const struct orc_result *r, void *const *cells, uint8_t *const *nulls)
{
	for (int c = 0; c < r->ncols; c++) {
		if (cells && cells[c])
			memcpy(cells[c], r->cells[c], r->nrows * 8);
		if (nulls && nulls[c])
			memcpy(nulls[c], r->nulls[c], r->nrows);
	}
	return MDBCU_OK;
}

void orc_result_free(struct orc_result *r)
{
	if (!r)
		return;
	for (int c = 0; c < MDBCU_MAX_OUT; c++) {
		free(r->cells[c]);
		free(r->nulls[c]);
	}
	free(r);
}

/* ------------------------------------------------------------------ cpu_baseline helpers (bench.py)
 * single-thread hash join + count of the README query over raw key arrays: the "port" baseline
 * for sizes the nested-loop reference cannot finish. */
int64_t orc_join_count_groups(const int64_t *a, size_t na, const int64_t *b, size_t nb_rows,
		int64_t *out_keys, int64_t *out_counts, size_t cap)
{
	size_t nb = 16;
	int64_t *keys, *ca, *cb;
	uint8_t *used;
	int64_t ngroups = 0;

	while (nb < (na < nb_rows ? na : nb_rows) * 2)
		nb *= 2;
	/* table over the smaller distinct-set bound is not known: size by na */
	while (nb < na * 2)
		nb *= 2;
	keys = malloc(nb * 8);
	ca = calloc(nb, 8);
	cb = calloc(nb, 8);
	used = calloc(nb, 1);
	if (!keys || !ca || !cb || !used) {
		free(keys); free(ca); free(cb); free(used);
		return -1;
	}
	for (size_t i = 0; i < na; i++) {
		size_t s = mix64((uint64_t)a[i]) & (nb - 1);
		while (used[s] && keys[s] != a[i])
			s = (s + 1) & (nb - 1);
		used[s] = 1;
		keys[s] = a[i];
		ca[s]++;
	}
	for (size_t i = 0; i < nb_rows; i++) {
		size_t s = mix64((uint64_t)b[i]) & (nb - 1);
		while (used[s] && keys[s] != b[i])
			s = (s + 1) & (nb - 1);
		if (used[s])
			cb[s]++;
	}
	for (size_t s = 0; s < nb; s++) {
		if (used[s] && cb[s]) {
			if ((size_t)ngroups < cap && out_keys && out_counts) {
				out_keys[ngroups] = keys[s];
				out_counts[ngroups] = ca[s] * cb[s];
			}
			ngroups++;
		}
	}
	free(keys); free(ca); free(cb); free(used);
	return ngroups;
}
