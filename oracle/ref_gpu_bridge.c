/*
 * ref_gpu_bridge.c - TEST INFRASTRUCTURE (oracle/): the binding INTEGRATION.md section B describes, as real code.
 *
 * Linked with the UNMODIFIED reference objects and `-Wl,--wrap=executor_run_select_stmt` (oracle/Makefile, target
 * `refgpu`, output oracle/_ref/libmidoridb_ref_gpu.so), this file replaces the one call the reference's executor
 * dispatch makes for SELECT (src/engine/executor.c:23-24 -> src/engine/executor_select.c:1655): the reference's own
 * parser stand-in, AST builder, semantic analysis and optimiser run as always, then
 *
 *   __wrap_executor_run_select_stmt   lowers the reference's OPTIMISED AST (include/parser/ast.h:456-720) to a
 *                                     `struct mdbcu_plan`, mirrors the FROM tables' datablocks on the device
 *                                     (mdbcu_table_append_page_ptrs on the reference's own malloc'd pages), calls
 *                                     mdbcu_select, and hands the result back as a reference `struct table` whose
 *                                     datablocks are the page images mdbcu_result_fetch_pages produced -
 *
 * so query_cur_step / query_column_int64 (src/engine/query.c:108,148) and the page-walk idiom of the reference's tests
 * read GPU results through untouched reference code.  tests/test_gpu_bridge.py runs the golden cases through it.
 *
 * What a maintainer would do differently in production: keep the mirrors alive between queries and feed them from the
 * INSERT / UPDATE / DELETE executors instead of re-mirroring per SELECT as this harness does.
 * Nothing here is shipped in the product.
 */
#define _GNU_SOURCE
#include <engine/executor.h>
#include <engine/query.h>
#include <engine/database.h>
#include <parser/ast.h>
#include <primitive/table.h>
#include <primitive/row.h>
#include <primitive/column.h>
#include <primitive/datablock.h>
#include <datastructure/hashtable.h>
#include <datastructure/linkedlist.h>

#include "../include/midoridb_cuda.h"

#define FQ_LEN (TABLE_MAX_NAME + 1 + TABLE_MAX_COLUMN_NAME + 1)

static mdbcu_ctx *bridge_ctx(void)
{
	static mdbcu_ctx *ctx;

	if (!ctx) {
		const char *dev = getenv("MIDORIDB_CUDA_DEVICE");

		if (mdbcu_init(dev ? atoi(dev) : 0, &ctx) != MDBCU_OK)
			ctx = NULL;
	}
	return ctx;
}

struct lower {
	struct database *db;
	int ntab;
	struct table *tab[MDBCU_MAX_TABLES];
	char name[MDBCU_MAX_TABLES][TABLE_MAX_NAME + 1];
	struct mdbcu_plan plan;
	char err[256];
};

#define FAIL(L, ...) (snprintf((L)->err, sizeof((L)->err), __VA_ARGS__), -1)

static int tab_index(struct lower *L, const char *name)
{
	for (int t = 0; t < L->ntab; t++)
		if (strcmp(L->name[t], name) == 0)
			return t;
	return -1;
}

static int col_index(struct table *t, const char *col)
{
	for (int c = 0; c < t->column_count; c++)
		if (strcmp(t->columns[c].name, col) == 0)
			return c;
	return -1;
}

/* "T.c" or (unqualified) "c" -> (table, column) */
static int resolve_name(struct lower *L, const char *fq, int *tbl, int *col)
{
	const char *dot = strchr(fq, '.');

	if (dot) {
		char tname[TABLE_MAX_NAME + 1] = {0};

		if ((size_t)(dot - fq) > TABLE_MAX_NAME)
			return FAIL(L, "name too long: %s", fq);
		memcpy(tname, fq, (size_t)(dot - fq));
		*tbl = tab_index(L, tname);
		if (*tbl < 0)
			return FAIL(L, "unknown table in %s", fq);
		*col = col_index(L->tab[*tbl], dot + 1);
		return *col < 0 ? FAIL(L, "unknown column %s", fq) : 0;
	}
	for (int t = 0; t < L->ntab; t++) {
		int c = col_index(L->tab[t], fq);

		if (c >= 0) {
			*tbl = t;
			*col = c;
			return 0;
		}
	}
	return FAIL(L, "unknown column %s", fq);
}

/* FIELDNAME node or EXPRVAL-with-a-name -> (table, column); 1 = the node is not a column reference */
static int resolve_node(struct lower *L, struct ast_node *n, int *tbl, int *col)
{
	if (n->node_type == AST_TYPE_SEL_FIELDNAME) {
		struct ast_sel_fieldname_node *f = (typeof(f))n;
		char fq[FQ_LEN];

		snprintf(fq, sizeof(fq), "%s.%s", f->table_name, f->col_name);
		return resolve_name(L, fq, tbl, col);
	}
	if (n->node_type == AST_TYPE_SEL_EXPRVAL && ((struct ast_sel_exprval_node*)n)->value_type.is_name)
		return resolve_name(L, ((struct ast_sel_exprval_node*)n)->name_val, tbl, col);
	return 1;
}

static int add_table(struct lower *L, struct ast_sel_table_node *tn)
{
	struct table *t = database_table_get(L->db, tn->table_name);

	if (!t)
		return FAIL(L, "table %s does not exist", tn->table_name);
	if (L->ntab == MDBCU_MAX_TABLES)
		return FAIL(L, "more than %d tables", MDBCU_MAX_TABLES);
	L->tab[L->ntab] = t;
	snprintf(L->name[L->ntab], sizeof(L->name[0]), "%s", tn->table_name);
	return L->ntab++;
}

static int push_pred(struct lower *L, int op, int arg, int tbl, int col, int64_t iv, double dv)
{
	struct mdbcu_pred_op *o;

	if (L->plan.n_pred >= MDBCU_MAX_PRED)
		return FAIL(L, "WHERE expression too long");
	o = &L->plan.pred[L->plan.n_pred++];
	memset(o, 0, sizeof(*o));
	o->op = op;
	o->arg = arg;
	o->tbl = tbl;
	o->col = col;
	o->ival = iv;
	o->dval = dv;
	return 0;
}

static int push_operand(struct lower *L, struct ast_node *n)
{
	int t, c, r = resolve_node(L, n, &t, &c);

	if (r < 0)
		return r;
	if (r == 0)
		return push_pred(L, MDBCU_P_COL, 0, t, c, 0, 0);
	if (n->node_type == AST_TYPE_SEL_EXPRVAL) {
		struct ast_sel_exprval_node *v = (typeof(v))n;

		if (v->value_type.is_null)
			return push_pred(L, MDBCU_P_NULL, 0, 0, 0, 0, 0);
		if (v->value_type.is_intnum)
			return push_pred(L, MDBCU_P_INT, 0, 0, 0, v->int_val, 0);
		if (v->value_type.is_approxnum)
			return push_pred(L, MDBCU_P_DBL, 0, 0, 0, 0, v->double_val);
		if (v->value_type.is_bool)
			return push_pred(L, MDBCU_P_INT, 0, 0, 0, v->bool_val, 0);
	}
	return FAIL(L, "unsupported operand in a condition");
}

/* a condition subtree (eval_row_cond, src/engine/executor_select.c:1027) -> postfix; leaves exactly one value on the stack */
static int lower_cond(struct lower *L, struct ast_node *n)
{
	struct list_head *pos;
	struct ast_node *kid, *kids[MDBCU_MAX_PRED];
	int nk = 0;

	list_for_each(pos, n->node_children_head)
	{
		kid = list_entry(pos, typeof(*kid), head);
		if (nk == MDBCU_MAX_PRED)
			return FAIL(L, "condition too wide");
		kids[nk++] = kid;
	}
	switch (n->node_type) {
	case AST_TYPE_SEL_CMP: {
		/* the reference compares its FIRST column operand with every literal operand (eval_cmp :985-1023): a column
		 * operand first, then the other operand, whatever side of the operator they were written on */
		int first_col = -1, t, c;

		if (nk != 2)
			return FAIL(L, "comparison with %d operands", nk);
		for (int k = 0; k < 2 && first_col < 0; k++)
			if (resolve_node(L, kids[k], &t, &c) == 0)
				first_col = k;
		if (first_col < 0)
			first_col = 0;
		if (push_operand(L, kids[first_col]) || push_operand(L, kids[1 - first_col]))
			return -1;
		return push_pred(L, MDBCU_P_CMP, ((struct ast_sel_cmp_node*)n)->cmp_type, 0, 0, 0, 0);
	}
	case AST_TYPE_SEL_LOGOP: {
		int op = ((struct ast_sel_logop_node*)n)->logop_type;

		if (nk < 1)
			return FAIL(L, "empty logical operation");
		for (int k = 0; k < nk; k++) {
			if (lower_cond(L, kids[k]))
				return -1;
			if (k > 0 && push_pred(L, op == AST_LOGOP_TYPE_AND ? MDBCU_P_AND : op == AST_LOGOP_TYPE_OR ? MDBCU_P_OR : MDBCU_P_XOR,
						0, 0, 0, 0, 0))
				return -1;
		}
		return 0;
	}
	case AST_TYPE_SEL_EXPRISXNULL:
		if (nk != 1 || push_operand(L, kids[0]))
			return L->err[0] ? -1 : FAIL(L, "malformed IS NULL");
		return push_pred(L, ((struct ast_sel_isxnull_node*)n)->is_negation ? MDBCU_P_ISNOTNULL : MDBCU_P_ISNULL, 0, 0, 0, 0, 0);
	case AST_TYPE_SEL_EXPRISXIN: {
		int probe = -1, t, c;

		for (int k = 0; k < nk && probe < 0; k++)
			if (resolve_node(L, kids[k], &t, &c) == 0)
				probe = k;
		if (probe < 0 || nk < 2)
			return FAIL(L, "malformed IN");
		if (push_operand(L, kids[probe]))
			return -1;
		for (int k = 0; k < nk; k++)
			if (k != probe && push_operand(L, kids[k]))
				return -1;
		return push_pred(L, ((struct ast_sel_isxin_node*)n)->is_negation ? MDBCU_P_NOTIN : MDBCU_P_IN, nk - 1, 0, 0, 0, 0);
	}
	default:
		/* wrappers (WHERE, ONEXPR): the conjunction of their children */
		if (nk < 1)
			return FAIL(L, "empty condition");
		for (int k = 0; k < nk; k++) {
			if (lower_cond(L, kids[k]))
				return -1;
			if (k > 0 && push_pred(L, MDBCU_P_AND, 0, 0, 0, 0, 0))
				return -1;
		}
		return 0;
	}
}

/* FROM: tables in the order the reference's nested loops bring them in (proc_from_clause_join :1232), one join each */
static int lower_from(struct lower *L, struct ast_node *n)
{
	struct list_head *pos;
	struct ast_node *kid, *left = NULL, *right = NULL, *on = NULL, *cmp = NULL, *a = NULL, *b = NULL;
	struct mdbcu_join *jn;
	int added, ta, ca, tb, cb;

	if (n->node_type == AST_TYPE_SEL_TABLE)
		return add_table(L, (struct ast_sel_table_node*)n) < 0 ? -1 : 0;
	if (n->node_type != AST_TYPE_SEL_JOIN)
		return FAIL(L, "unexpected node in FROM");
	list_for_each(pos, n->node_children_head)
	{
		kid = list_entry(pos, typeof(*kid), head);
		if (kid->node_type == AST_TYPE_SEL_ONEXPR)
			on = kid;
		else if (!left)
			left = kid;
		else
			right = kid;
	}
	if (!left || !right)
		return FAIL(L, "JOIN without two operands");
	if (left->node_type == AST_TYPE_SEL_TABLE && right->node_type == AST_TYPE_SEL_JOIN) {
		struct ast_node *tmp = left; /* the reference materialises the nested join first, then joins the table to it */

		left = right;
		right = tmp;
	}
	if (right->node_type != AST_TYPE_SEL_TABLE)
		return FAIL(L, "unsupported JOIN shape");
	if (lower_from(L, left))
		return -1;
	added = add_table(L, (struct ast_sel_table_node*)right);
	if (added < 0)
		return -1;
	jn = &L->plan.joins[added - 1];
	memset(jn, 0, sizeof(*jn));
	if (on)
		list_for_each(pos, on->node_children_head)
		{
			cmp = list_entry(pos, typeof(*cmp), head);
			break;
		}
	if (cmp && cmp->node_type == AST_TYPE_SEL_CMP)
		list_for_each(pos, cmp->node_children_head)
		{
			kid = list_entry(pos, typeof(*kid), head);
			if (!a)
				a = kid;
			else
				b = kid;
		}
	if (!cmp || !a || !b || resolve_node(L, a, &ta, &ca) != 0 || resolve_node(L, b, &tb, &cb) != 0) {
		L->err[0] = 0;
		jn->cross = 1; /* comma list / ON 1=1: the optimiser's synthetic join (optimiser_select.c:395) */
		return 0;
	}
	if (((struct ast_sel_cmp_node*)cmp)->cmp_type != AST_CMP_EQUALS_OP)
		return FAIL(L, "only ON <column> = <column> joins run on the GPU path");
	if (ta == added) {
		int t = ta, c = ca;

		ta = tb;
		ca = cb;
		tb = t;
		cb = c;
	}
	if (tb != added || ta >= added)
		return FAIL(L, "a JOIN condition must compare the joined table with an earlier one");
	jn->left.tbl = ta;
	jn->left.col = ca;
	jn->right.tbl = tb;
	jn->right.col = cb;
	return 0;
}

/* scaffold: the keys build_cols_hashtable puts (src/engine/executor_select.c:267-291), in its traversal order, into the
 * reference's OWN hashtable - iterating it gives the reference's result-column order without restating its hash */
struct scaffold {
	struct lower *L;
	struct hashtable ht;
	int rc;
};

static void scaffold_put(struct scaffold *s, struct ast_node *n)
{
	struct list_head *pos;
	struct ast_node *kid;

	if (n->node_type == AST_TYPE_SEL_TABLE) {
		struct ast_sel_table_node *tn = (typeof(tn))n;
		struct table *t = database_table_get(s->L->db, tn->table_name);

		for (int c = 0; t && c < t->column_count; c++) {
			char key[FQ_LEN] = {0};

			snprintf(key, sizeof(key) - 1, "%s.%s", tn->table_name, t->columns[c].name);
			if (!hashtable_put(&s->ht, key, strlen(key) + 1, &t->columns[c], sizeof(t->columns[c])))
				s->rc = -1;
		}
	} else if (n->node_type == AST_TYPE_SEL_COUNT) {
		char key[] = "COUNT(*)";
		struct column col = {0};

		col.type = CT_INTEGER;
		col.precision = table_calc_column_precision(col.type);
		col.is_count = true;
		if (!hashtable_put(&s->ht, key, strlen(key) + 1, &col, sizeof(col)))
			s->rc = -1;
	} else if (n->node_type == AST_TYPE_SEL_ALIAS) {
		s->rc = -2; /* aliases: not on the GPU path */
	} else {
		list_for_each(pos, n->node_children_head)
		{
			kid = list_entry(pos, typeof(*kid), head);
			scaffold_put(s, kid);
		}
	}
}

struct out_build {
	struct lower *L;
	struct ast_sel_select_node *sel;
	struct table *result;
	int rc;
};

/* is this scaffold key one of the SELECT expressions (proc_select_clause :1369)?  Fills the plan's output entry. */
static bool selected(struct out_build *ob, const char *key, struct mdbcu_out *out)
{
	struct list_head *pos;
	struct ast_node *n;

	list_for_each(pos, ob->sel->node_children_head)
	{
		char name[FQ_LEN] = {0};

		n = list_entry(pos, typeof(*n), head);
		if (n->node_type == AST_TYPE_SEL_EXPRVAL && ((struct ast_sel_exprval_node*)n)->value_type.is_name) {
			snprintf(name, sizeof(name), "%s", ((struct ast_sel_exprval_node*)n)->name_val);
		} else if (n->node_type == AST_TYPE_SEL_FIELDNAME) {
			struct ast_sel_fieldname_node *f = (typeof(f))n;

			snprintf(name, sizeof(name), "%s.%s", f->table_name, f->col_name);
		} else if (n->node_type == AST_TYPE_SEL_COUNT) {
			strcpy(name, "COUNT(*)");
		} else {
			continue;
		}
		if (strcmp(name, key) != 0)
			continue;
		memset(out, 0, sizeof(*out));
		if (n->node_type == AST_TYPE_SEL_COUNT) {
			out->kind = MDBCU_OUT_COUNT_STAR; /* the reference counts rows whatever the argument (inc_count_cols :1501) */
		} else {
			int t, c;

			if (resolve_name(ob->L, key, &t, &c))
				return false;
			out->kind = MDBCU_OUT_COLUMN;
			out->ref.tbl = t;
			out->ref.col = c;
		}
		return true;
	}
	return false;
}

static void out_each(struct hashtable *ht, const void *key, size_t klen, const void *value, size_t vlen, void *arg)
{
	struct out_build *ob = arg;
	struct mdbcu_out out;
	struct column col = {0};

	(void)ht;
	(void)vlen;
	if (ob->rc || !selected(ob, key, &out))
		return;
	if (ob->L->plan.n_out == MDBCU_MAX_OUT) {
		ob->rc = -1;
		return;
	}
	ob->L->plan.out[ob->L->plan.n_out++] = out;
	strncpy(col.name, key, MIN(sizeof(col.name), klen) - 1);
	col.type = ((const struct column*)value)->type;
	col.precision = ((const struct column*)value)->precision;
	col.is_count = ((const struct column*)value)->is_count;
	if (!table_add_column(ob->result, &col))
		ob->rc = -1;
}

/* (the reference releases its column hashtable the same way: free_hashmap_entries, executor_select.c:66-76) */
static void free_entry(struct hashtable *ht, const void *key, size_t klen, const void *value, size_t vlen, void *arg)
{
	(void)value;
	(void)vlen;
	(void)arg;
	hashtable_free_entry(hashtable_remove(ht, key, klen));
}

int __real_executor_run_select_stmt(struct database *db, struct ast_sel_select_node *select_node, struct query_output *output);

int __wrap_executor_run_select_stmt(struct database *db, struct ast_sel_select_node *select_node, struct query_output *output)
{
	struct lower *L = calloc(1, sizeof(*L));
	struct scaffold sc;
	struct out_build ob;
	struct list_head *pos;
	struct ast_node *n, *from = NULL, *where = NULL, *groupby = NULL;
	mdbcu_ctx *ctx = bridge_ctx();
	mdbcu_table *mirror[MDBCU_MAX_TABLES] = {0};
	mdbcu_result *res = NULL;
	struct table *result = NULL;
	unsigned char *pages = NULL;
	int ret = -MIDORIDB_INTERNAL;

	if (!L || !ctx) {
		snprintf(output->error.message, sizeof(output->error.message), "execution phase: %s\n",
				L ? mdbcu_last_error(NULL) : "out of memory");
		free(L);
		return -MIDORIDB_INTERNAL;
	}
	L->db = db;
	list_for_each(pos, select_node->node_children_head)
	{
		n = list_entry(pos, typeof(*n), head);
		if ((n->node_type == AST_TYPE_SEL_TABLE || n->node_type == AST_TYPE_SEL_JOIN) && !from)
			from = n;
		else if (n->node_type == AST_TYPE_SEL_WHERE)
			where = n;
		else if (n->node_type == AST_TYPE_SEL_GROUPBY)
			groupby = n;
	}
	if (!from || lower_from(L, from))
		goto lower_failed;
	L->plan.n_tables = L->ntab;
	L->plan.n_joins = L->ntab - 1;
	if (where && lower_cond(L, where))
		goto lower_failed;
	if (groupby)
		list_for_each(pos, groupby->node_children_head)
		{
			int t, c;

			n = list_entry(pos, typeof(*n), head);
			if (L->plan.n_group == MDBCU_MAX_GROUP || resolve_node(L, n, &t, &c) != 0) {
				if (!L->err[0])
					snprintf(L->err, sizeof(L->err), "unsupported GROUP BY");
				goto lower_failed;
			}
			L->plan.group[L->plan.n_group].tbl = t;
			L->plan.group[L->plan.n_group].col = c;
			L->plan.n_group++;
		}

	/* result columns in the reference's scaffold order; the result table is a genuine reference table */
	memset(&sc, 0, sizeof(sc));
	sc.L = L;
	if (!hashtable_init(&sc.ht, &hashtable_str_compare, &hashtable_str_hash))
		goto lower_failed;
	scaffold_put(&sc, (struct ast_node*)select_node);
	result = table_init("early_mat_tbl");
	memset(&ob, 0, sizeof(ob));
	ob.L = L;
	ob.sel = select_node;
	ob.result = result;
	if (!sc.rc && result)
		hashtable_foreach(&sc.ht, &out_each, &ob);
	hashtable_foreach(&sc.ht, &free_entry, NULL);
	hashtable_free(&sc.ht);
	if (sc.rc || !result || ob.rc || L->plan.n_out == 0) {
		snprintf(L->err, sizeof(L->err), "unsupported select list");
		goto lower_failed;
	}

	/* device mirrors of the FROM tables: the reference's own datablocks, page by page */
	for (int t = 0; t < L->ntab; t++) {
		int32_t types[MDBCU_MAX_COLUMNS];
		const void **ptrs;
		size_t np = 0, k = 0;

		for (int c = 0; c < L->tab[t]->column_count; c++)
			types[c] = (int32_t)L->tab[t]->columns[c].type;
		if (mdbcu_table_create(ctx, L->name[t], L->tab[t]->column_count, types, &mirror[t]) != MDBCU_OK)
			goto gpu_failed;
		list_for_each(pos, L->tab[t]->datablock_head)
			np++;
		ptrs = calloc(np ? np : 1, sizeof(*ptrs));
		if (!ptrs)
			goto gpu_failed;
		list_for_each(pos, L->tab[t]->datablock_head)
			ptrs[k++] = list_entry(pos, struct datablock, head)->data;
		if (mdbcu_table_append_page_ptrs(mirror[t], ptrs, np) != MDBCU_OK) {
			free(ptrs);
			goto gpu_failed;
		}
		free(ptrs);
		L->plan.tables[t] = mirror[t];
	}
	if (mdbcu_select(ctx, &L->plan, &res) != MDBCU_OK)
		goto gpu_failed;

	/* page images in the reference's row format -> datablocks of the result table */
	{
		size_t n_pages = mdbcu_result_page_count(res), rs = mdbcu_result_row_size(res), rpp;
		uint64_t nrows = mdbcu_result_rows(res);

		if (rs != table_calc_row_size(result)) {
			snprintf(L->err, sizeof(L->err), "row size %zu, the reference computes %zu", rs, table_calc_row_size(result));
			goto lower_failed;
		}
		rpp = (DATABLOCK_PAGE_SIZE - 1) / rs;
		pages = malloc(n_pages * DATABLOCK_PAGE_SIZE);
		if (!pages || mdbcu_result_fetch_pages(res, pages, n_pages) != MDBCU_OK)
			goto gpu_failed;
		for (size_t p = 0; p < n_pages; p++) {
			struct datablock *blk = datablock_alloc(result->datablock_head);

			if (!blk)
				goto gpu_failed;
			memcpy(blk->data, pages + p * DATABLOCK_PAGE_SIZE, DATABLOCK_PAGE_SIZE);
		}
		result->free_dtbkl_offset = nrows ? (size_t)(nrows - (n_pages - 1) * rpp) * rs : 0;
	}
	output->results.table = result;
	result = NULL;
	ret = MIDORIDB_OK;
	goto out;

gpu_failed:
	snprintf(output->error.message, sizeof(output->error.message), "execution phase: %s\n", mdbcu_last_error(ctx));
	goto out;
lower_failed:
	snprintf(output->error.message, sizeof(output->error.message), "execution phase: cannot lower the statement to a GPU plan: %s\n",
			L->err[0] ? L->err : "unsupported shape");
out:
	if (res)
		mdbcu_result_free(res);
	for (int t = 0; t < MDBCU_MAX_TABLES; t++)
		if (mirror[t])
			mdbcu_table_drop(mirror[t]);
	if (result)
		table_destroy(&result);
	free(pages);
	free(L);
	return ret;
}

/* which implementation answers SELECT in this library (the test asserts it is the bridge) */
const char *refh_select_backend(void)
{
	return "libmidoridb_cuda.so via __wrap_executor_run_select_stmt";
}
