/*
 * ref_shim.c - TEST INFRASTRUCTURE (oracle/): glue that lets the UNMODIFIED reference
 * (compiled in place from $MIDORIDB_REF/src by oracle/Makefile, output oracle/_ref/) run
 * in this container, where flex/bison are missing.
 *
 * It supplies:
 *   - syntax_parse()  (declared in the reference's include/parser/syntax.h, implemented by the
 *     reference in src/parser/syntax.c:13 on top of flex/bison).  This stand-in feeds the genuine
 *     ast_build_tree -> semantic_analyse -> optimiser_run -> executor_run chain (src/engine/query.c:35-103)
 *     from this repo's hand-written SQL front-end (midoridb_b200/host/sqlfront.c) or from a raw token
 *     script ("\x01" + newline-separated tokens).
 *   - refh_*() helpers for ctypes: bulk table load through the reference's own table_insert_row()
 *     (src/primitive/row.c:26), page export, and a page-walk result dump using the executor idiom
 *     (src/engine/executor_select.c:1096-1106) because query_cur_step() mis-steps on multi-page
 *     results.
 *
 * Nothing here is shipped in the product; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load oracle/_ref/libmidoridb_ref.so.
 */
#define _GNU_SOURCE
#include <parser/syntax.h>
#include <engine/query.h>
#include <engine/database.h>
#include <primitive/table.h>
#include <primitive/row.h>
#include <primitive/column.h>
#include <lib/bit.h>

#include "../midoridb_b200/host/sqlfront.h"

static int offer_token(void *ctx, const char *token)
{
	struct queue *q = ctx;
	return queue_offer(q, (void*)token, strlen(token) + 1) ? 0 : 1;
}

int syntax_parse(char *in, struct queue *out)
{
	char err[256] = {0};

	BUG_ON(!in || !out);

	if (in[0] == '\x01') {
		/* raw token script, one token per line */
		char *copy = strdup(in + 1);
		char *save = NULL;

		if (!copy)
			return 1;
		for (char *tok = strtok_r(copy, "\n", &save); tok; tok = strtok_r(NULL, "\n", &save)) {
			if (!queue_offer(out, tok, strlen(tok) + 1)) {
				free(copy);
				return 1;
			}
		}
		free(copy);
		return 0;
	}

	if (mdb_sql_to_tokens(in, offer_token, out, err, sizeof(err))) {
		/* query_execute copies the message from the head of the queue (query.c:55-58) */
		while (queue_peek(out))
			free(queue_poll(out));
		if (!queue_offer(out, err, strlen(err) + 1))
			fprintf(stderr, "%s\n", err);
		return 1;
	}
	return 0;
}

/* ------------------------------------------------------------------ helpers for ctypes */

struct database* refh_database_new(void)
{
	struct database *db = zalloc(sizeof(*db));
	if (db && database_open(db) != MIDORIDB_OK) {
		free(db);
		db = NULL;
	}
	return db;
}

void refh_database_free(struct database *db)
{
	if (db) {
		database_close(db);
		free(db);
	}
}

/* types use enum COLUMN_TYPE values (include/primitive/column.h:17-25) */
struct table* refh_table_create(struct database *db, const char *name, int ncols, const char *const *col_names,
		const int *col_types)
{
	struct table *table = table_init((char*)name);

	if (!table)
		return NULL;

	for (int i = 0; i < ncols; i++) {
		struct column column = {0};

		strncpy(column.name, col_names[i], sizeof(column.name) - 1);
		column.type = col_types[i];
		column.precision = table_calc_column_precision(column.type);
		column.nullable = true;
		if (!table_add_column(table, &column))
			goto err;
	}

	if (database_table_add(db, table) != MIDORIDB_OK)
		goto err;
	return table;
err:
	table_destroy(&table);
	return NULL;
}

struct table* refh_table_get(struct database *db, const char *name)
{
	if (!database_table_exists(db, (char*)name))
		return NULL;
	return database_table_get(db, (char*)name);
}

/* cells: row-major nrows x ncols raw 8-byte cells; nulls: row-major nrows x ncols bytes or NULL */
long refh_table_append(struct table *table, size_t nrows, const int64_t *cells, const uint8_t *nulls)
{
	size_t row_size = table_calc_row_size(table);
	struct row *row = zalloc(row_size);
	int ncols = table->column_count;

	if (!row)
		return -1;

	for (size_t r = 0; r < nrows; r++) {
		memzero(row, row_size);
		for (int c = 0; c < ncols; c++) {
			if (nulls && nulls[r * ncols + c])
				bit_set(row->null_bitmap, c, sizeof(row->null_bitmap));
			else
				memcpy(row->data + 8 * c, &cells[r * ncols + c], 8);
		}
		if (!table_insert_row(table, row, row_size)) {
			free(row);
			return -1;
		}
	}
	free(row);
	return (long)nrows;
}

size_t refh_table_row_size(struct table *table)
{
	return table_calc_row_size(table);
}

int refh_table_ncols(struct table *table)
{
	return table->column_count;
}

const char* refh_table_colname(struct table *table, int idx)
{
	return table->columns[idx].name;
}

int refh_table_coltype(struct table *table, int idx)
{
	return table->columns[idx].type;
}

size_t refh_table_npages(struct table *table)
{
	struct list_head *pos;
	size_t n = 0;

	list_for_each(pos, table->datablock_head)
		n++;
	return n;
}

/* export pointers to each page's 4096-byte data area, in list order */
size_t refh_table_pages(struct table *table, void **out, size_t cap)
{
	struct list_head *pos;
	size_t n = 0;

	list_for_each(pos, table->datablock_head)
	{
		struct datablock *blk = list_entry(pos, typeof(*blk), head);
		if (n < cap)
			out[n] = blk->data;
		n++;
	}
	return n;
}

/* tombstone the row at (page index, slot) exactly like scan_delete does (executor_delete.c:430) */
int refh_table_delete_slot(struct table *table, size_t page_idx, size_t slot)
{
	struct list_head *pos;
	size_t n = 0, row_size = table_calc_row_size(table);

	list_for_each(pos, table->datablock_head)
	{
		if (n++ == page_idx) {
			struct datablock *blk = list_entry(pos, typeof(*blk), head);
			return table_delete_row(table, blk, slot * row_size) ? 0 : -1;
		}
	}
	return -1;
}

/*
 * walk pages with the executor idiom; writes up to cap rows of raw 8-byte cells (row-major) and
 * null flags.  Returns the total number of live rows.
 */
size_t refh_table_dump(struct table *table, int64_t *cells, uint8_t *nulls, size_t cap)
{
	struct list_head *pos;
	size_t row_size = table_calc_row_size(table);
	size_t n = 0;
	int ncols = table->column_count;

	list_for_each(pos, table->datablock_head)
	{
		struct datablock *blk = list_entry(pos, typeof(*blk), head);
		for (size_t i = 0; i < DATABLOCK_PAGE_SIZE / row_size; i++) {
			struct row *row = (struct row*)&blk->data[row_size * i];

			if (row->flags.empty)
				break;
			if (row->flags.deleted)
				continue;
			if (n < cap) {
				size_t off = 0;
				for (int c = 0; c < ncols; c++) {
					size_t w = table_calc_column_space(&table->columns[c]);
					int64_t v = 0;
					memcpy(&v, row->data + off, MIN(w, sizeof(v)));
					if (cells)
						cells[n * ncols + c] = v;
					if (nulls)
						nulls[n * ncols + c] = bit_test(row->null_bitmap, c,
								sizeof(row->null_bitmap));
					off += w;
				}
			}
			n++;
		}
	}
	return n;
}

/* query_output accessors so ctypes needs no struct layout knowledge */
int refh_output_status(struct query_output *out)
{
	return out->status;
}

const char* refh_output_error(struct query_output *out)
{
	return out->error.message;
}

size_t refh_output_rows_affected(struct query_output *out)
{
	return out->n_rows_aff;
}

struct table* refh_output_table(struct query_output *out)
{
	return out->status == ST_OK_WITH_RESULTS ? out->results.table : NULL;
}

struct result_set* refh_output_results(struct query_output *out)
{
	return &out->results;
}
