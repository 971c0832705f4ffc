/*
 * midoridb.h - public C API of libmidoridb_b200.so: MidoriDB's own entry points, served by the B200 backend.
 *
 * Names, signatures, status codes and the layout of every struct a caller can reach through
 * `struct query_output` are those of the reference (include/engine/query.h:15-69,
 * include/engine/database.h:18-32, include/engine/error.h:11-15, include/primitive/{table,column,row,
 * datablock}.h), so a program written against the reference (README.md:48-77) links against this library
 * unchanged.  What differs is behind query_execute(): statements are parsed by the hand-written front-end
 * (sqlfront.c), row storage keeps the reference's page format on the host, and SELECT runs on the GPU through
 * libmidoridb_cuda.so (include/midoridb_cuda.h).  There is no CPU fallback for SELECT.
 */
#ifndef MIDORIDB_B200_H
#define MIDORIDB_B200_H

#include <pthread.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/engine/error.h:11-15 */
#define MIDORIDB_OK 0
#define MIDORIDB_ERROR 1
#define MIDORIDB_INTERNAL 2
#define MIDORIDB_NOMEM 3
#define MIDORIDB_ROW 4

/* include/datastructure/linkedlist.h */
struct list_head {
	struct list_head *next;
	struct list_head *prev;
};

/* include/primitive/datablock.h:7-13 */
#define DATABLOCK_PAGE_SIZE 4096
struct datablock {
	uint64_t block_id;
	char data[DATABLOCK_PAGE_SIZE];
	struct list_head head;
};

/* include/primitive/column.h:13-49 */
#define TABLE_MAX_COLUMN_NAME 127
enum COLUMN_TYPE { CT_VARCHAR, CT_INTEGER, CT_TINYINT, CT_DOUBLE, CT_DATE, CT_DATETIME };
struct column {
	char name[TABLE_MAX_COLUMN_NAME + 1];
	enum COLUMN_TYPE type;
	int precision;
	bool indexed;
	bool nullable;
	bool unique;
	bool auto_inc;
	bool primary_key;
	bool is_count;
};

/* include/primitive/table.h:16-42 */
#define TABLE_MAX_COLUMNS 128
#define TABLE_MAX_NAME 127
struct table {
	char name[TABLE_MAX_NAME + 1];
	struct column columns[TABLE_MAX_COLUMNS];
	int column_count;
	struct list_head *datablock_head;
	size_t free_dtbkl_offset;
	pthread_mutex_t mutex;
};

/* include/primitive/row.h:15-28 */
struct row_header_flags {
	bool empty;
	bool deleted;
};
struct row {
	struct row_header_flags flags;
	char null_bitmap[TABLE_MAX_COLUMNS / 8];
	__attribute__((aligned(8))) char data[];
};

/* include/engine/database.h:18-21 (`tables` is this library's catalog instead of the reference's hashtable) */
struct database {
	void *tables;
	pthread_mutex_t mutex;
};

/* include/engine/query.h:15-40 */
enum query_output_status { ST_OK_WITH_RESULTS, ST_OK_EXECUTED, ST_ERROR };
struct result_set {
	struct table *table;
	struct datablock *cursor_blk;
	size_t cursor_offset;
};
struct query_output_error {
	char message[1024];
};
struct query_output {
	enum query_output_status status;
	struct result_set results;
	struct query_output_error error;
	size_t n_rows_aff;
};

int database_open(struct database *db);
void database_close(struct database *db);
struct query_output *query_execute(struct database *db, char *query);
int query_cur_step(struct result_set *res);
int64_t query_column_int64(struct result_set *res, int col_idx);
void query_free(struct query_output *output);

/* extensions: typed accessors next to query_column_int64 */
double query_column_double(struct result_set *res, int col_idx);
bool query_column_is_null(struct result_set *res, int col_idx);
/* last physical path the GPU backend took (MDBCU_PATH_*) and its kernel launch count, for tests/benchmarks */
int midoridb_b200_last_path(struct database *db, uint64_t *kernel_launches);
/* test hook: result-column (scaffold) order of the reference for keys put in the given order */
int midoridb_b200_scaffold_order(const char *const *keys, int n, int *position);

#ifdef __cplusplus
}
#endif

#endif /* MIDORIDB_B200_H */
