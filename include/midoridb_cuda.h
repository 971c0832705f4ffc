/*
 * midoridb_cuda.h - C ABI of libmidoridb_cuda.so, the B200 execution backend for MidoriDB's
 * scan -> WHERE -> INNER JOIN -> GROUP BY/aggregate path.
 *
 * The reference has no plugin/FFI interface; its executor reaches row storage through
 * direct C calls.  Each entry point below therefore replaces one direct call site of the reference
 * (cited per function, paths relative to the reference tree).  A maintainer's binding is shown in
 * `executor_run_select_stmt` (src/engine/executor_select.c:1655) lowers its optimised
 * AST into a `struct mdbcu_plan`, the DML executors notify the mirror, and results come back as page
 * images in the reference's own row format so `query_cur_step` / `query_column_int64`
 * (src/engine/query.c:108,148) keep working unchanged.
 *
 * Conventions: plain C, no C++/torch types, caller-owned input buffers, library-owned handles.
 * Every function returns MDBCU_OK (0) or a negative MDBCU_E* code and never aborts the process;
 * `mdbcu_last_error()` returns the message (the reference's error.message convention,
 * include/engine/query.h:30-32).  There is no CPU fallback: without a CUDA device `mdbcu_init` fails.
 * One in-flight query per context (the reference is single-threaded per database).
 */
#ifndef MIDORIDB_CUDA_H
#define MIDORIDB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDBCU_OK            0
#define MDBCU_EERROR       (-1)  /* -MIDORIDB_ERROR    include/engine/error.h:12 */
#define MDBCU_EINTERNAL    (-2)  /* -MIDORIDB_INTERNAL include/engine/error.h:13 */
#define MDBCU_ENOMEM       (-3)  /* -MIDORIDB_NOMEM    include/engine/error.h:14 */
#define MDBCU_EUNSUPPORTED (-16) /* plan shape outside the GPU hot path */
#define MDBCU_ECUDA        (-17) /* CUDA / NCCL runtime failure */

/* enum COLUMN_TYPE values, include/primitive/column.h:17-25 */
#define MDBCU_CT_VARCHAR  0
#define MDBCU_CT_INTEGER  1
#define MDBCU_CT_TINYINT  2
#define MDBCU_CT_DOUBLE   3
#define MDBCU_CT_DATE     4
#define MDBCU_CT_DATETIME 5

#define MDBCU_PAGE_SIZE      4096 /* DATABLOCK_PAGE_SIZE include/primitive/datablock.h:7 */
#define MDBCU_ROW_HEADER     24   /* offsetof(struct row, data) include/primitive/row.h:22-28 */
#define MDBCU_NULL_BITMAP_OFF 2   /* offsetof(struct row, null_bitmap) */
#define MDBCU_MAX_COLUMNS    128  /* TABLE_MAX_COLUMNS include/primitive/table.h:16 */
#define MDBCU_MAX_TABLES     4
#define MDBCU_MAX_PRED       64
#define MDBCU_MAX_OUT        32
#define MDBCU_MAX_GROUP      2
#define MDBCU_MAX_HAVING     16
#define MDBCU_MAX_ORDER      4

typedef struct mdbcu_ctx mdbcu_ctx;
typedef struct mdbcu_table mdbcu_table;
typedef struct mdbcu_result mdbcu_result;

/* ------------------------------------------------------------------ context */

/* replaces: nothing (new); lifetime follows database_open/database_close, src/engine/database.c:10,45 */
int mdbcu_init(int device, mdbcu_ctx **out);
void mdbcu_shutdown(mdbcu_ctx *ctx);
/* message of the last failure on this context (or of a failed mdbcu_init when ctx == NULL) */
const char *mdbcu_last_error(mdbcu_ctx *ctx);
int mdbcu_device_sync(mdbcu_ctx *ctx);
/* page-locked host memory for page images / result buffers (full PCIe speed for the copies below);
 * plain malloc'd memory is accepted everywhere too, just slower */
void *mdbcu_host_alloc(mdbcu_ctx *ctx, size_t bytes);
void mdbcu_host_free(mdbcu_ctx *ctx, void *p);

/* ------------------------------------------------------------------ device mirror of row storage */

/* replaces: table_init + table_add_column, src/primitive/table.c:58, src/primitive/column.c:119
 * (hooked at executor_run_create_stmt, src/engine/executor_create.c:99) */
int mdbcu_table_create(mdbcu_ctx *ctx, const char *name, int ncols, const int32_t *col_types, mdbcu_table **out);
void mdbcu_table_drop(mdbcu_table *t);

/* replaces: the per-row reads of cpy_cols (src/engine/executor_select.c:340) over pages filled by
 * table_insert_row (src/primitive/row.c:26; hooked at executor_insert.c:230).
 * `pages` = n_pages page data areas (struct datablock.data, 4096 B each) laid out `page_stride`
 * bytes apart (4096 for packed images, sizeof(struct datablock)=4120 for an array of datablocks).
 * Slots are decoded with the executor idiom: stop at flags.empty, skip flags.deleted
 * (executor_select.c:1099-1106).  Host memory; pinned memory gives full PCIe speed. */
int mdbcu_table_append_pages(mdbcu_table *t, const void *pages, size_t n_pages, size_t page_stride);
/* same, for the reference's individually malloc'd datablocks (src/primitive/datablock.c:17):
 * page_ptrs[i] points at datablock i's data area */
int mdbcu_table_append_page_ptrs(mdbcu_table *t, const void *const *page_ptrs, size_t n_pages);
/* replace the mirror of pages [first_page, first_page+n) after in-place UPDATE/DELETE
 * (executor_update.c:477, executor_delete.c:430) */
int mdbcu_table_reload_pages(mdbcu_table *t, size_t first_page, const void *const *page_ptrs, size_t n_pages);
/* tombstone rows by (page, slot): replaces table_delete_row, src/primitive/row.c:126 */
int mdbcu_table_tombstone(mdbcu_table *t, const uint64_t *page_idx, const uint32_t *slot_idx, size_t n);

/* bulk columnar load (bench/e2e helper, no reference counterpart): col_data[c] = n_rows 8-byte cells,
 * col_nulls[c] = n_rows byte flags or NULL; the arrays may be host or device memory (the library's own distributed GROUP BY
 * appends device-resident partial rows) */
int mdbcu_table_append_columns(mdbcu_table *t, size_t n_rows, const void *const *col_data,
		const uint8_t *const *col_nulls);

/* device-side synthetic load for scale configs whose page images would not fit host memory
 * (SURVEY.md 7.2 "Host capacity").  Row r (global index row_offset + r) of column c gets a
 * counter-based value, identical for any sharding of the same (seed, global index). */
#define MDBCU_GEN_UNIFORM_INT   0 /* uniform int64 in [lo, hi]                       */
#define MDBCU_GEN_UNIFORM_DBL   1 /* uniform double in [0,1)                         */
#define MDBCU_GEN_PERMUTATION   2 /* bijection of [0, hi-lo] + lo (needs hi-lo+1 = 2^k) */
#define MDBCU_GEN_ZIPF          3 /* Zipf(param) over [lo, hi]                       */
#define MDBCU_GEN_SEQUENCE      4 /* lo + global index                               */
struct mdbcu_gen_spec {
	int32_t kind;
	int32_t null_permille; /* fraction of NULL cells, in 1/1000 */
	int64_t lo, hi;
	double param;
	uint64_t seed;
};
int mdbcu_table_generate(mdbcu_table *t, uint64_t n_rows, uint64_t row_offset, const struct mdbcu_gen_spec *specs);

uint64_t mdbcu_table_slots(const mdbcu_table *t); /* device row slots (live + tombstoned + unused tail slots) */
uint64_t mdbcu_table_live_rows(mdbcu_table *t);   /* counts on device */
/* copy one mirrored column back (test hook): cells = 8-byte values, valid = byte flags (1 = live and not NULL) */
int mdbcu_table_read_column(mdbcu_table *t, int col, uint64_t first, uint64_t n, void *cells, uint8_t *valid);

/* zero-copy view of one mirrored column for a harness that has its own CUDA code (bench.py verifies results with an
 * independent torch histogram): *cells = n_slots 8-byte cells in device memory, valid until the table changes or is
 * dropped.  Rows are in slot order; NULL / tombstoned slots hold unspecified values (see mdbcu_table_read_column). */
int mdbcu_table_column_device_ptr(mdbcu_table *t, int col, const void **cells, uint64_t *n_slots);

/* ------------------------------------------------------------------ query plan */

/* predicate program, postfix, mirroring the reference's RPN token stream (midorisql.y:250-282) as
 * evaluated by eval_row_cond / eval_cmp (src/engine/executor_select.c:1027,865).  NULL operand =>
 * comparison is false (executor_select.c:557-579,629-631). */
#define MDBCU_P_COL     1 /* push column (tbl, col)                    */
#define MDBCU_P_INT     2 /* push int64 literal ival                   */
#define MDBCU_P_DBL     3 /* push double literal dval                  */
#define MDBCU_P_NULL    4 /* push NULL literal                         */
#define MDBCU_P_CMP     5 /* pop b, pop a, push a <cmp> b; arg = 1 '<' 2 '>' 3 '<>' 4 '=' 5 '<=' 6 '>=' (midorisql.l:122-128) */
#define MDBCU_P_AND     6
#define MDBCU_P_OR      7
#define MDBCU_P_XOR     8
#define MDBCU_P_ISNULL  9 /* pop operand, push is-null                 */
#define MDBCU_P_ISNOTNULL 10
#define MDBCU_P_IN      11 /* arg = n: pop n values then the probe; true if ANY equal (SQL; see DESIGN.md on D4) */
#define MDBCU_P_NOTIN   12 /* arg = n: true if ALL differ               */
#define MDBCU_P_BOOL    13 /* push boolean literal ival                 */
#define MDBCU_P_OUT     14 /* HAVING programs only: push result column `col` (an index into plan.out) of the candidate row */

struct mdbcu_pred_op {
	int32_t op;
	int32_t arg;
	int32_t tbl, col;
	int64_t ival;
	double dval;
};

struct mdbcu_colref {
	int32_t tbl; /* index into plan.tables */
	int32_t col;
};

/* INNER equi-join step (JOIN 1 + ONEXPR FIELDNAME = FIELDNAME, midorisql.y:231,247;
 * _join_nested_loop_tbl2tbl / _tbl2mat, executor_select.c:1076,1151).  tables[0..i+1] joined so far. */
struct mdbcu_join {
	struct mdbcu_colref left;  /* a column of tables[0..i]  */
	struct mdbcu_colref right; /* a column of tables[i+1]   */
	int32_t cross;             /* 1 = no condition (comma list / ON 1=1, optimiser_select.c:395) */
	int32_t _pad;
};

/* output column kinds */
#define MDBCU_OUT_COLUMN     0 /* plain column; under GROUP BY: the value of the group's first row
                                  (the "survivor" row_1 of proc_groupby_clause, executor_select.c:1546) */
#define MDBCU_OUT_COUNT_STAR 1 /* COUNT(*)  init_count_cols / inc_count_cols, executor_select.c:324,1501 */
#define MDBCU_OUT_COUNT_COL  2 /* COUNT(x): non-NULL x (SQL)                                              */
#define MDBCU_OUT_SUM        3 /* extension: INT -> int64 (wrapping), DOUBLE -> double                    */
#define MDBCU_OUT_MIN        4
#define MDBCU_OUT_MAX        5
#define MDBCU_OUT_AVG        6 /* extension: always double                                                */

struct mdbcu_out {
	int32_t kind;
	struct mdbcu_colref ref;
};

#define MDBCU_PLAN_DISTRIBUTED 1u /* tables are this rank's shards (needs mdbcu_comm_init).  Distributed plan shapes: join + GROUP BY join
                                   * key + COUNT(*) (join sides exchanged by key, every rank returns the keys it owns); filter + aggregate
                                   * scans, small-dimension star joins and GROUP BY / aggregates over one sharded table (partials merged,
                                   * rank 0 returns the rows).  Anything else: MDBCU_EUNSUPPORTED */
#define MDBCU_PLAN_NO_FASTPATH 2u /* test hook: force the general operators                                */

/* what executor_run_select_stmt (executor_select.c:1655) receives as an optimised AST, flattened.
 * Limits of the device operators - such a plan returns MDBCU_EUNSUPPORTED with the reason in mdbcu_last_error(), there is
 * no CPU fallback: VARCHAR columns in predicates, join keys, GROUP BY or aggregates (not mirrored on the device); join keys
 * of different kinds (INT vs DOUBLE); a two-column GROUP BY whose columns are not both integers within the int32 range and
 * free of NULLs (one-column keys of any kind, NULLs included, are fine); plain non-key columns under GROUP BY when the row
 * ids of the joined tables do not pack into 64 bits; more than 2^32-1 joined rows in the general operators; WHERE programs
 * nesting deeper than 24 operands. */
struct mdbcu_plan {
	int32_t n_tables;
	mdbcu_table *tables[MDBCU_MAX_TABLES];
	int32_t n_joins; /* == n_tables - 1 */
	struct mdbcu_join joins[MDBCU_MAX_TABLES - 1];
	int32_t n_pred; /* 0 = no WHERE */
	struct mdbcu_pred_op pred[MDBCU_MAX_PRED];
	int32_t n_group; /* 0 = no GROUP BY */
	struct mdbcu_colref group[MDBCU_MAX_GROUP];
	int32_t n_out;
	struct mdbcu_out out[MDBCU_MAX_OUT];
	uint32_t flags;
	/* Tail operators on the result rows.  The reference's grammar accepts them (DISTINCT midorisql.y:203, HAVING :180,
	 * ORDER BY :183-191, LIMIT :193-196) and its semantic phase validates them (semantic_select.c:1895,2004), but its
	 * executor never runs them (`TODO process distinct`, executor_select.c:1723): there is no reference
	 * behaviour to match, the semantics are SQL's as sqlite3 implements them (the oracle is anchored on it).  Applied in
	 * this order: HAVING, DISTINCT, ORDER BY, LIMIT.  All-zero = none.  In MDBCU_PLAN_DISTRIBUTED plans only over ONE sharded
	 * table (rank 0 returns the whole result and applies them); a distributed join's result is spread over the ranks. */
	int32_t distinct;                               /* SELECT DISTINCT: one row per distinct combination of ALL output columns (NULL = NULL) */
	int32_t n_having;                               /* HAVING: postfix program over the result row, columns pushed with MDBCU_P_OUT */
	struct mdbcu_pred_op having[MDBCU_MAX_HAVING];
	int32_t n_order;                                /* ORDER BY: result columns, most significant first; NULLs first when ascending, */
	struct mdbcu_order {                            /* last when descending; ties keep no particular order                              */
		int32_t out_col;                        /* index into plan.out */
		int32_t desc;
	} order[MDBCU_MAX_ORDER];
	int32_t has_limit;                              /* LIMIT [offset,] count */
	int32_t _pad2;
	int64_t limit, offset;
};

/* replaces: executor_run_select_stmt's data path (proc_from_clause :1345, proc_where_clause :1435,
 * proc_groupby_clause :1526, proc_select_clause :1369, handle_countonly_case :1590) */
int mdbcu_select(mdbcu_ctx *ctx, const struct mdbcu_plan *plan, mdbcu_result **out);

/* ------------------------------------------------------------------ results */

uint64_t mdbcu_result_rows(const mdbcu_result *r);
int mdbcu_result_cols(const mdbcu_result *r);
int mdbcu_result_col_type(const mdbcu_result *r, int col); /* MDBCU_CT_INTEGER or MDBCU_CT_DOUBLE */
/* columnar fetch into caller buffers: cells[c] = rows x 8 bytes, nulls[c] = rows bytes (may be NULL) */
int mdbcu_result_fetch_columns(mdbcu_result *r, void *const *cells, uint8_t *const *nulls);
/* zero-copy view of one result column (device memory, rows in device order, valid until mdbcu_result_free):
 * *cells = rows x 8 bytes, *nulls = rows byte flags or NULL when the column has no NULL cell */
int mdbcu_result_column_device_ptr(mdbcu_result *r, int col, const void **cells, const uint8_t **nulls);
/* replaces: table_insert_row on the result + table_vacuum (src/primitive/row.c:26, vacuum.c:12):
 * result rows laid out as page images in the reference row format (24-byte header, null bitmap,
 * packed 8-byte cells, floor(4095/row_size) rows per page, tail slots flags.empty=1; an empty
 * result is ONE page of empty slots).  `pages` = n_pages*4096 bytes, caller-owned. */
size_t mdbcu_result_page_count(const mdbcu_result *r);
size_t mdbcu_result_row_size(const mdbcu_result *r);
int mdbcu_result_fetch_pages(mdbcu_result *r, void *pages, size_t n_pages);
void mdbcu_result_free(mdbcu_result *r);

/* ------------------------------------------------------------------ stats (opt-in metrics) */

struct mdbcu_stats {
	double total_ms;     /* device time of the last mdbcu_select, CUDA events on the context's stream */
	double phase_ms[8];  /* 0 scan/filter 1 partition 2 build/count 3 probe/emit 4 aggregate 5 materialise 6 exchange 7 other */
	uint64_t kernel_launches; /* kernels launched by the last mdbcu_select */
	uint64_t total_kernel_launches; /* since mdbcu_init */
	uint64_t input_rows; /* sum of slots of the scanned tables */
	uint64_t result_rows;
	uint64_t algorithmic_bytes; /* compulsory input + output bytes of the last select */
	int32_t path; /* which physical path ran: see MDBCU_PATH_* */
	int32_t _pad;
	double dominant_ms; /* device time of the dominant kernel(s) of the path (roofline numerator's denominator) */
	uint64_t dominant_bytes; /* algorithmic bytes attributed to the dominant kernel(s) */
	uint64_t exchange_bytes; /* bytes this rank sent to OTHER ranks over NVLink in the last select */
};
#define MDBCU_PATH_GENERAL        0
#define MDBCU_PATH_SCAN_AGG       1 /* fused filter + aggregate scan (no join, no GROUP BY) */
#define MDBCU_PATH_RADIX_JOINCOUNT 2 /* radix-partitioned join + GROUP BY join key, COUNT(*) */
#define MDBCU_PATH_DIRECT_STAR    3 /* small build side, direct-addressed probe + grouped MIN/MAX/SUM/COUNT */
#define MDBCU_PATH_FUSED_MULTIWAY 4 /* scan of tables[0] with direct row tables for every joined table, WHERE and grouped
                                      aggregates in one kernel (no tuple arrays); MDBCU_FUSED_MULTIWAY=0 disables it */
#define MDBCU_PATH_DIRECT_COUNT   5 /* join + GROUP BY join key + COUNT(*) with one 32-bit counter per key value: any key
                                      multiplicity, skew or order (what the radix path hands back); distributed plans sum
                                      the counters over the ranks */
int mdbcu_get_stats(mdbcu_ctx *ctx, struct mdbcu_stats *out);

/* CUDA events on the context's own stream (the stream every kernel of this library is launched on), so a
 * harness can time a region of calls on the device: record slot a, run, record slot b, read the elapsed time.
 * MDBCU_EVENT_SLOTS slots. */
#define MDBCU_EVENT_SLOTS 8
int mdbcu_event_record(mdbcu_ctx *ctx, int slot);
int mdbcu_event_elapsed_ms(mdbcu_ctx *ctx, int slot_start, int slot_stop, double *ms);

/* ------------------------------------------------------------------ multi-GPU (one process per GPU) */

/* 128-byte NCCL unique id, created on rank 0 and distributed by the host program (torch.distributed,
 * MPI, a file...).  The join exchange is the only collective on the path. */
int mdbcu_comm_unique_id(mdbcu_ctx *ctx, void *id128);
int mdbcu_comm_init(mdbcu_ctx *ctx, int rank, int world, const void *id128);
int mdbcu_comm_world(mdbcu_ctx *ctx, int *rank, int *world);
/* Several contexts of ONE process as the ranks 0..world-1 of a communicator (instead of mdbcu_comm_init): one host thread
 * per context drives its rank, collectives meet in a host rendezvous, exchange arenas are plain peer pointers.  The
 * contexts may sit on different GPUs (peer access is enabled) or all on the same GPU - the loop-back mode that lets a
 * single-GPU box run the exchange kernels of a distributed plan.  Call once, from one thread, before the
 * contexts are used; every collective call (mdbcu_table_sync_stats, a MDBCU_PLAN_DISTRIBUTED select) must then be made
 * by all ranks concurrently, each from its own thread. */
int mdbcu_comm_init_local(mdbcu_ctx *const *ctxs, int world);
/* COLLECTIVE (every rank calls it for its shard of the same table): exchanges the zone-map statistics
 * (min / max per integer column) so every rank partitions by the same global key range.  Call after loading
 * or changing a sharded table, before a MDBCU_PLAN_DISTRIBUTED select. */
int mdbcu_table_sync_stats(mdbcu_table *t);

/* How a distributed radix join (join + GROUP BY join key + COUNT(*)) lays out its exchange; pure host arithmetic, usable
 * without a device (tests).  Rank r owns partitions [part_first[r], part_first[r + 1]).  Every rank's arena holds, per
 * query half and join side, the streams pass 1 produced from the rank's own shard for ALL partitions; the owner of a
 * partition reads them from every rank's arena. */
struct mdbcu_dist_layout {
	uint32_t part_first[9];        /* world + 1 entries */
	uint32_t stream_cap, tail_cap; /* entries per partition in the main / tail stream (identical on every rank) */
	uint32_t _pad;
	uint64_t region_main_off, region_tail_off, region_cursor_off, region_bytes; /* one join side inside an arena half */
	uint64_t arena_half_bytes;     /* one of the two halves alternate queries use */
};
int mdbcu_dist_describe(int nparts, int world, uint64_t global_rows, int sms, struct mdbcu_dist_layout *out);
int mdbcu_dist_owner(uint32_t partition, int nparts, int world); /* rank that owns `partition`, -1 if out of range */

const char *mdbcu_version(void);

#ifdef __cplusplus
}
#endif

#endif /* MIDORIDB_CUDA_H */
