/*
 * sqlfront.h - hand-written SQL front-end for the hot-path subset.
 *
 * Replaces the reference's flex/bison parser (src/parser/midorisql.l,
 * src/parser/midorisql.y, src/parser/syntax.c:13) where flex/bison are not
 * available.  It produces exactly the postfix text-token protocol that the
 * reference grammar's emit() calls produce (midorisql.y:517), so the token
 * stream can be fed either to the reference's own ast_build_tree() (via the
 * syntax_parse() stand-in in oracle/ref_shim.c) or to this repo's planner
 * (midoridb_b200/host/planner.cpp).
 *
 * Extensions over the reference grammar (the reference cannot parse these):
 *   SUM(x) / MIN(x) / MAX(x) / AVG(x)  -> tokens SUMFIELD / MINFIELD / MAXFIELD / AVGFIELD
 *   x BETWEEN a AND b                  -> lowered to  x a CMP 6  x b CMP 5  AND
 */
#ifndef MIDORIDB_B200_SQLFRONT_H
#define MIDORIDB_B200_SQLFRONT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* called once per token, in emission order; return non-zero to abort */
typedef int (*mdb_sql_emit_fn)(void *ctx, const char *token);

/*
 * mdb_sql_to_tokens - parse ONE statement (must end in ';', midorisql.y:148)
 * Returns 0 on success, non-zero on syntax error (err filled, NUL-terminated).
 */
int mdb_sql_to_tokens(const char *sql, mdb_sql_emit_fn emit, void *ctx, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif

#endif
