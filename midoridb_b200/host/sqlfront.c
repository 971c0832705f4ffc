/*
 * sqlfront.c - hand-written SQL -> postfix token stream (see sqlfront.h).
 *
 * Token protocol restated from the reference grammar's emit() actions:
 *   operands   midorisql.y:250-256   NAME / FIELDNAME / STRING / NUMBER / FLOAT / BOOL / NULL
 *   operators  midorisql.y:259-282   ADD SUB MUL DIV MOD NEG AND OR XOR "CMP k" ISNULL ISNOTNULL "ISIN n" "ISNOTIN n"
 *   compare k  midorisql.l:122-128   1 '<'  2 '>'  3 '<>' '!='  4 '='  5 '<='  6 '>='
 *   select     midorisql.y:157-160,164,168,208,221,224-225,231-233,247,286-287
 *   delete     midorisql.y:315     insert midorisql.y:350-364     update midorisql.y:392-407
 *   create     midorisql.y:451-483 (type codes 50000 INT, 60000 TINYINT, 80000 DOUBLE, ...)
 * Emission order equals bison's reduction order, i.e. plain postfix.
 * Precedence table: midorisql.y:49-63.
 */
#include "sqlfront.h"

#include <ctype.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

enum tk {
	T_EOF, T_NAME, T_INT, T_FLOAT, T_STRING, T_BOOL, T_CMP, T_PUNCT, T_KW, T_ANDOP, T_OROP
};

struct tok {
	enum tk kind;
	char *text;      /* NAME / STRING / KW (upper-cased) / PUNCT char */
	long long ival;  /* INT / BOOL / CMP subtok */
	double fval;
};

struct parser {
	struct tok *toks;
	size_t ntoks, pos;
	char **out;
	size_t nout, capout;
	char err[256];
	int failed;
};

static const char *const KEYWORDS[] = {
	"AND", "AS", "ASC", "AUTO_INCREMENT", "BETWEEN", "BY", "CASE", "CREATE", "CURRENT_DATE",
	"CURRENT_TIMESTAMP", "DATE", "DATETIME", "DELETE", "DESC", "DISTINCT", "DOUBLE", "ELSE", "END",
	"EXISTS", "FROM", "GROUP", "HAVING", "IF", "IN", "INDEX", "INNER", "INSERT", "INT", "INT4",
	"INTEGER", "INTO", "IS", "JOIN", "KEY", "LEFT", "LIKE", "LIMIT", "MOD", "NOT", "NULL", "ON", "OR",
	"ORDER", "OUTER", "PRIMARY", "RIGHT", "SELECT", "SET", "TABLE", "THEN", "TINYINT", "UNIQUE",
	"UPDATE", "VALUES", "VALUE", "VARCHAR", "VARCHARACTER", "WHEN", "WHERE", "XOR", NULL
};

static void fail(struct parser *p, const char *fmt, ...)
{
	va_list ap;

	if (p->failed)
		return;
	p->failed = 1;
	va_start(ap, fmt);
	vsnprintf(p->err, sizeof(p->err), fmt, ap);
	va_end(ap);
}

static void emit(struct parser *p, const char *fmt, ...)
{
	char buf[256];
	va_list ap;

	if (p->failed)
		return;
	memset(buf, 0, sizeof(buf));
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);

	if (p->nout == p->capout) {
		size_t ncap = p->capout ? p->capout * 2 : 64;
		char **n = realloc(p->out, ncap * sizeof(*n));
		if (!n) {
			fail(p, "out of memory");
			return;
		}
		p->out = n;
		p->capout = ncap;
	}
	p->out[p->nout] = strdup(buf);
	if (!p->out[p->nout]) {
		fail(p, "out of memory");
		return;
	}
	p->nout++;
}

/* re-emit a copy of already emitted tokens [from, to) - used by the BETWEEN lowering */
static void emit_copy(struct parser *p, size_t from, size_t to)
{
	for (size_t i = from; i < to && !p->failed; i++)
		emit(p, "%s", p->out[i]);
}

/* ---------------------------------------------------------------- lexer */

static int is_keyword(const char *up)
{
	for (int i = 0; KEYWORDS[i]; i++)
		if (strcmp(KEYWORDS[i], up) == 0)
			return 1;
	return 0;
}

static int push_tok(struct parser *p, size_t *cap, struct tok t)
{
	if (p->ntoks == *cap) {
		size_t ncap = *cap ? *cap * 2 : 64;
		struct tok *n = realloc(p->toks, ncap * sizeof(*n));
		if (!n)
			return -1;
		p->toks = n;
		*cap = ncap;
	}
	p->toks[p->ntoks++] = t;
	return 0;
}

static char *dup_range(const char *s, size_t n)
{
	char *r = malloc(n + 1);
	if (r) {
		memcpy(r, s, n);
		r[n] = 0;
	}
	return r;
}

static int prev_is_operand(struct parser *p)
{
	if (p->ntoks == 0)
		return 0;
	struct tok *t = &p->toks[p->ntoks - 1];
	if (t->kind == T_NAME || t->kind == T_INT || t->kind == T_FLOAT || t->kind == T_STRING || t->kind == T_BOOL)
		return 1;
	if (t->kind == T_PUNCT && t->text[0] == ')')
		return 1;
	if (t->kind == T_KW && strcmp(t->text, "NULL") == 0)
		return 1;
	return 0;
}

static int lex(struct parser *p, const char *s)
{
	size_t cap = 0;
	size_t i = 0, n = strlen(s);

	while (i < n) {
		struct tok t = {0};
		char c = s[i];

		if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
			i++;
			continue;
		}
		/* comments: '#...', '-- ...', C-style (midorisql.l:158-166) */
		if (c == '#' || (c == '-' && i + 2 < n && s[i + 1] == '-' && (s[i + 2] == ' ' || s[i + 2] == '\t'))) {
			while (i < n && s[i] != '\n')
				i++;
			continue;
		}
		if (c == '/' && i + 1 < n && s[i + 1] == '*') {
			i += 2;
			while (i + 1 < n && !(s[i] == '*' && s[i + 1] == '/'))
				i++;
			if (i + 1 >= n) {
				fail(p, "unclosed comment");
				return -1;
			}
			i += 2;
			continue;
		}
		/* numbers; a leading '-' belongs to the literal (midorisql.l:85-91) unless it follows an operand */
		if (isdigit((unsigned char)c) || (c == '.' && i + 1 < n && isdigit((unsigned char)s[i + 1])) ||
				(c == '-' && !prev_is_operand(p) && i + 1 < n &&
				(isdigit((unsigned char)s[i + 1]) ||
				(s[i + 1] == '.' && i + 2 < n && isdigit((unsigned char)s[i + 2]))))) {
			size_t j = i + (c == '-');
			int isfloat = 0;
			while (j < n && isdigit((unsigned char)s[j]))
				j++;
			if (j < n && s[j] == '.') {
				isfloat = 1;
				j++;
				while (j < n && isdigit((unsigned char)s[j]))
					j++;
			}
			if (j < n && (s[j] == 'E' || s[j] == 'e')) {
				size_t k = j + 1;
				if (k < n && (s[k] == '+' || s[k] == '-'))
					k++;
				if (k < n && isdigit((unsigned char)s[k])) {
					isfloat = 1;
					while (k < n && isdigit((unsigned char)s[k]))
						k++;
					j = k;
				}
			}
			t.text = dup_range(s + i, j - i);
			if (!t.text)
				goto nomem;
			if (isfloat) {
				t.kind = T_FLOAT;
				t.fval = atof(t.text);
			} else {
				t.kind = T_INT;
				t.ival = strtoll(t.text, NULL, 10);
			}
			i = j;
			if (push_tok(p, &cap, t))
				goto nomem;
			continue;
		}
		if (c == '\'' || c == '"') {
			size_t j = i + 1;
			while (j < n && s[j] != '\n') {
				if (s[j] == '\\' && j + 1 < n) {
					j += 2;
					continue;
				}
				if (s[j] == c) {
					if (j + 1 < n && s[j + 1] == c) {
						j += 2;
						continue;
					}
					break;
				}
				j++;
			}
			if (j >= n || s[j] != c) {
				fail(p, "Unterminated string");
				return -1;
			}
			t.kind = T_STRING;
			t.text = dup_range(s + i, j + 1 - i); /* quotes kept, as yytext is (midorisql.l:101) */
			if (!t.text)
				goto nomem;
			i = j + 1;
			if (push_tok(p, &cap, t))
				goto nomem;
			continue;
		}
		if (c == '`') {
			size_t j = i + 1;
			while (j < n && s[j] != '`' && s[j] != '\n')
				j++;
			if (j >= n || s[j] != '`') {
				fail(p, "unterminated quoted name");
				return -1;
			}
			t.kind = T_NAME;
			t.text = dup_range(s + i + 1, j - i - 1);
			if (!t.text)
				goto nomem;
			i = j + 1;
			if (push_tok(p, &cap, t))
				goto nomem;
			continue;
		}
		if (isalpha((unsigned char)c)) {
			size_t j = i;
			char up[64];
			while (j < n && (isalnum((unsigned char)s[j]) || s[j] == '_'))
				j++;
			t.text = dup_range(s + i, j - i);
			if (!t.text)
				goto nomem;
			if (j - i < sizeof(up)) {
				for (size_t k = 0; k < j - i; k++)
					up[k] = (char)toupper((unsigned char)s[i + k]);
				up[j - i] = 0;
				if (strcmp(up, "TRUE") == 0 || strcmp(up, "FALSE") == 0 || strcmp(up, "UNKNOWN") == 0) {
					t.kind = T_BOOL;
					t.ival = up[0] == 'T' ? 1 : (up[0] == 'F' ? 0 : -1);
				} else if (is_keyword(up)) {
					t.kind = T_KW;
					strcpy(t.text, up);
				} else {
					t.kind = T_NAME;
				}
			} else {
				t.kind = T_NAME;
			}
			/* COUNT/SUM/MIN/MAX/AVG are functions only when directly followed by '(' (midorisql.l:138-142) */
			i = j;
			if (push_tok(p, &cap, t))
				goto nomem;
			continue;
		}
		/* operators */
		t.text = NULL;
		if (c == '&' && i + 1 < n && s[i + 1] == '&') {
			t.kind = T_ANDOP;
			i += 2;
		} else if (c == '|' && i + 1 < n && s[i + 1] == '|') {
			t.kind = T_OROP;
			i += 2;
		} else if (c == '=') {
			t.kind = T_CMP; t.ival = 4; i++;
		} else if (c == '>' && i + 1 < n && s[i + 1] == '=') {
			t.kind = T_CMP; t.ival = 6; i += 2;
		} else if (c == '>') {
			t.kind = T_CMP; t.ival = 2; i++;
		} else if (c == '<' && i + 1 < n && s[i + 1] == '=') {
			t.kind = T_CMP; t.ival = 5; i += 2;
		} else if (c == '<' && i + 1 < n && s[i + 1] == '>') {
			t.kind = T_CMP; t.ival = 3; i += 2;
		} else if (c == '<') {
			t.kind = T_CMP; t.ival = 1; i++;
		} else if (c == '!' && i + 1 < n && s[i + 1] == '=') {
			t.kind = T_CMP; t.ival = 3; i += 2;
		} else if (strchr("-+&~|^/%*(),.;!", c)) {
			t.kind = T_PUNCT;
			t.text = dup_range(s + i, 1);
			if (!t.text)
				goto nomem;
			i++;
		} else {
			fail(p, "mystery character '%c'", c);
			return -1;
		}
		if (push_tok(p, &cap, t))
			goto nomem;
	}
	{
		struct tok t = {0};
		t.kind = T_EOF;
		if (push_tok(p, &cap, t))
			goto nomem;
	}
	return 0;
nomem:
	fail(p, "out of memory");
	return -1;
}

/* --------------------------------------------------------------- parser */

static struct tok *peek(struct parser *p)
{
	return &p->toks[p->pos < p->ntoks ? p->pos : p->ntoks - 1];
}

static struct tok *peek2(struct parser *p)
{
	return &p->toks[p->pos + 1 < p->ntoks ? p->pos + 1 : p->ntoks - 1];
}

static int is_kw(struct tok *t, const char *kw)
{
	return t->kind == T_KW && strcmp(t->text, kw) == 0;
}

static int is_punct(struct tok *t, char c)
{
	return t->kind == T_PUNCT && t->text[0] == c;
}

static int accept_kw(struct parser *p, const char *kw)
{
	if (is_kw(peek(p), kw)) {
		p->pos++;
		return 1;
	}
	return 0;
}

static int accept_punct(struct parser *p, char c)
{
	if (is_punct(peek(p), c)) {
		p->pos++;
		return 1;
	}
	return 0;
}

static void expect_kw(struct parser *p, const char *kw)
{
	if (!accept_kw(p, kw))
		fail(p, "syntax error: expected %s", kw);
}

static void expect_punct(struct parser *p, char c)
{
	if (!accept_punct(p, c))
		fail(p, "syntax error: expected '%c'", c);
}

static char *expect_name(struct parser *p)
{
	struct tok *t = peek(p);
	if (t->kind != T_NAME) {
		fail(p, "syntax error: expected a name");
		return "";
	}
	p->pos++;
	return t->text;
}

/* binding powers, midorisql.y:49-63 */
enum {
	P_OR = 1, P_XOR, P_AND, P_INISLIKE, P_NOT, P_BETWEEN, P_CMP, P_BITOR, P_BITAND, P_SHIFT, P_ADD, P_MUL,
	P_POW, P_UMINUS
};

static void parse_expr(struct parser *p, int min_prec);

static int func_token(struct tok *t, struct tok *next, const char **out)
{
	static const char *const fn[][2] = {
		{"SUM", "SUMFIELD"}, {"MIN", "MINFIELD"}, {"MAX", "MAXFIELD"}, {"AVG", "AVGFIELD"}, {NULL, NULL}
	};
	if (t->kind != T_NAME || !is_punct(next, '('))
		return 0;
	for (int i = 0; fn[i][0]; i++) {
		if (strcasecmp(t->text, fn[i][0]) == 0) {
			*out = fn[i][1];
			return 1;
		}
	}
	return 0;
}

static void parse_primary(struct parser *p)
{
	struct tok *t = peek(p);
	const char *fn;

	if (p->failed)
		return;

	if (t->kind == T_NAME && strcasecmp(t->text, "COUNT") == 0 && is_punct(peek2(p), '(')) {
		p->pos += 2;
		if (accept_punct(p, '*')) {
			expect_punct(p, ')');
			emit(p, "COUNTALL");
		} else {
			parse_expr(p, P_OR);
			expect_punct(p, ')');
			emit(p, "COUNTFIELD");
		}
	} else if (func_token(t, peek2(p), &fn)) {
		p->pos += 2;
		parse_expr(p, P_OR);
		expect_punct(p, ')');
		emit(p, "%s", fn);
	} else if (t->kind == T_NAME) {
		p->pos++;
		if (is_punct(peek(p), '.') && peek2(p)->kind == T_NAME) {
			emit(p, "FIELDNAME %s.%s", t->text, peek2(p)->text);
			p->pos += 2;
		} else {
			emit(p, "NAME %s", t->text);
		}
	} else if (t->kind == T_STRING) {
		p->pos++;
		emit(p, "STRING %s", t->text);
	} else if (t->kind == T_INT) {
		p->pos++;
		emit(p, "NUMBER %lld", t->ival);
	} else if (t->kind == T_FLOAT) {
		p->pos++;
		emit(p, "FLOAT %g", t->fval);
	} else if (t->kind == T_BOOL) {
		p->pos++;
		emit(p, "BOOL %d", (int)t->ival);
	} else if (is_kw(t, "NULL")) {
		p->pos++;
		emit(p, "NULL");
	} else if (is_kw(t, "CURRENT_TIMESTAMP") || is_kw(t, "CURRENT_DATE")) {
		p->pos++;
		emit(p, "NOW");
	} else if (is_punct(t, '(')) {
		p->pos++;
		parse_expr(p, P_OR);
		expect_punct(p, ')');
	} else if (is_punct(t, '-')) {
		p->pos++;
		parse_expr(p, P_UMINUS);
		emit(p, "NEG");
	} else {
		fail(p, "syntax error near token %zu", p->pos);
	}
}

static int parse_val_list(struct parser *p)
{
	int n = 0;
	expect_punct(p, '(');
	do {
		parse_expr(p, P_OR);
		n++;
	} while (!p->failed && accept_punct(p, ','));
	expect_punct(p, ')');
	return n;
}

static void parse_expr(struct parser *p, int min_prec)
{
	size_t lhs_start = p->nout;

	parse_primary(p);

	while (!p->failed) {
		struct tok *t = peek(p);
		int prec;
		const char *op = NULL;

		if (is_kw(t, "OR") || t->kind == T_OROP) {
			prec = P_OR; op = "OR";
		} else if (is_kw(t, "XOR")) {
			prec = P_XOR; op = "XOR";
		} else if (is_kw(t, "AND") || t->kind == T_ANDOP) {
			prec = P_AND; op = "AND";
		} else if (t->kind == T_CMP) {
			prec = P_CMP;
		} else if (is_punct(t, '+')) {
			prec = P_ADD; op = "ADD";
		} else if (is_punct(t, '-')) {
			prec = P_ADD; op = "SUB";
		} else if (is_punct(t, '*')) {
			prec = P_MUL; op = "MUL";
		} else if (is_punct(t, '/')) {
			prec = P_MUL; op = "DIV";
		} else if (is_punct(t, '%') || is_kw(t, "MOD")) {
			prec = P_MUL; op = "MOD";
		} else if (is_kw(t, "IS") || is_kw(t, "IN") || is_kw(t, "LIKE")) {
			prec = P_INISLIKE;
		} else if (is_kw(t, "NOT") && (is_kw(peek2(p), "IN") || is_kw(peek2(p), "LIKE"))) {
			prec = P_INISLIKE;
		} else if (is_kw(t, "BETWEEN")) {
			prec = P_BETWEEN;
		} else {
			break;
		}

		if (prec < min_prec)
			break;

		if (op) {
			p->pos++;
			parse_expr(p, prec + 1);
			emit(p, "%s", op);
		} else if (t->kind == T_CMP) {
			p->pos++;
			parse_expr(p, prec + 1);
			emit(p, "CMP %d", (int)t->ival);
		} else if (is_kw(t, "IS")) {
			p->pos++;
			if (accept_kw(p, "NOT")) {
				expect_kw(p, "NULL");
				emit(p, "ISNOTNULL");
			} else {
				expect_kw(p, "NULL");
				emit(p, "ISNULL");
			}
		} else if (is_kw(t, "IN")) {
			p->pos++;
			emit(p, "ISIN %d", parse_val_list(p));
		} else if (is_kw(t, "LIKE")) {
			p->pos++;
			parse_expr(p, P_NOT);
			emit(p, "LIKE");
		} else if (is_kw(t, "NOT")) {
			p->pos++;
			if (accept_kw(p, "IN")) {
				emit(p, "ISNOTIN %d", parse_val_list(p));
			} else {
				expect_kw(p, "LIKE");
				parse_expr(p, P_NOT);
				emit(p, "NOTLIKE");
			}
		} else if (is_kw(t, "BETWEEN")) {
			/* extension: the reference lexes BETWEEN but has no rule for it (midorisql.y:55,69) */
			size_t lhs_end = p->nout;
			p->pos++;
			parse_expr(p, P_CMP + 1);
			emit(p, "CMP 6");
			expect_kw(p, "AND");
			emit_copy(p, lhs_start, lhs_end);
			parse_expr(p, P_CMP + 1);
			emit(p, "CMP 5");
			emit(p, "AND");
		}
	}
}

static void parse_opt_alias(struct parser *p)
{
	if (accept_kw(p, "AS")) {
		emit(p, "ALIAS %s", expect_name(p));
	} else if (peek(p)->kind == T_NAME) {
		emit(p, "ALIAS %s", expect_name(p));
	}
}

static void parse_table_factor(struct parser *p)
{
	emit(p, "TABLE %s", expect_name(p));
	parse_opt_alias(p);
}

static void parse_table_reference(struct parser *p)
{
	parse_table_factor(p);

	while (!p->failed) {
		int code;

		if (is_kw(peek(p), "JOIN")) {
			p->pos++;
			code = 1;
		} else if (is_kw(peek(p), "INNER")) {
			p->pos++;
			expect_kw(p, "JOIN");
			code = 1;
		} else if (is_kw(peek(p), "LEFT") || is_kw(peek(p), "RIGHT")) {
			code = is_kw(peek(p), "LEFT") ? 2 : 4;
			p->pos++;
			if (accept_kw(p, "OUTER"))
				code += 6;
			expect_kw(p, "JOIN");
		} else {
			break;
		}

		parse_table_factor(p);
		expect_kw(p, "ON");
		parse_expr(p, P_OR);
		emit(p, "ONEXPR");
		emit(p, "JOIN %d", code);
	}
}

static int parse_opt_asc_desc(struct parser *p)
{
	if (accept_kw(p, "ASC"))
		return 0;
	if (accept_kw(p, "DESC"))
		return 1;
	return 0;
}

static void parse_select(struct parser *p)
{
	int opts = 0, n = 0;

	expect_kw(p, "SELECT");
	while (accept_kw(p, "DISTINCT")) {
		if (opts & 2)
			fail(p, "duplicate DISTINCT option");
		opts |= 2;
	}

	if (accept_punct(p, '*')) {
		emit(p, "SELECTALL");
		n = 1;
	} else {
		do {
			parse_expr(p, P_OR);
			parse_opt_alias(p);
			n++;
		} while (!p->failed && accept_punct(p, ','));
	}

	if (accept_kw(p, "FROM")) {
		do {
			parse_table_reference(p);
			n++;
		} while (!p->failed && accept_punct(p, ','));

		if (accept_kw(p, "WHERE")) {
			parse_expr(p, P_OR);
			emit(p, "WHERE");
			n++;
		}
		if (accept_kw(p, "GROUP")) {
			int g = 0;
			expect_kw(p, "BY");
			do {
				parse_expr(p, P_OR);
				parse_opt_asc_desc(p);
				g++;
			} while (!p->failed && accept_punct(p, ','));
			emit(p, "GROUPBYLIST %d", g);
			n++;
		}
		if (accept_kw(p, "HAVING")) {
			parse_expr(p, P_OR);
			emit(p, "HAVING");
			n++;
		}
		if (accept_kw(p, "ORDER")) {
			int o = 0;
			expect_kw(p, "BY");
			do {
				parse_expr(p, P_OR);
				emit(p, "ORDERBYITEM %d", parse_opt_asc_desc(p));
				o++;
			} while (!p->failed && accept_punct(p, ','));
			emit(p, "ORDERBYLIST %d", o);
			n++;
		}
		if (accept_kw(p, "LIMIT")) {
			parse_expr(p, P_OR);
			if (accept_punct(p, ',')) {
				parse_expr(p, P_OR);
				emit(p, "LIMIT 2");
			} else {
				emit(p, "LIMIT 1");
			}
			n++;
		}
	}
	emit(p, "SELECT %d %d", opts, n);
}

static void parse_delete(struct parser *p)
{
	char *name;

	expect_kw(p, "DELETE");
	expect_kw(p, "FROM");
	name = expect_name(p);
	if (accept_kw(p, "WHERE")) {
		parse_expr(p, P_OR);
		emit(p, "WHERE");
	}
	emit(p, "DELETEONE %s", name);
}

static int parse_column_list(struct parser *p)
{
	int n = 0;
	do {
		emit(p, "COLUMN %s", expect_name(p));
		n++;
	} while (!p->failed && accept_punct(p, ','));
	return n;
}

static void parse_insert(struct parser *p)
{
	char *name;
	int hascols = 0, ntuples = 0;

	expect_kw(p, "INSERT");
	accept_kw(p, "INTO");
	name = expect_name(p);

	if (accept_punct(p, '(')) {
		int n = parse_column_list(p);
		expect_punct(p, ')');
		emit(p, "INSERTCOLS %d", n);
		hascols = 1;
	}

	if (!accept_kw(p, "VALUES") && !accept_kw(p, "VALUE")) {
		fail(p, "syntax error: expected VALUES");
		return;
	}

	do {
		int n = 0;
		expect_punct(p, '(');
		do {
			parse_expr(p, P_ADD); /* insert_expr: literals and arithmetic only (midorisql.y:372-386) */
			n++;
		} while (!p->failed && accept_punct(p, ','));
		expect_punct(p, ')');
		emit(p, "VALUES %d", n);
		ntuples++;
	} while (!p->failed && accept_punct(p, ','));

	emit(p, "INSERTVALS %d %d %s", hascols, ntuples, name);
}

static void parse_update(struct parser *p)
{
	char *name;
	int nassign = 0, haswhere = 0;

	expect_kw(p, "UPDATE");
	name = expect_name(p);
	expect_kw(p, "SET");
	do {
		char *col = expect_name(p);
		struct tok *t = peek(p);
		if (t->kind != T_CMP || t->ival != 4) {
			fail(p, "bad insert assignment to %s", col);
			return;
		}
		p->pos++;
		parse_expr(p, P_OR);
		emit(p, "ASSIGN %s", col);
		nassign++;
	} while (!p->failed && accept_punct(p, ','));

	if (accept_kw(p, "WHERE")) {
		parse_expr(p, P_OR);
		emit(p, "WHERE");
		haswhere = 1;
	}
	emit(p, "UPDATE %s %d %d", name, nassign, haswhere);
}

static void parse_create(struct parser *p)
{
	char *name;
	int ifnotexists = 0, ncols = 0;

	expect_kw(p, "CREATE");
	expect_kw(p, "TABLE");
	if (accept_kw(p, "IF")) {
		expect_kw(p, "NOT");
		expect_kw(p, "EXISTS");
		ifnotexists = 1;
	}
	name = expect_name(p);
	expect_punct(p, '(');
	do {
		if (accept_kw(p, "PRIMARY")) {
			int n;
			expect_kw(p, "KEY");
			expect_punct(p, '(');
			n = parse_column_list(p);
			expect_punct(p, ')');
			emit(p, "PRIKEY %d", n);
		} else if (accept_kw(p, "INDEX")) {
			int n;
			expect_punct(p, '(');
			n = parse_column_list(p);
			expect_punct(p, ')');
			emit(p, "KEY %d", n);
		} else {
			char *col;
			int type = 0;

			emit(p, "STARTCOL");
			col = expect_name(p);
			if (accept_kw(p, "INT") || accept_kw(p, "INT4") || accept_kw(p, "INTEGER")) {
				type = 50000; /* INT4?|INTEGER all lex to INTEGER, midorisql.l:53 */
			} else if (accept_kw(p, "TINYINT")) {
				type = 60000;
			} else if (accept_kw(p, "DOUBLE")) {
				type = 80000;
			} else if (accept_kw(p, "DATE")) {
				type = 100000;
			} else if (accept_kw(p, "DATETIME")) {
				type = 110000;
			} else if (accept_kw(p, "VARCHAR") || accept_kw(p, "VARCHARACTER")) {
				struct tok *t;
				expect_punct(p, '(');
				t = peek(p);
				if (t->kind != T_INT) {
					fail(p, "syntax error: expected VARCHAR length");
					return;
				}
				p->pos++;
				expect_punct(p, ')');
				type = 130000 + (int)t->ival;
			} else {
				fail(p, "syntax error: unknown data type");
				return;
			}
			while (!p->failed) {
				if (accept_kw(p, "NOT")) {
					expect_kw(p, "NULL");
					emit(p, "ATTR NOTNULL");
				} else if (accept_kw(p, "NULL")) {
					/* nothing emitted, midorisql.y:467 */
				} else if (accept_kw(p, "AUTO_INCREMENT")) {
					emit(p, "ATTR AUTOINC");
				} else if (accept_kw(p, "UNIQUE")) {
					emit(p, "ATTR UNIQUEKEY");
				} else if (accept_kw(p, "PRIMARY")) {
					expect_kw(p, "KEY");
					emit(p, "ATTR PRIKEY");
				} else {
					break;
				}
			}
			emit(p, "COLUMNDEF %d %s", type, col);
		}
		ncols++;
	} while (!p->failed && accept_punct(p, ','));
	expect_punct(p, ')');
	emit(p, "CREATE %d %d %s", ifnotexists, ncols, name);
}

int mdb_sql_to_tokens(const char *sql, mdb_sql_emit_fn emit_fn, void *ctx, char *err, size_t errlen)
{
	struct parser p = {0};
	int rc = 0;

	if (!sql || !emit_fn) {
		if (err && errlen)
			snprintf(err, errlen, "invalid argument");
		return 1;
	}

	if (lex(&p, sql) == 0) {
		struct tok *t = peek(&p);
		if (is_kw(t, "SELECT"))
			parse_select(&p);
		else if (is_kw(t, "DELETE"))
			parse_delete(&p);
		else if (is_kw(t, "INSERT"))
			parse_insert(&p);
		else if (is_kw(t, "UPDATE"))
			parse_update(&p);
		else if (is_kw(t, "CREATE"))
			parse_create(&p);
		else
			fail(&p, "syntax error: unknown statement");

		emit(&p, "STMT");
		expect_punct(&p, ';');
		if (!p.failed && peek(&p)->kind != T_EOF)
			fail(&p, "syntax error: trailing input after ';'");
	}

	if (p.failed) {
		rc = 1;
		if (err && errlen)
			snprintf(err, errlen, "%s", p.err);
	} else {
		for (size_t i = 0; i < p.nout; i++) {
			if (emit_fn(ctx, p.out[i])) {
				rc = 1;
				if (err && errlen)
					snprintf(err, errlen, "token sink failed");
				break;
			}
		}
	}

	for (size_t i = 0; i < p.nout; i++)
		free(p.out[i]);
	free(p.out);
	for (size_t i = 0; i < p.ntoks; i++)
		free(p.toks[i].text);
	free(p.toks);
	return rc;
}
