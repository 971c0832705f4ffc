// mdb_ops.cu - general physical operators of the SELECT path (any plan shape the ABI can express).
//
// The reference materialises early and interprets per row (src/engine/executor_select.c); here the same
// stages run as data-parallel kernels over row-id tuples (late materialisation):
//   proc_from_clause_table :1282            -> live-row scan (bitmap compaction)
//   _join_nested_loop_tbl2tbl/_tbl2mat :1076,:1151 -> hash build (CSR multimap) + probe count/emit
//   proc_where_clause :1435, eval_row_cond :1027    -> k_eval_pred (postfix program) + compaction
//   proc_groupby_clause :1526, inc_count_cols :1501 -> k_group_update (open-addressing hash aggregate)
//   handle_countonly_case :1590                     -> the same aggregate with a single constant group
//   proc_select_clause :1369                        -> k_gather_out
// The fused fast paths for the benchmark shapes live in mdb_fast.cu.
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>
#include <new>

// =========================================================================================== scan

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <typename T>
__device__ static inline T block_exclusive_scan(T v, T *total, T *smem /* 32 entries */)
{
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	T incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		T n = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o)
			incl += n;
	}
	if (lane == 31)
		smem[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		T w = lane < nwarps ? smem[lane] : (T)0;
		T wi = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			T n = __shfl_up_sync(0xffffffffu, wi, o);
			if (lane >= o)
				wi += n;
		}
		smem[lane] = wi - w; // exclusive warp offsets
		if (lane == 31)
			smem[32] = wi; // block total
	}
	__syncthreads();
	T res = smem[warp] + incl - v;
	if (total)
		*total = smem[32];
	__syncthreads();
	return res;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const TIn *__restrict__ in, size_t n, uint64_t *__restrict__ block_sums)
{
	__shared__ uint64_t sm[33];
	size_t base = (size_t)blockIdx.x * SCAN_TILE;
	uint64_t acc = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
		if (i < n)
			acc += (uint64_t)in[i];
	}
	uint64_t total;
	block_exclusive_scan<uint64_t>(acc, &total, sm);
	if (threadIdx.x == 0)
		block_sums[blockIdx.x] = total;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const TIn *__restrict__ in, uint64_t *__restrict__ out, size_t n,
		const uint64_t *__restrict__ block_offs, uint64_t *__restrict__ d_total)
{
	__shared__ uint64_t sm[33];
	size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
	uint64_t v[SCAN_ITEMS];
	uint64_t acc = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		size_t i = base + k;
		v[k] = i < n ? (uint64_t)in[i] : 0;
		acc += v[k];
	}
	uint64_t total;
	uint64_t off = block_exclusive_scan<uint64_t>(acc, &total, sm) + (block_offs ? block_offs[blockIdx.x] : 0);
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		size_t i = base + k;
		if (i < n)
			out[i] = off;
		off += v[k];
	}
	if (d_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1)
		*d_total = off;
}

template <typename TIn>
static int scan_impl(mdbcu_ctx *ctx, const TIn *in, uint64_t *out, size_t n, uint64_t *d_total)
{
	if (n == 0) {
		if (d_total)
			CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint64_t), ctx->stream));
		return MDBCU_OK;
	}
	size_t nblocks = mdb_div_up(n, SCAN_TILE);
	if (nblocks == 1) {
		MDB_LAUNCH(ctx, k_scan_apply<TIn>, 1, SCAN_THREADS, 0, in, out, n, (const uint64_t*)nullptr, d_total);
		CUDA_CHECK_LAUNCH(ctx);
		return MDBCU_OK;
	}
	DevTemp tmp(ctx, true);
	uint64_t *sums, *offs;
	MDB_TRY(tmp.alloc(&sums, nblocks));
	MDB_TRY(tmp.alloc(&offs, nblocks));
	MDB_LAUNCH(ctx, k_scan_reduce<TIn>, (unsigned)nblocks, SCAN_THREADS, 0, in, n, sums);
	CUDA_CHECK_LAUNCH(ctx);
	MDB_TRY(scan_impl<uint64_t>(ctx, sums, offs, nblocks, nullptr));
	MDB_LAUNCH(ctx, k_scan_apply<TIn>, (unsigned)nblocks, SCAN_THREADS, 0, in, out, n, (const uint64_t*)offs, d_total);
	CUDA_CHECK_LAUNCH(ctx);
	return MDBCU_OK;
}

int mdb_scan_u32_u64(mdbcu_ctx *ctx, const uint32_t *in, uint64_t *out, size_t n, uint64_t *d_total)
{
	return scan_impl<uint32_t>(ctx, in, out, n, d_total);
}

int mdb_scan_u64_u64(mdbcu_ctx *ctx, const uint64_t *in, uint64_t *out, size_t n, uint64_t *d_total)
{
	return scan_impl<uint64_t>(ctx, in, out, n, d_total);
}

// =========================================================================================== tuples

struct Tuples {
	int ntab = 0;
	uint64_t n = 0;
	uint32_t *rid[MDBCU_MAX_TABLES] = {nullptr, nullptr, nullptr, nullptr};
};

struct TuplesDev {
	int ntab;
	uint64_t n;
	const uint32_t *rid[MDBCU_MAX_TABLES];
};

static TuplesDev to_dev(const Tuples &t)
{
	TuplesDev d;
	d.ntab = t.ntab;
	d.n = t.n;
	for (int i = 0; i < MDBCU_MAX_TABLES; i++)
		d.rid[i] = t.rid[i];
	return d;
}

static void free_tuples(mdbcu_ctx *ctx, Tuples &t)
{
	for (int i = 0; i < MDBCU_MAX_TABLES; i++) {
		mdb_free(ctx, t.rid[i]);
		t.rid[i] = nullptr;
	}
	t.n = 0;
}

// ------------------------------------------------------------------------------- bitmap compaction
// items flagged in `bits` (bit i <-> item i, i < n) are written in order to dst[k] = src[k] ? src[k][i] : i

#define CMP_THREADS 256

__global__ void __launch_bounds__(CMP_THREADS) k_compact_count(const uint32_t *__restrict__ bits, uint64_t n, uint32_t *__restrict__ block_counts)
{
	__shared__ uint32_t sm[33];
	uint64_t words = (n + 31) / 32;
	uint64_t w = (uint64_t)blockIdx.x * CMP_THREADS + threadIdx.x;
	uint32_t c = 0;
	if (w < words) {
		uint32_t valid = (w == words - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
		c = __popc(bits[w] & valid);
	}
	uint32_t total;
	block_exclusive_scan<uint32_t>(c, &total, sm);
	if (threadIdx.x == 0)
		block_counts[blockIdx.x] = total;
}

struct CompactArrays {
	int narr;
	const uint32_t *src[MDBCU_MAX_TABLES];
	uint32_t *dst[MDBCU_MAX_TABLES];
};

__global__ void __launch_bounds__(CMP_THREADS) k_compact_write(const uint32_t *__restrict__ bits, uint64_t n,
		const uint64_t *__restrict__ block_offs, CompactArrays arr)
{
	__shared__ uint32_t sm[33];
	uint64_t words = (n + 31) / 32;
	uint64_t w = (uint64_t)blockIdx.x * CMP_THREADS + threadIdx.x;
	uint32_t b = 0;
	if (w < words) {
		uint32_t valid = (w == words - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
		b = bits[w] & valid;
	}
	uint32_t off = block_exclusive_scan<uint32_t>(__popc(b), nullptr, sm);
	uint64_t o = block_offs[blockIdx.x] + off;
	while (b) {
		int bit = __ffs(b) - 1;
		b &= b - 1;
		uint64_t i = w * 32 + bit;
		for (int k = 0; k < arr.narr; k++)
			arr.dst[k][o] = arr.src[k] ? arr.src[k][i] : (uint32_t)i;
		o++;
	}
}

// compacts `in` (or the identity when in == nullptr, ntab = 1) by `bits`; result arrays are freshly allocated
static int compact_tuples(mdbcu_ctx *ctx, const uint32_t *bits, uint64_t n, const Tuples *in, int ntab, Tuples *out)
{
	out->ntab = ntab;
	out->n = 0;
	if (n == 0)
		return MDBCU_OK;
	DevTemp tmp(ctx, true);
	uint64_t words = (n + 31) / 32;
	size_t nblocks = mdb_div_up(words, CMP_THREADS);
	uint32_t *counts;
	uint64_t *offs, *d_total;
	MDB_TRY(tmp.alloc(&counts, nblocks));
	MDB_TRY(tmp.alloc(&offs, nblocks));
	MDB_TRY(tmp.alloc(&d_total, 1));
	MDB_LAUNCH(ctx, k_compact_count, (unsigned)nblocks, CMP_THREADS, 0, bits, n, counts);
	CUDA_CHECK_LAUNCH(ctx);
	MDB_TRY(mdb_scan_u32_u64(ctx, counts, offs, nblocks, d_total));
	uint64_t total = 0;
	MDB_TRY(mdb_read_u64(ctx, d_total, &total));
	out->n = total;
	if (total == 0)
		return MDBCU_OK;
	CompactArrays arr;
	arr.narr = ntab;
	for (int k = 0; k < MDBCU_MAX_TABLES; k++) {
		arr.src[k] = nullptr;
		arr.dst[k] = nullptr;
	}
	for (int k = 0; k < ntab; k++) {
		int rc = mdb_alloc(ctx, &out->rid[k], total);
		if (rc != MDBCU_OK) {
			free_tuples(ctx, *out);
			return rc;
		}
		arr.src[k] = in ? in->rid[k] : nullptr;
		arr.dst[k] = out->rid[k];
	}
	MDB_LAUNCH(ctx, k_compact_write, (unsigned)nblocks, CMP_THREADS, 0, bits, n, (const uint64_t*)offs, arr);
	CUDA_CHECK_LAUNCH(ctx);
	return MDBCU_OK;
}

__global__ void k_iota(uint32_t *out, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = (uint32_t)i;
}

static int grid_for(mdbcu_ctx *ctx, uint64_t n, int threads)
{
	uint64_t g = mdb_div_up(n ? n : 1, threads);
	return (int)std::min<uint64_t>(g, (uint64_t)ctx->num_sms * 16);
}

// FROM <table>: live rows in storage order (proc_from_clause_table, executor_select.c:1295-1306)
// identity_ok: the caller's operators all read row ids through TupleRows / k_gather_out (plans without a join): a table
// without tombstones then needs no row-id array - tuple i IS row i (rid[0] stays nullptr).  K1 of the north star
// (predicate scan -> selection vector) ran at 4.2 ms for 2^28 rows with the 1 GiB iota array written, read by the
// predicate kernel and read again by the compaction; without it the predicate kernel reads the columns coalesced.
static int scan_live(mdbcu_ctx *ctx, const mdbcu_table *t, Tuples *out, bool identity_ok)
{
	out->ntab = 1;
	out->n = 0;
	if (t->n_slots == 0)
		return MDBCU_OK;
	if (t->all_live && identity_ok) {
		out->n = t->n_slots;
		return MDBCU_OK;
	}
	if (t->all_live) {
		MDB_TRY(mdb_alloc(ctx, &out->rid[0], t->n_slots));
		MDB_LAUNCH(ctx, k_iota, grid_for(ctx, t->n_slots, 256), 256, 0, out->rid[0], t->n_slots);
		CUDA_CHECK_LAUNCH(ctx);
		out->n = t->n_slots;
		return MDBCU_OK;
	}
	return compact_tuples(ctx, t->live, t->n_slots, nullptr, 1, out);
}

// =========================================================================================== predicate

struct DPredOp {
	int32_t op, arg;
	int32_t tbl, is_dbl;
	const int64_t *data;
	const uint32_t *present; // nullptr = every cell present
	int64_t ival;
	double dval;
};

// one `<column> <cmp> <literal>` / `<column> IS [NOT] NULL` term (see DPredProgram::n_terms); cmp 7 = IS NULL, 8 = IS NOT NULL
struct DPredTerm {
	int32_t tbl, cmp;
	int32_t as_dbl;  // compare as doubles (DOUBLE column or DOUBLE literal), else as int64
	int32_t col_dbl; // the column holds doubles (raw bits)
	const int64_t *data;
	const uint32_t *present;
	int64_t ilit;
	double dlit;
};

#define PRED_MAX_TERMS 8

struct DPredProgram {
	int32_t n;
	// > 0: the program is a boolean combination (AND / OR / XOR) of n_terms <= 8 terms `<column> <cmp> <literal>` or
	// `<column> IS [NOT] NULL`, compiled on the host: the kernels evaluate the terms directly and look the verdict up in the
	// combination's truth table (bit m of `truth` = verdict when the terms' outcomes are the bits of m) - no operand stack in
	// local memory, no decoding.  conj: the combination is the AND of all terms (the common WHERE shape: BASELINE configs 2
	// and 4) - the first false term ends the evaluation.
	int32_t n_terms;
	int32_t conj;
	uint32_t truth[8];
	DPredOp ops[MDBCU_MAX_PRED];
	DPredTerm terms[PRED_MAX_TERMS];
};

struct PVal {
	int kind; // 0 int, 1 double, 2 null, 3 bool
	int64_t i;
};

__device__ static inline bool pv_cmp(int cmp, PVal a, PVal b)
{
	if (a.kind == 2 || b.kind == 2)
		return false; // executor_select.c:629-631
	if (a.kind == 1 || b.kind == 1) {
		double x = a.kind == 1 ? __longlong_as_double(a.i) : (double)a.i;
		double y = b.kind == 1 ? __longlong_as_double(b.i) : (double)b.i;
		switch (cmp) {
		case 1: return x < y;
		case 2: return x > y;
		case 3: return x != y;
		case 4: return x == y;
		case 5: return x <= y;
		case 6: return x >= y;
		}
		return false;
	}
	switch (cmp) {
	case 1: return a.i < b.i;
	case 2: return a.i > b.i;
	case 3: return a.i != b.i;
	case 4: return a.i == b.i;
	case 5: return a.i <= b.i;
	case 6: return a.i >= b.i;
	}
	return false;
}

#define PRED_STACK 24

// Row ids of one joined tuple, by table. The general operators read them from the materialised tuple arrays ...
struct TupleRows {
	const TuplesDev *ts;
	uint64_t i;
	// (a missing array = the identity: the rows of a single, fully live table need no row-id array at all)
	__device__ uint32_t operator()(int t) const
	{
		const uint32_t *c = ts->rid[t];
		return c ? c[i] : (uint32_t)i;
	}
};
// ... the fused multiway aggregate holds them in registers.
struct StarRows {
	uint32_t r[MDBCU_MAX_TABLES];
	__device__ uint32_t operator()(int t) const { return r[t]; }
};

template <typename Rows>
__device__ static bool eval_program(const DPredProgram *__restrict__ prog, const Rows &rows)
{
	if (prog->n_terms > 0) {
		uint32_t outcome = 0;
		for (int k = 0; k < prog->n_terms; k++) {
			const DPredTerm &t = prog->terms[k];
			const uint32_t r = rows(t.tbl);
			const bool present = !t.present || mdb_bit(t.present, r);
			bool ok;
			if (t.cmp >= 7) {
				ok = present == (t.cmp == 8);
			} else if (!present) {
				ok = false; // NULL operand: the comparison is not true (executor_select.c:629-631)
			} else {
				const long long v = t.data[r];
				if (t.as_dbl) {
					const double x = t.col_dbl ? __longlong_as_double(v) : (double)v, y = t.dlit;
					ok = t.cmp == 1 ? x < y : t.cmp == 2 ? x > y : t.cmp == 3 ? x != y : t.cmp == 4 ? x == y : t.cmp == 5 ? x <= y : x >= y;
				} else {
					const long long y = t.ilit;
					ok = t.cmp == 1 ? v < y : t.cmp == 2 ? v > y : t.cmp == 3 ? v != y : t.cmp == 4 ? v == y : t.cmp == 5 ? v <= y : v >= y;
				}
			}
			if (prog->conj && !ok)
				return false;
			outcome |= (ok ? 1u : 0u) << k;
		}
		return prog->conj || ((prog->truth[outcome >> 5] >> (outcome & 31u)) & 1u) != 0;
	}
	PVal st[PRED_STACK];
	int sp = 0;
	for (int k = 0; k < prog->n; k++) {
		const DPredOp &op = prog->ops[k];
		PVal v;
		v.kind = 3;
		v.i = 0;
		switch (op.op) {
		case MDBCU_P_COL: {
			uint32_t r = rows(op.tbl);
			if (op.present && !mdb_bit(op.present, r)) {
				v.kind = 2;
			} else {
				v.kind = op.is_dbl ? 1 : 0;
				v.i = op.data[r];
			}
			st[sp++] = v;
			break;
		}
		case MDBCU_P_INT:
			v.kind = 0; v.i = op.ival; st[sp++] = v;
			break;
		case MDBCU_P_DBL:
			v.kind = 1; v.i = __double_as_longlong(op.dval); st[sp++] = v;
			break;
		case MDBCU_P_NULL:
			v.kind = 2; st[sp++] = v;
			break;
		case MDBCU_P_BOOL:
			v.kind = 3; v.i = op.ival != 0; st[sp++] = v;
			break;
		case MDBCU_P_CMP: {
			PVal b = st[--sp], a = st[--sp];
			v.i = pv_cmp(op.arg, a, b); st[sp++] = v;
			break;
		}
		case MDBCU_P_AND: case MDBCU_P_OR: case MDBCU_P_XOR: {
			PVal b = st[--sp], a = st[--sp];
			bool x = a.i != 0, y = b.i != 0;
			v.i = op.op == MDBCU_P_AND ? (x && y) : (op.op == MDBCU_P_OR ? (x || y) : (x != y));
			st[sp++] = v;
			break;
		}
		case MDBCU_P_ISNULL: case MDBCU_P_ISNOTNULL: {
			PVal a = st[--sp];
			v.i = (a.kind == 2) != (op.op == MDBCU_P_ISNOTNULL); st[sp++] = v;
			break;
		}
		case MDBCU_P_IN: case MDBCU_P_NOTIN: {
			int n = op.arg;
			PVal probe = st[sp - n - 1];
			bool any = false, all_diff = true;
			for (int j = 0; j < n; j++) {
				PVal e = st[sp - n + j];
				any = any || pv_cmp(4, probe, e);
				all_diff = all_diff && pv_cmp(3, probe, e);
			}
			sp -= n + 1;
			v.i = op.op == MDBCU_P_IN ? any : all_diff; st[sp++] = v;
			break;
		}
		}
	}
	return sp == 1 && st[0].i != 0;
}

__global__ void k_eval_pred(const __grid_constant__ DPredProgram prog_, const __grid_constant__ TuplesDev ts, uint32_t *__restrict__ bits)
{
	const DPredProgram *prog = &prog_; // (a kernel parameter: the program is read from the constant bank, not from global memory)
	// one aligned group of 32 tuples per warp iteration, so every ballot is exactly one bitmap word
	uint64_t groups = (ts.n + 31) / 32;
	uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	int lane = threadIdx.x & 31;
	for (uint64_t g = warp; g < groups; g += nwarps) {
		uint64_t i = g * 32 + lane;
		TupleRows rows = {&ts, i};
		bool keep = i < ts.n && eval_program(prog, rows);
		uint32_t b = __ballot_sync(0xffffffffu, keep);
		if (lane == 0)
			bits[g] = b;
	}
}

// K1 of the north star - the vectorised predicate scan: the rows of ONE fully live table (identity tuples, scan_live) against
// a host-compiled term program (DPredProgram::n_terms > 0).  Every thread takes 8 consecutive rows per step: two 256-bit
// loads per referenced column (a term on the column the previous term used reuses the registers), the terms' 8 outcomes
// as byte masks, the truth table applied per row, and the warp's 256 verdicts leave as 8 coalesced bitmap words.  The
// program is decoded once per 8 rows from shared memory.  HBM-bound: 8 B per row and referenced column in, 1 bit out
// (the per-row kernel read the terms' fields from global memory for every row and ran at 1.2 TB/s).
#define EVT_THREADS 256

__device__ __forceinline__ void evt_load256(const void *p, uint32_t *w)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			: "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

__global__ void __launch_bounds__(EVT_THREADS) k_eval_terms_scan(const DPredProgram *__restrict__ prog, uint64_t n, uint32_t *__restrict__ bits)
{
	__shared__ DPredTerm s_terms[PRED_MAX_TERMS];
	__shared__ uint32_t s_truth[8];
	__shared__ int s_nt, s_conj;
	if (threadIdx.x == 0) {
		s_nt = prog->n_terms;
		s_conj = prog->conj;
	}
	if (threadIdx.x < 8)
		s_truth[threadIdx.x] = prog->truth[threadIdx.x];
	for (uint32_t i = threadIdx.x; i < PRED_MAX_TERMS * sizeof(DPredTerm) / 4; i += EVT_THREADS)
		reinterpret_cast<uint32_t*>(s_terms)[i] = reinterpret_cast<const uint32_t*>(prog->terms)[i];
	__syncthreads();
	const int nt = s_nt;
	const bool conj = s_conj != 0;
	const uint32_t lane = threadIdx.x & 31u;
	// one warp step = 256 rows = 8 bitmap words; whole steps only (the caller's n is padded: rows beyond n are masked below)
	const uint64_t steps = (n + 255) / 256;
	const uint64_t warp = (blockIdx.x * (uint64_t)EVT_THREADS + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * EVT_THREADS) >> 5;
	for (uint64_t st = warp; st < steps; st += nwarps) {
		const uint64_t row0 = st * 256 + lane * 8u;
		uint32_t raw[16]; // 8 cells of the current column: [2 * j] low word, [2 * j + 1] high word
		const int64_t *loaded = nullptr;
		uint32_t outcome[8]; // per row: bit k = term k true
#pragma unroll
		for (int j = 0; j < 8; j++)
			outcome[j] = 0;
		uint32_t alive = 0xffu; // conj: rows that have not failed a term yet
		for (int k = 0; k < nt; k++) {
			const DPredTerm &t = s_terms[k];
			if (t.data != loaded) { // (uniform)
				if (row0 + 8 <= n) {
					evt_load256(t.data + row0, raw);
					evt_load256(t.data + row0 + 4, raw + 8);
				} else {
#pragma unroll
					for (int j = 0; j < 8; j++) {
						const unsigned long long v = row0 + j < n ? (unsigned long long)t.data[row0 + j] : 0ull;
						raw[2 * j] = (uint32_t)v;
						raw[2 * j + 1] = (uint32_t)(v >> 32);
					}
				}
				loaded = t.data;
			}
			// present bits of the 8 rows (bit = live and not NULL); row0 is a multiple of 8: one byte of the bitmap
			const uint32_t pres = t.present ? (row0 < n ? reinterpret_cast<const uint8_t*>(t.present)[row0 >> 3] : 0u) : 0xffu;
			uint32_t okmask = 0;
			if (t.cmp >= 7) {
				okmask = t.cmp == 8 ? pres : (~pres & 0xffu);
			} else {
				// the comparison is chosen ONCE per term and 8 rows (all branches here are uniform): per row one compare and one OR
#define EVT_ROWS(EXPR)                                                                                                   \
	_Pragma("unroll") for (int j = 0; j < 8; j++) {                                                                  \
		const long long v = (long long)(((unsigned long long)raw[2 * j + 1] << 32) | raw[2 * j]);                \
		okmask |= ((EXPR) ? 1u : 0u) << j;                                                                       \
	}
#define EVT_CMPS(X, Y)                                                                                                   \
	switch (t.cmp) {                                                                                                 \
	case 1: EVT_ROWS((X) < (Y)) break;                                                                               \
	case 2: EVT_ROWS((X) > (Y)) break;                                                                               \
	case 3: EVT_ROWS((X) != (Y)) break;                                                                              \
	case 4: EVT_ROWS((X) == (Y)) break;                                                                              \
	case 5: EVT_ROWS((X) <= (Y)) break;                                                                              \
	default: EVT_ROWS((X) >= (Y)) break;                                                                             \
	}
				if (!t.as_dbl) {
					const long long y = t.ilit;
					EVT_CMPS(v, y)
				} else if (t.col_dbl) {
					const double y = t.dlit;
					EVT_CMPS(__longlong_as_double(v), y)
				} else {
					const double y = t.dlit;
					EVT_CMPS((double)v, y)
				}
#undef EVT_CMPS
#undef EVT_ROWS
				okmask &= pres; // NULL operand: the comparison is not true (executor_select.c:629-631)
			}
			if (conj) {
				alive &= okmask;
			} else {
#pragma unroll
				for (int j = 0; j < 8; j++)
					outcome[j] |= ((okmask >> j) & 1u) << k;
			}
		}
		uint32_t verdict = 0; // 8 bits
		if (conj) {
			verdict = alive;
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++)
				verdict |= ((s_truth[outcome[j] >> 5] >> (outcome[j] & 31u)) & 1u) << j;
		}
		// rows beyond n
		if (row0 + 8 > n)
			verdict &= row0 < n ? ((1u << (n - row0)) - 1u) : 0u;
		// lanes 4g .. 4g+3 hold the four bytes of bitmap word g of this step
		uint32_t word = verdict << (8u * (lane & 3u));
		word |= __shfl_xor_sync(0xffffffffu, word, 1);
		word |= __shfl_xor_sync(0xffffffffu, word, 2);
		const uint64_t w = st * 8 + (lane >> 2);
		if ((lane & 3u) == 0 && w < (n + 31) / 32)
			bits[w] = word;
	}
}

static int check_colref(mdbcu_ctx *ctx, const mdbcu_plan *plan, int tbl, int col, const char *what)
{
	if (tbl < 0 || tbl >= plan->n_tables || col < 0 || col >= plan->tables[tbl]->ncols)
		return mdb_fail(ctx, MDBCU_EERROR, "%s references table %d column %d outside the plan", what, tbl, col);
	if (plan->tables[tbl]->cols[col].type == MDBCU_CT_VARCHAR)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "%s uses a VARCHAR column (not mirrored on the device)", what);
	return MDBCU_OK;
}

static bool col_all_present(const mdbcu_table *t, int col)
{
	return t->all_live && !t->cols[col].has_nulls;
}

// Is the postfix program a boolean combination (AND / OR / XOR, any nesting) of at most 8 terms `<column> <literal> CMP`,
// `<literal> <column> CMP`, `<column> IS NULL`, `<column> IS NOT NULL`?  Then fill h->terms and the combination's truth table.
// A host-side stack walks the program once; a boolean item carries ITS truth table over the terms seen so far, so AND / OR /
// XOR of two items is the bitwise operation on their tables.
static void compile_terms(DPredProgram *h)
{
	enum { COL, LIT, BOOL };
	struct Item {
		int kind, op;
		uint32_t tt[8];
	} st[PRED_STACK + 1];
	int sp = 0, nt = 0;
	DPredTerm terms[PRED_MAX_TERMS];
	auto push_term = [&](int k) {
		Item it;
		it.kind = BOOL;
		it.op = 0;
		for (uint32_t m = 0; m < 256; m++) {
			if ((m & 31u) == 0)
				it.tt[m >> 5] = 0;
			if ((m >> k) & 1u)
				it.tt[m >> 5] |= 1u << (m & 31u);
		}
		st[sp++] = it;
	};
	for (int k = 0; k < h->n; k++) {
		const DPredOp &o = h->ops[k];
		switch (o.op) {
		case MDBCU_P_COL:
			st[sp].kind = COL;
			st[sp++].op = k;
			break;
		case MDBCU_P_INT: case MDBCU_P_DBL:
			st[sp].kind = LIT;
			st[sp++].op = k;
			break;
		case MDBCU_P_CMP: {
			if (sp < 2 || nt == PRED_MAX_TERMS)
				return;
			const Item b = st[--sp], a = st[--sp];
			if (!((a.kind == COL && b.kind == LIT) || (a.kind == LIT && b.kind == COL)))
				return;
			const DPredOp &c = h->ops[a.kind == COL ? a.op : b.op], &l = h->ops[a.kind == COL ? b.op : a.op];
			static const int flipped[7] = {0, 2, 1, 3, 4, 6, 5}; // literal on the left: a < b  <=>  b > a
			DPredTerm &t = terms[nt];
			t.tbl = c.tbl;
			t.cmp = a.kind == COL ? o.arg : flipped[o.arg];
			t.col_dbl = c.is_dbl;
			t.as_dbl = c.is_dbl || l.op == MDBCU_P_DBL;
			t.data = c.data;
			t.present = c.present;
			t.ilit = l.ival;
			t.dlit = l.op == MDBCU_P_DBL ? l.dval : (double)l.ival;
			push_term(nt++);
			break;
		}
		case MDBCU_P_ISNULL: case MDBCU_P_ISNOTNULL: {
			if (sp < 1 || nt == PRED_MAX_TERMS || st[sp - 1].kind != COL)
				return;
			const DPredOp &c = h->ops[st[--sp].op];
			DPredTerm &t = terms[nt];
			memset(&t, 0, sizeof(t));
			t.tbl = c.tbl;
			t.cmp = o.op == MDBCU_P_ISNULL ? 7 : 8;
			t.data = c.data;
			t.present = c.present;
			push_term(nt++);
			break;
		}
		case MDBCU_P_AND: case MDBCU_P_OR: case MDBCU_P_XOR: {
			if (sp < 2 || st[sp - 1].kind != BOOL || st[sp - 2].kind != BOOL)
				return;
			const Item b = st[--sp];
			Item &a = st[sp - 1];
			for (int w = 0; w < 8; w++)
				a.tt[w] = o.op == MDBCU_P_AND ? (a.tt[w] & b.tt[w]) : o.op == MDBCU_P_OR ? (a.tt[w] | b.tt[w]) : (a.tt[w] ^ b.tt[w]);
			break;
		}
		default:
			return;
		}
	}
	if (sp != 1 || st[0].kind != BOOL || nt == 0)
		return;
	for (int k = 0; k < nt; k++)
		h->terms[k] = terms[k];
	// only the outcomes of the nt terms that exist matter: bit m of the table for m < 2^nt
	const uint32_t all = (1u << nt) - 1u;
	bool conj = true;
	for (uint32_t m = 0; m <= all; m++)
		if ((((st[0].tt[m >> 5] >> (m & 31u)) & 1u) != 0) != (m == all))
			conj = false;
	for (int w = 0; w < 8; w++)
		h->truth[w] = st[0].tt[w];
	h->conj = conj ? 1 : 0;
	h->n_terms = nt;
}

static int build_pred(mdbcu_ctx *ctx, const mdbcu_plan *plan, DPredProgram *h)
{
	int depth = 0;
	h->n = plan->n_pred;
	h->n_terms = 0;
	h->conj = 0;
	if (plan->n_pred < 0 || plan->n_pred > MDBCU_MAX_PRED)
		return mdb_fail(ctx, MDBCU_EERROR, "predicate program too long");
	for (int k = 0; k < plan->n_pred; k++) {
		const mdbcu_pred_op &s = plan->pred[k];
		DPredOp &d = h->ops[k];
		d.op = s.op;
		d.arg = s.arg;
		d.tbl = s.tbl;
		d.is_dbl = 0;
		d.data = nullptr;
		d.present = nullptr;
		d.ival = s.ival;
		d.dval = s.dval;
		switch (s.op) {
		case MDBCU_P_COL: {
			MDB_TRY(check_colref(ctx, plan, s.tbl, s.col, "WHERE"));
			const mdbcu_table *t = plan->tables[s.tbl];
			d.data = t->cols[s.col].data;
			d.present = col_all_present(t, s.col) ? nullptr : t->cols[s.col].present;
			d.is_dbl = t->cols[s.col].type == MDBCU_CT_DOUBLE;
			depth++;
			break;
		}
		case MDBCU_P_INT: case MDBCU_P_DBL: case MDBCU_P_NULL: case MDBCU_P_BOOL:
			depth++;
			break;
		case MDBCU_P_CMP:
			if (s.arg < 1 || s.arg > 6)
				return mdb_fail(ctx, MDBCU_EERROR, "bad comparison code %d", s.arg);
			/* fallthrough */
		case MDBCU_P_AND: case MDBCU_P_OR: case MDBCU_P_XOR:
			if (depth < 2)
				return mdb_fail(ctx, MDBCU_EERROR, "malformed predicate program");
			depth--;
			break;
		case MDBCU_P_ISNULL: case MDBCU_P_ISNOTNULL:
			if (depth < 1)
				return mdb_fail(ctx, MDBCU_EERROR, "malformed predicate program");
			break;
		case MDBCU_P_IN: case MDBCU_P_NOTIN:
			if (s.arg < 1 || depth < s.arg + 1)
				return mdb_fail(ctx, MDBCU_EERROR, "malformed predicate program");
			depth -= s.arg;
			break;
		default:
			return mdb_fail(ctx, MDBCU_EERROR, "unknown predicate op %d", s.op);
		}
		if (depth > PRED_STACK)
			return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "predicate nests deeper than %d operands", PRED_STACK);
	}
	if (plan->n_pred && depth != 1)
		return mdb_fail(ctx, MDBCU_EERROR, "malformed predicate program");
	compile_terms(h);
	return MDBCU_OK;
}

// WHERE verdicts of the tuples as a bitmap in `tmp` (nullptr: no predicate, every tuple qualifies)
static int eval_pred_bits(mdbcu_ctx *ctx, const mdbcu_plan *plan, const Tuples &ts, DevTemp &tmp, uint32_t **bits_out)
{
	*bits_out = nullptr;
	if (plan->n_pred == 0 || ts.n == 0)
		return MDBCU_OK;
	DPredProgram h;
	MDB_TRY(build_pred(ctx, plan, &h));
	DPredProgram *d_prog;
	uint32_t *bits;
	MDB_TRY(tmp.alloc(&d_prog, 1));
	MDB_TRY(tmp.alloc(&bits, (ts.n + 31) / 32));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_prog, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
	bool vectorised = ts.ntab == 1 && ts.rid[0] == nullptr && h.n_terms > 0; // identity tuples + term program: the predicate SCAN
	for (int k = 0; vectorised && k < h.n_terms; k++)
		vectorised = ((uintptr_t)h.terms[k].data & 31u) == 0;
	if (vectorised)
		MDB_LAUNCH(ctx, k_eval_terms_scan, (int)std::min<uint64_t>(mdb_div_up(ts.n, (uint64_t)EVT_THREADS * 8), (uint64_t)ctx->num_sms * 8),
				EVT_THREADS, 0, (const DPredProgram*)d_prog, (uint64_t)ts.n, bits);
	else
		MDB_LAUNCH(ctx, k_eval_pred, grid_for(ctx, ts.n, 256), 256, 0, h, to_dev(ts), bits);
	CUDA_CHECK_LAUNCH(ctx);
	*bits_out = bits;
	return MDBCU_OK;
}

static int filter_tuples(mdbcu_ctx *ctx, const mdbcu_plan *plan, Tuples *ts)
{
	DevTemp tmp(ctx, true);
	uint32_t *bits;
	MDB_TRY(eval_pred_bits(ctx, plan, *ts, tmp, &bits));
	if (!bits)
		return MDBCU_OK;
	Tuples out;
	MDB_TRY(compact_tuples(ctx, bits, ts->n, ts, ts->ntab, &out));
	free_tuples(ctx, *ts);
	*ts = out;
	return MDBCU_OK;
}

// =========================================================================================== hash join

#define HT_EMPTY ((long long)0x8000000000000000LL) // INT64_MIN is the empty marker; that key value uses slot `cap`

// normalise a key for equality hashing: -0.0 == +0.0 for doubles
__device__ static inline long long norm_key(int64_t v, int is_dbl)
{
	if (is_dbl && v == (int64_t)0x8000000000000000LL)
		return 0;
	return v;
}

__device__ static inline bool key_is_nan(int64_t v)
{
	return (v & 0x7ff0000000000000LL) == 0x7ff0000000000000LL && (v & 0x000fffffffffffffLL) != 0;
}

__device__ static inline uint64_t ht_insert(long long *keys, uint64_t cap_mask, long long key)
{
	if (key == HT_EMPTY)
		return cap_mask + 1;
	uint64_t s = mdb_mix64((uint64_t)key) & cap_mask;
	while (true) {
		long long prev = keys[s];
		if (prev == key)
			return s;
		if (prev == HT_EMPTY) {
			prev = atomicCAS((unsigned long long*)&keys[s], (unsigned long long)HT_EMPTY, (unsigned long long)key);
			if (prev == HT_EMPTY || prev == key)
				return s;
		}
		s = (s + 1) & cap_mask;
	}
}

__device__ static inline bool ht_find(const long long *keys, uint64_t cap_mask, long long key, uint64_t *slot)
{
	if (key == HT_EMPTY) {
		*slot = cap_mask + 1;
		return true; // the caller checks the count of the special slot
	}
	uint64_t s = mdb_mix64((uint64_t)key) & cap_mask;
	while (true) {
		long long prev = keys[s];
		if (prev == key) {
			*slot = s;
			return true;
		}
		if (prev == HT_EMPTY)
			return false;
		s = (s + 1) & cap_mask;
	}
}

__global__ void k_fill_i64(long long *p, uint64_t n, long long v)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		p[i] = v;
}

__global__ void k_join_build_count(const int64_t *__restrict__ data, const uint32_t *__restrict__ present, uint64_t n, int is_dbl,
		long long *__restrict__ keys, uint64_t cap_mask, uint32_t *__restrict__ cnt, unsigned long long *__restrict__ dup)
{
	for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
		if (present && !mdb_bit(present, r))
			continue; // NULL (or tombstoned) rows never match, executor_select.c:716-738
		int64_t v = data[r];
		if (is_dbl && key_is_nan(v))
			continue;
		uint64_t s = ht_insert(keys, cap_mask, norm_key(v, is_dbl));
		if (atomicAdd(&cnt[s], 1u))
			*dup = 1; // some key occurs twice
	}
}

__global__ void k_join_build_fill(const int64_t *__restrict__ data, const uint32_t *__restrict__ present, uint64_t n, int is_dbl,
		const long long *__restrict__ keys, uint64_t cap_mask, const uint64_t *__restrict__ off, uint32_t *__restrict__ fill,
		uint32_t *__restrict__ rows)
{
	for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
		if (present && !mdb_bit(present, r))
			continue;
		int64_t v = data[r];
		if (is_dbl && key_is_nan(v))
			continue;
		uint64_t s;
		if (!ht_find(keys, cap_mask, norm_key(v, is_dbl), &s))
			continue;
		uint32_t pos = atomicAdd(&fill[s], 1u);
		rows[off[s] + pos] = (uint32_t)r;
	}
}

__global__ void k_join_probe_count(TuplesDev ts, int ltbl, const int64_t *__restrict__ data, const uint32_t *__restrict__ present,
		int is_dbl, const long long *__restrict__ keys, uint64_t cap_mask, const uint32_t *__restrict__ cnt,
		const uint64_t *__restrict__ row_off, uint32_t *__restrict__ matches, uint64_t *__restrict__ first_row)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < ts.n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t r = ts.rid[ltbl][i];
		uint32_t m = 0;
		uint64_t s = 0;
		if (!present || mdb_bit(present, r)) {
			int64_t v = data[r];
			if (!(is_dbl && key_is_nan(v)) && ht_find(keys, cap_mask, norm_key(v, is_dbl), &s))
				m = cnt[s];
		}
		matches[i] = m;
		first_row[i] = m ? row_off[s] : 0;
	}
}

// how a probe finds the partner of a key when the build side has no duplicate keys
struct JoinLookup {
	// direct != nullptr: narrow INT key range, row = direct[key - dmin] (0xffffffff = no such key)
	const uint32_t *direct;
	long long dmin;
	uint64_t drange;
	// otherwise the hash table of the general route
	const long long *keys;
	uint64_t cap_mask;
	const uint32_t *cnt;
	const uint64_t *row_off;
	const uint32_t *rows;
};

#define JOIN_NO_ROW 0xffffffffu

__global__ void k_join_build_direct(const int64_t *__restrict__ data, const uint32_t *__restrict__ present, uint64_t n, long long kmin,
		uint64_t range, uint32_t *__restrict__ direct, unsigned long long *__restrict__ dup)
{
	for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
		if (present && !mdb_bit(present, r))
			continue; // NULL (or tombstoned) rows never match, executor_select.c:716-738
		uint64_t d = (uint64_t)data[r] - (uint64_t)kmin;
		if (d >= range)
			continue; // outside the zone map: cannot happen, the bounds are supersets
		if (atomicCAS(&direct[d], JOIN_NO_ROW, (uint32_t)r) != JOIN_NO_ROW)
			*dup = 1; // some key occurs twice
	}
}

// build side without duplicate keys: a tuple matches at most one row, so the join appends one row-id column (JOIN_NO_ROW =
// no match) and a keep bitmap instead of expanding the tuples; n_match tells the host whether anything has to be dropped
__global__ void k_join_probe_unique(TuplesDev ts, int ltbl, const int64_t *__restrict__ data, const uint32_t *__restrict__ present,
		int is_dbl, JoinLookup lk, uint32_t *__restrict__ newcol, uint32_t *__restrict__ keep, unsigned long long *__restrict__ n_match)
{
	uint32_t mine = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x; // a multiple of 32: the lanes of a warp stay on one bitmap word
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; (i & ~31ull) < ts.n; i += stride) {
		uint32_t row = JOIN_NO_ROW;
		if (i < ts.n) {
			uint32_t r = ts.rid[ltbl][i];
			if (!present || mdb_bit(present, r)) {
				int64_t v = data[r];
				uint64_t s = 0;
				if (lk.direct) {
					uint64_t d = (uint64_t)v - (uint64_t)lk.dmin;
					if (d < lk.drange)
						row = lk.direct[d];
				} else if (!(is_dbl && key_is_nan(v)) && ht_find(lk.keys, lk.cap_mask, norm_key(v, is_dbl), &s) && lk.cnt[s]) {
					row = lk.rows[lk.row_off[s]];
				}
			}
			newcol[i] = row;
		}
		uint32_t b = __ballot_sync(0xffffffffu, row != JOIN_NO_ROW);
		if ((threadIdx.x & 31) == 0) {
			keep[i >> 5] = b;
			mine += __popc(b);
		}
	}
	if ((threadIdx.x & 31) == 0 && mine)
		atomicAdd(n_match, (unsigned long long)mine);
}

__global__ void k_join_probe_emit(TuplesDev ts, const uint32_t *__restrict__ matches, const uint64_t *__restrict__ first_row,
		const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ rows, CompactArrays dst)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < ts.n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t m = matches[i];
		if (!m)
			continue;
		uint64_t o = out_off[i], ro = first_row[i];
		for (uint32_t j = 0; j < m; j++) {
			for (int k = 0; k < ts.ntab; k++)
				dst.dst[k][o + j] = ts.rid[k][i];
			dst.dst[ts.ntab][o + j] = rows[ro + j];
		}
	}
}

__global__ void k_cross_emit(TuplesDev ts, const uint32_t *__restrict__ right, uint64_t nright, CompactArrays dst)
{
	uint64_t total = ts.n * nright;
	for (uint64_t o = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; o < total; o += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t i = o / nright, j = o - i * nright;
		for (int k = 0; k < ts.ntab; k++)
			dst.dst[k][o] = ts.rid[k][i];
		dst.dst[ts.ntab][o] = right[j];
	}
}

static int alloc_out_tuples(mdbcu_ctx *ctx, Tuples *out, int ntab, uint64_t n, CompactArrays *arr)
{
	if (n >= (1ull << 32))
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "join produces %llu rows; the general path is limited to 2^32-1",
				(unsigned long long)n);
	out->ntab = ntab;
	out->n = n;
	arr->narr = ntab;
	for (int k = 0; k < MDBCU_MAX_TABLES; k++) {
		arr->src[k] = nullptr;
		arr->dst[k] = nullptr;
	}
	for (int k = 0; k < ntab; k++) {
		int rc = mdb_alloc(ctx, &out->rid[k], n);
		if (rc != MDBCU_OK) {
			free_tuples(ctx, *out);
			return rc;
		}
		arr->dst[k] = out->rid[k];
	}
	return MDBCU_OK;
}

static int join_step(mdbcu_ctx *ctx, const mdbcu_plan *plan, int j, Tuples *ts)
{
	const mdbcu_join &jn = plan->joins[j];
	const mdbcu_table *rt = plan->tables[j + 1];
	Tuples out;
	CompactArrays arr;

	if (ts->ntab != j + 1)
		return mdb_fail(ctx, MDBCU_EINTERNAL, "join order mismatch");

	if (jn.cross) {
		// comma list / ON 1=1: every pair (wrap_on_join_node, optimiser_select.c:395)
		Tuples right;
		MDB_TRY(scan_live(ctx, rt, &right, false));
		uint64_t total = ts->n * right.n;
		int rc = alloc_out_tuples(ctx, &out, ts->ntab + 1, total, &arr);
		if (rc == MDBCU_OK && total) {
			MDB_LAUNCH(ctx, k_cross_emit, grid_for(ctx, total, 256), 256, 0, to_dev(*ts), (const uint32_t*)right.rid[0],
					right.n, arr);
			cudaError_t e = cudaGetLastError();
			if (e != cudaSuccess)
				rc = mdb_fail(ctx, MDBCU_ECUDA, "k_cross_emit: %s", cudaGetErrorString(e));
		}
		free_tuples(ctx, right);
		if (rc != MDBCU_OK)
			return rc;
		free_tuples(ctx, *ts);
		*ts = out;
		return MDBCU_OK;
	}

	MDB_TRY(check_colref(ctx, plan, jn.left.tbl, jn.left.col, "JOIN"));
	MDB_TRY(check_colref(ctx, plan, jn.right.tbl, jn.right.col, "JOIN"));
	if (jn.right.tbl != j + 1 || jn.left.tbl > j)
		return mdb_fail(ctx, MDBCU_EERROR, "join %d must compare a column of tables[0..%d] with one of tables[%d]", j, j, j + 1);
	const mdbcu_table *lt = plan->tables[jn.left.tbl];
	const DevColumn &lc = lt->cols[jn.left.col], &rc_ = rt->cols[jn.right.col];
	int l_dbl = lc.type == MDBCU_CT_DOUBLE, r_dbl = rc_.type == MDBCU_CT_DOUBLE;
	if (l_dbl != r_dbl)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "join key types differ (INT vs DOUBLE)");

	if (ts->n == 0 || rt->n_slots == 0) {
		free_tuples(ctx, *ts);
		ts->ntab = j + 2;
		return MDBCU_OK;
	}

	HostLap lap;
	DevTemp tmp(ctx, true);
	const uint32_t *r_present = col_all_present(rt, jn.right.col) ? nullptr : rc_.present;
	const uint32_t *l_present = col_all_present(lt, jn.left.col) ? nullptr : lc.present;
	const int gr = grid_for(ctx, rt->n_slots, 256), gp = grid_for(ctx, ts->n, 256);
	uint64_t *d_total;
	unsigned long long *d_dup;
	MDB_TRY(tmp.alloc(&d_total, 1));
	MDB_TRY(tmp.alloc(&d_dup, 1));

	// Build sides without duplicate keys (the foreign-key shape): the tuples keep their arrays and gain one column; they
	// are compacted only if some tuple has no partner.
	JoinLookup lk;
	memset(&lk, 0, sizeof(lk));
	auto probe_unique = [&]() -> int {
		uint32_t *newcol = nullptr, *keep;
		MDB_TRY(tmp.alloc(&keep, (ts->n + 31) / 32));
		CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(*d_total), ctx->stream));
		MDB_TRY(mdb_alloc(ctx, &newcol, ts->n));
		MDB_LAUNCH(ctx, k_join_probe_unique, gp, 256, 0, to_dev(*ts), jn.left.tbl, (const int64_t*)lc.data, l_present, l_dbl, lk,
				newcol, keep, (unsigned long long*)d_total);
		uint64_t matched = 0;
		int rc = cudaGetLastError() == cudaSuccess ? mdb_read_u64(ctx, d_total, &matched)
				: mdb_fail(ctx, MDBCU_ECUDA, "k_join_probe_unique failed to launch");
		lap(lk.direct ? "join: direct probe (sync)" : "join: unique probe (sync)", matched);
		ts->rid[ts->ntab] = newcol; // from here on the column belongs to the tuples
		ts->ntab++;
		if (rc != MDBCU_OK || matched == ts->n)
			return rc;
		Tuples kept;
		rc = compact_tuples(ctx, keep, ts->n, ts, ts->ntab, &kept);
		free_tuples(ctx, *ts);
		if (rc == MDBCU_OK)
			*ts = kept;
		else
			ts->ntab = j + 2;
		return rc;
	};

	// 1. INT keys whose zone map is narrow: a direct row table instead of a hash table (one 4-byte lookup per tuple)
	uint64_t dup = 0;
	bool dup_known = false;
	if (!r_dbl && rc_.stats_ok && rc_.imin <= rc_.imax) {
		const unsigned long long range = (unsigned long long)rc_.imax - (unsigned long long)rc_.imin + 1ull;
		if (range != 0 && range <= std::max<unsigned long long>(1ull << 16, 4ull * rt->n_slots)) {
			uint32_t *direct;
			MDB_TRY(tmp.alloc(&direct, range));
			CUDA_TRY(ctx, cudaMemsetAsync(direct, 0xff, range * sizeof(uint32_t), ctx->stream));
			CUDA_TRY(ctx, cudaMemsetAsync(d_dup, 0, sizeof(*d_dup), ctx->stream));
			MDB_LAUNCH(ctx, k_join_build_direct, gr, 256, 0, (const int64_t*)rc_.data, r_present, rt->n_slots, (long long)rc_.imin,
					(uint64_t)range, direct, d_dup);
			CUDA_CHECK_LAUNCH(ctx);
			MDB_TRY(mdb_read_u64(ctx, (const uint64_t*)d_dup, &dup));
			dup_known = true;
			lap("join: direct build (sync)", range);
			if (!dup) {
				lk.direct = direct;
				lk.dmin = rc_.imin;
				lk.drange = range;
				return probe_unique();
			}
		}
	}

	// 2. hash table of the build side as a CSR multimap (count -> scan -> fill)
	uint64_t cap = 1024;
	while (cap < rt->n_slots * 2)
		cap <<= 1;
	long long *keys;
	uint32_t *cnt, *fill, *rows;
	uint64_t *off;
	MDB_TRY(tmp.alloc(&keys, cap));
	MDB_TRY(tmp.alloc(&cnt, cap + 2));
	MDB_TRY(tmp.alloc(&fill, cap + 2));
	MDB_TRY(tmp.alloc(&off, cap + 2));
	MDB_TRY(tmp.alloc(&rows, rt->n_slots));
	MDB_LAUNCH(ctx, k_fill_i64, grid_for(ctx, cap, 256), 256, 0, keys, cap, HT_EMPTY);
	CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, (cap + 2) * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(fill, 0, (cap + 2) * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(d_dup, 0, sizeof(*d_dup), ctx->stream));
	MDB_LAUNCH(ctx, k_join_build_count, gr, 256, 0, (const int64_t*)rc_.data, r_present, rt->n_slots, r_dbl, keys, cap - 1, cnt,
			d_dup);
	CUDA_CHECK_LAUNCH(ctx);
	MDB_TRY(mdb_scan_u32_u64(ctx, cnt, off, cap + 2, nullptr));
	MDB_LAUNCH(ctx, k_join_build_fill, gr, 256, 0, (const int64_t*)rc_.data, r_present, rt->n_slots, r_dbl,
			(const long long*)keys, cap - 1, (const uint64_t*)off, fill, rows);
	CUDA_CHECK_LAUNCH(ctx);
	if (!dup_known) {
		MDB_TRY(mdb_read_u64(ctx, (const uint64_t*)d_dup, &dup));
		lap("join: hash build (sync)", cap);
	}
	if (!dup) {
		lk.keys = keys;
		lk.cap_mask = cap - 1;
		lk.cnt = cnt;
		lk.row_off = off;
		lk.rows = rows;
		return probe_unique();
	}

	// 3. duplicates on the build side: count the matches of every tuple, scan, emit into new arrays
	uint32_t *matches;
	uint64_t *out_off, *first_row;
	MDB_TRY(tmp.alloc(&matches, ts->n));
	MDB_TRY(tmp.alloc(&first_row, ts->n));
	MDB_TRY(tmp.alloc(&out_off, ts->n));
	MDB_LAUNCH(ctx, k_join_probe_count, gp, 256, 0, to_dev(*ts), jn.left.tbl, (const int64_t*)lc.data, l_present, l_dbl,
			(const long long*)keys, cap - 1, (const uint32_t*)cnt, (const uint64_t*)off, matches, first_row);
	CUDA_CHECK_LAUNCH(ctx);
	MDB_TRY(mdb_scan_u32_u64(ctx, matches, out_off, ts->n, d_total));
	uint64_t total = 0;
	MDB_TRY(mdb_read_u64(ctx, d_total, &total));
	lap("join: count matches (sync)", total);

	MDB_TRY(alloc_out_tuples(ctx, &out, ts->ntab + 1, total, &arr));
	if (total) {
		MDB_LAUNCH(ctx, k_join_probe_emit, gp, 256, 0, to_dev(*ts), (const uint32_t*)matches, (const uint64_t*)first_row,
				(const uint64_t*)out_off, (const uint32_t*)rows, arr);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) {
			free_tuples(ctx, out);
			return mdb_fail(ctx, MDBCU_ECUDA, "k_join_probe_emit: %s", cudaGetErrorString(e));
		}
	}
	free_tuples(ctx, *ts);
	*ts = out;
	return MDBCU_OK;
}

// =========================================================================================== output / aggregation

struct DOut {
	int32_t kind;
	int32_t tbl;
	int32_t is_dbl;
	int32_t key_idx; // >= 0: this plain column equals group key `key_idx` (no first-row lookup needed)
	const int64_t *data;
	const uint32_t *present;
	long long *acc;           // per-slot accumulator (sum / min / max / count)
	unsigned long long *nn;   // per-slot count of non-NULL inputs (or COUNT(*) rows)
};

struct DGroupSpec {
	int32_t n_group;
	int32_t n_out;
	int32_t ntab;
	int32_t pack_ok;         // order keys are available
	int32_t dense;           // single INT group key with a narrow zone map: slot = key - dense_min, no hash table
	long long dense_min;
	int32_t pack_shift[MDBCU_MAX_TABLES];
	struct {
		int32_t tbl, is_dbl, mode; // mode 0: whole 64-bit key (single group column), 1: low 32 bits packed
		int32_t _pad;
		const int64_t *data;
		const uint32_t *present;
	} g[MDBCU_MAX_GROUP];
	DOut out[MDBCU_MAX_OUT];
};

template <typename Rows>
__device__ static inline unsigned long long pack_rids(const DGroupSpec *sp, const Rows &rows)
{
	unsigned long long k = 0;
	for (int t = 0; t < sp->ntab; t++)
		k |= (unsigned long long)rows(t) << sp->pack_shift[t];
	return k;
}

__device__ static inline unsigned long long pack_rids(const DGroupSpec *sp, const TuplesDev &ts, uint64_t i)
{
	TupleRows rows = {&ts, i};
	return pack_rids(sp, rows);
}

// slot layout: [0, cap) hashed keys, cap = key INT64_MIN, cap+1 = NULL group (NULLs collate equal, :1476-1482)
//
// Hot groups (a Zipf foreign key sends 12 % of the rows to one slot) would serialise on one L2 atomic unit, so each
// CTA keeps a direct-mapped cache of `cache_entries` groups in shared memory: [tag | first row | (acc, nn) per
// aggregate], E words each. A row whose slot owns (or can claim) its cache line updates shared memory; any other row
// goes to the global accumulators as before. The cache is merged into the global accumulators once per CTA.
// cache_entries == 0: no cache (more aggregates than 48 KiB of shared memory hold).
__device__ static inline long long group_acc_init(int kind)
{
	return kind == MDBCU_OUT_MIN ? INT64_MAX : kind == MDBCU_OUT_MAX ? INT64_MIN : 0;
}

__device__ static inline void group_acc_merge(const DOut &out, long long *acc, long long v)
{
	switch (out.kind) {
	case MDBCU_OUT_SUM: case MDBCU_OUT_AVG:
		if (out.is_dbl)
			atomicAdd((double*)acc, __longlong_as_double(v));
		else
			atomicAdd((unsigned long long*)acc, (unsigned long long)v);
		break;
	case MDBCU_OUT_MIN:
		atomicMin(acc, v);
		break;
	case MDBCU_OUT_MAX:
		atomicMax(acc, v);
		break;
	}
}

// Where the aggregate gets its joined tuples from.
// (1) The general operators: materialised tuple arrays, optionally with the WHERE verdicts as a bitmap.
struct TupleSource {
	typedef TupleRows Rows;
	TuplesDev ts;
	const uint32_t *keep; // bit i = tuple i qualifies (nullptr: all do; the filter was not materialised)
	__device__ uint64_t count() const { return ts.n; }
	__device__ bool fetch(uint64_t i, Rows &rows) const
	{
		if (keep && !mdb_bit(keep, i))
			return false;
		rows.ts = &ts;
		rows.i = i;
		return true;
	}
};

// (2) The fused multiway aggregate: the rows of tables[0] are scanned once; the partner in every further table is looked up
// in a direct row table (build keys duplicate-free with a narrow zone map, see k_join_build_direct), WHERE is evaluated on
// the row ids in registers. No tuple array is ever written.
struct StarDim {
	int32_t ltbl; // the join's left operand is a column of this (earlier) table
	int32_t _pad;
	const int64_t *ldata;
	const uint32_t *lpresent;
	const uint32_t *direct;
	long long dmin;
	uint64_t drange;
};

struct StarSource {
	typedef StarRows Rows;
	uint64_t n_fact;
	const uint32_t *live0; // nullptr: every slot of tables[0] is a live row
	int32_t n_dims;
	int32_t _pad;
	StarDim dim[MDBCU_MAX_TABLES - 1];
	int32_t has_prog; // 0: no WHERE
	int32_t _pad2;
	DPredProgram prog; // by value: the source is a kernel parameter, the program is read from the constant bank
	__device__ uint64_t count() const { return n_fact; }
	__device__ bool fetch(uint64_t i, Rows &rows) const
	{
		if (live0 && !mdb_bit(live0, i))
			return false;
		rows.r[0] = (uint32_t)i;
		for (int d = 0; d < n_dims; d++) {
			const StarDim &m = dim[d];
			uint32_t r = rows.r[m.ltbl];
			if (m.lpresent && !mdb_bit(m.lpresent, r))
				return false; // NULL keys never match
			uint64_t k = (uint64_t)m.ldata[r] - (uint64_t)m.dmin;
			if (k >= m.drange)
				return false;
			uint32_t partner = m.direct[k];
			if (partner == JOIN_NO_ROW)
				return false;
			rows.r[d + 1] = partner;
		}
		return !has_prog || eval_program(&prog, rows);
	}
};

template <typename Source>
__global__ void k_group_update(const __grid_constant__ DGroupSpec sp_, const __grid_constant__ Source src, long long *__restrict__ keys, uint64_t cap_mask,
		unsigned long long *__restrict__ first_key, uint32_t *__restrict__ used, uint32_t cache_entries)
{
	// the group / output descriptors are a kernel PARAMETER (constant bank): the ~40 descriptor reads per row are operands or
	// uniform loads instead of global loads that sit in front of the data loads they address (ncu: 45 LDG requests per warp-row)
	const DGroupSpec *sp = &sp_;
	extern __shared__ unsigned long long s_cache[];
	const uint32_t E = cache_entries;
	const unsigned long long NO_TAG = ~0ull;

	if (E) {
		for (uint32_t e = threadIdx.x; e < 2 * E; e += blockDim.x)
			s_cache[e] = NO_TAG; // tags, then first rows
		int a = 0;
		for (int o = 0; o < sp->n_out; o++) {
			if (sp->out[o].kind == MDBCU_OUT_COLUMN)
				continue;
			long long init = group_acc_init(sp->out[o].kind);
			for (uint32_t e = threadIdx.x; e < E; e += blockDim.x) {
				s_cache[(2 + 2 * a) * E + e] = (unsigned long long)init;
				s_cache[(3 + 2 * a) * E + e] = 0;
			}
			a++;
		}
		__syncthreads();
	}

	const uint64_t n_in = src.count();
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_in; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t slot;
		typename Source::Rows rows;
		if (!src.fetch(i, rows))
			continue;
		if (sp->n_group == 0) {
			slot = 0;
		} else {
			bool any_null = false;
			long long key = 0;
			for (int g = 0; g < sp->n_group; g++) {
				uint32_t r = rows(sp->g[g].tbl);
				bool isnull = sp->g[g].present && !mdb_bit(sp->g[g].present, r);
				long long v = isnull ? 0 : norm_key(sp->g[g].data[r], sp->g[g].is_dbl);
				if (sp->g[g].mode == 0) {
					key = v;
					any_null = isnull;
				} else {
					// two int32-range columns packed into one 64-bit key; NULL is encoded out of band below
					key = (long long)(((unsigned long long)key << 32) | ((unsigned long long)v & 0xffffffffull));
					if (isnull)
						any_null = true;
				}
			}
			if (any_null) {
				// the NULL group; composite keys containing NULLs are rejected on the host (build_group_spec)
				slot = cap_mask + 2;
			} else if (sp->dense) {
				slot = (uint64_t)key - (uint64_t)sp->dense_min;
			} else {
				slot = ht_insert(keys, cap_mask, key);
			}
		}

		// slots are hash positions already: their low bits index the cache
		bool hit = false;
		uint32_t e = (uint32_t)slot & (E - 1);
		if (E) {
			unsigned long long tag = s_cache[e];
			if (tag == NO_TAG) {
				tag = atomicCAS(&s_cache[e], NO_TAG, (unsigned long long)slot);
				if (tag == NO_TAG)
					tag = slot;
			}
			hit = tag == slot;
		}

		if (!hit)
			used[slot] = 1;
		if (sp->pack_ok)
			atomicMin(hit ? &s_cache[E + e] : &first_key[slot], pack_rids(sp, rows));
		int a = 0;
		for (int o = 0; o < sp->n_out; o++) {
			const DOut &out = sp->out[o];
			if (out.kind == MDBCU_OUT_COLUMN)
				continue;
			unsigned long long *nn = hit ? &s_cache[(3 + 2 * a) * E + e] : &out.nn[slot];
			long long *acc = hit ? (long long*)&s_cache[(2 + 2 * a) * E + e] : &out.acc[slot];
			a++;
			if (out.kind == MDBCU_OUT_COUNT_STAR) {
				atomicAdd(nn, 1ull);
				continue;
			}
			uint32_t r = rows(out.tbl);
			if (out.present && !mdb_bit(out.present, r))
				continue;
			long long v = out.data[r];
			atomicAdd(nn, 1ull);
			if (out.kind == MDBCU_OUT_MIN || out.kind == MDBCU_OUT_MAX)
				v = out.is_dbl ? mdb_dbl_to_ordered(v) : v;
			group_acc_merge(out, acc, v);
		}
	}

	if (!E)
		return;
	__syncthreads();
	for (uint32_t e = threadIdx.x; e < E; e += blockDim.x) {
		unsigned long long slot = s_cache[e];
		if (slot == NO_TAG)
			continue;
		used[slot] = 1;
		if (sp->pack_ok)
			atomicMin(&first_key[slot], s_cache[E + e]);
		int a = 0;
		for (int o = 0; o < sp->n_out; o++) {
			const DOut &out = sp->out[o];
			if (out.kind == MDBCU_OUT_COLUMN)
				continue;
			long long acc = (long long)s_cache[(2 + 2 * a) * E + e];
			unsigned long long nn = s_cache[(3 + 2 * a) * E + e];
			a++;
			if (nn == 0)
				continue;
			atomicAdd(&out.nn[slot], nn);
			if (out.acc)
				group_acc_merge(out, &out.acc[slot], acc);
		}
	}
}

__global__ void k_flags_to_bits(const uint32_t *__restrict__ used, uint64_t n, uint32_t *__restrict__ bits)
{
	uint64_t groups = (n + 31) / 32;
	uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	int lane = threadIdx.x & 31;
	for (uint64_t g = warp; g < groups; g += nwarps) {
		uint64_t i = g * 32 + lane;
		uint32_t b = __ballot_sync(0xffffffffu, i < n && used[i]);
		if (lane == 0)
			bits[g] = b;
	}
}

struct DResultCols {
	int64_t *cells[MDBCU_MAX_OUT];
	uint8_t *nulls[MDBCU_MAX_OUT];
};

__device__ static inline void unpack_rids(const DGroupSpec *sp, unsigned long long k, uint32_t *rid)
{
	for (int t = 0; t < sp->ntab; t++) {
		int hi = t == 0 ? 64 : sp->pack_shift[t - 1];
		int width = hi - sp->pack_shift[t];
		unsigned long long m = width >= 64 ? ~0ull : ((1ull << width) - 1ull);
		rid[t] = (uint32_t)((k >> sp->pack_shift[t]) & m);
	}
}

__global__ void k_group_emit(const DGroupSpec *__restrict__ sp, const uint32_t *__restrict__ slots, uint64_t ngroups,
		const long long *__restrict__ keys, uint64_t cap_mask, const unsigned long long *__restrict__ first_key,
		DResultCols res, unsigned long long *__restrict__ order_out)
{
	for (uint64_t gidx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; gidx < ngroups; gidx += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t slot = slots[gidx];
		uint32_t rid[MDBCU_MAX_TABLES] = {0, 0, 0, 0};
		if (sp->pack_ok) {
			unpack_rids(sp, first_key[slot], rid);
			if (order_out)
				order_out[gidx] = first_key[slot];
		}
		bool key_null = sp->n_group > 0 && slot == cap_mask + 2;
		long long key = 0;
		if (sp->n_group > 0 && !key_null)
			key = slot == cap_mask + 1 ? HT_EMPTY : sp->dense ? (long long)(slot + (uint64_t)sp->dense_min) : keys[slot];
		for (int o = 0; o < sp->n_out; o++) {
			const DOut &out = sp->out[o];
			long long cell = 0;
			bool isnull = false;
			unsigned long long nn = out.nn ? out.nn[slot] : 0;
			switch (out.kind) {
			case MDBCU_OUT_COLUMN:
				if (out.key_idx >= 0 && sp->g[out.key_idx].mode == 0) {
					cell = key;
					isnull = key_null;
				} else {
					uint32_t r = rid[out.tbl];
					isnull = out.present && !mdb_bit(out.present, r);
					cell = isnull ? 0 : out.data[r];
				}
				break;
			case MDBCU_OUT_COUNT_STAR: case MDBCU_OUT_COUNT_COL:
				cell = (long long)nn;
				break;
			case MDBCU_OUT_SUM:
				isnull = nn == 0;
				cell = out.acc[slot];
				break;
			case MDBCU_OUT_MIN: case MDBCU_OUT_MAX:
				isnull = nn == 0;
				cell = out.is_dbl ? mdb_ordered_to_dbl(out.acc[slot]) : out.acc[slot];
				break;
			case MDBCU_OUT_AVG:
				isnull = nn == 0;
				if (!isnull) {
					double s = out.is_dbl ? __longlong_as_double(out.acc[slot]) : (double)out.acc[slot];
					cell = __double_as_longlong(s / (double)nn);
				}
				break;
			}
			res.cells[o][gidx] = isnull ? 0 : cell;
			res.nulls[o][gidx] = isnull;
		}
	}
}

// projection: result cell (o, i) = column o of tuple i's row.  The output descriptors are decoded from shared memory and
// every thread has four tuples in flight (their gathers are independent): 2^27 cells in 0.69 ms before, one tuple per thread
// with the descriptors read from global memory per cell.
#define GO_ROWS 4
__global__ void __launch_bounds__(256) k_gather_out(const DGroupSpec *__restrict__ sp, TuplesDev ts, DResultCols res,
		unsigned long long *__restrict__ order_out)
{
	__shared__ DOut s_out[MDBCU_MAX_OUT];
	const int n_out = sp->n_out;
	for (uint32_t i = threadIdx.x; i < n_out * sizeof(DOut) / 4; i += blockDim.x)
		reinterpret_cast<uint32_t*>(s_out)[i] = reinterpret_cast<const uint32_t*>(sp->out)[i];
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < ts.n; i0 += GO_ROWS * stride) {
		for (int o = 0; o < n_out; o++) {
			const DOut &out = s_out[o];
			const uint32_t *c = ts.rid[out.tbl];
			uint32_t r[GO_ROWS];
			long long cell[GO_ROWS];
			bool isnull[GO_ROWS];
#pragma unroll
			for (int j = 0; j < GO_ROWS; j++) {
				const uint64_t i = i0 + j * stride;
				r[j] = i < ts.n ? (c ? c[i] : (uint32_t)i) : 0u; // (identity tuples: scan_live)
			}
#pragma unroll
			for (int j = 0; j < GO_ROWS; j++) {
				const bool in = i0 + j * stride < ts.n;
				isnull[j] = in && out.present && !mdb_bit(out.present, r[j]);
				cell[j] = in && !isnull[j] ? out.data[r[j]] : 0;
			}
#pragma unroll
			for (int j = 0; j < GO_ROWS; j++) {
				const uint64_t i = i0 + j * stride;
				if (i < ts.n) {
					res.cells[o][i] = cell[j];
					res.nulls[o][i] = isnull[j];
				}
			}
		}
		if (order_out) {
#pragma unroll
			for (int j = 0; j < GO_ROWS; j++) {
				const uint64_t i = i0 + j * stride;
				if (i < ts.n)
					order_out[i] = pack_rids(sp, ts, i);
			}
		}
	}
}

int out_result_type(const mdbcu_plan *plan, int o)
{
	const mdbcu_out &out = plan->out[o];
	if (out.kind == MDBCU_OUT_COUNT_STAR || out.kind == MDBCU_OUT_COUNT_COL)
		return MDBCU_CT_INTEGER;
	if (out.kind == MDBCU_OUT_AVG)
		return MDBCU_CT_DOUBLE;
	int type = plan->tables[out.ref.tbl]->cols[out.ref.col].type;
	if (out.kind == MDBCU_OUT_COLUMN)
		return type;
	return type == MDBCU_CT_DOUBLE ? MDBCU_CT_DOUBLE : MDBCU_CT_INTEGER;
}

int mdb_result_alloc(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res, uint64_t nrows, bool with_order)
{
	res->nrows = nrows;
	res->cols.resize(plan->n_out);
	for (int o = 0; o < plan->n_out; o++) {
		res->cols[o].type = out_result_type(plan, o);
		MDB_TRY(mdb_alloc(ctx, &res->cols[o].cells, nrows));
		MDB_TRY(mdb_alloc(ctx, &res->cols[o].nulls, nrows));
	}
	if (with_order)
		MDB_TRY(mdb_alloc(ctx, &res->order_key, nrows));
	return MDBCU_OK;
}

static bool colref_eq(const mdbcu_colref &a, const mdbcu_colref &b)
{
	return a.tbl == b.tbl && a.col == b.col;
}

// fills everything of the spec except accumulator pointers
static int build_group_spec(mdbcu_ctx *ctx, const mdbcu_plan *plan, int ntab, DGroupSpec *sp)
{
	memset(sp, 0, sizeof(*sp));
	sp->n_group = plan->n_group;
	sp->n_out = plan->n_out;
	sp->ntab = ntab;

	// order keys: row ids of all joined tables packed left-major = the reference's nested-loop row order
	int bits_total = 0, bits[MDBCU_MAX_TABLES];
	for (int t = 0; t < ntab; t++) {
		uint64_t n = std::max<uint64_t>(plan->tables[t]->n_slots, 2);
		bits[t] = 1;
		while ((1ull << bits[t]) < n)
			bits[t]++;
		bits_total += bits[t];
	}
	sp->pack_ok = bits_total <= 64;
	if (sp->pack_ok) {
		int sh = 64;
		for (int t = 0; t < ntab; t++) {
			sh -= bits[t];
			sp->pack_shift[t] = sh;
		}
		// put the last table at bit 0 so single-table keys are plain row ids
		int slack = sp->pack_shift[ntab - 1];
		for (int t = 0; t < ntab; t++)
			sp->pack_shift[t] -= slack;
	}

	if (plan->n_group < 0 || plan->n_group > MDBCU_MAX_GROUP)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "GROUP BY supports at most %d columns", MDBCU_MAX_GROUP);
	for (int g = 0; g < plan->n_group; g++) {
		MDB_TRY(check_colref(ctx, plan, plan->group[g].tbl, plan->group[g].col, "GROUP BY"));
		const mdbcu_table *t = plan->tables[plan->group[g].tbl];
		const DevColumn &c = t->cols[plan->group[g].col];
		sp->g[g].tbl = plan->group[g].tbl;
		sp->g[g].is_dbl = c.type == MDBCU_CT_DOUBLE;
		sp->g[g].data = c.data;
		sp->g[g].present = col_all_present(t, plan->group[g].col) ? nullptr : c.present;
		sp->g[g].mode = plan->n_group == 1 ? 0 : 1;
		if (plan->n_group > 1) {
			// composite keys are packed 2 x 32 bits: the reference's own integer domain is int32
			if (c.type == MDBCU_CT_DOUBLE || !c.stats_ok || (c.imin <= c.imax && (c.imin < INT32_MIN || c.imax > INT32_MAX)))
				return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "composite GROUP BY needs INT columns within int32 range");
			if (c.has_nulls)
				return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "composite GROUP BY over columns containing NULLs");
		}
	}

	for (int o = 0; o < plan->n_out; o++) {
		const mdbcu_out &po = plan->out[o];
		DOut &d = sp->out[o];
		d.kind = po.kind;
		d.key_idx = -1;
		if (po.kind < MDBCU_OUT_COLUMN || po.kind > MDBCU_OUT_AVG)
			return mdb_fail(ctx, MDBCU_EERROR, "unknown output kind %d", po.kind);
		if (po.kind == MDBCU_OUT_COUNT_STAR)
			continue;
		MDB_TRY(check_colref(ctx, plan, po.ref.tbl, po.ref.col, "SELECT list"));
		const mdbcu_table *t = plan->tables[po.ref.tbl];
		const DevColumn &c = t->cols[po.ref.col];
		d.tbl = po.ref.tbl;
		d.is_dbl = c.type == MDBCU_CT_DOUBLE;
		d.data = c.data;
		d.present = col_all_present(t, po.ref.col) ? nullptr : c.present;
		if (po.kind == MDBCU_OUT_COLUMN) {
			for (int g = 0; g < plan->n_group; g++)
				if (colref_eq(po.ref, plan->group[g]))
					d.key_idx = g;
		}
	}
	return MDBCU_OK;
}

__global__ void k_fill_u64(unsigned long long *p, uint64_t n, unsigned long long v)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		p[i] = v;
}

static bool plan_has_aggregate(const mdbcu_plan *plan)
{
	for (int o = 0; o < plan->n_out; o++)
		if (plan->out[o].kind != MDBCU_OUT_COLUMN)
			return true;
	return false;
}

// keep: optional WHERE verdict bitmap over the tuples (bit i = tuple i qualifies)
// star: the joined tuples are not materialised at all but produced on the fly (StarSource); ts is ignored then
static int aggregate_tuples(mdbcu_ctx *ctx, const mdbcu_plan *plan, const Tuples &ts, const uint32_t *keep, const StarSource *star,
		mdbcu_result *res)
{
	const uint64_t n_in = star ? star->n_fact : ts.n;
	DGroupSpec sp;
	MDB_TRY(build_group_spec(ctx, plan, star ? star->n_dims + 1 : ts.ntab, &sp));

	bool need_first = false;
	for (int o = 0; o < plan->n_out; o++)
		if (sp.out[o].kind == MDBCU_OUT_COLUMN && !(sp.out[o].key_idx >= 0 && sp.g[sp.out[o].key_idx].mode == 0))
			need_first = true;
	if (need_first && !sp.pack_ok)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "plain columns under GROUP BY need row ids that pack into 64 bits");
	// The first row of every group (one 64-bit atomicMin per input row) serves two purposes: plain columns under GROUP BY, and
	// the reference's row order for results small enough to be ordered at all.  Above 2^24 input rows - sizes the reference
	// itself could never run - the order is not kept, so without plain columns the atomic is dropped.
	if (!need_first && n_in > (1ull << 24))
		sp.pack_ok = 0;

	if (n_in == 0)
		return mdb_result_alloc(ctx, plan, res, 0, false); // no qualifying row: no result row (executor keeps zero rows)

	// a single INT key whose zone map is narrow is its own slot number (tables small enough to stay in L2, nothing to probe)
	if (plan->n_group == 1 && sp.g[0].mode == 0 && !sp.g[0].is_dbl) {
		const DevColumn &gc = plan->tables[plan->group[0].tbl]->cols[plan->group[0].col];
		if (gc.stats_ok && gc.imin <= gc.imax && gc.imin > INT64_MIN) {
			unsigned long long range = (unsigned long long)gc.imax - (unsigned long long)gc.imin + 1ull;
			if (range != 0 && range <= std::max<unsigned long long>(1ull << 16, 4ull * n_in)) {
				sp.dense = 1;
				sp.dense_min = gc.imin;
			}
		}
	}

	uint64_t cap = 1024;
	if (sp.dense) {
		const DevColumn &gc = plan->tables[plan->group[0].tbl]->cols[plan->group[0].col];
		while (cap < (unsigned long long)gc.imax - (unsigned long long)gc.imin + 1ull)
			cap <<= 1;
	} else if (plan->n_group > 0) {
		while (cap < n_in * 2)
			cap <<= 1;
	}
	uint64_t nslots = cap + 3;

	HostLap lap;
	DevTemp tmp(ctx, true);
	long long *keys;
	unsigned long long *first_key;
	uint32_t *used, *bits;
	DGroupSpec *d_sp;
	MDB_TRY(tmp.alloc(&keys, sp.dense ? 1 : cap));
	MDB_TRY(tmp.alloc(&first_key, nslots));
	MDB_TRY(tmp.alloc(&used, nslots));
	MDB_TRY(tmp.alloc(&bits, (nslots + 31) / 32));
	MDB_TRY(tmp.alloc(&d_sp, 1));
	if (!sp.dense)
		MDB_LAUNCH(ctx, k_fill_i64, grid_for(ctx, cap, 256), 256, 0, keys, cap, HT_EMPTY);
	MDB_LAUNCH(ctx, k_fill_u64, grid_for(ctx, nslots, 256), 256, 0, first_key, nslots, ~0ull);
	CUDA_TRY(ctx, cudaMemsetAsync(used, 0, nslots * sizeof(uint32_t), ctx->stream));

	for (int o = 0; o < plan->n_out; o++) {
		DOut &d = sp.out[o];
		if (d.kind == MDBCU_OUT_COLUMN)
			continue;
		MDB_TRY(tmp.alloc(&d.nn, nslots));
		CUDA_TRY(ctx, cudaMemsetAsync(d.nn, 0, nslots * sizeof(unsigned long long), ctx->stream));
		if (d.kind == MDBCU_OUT_COUNT_STAR || d.kind == MDBCU_OUT_COUNT_COL)
			continue;
		MDB_TRY(tmp.alloc(&d.acc, nslots));
		long long init = 0;
		if (d.kind == MDBCU_OUT_MIN)
			init = INT64_MAX;
		else if (d.kind == MDBCU_OUT_MAX)
			init = INT64_MIN;
		MDB_LAUNCH(ctx, k_fill_i64, grid_for(ctx, nslots, 256), 256, 0, d.acc, nslots, init);
	}
	CUDA_TRY(ctx, cudaMemcpyAsync(d_sp, &sp, sizeof(sp), cudaMemcpyHostToDevice, ctx->stream));
	lap("aggregate: allocate + init", nslots);

	// per-CTA cache of hot groups: the largest power of two of entries that fits 48 KiB (none if even 256 do not)
	int n_agg = 0;
	for (int o = 0; o < plan->n_out; o++)
		n_agg += sp.out[o].kind != MDBCU_OUT_COLUMN;
	const size_t entry_bytes = 8 * (2 + 2 * (size_t)n_agg);
	// (MDBCU_GROUP_CACHE_KB: A/B switch.  The kernel needs 32 registers, so shared memory decides how many CTAs an SM holds -
	// 48 KiB: 1024 threads, 24 KiB: 2048 - but config 4 runs 2.82 / 2.84 / 2.94 ms with 48 / 24 / 12 KiB: occupancy is not
	// what limits it, the hit rate of the cache matters a little)
	static const size_t cache_kb = getenv("MDBCU_GROUP_CACHE_KB") ? (size_t)atoi(getenv("MDBCU_GROUP_CACHE_KB")) : 48;
	uint32_t cache_entries = 2048;
	while (cache_entries >= 256 && cache_entries * entry_bytes > cache_kb * 1024)
		cache_entries >>= 1;
	if (cache_entries < 256 || getenv("MDBCU_NO_GROUP_CACHE")) // the switch is for A/B measurements
		cache_entries = 0;
	if (star) {
		MDB_LAUNCH(ctx, k_group_update<StarSource>, grid_for(ctx, n_in, 256), 256, cache_entries * entry_bytes,
				sp, *star, keys, cap - 1, first_key, used, cache_entries);
	} else {
		TupleSource src = {to_dev(ts), keep};
		MDB_LAUNCH(ctx, k_group_update<TupleSource>, grid_for(ctx, n_in, 256), 256, cache_entries * entry_bytes,
				sp, src, keys, cap - 1, first_key, used, cache_entries);
	}
	CUDA_CHECK_LAUNCH(ctx);
	MDB_LAUNCH(ctx, k_flags_to_bits, grid_for(ctx, nslots, 256), 256, 0, (const uint32_t*)used, nslots, bits);
	CUDA_CHECK_LAUNCH(ctx);

	Tuples slots;
	lap("aggregate: launches", n_in);
	MDB_TRY(compact_tuples(ctx, bits, nslots, nullptr, 1, &slots));
	lap("aggregate: compact (sync)", slots.n);
	int rc = mdb_result_alloc(ctx, plan, res, slots.n, sp.pack_ok != 0);
	lap("aggregate: result alloc", slots.n);
	if (rc == MDBCU_OK && slots.n) {
		DResultCols rcols;
		memset(&rcols, 0, sizeof(rcols));
		for (int o = 0; o < plan->n_out; o++) {
			rcols.cells[o] = res->cols[o].cells;
			rcols.nulls[o] = res->cols[o].nulls;
		}
		MDB_LAUNCH(ctx, k_group_emit, grid_for(ctx, slots.n, 256), 256, 0, (const DGroupSpec*)d_sp,
				(const uint32_t*)slots.rid[0], slots.n, (const long long*)keys, cap - 1,
				(const unsigned long long*)first_key, rcols, (unsigned long long*)res->order_key);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			rc = mdb_fail(ctx, MDBCU_ECUDA, "k_group_emit: %s", cudaGetErrorString(e));
	}
	free_tuples(ctx, slots);
	return rc;
}

static int project_tuples(mdbcu_ctx *ctx, const mdbcu_plan *plan, const Tuples &ts, mdbcu_result *res)
{
	DGroupSpec sp;
	MDB_TRY(build_group_spec(ctx, plan, ts.ntab, &sp));
	MDB_TRY(mdb_result_alloc(ctx, plan, res, ts.n, sp.pack_ok != 0));
	if (ts.n == 0)
		return MDBCU_OK;
	DevTemp tmp(ctx, true);
	DGroupSpec *d_sp;
	MDB_TRY(tmp.alloc(&d_sp, 1));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_sp, &sp, sizeof(sp), cudaMemcpyHostToDevice, ctx->stream));
	DResultCols rcols;
	memset(&rcols, 0, sizeof(rcols));
	for (int o = 0; o < plan->n_out; o++) {
		rcols.cells[o] = res->cols[o].cells;
		rcols.nulls[o] = res->cols[o].nulls;
	}
	MDB_LAUNCH(ctx, k_gather_out, grid_for(ctx, ts.n, 256), 256, 0, (const DGroupSpec*)d_sp, to_dev(ts), rcols,
			(unsigned long long*)res->order_key);
	CUDA_CHECK_LAUNCH(ctx);
	return MDBCU_OK;
}

// =========================================================================================== fused multiway aggregate

// Joins whose build sides are duplicate-free INT keys with a narrow zone map need no tuple arrays: one direct row table per
// joined table, then ONE kernel scans tables[0], looks the partners up, evaluates WHERE and aggregates (StarSource).
// *ok = false (and MDBCU_OK): the plan does not have that shape; the general operators take it.
static int star_source(mdbcu_ctx *ctx, const mdbcu_plan *plan, DevTemp &tmp, StarSource *src, bool *ok)
{
	*ok = false;
	memset(src, 0, sizeof(*src));
	const mdbcu_table *fact = plan->tables[0];
	if (plan->n_joins < 1 || plan->n_tables != plan->n_joins + 1 || fact->n_slots == 0)
		return MDBCU_OK;
	for (int j = 0; j < plan->n_joins; j++) {
		const mdbcu_join &jn = plan->joins[j];
		if (jn.cross || jn.right.tbl != j + 1 || jn.left.tbl > j || jn.left.tbl < 0)
			return MDBCU_OK;
		MDB_TRY(check_colref(ctx, plan, jn.left.tbl, jn.left.col, "JOIN"));
		MDB_TRY(check_colref(ctx, plan, jn.right.tbl, jn.right.col, "JOIN"));
		const mdbcu_table *rt = plan->tables[j + 1];
		const DevColumn &lc = plan->tables[jn.left.tbl]->cols[jn.left.col], &rc = rt->cols[jn.right.col];
		if (lc.type == MDBCU_CT_DOUBLE || rc.type == MDBCU_CT_DOUBLE || !rc.stats_ok || rc.imin > rc.imax || rt->n_slots == 0)
			return MDBCU_OK;
		const unsigned long long range = (unsigned long long)rc.imax - (unsigned long long)rc.imin + 1ull;
		if (range == 0 || range > std::max<unsigned long long>(1ull << 16, 4ull * rt->n_slots))
			return MDBCU_OK;
	}

	unsigned long long *d_dup;
	MDB_TRY(tmp.alloc(&d_dup, 1));
	CUDA_TRY(ctx, cudaMemsetAsync(d_dup, 0, sizeof(*d_dup), ctx->stream));
	src->n_fact = fact->n_slots;
	src->live0 = fact->all_live ? nullptr : fact->live;
	src->n_dims = plan->n_joins;
	for (int j = 0; j < plan->n_joins; j++) {
		const mdbcu_join &jn = plan->joins[j];
		const mdbcu_table *lt = plan->tables[jn.left.tbl], *rt = plan->tables[j + 1];
		const DevColumn &lc = lt->cols[jn.left.col], &rc = rt->cols[jn.right.col];
		const uint64_t range = (uint64_t)rc.imax - (uint64_t)rc.imin + 1ull;
		uint32_t *direct;
		MDB_TRY(tmp.alloc(&direct, range));
		CUDA_TRY(ctx, cudaMemsetAsync(direct, 0xff, range * sizeof(uint32_t), ctx->stream));
		MDB_LAUNCH(ctx, k_join_build_direct, grid_for(ctx, rt->n_slots, 256), 256, 0, (const int64_t*)rc.data,
				col_all_present(rt, jn.right.col) ? (const uint32_t*)nullptr : (const uint32_t*)rc.present, rt->n_slots,
				(long long)rc.imin, range, direct, d_dup);
		CUDA_CHECK_LAUNCH(ctx);
		StarDim &m = src->dim[j];
		m.ltbl = jn.left.tbl;
		m.ldata = lc.data;
		m.lpresent = col_all_present(lt, jn.left.col) ? nullptr : lc.present;
		m.direct = direct;
		m.dmin = rc.imin;
		m.drange = range;
	}
	uint64_t dup = 1;
	MDB_TRY(mdb_read_u64(ctx, (const uint64_t*)d_dup, &dup));
	if (dup)
		return MDBCU_OK; // some build key occurs twice: tuples multiply, the general join handles that

	if (plan->n_pred > 0) {
		MDB_TRY(build_pred(ctx, plan, &src->prog));
		src->has_prog = 1;
	}
	*ok = true;
	return MDBCU_OK;
}

// =========================================================================================== driver

// the fused multiway aggregate (scan of tables[0], direct row tables for the joined tables, WHERE and grouped aggregates
// in one kernel) is the default for the plans it fits; MDBCU_FUSED_MULTIWAY=0 sends them to the tuple operators instead
static bool mdb_fused_multiway_enabled()
{
	const char *e = getenv("MDBCU_FUSED_MULTIWAY");
	return !e || atoi(e) != 0;
}

int mdb_select_general(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	Tuples ts;
	PhaseClock clock(ctx);
	int rc;

	HostLap lap; // MDBCU_TRACE=2: host wall time of each operator (includes the stream synchronisations inside it)

	const bool aggregates = plan->n_group > 0 || plan_has_aggregate(plan);
	if (aggregates && plan->n_joins > 0 && mdb_fused_multiway_enabled()) {
		// opt-in until the whole GPU suite has run with it
		DevTemp star_tmp(ctx, true);
		StarSource star;
		bool ok = false;
		clock.begin(2);
		rc = star_source(ctx, plan, star_tmp, &star, &ok);
		lap("fused: direct tables (sync)", plan->n_joins);
		if (rc != MDBCU_OK || ok) {
			if (rc == MDBCU_OK) {
				ctx->stats.path = MDBCU_PATH_FUSED_MULTIWAY;
				clock.begin(4);
				rc = aggregate_tuples(ctx, plan, ts, nullptr, &star, res);
				lap("fused: aggregate", star.n_fact);
			}
			clock.finish();
			return rc;
		}
	}

	ctx->stats.path = MDBCU_PATH_GENERAL;
	clock.begin(0);
	rc = scan_live(ctx, plan->tables[0], &ts, plan->n_joins == 0);
	lap("general: scan", ts.n);
	for (int j = 0; rc == MDBCU_OK && j < plan->n_joins; j++) {
		clock.begin(j == 0 ? 2 : 3);
		rc = join_step(ctx, plan, j, &ts);
		lap("general: join", ts.n);
	}
	DevTemp verdicts(ctx, true);
	uint32_t *keep = nullptr;
	if (rc == MDBCU_OK) {
		clock.begin(0);
		// an aggregate consumes the verdict bitmap directly; a projection needs the surviving tuples materialised
		rc = aggregates ? eval_pred_bits(ctx, plan, ts, verdicts, &keep) : filter_tuples(ctx, plan, &ts);
		lap("general: filter", ts.n);
	}
	if (rc == MDBCU_OK) {
		if (aggregates) {
			clock.begin(4);
			rc = aggregate_tuples(ctx, plan, ts, keep, nullptr, res);
			lap("general: aggregate", ts.n);
		} else {
			clock.begin(5);
			rc = project_tuples(ctx, plan, ts, res);
			lap("general: project", ts.n);
		}
	}
	free_tuples(ctx, ts);
	clock.finish();
	lap("general: finish", ts.n);
	return rc;
}
