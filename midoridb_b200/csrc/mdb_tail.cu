// mdb_tail.cu - tail operators on a result set that is already on the device: HAVING, DISTINCT, ORDER BY, LIMIT.
//
// The reference parses and validates all four (midorisql.y:180-196,203; semantic_select.c:1895,2004) and executes none
// (executor_select.c:1723 `TODO process distinct`), so there is no reference behaviour to reproduce: the
// semantics are SQL's as sqlite3 implements them, which is also what the oracle (oracle/mdb_oracle.c) is anchored on.
//
// The result is columnar (8-byte cells + one NULL byte per cell).  Everything here works on a PERMUTATION of row numbers
// and touches the cells once at the end:
//   HAVING    k_having_eval: a postfix program per row (the WHERE interpreter's operators, operands = result columns),
//             then a stable compaction of the permutation;
//   ORDER BY  least-significant-column-first stable sort of the permutation.  One column = one gather of its cells
//             into order-preserving unsigned keys, then one stable 1-bit split per key bit that actually varies
//             (an OR / AND reduction finds them: a COUNT column of small numbers costs a handful of passes, not 64),
//             then one more split on the NULL flag (NULLs first ascending, last descending);
//   DISTINCT  the same sort over ALL columns, then "differs from its predecessor" flags and a compaction;
//   LIMIT     a slice of the permutation;
//   finally   k_gather_rows writes the surviving rows in their final order.
// HBM-bound streaming passes (12 bytes per row and pass); results are at most a few 10^8 rows and usually tiny.
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>

#define TL_THREADS 256
#define TL_PER_THREAD 8
#define TL_TILE (TL_THREADS * TL_PER_THREAD)

struct TLCol {
	const int64_t *cells;
	const uint8_t *nulls; // may be nullptr
	int is_dbl;
};

struct TLCols {
	int n;
	TLCol c[MDBCU_MAX_OUT];
};

// ------------------------------------------------------------------ HAVING

struct TLVal {
	int kind; // 0 int, 1 double, 2 NULL, 3 bool
	long long i;
};

__device__ static double tl_as_dbl(const TLVal &v)
{
	return v.kind == 1 ? __longlong_as_double(v.i) : (double)v.i;
}

// comparison with a NULL operand is false (the WHERE rule: executor_select.c:557-579,629-631)
__device__ static bool tl_cmp(int cmp, const TLVal &a, const TLVal &b)
{
	if (a.kind == 2 || b.kind == 2)
		return false;
	int r;
	if (a.kind == 1 || b.kind == 1) {
		const double x = tl_as_dbl(a), y = tl_as_dbl(b);
		r = x < y ? -1 : (x > y ? 1 : 0);
	} else {
		r = a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
	}
	switch (cmp) {
	case 1: return r < 0;
	case 2: return r > 0;
	case 3: return r != 0;
	case 4: return r == 0;
	case 5: return r <= 0;
	case 6: return r >= 0;
	}
	return false;
}

struct TLProgram {
	int n;
	mdbcu_pred_op ops[MDBCU_MAX_HAVING];
};

__global__ void k_having_eval(TLProgram prog, TLCols cols, uint64_t n, uint32_t *__restrict__ keep)
{
	for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < n; row += (uint64_t)gridDim.x * blockDim.x) {
		TLVal st[MDBCU_MAX_HAVING];
		int sp = 0;
		for (int k = 0; k < prog.n; k++) {
			const mdbcu_pred_op &op = prog.ops[k];
			TLVal v;
			v.kind = 3;
			v.i = 0;
			switch (op.op) {
			case MDBCU_P_OUT: {
				const TLCol &c = cols.c[op.col];
				if (c.nulls && c.nulls[row]) {
					v.kind = 2;
				} else {
					v.kind = c.is_dbl ? 1 : 0;
					v.i = c.cells[row];
				}
				st[sp++] = v;
				break;
			}
			case MDBCU_P_INT: v.kind = 0; v.i = op.ival; st[sp++] = v; break;
			case MDBCU_P_DBL: v.kind = 1; v.i = __double_as_longlong(op.dval); st[sp++] = v; break;
			case MDBCU_P_NULL: v.kind = 2; st[sp++] = v; break;
			case MDBCU_P_BOOL: v.i = op.ival != 0; st[sp++] = v; break;
			case MDBCU_P_CMP: {
				const TLVal b = st[--sp], a = st[--sp];
				v.i = tl_cmp(op.arg, a, b);
				st[sp++] = v;
				break;
			}
			case MDBCU_P_AND: case MDBCU_P_OR: case MDBCU_P_XOR: {
				const TLVal b = st[--sp], a = st[--sp];
				const bool x = a.i != 0, y = b.i != 0;
				v.i = op.op == MDBCU_P_AND ? (x && y) : (op.op == MDBCU_P_OR ? (x || y) : (x != y));
				st[sp++] = v;
				break;
			}
			case MDBCU_P_ISNULL: case MDBCU_P_ISNOTNULL: {
				const TLVal a = st[--sp];
				v.i = (a.kind == 2) != (op.op == MDBCU_P_ISNOTNULL);
				st[sp++] = v;
				break;
			}
			case MDBCU_P_IN: case MDBCU_P_NOTIN: {
				const int cnt = op.arg;
				const TLVal probe = st[sp - cnt - 1];
				bool any = false, all_diff = true;
				for (int j = 0; j < cnt; j++) {
					any = any || tl_cmp(4, probe, st[sp - cnt + j]);
					all_diff = all_diff && tl_cmp(3, probe, st[sp - cnt + j]);
				}
				sp -= cnt + 1;
				v.i = op.op == MDBCU_P_IN ? any : all_diff;
				st[sp++] = v;
				break;
			}
			}
		}
		keep[row] = (sp == 1 && st[0].i != 0) ? 1u : 0u;
	}
}

// host-side check of a HAVING program: operators known, operands inside the result, stack never under- or overflows
static bool tl_program_ok(const mdbcu_plan *plan)
{
	int sp = 0;
	for (int k = 0; k < plan->n_having; k++) {
		const mdbcu_pred_op &op = plan->having[k];
		switch (op.op) {
		case MDBCU_P_OUT:
			if (op.col < 0 || op.col >= plan->n_out)
				return false;
			sp++;
			break;
		case MDBCU_P_INT: case MDBCU_P_DBL: case MDBCU_P_NULL: case MDBCU_P_BOOL:
			sp++;
			break;
		case MDBCU_P_CMP:
			if (op.arg < 1 || op.arg > 6 || sp < 2)
				return false;
			sp--;
			break;
		case MDBCU_P_AND: case MDBCU_P_OR: case MDBCU_P_XOR:
			if (sp < 2)
				return false;
			sp--;
			break;
		case MDBCU_P_ISNULL: case MDBCU_P_ISNOTNULL:
			if (sp < 1)
				return false;
			break;
		case MDBCU_P_IN: case MDBCU_P_NOTIN:
			if (op.arg < 1 || sp < op.arg + 1)
				return false;
			sp -= op.arg;
			break;
		default:
			return false;
		}
		if (sp > MDBCU_MAX_HAVING)
			return false;
	}
	return sp == 1;
}

// ------------------------------------------------------------------ stable split / compaction of (key, row) pairs

// zeros of one tile: flag(i) = 0 means "goes to the front"
template <typename FlagOf>
__device__ __forceinline__ uint32_t tl_tile_zeros(uint64_t base, uint64_t n, FlagOf flag_of, uint32_t *mine /* [TL_PER_THREAD] flags */)
{
	uint32_t zeros = 0;
#pragma unroll
	for (int j = 0; j < TL_PER_THREAD; j++) {
		const uint64_t i = base + (uint64_t)threadIdx.x * TL_PER_THREAD + j;
		mine[j] = i < n ? flag_of(i) : 1u;
		zeros += (i < n && mine[j] == 0) ? 1u : 0u;
	}
	return zeros;
}

// exclusive prefix of `v` over the block (TL_THREADS threads); *total = block sum
__device__ __forceinline__ uint32_t tl_block_exclusive(uint32_t v, uint32_t *total)
{
	__shared__ uint32_t s_warp[TL_THREADS / 32];
	__shared__ uint32_t s_total;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (uint32_t)o)
			incl += x;
	}
	if (lane == 31)
		s_warp[warp] = incl;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < TL_THREADS / 32; w++) {
			const uint32_t x = s_warp[w];
			s_warp[w] = t;
			t += x;
		}
		s_total = t;
	}
	__syncthreads();
	const uint32_t r = s_warp[warp] + incl - v;
	*total = s_total;
	__syncthreads();
	return r;
}

// pass A: zeros per tile.  key == nullptr: the flag is flags[row] (compaction: 0 = keep ... see callers)
__global__ void __launch_bounds__(TL_THREADS) k_split_count(const uint64_t *__restrict__ key, int bit, const uint32_t *__restrict__ flags, int invert,
		uint64_t n, uint32_t *__restrict__ tile_zeros)
{
	uint32_t mine[TL_PER_THREAD];
	const uint64_t base = (uint64_t)blockIdx.x * TL_TILE;
	const uint32_t z = tl_tile_zeros(base, n, [&](uint64_t i) -> uint32_t {
		const uint32_t f = key ? (uint32_t)((key[i] >> bit) & 1ull) : (flags[i] != 0 ? 1u : 0u);
		return f ^ (uint32_t)invert;
	}, mine);
	uint32_t total;
	tl_block_exclusive(z, &total);
	if (threadIdx.x == 0)
		tile_zeros[blockIdx.x] = total;
}

// pass B: stable scatter.  Elements with flag 0 go to [0, total_zeros) in order, the others behind them in order
// (drop_ones: they are dropped - compaction).  Moves the keys (if any) and the row numbers together.
__global__ void __launch_bounds__(TL_THREADS) k_split_scatter(const uint64_t *__restrict__ key, int bit, const uint32_t *__restrict__ flags, int invert,
		uint64_t n, const uint64_t *__restrict__ tile_off, const uint64_t *__restrict__ d_total_zeros, int drop_ones,
		const uint32_t *__restrict__ rows_in, uint64_t *__restrict__ key_out, uint32_t *__restrict__ rows_out)
{
	uint32_t mine[TL_PER_THREAD];
	const uint64_t base = (uint64_t)blockIdx.x * TL_TILE;
	const uint32_t z = tl_tile_zeros(base, n, [&](uint64_t i) -> uint32_t {
		const uint32_t f = key ? (uint32_t)((key[i] >> bit) & 1ull) : (flags[i] != 0 ? 1u : 0u);
		return f ^ (uint32_t)invert;
	}, mine);
	uint32_t total;
	uint32_t zeros_before = tl_block_exclusive(z, &total);
	const uint64_t zero_base = tile_off[blockIdx.x], total_zeros = *d_total_zeros;
	const uint64_t first = base + (uint64_t)threadIdx.x * TL_PER_THREAD;
#pragma unroll
	for (int j = 0; j < TL_PER_THREAD; j++) {
		const uint64_t i = first + j;
		if (i >= n)
			break;
		uint64_t dst;
		if (mine[j] == 0) {
			dst = zero_base + zeros_before;
			zeros_before++;
		} else {
			if (drop_ones)
				continue;
			// ones before element i = i - zeros before i
			dst = total_zeros + (i - (zero_base + zeros_before));
		}
		rows_out[dst] = rows_in[i];
		if (key_out)
			key_out[dst] = key[i];
	}
}

// ------------------------------------------------------------------ sort keys

__global__ void k_iota_u32(uint32_t *__restrict__ p, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		p[i] = (uint32_t)i;
}

// order-preserving unsigned key of one column for the rows in permutation order; NULL cells get key 0 and flag 1.
// Also reduces OR and AND of the keys of non-NULL rows so that the host can skip the bits that never vary.
__global__ void k_make_keys(TLCol col, int desc, const uint32_t *__restrict__ rows, uint64_t n, uint64_t *__restrict__ key,
		uint32_t *__restrict__ isnull, unsigned long long *__restrict__ or_and /* [0] OR, [1] AND, [2] NULL count */)
{
	unsigned long long o = 0, a = ~0ull, nn = 0;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t r = rows[i];
		const bool nul = col.nulls && col.nulls[r];
		unsigned long long k = 0;
		if (!nul) {
			long long v = col.cells[r];
			if (col.is_dbl) {
				if (v == (long long)0x8000000000000000ull)
					v = 0; // -0.0 sorts (and compares) equal to +0.0
				v = mdb_dbl_to_ordered(v);
			}
			k = (unsigned long long)v ^ 0x8000000000000000ull;
			if (desc)
				k = ~k;
			o |= k;
			a &= k;
		} else {
			nn++;
		}
		key[i] = k;
		isnull[i] = nul ? 1u : 0u;
	}
	for (int s = 16; s > 0; s >>= 1) {
		o |= __shfl_xor_sync(0xffffffffu, o, s);
		a &= __shfl_xor_sync(0xffffffffu, a, s);
		nn += __shfl_xor_sync(0xffffffffu, nn, s);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicOr(&or_and[0], o);
		atomicAnd(&or_and[1], a);
		if (nn)
			atomicAdd(&or_and[2], nn);
	}
}

// gather 4-byte flags through a permutation (the NULL flags follow the rows between the value passes and the NULL pass)
__global__ void k_gather_u32(const uint32_t *__restrict__ src_by_row, const uint32_t *__restrict__ rows, uint64_t n, uint32_t *__restrict__ dst)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		dst[i] = src_by_row[rows[i]];
}

// NULL flag of the row at every position of the permutation
__global__ void k_null_flags(TLCol col, const uint32_t *__restrict__ rows, uint64_t n, uint32_t *__restrict__ out)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = (col.nulls && col.nulls[rows[i]]) ? 1u : 0u;
}

// DISTINCT: differs[i] = 0 when row rows[i] equals row rows[i - 1] in every column (NULL = NULL), else 1 ... inverted for the
// compaction, whose flag 0 means "keep": keep_flag[i] = 0 for the first row of every run
__global__ void k_mark_duplicates(TLCols cols, const uint32_t *__restrict__ rows, uint64_t n, uint32_t *__restrict__ drop)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		bool same = i > 0;
		if (same) {
			const uint32_t r = rows[i], q = rows[i - 1];
			for (int c = 0; c < cols.n && same; c++) {
				const bool nr = cols.c[c].nulls && cols.c[c].nulls[r], nq = cols.c[c].nulls && cols.c[c].nulls[q];
				if (nr || nq) {
					same = nr && nq;
					continue;
				}
				long long x = cols.c[c].cells[r], y = cols.c[c].cells[q];
				if (cols.c[c].is_dbl) { // -0.0 = +0.0
					x = x == (long long)0x8000000000000000ull ? 0 : x;
					y = y == (long long)0x8000000000000000ull ? 0 : y;
				}
				same = x == y;
			}
		}
		drop[i] = same ? 1u : 0u;
	}
}

__global__ void k_gather_rows(const int64_t *__restrict__ cells, const uint8_t *__restrict__ nulls, const uint32_t *__restrict__ rows, uint64_t n,
		int64_t *__restrict__ out_cells, uint8_t *__restrict__ out_nulls)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t r = rows[i];
		out_cells[i] = cells[r];
		if (out_nulls)
			out_nulls[i] = nulls ? nulls[r] : 0;
	}
}

__global__ void k_gather_u64(const uint64_t *__restrict__ src, const uint32_t *__restrict__ rows, uint64_t n, uint64_t *__restrict__ dst)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		dst[i] = src[rows[i]];
}

// ------------------------------------------------------------------ host side

namespace {

struct Tail {
	mdbcu_ctx *ctx;
	DevTemp tmp;
	uint64_t n;            // rows in the permutation
	uint32_t *rows[2];     // ping-pong permutation buffers; rows[cur] is current
	uint64_t *key[2];
	uint32_t *flag[2];     // scratch flags (per position) / NULL flags
	uint32_t *tile_zeros;
	uint64_t *tile_off, *d_total;
	unsigned long long *d_red;
	int cur = 0;
	explicit Tail(mdbcu_ctx *c) : ctx(c), tmp(c) {}

	int grid_for(uint64_t count) const
	{
		return (int)std::max<uint64_t>(1, std::min<uint64_t>(mdb_div_up(count, 256), (uint64_t)ctx->num_sms * 8));
	}

	int init(uint64_t nrows)
	{
		n = nrows;
		const size_t tiles = mdb_div_up(nrows, TL_TILE) + 1;
		for (int b = 0; b < 2; b++) {
			MDB_TRY(tmp.alloc(&rows[b], nrows));
			MDB_TRY(tmp.alloc(&key[b], nrows));
			MDB_TRY(tmp.alloc(&flag[b], nrows));
		}
		MDB_TRY(tmp.alloc(&tile_zeros, tiles));
		MDB_TRY(tmp.alloc(&tile_off, tiles));
		MDB_TRY(tmp.alloc(&d_total, 1));
		MDB_TRY(tmp.alloc(&d_red, 4));
		MDB_LAUNCH(ctx, k_iota_u32, grid_for(nrows), 256, 0, rows[0], nrows);
		return MDBCU_OK;
	}

	// one stable split of the current permutation: by bit `bit` of key[cur] (with_key), or by flags[] != 0.  drop: compaction
	int split(bool with_key, int bit, const uint32_t *flags, int invert, bool drop)
	{
		if (n == 0)
			return MDBCU_OK;
		const unsigned tiles = (unsigned)mdb_div_up(n, TL_TILE);
		const uint64_t *k = with_key ? key[cur] : nullptr;
		MDB_LAUNCH(ctx, k_split_count, tiles, TL_THREADS, 0, k, bit, flags, invert, n, tile_zeros);
		MDB_TRY(mdb_scan_u32_u64(ctx, tile_zeros, tile_off, tiles, d_total));
		MDB_LAUNCH(ctx, k_split_scatter, tiles, TL_THREADS, 0, k, bit, flags, invert, n, (const uint64_t*)tile_off, (const uint64_t*)d_total,
				drop ? 1 : 0, (const uint32_t*)rows[cur], with_key ? key[cur ^ 1] : (uint64_t*)nullptr, rows[cur ^ 1]);
		CUDA_CHECK_LAUNCH(ctx);
		cur ^= 1;
		if (drop) {
			uint64_t kept = 0;
			MDB_TRY(mdb_read_u64(ctx, d_total, &kept));
			n = kept;
		}
		return MDBCU_OK;
	}

	// stable sort of the current permutation by one column
	int sort_by(const TLCol &col, bool desc)
	{
		if (n < 2)
			return MDBCU_OK;
		CUDA_TRY(ctx, cudaMemsetAsync(d_red, 0, 4 * sizeof(unsigned long long), ctx->stream));
		CUDA_TRY(ctx, cudaMemsetAsync(d_red + 1, 0xff, sizeof(unsigned long long), ctx->stream));
		MDB_LAUNCH(ctx, k_make_keys, grid_for(n), 256, 0, col, desc ? 1 : 0, (const uint32_t*)rows[cur], n, key[cur], flag[0], d_red);
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, d_red, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		const uint64_t vor = ctx->h_scalar[0], vand = ctx->h_scalar[1], nnull = ctx->h_scalar[2];
		// NULL rows carry key 0: a bit varies if it varies among the values, or if it is set in some value while NULLs exist
		const uint64_t varying = nnull == n ? 0 : ((vor & ~vand) | (nnull ? vor : 0));
		for (int bit = 0; bit < 64; bit++)
			if ((varying >> bit) & 1ull)
				MDB_TRY(split(true, bit, nullptr, 0, false));
		if (nnull && nnull != n)
			MDB_TRY(null_pass(col, desc));
		return MDBCU_OK;
	}

	// NULLs first when ascending, last when descending (sqlite3)
	int null_pass(const TLCol &col, bool desc)
	{
		MDB_LAUNCH(ctx, k_null_flags, grid_for(n), 256, 0, col, (const uint32_t*)rows[cur], n, flag[1]);
		// ascending: NULL rows (flag 1) must go first = they must be the "zeros": invert
		return split(false, 0, flag[1], desc ? 0 : 1, false);
	}
};

} // namespace

bool mdb_plan_has_tail(const mdbcu_plan *plan)
{
	return plan->distinct || plan->n_having > 0 || plan->n_order > 0 || plan->has_limit;
}

int mdb_validate_tail(mdbcu_ctx *ctx, const mdbcu_plan *plan)
{
	if (plan->n_having < 0 || plan->n_having > MDBCU_MAX_HAVING || plan->n_order < 0 || plan->n_order > MDBCU_MAX_ORDER)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: HAVING / ORDER BY list too long");
	if (plan->n_having > 0 && !tl_program_ok(plan))
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: malformed HAVING program");
	for (int k = 0; k < plan->n_order; k++)
		if (plan->order[k].out_col < 0 || plan->order[k].out_col >= plan->n_out)
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: ORDER BY column %d is not a result column", plan->order[k].out_col);
	if (plan->has_limit && (plan->limit < 0 || plan->offset < 0))
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_select: negative LIMIT / OFFSET");
	// distributed plans: a plan over ONE sharded table returns all of its rows on rank 0 (mdb_fast.cu, mdb_dist_group.cu), where
	// the tail operators then see the whole result; a join's result stays spread over the ranks that own its keys
	if (mdb_plan_has_tail(plan) && (plan->flags & MDBCU_PLAN_DISTRIBUTED) && plan->n_tables != 1)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "HAVING / DISTINCT / ORDER BY / LIMIT are not available in distributed joins "
				"(every rank holds a part of the result)");
	return MDBCU_OK;
}

// HAVING, DISTINCT, ORDER BY, LIMIT on `res` (in that order), in place
int mdb_apply_tail(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (!mdb_plan_has_tail(plan))
		return MDBCU_OK;
	if (res->nrows >= (1ull << 32))
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "tail operators work on results of fewer than 2^32 rows");
	const bool reorder = plan->distinct || plan->n_order > 0;
	if (res->nrows > 0) {
		Tail t(ctx);
		MDB_TRY(t.init(res->nrows));
		TLCols cols;
		memset(&cols, 0, sizeof(cols));
		cols.n = (int)res->cols.size();
		for (int c = 0; c < cols.n; c++) {
			cols.c[c].cells = res->cols[c].cells;
			cols.c[c].nulls = res->cols[c].nulls;
			cols.c[c].is_dbl = res->cols[c].type == MDBCU_CT_DOUBLE;
		}
		// results of the general operators carry a row-order key (the reference's row order is produced when the rows are
		// fetched): LIMIT without ORDER BY must cut in THAT order, so it becomes an explicit sort here
		if (res->order_key && !reorder && plan->has_limit) {
			TLCol oc;
			oc.cells = reinterpret_cast<const int64_t*>(res->order_key);
			oc.nulls = nullptr;
			oc.is_dbl = 0;
			// (the key is unsigned: as a signed cell with the sign bit flipped back it keeps its order below 2^63)
			MDB_TRY(t.sort_by(oc, false));
		}
		if (plan->n_having > 0) {
			TLProgram prog;
			memset(&prog, 0, sizeof(prog));
			prog.n = plan->n_having;
			memcpy(prog.ops, plan->having, sizeof(mdbcu_pred_op) * plan->n_having);
			// evaluated per ROW NUMBER; the permutation is still the identity (or the row-order sort above: flags go by position)
			MDB_LAUNCH(ctx, k_having_eval, t.grid_for(res->nrows), 256, 0, prog, cols, res->nrows, t.flag[0]);
			MDB_LAUNCH(ctx, k_gather_u32, t.grid_for(t.n), 256, 0, (const uint32_t*)t.flag[0], (const uint32_t*)t.rows[t.cur], t.n, t.flag[1]);
			MDB_TRY(t.split(false, 0, t.flag[1], 1, true)); // keep = flag 1 -> inverted: zeros are kept
		}
		if (plan->distinct && t.n > 1) {
			for (int c = cols.n - 1; c >= 0; c--)
				MDB_TRY(t.sort_by(cols.c[c], false));
			MDB_LAUNCH(ctx, k_mark_duplicates, t.grid_for(t.n), 256, 0, cols, (const uint32_t*)t.rows[t.cur], t.n, t.flag[1]);
			MDB_TRY(t.split(false, 0, t.flag[1], 0, true)); // drop flag 1 = duplicates
		}
		for (int k = plan->n_order - 1; k >= 0; k--)
			MDB_TRY(t.sort_by(cols.c[plan->order[k].out_col], plan->order[k].desc != 0));
		uint64_t first = 0, count = t.n;
		if (plan->has_limit) {
			first = std::min<uint64_t>((uint64_t)plan->offset, t.n);
			count = std::min<uint64_t>((uint64_t)plan->limit, t.n - first);
		}
		const uint32_t *perm = t.rows[t.cur] + first;
		// the surviving rows in their final order
		for (auto &c : res->cols) {
			int64_t *nc = nullptr;
			uint8_t *nn = nullptr;
			MDB_TRY(mdb_alloc(ctx, &nc, count));
			if (c.nulls)
				MDB_TRY(mdb_alloc(ctx, &nn, count));
			if (count)
				MDB_LAUNCH(ctx, k_gather_rows, t.grid_for(count), 256, 0, (const int64_t*)c.cells, (const uint8_t*)c.nulls, perm, count, nc, nn);
			mdb_free(ctx, c.cells);
			mdb_free(ctx, c.nulls);
			c.cells = nc;
			c.nulls = nn;
		}
		if (res->order_key) {
			if (reorder || plan->has_limit) {
				// the rows are now physically in their final order
				mdb_free(ctx, res->order_key);
				res->order_key = nullptr;
			} else {
				uint64_t *nk = nullptr;
				MDB_TRY(mdb_alloc(ctx, &nk, count));
				if (count)
					MDB_LAUNCH(ctx, k_gather_u64, t.grid_for(count), 256, 0, (const uint64_t*)res->order_key, perm, count, nk);
				mdb_free(ctx, res->order_key);
				res->order_key = nk;
			}
		}
		CUDA_CHECK_LAUNCH(ctx);
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // the temporaries of `t` are released when it goes out of scope
		res->nrows = count;
	}
	return MDBCU_OK;
}
