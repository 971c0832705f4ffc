// mdb_star.cu - small-build / large-probe star join with grouped aggregates (BASELINE config 5):
//
//   SELECT d.g, MIN(f.m), MAX(f.m) [, SUM / AVG / COUNT ...] FROM D INNER JOIN F ON d.id = f.fk GROUP BY d.g
//
// replaces _join_nested_loop_tbl2tbl (src/engine/executor_select.c:1076) + proc_groupby_clause (:1526) for a
// dimension whose keys are unique and dense enough for a direct-addressed table: key - kmin -> 16-bit group id,
// at most 65536 keys = 128 KiB of shared memory per CTA.  The fact table is streamed ONCE (8 B per referenced cell,
// 256-bit loads); every row costs one shared-memory lookup and one shared-memory atomic per aggregate into the CTA's
// private per-group accumulators, which are merged into global memory when the CTA is done.  No candidate pair, no
// joined row, no hash probe in HBM.  Anything outside that shape (duplicate dimension keys, sparse keys, too many
// groups, NULL group values, predicates) returns MDBCU_EUNSUPPORTED and the general operators answer.
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>

#define ST_MAX_KEYS 65536          // dimension key range covered by the direct-addressed table
#define ST_MAX_AGGS 4
#define ST_MAX_COLS 3              // fact columns referenced: the foreign key + up to two aggregated columns
#define ST_THREADS 1024
#ifndef ST_UNROLL
#define ST_UNROLL 2                // 32-byte vectors per column a thread has in flight
#endif
#define ST_EMPTY 0xffffu
#define ST_ACC_BYTES (64 * 1024)   // shared memory for the per-group accumulators (table + accumulators stay under 195 KiB)

struct StarAgg {
	int32_t kind; // MDBCU_OUT_*
	int32_t col;  // index into StarSpec::data (0 = the foreign key), -1 for COUNT(*)
	int32_t is_dbl;
	int32_t track_nn; // the column may hold NULLs: count the non-NULL inputs (otherwise that count = rows of the group)
};

struct StarSpec {
	const int64_t *data[ST_MAX_COLS];
	const uint32_t *present[ST_MAX_COLS]; // nullptr: all present
	int32_t ncols, naggs;
	long long kmin;
	uint32_t range;    // dimension keys lie in [kmin, kmin + range)
	uint32_t ngroups;  // group ids are g - gmin in [0, ngroups)
	StarAgg aggs[ST_MAX_AGGS];
	// global accumulators (merged from the CTAs' shared-memory copies)
	unsigned int *g_rows;         // [ngroups] joined rows per group
	long long *g_acc;             // [naggs][ngroups]
	unsigned long long *g_nn;     // [naggs][ngroups] non-NULL values per aggregate
};

static bool st_all_present(const mdbcu_table *t, int col)
{
	return t->all_live && !t->cols[col].has_nulls;
}

// dimension -> table[key - kmin] = g - gmin.  A key seen twice, or a NULL group value, makes the shape unsupported.
__global__ void k_star_build(const int64_t *__restrict__ keys, const uint32_t *__restrict__ key_present, const int64_t *__restrict__ grp,
		const uint32_t *__restrict__ grp_present, uint64_t n, long long kmin, uint32_t range, long long gmin, uint32_t ngroups,
		unsigned int *__restrict__ table, unsigned int *__restrict__ flags)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		if (key_present && !mdb_bit(key_present, i))
			continue; // dead row or NULL key: never joins (executor_select.c:716-738)
		if (grp_present && !mdb_bit(grp_present, i)) {
			atomicOr(flags, 1u); // NULL group value
			continue;
		}
		const unsigned long long off = (unsigned long long)keys[i] - (unsigned long long)kmin;
		const unsigned long long gid = (unsigned long long)grp[i] - (unsigned long long)gmin;
		if (off >= range || gid >= ngroups) {
			atomicOr(flags, 2u); // statistics out of date
			continue;
		}
		if (atomicCAS(&table[off], 0xffffffffu, (unsigned int)gid) != 0xffffffffu) // (32-bit build table; the probe keeps 16 bits)
			atomicOr(flags, 4u); // duplicate dimension key
	}
}

__device__ __forceinline__ void st_load256(const void *p, uint32_t *w)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			: "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

// group id of one fact row (ST_EMPTY: no partner in the dimension)
__device__ __forceinline__ uint32_t st_lookup(const StarSpec &sp, const uint16_t *tab, long long key, bool present)
{
	const unsigned long long off = (unsigned long long)key - (unsigned long long)sp.kmin;
	return (present && off < sp.range) ? tab[off] : ST_EMPTY;
}

// one value into one accumulator slot.  KIND and DBL are compile-time: the plan is decoded once per batch of rows, not per row
template <int KIND, bool DBL>
__device__ __forceinline__ void st_update(long long *slot, long long x)
{
	if (KIND == MDBCU_OUT_SUM) { // (AVG accumulates like SUM)
		if (DBL)
			atomicAdd(reinterpret_cast<double*>(slot), __longlong_as_double(x));
		else
			atomicAdd(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)x);
	} else if (KIND == MDBCU_OUT_MIN || KIND == MDBCU_OUT_MAX) {
		// a 64-bit shared-memory min/max is a compare-and-swap loop; after the first few rows of a group almost no row
		// improves the extremum, so look first (a stale read only costs an atomic that changes nothing)
		const long long y = DBL ? mdb_dbl_to_ordered(x) : x;
		const long long cur = *reinterpret_cast<volatile long long*>(slot);
		if (KIND == MDBCU_OUT_MIN ? y < cur : y > cur) {
			if (KIND == MDBCU_OUT_MIN)
				atomicMin(slot, y);
			else
				atomicMax(slot, y);
		}
	}
}

// a batch of ROWS rows (group ids g[], cells of the aggregated column lo[]/hi[], presence bits) into aggregate a
template <int KIND, bool DBL, int ROWS>
__device__ __forceinline__ void st_batch(const uint32_t *g, const uint32_t *lo, const uint32_t *hi, uint32_t present_bits, bool track_nn,
		long long *acc, unsigned int *nn)
{
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		if (g[r] == ST_EMPTY || !((present_bits >> r) & 1u))
			continue;
		if (track_nn)
			atomicAdd(&nn[g[r]], 1u);
		if (KIND != MDBCU_OUT_COUNT_COL)
			st_update<KIND, DBL>(&acc[g[r]], (long long)(((unsigned long long)hi[r] << 32) | lo[r]));
	}
}

template <int ROWS>
__device__ __forceinline__ void st_batch_dispatch(const StarAgg &ag, const uint32_t *g, const uint32_t *lo, const uint32_t *hi,
		uint32_t present_bits, long long *acc, unsigned int *nn)
{
	const bool t = ag.track_nn != 0;
	switch (ag.kind) { // uniform: one decision per batch
	case MDBCU_OUT_SUM: case MDBCU_OUT_AVG:
		if (ag.is_dbl)
			st_batch<MDBCU_OUT_SUM, true, ROWS>(g, lo, hi, present_bits, t, acc, nn);
		else
			st_batch<MDBCU_OUT_SUM, false, ROWS>(g, lo, hi, present_bits, t, acc, nn);
		break;
	case MDBCU_OUT_MIN:
		if (ag.is_dbl)
			st_batch<MDBCU_OUT_MIN, true, ROWS>(g, lo, hi, present_bits, t, acc, nn);
		else
			st_batch<MDBCU_OUT_MIN, false, ROWS>(g, lo, hi, present_bits, t, acc, nn);
		break;
	case MDBCU_OUT_MAX:
		if (ag.is_dbl)
			st_batch<MDBCU_OUT_MAX, true, ROWS>(g, lo, hi, present_bits, t, acc, nn);
		else
			st_batch<MDBCU_OUT_MAX, false, ROWS>(g, lo, hi, present_bits, t, acc, nn);
		break;
	case MDBCU_OUT_COUNT_COL:
		st_batch<MDBCU_OUT_COUNT_COL, false, ROWS>(g, lo, hi, present_bits, true, acc, nn);
		break;
	default: // COUNT(*): the rows counter of the group
		break;
	}
}

// Probe: one persistent CTA per SM.  Shared memory: the 16-bit direct table, then rows / acc / nn per group.
// Rows are taken eight at a time (two 256-bit loads per referenced column): first their group ids, then aggregate by
// aggregate - the plan (which aggregate, which column, which type) is decoded once per eight rows.
template <int NC, int NA>
__global__ void __launch_bounds__(ST_THREADS, 1) k_star_probe(StarSpec sp, const unsigned int *__restrict__ table, uint64_t n)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint16_t *tab = reinterpret_cast<uint16_t*>(smem_raw);
	const uint32_t tab_bytes = ((sp.range * 2u) + 15u) & ~15u;
	long long *s_acc = reinterpret_cast<long long*>(smem_raw + tab_bytes);
	unsigned int *s_nn = reinterpret_cast<unsigned int*>(s_acc + (size_t)NA * sp.ngroups);
	unsigned int *s_rows = s_nn + (size_t)NA * sp.ngroups;
	const uint32_t tid = threadIdx.x, G = sp.ngroups;

	for (uint32_t i = tid; i < sp.range; i += ST_THREADS)
		tab[i] = (uint16_t)table[i];
	for (uint32_t i = tid; i < NA * G; i += ST_THREADS) {
		const int kind = sp.aggs[i / G].kind;
		s_acc[i] = kind == MDBCU_OUT_MIN ? INT64_MAX : kind == MDBCU_OUT_MAX ? INT64_MIN : 0; // 0 is also +0.0
		s_nn[i] = 0;
	}
	for (uint32_t i = tid; i < G; i += ST_THREADS)
		s_rows[i] = 0;
	__syncthreads();

	const uint64_t nquads = n / 4, stride = (uint64_t)gridDim.x * ST_THREADS;
	uint64_t quad = (uint64_t)blockIdx.x * ST_THREADS + tid;
	constexpr int UNROLL = ST_UNROLL, ROWS = 4 * UNROLL;
	for (; quad + (UNROLL - 1) * stride < nquads; quad += UNROLL * stride) {
		uint32_t raw[NC][UNROLL][8], pbits[NC];
#pragma unroll
		for (int c = 0; c < NC; c++) {
			pbits[c] = 0;
#pragma unroll
			for (int u = 0; u < UNROLL; u++) {
				const uint64_t qi = quad + u * stride;
				st_load256(reinterpret_cast<const char*>(sp.data[c]) + qi * 32, raw[c][u]);
				const uint32_t pw = sp.present[c] ? sp.present[c][qi >> 3] : 0xffffffffu;
				pbits[c] |= ((pw >> ((qi & 7) * 4)) & 0xfu) << (4 * u);
			}
		}
		uint32_t g[ROWS], lo[ROWS], hi[ROWS];
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			const long long key = (long long)(((unsigned long long)raw[0][r / 4][2 * (r % 4) + 1] << 32) | raw[0][r / 4][2 * (r % 4)]);
			g[r] = st_lookup(sp, tab, key, (pbits[0] >> r) & 1u);
		}
#pragma unroll
		for (int r = 0; r < ROWS; r++)
			if (g[r] != ST_EMPTY)
				atomicAdd(&s_rows[g[r]], 1u);
#pragma unroll
		for (int a = 0; a < NA; a++) {
			const StarAgg &ag = sp.aggs[a];
#pragma unroll
			for (int c = 0; c < NC; c++) {
				if (ag.col != c) // uniform
					continue;
#pragma unroll
				for (int r = 0; r < ROWS; r++) {
					lo[r] = raw[c][r / 4][2 * (r % 4)];
					hi[r] = raw[c][r / 4][2 * (r % 4) + 1];
				}
				st_batch_dispatch<ROWS>(ag, g, lo, hi, pbits[c], s_acc + (size_t)a * G, s_nn + (size_t)a * G);
			}
		}
	}
	for (uint64_t r = quad * 4; r < n; r += 4 * stride) { // remainder: single rows through the same code
		for (int k = 0; k < 4 && r + k < n; k++) {
			const uint64_t row = r + k;
			const bool p0 = sp.present[0] ? mdb_bit(sp.present[0], row) : true;
			const uint32_t g1 = st_lookup(sp, tab, sp.data[0][row], p0);
			if (g1 == ST_EMPTY)
				continue;
			atomicAdd(&s_rows[g1], 1u);
			for (int a = 0; a < NA; a++) {
				const StarAgg &ag = sp.aggs[a];
				if (ag.col < 0)
					continue;
				const long long x = sp.data[ag.col][row];
				const uint32_t lo1 = (uint32_t)(unsigned long long)x, hi1 = (uint32_t)((unsigned long long)x >> 32);
				const uint32_t pb = (sp.present[ag.col] ? mdb_bit(sp.present[ag.col], row) : true) ? 1u : 0u;
				st_batch_dispatch<1>(ag, &g1, &lo1, &hi1, pb, s_acc + (size_t)a * G, s_nn + (size_t)a * G);
			}
		}
	}
	__syncthreads();

	// merge this CTA's groups into the global accumulators
	for (uint32_t g = tid; g < G; g += ST_THREADS)
		if (s_rows[g])
			atomicAdd(&sp.g_rows[g], s_rows[g]);
	for (uint32_t i = tid; i < NA * G; i += ST_THREADS) {
		const StarAgg &ag = sp.aggs[i / G];
		if (ag.kind == MDBCU_OUT_COUNT_STAR || (ag.track_nn ? s_nn[i] == 0 : s_rows[i % G] == 0))
			continue;
		if (ag.track_nn)
			atomicAdd(&sp.g_nn[i], (unsigned long long)s_nn[i]);
		switch (ag.kind) {
		case MDBCU_OUT_SUM: case MDBCU_OUT_AVG:
			if (ag.is_dbl)
				atomicAdd(reinterpret_cast<double*>(&sp.g_acc[i]), __longlong_as_double(s_acc[i]));
			else
				atomicAdd(reinterpret_cast<unsigned long long*>(&sp.g_acc[i]), (unsigned long long)s_acc[i]);
			break;
		case MDBCU_OUT_MIN:
			atomicMin(&sp.g_acc[i], s_acc[i]);
			break;
		case MDBCU_OUT_MAX:
			atomicMax(&sp.g_acc[i], s_acc[i]);
			break;
		default:
			break;
		}
	}
}

struct StarOut {
	int nout;
	int kind[MDBCU_MAX_OUT];   // MDBCU_OUT_COLUMN = the group value
	int agg[MDBCU_MAX_OUT];    // index into StarSpec::aggs
	int64_t *cells[MDBCU_MAX_OUT];
	uint8_t *nulls[MDBCU_MAX_OUT];
	unsigned long long *nrows;
	long long gmin;
};

// one result row per group that joined at least one fact row (any order: the caller canonicalises)
__global__ void k_star_emit(StarSpec sp, StarOut out)
{
	for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < sp.ngroups; g += gridDim.x * blockDim.x) {
		const unsigned int rows = sp.g_rows[g];
		if (!rows)
			continue;
		const unsigned long long r = atomicAdd(out.nrows, 1ull);
		for (int o = 0; o < out.nout; o++) {
			long long cell = 0;
			bool isnull = false;
			if (out.kind[o] == MDBCU_OUT_COLUMN) {
				cell = out.gmin + (long long)g;
			} else if (out.kind[o] == MDBCU_OUT_COUNT_STAR) {
				cell = rows;
			} else {
				const int a = out.agg[o];
				const unsigned long long nn = sp.aggs[a].track_nn ? sp.g_nn[(size_t)a * sp.ngroups + g] : rows;
				const long long acc = sp.g_acc[(size_t)a * sp.ngroups + g];
				const bool dbl = sp.aggs[a].is_dbl;
				switch (out.kind[o]) {
				case MDBCU_OUT_COUNT_COL:
					cell = (long long)nn;
					break;
				case MDBCU_OUT_SUM:
					isnull = nn == 0;
					cell = acc;
					break;
				case MDBCU_OUT_MIN: case MDBCU_OUT_MAX:
					isnull = nn == 0;
					cell = dbl ? mdb_ordered_to_dbl(acc) : acc;
					break;
				case MDBCU_OUT_AVG:
					isnull = nn == 0;
					if (!isnull)
						cell = __double_as_longlong((dbl ? __longlong_as_double(acc) : (double)acc) / (double)nn);
					break;
				}
			}
			out.cells[o][r] = isnull ? 0 : cell;
			out.nulls[o][r] = isnull;
		}
	}
}

// distributed plan: the ranks' direct tables ([range entries | flag word] each) -> the whole dimension's table
__global__ void k_star_merge_tables(const unsigned int *__restrict__ all, int world, uint32_t range, unsigned int *__restrict__ table,
		unsigned int *__restrict__ flags)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) {
		unsigned int f = 0;
		for (int r = 0; r < world; r++)
			f |= all[(size_t)r * (range + 1) + range];
		if (f)
			atomicOr(flags, f);
	}
	if (i >= range)
		return;
	unsigned int v = 0xffffffffu;
	for (int r = 0; r < world; r++) {
		const unsigned int x = all[(size_t)r * (range + 1) + i];
		if (x == 0xffffffffu)
			continue;
		if (v != 0xffffffffu)
			atomicOr(flags, 4u); // the key is in two shards: duplicate dimension key
		v = x;
	}
	table[i] = v;
}

// distributed plan: all ranks' accumulator blocks folded into sp.g_* in rank order (layout: mdb_select_direct_star)
__global__ void k_star_merge_acc(StarSpec sp, const unsigned char *__restrict__ all, int world, size_t rows_bytes, size_t block_bytes)
{
	const uint32_t G = sp.ngroups;
	for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
		unsigned int rows = 0;
		for (int r = 0; r < world; r++)
			rows += reinterpret_cast<const unsigned int*>(all + (size_t)r * block_bytes)[g];
		sp.g_rows[g] = rows;
		for (int a = 0; a < sp.naggs; a++) {
			const int kind = sp.aggs[a].kind;
			const bool dsum = (kind == MDBCU_OUT_SUM || kind == MDBCU_OUT_AVG) && sp.aggs[a].is_dbl;
			long long acc = kind == MDBCU_OUT_MIN ? INT64_MAX : kind == MDBCU_OUT_MAX ? INT64_MIN : 0;
			double dacc = 0.0;
			unsigned long long nn = 0;
			for (int r = 0; r < world; r++) {
				const unsigned char *blk = all + (size_t)r * block_bytes + rows_bytes;
				const long long x = reinterpret_cast<const long long*>(blk)[(size_t)a * G + g];
				nn += reinterpret_cast<const unsigned long long*>(blk + (size_t)sp.naggs * G * 8)[(size_t)a * G + g];
				if (kind == MDBCU_OUT_MIN)
					acc = x < acc ? x : acc;
				else if (kind == MDBCU_OUT_MAX)
					acc = x > acc ? x : acc;
				else if (dsum)
					dacc += __longlong_as_double(x);
				else
					acc = (long long)((unsigned long long)acc + (unsigned long long)x);
			}
			sp.g_acc[(size_t)a * G + g] = dsum ? __double_as_longlong(dacc) : acc;
			sp.g_nn[(size_t)a * G + g] = nn;
		}
	}
}

__global__ void k_star_init_acc(StarSpec sp)
{
	const uint32_t G = sp.ngroups;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sp.naggs * G; i += gridDim.x * blockDim.x) {
		const int kind = sp.aggs[i / G].kind;
		sp.g_acc[i] = kind == MDBCU_OUT_MIN ? INT64_MAX : kind == MDBCU_OUT_MAX ? INT64_MIN : 0;
	}
}

template <int NC>
static void st_launch_probe(mdbcu_ctx *ctx, int grid, size_t smem, int na, const StarSpec &sp, const unsigned int *table, uint64_t n)
{
	auto launch = [&](auto kernel) {
		// (set on every launch: the limit belongs to the current device, one process may drive several, and the call is cheap)
		cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 195 * 1024);
		MDB_LAUNCH(ctx, kernel, grid, ST_THREADS, smem, sp, table, n);
	};
	if (na <= 1)
		launch(k_star_probe<NC, 1>);
	else if (na <= 2)
		launch(k_star_probe<NC, 2>);
	else
		launch(k_star_probe<NC, 4>);
}

int mdb_select_direct_star(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (plan->n_tables != 2 || plan->n_joins != 1 || plan->joins[0].cross || plan->n_pred != 0 || plan->n_group != 1 ||
			plan->n_out < 1 || plan->n_out > MDBCU_MAX_OUT)
		return MDBCU_EUNSUPPORTED;
	// Distributed plan ("C5: replicate the small dimension, shard the fact - no all-to-all"): every rank holds a
	// shard of BOTH tables; the ranks' direct tables (<= 256 KiB each) are all-gathered and merged into the whole dimension,
	// every rank probes it with its fact shard, the per-group accumulators (<= 64 KiB per rank) are all-gathered and folded
	// in rank order on rank 0, which returns the groups; the other ranks return no row.  All decisions use the plan and
	// the GLOBAL statistics (mdbcu_table_sync_stats): the ranks take this path together or not at all.
	const bool dist = (plan->flags & MDBCU_PLAN_DISTRIBUTED) != 0;
	if (dist && !mdb_comm_ready(ctx))
		return mdb_fail(ctx, MDBCU_EERROR, "MDBCU_PLAN_DISTRIBUTED needs mdbcu_comm_init first");
	const int W = dist ? ctx->world : 1;
	const mdbcu_join &jn = plan->joins[0];
	if (jn.left.tbl == jn.right.tbl || jn.left.tbl < 0 || jn.left.tbl > 1 || jn.right.tbl < 0 || jn.right.tbl > 1)
		return MDBCU_EUNSUPPORTED;
	// the dimension is the table that carries the GROUP BY column, the fact table is the other one
	const int dt = plan->group[0].tbl, ft = 1 - dt;
	if (dt < 0 || dt > 1)
		return MDBCU_EUNSUPPORTED;
	const mdbcu_table *D = plan->tables[dt], *F = plan->tables[ft];
	const int dk = jn.left.tbl == dt ? jn.left.col : jn.right.col, fk = jn.left.tbl == ft ? jn.left.col : jn.right.col;
	const int gc = plan->group[0].col;
	if (dk < 0 || dk >= D->ncols || fk < 0 || fk >= F->ncols || gc < 0 || gc >= D->ncols)
		return MDBCU_EUNSUPPORTED;
	auto intlike = [](int type) { return type == MDBCU_CT_INTEGER || type == MDBCU_CT_DATE || type == MDBCU_CT_DATETIME || type == MDBCU_CT_TINYINT; };
	DevColumn ck = D->cols[dk], cg = D->cols[gc]; // copies: a distributed plan uses the bounds over all shards
	const DevColumn &cf = F->cols[fk];
	if (!intlike(ck.type) || !intlike(cg.type) || !intlike(cf.type))
		return MDBCU_EUNSUPPORTED;
	if (dist) {
		if (!ck.gstats_ok || !cg.gstats_ok || !D->global_slots || !F->global_slots)
			return mdb_fail(ctx, MDBCU_EERROR, "distributed plan: call mdbcu_table_sync_stats on every sharded table first");
		ck.imin = ck.gmin;
		ck.imax = ck.gmax;
		cg.imin = cg.gmin;
		cg.imax = cg.gmax;
	} else if (!ck.stats_ok || !cg.stats_ok) {
		return MDBCU_EUNSUPPORTED;
	}
	const uint64_t f_rows = dist ? F->global_slots : F->n_slots, d_rows = dist ? D->global_slots : D->n_slots;
	if (f_rows < (1ull << 20) || d_rows == 0 || d_rows > (1ull << 22))
		return MDBCU_EUNSUPPORTED; // small inputs: general operators (they also return the reference's row order)
	if (ck.imin > ck.imax || cg.imin > cg.imax)
		return MDBCU_EUNSUPPORTED;
	const unsigned long long range = (unsigned long long)ck.imax - (unsigned long long)ck.imin + 1ull;
	const unsigned long long ngroups = (unsigned long long)cg.imax - (unsigned long long)cg.imin + 1ull;
	if (range == 0 || range > ST_MAX_KEYS || ngroups == 0 || ngroups >= ST_EMPTY)
		return MDBCU_EUNSUPPORTED;

	// fact columns and aggregates
	StarSpec sp;
	memset(&sp, 0, sizeof(sp));
	StarOut out;
	memset(&out, 0, sizeof(out));
	int col_of[MDBCU_MAX_COLUMNS];
	for (int c = 0; c < MDBCU_MAX_COLUMNS; c++)
		col_of[c] = -1;
	sp.data[0] = cf.data;
	sp.present[0] = st_all_present(F, fk) ? nullptr : cf.present;
	sp.ncols = 1;
	col_of[fk] = 0;
	for (int o = 0; o < plan->n_out; o++) {
		const mdbcu_out &po = plan->out[o];
		out.kind[o] = po.kind;
		out.agg[o] = -1;
		if (po.kind == MDBCU_OUT_COLUMN) {
			if (po.ref.tbl != dt || po.ref.col != gc)
				return MDBCU_EUNSUPPORTED;
			continue;
		}
		if (po.kind == MDBCU_OUT_COUNT_STAR)
			continue;
		if (po.ref.tbl != ft || po.ref.col < 0 || po.ref.col >= F->ncols || F->cols[po.ref.col].type == MDBCU_CT_VARCHAR)
			return MDBCU_EUNSUPPORTED;
		if (sp.naggs == ST_MAX_AGGS)
			return MDBCU_EUNSUPPORTED;
		int sc = col_of[po.ref.col];
		if (sc < 0) {
			if (sp.ncols == ST_MAX_COLS)
				return MDBCU_EUNSUPPORTED;
			sc = sp.ncols++;
			col_of[po.ref.col] = sc;
			sp.data[sc] = F->cols[po.ref.col].data;
			sp.present[sc] = st_all_present(F, po.ref.col) ? nullptr : F->cols[po.ref.col].present;
		}
		StarAgg &ag = sp.aggs[sp.naggs];
		ag.kind = po.kind;
		ag.col = sc;
		ag.is_dbl = F->cols[po.ref.col].type == MDBCU_CT_DOUBLE;
		ag.track_nn = sp.present[sc] != nullptr;
		out.agg[o] = sp.naggs++;
	}
	for (int c = 0; c < sp.ncols; c++)
		if (((uintptr_t)sp.data[c] & 31u) != 0)
			return MDBCU_EUNSUPPORTED;
	const int na = sp.naggs <= 1 ? 1 : sp.naggs <= 2 ? 2 : 4;
	for (int a = sp.naggs; a < na; a++) { // padding slots of the template: COUNT(*) costs nothing per row
		sp.aggs[a].kind = MDBCU_OUT_COUNT_STAR;
		sp.aggs[a].col = -1;
	}
	const size_t acc_bytes = (size_t)ngroups * (na * 12 + 4);
	if (acc_bytes > ST_ACC_BYTES)
		return MDBCU_EUNSUPPORTED;
	sp.kmin = ck.imin;
	sp.range = (uint32_t)range;
	sp.ngroups = (uint32_t)ngroups;

	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	unsigned int *d_table, *d_flags;
	unsigned long long *d_nrows;
	MDB_TRY(tmp.alloc(&d_table, (size_t)range));
	MDB_TRY(tmp.alloc(&d_flags, 1));
	MDB_TRY(tmp.alloc(&d_nrows, 1));
	// the accumulators of all groups in ONE block [joined rows u32 x G, padded | acc i64 x na x G | non-NULL counts u64 x na x G]:
	// it is what a distributed plan exchanges
	const size_t rows_bytes = ((size_t)ngroups * sizeof(unsigned int) + 7) & ~(size_t)7;
	const size_t block_bytes = rows_bytes + 2 * (size_t)na * ngroups * 8;
	unsigned char *acc_block;
	MDB_TRY(tmp.alloc(&acc_block, block_bytes));
	sp.g_rows = reinterpret_cast<unsigned int*>(acc_block);
	sp.g_acc = reinterpret_cast<long long*>(acc_block + rows_bytes);
	sp.g_nn = reinterpret_cast<unsigned long long*>(acc_block + rows_bytes + (size_t)na * ngroups * 8);
	CUDA_TRY(ctx, cudaMemsetAsync(d_table, 0xff, range * sizeof(unsigned int), ctx->stream)); // 0xffffffff: (uint16_t) = ST_EMPTY
	CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, sizeof(unsigned int), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(d_nrows, 0, sizeof(unsigned long long), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(sp.g_rows, 0, ngroups * sizeof(unsigned int), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(sp.g_nn, 0, (size_t)na * ngroups * sizeof(unsigned long long), ctx->stream));
	const int naggs_real = sp.naggs;
	sp.naggs = na;
	MDB_LAUNCH(ctx, k_star_init_acc, 8, 256, 0, sp);

	// ---- build: dimension -> direct table (and the checks that decide whether this path may answer at all)
	clock.begin(2);
	MDB_LAUNCH(ctx, k_star_build, std::max(1, (int)std::min<uint64_t>(ctx->num_sms * 4, (D->n_slots + 255) / 256)), 256, 0,
			(const int64_t*)ck.data, st_all_present(D, dk) ? (const uint32_t*)nullptr : (const uint32_t*)ck.present, (const int64_t*)cg.data,
			st_all_present(D, gc) ? (const uint32_t*)nullptr : (const uint32_t*)cg.present, (uint64_t)D->n_slots, (long long)ck.imin, (uint32_t)range,
			(long long)cg.imin, (uint32_t)ngroups, d_table, d_flags);
	if (W > 1) {
		// [table | flags] of every rank -> every rank; the merged table is the whole dimension (a key present in two shards
		// is a duplicate), the merged flags make all ranks take the same decision
		unsigned int *mine, *all;
		const size_t words = (size_t)range + 1;
		MDB_TRY(tmp.alloc(&mine, words));
		MDB_TRY(tmp.alloc(&all, words * W));
		CUDA_TRY(ctx, cudaMemcpyAsync(mine, d_table, range * sizeof(unsigned int), cudaMemcpyDeviceToDevice, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(mine + range, d_flags, sizeof(unsigned int), cudaMemcpyDeviceToDevice, ctx->stream));
		clock.begin(6);
		MDB_TRY(mdb_comm_allgather_bytes(ctx, mine, all, words * sizeof(unsigned int)));
		ctx->stats.exchange_bytes += (uint64_t)(W - 1) * words * sizeof(unsigned int);
		clock.begin(2);
		MDB_LAUNCH(ctx, k_star_merge_tables, std::max(1, (int)((range + 255) / 256)), 256, 0, (const unsigned int*)all, W, (uint32_t)range,
				d_table, d_flags);
	}
	uint64_t flags64 = 0;
	{
		// (4-byte flag word read through the 8-byte helper: the next word belongs to d_nrows' allocation or padding)
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, d_flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		flags64 = ctx->h_scalar[0] & 0xffffffffull;
	}
	if (flags64)
		return MDBCU_EUNSUPPORTED; // duplicate keys / NULL groups: not a star dimension, the general operators answer

	// ---- probe: stream the fact table once
	clock.begin(3);
	const size_t smem = (((size_t)range * 2 + 15) & ~(size_t)15) + (size_t)ngroups * (na * 12 + 4);
	KernelTimer ktimer;
	ktimer.start(ctx->stream);
	if (sp.ncols == 1)
		st_launch_probe<1>(ctx, ctx->num_sms, smem, na, sp, d_table, F->n_slots);
	else if (sp.ncols == 2)
		st_launch_probe<2>(ctx, ctx->num_sms, smem, na, sp, d_table, F->n_slots);
	else
		st_launch_probe<3>(ctx, ctx->num_sms, smem, na, sp, d_table, F->n_slots);
	ktimer.stop(ctx->stream);
	if (W > 1) {
		unsigned char *all;
		MDB_TRY(tmp.alloc(&all, block_bytes * W));
		clock.begin(6);
		MDB_TRY(mdb_comm_allgather_bytes(ctx, acc_block, all, block_bytes));
		ctx->stats.exchange_bytes += (uint64_t)(W - 1) * block_bytes;
		// rank order, starting from rank 0's block: the same fold on every run
		MDB_LAUNCH(ctx, k_star_merge_acc, 8, 256, 0, sp, (const unsigned char*)all, W, rows_bytes, block_bytes);
	}

	// ---- emit
	clock.begin(4);
	MDB_TRY(mdb_result_alloc(ctx, plan, res, ngroups, false));
	out.nout = plan->n_out;
	out.nrows = d_nrows;
	out.gmin = cg.imin;
	for (int o = 0; o < plan->n_out; o++) {
		out.cells[o] = res->cols[o].cells;
		out.nulls[o] = res->cols[o].nulls;
	}
	MDB_LAUNCH(ctx, k_star_emit, 8, 256, 0, sp, out);
	cudaError_t e = cudaGetLastError();
	uint64_t nrows = 0;
	MDB_TRY(mdb_read_u64(ctx, (const uint64_t*)d_nrows, &nrows));
	clock.finish();
	const float kms = ktimer.ms();
	if (e != cudaSuccess)
		return mdb_fail(ctx, MDBCU_ECUDA, "star join launch failed: %s", cudaGetErrorString(e));
	res->nrows = (W > 1 && ctx->rank != 0) ? 0 : nrows; // distributed: rank 0 returns the groups
	(void)naggs_real;

	ctx->stats.path = MDBCU_PATH_DIRECT_STAR;
	ctx->stats.algorithmic_bytes = 8ull * F->n_slots * sp.ncols + 16ull * D->n_slots + 8ull * plan->n_out * nrows;
	ctx->stats.dominant_ms = kms;
	ctx->stats.dominant_bytes = 8ull * F->n_slots * sp.ncols;
	return MDBCU_OK;
}
