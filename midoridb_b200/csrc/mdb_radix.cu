// mdb_radix.cu - radix-partitioned join + GROUP BY join key + COUNT(*)   (the README query; BASELINE configs 1 and 3)
//
//   SELECT k, COUNT(*) FROM A INNER JOIN B ON A.k = B.k GROUP BY k
//
// replaces _join_nested_loop_tbl2tbl (src/engine/executor_select.c:1076) + proc_groupby_clause (:1526):
// result(k) = cntA[k] * cntB[k]; no candidate pair and no joined row is ever materialised.
//
//   pass 1  k_radix_partition   streams the 8-byte keys ONCE, writes 2-byte remainders into per-partition
//                               512-byte chunks (partition = high bits of key - kmin, <= 4096 partitions);
//   (dir)   k_radix_dir_*       counting sort of chunk ids by partition (a few microseconds);
//   pass 2  k_radix_joincount   per partition: both sides' remainders -> packed 4- or 8-bit counters in shared
//                               memory, checksum against the number of remainders, multiply, emit groups.
//
// Algorithmic bytes: 8|A| + 8|B| in, 16 G out.  Extra traffic of this design: 2 bytes per key
// written + read back (the remainders).  HBM-bound integer work; no tensor cores (nothing here is a contraction).
#include "mdb_common.cuh"

#include <string.h>
#include <algorithm>

#include "mdb_radix_types.cuh"

static bool col_all_present(const mdbcu_table *t, int col)
{
	return t->all_live && !t->cols[col].has_nulls;
}

#include "mdb_radix_pass1.cuh"


// exclusive scan of the per-partition chunk counts (single block, nparts <= 4096)
__global__ void k_radix_dir_scan(RJSide s, int nparts)
{
	__shared__ uint64_t warp_tot[33];
	uint64_t carry = 0;
	for (int base = 0; base < nparts + 1; base += blockDim.x) {
		int i = base + threadIdx.x;
		uint64_t v = i < nparts ? s.dst[s.self].dir_cnt[i] : 0;
		uint64_t incl = v;
		int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		for (int o = 1; o < 32; o <<= 1) {
			uint64_t n = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o)
				incl += n;
		}
		if (lane == 31)
			warp_tot[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint64_t w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0, wi = w;
			for (int o = 1; o < 32; o <<= 1) {
				uint64_t n = __shfl_up_sync(0xffffffffu, wi, o);
				if (lane >= o)
					wi += n;
			}
			warp_tot[lane] = wi - w;
			if (lane == 31)
				warp_tot[32] = wi;
		}
		__syncthreads();
		if (i < nparts + 1)
			s.dir_off[i] = carry + warp_tot[warp] + incl - v;
		carry += warp_tot[32];
		__syncthreads();
	}
}

#include "mdb_radix_pass2.cuh"
#include "mdb_radix_dist.cuh"

// chunk-id reservation sizes: a CTA needs at most one fresh chunk per owned partition at start-up
static void rj_id_policy(int world, uint32_t *batch, uint32_t *low)
{
	*batch = world == 1 ? 8192u : std::max(1024u, 2u * RJ_MAX_PART / (uint32_t)world);
	*low = *batch / 4;
}

// chunks one owner's pool must hold: data chunks (with slack for imbalance) + one ragged chunk per
// (source CTA, owned partition) + ids abandoned at refills + one reserve per (source CTA, owner)
static uint64_t rj_pool_chunks(uint64_t rows_for_owner, int world, int grid)
{
	uint32_t batch, low;
	rj_id_policy(world, &batch, &low);
	uint64_t chunks = rows_for_owner / RJ_CHUNK + (uint64_t)grid * RJ_MAX_PART;
	chunks += chunks / 2 + (uint64_t)world * grid * batch + 1024;
	return chunks;
}

// byte layout of one side inside an exchange arena (identical on every rank)
struct RJArenaLayout {
	size_t pool, chunk_part, chunk_entries, dir_cnt, pool_next, bytes;
};

static RJArenaLayout rj_arena_layout(uint64_t chunks)
{
	auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
	RJArenaLayout l;
	l.pool = 0;
	l.chunk_part = up(l.pool + chunks * RJ_CHUNK * sizeof(uint16_t));
	l.chunk_entries = up(l.chunk_part + chunks * sizeof(uint16_t));
	l.dir_cnt = up(l.chunk_entries + chunks * sizeof(uint16_t));
	l.pool_next = up(l.dir_cnt + (RJ_MAX_PART + 1) * sizeof(uint32_t));
	l.bytes = up(l.pool_next + 256);
	return l;
}

static void rj_target_at(RJTarget *t, void *base, const RJArenaLayout &l)
{
	char *b = (char*)base;
	t->pool = (uint16_t*)(b + l.pool);
	t->chunk_part = (uint16_t*)(b + l.chunk_part);
	t->chunk_entries = (uint16_t*)(b + l.chunk_entries);
	t->dir_cnt = (uint32_t*)(b + l.dir_cnt);
	t->pool_next = (uint32_t*)(b + l.pool_next);
}

// arena_bases == nullptr: all chunks stay on this GPU (single-GPU plan, or NCCL exchange afterwards);
// otherwise dst[r] points into rank r's arena (at byte offset arena_off) and pass 1 writes over NVLink
static int rj_side_setup(mdbcu_ctx *ctx, DevTemp &tmp, RJSide *s, const mdbcu_table *t, int col, int grid, uint64_t chunks,
		void *const *arena_bases, size_t arena_off)
{
	memset(s, 0, sizeof(*s));
	s->keys = t->cols[col].data;
	s->present = col_all_present(t, col) ? nullptr : t->cols[col].present;
	s->n = t->n_slots;
	if (chunks >= (1ull << 27))
		return MDBCU_EUNSUPPORTED;
	s->pool_chunks = (uint32_t)chunks;
	s->world = arena_bases ? ctx->world : 1;
	s->self = arena_bases ? ctx->rank : 0;
	rj_id_policy(s->world, &s->id_batch, &s->id_low);
	static const uint32_t hints = getenv("MDBCU_P1_HINTS") ? (uint32_t)atoi(getenv("MDBCU_P1_HINTS")) : RJ_HINT_DEFAULT;
	s->hints = hints;
	RJTarget &own = s->dst[s->self];
	if (arena_bases) {
		const RJArenaLayout l = rj_arena_layout(chunks);
		for (int r = 0; r < ctx->world; r++)
			rj_target_at(&s->dst[r], (char*)arena_bases[r] + arena_off, l);
	} else {
		MDB_TRY(tmp.alloc(&own.pool, chunks * RJ_CHUNK));
		MDB_TRY(tmp.alloc(&own.pool_next, 1));
		MDB_TRY(tmp.alloc(&own.chunk_part, chunks));
		MDB_TRY(tmp.alloc(&own.chunk_entries, chunks));
		MDB_TRY(tmp.alloc(&own.dir_cnt, RJ_MAX_PART + 1));
	}
	s->pool = own.pool;
	MDB_TRY(tmp.alloc(&s->dir_fill, RJ_MAX_PART + 1));
	MDB_TRY(tmp.alloc(&s->dir_off, RJ_MAX_PART + 2));
	MDB_TRY(tmp.alloc(&s->dir, chunks));
	CUDA_TRY(ctx, cudaMemsetAsync(own.pool_next, 0, sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(own.chunk_part, 0xff, chunks * sizeof(uint16_t), ctx->stream)); // 0xffff = never allocated
	CUDA_TRY(ctx, cudaMemsetAsync(own.dir_cnt, 0, (RJ_MAX_PART + 1) * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(s->dir_fill, 0, (RJ_MAX_PART + 1) * sizeof(uint32_t), ctx->stream));
	return MDBCU_OK;
}

static void launch_partition(mdbcu_ctx *ctx, int grid, const RJSide &s, const RJParams &pr)
{
	if (s.present)
		MDB_LAUNCH(ctx, k_radix_partition<true>, grid, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
	else if (s.all_in_range && pr.range <= 0xffffffffull && ((uintptr_t)s.keys & 31u) == 0)
		MDB_LAUNCH(ctx, k_radix_partition_fast, grid, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
	else
		MDB_LAUNCH(ctx, k_radix_partition<false>, grid, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
}

int mdb_select_radix_joincount(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (plan->n_tables != 2 || plan->n_joins != 1 || plan->joins[0].cross || plan->n_pred != 0 || plan->n_group != 1 ||
			plan->n_out < 1 || plan->n_out > 4)
		return MDBCU_EUNSUPPORTED;
	const bool dist = (plan->flags & MDBCU_PLAN_DISTRIBUTED) != 0;
	if (dist && !ctx->nccl_comm)
		return mdb_fail(ctx, MDBCU_EERROR, "MDBCU_PLAN_DISTRIBUTED needs mdbcu_comm_init first");
	const mdbcu_join &jn = plan->joins[0];
	if (jn.left.tbl != 0 || jn.right.tbl != 1)
		return MDBCU_EUNSUPPORTED;
	const mdbcu_table *ta = plan->tables[0], *tb = plan->tables[1];
	if (jn.left.col < 0 || jn.left.col >= ta->ncols || jn.right.col < 0 || jn.right.col >= tb->ncols)
		return MDBCU_EUNSUPPORTED;
	DevColumn ca = ta->cols[jn.left.col], cb = tb->cols[jn.right.col]; // copies: distributed plans overwrite the bounds
	auto intlike = [](int type) { return type == MDBCU_CT_INTEGER || type == MDBCU_CT_DATE || type == MDBCU_CT_DATETIME; };
	if (!intlike(ca.type) || !intlike(cb.type) || !ca.stats_ok || !cb.stats_ok)
		return MDBCU_EUNSUPPORTED;
	if (dist) {
		// every rank must take the same decisions: use the bounds over ALL shards
		if (!ca.gstats_ok || !cb.gstats_ok)
			return mdb_fail(ctx, MDBCU_EERROR, "distributed plan: call mdbcu_table_sync_stats on every sharded table first");
		ca.imin = ca.gmin;
		ca.imax = ca.gmax;
		cb.imin = cb.gmin;
		cb.imax = cb.gmax;
	}
	auto is_key = [&](const mdbcu_colref &r) {
		return (r.tbl == 0 && r.col == jn.left.col) || (r.tbl == 1 && r.col == jn.right.col);
	};
	if (!is_key(plan->group[0]))
		return MDBCU_EUNSUPPORTED;
	for (int o = 0; o < plan->n_out; o++) {
		if (plan->out[o].kind == MDBCU_OUT_COUNT_STAR)
			continue;
		if (plan->out[o].kind != MDBCU_OUT_COLUMN || !is_key(plan->out[o].ref))
			return MDBCU_EUNSUPPORTED;
	}
	if (!dist && ta->n_slots + tb->n_slots < (1ull << 20))
		return MDBCU_EUNSUPPORTED; // small inputs: general operators (they also return the reference's row order)

	// only keys inside both columns' [min, max] (zone-map statistics kept by the mirror) can ever match
	long long kmin = std::max(ca.imin, cb.imin), kmax = std::min(ca.imax, cb.imax);
	if (ca.imin > ca.imax || cb.imin > cb.imax || kmin > kmax) {
		ctx->stats.path = MDBCU_PATH_RADIX_JOINCOUNT;
		return mdb_result_alloc(ctx, plan, res, 0, false);
	}
	unsigned long long range = (unsigned long long)kmax - (unsigned long long)kmin + 1ull;
	if (range == 0 || range > ((unsigned long long)RJ_MAX_PART << RJ_MAX_SHIFT) || range < 4096)
		return MDBCU_EUNSUPPORTED;
	{
		// when the two columns cover almost the same interval, partition over the UNION of the intervals instead:
		// every key of both sides is then in range and the hot loop needs no per-key range test
		long long umin = std::min(ca.imin, cb.imin), umax = std::max(ca.imax, cb.imax);
		unsigned long long urange = (unsigned long long)umax - (unsigned long long)umin + 1ull;
		if (urange != 0 && urange <= ((unsigned long long)RJ_MAX_PART << RJ_MAX_SHIFT) && urange <= range + range / 4) {
			kmin = umin;
			kmax = umax;
			range = urange;
		}
	}
	int bits = 0;
	while ((1ull << bits) < range)
		bits++;
	int shift = std::max(0, bits - 12);
	int nparts = (int)((range + (1ull << shift) - 1) >> shift);

	ctx->stats.path = MDBCU_PATH_RADIX_JOINCOUNT;
	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	const int grid1 = ctx->num_sms;
	RJSide sa, sb;
	RJParams pr;
	// multi-GPU plans: pass 1 writes every chunk straight into the owner GPU's arena over NVLink (CUDA IPC);
	// MDBCU_EXCHANGE=nccl selects the staged variant instead (partition locally, then a grouped send/recv)
	static const bool use_nccl_exchange = getenv("MDBCU_EXCHANGE") && strcmp(getenv("MDBCU_EXCHANGE"), "nccl") == 0;
	const bool p2p = dist && !use_nccl_exchange;
	if (p2p) {
		if (!ta->global_slots || !tb->global_slots)
			return mdb_fail(ctx, MDBCU_EERROR, "distributed plan: call mdbcu_table_sync_stats on every sharded table first");
		const uint64_t ca_chunks = rj_pool_chunks(ta->global_slots / ctx->world + 1, ctx->world, grid1);
		const uint64_t cb_chunks = rj_pool_chunks(tb->global_slots / ctx->world + 1, ctx->world, grid1);
		const RJArenaLayout la = rj_arena_layout(ca_chunks), lb = rj_arena_layout(cb_chunks);
		void *bases[MDB_MAX_RANKS];
		MDB_TRY(mdb_comm_arena(ctx, la.bytes + lb.bytes, bases));
		MDB_TRY(rj_side_setup(ctx, tmp, &sa, ta, jn.left.col, grid1, ca_chunks, bases, 0));
		MDB_TRY(rj_side_setup(ctx, tmp, &sb, tb, jn.right.col, grid1, cb_chunks, bases, la.bytes));
	} else {
		MDB_TRY(rj_side_setup(ctx, tmp, &sa, ta, jn.left.col, grid1, rj_pool_chunks(ta->n_slots, 1, grid1), nullptr, 0));
		MDB_TRY(rj_side_setup(ctx, tmp, &sb, tb, jn.right.col, grid1, rj_pool_chunks(tb->n_slots, 1, grid1), nullptr, 0));
	}
	sa.all_in_range = ca.imin >= kmin && ca.imax <= kmax;
	sb.all_in_range = cb.imin >= kmin && cb.imax <= kmax;
	pr.kmin = kmin;
	pr.range = range;
	pr.shift = shift;
	pr.mask = (1u << shift) - 1u;
	pr.nparts = nparts;
	pr.part_first = 0;
	pr.part_end = nparts;
	if (dist) {
		pr.part_first = (int)((uint64_t)ctx->rank * nparts / ctx->world);
		pr.part_end = (int)((uint64_t)(ctx->rank + 1) * nparts / ctx->world);
	}
	uint32_t *d_flags; // [0] error flags, [1] partition counter
	unsigned long long *d_cursor;
	MDB_TRY(tmp.alloc(&d_flags, 2));
	MDB_TRY(tmp.alloc(&d_cursor, 1));
	CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
	pr.error_flag = d_flags;

	// upper bound of groups this rank can emit: one per key of the partitions it owns
	uint64_t cap_groups = std::min<uint64_t>((uint64_t)(pr.part_end - pr.part_first) << shift, range);
	if (!dist)
		cap_groups = std::min<uint64_t>(cap_groups, std::min<uint64_t>(ta->n_slots, tb->n_slots));
	MDB_TRY(mdb_result_alloc(ctx, plan, res, 0, false));
	RJOut out;
	memset(&out, 0, sizeof(out));
	out.nout = plan->n_out;
	out.cursor = d_cursor;
	out.cap = cap_groups;
	for (int o = 0; o < plan->n_out; o++) {
		mdb_free(ctx, res->cols[o].cells);
		mdb_free(ctx, res->cols[o].nulls);
		res->cols[o].nulls = nullptr; // NULL keys never join (executor_select.c:716-738): no NULL cells in this result
		res->cols[o].cells = nullptr;
		MDB_TRY(mdb_alloc(ctx, &res->cols[o].cells, cap_groups));
		out.cells[o] = res->cols[o].cells;
		out.is_count[o] = plan->out[o].kind == MDBCU_OUT_COUNT_STAR;
	}

	static bool attr_done = false;
	if (!attr_done) {
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2 * RJ_DESC_CAP * (int)sizeof(RJDesc)));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 2 * RJ_DESC_CAP * (int)sizeof(RJDesc)));
		attr_done = true;
	}

	if (p2p) {
		// every rank's arena must be reset before any peer starts writing into it
		clock.begin(6);
		MDB_TRY(mdb_comm_barrier_or(ctx, d_flags, nullptr));
	}
	clock.begin(1);
	launch_partition(ctx, grid1, sa, pr);
	launch_partition(ctx, grid1, sb, pr);
	if (p2p) {
		// all remote stores have landed once every rank's pass 1 has completed; error flags are shared so that
		// every rank takes the same exit
		clock.begin(6);
		uint32_t any = 0;
		MDB_TRY(mdb_comm_barrier_or(ctx, d_flags, &any));
		if (any)
			return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "distributed radix join: pass 1 failed on some rank (flags %u: "
					"1 = chunk pool exhausted, 4 = extreme skew)", any);
	}
	clock.begin(7);
	MDB_LAUNCH(ctx, k_radix_dir_scan, 1, 1024, 0, sa, nparts);
	MDB_LAUNCH(ctx, k_radix_dir_scan, 1, 1024, 0, sb, nparts);
	MDB_LAUNCH(ctx, k_radix_dir_fill, ctx->num_sms * 4, 256, 0, sa);
	MDB_LAUNCH(ctx, k_radix_dir_fill, ctx->num_sms * 4, 256, 0, sb);
	uint64_t exchanged = 0;
	if (dist && !p2p) {
		clock.begin(6);
		RJSide *both[2] = {&sa, &sb};
		MDB_TRY(rj_exchange(ctx, tmp, &pr, both, &exchanged));
	}
	clock.begin(2);

	// 4-bit counters (two CTAs per SM) when keys are mostly unique per side, 8-bit otherwise; a wrapped
	// counter is detected by the checksum and the pass is repeated one width up before giving up
	const uint64_t D = 1ull << shift;
	bool try4 = std::max(ta->n_slots, tb->n_slots) <= 2 * range;
	uint64_t ngroups = 0;
	uint32_t flags = 0;
	for (int attempt = try4 ? 0 : 1; attempt < 2; attempt++) {
		const int bitsw = attempt == 0 ? 4 : 8;
		const size_t smem2 = 2 * (size_t)std::max<uint64_t>(1, D * bitsw / 32) * sizeof(uint32_t) + 2 * RJ_DESC_CAP * sizeof(RJDesc);
		const int grid2 = std::max(1, std::min(pr.part_end - pr.part_first, ctx->num_sms * (bitsw == 4 ? 2 : 1)));
		if (bitsw == 4)
			MDB_LAUNCH(ctx, (k_radix_joincount<4, 512>), grid2, 512, smem2, sa, sb, pr, out, d_flags + 1);
		else
			MDB_LAUNCH(ctx, (k_radix_joincount<8, 1024>), grid2, 1024, smem2, sa, sb, pr, out, d_flags + 1);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			return mdb_fail(ctx, MDBCU_ECUDA, "radix join launch failed: %s", cudaGetErrorString(e));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, d_cursor, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar + 1, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		ngroups = ctx->h_scalar[0];
		flags = (uint32_t)(ctx->h_scalar[1] & 0xffffffffu);
		if (flags != RJ_ERR_COUNTER || attempt == 1)
			break;
		// a 4-bit counter wrapped: repeat pass 2 with 8-bit counters (the partitioned remainders are still valid)
		CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), ctx->stream));
		CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
	}
	clock.finish();
	if (flags || ngroups > cap_groups) {
		if (dist)
			return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "distributed radix join: key multiplicity or skew beyond the counter width "
					"(flags %u); the general operators are single-GPU only", flags);
		return MDBCU_EUNSUPPORTED; // heavy duplicates / skew / pool exhaustion: the general operators redo the query
	}
	res->nrows = ngroups;
	ctx->stats.exchange_bytes = exchanged;

	ctx->stats.algorithmic_bytes = 8ull * (ta->n_slots + tb->n_slots) + 8ull * plan->n_out * ngroups;
	ctx->stats.dominant_ms = ctx->stats.phase_ms[1] + ctx->stats.phase_ms[2];
	ctx->stats.dominant_bytes = ctx->stats.algorithmic_bytes;
	return MDBCU_OK;
}
