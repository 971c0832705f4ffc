// mdb_radix.cu - radix-partitioned join + GROUP BY join key + COUNT(*)   (the README query; BASELINE configs 1 and 3)
//
//   SELECT k, COUNT(*) FROM A INNER JOIN B ON A.k = B.k GROUP BY k
//
// replaces _join_nested_loop_tbl2tbl (src/engine/executor_select.c:1076) + proc_groupby_clause (:1526):
// result(k) = cntA[k] * cntB[k]; no candidate pair and no joined row is ever materialised.
//
//   pass 1  k_radix_partition   streams the 8-byte keys ONCE and appends 2-byte remainders to per-partition
//                               streams (partition = (key - kmin) / width, <= 4096 partitions of <= 65536 key values);
//   (barrier)                   multi-GPU plans only: the streams live in an arena every peer has mapped; pass 2 of the rank
//                               that owns a partition reads all ranks' streams of it over NVLink (mdb_radix_dist.cuh);
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
#include "mdb_radix_pass2.cuh"
#include "mdb_radix_dist.cuh"
#include "mdb_radix_sorted.cuh"

// dynamic shared memory the push kernel asks for: it is not used, it makes a push CTA own its SM, so that the push takes
// push_sms SMs and pass 1 of the other side gets all the others
#define RJ_SHIP_SMEM (160 * 1024)

// entries one partition's main stream can hold: twice the average (uniform keys fill half of it; a partition that
// receives more than twice its share sets RJ_ERR_STREAM and the general operators take over)
static uint32_t rj_stream_cap(uint64_t rows, int nparts)
{
	uint64_t cap = 2 * (rows / (uint64_t)nparts) + 2048;
	cap = (cap + 63) & ~63ull; // whole 128-byte lines
	return (uint32_t)std::min<uint64_t>(cap, 0x7fffffc0ull);
}

// tail sectors (16 entries each) a partition has room for: one per CTA for the partial staging rows at the end of pass 1, one
// per key that found its staging row full (about 1 in 1000 on uniform keys): 1/16 of the main capacity on top
static uint32_t rj_tail_cap(int grid, uint32_t cap)
{
	return ((uint32_t)grid * RJ_FLUSH + cap / 16 + 15u) & ~15u;
}

// cursors: RJ_MAX_PART * RJ_CUR_STRIDE zeroed words inside the query's control block
// `region` != nullptr (multi-GPU plans): the streams and the cursors live in this rank's exchange arena, at the offsets of
// rj_region_layout, where the peers' pass 2 reads them over NVLink; otherwise they are query temporaries.
static int rj_side_setup(mdbcu_ctx *ctx, DevTemp &tmp, RJSide *s, const mdbcu_table *t, int col, int grid, int nparts, uint32_t cap,
		uint32_t *cursors, char *region = nullptr)
{
	memset(s, 0, sizeof(*s));
	s->keys = t->cols[col].data;
	s->present = col_all_present(t, col) ? nullptr : t->cols[col].present;
	s->n = t->n_slots;
	static const uint32_t hints = getenv("MDBCU_P1_HINTS") ? (uint32_t)atoi(getenv("MDBCU_P1_HINTS")) : RJ_HINT_DEFAULT;
	s->hints = hints;
	s->cap = cap;
	s->tail_cap = rj_tail_cap(grid, cap);
	if ((uint64_t)nparts * cap >= (1ull << 40))
		return MDBCU_EUNSUPPORTED;
	if (region) {
		const RJRegionLayout l = rj_region_layout((uint32_t)nparts, cap, s->tail_cap);
		s->stream = (uint16_t*)(region + l.main);
		s->tail = (uint16_t*)(region + l.tail);
		s->cursor = (uint32_t*)(region + l.cursor);
		CUDA_TRY(ctx, cudaMemsetAsync(s->cursor, 0, (size_t)RJ_MAX_PART * RJ_CUR_STRIDE * sizeof(uint32_t), ctx->stream));
	} else {
		MDB_TRY(tmp.alloc(&s->stream, (size_t)nparts * cap));
		MDB_TRY(tmp.alloc(&s->tail, (size_t)nparts * s->tail_cap));
		s->cursor = cursors;
	}
	s->tail_cursor = s->cursor + 1;
	return MDBCU_OK;
}

// Is the key column sorted (non-decreasing, no NULLs, no tombstones)?  Checked once per table version: a cheap look at
// the first 2^20 rows settles it for unordered data, only a column that passes is read completely.
static int rj_column_sorted(mdbcu_ctx *ctx, DevTemp &tmp, const mdbcu_table *t, int col, bool *sorted)
{
	DevColumn &c = const_cast<mdbcu_table*>(t)->cols[col];
	if (c.sorted_version == t->version) {
		*sorted = c.sorted;
		return MDBCU_OK;
	}
	*sorted = false;
	if (col_all_present(t, col) && t->n_slots >= 2 && t->n_slots < (1ull << 32) && ((uintptr_t)c.data & 31u) == 0) {
		uint32_t *d_desc;
		MDB_TRY(tmp.alloc(&d_desc, 1));
		CUDA_TRY(ctx, cudaMemsetAsync(d_desc, 0, sizeof(uint32_t), ctx->stream));
		const uint64_t sample = std::min<uint64_t>(t->n_slots, 1ull << 20);
		for (uint64_t rows : {sample, (uint64_t)t->n_slots}) {
			MDB_LAUNCH(ctx, k_is_sorted, ctx->num_sms * 8, 256, 0, c.data, rows, d_desc);
			CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, d_desc, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
			CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
			*sorted = (uint32_t)(ctx->h_scalar[0] & 0xffffffffu) == 0;
			if (!*sorted || rows == t->n_slots)
				break;
		}
	}
	c.sorted_version = t->version;
	c.sorted = *sorted;
	return MDBCU_OK;
}

// pass 2 input of a single-GPU plan: the local streams
static void rj_runs_local(RJRuns *r, const RJSide &s)
{
	memset(r, 0, sizeof(*r));
	r->nsrc = 1;
	r->cap = s.cap;
	r->tail_cap = s.tail_cap;
	r->stream[0] = s.stream;
	r->tail[0] = s.tail;
	r->cursor[0] = s.cursor;
	r->tail_cursor[0] = s.tail_cursor;
	r->first[0] = 0;
	r->cur_stride[0] = RJ_CUR_STRIDE;
}

// grid = SMs to use: RJ_SPLIT CTAs are launched for each (together they fill one SM's shared memory and registers)
static void launch_partition(mdbcu_ctx *ctx, int grid, const RJSide &s, const RJParams &pr)
{
	if (s.present)
		MDB_LAUNCH(ctx, k_radix_partition<true>, grid * RJ_SPLIT, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
	else if (s.all_in_range && pr.range <= 0xffffffffull && ((uintptr_t)s.keys & 31u) == 0)
	{
		static const bool no_w16 = getenv("MDBCU_P1_NO_W16") != nullptr; // (A/B measurements)
		if (pr.width == 65536u && !no_w16)
			MDB_LAUNCH(ctx, k_radix_partition_fast<true>, grid * RJ_SPLIT, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
		else
			MDB_LAUNCH(ctx, k_radix_partition_fast<false>, grid * RJ_SPLIT, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
	}
	else
		MDB_LAUNCH(ctx, k_radix_partition<false>, grid * RJ_SPLIT, RJ_P1_THREADS, sizeof(RJP1Smem), s, pr);
}

int mdb_select_radix_joincount(mdbcu_ctx *ctx, const mdbcu_plan *plan, mdbcu_result *res)
{
	if (plan->n_tables != 2 || plan->n_joins != 1 || plan->joins[0].cross || plan->n_pred != 0 || plan->n_group != 1 ||
			plan->n_out < 1 || plan->n_out > 4)
		return MDBCU_EUNSUPPORTED;
	const bool dist = (plan->flags & MDBCU_PLAN_DISTRIBUTED) != 0;
	if (dist && !mdb_comm_ready(ctx))
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
		if (!ca.gstats_ok || !cb.gstats_ok || !ta->global_slots || !tb->global_slots)
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
	if (ta->n_slots >= (1ull << 32) || tb->n_slots >= (1ull << 32))
		return MDBCU_EUNSUPPORTED; // stream positions and per-partition totals are 32-bit

	// only keys inside both columns' [min, max] (zone-map statistics kept by the mirror) can ever match
	long long kmin = std::max(ca.imin, cb.imin), kmax = std::min(ca.imax, cb.imax);
	if (ca.imin > ca.imax || cb.imin > cb.imax || kmin > kmax) {
		ctx->stats.path = MDBCU_PATH_RADIX_JOINCOUNT;
		return mdb_result_alloc(ctx, plan, res, 0, false);
	}
	unsigned long long range = (unsigned long long)kmax - (unsigned long long)kmin + 1ull;
	if (range == 0 || range > ((unsigned long long)RJ_MAX_PART << RJ_MAX_SHIFT) || range < 4096) {
		// a key range beyond 4096 partitions x 2^16 remainders: large inputs go to the direct-count path (one counter per key
		// value, up to 2^30 values) instead of the general operators, which would build a hash table of every row
		if (range > ((unsigned long long)RJ_MAX_PART << RJ_MAX_SHIFT))
			ctx->radix_gave_up = true;
		return MDBCU_EUNSUPPORTED;
	}
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
	// partitions of `width` consecutive key values (mdb_radix_types.cuh): all 4096 staging rows are used whatever the range
	const uint32_t width = (uint32_t)std::max<unsigned long long>(2, (range + RJ_MAX_PART - 1) / RJ_MAX_PART);
	const int nparts = (int)((range + width - 1) / width);

	ctx->stats.path = MDBCU_PATH_RADIX_JOINCOUNT;
	PhaseClock clock(ctx);
	DevTemp tmp(ctx);
	const int grid1 = ctx->num_sms;
	const int W = dist ? ctx->world : 1, me = dist ? ctx->rank : 0;
	// Every decision up to here depends on the plan and on GLOBAL statistics only: all ranks of a distributed plan are here
	// together.  From now on a rank that fails on its own (out of memory ...) must not leave its peers spinning in the
	// cross-rank barrier: the guard publishes MDB_PEER_ABORT in place of the barrier this rank will not reach.
	struct AbortGuard {
		mdbcu_ctx *ctx;
		bool armed;
		~AbortGuard()
		{
			if (armed)
				mdb_comm_arena_abort(ctx);
		}
	} guard = {ctx, false};
	const uint32_t arena_query = (dist && W > 1) ? ctx->arena_queries++ : 0u; // (counted before anything can fail: the ranks stay in step)
	guard.armed = dist && W > 1 && ctx->arena_local != nullptr;
	RJSide sa, sb;
	RJParams pr;
	// stream capacities must be identical on every rank (the arena slots mirror the local layout)
	const uint32_t cap_a = rj_stream_cap(dist ? (ta->global_slots + W - 1) / W : ta->n_slots, nparts);
	const uint32_t cap_b = rj_stream_cap(dist ? (tb->global_slots + W - 1) / W : tb->n_slots, nparts);
	// a side that is already sorted skips pass 1 (single-GPU plans; sorted shards of a distributed plan are not handled)
	bool sorted_a = false, sorted_b = false;
	if (!dist) {
		MDB_TRY(rj_column_sorted(ctx, tmp, ta, jn.left.col, &sorted_a));
		MDB_TRY(rj_column_sorted(ctx, tmp, tb, jn.right.col, &sorted_b));
	}
	// control block of the query: ONE allocation, ONE memset, ONE read-back
	//   bytes  0..15  groups emitted, bytes pushed to peers (u64 each)      16..23  error flags, partition counter (u32 each)
	//   bytes 32..63  every rank's error flags after the exchange (u32 x 8)  64..95   every rank's error flags after pass 2
	//   bytes 128..   cursors of side A, then of side B
	const size_t cursor_words = (size_t)RJ_MAX_PART * RJ_CUR_STRIDE;
	const size_t ctl_head = 32;
	uint32_t *ctl;
	MDB_TRY(tmp.alloc(&ctl, ctl_head + 2 * cursor_words));
	CUDA_TRY(ctx, cudaMemsetAsync(ctl, 0, (ctl_head + 2 * cursor_words) * sizeof(uint32_t), ctx->stream));
	memset(&sa, 0, sizeof(sa));
	memset(&sb, 0, sizeof(sb));
	// Multi-GPU plans: every rank partitions ITS shard into streams inside its own exchange arena (one cudaMalloc block per
	// rank, mapped into all peers: CUDA IPC between processes, plain pointers inside one process); after ONE cross-rank
	// barrier the owner of a partition has side A's runs in local slots (pushed by the peers' k_radix_ship on a few SMs while
	// pass 1 of side B still ran) and reads side B's runs out of all W arenas in pass 2 itself - every byte crosses
	// NVLink under cover of some computation (mdb_radix_dist.cuh).  Alternate queries use alternate halves of the arena: a peer that is still reading this query's streams is never overtaken by the
	// next query's pass 1 (the barrier after pass 2 orders query k before query k + 2).
	void *bases[MDB_MAX_RANKS] = {};
	size_t half_off = 0, side_b_off = 0, slots_off = 0;
	const bool multi_gpu = dist && W > 1;
	const uint32_t pown = (uint32_t)((nparts + W - 1) / W) + 1; // partitions a rank owns at most
	if (multi_gpu) {
		// one half of the arena: [side A's streams | side B's streams | W slots that receive the peers' side-A runs]
		const uint32_t tc_a = rj_tail_cap(grid1, cap_a), tc_b = rj_tail_cap(grid1, cap_b);
		const RJRegionLayout la = rj_region_layout((uint32_t)nparts, cap_a, tc_a), lb = rj_region_layout((uint32_t)nparts, cap_b, tc_b);
		const RJSlotLayout ls = rj_slot_layout(pown, cap_a, tc_a);
		const size_t half_bytes = la.bytes + lb.bytes + ls.bytes * (size_t)W;
		MDB_TRY(mdb_comm_arena(ctx, 2 * half_bytes, bases));
		guard.armed = true; // (the arena exists from here on, if it did not before)
		half_off = (arena_query & 1u) ? half_bytes : 0;
		side_b_off = la.bytes;
		slots_off = la.bytes + lb.bytes;
	}
	if (!sorted_a)
		MDB_TRY(rj_side_setup(ctx, tmp, &sa, ta, jn.left.col, grid1, nparts, cap_a, ctl + ctl_head,
				multi_gpu ? (char*)bases[me] + half_off : nullptr));
	if (!sorted_b)
		MDB_TRY(rj_side_setup(ctx, tmp, &sb, tb, jn.right.col, grid1, nparts, cap_b, ctl + ctl_head + cursor_words,
				multi_gpu ? (char*)bases[me] + half_off + side_b_off : nullptr));
	sa.all_in_range = ca.imin >= kmin && ca.imax <= kmax;
	sb.all_in_range = cb.imin >= kmin && cb.imax <= kmax;
	pr.kmin = kmin;
	pr.range = range;
	pr.width = width;
	pr.magic = (uint32_t)((1ull << 32) / width);
	pr.nparts = nparts;
	pr.part_first = (int)rj_part_first((uint32_t)me, (uint32_t)nparts, (uint32_t)W);
	pr.part_end = (int)rj_part_first((uint32_t)me + 1u, (uint32_t)nparts, (uint32_t)W);
	unsigned long long *d_cursor = reinterpret_cast<unsigned long long*>(ctl); // [0] groups emitted, [1] bytes read from peers
	uint32_t *d_flags = ctl + 4;                                                // [0] error flags, [1] partition counter
	pr.error_flag = d_flags;
	pr.peer_flags = nullptr;
	pr.n_peer_flags = 0;
	static const int plain_emit = getenv("MDBCU_P2_PLAIN_EMIT") ? atoi(getenv("MDBCU_P2_PLAIN_EMIT")) : 0;
	pr.plain_emit = plain_emit;
	uint32_t *d_peer_flags = nullptr, *d_peer_flags2 = nullptr;
	if (dist && W > 1) {
		d_peer_flags = ctl + 8;
		d_peer_flags2 = ctl + 16;
		pr.peer_flags = d_peer_flags;
		pr.n_peer_flags = W;
	}

	RJRuns ra, rb;
	auto runs_of = [&](RJRuns *r, const RJSide &side, bool sorted, const mdbcu_table *t, int col) -> int {
		if (!sorted) {
			rj_runs_local(r, side);
			return MDBCU_OK;
		}
		memset(r, 0, sizeof(*r)); // nsrc = 0: pass 2 reads rows [bnd[p], bnd[p + 1]) of the column itself
		uint64_t *bnd;
		MDB_TRY(tmp.alloc(&bnd, (size_t)nparts + 1));
		MDB_LAUNCH(ctx, k_sorted_bounds, (nparts + 1 + 255) / 256, 256, 0, (const int64_t*)t->cols[col].data, (uint64_t)t->n_slots, pr, bnd);
		r->sorted_keys = t->cols[col].data;
		r->sorted_bnd = bnd;
		return MDBCU_OK;
	};
	MDB_TRY(runs_of(&ra, sa, sorted_a, ta, jn.left.col));
	MDB_TRY(runs_of(&rb, sb, sorted_b, tb, jn.right.col));
	RJShip ship;
	memset(&ship, 0, sizeof(ship));
	if (multi_gpu) {
		const RJRegionLayout lb = rj_region_layout((uint32_t)nparts, sb.cap, sb.tail_cap);
		const RJSlotLayout ls = rj_slot_layout(pown, sa.cap, sa.tail_cap);
		// side A: source o = the slot of THIS rank's arena that rank o's k_radix_ship fills (partition q = p - first), source
		// `me` = this rank's own streams; side B: source o = rank o's streams in ITS arena, read over NVLink by pass 2
		ra.nsrc = rb.nsrc = W;
		ra.self = rb.self = me;
		ship.world = W;
		ship.self = me;
		ship.nparts = nparts;
		ship.shipped_bytes = d_cursor + 1;
		for (int o = 0; o < W; o++) {
			const char *peer_b = (const char*)bases[o] + half_off + side_b_off;
			char *there = (char*)bases[o] + half_off + slots_off + (size_t)me * ls.bytes; // slot [me] of rank o's arena: where this rank pushes
			const char *here = (const char*)bases[me] + half_off + slots_off + (size_t)o * ls.bytes; // slot [o] of this rank's arena
			rb.stream[o] = (const uint16_t*)(peer_b + lb.main);
			rb.tail[o] = (const uint16_t*)(peer_b + lb.tail);
			rb.cursor[o] = (const uint32_t*)(peer_b + lb.cursor);
			rb.tail_cursor[o] = rb.cursor[o] + 1;
			rb.first[o] = 0;
			rb.cur_stride[o] = RJ_CUR_STRIDE;
			ship.main[o] = (uint16_t*)(there + ls.main);
			ship.tail[o] = (uint16_t*)(there + ls.tail);
			ship.cursor[o] = (uint32_t*)(there + ls.cursor);
			ship.tail_cursor[o] = (uint32_t*)(there + ls.tail_cursor);
			ra.stream[o] = (const uint16_t*)(here + ls.main);
			ra.tail[o] = (const uint16_t*)(here + ls.tail);
			ra.cursor[o] = (const uint32_t*)(here + ls.cursor);
			ra.tail_cursor[o] = (const uint32_t*)(here + ls.tail_cursor);
			ra.first[o] = (uint32_t)pr.part_first;
			ra.cur_stride[o] = 1;
		}
		ra.stream[me] = sa.stream;
		ra.tail[me] = sa.tail;
		ra.cursor[me] = sa.cursor;
		ra.tail_cursor[me] = sa.tail_cursor;
		ra.first[me] = 0;
		ra.cur_stride[me] = RJ_CUR_STRIDE;
		ra.pulled_bytes = nullptr; // (side A's bytes are counted by the push kernel)
		rb.pulled_bytes = d_cursor + 1;
	}

	// (per context: the attribute belongs to the device, and one process may drive several)
	if (!ctx->radix_attr_done) {
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_partition_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RJP1Smem)));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512, 0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512, 1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512, 2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<4, 512, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024, 0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024, 1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024, 2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute((k_radix_joincount<8, 1024, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
		CUDA_TRY(ctx, cudaFuncSetAttribute(k_radix_ship, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_SHIP_SMEM));
		ctx->radix_attr_done = true;
	}

	clock.begin(1);
	if (!sorted_a)
		launch_partition(ctx, grid1, sa, pr);
	if (multi_gpu) {
		// second stream: push side A's streams to their owners.  It takes a few SMs (CTAs that own their SM through a dynamic
		// shared-memory request) WHILE pass 1 of side B runs on the others: pass 1 is bound by shared memory, not by the SM
		// count, and loses about an eighth for it
		if (!ctx->side_stream) {
			CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
			CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->side_ev[0], cudaEventDisableTiming));
			CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->side_ev[1], cudaEventDisableTiming));
		}
		const int push_sms = std::max(4, ctx->num_sms / 9);
		CUDA_TRY(ctx, cudaEventRecord(ctx->side_ev[0], ctx->stream));
		CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->side_ev[0], 0));
		k_radix_ship<<<push_sms, 1024, RJ_SHIP_SMEM, ctx->side_stream>>>(sa, ship); // one CTA per SM (see RJ_SHIP_SMEM)
		ctx->stats.kernel_launches++;
		ctx->total_launches++;
		CUDA_TRY(ctx, cudaEventRecord(ctx->side_ev[1], ctx->side_stream));
		launch_partition(ctx, grid1 - push_sms, sb, pr);
		// ONE barrier: this rank's side A has landed in the owners' slots and its side B is partitioned.  The error flags
		// travel with it and stay on the device: pass 2 checks them itself, the host reads them with the result count
		clock.begin(7);
		CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->side_ev[1], 0));
		MDB_TRY(mdb_comm_arena_barrier(ctx, d_flags, d_peer_flags));
		guard.armed = false; // the barrier is in the stream: the peers get this rank's word
	} else if (!sorted_b) {
		launch_partition(ctx, grid1, sb, pr);
	}
	// (result columns are allocated here, after pass 1 is on its way, so that the GPU starts the query's first kernel as
	// early as the host allows)
	// upper bound of groups this rank can emit: one per key of the partitions it owns
	uint64_t cap_groups = std::min<uint64_t>((uint64_t)(pr.part_end - pr.part_first) * width, range);
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

	clock.begin(2);

	// 4-bit counters (two CTAs per SM) when keys are mostly unique per side, 8-bit otherwise; a wrapped
	// counter is detected by the checksum and the pass is repeated one width up before giving up
	const uint64_t D = ((uint64_t)width + 7) & ~7ull; // counters per side (whole 32-bit words of 4-bit fields)
	const uint64_t rows_a = dist ? ta->global_slots : ta->n_slots, rows_b = dist ? tb->global_slots : tb->n_slots;
	bool try4 = std::max(rows_a, rows_b) <= 2 * range;
	uint64_t ngroups = 0;
	uint32_t flags = 0, any1 = 0, any2 = 0; // this rank's flags; all ranks' flags after pass 1 / after pass 2
	const bool multi = d_peer_flags != nullptr;
	int attempt = try4 ? 0 : 1;
	bool run = true;
	// Single GPU: a wrapped 4-bit counter repeats pass 2 once with 8-bit counters.  Several GPUs: every round ends with a
	// barrier that carries the ranks' pass-2 flags, so that ALL ranks take the same decision (done / one more round for
	// the ranks whose 4-bit counters wrapped / give the query to the direct-count path, which is a collective).
	for (int round = 0; round < 2; round++) {
		if (run) {
			const int bitsw = attempt == 0 ? 4 : 8;
			const size_t smem2 = 2 * (size_t)std::max<uint64_t>(1, D * bitsw / 32) * sizeof(uint32_t);
			const int grid2 = std::max(1, std::min(pr.part_end - pr.part_first, ctx->num_sms * (bitsw == 4 ? 2 : 1)));
			// result layout known at compile time for the two common shapes: [key, count] and [count, key]
			const int layout = out.nout != 2 ? 0 : (!out.is_count[0] && out.is_count[1]) ? 1 : (out.is_count[0] && !out.is_count[1]) ? 2 : 0;
#define RJ_LAUNCH2(B, T, L)                                                                                   \
	do {                                                                                                  \
		if (ra.nsrc > 1 || rb.nsrc > 1)                                                               \
			MDB_LAUNCH(ctx, (k_radix_joincount<B, T, L, true>), grid2, T, smem2, ra, rb, pr, out, d_flags + 1);  \
		else                                                                                          \
			MDB_LAUNCH(ctx, (k_radix_joincount<B, T, L, false>), grid2, T, smem2, ra, rb, pr, out, d_flags + 1); \
	} while (0)
			if (bitsw == 4) {
				if (layout == 1)
					RJ_LAUNCH2(4, 512, 1);
				else if (layout == 2)
					RJ_LAUNCH2(4, 512, 2);
				else
					RJ_LAUNCH2(4, 512, 0);
			} else {
				if (layout == 1)
					RJ_LAUNCH2(8, 1024, 1);
				else if (layout == 2)
					RJ_LAUNCH2(8, 1024, 2);
				else
					RJ_LAUNCH2(8, 1024, 0);
			}
#undef RJ_LAUNCH2
			cudaError_t e = cudaGetLastError();
			if (e != cudaSuccess) {
				guard.armed = multi;
				return mdb_fail(ctx, MDBCU_ECUDA, "radix join launch failed: %s", cudaGetErrorString(e));
			}
		}
		if (multi) {
			guard.armed = true;
			MDB_TRY(mdb_comm_arena_barrier(ctx, d_flags, d_peer_flags2));
			guard.armed = false;
		}
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, ctl, 128, cudaMemcpyDeviceToHost, ctx->stream)); // the control block's head
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		ngroups = ctx->h_scalar[0];
		flags = (uint32_t)(ctx->h_scalar[2] & 0xffffffffu);
		any1 = any2 = 0;
		for (int r = 0; multi && r < W; r++) {
			any1 |= reinterpret_cast<const uint32_t*>(ctx->h_scalar + 4)[r];
			any2 |= reinterpret_cast<const uint32_t*>(ctx->h_scalar + 8)[r];
		}
		const uint32_t seen = multi ? any2 : flags;
		if (any1 || seen != RJ_ERR_COUNTER || !try4 || round == 1)
			break;
		// somebody's 4-bit counters wrapped: those ranks repeat pass 2 with 8-bit counters (the partitioned remainders are
		// still valid), the others only take part in the next barrier
		run = flags == RJ_ERR_COUNTER;
		if (run) {
			attempt = 1;
			CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), ctx->stream));
			CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
		}
	}
	clock.finish();
	if (flags || any1 || any2) {
		if (getenv("MDBCU_TRACE"))
			fprintf(stderr, "[mdbcu] radix join gives up: flags %u, any rank after pass 1 %u, after pass 2 %u (1 stream full, 2 counter "
					"wrapped, 4 clustered keys, 8 a rank aborted), groups %llu of %llu\n", flags, any1, any2, (unsigned long long)ngroups,
					(unsigned long long)cap_groups);
		if ((any1 | any2) & MDB_PEER_ABORT)
			return mdb_fail(ctx, MDBCU_EERROR, "distributed radix join: a peer rank aborted the query or never reached the barrier");
		// heavy duplicates / skew / stream overflow: the direct-count path (mdb_direct.cu) answers such inputs exactly; in a
		// distributed plan every rank is here (the flags travelled with the barriers)
		ctx->radix_gave_up = true;
		return MDBCU_EUNSUPPORTED;
	}
	if (ngroups > cap_groups)
		return mdb_fail(ctx, MDBCU_EINTERNAL, "radix join emitted %llu groups into %llu rows", (unsigned long long)ngroups,
				(unsigned long long)cap_groups);
	res->nrows = ngroups;
	ctx->stats.exchange_bytes = ctx->h_scalar[1];

	ctx->stats.algorithmic_bytes = 8ull * (ta->n_slots + tb->n_slots) + 8ull * plan->n_out * ngroups;
	ctx->stats.dominant_ms = ctx->stats.phase_ms[1] + ctx->stats.phase_ms[2];
	ctx->stats.dominant_bytes = ctx->stats.algorithmic_bytes;
	return MDBCU_OK;
}

// Host-side description of how a distributed radix join lays its exchange out (no device needed): which partitions every
// rank owns and where a join side's streams sit inside a rank's arena half.  The numbers come from the functions the kernels
// and mdb_select_radix_joincount use (rj_part_first / rj_owner_of / rj_stream_cap / rj_tail_cap / rj_region_layout), so the CPU
// test of the multi-rank arithmetic (tests/test_dist_cpu.py, world_size 2 over gloo) checks the shipped code, not a copy of it.
extern "C" int mdbcu_dist_describe(int nparts, int world, uint64_t global_rows, int sms, struct mdbcu_dist_layout *out)
{
	if (!out || nparts < 1 || nparts > RJ_MAX_PART || world < 1 || world > RJ_MAX_RANKS || sms < 1)
		return MDBCU_EERROR;
	memset(out, 0, sizeof(*out));
	for (int r = 0; r <= world; r++)
		out->part_first[r] = rj_part_first((uint32_t)r, (uint32_t)nparts, (uint32_t)world);
	out->stream_cap = rj_stream_cap((global_rows + world - 1) / world, nparts);
	out->tail_cap = rj_tail_cap(sms, out->stream_cap);
	const RJRegionLayout l = rj_region_layout((uint32_t)nparts, out->stream_cap, out->tail_cap);
	out->region_main_off = l.main;
	out->region_tail_off = l.tail;
	out->region_cursor_off = l.cursor;
	out->region_bytes = l.bytes;
	out->arena_half_bytes = 2 * l.bytes; // both join sides (equal row counts per side)
	return MDBCU_OK;
}

extern "C" int mdbcu_dist_owner(uint32_t partition, int nparts, int world)
{
	if (nparts < 1 || world < 1 || partition >= (uint32_t)nparts)
		return -1;
	return (int)rj_owner_of(partition, (uint32_t)nparts, (uint32_t)world);
}
