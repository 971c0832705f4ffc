// mdb_radix_dist.cuh - multi-GPU exchange of the radix join+count (included by mdb_radix.cu).
//
// One process per GPU.  Every rank partitions ITS shard of both join sides with pass 1 over the same global key
// range (mdbcu_table_sync_stats), so partition p means the same key interval everywhere.  Rank r owns the
// contiguous partition block [r*P/W, (r+1)*P/W): key ranges are disjoint, so after the exchange every rank
// runs pass 2 on its own partitions and the per-rank results simply concatenate - no merge step.
//
// What crosses NVLink is the 2-byte remainders, not the 8-byte keys: per chunk only the 16-byte granules that
// hold data, plus 4 bytes of descriptor.  The exchange is one NCCL grouped send/recv (all-to-all) per query for
// both sides together, preceded by one small all-gather of per-partition chunk counts.
#pragma once

#include <chrono>
#include <cstdlib>
#include <vector>

__global__ void k_dist_granules(const RJDesc *__restrict__ dir, uint64_t n, uint32_t *__restrict__ gran, uint32_t *__restrict__ ne)
{
	for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < n; c += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t e = dir[c].ne;
		ne[c] = e;
		gran[c] = (e + 7) / 8;
	}
}

__global__ void k_dist_granules_of(const uint32_t *__restrict__ ne, uint64_t n, uint32_t *__restrict__ gran)
{
	for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < n; c += (uint64_t)gridDim.x * blockDim.x)
		gran[c] = (ne[c] + 7) / 8;
}

// one warp per chunk: the lanes that hold data copy one 16-byte granule each into the send buffer
__global__ void k_dist_gather(const uint16_t *__restrict__ pool, const RJDesc *__restrict__ dir, const uint64_t *__restrict__ goff,
		uint64_t n, int4 *__restrict__ send)
{
	const int lane = threadIdx.x & 31;
	const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t c = warp; c < n; c += nwarps) {
		const RJDesc d = dir[c];
		if ((uint32_t)lane * 8 < d.ne)
			send[goff[c] + lane] = mdb_ldg_stream(reinterpret_cast<const int4*>(pool) + (size_t)d.off16 + lane);
	}
}

// header of one side: [chunks per partition (P) | granules per destination rank (W)] as u64
__global__ void k_dist_header(const uint32_t *__restrict__ dir_cnt, const uint64_t *__restrict__ dir_off, const uint64_t *__restrict__ goff,
		const uint64_t *__restrict__ gtotal, uint64_t nchunks, int P, int W, uint64_t *__restrict__ hdr)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P + W; i += gridDim.x * blockDim.x) {
		if (i < P) {
			hdr[i] = dir_cnt[i];
		} else {
			const int d = i - P;
			const uint64_t c0 = dir_off[(uint64_t)d * P / W], c1 = dir_off[(uint64_t)(d + 1) * P / W];
			const uint64_t g0 = c0 < nchunks ? goff[c0] : *gtotal, g1 = c1 < nchunks ? goff[c1] : *gtotal;
			hdr[i] = g1 - g0;
		}
	}
}

// one block per (source rank, owned partition): received chunk k of the pair becomes descriptor dst_start + k
__global__ void k_dist_remote_dir(const uint32_t *__restrict__ ne_recv, const uint64_t *__restrict__ goff_recv,
		const uint64_t *__restrict__ src_start, const uint64_t *__restrict__ dst_start, const uint64_t *__restrict__ cnt, RJDesc *__restrict__ dir2)
{
	const uint64_t s0 = src_start[blockIdx.x], d0 = dst_start[blockIdx.x], n = cnt[blockIdx.x];
	for (uint64_t k = threadIdx.x; k < n; k += blockDim.x) {
		RJDesc d;
		d.off16 = (uint32_t)goff_recv[s0 + k];
		d.ne = ne_recv[s0 + k];
		dir2[d0 + k] = d;
	}
}

struct RJDistSide {
	uint64_t nchunks = 0;
	uint32_t *gran = nullptr, *ne = nullptr;
	uint64_t *goff = nullptr, *gtotal = nullptr;
	int4 *send = nullptr;
	// receive side
	uint32_t *ne_recv = nullptr, *gran_recv_cnt = nullptr;
	uint64_t *goff_recv = nullptr;
	int4 *gran_recv = nullptr;
	RJDesc *dir2 = nullptr;
	uint64_t *dir_off2 = nullptr;
};

// sides[0..1] are rewritten to describe the RECEIVED remainders of the partitions this rank owns
static int rj_exchange(mdbcu_ctx *ctx, DevTemp &tmp, RJParams *pr, RJSide *sides[2], uint64_t *exchanged_bytes)
{
	const int W = ctx->world, rank = ctx->rank, P = pr->nparts;
	const int plo = (int)((uint64_t)rank * P / W), phi = (int)((uint64_t)(rank + 1) * P / W);
	const int nmine = phi - plo;
	const size_t hdr_len = (size_t)P + W; // per side
	RJDistSide ds[2];
	// MDBCU_TRACE=1: wall-clock of every step of the exchange on stderr (synchronises the stream; debugging aid only)
	static const bool trace = getenv("MDBCU_TRACE") != nullptr;
	auto t_last = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) {
		if (!trace)
			return;
		cudaStreamSynchronize(ctx->stream);
		auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[mdbcu rank %d] exchange %-28s %8.1f us\n", rank, what,
				std::chrono::duration<double, std::micro>(now - t_last).count());
		t_last = now;
	};
	lap("start (waits for pass 1)");

	// 1. local chunk counts: dir_off[P] of each side (two 8-byte reads), then granules / descriptors in directory order
	for (int s = 0; s < 2; s++) {
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar + s, sides[s]->dir_off + P, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
	}
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	uint64_t *hdr_local, *hdr_all;
	MDB_TRY(tmp.alloc(&hdr_local, 2 * hdr_len));
	MDB_TRY(tmp.alloc(&hdr_all, 2 * hdr_len * W));
	for (int s = 0; s < 2; s++) {
		RJDistSide &d = ds[s];
		d.nchunks = ctx->h_scalar[s];
		MDB_TRY(tmp.alloc(&d.gran, d.nchunks));
		MDB_TRY(tmp.alloc(&d.ne, d.nchunks));
		MDB_TRY(tmp.alloc(&d.goff, d.nchunks));
		MDB_TRY(tmp.alloc(&d.gtotal, 1));
		if (d.nchunks) {
			int grid = (int)std::min<uint64_t>(mdb_div_up(d.nchunks, 256), (uint64_t)ctx->num_sms * 8);
			MDB_LAUNCH(ctx, k_dist_granules, grid, 256, 0, (const RJDesc*)sides[s]->dir, d.nchunks, d.gran, d.ne);
		}
		MDB_TRY(mdb_scan_u32_u64(ctx, d.gran, d.goff, d.nchunks, d.gtotal));
		MDB_LAUNCH(ctx, k_dist_header, 8, 256, 0, (const uint32_t*)sides[s]->dst[sides[s]->self].dir_cnt, (const uint64_t*)sides[s]->dir_off,
				(const uint64_t*)d.goff, (const uint64_t*)d.gtotal, d.nchunks, P, W, hdr_local + s * hdr_len);
	}
	CUDA_CHECK_LAUNCH(ctx);
	lap("granules+scan+header");

	// 2. everyone learns everyone's per-partition chunk counts and per-destination granule counts
	MDB_TRY(mdb_comm_allgather_bytes(ctx, hdr_local, hdr_all, 2 * hdr_len * sizeof(uint64_t)));
	std::vector<uint64_t> h((size_t)2 * hdr_len * W);
	CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), hdr_all, h.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	lap("allgather header + D2H");
	auto cnt_of = [&](int src, int side, int p) { return h[((size_t)src * 2 + side) * hdr_len + p]; };
	auto gran_to = [&](int src, int side, int dst) { return h[((size_t)src * 2 + side) * hdr_len + P + dst]; };

	// 3. offsets (host; a few thousand integers)
	std::vector<uint64_t> send_c[2], send_g[2], recv_c[2], recv_g[2];
	for (int s = 0; s < 2; s++) {
		RJDistSide &d = ds[s];
		send_c[s].assign(W + 1, 0);
		send_g[s].assign(W + 1, 0);
		recv_c[s].assign(W + 1, 0);
		recv_g[s].assign(W + 1, 0);
		for (int dst = 0; dst < W; dst++) {
			uint64_t c = 0;
			for (int p = (int)((uint64_t)dst * P / W); p < (int)((uint64_t)(dst + 1) * P / W); p++)
				c += cnt_of(rank, s, p);
			send_c[s][dst + 1] = send_c[s][dst] + c;
			send_g[s][dst + 1] = send_g[s][dst] + gran_to(rank, s, dst);
		}
		for (int src = 0; src < W; src++) {
			uint64_t c = 0;
			for (int p = plo; p < phi; p++)
				c += cnt_of(src, s, p);
			recv_c[s][src + 1] = recv_c[s][src] + c;
			recv_g[s][src + 1] = recv_g[s][src] + gran_to(src, s, rank);
		}
		const uint64_t RC = recv_c[s][W], RG = recv_g[s][W];
		if (RG >= (1ull << 32))
			return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "a rank would receive more than 64 GiB of remainders");
		MDB_TRY(tmp.alloc(&d.send, send_g[s][W]));
		MDB_TRY(tmp.alloc(&d.ne_recv, RC));
		MDB_TRY(tmp.alloc(&d.gran_recv_cnt, RC));
		MDB_TRY(tmp.alloc(&d.goff_recv, RC));
		MDB_TRY(tmp.alloc(&d.gran_recv, RG));
		MDB_TRY(tmp.alloc(&d.dir2, RC));
		MDB_TRY(tmp.alloc(&d.dir_off2, (size_t)P + 2));
		lap("  offsets + alloc");
		if (d.nchunks) {
			int grid = (int)std::min<uint64_t>(mdb_div_up(d.nchunks * 32, 256), (uint64_t)ctx->num_sms * 16);
			MDB_LAUNCH(ctx, k_dist_gather, grid, 256, 0, (const uint16_t*)sides[s]->pool, (const RJDesc*)sides[s]->dir,
					(const uint64_t*)d.goff, d.nchunks, d.send);
		}
	}
	CUDA_CHECK_LAUNCH(ctx);
	lap("offsets + alloc + gather");

	// 4. the exchange: descriptors (4 B per chunk) and granules (16 B each) of both sides in ONE grouped all-to-all
	MDB_TRY(mdb_comm_group_begin(ctx));
	for (int peer = 0; peer < W; peer++) {
		for (int s = 0; s < 2; s++) {
			RJDistSide &d = ds[s];
			MDB_TRY(mdb_comm_send(ctx, d.ne + send_c[s][peer], (send_c[s][peer + 1] - send_c[s][peer]) * sizeof(uint32_t), peer));
			MDB_TRY(mdb_comm_recv(ctx, d.ne_recv + recv_c[s][peer], (recv_c[s][peer + 1] - recv_c[s][peer]) * sizeof(uint32_t), peer));
			MDB_TRY(mdb_comm_send(ctx, d.send + send_g[s][peer], (send_g[s][peer + 1] - send_g[s][peer]) * sizeof(int4), peer));
			MDB_TRY(mdb_comm_recv(ctx, d.gran_recv + recv_g[s][peer], (recv_g[s][peer + 1] - recv_g[s][peer]) * sizeof(int4), peer));
			if (peer != rank)
				*exchanged_bytes += (send_c[s][peer + 1] - send_c[s][peer]) * sizeof(uint32_t) +
						    (send_g[s][peer + 1] - send_g[s][peer]) * sizeof(int4);
		}
	}
	MDB_TRY(mdb_comm_group_end(ctx));
	lap("grouped send/recv");

	// 5. descriptors of the received chunks, grouped by (owned) partition
	for (int s = 0; s < 2; s++) {
		RJDistSide &d = ds[s];
		const uint64_t RC = recv_c[s][W];
		std::vector<uint64_t> off2((size_t)P + 2, 0), src_start, dst_start, cnt;
		uint64_t run = 0;
		for (int p = 0; p <= P; p++) {
			off2[p] = run;
			if (p >= plo && p < phi)
				for (int src = 0; src < W; src++)
					run += cnt_of(src, s, p);
		}
		off2[P + 1] = run;
		src_start.reserve((size_t)W * nmine);
		dst_start.reserve((size_t)W * nmine);
		cnt.reserve((size_t)W * nmine);
		std::vector<uint64_t> within(nmine, 0); // chunks of partition p already placed by lower ranks
		for (int src = 0; src < W; src++) {
			uint64_t sc = recv_c[s][src];
			for (int p = plo; p < phi; p++) {
				const uint64_t c = cnt_of(src, s, p);
				src_start.push_back(sc);
				dst_start.push_back(off2[p] + within[p - plo]);
				cnt.push_back(c);
				sc += c;
				within[p - plo] += c;
			}
		}
		uint64_t *d_pairs;
		const size_t np = cnt.size();
		MDB_TRY(tmp.alloc(&d_pairs, 3 * np + 1));
		CUDA_TRY(ctx, cudaMemcpyAsync(d.dir_off2, off2.data(), off2.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
		if (np) {
			CUDA_TRY(ctx, cudaMemcpyAsync(d_pairs, src_start.data(), np * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
			CUDA_TRY(ctx, cudaMemcpyAsync(d_pairs + np, dst_start.data(), np * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
			CUDA_TRY(ctx, cudaMemcpyAsync(d_pairs + 2 * np, cnt.data(), np * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
		}
		// the host vectors die at the end of this iteration: wait for the uploads (pageable memory copies are staged anyway)
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		if (RC) {
			int grid = (int)std::min<uint64_t>(mdb_div_up(RC, 256), (uint64_t)ctx->num_sms * 8);
			MDB_LAUNCH(ctx, k_dist_granules_of, grid, 256, 0, (const uint32_t*)d.ne_recv, RC, d.gran_recv_cnt);
			MDB_TRY(mdb_scan_u32_u64(ctx, d.gran_recv_cnt, d.goff_recv, RC, nullptr));
			if (np)
				MDB_LAUNCH(ctx, k_dist_remote_dir, (unsigned)np, 128, 0, (const uint32_t*)d.ne_recv, (const uint64_t*)d.goff_recv,
						(const uint64_t*)d_pairs, (const uint64_t*)(d_pairs + np), (const uint64_t*)(d_pairs + 2 * np), d.dir2);
		}
		sides[s]->pool = reinterpret_cast<uint16_t*>(d.gran_recv);
		sides[s]->dir = d.dir2;
		sides[s]->dir_off = d.dir_off2;
	}
	CUDA_CHECK_LAUNCH(ctx);
	lap("remote directory");
	pr->part_first = plo;
	pr->part_end = phi;
	return MDBCU_OK;
}
