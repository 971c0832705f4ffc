// mdb_comm.cu - multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-process and has no communication layer; the only exchange on
// the path is the key-partitioned shuffle of join sides.  NCCL is resolved with dlopen
// at mdbcu_comm_init time so the single-GPU path carries no link-time dependency on libnccl and a
// process that already loaded torch's bundled libnccl.so.2 keeps using that copy.
//
// Two kinds of communicator sit behind the same internal calls (all-gather, peer-mapped arena, owner reduction):
//   * mdbcu_comm_init        one process per GPU; NCCL for the plumbing, CUDA IPC maps the arenas;
//   * mdbcu_comm_init_local  several contexts of ONE process (one host thread each), on different GPUs or - the
//                            loop-back mode of SURVEY.md 4.3 - all on the same GPU: a host rendezvous replaces NCCL and
//                            the peers' arenas are ordinary pointers.  The exchange kernels (k_arena_barrier, the multi-source
//                            pass 2 reading peer arenas, k_reduce_add_u32) are exactly the ones a multi-process run uses,
//                            so a single-GPU test box exercises them.
#include "mdb_common.cuh"

#include <dlfcn.h>
#include <string.h>
#include <chrono>
#include <condition_variable>
#include <mutex>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
#define NCCL_UINT8 1
#define NCCL_UINT32 3
#define NCCL_UINT64 5
#define NCCL_SUM 0

struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl(mdbcu_ctx *ctx)
{
	if (g_nccl.handle)
		return MDBCU_OK;
	// Order: an explicit path (MDBCU_NCCL_LIB; the Python binding points it at the NCCL wheel torch links against), a copy
	// the process has loaded already (torch's), then the system library.  One process must not end up with two different
	// NCCL builds behind the same SONAME: whoever loads second gets the first one's symbols.
	const char *names[] = {getenv("MDBCU_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
	for (int i = 0; i < 3 && !h; i++)
		if (names[i] && names[i][0])
			h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
	if (!h)
		return mdb_fail(ctx, MDBCU_ECUDA, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                          \
	do {                                                                                      \
		*(void**)(&g_nccl.field) = dlsym(h, name);                                        \
		if (!g_nccl.field)                                                                \
			return mdb_fail(ctx, MDBCU_ECUDA, "libnccl is missing symbol %s", name);  \
	} while (0)
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(GroupStart, "ncclGroupStart");
	SYM(GroupEnd, "ncclGroupEnd");
	SYM(Send, "ncclSend");
	SYM(Recv, "ncclRecv");
	SYM(AllGather, "ncclAllGather");
	SYM(Reduce, "ncclReduce");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
	g_nccl.handle = h;
	return MDBCU_OK;
}

#define NCCL_TRY(ctx, call)                                                                       \
	do {                                                                                      \
		ncclResult_t _r = (call);                                                         \
		if (_r != 0)                                                                      \
			return mdb_fail((ctx), MDBCU_ECUDA, "%s failed: %s", #call, g_nccl.GetErrorString(_r)); \
	} while (0)

// ---- in-process communicator: W contexts of this process, one host thread each
struct mdb_local_group {
	std::mutex m;
	std::condition_variable cv;
	int world = 0;
	int arrived = 0;
	uint64_t generation = 0;
	int refs = 0;
	bool broken = false; // a rank timed out at the rendezvous: every later collective fails at once
	mdbcu_ctx *ctxs[MDB_MAX_RANKS] = {};
	const void *posted[MDB_MAX_RANKS] = {}; // what every rank offers to its peers in the current collective
};

// host rendezvous of all ranks (each from its own thread); false if a peer does not show up within the time limit
static bool local_barrier(mdb_local_group *g)
{
	std::unique_lock<std::mutex> lk(g->m);
	if (g->broken)
		return false;
	const uint64_t gen = g->generation;
	if (++g->arrived == g->world) {
		g->arrived = 0;
		g->generation++;
		g->cv.notify_all();
		return true;
	}
	if (!g->cv.wait_for(lk, std::chrono::seconds(120), [&] { return g->generation != gen || g->broken; })) {
		g->broken = true;
		g->cv.notify_all();
		return false;
	}
	return !g->broken;
}

#define LOCAL_BARRIER(ctx)                                                                                          \
	do {                                                                                                        \
		if (!local_barrier((ctx)->local_group))                                                             \
			return mdb_fail((ctx), MDBCU_EERROR, "in-process communicator: a rank did not reach the collective"); \
	} while (0)

bool mdb_comm_ready(const mdbcu_ctx *ctx)
{
	return ctx->nccl_comm != nullptr || ctx->local_group != nullptr;
}

extern "C" int mdbcu_comm_init_local(mdbcu_ctx *const *ctxs, int world)
{
	if (!ctxs || world < 1 || world > MDB_MAX_RANKS)
		return MDBCU_EERROR;
	for (int r = 0; r < world; r++)
		if (!ctxs[r] || mdb_comm_ready(ctxs[r]))
			return ctxs[r] ? mdb_fail(ctxs[r], MDBCU_EERROR, "mdbcu_comm_init_local: context %d already has a communicator", r) : MDBCU_EERROR;
	mdb_local_group *g = new (std::nothrow) mdb_local_group();
	if (!g)
		return mdb_fail(ctxs[0], MDBCU_ENOMEM, "out of host memory");
	g->world = world;
	g->refs = world;
	for (int r = 0; r < world; r++) {
		g->ctxs[r] = ctxs[r];
		// contexts on different GPUs store into each other's arenas: peer access both ways (a no-op on one GPU)
		for (int o = 0; o < world; o++) {
			if (ctxs[o]->device == ctxs[r]->device)
				continue;
			cudaSetDevice(ctxs[r]->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[o]->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
				delete g;
				return mdb_fail(ctxs[r], MDBCU_ECUDA, "no peer access from GPU %d to GPU %d: %s", ctxs[r]->device, ctxs[o]->device,
						cudaGetErrorString(e));
			}
			cudaGetLastError();
			// query temporaries come from the stream-ordered pool, which peers cannot touch unless told so
			// (the all-gather and the owner reduction read them through peer pointers)
			cudaMemPool_t pool;
			if (cudaDeviceGetDefaultMemPool(&pool, ctxs[r]->device) == cudaSuccess) {
				cudaMemAccessDesc desc;
				memset(&desc, 0, sizeof(desc));
				desc.location.type = cudaMemLocationTypeDevice;
				desc.location.id = ctxs[o]->device;
				desc.flags = cudaMemAccessFlagsProtReadWrite;
				e = cudaMemPoolSetAccess(pool, &desc, 1);
				if (e != cudaSuccess) {
					delete g;
					return mdb_fail(ctxs[r], MDBCU_ECUDA, "cannot open GPU %d's memory pool to GPU %d: %s", ctxs[r]->device,
							ctxs[o]->device, cudaGetErrorString(e));
				}
			}
		}
	}
	for (int r = 0; r < world; r++) {
		ctxs[r]->local_group = g;
		ctxs[r]->rank = r;
		ctxs[r]->world = world;
	}
	return MDBCU_OK;
}

extern "C" int mdbcu_comm_unique_id(mdbcu_ctx *ctx, void *id128)
{
	if (!ctx || !id128)
		return MDBCU_EERROR;
	MDB_TRY(load_nccl(ctx));
	ncclUniqueId id;
	NCCL_TRY(ctx, g_nccl.GetUniqueId(&id));
	memcpy(id128, &id, sizeof(id));
	return MDBCU_OK;
}

extern "C" int mdbcu_comm_init(mdbcu_ctx *ctx, int rank, int world, const void *id128)
{
	if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world)
		return ctx ? mdb_fail(ctx, MDBCU_EERROR, "mdbcu_comm_init: bad arguments") : MDBCU_EERROR;
	cudaSetDevice(ctx->device);
	MDB_TRY(load_nccl(ctx));
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclComm_t comm = nullptr;
	NCCL_TRY(ctx, g_nccl.CommInitRank(&comm, world, id, rank));
	ctx->nccl_comm = comm;
	ctx->rank = rank;
	ctx->world = world;
	return MDBCU_OK;
}

extern "C" int mdbcu_comm_world(mdbcu_ctx *ctx, int *rank, int *world)
{
	if (!ctx)
		return MDBCU_EERROR;
	if (rank)
		*rank = ctx->rank;
	if (world)
		*world = ctx->world;
	return MDBCU_OK;
}

void mdb_comm_destroy(mdbcu_ctx *ctx)
{
	if (ctx->nccl_comm && g_nccl.CommDestroy)
		g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
	ctx->nccl_comm = nullptr;
	if (ctx->local_group) {
		mdb_local_group *g = ctx->local_group;
		bool last;
		{
			std::lock_guard<std::mutex> lk(g->m);
			g->broken = true; // a collective with fewer ranks than `world` cannot complete
			g->cv.notify_all();
			last = --g->refs == 0;
		}
		if (last)
			delete g;
		ctx->local_group = nullptr;
	}
}

int mdb_comm_allgather_u64(mdbcu_ctx *ctx, const uint64_t *send, uint64_t *recv, size_t count)
{
	return mdb_comm_allgather_bytes(ctx, send, recv, count * sizeof(uint64_t));
}

// ---- NCCL is plumbing here (statistics, IPC handles); the join's data moves through the arena below

int mdb_comm_allgather_bytes(mdbcu_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank)
{
	if (ctx->local_group) {
		// what this rank offers must be complete before a peer's copy reads it
		mdb_local_group *g = ctx->local_group;
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		g->posted[ctx->rank] = send;
		LOCAL_BARRIER(ctx);
		for (int r = 0; r < ctx->world; r++)
			CUDA_TRY(ctx, cudaMemcpyAsync((char*)recv + (size_t)r * bytes_per_rank, g->posted[r], bytes_per_rank, cudaMemcpyDefault, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		LOCAL_BARRIER(ctx); // nobody reuses its send buffer while a peer still copies from it
		return MDBCU_OK;
	}
	if (!ctx->nccl_comm)
		return mdb_fail(ctx, MDBCU_EERROR, "distributed plan without mdbcu_comm_init");
	NCCL_TRY(ctx, g_nccl.AllGather(send, recv, bytes_per_rank, NCCL_UINT8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	return MDBCU_OK;
}

// ---- peer-mapped exchange arena: one cudaMalloc'd block per rank, mapped into every other rank's address
// space with CUDA IPC, so a kernel can store straight into the owner GPU's memory over NVLink / NVSwitch.
// COLLECTIVE: every rank calls it with the same size; the mapping is cached until a larger one is needed.

#define MDB_ARENA_HEADER 256 // bytes in front of every arena: two barrier flag words per source rank (k_arena_barrier)

static void arena_release(mdbcu_ctx *ctx)
{
	for (int r = 0; r < MDB_MAX_RANKS; r++) {
		if (ctx->arena_peer[r] && r != ctx->rank && !ctx->local_group)
			cudaIpcCloseMemHandle(ctx->arena_peer[r]);
		ctx->arena_peer[r] = nullptr;
	}
	if (ctx->arena_local)
		cudaFree(ctx->arena_local);
	ctx->arena_local = nullptr;
	ctx->arena_bytes = 0;
}

void mdb_comm_arena_destroy(mdbcu_ctx *ctx)
{
	arena_release(ctx);
}

int mdb_comm_arena(mdbcu_ctx *ctx, size_t bytes, void **bases)
{
	const int W = ctx->world;
	if (W > MDB_MAX_RANKS)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "at most %d ranks", MDB_MAX_RANKS);
	if (ctx->arena_bytes < bytes) {
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		arena_release(ctx);
		bytes += bytes / 8 + MDB_ARENA_HEADER; // head-room so slightly larger follow-up queries reuse the mapping
		cudaError_t e = cudaMalloc(&ctx->arena_local, bytes);
		if (e != cudaSuccess) {
			ctx->arena_local = nullptr;
			return mdb_fail(ctx, MDBCU_ENOMEM, "cannot allocate the %zu-byte exchange arena: %s", bytes, cudaGetErrorString(e));
		}
		ctx->arena_bytes = bytes - MDB_ARENA_HEADER;
		ctx->arena_peer[ctx->rank] = ctx->arena_local;
		CUDA_TRY(ctx, cudaMemsetAsync(ctx->arena_local, 0, MDB_ARENA_HEADER, ctx->stream)); // barrier flag words
		if (W > 1 && ctx->local_group) {
			// same process: the peers' blocks are ordinary device pointers (peer access is on, mdbcu_comm_init_local)
			mdb_local_group *g = ctx->local_group;
			CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // header zeroed before any peer can store a flag
			g->posted[ctx->rank] = ctx->arena_local;
			LOCAL_BARRIER(ctx);
			for (int r = 0; r < W; r++)
				ctx->arena_peer[r] = const_cast<void*>(g->posted[r]);
			LOCAL_BARRIER(ctx);
		} else if (W > 1) {
			cudaIpcMemHandle_t mine;
			CUDA_TRY(ctx, cudaIpcGetMemHandle(&mine, ctx->arena_local));
			DevTemp tmp(ctx);
			unsigned char *d_mine, *d_all;
			MDB_TRY(tmp.alloc(&d_mine, sizeof(mine)));
			MDB_TRY(tmp.alloc(&d_all, sizeof(mine) * W));
			CUDA_TRY(ctx, cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
			MDB_TRY(mdb_comm_allgather_bytes(ctx, d_mine, d_all, sizeof(mine)));
			std::vector<cudaIpcMemHandle_t> all(W);
			CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * W, cudaMemcpyDeviceToHost, ctx->stream));
			CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
			for (int r = 0; r < W; r++) {
				if (r == ctx->rank)
					continue;
				e = cudaIpcOpenMemHandle(&ctx->arena_peer[r], all[r], cudaIpcMemLazyEnablePeerAccess);
				if (e != cudaSuccess) {
					ctx->arena_peer[r] = nullptr;
					return mdb_fail(ctx, MDBCU_ECUDA, "cannot map rank %d's exchange arena (CUDA IPC / peer access): %s", r,
							cudaGetErrorString(e));
				}
			}
		}
	}
	for (int r = 0; r < W; r++)
		bases[r] = (char*)ctx->arena_peer[r] + MDB_ARENA_HEADER;
	return MDBCU_OK;
}

// ---- cross-rank barrier on the arena itself: every rank stores (epoch << 4 | its 4 error bits) into word [self] of every
// peer's arena header over NVLink and spins until all `world` words of its own header have reached the epoch.  A few
// microseconds where an NCCL all-gather of 4 bytes costs tens; no host involvement.  d_all[r] receives rank r's error bits.
// The spin is bounded: a peer that never arrives (it failed on the host before launching its barrier and could not even
// publish MDB_PEER_ABORT, see mdb_comm_arena_abort) is reported as error bit MDB_PEER_TIMEOUT instead of hanging the GPU.
struct ArenaBarrierArgs {
	uint32_t *peer_hdr[MDB_MAX_RANKS];
	int world, self;
	uint32_t epoch;
	unsigned long long timeout_ns;
};

__global__ void k_arena_barrier(ArenaBarrierArgs a, const uint32_t *__restrict__ err, uint32_t *__restrict__ all)
{
	const int r = threadIdx.x;
	if (r >= a.world)
		return;
	const uint32_t mine = (a.epoch << 4) | ((err ? *err : MDB_PEER_ABORT) & 0xfu);
	// two words per source rank, used by alternate epochs: a rank that is already at its NEXT barrier writes the other word,
	// so the word (and the error bits) of THIS epoch stay readable until every peer has passed it
	const uint32_t word = (a.epoch & 1u) * MDB_MAX_RANKS;
	__threadfence_system(); // everything this GPU wrote before this kernel is visible first
	if (!all) {
		// abort notice (mdb_comm_arena_abort): both words, and do not wait for anybody
		asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_hdr[r] + a.self), "r"(mine) : "memory");
		asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_hdr[r] + MDB_MAX_RANKS + a.self), "r"(mine) : "memory");
		return;
	}
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_hdr[r] + word + a.self), "r"(mine) : "memory");
	const uint32_t *slot = a.peer_hdr[a.self] + word + r;
	uint32_t v;
	unsigned long long t0 = 0, now = 0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	do {
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
		if ((v >> 4) >= a.epoch)
			break;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
	} while (now - t0 < a.timeout_ns);
	all[r] = (v >> 4) >= a.epoch ? (v & 0xfu) : MDB_PEER_TIMEOUT;
}

static void arena_barrier_args(mdbcu_ctx *ctx, ArenaBarrierArgs *a)
{
	memset(a, 0, sizeof(*a));
	for (int r = 0; r < ctx->world; r++)
		a->peer_hdr[r] = (uint32_t*)ctx->arena_peer[r];
	a->world = ctx->world;
	a->self = ctx->rank;
	a->epoch = ++ctx->arena_epoch;
	static const unsigned long long limit_ms = getenv("MDBCU_BARRIER_TIMEOUT_MS") ? strtoull(getenv("MDBCU_BARRIER_TIMEOUT_MS"), nullptr, 10) : 20000ull;
	a->timeout_ns = limit_ms * 1000000ull;
}

int mdb_comm_arena_barrier_on(mdbcu_ctx *ctx, cudaStream_t stream, const uint32_t *d_err, uint32_t *d_all)
{
	if (!ctx->arena_local)
		return mdb_fail(ctx, MDBCU_EERROR, "arena barrier without an arena");
	ArenaBarrierArgs a;
	arena_barrier_args(ctx, &a);
	k_arena_barrier<<<1, 32, 0, stream>>>(a, d_err, d_all);
	ctx->stats.kernel_launches++;
	ctx->total_launches++;
	return MDBCU_OK;
}

int mdb_comm_arena_barrier(mdbcu_ctx *ctx, const uint32_t *d_err, uint32_t *d_all)
{
	return mdb_comm_arena_barrier_on(ctx, ctx->stream, d_err, d_all);
}

// This rank gives up on a distributed query AFTER its peers may have started to wait for it: publish MDB_PEER_ABORT with the
// LARGEST epoch in place of every barrier this rank will not reach, so that the peers' barrier kernels return at once (their
// pass 2 sees the flag, the query fails on every rank) - now and in every later query: like an NCCL communicator after a
// rank failure, the communicator is unusable from here on, but nobody hangs.
void mdb_comm_arena_abort(mdbcu_ctx *ctx)
{
	if (!ctx->arena_local || ctx->world < 2)
		return;
	cudaSetDevice(ctx->device);
	cudaGetLastError();
	ArenaBarrierArgs a;
	arena_barrier_args(ctx, &a);
	a.epoch = 0x0fffffffu;
	ctx->arena_epoch = 0x0ffffff0u;
	k_arena_barrier<<<1, 32, 0, ctx->stream>>>(a, nullptr, nullptr);
	cudaStreamSynchronize(ctx->stream);
	cudaGetLastError();
}

// ---- owner reduction of 32-bit counters (direct-count join, mdb_direct.cu): `buf` holds `nsides` arrays of `n` counters
// on every rank; afterwards slice [n*r/W, n*(r+1)/W) of every array on rank r holds the sum over all ranks (the other
// slices are unspecified).  NCCL: one grouped ncclReduce per (array, owner) - the slices differ by at most one element,
// which ncclReduceScatter cannot express.  In-process communicator: every rank adds its peers' slices with plain loads
// through the peer pointers.
__global__ void k_reduce_add_u32(uint32_t *__restrict__ mine, const uint32_t *__restrict__ peer, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		mine[i] += peer[i];
}

int mdb_comm_reduce_owned_u32(mdbcu_ctx *ctx, uint32_t *buf, uint64_t n, int nsides, uint64_t *bytes_sent)
{
	const int W = ctx->world, me = ctx->rank;
	auto first_of = [&](int r) { return (uint64_t)((unsigned __int128)n * r / W); };
	if (bytes_sent)
		*bytes_sent = (uint64_t)nsides * (n - (first_of(me + 1) - first_of(me))) * sizeof(uint32_t);
	if (W < 2)
		return MDBCU_OK;
	if (ctx->local_group) {
		mdb_local_group *g = ctx->local_group;
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // my counters are complete
		g->posted[me] = buf;
		LOCAL_BARRIER(ctx);
		const uint64_t lo = first_of(me), cnt = first_of(me + 1) - lo;
		for (int o = 0; o < W && cnt; o++) {
			if (o == me)
				continue;
			const uint32_t *peer = (const uint32_t*)g->posted[o];
			for (int s = 0; s < nsides; s++) {
				const int grid = (int)std::min<uint64_t>(mdb_div_up(cnt, 256), (uint64_t)ctx->num_sms * 8);
				MDB_LAUNCH(ctx, k_reduce_add_u32, grid, 256, 0, buf + (size_t)s * n + lo, peer + (size_t)s * n + lo, cnt);
			}
		}
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		LOCAL_BARRIER(ctx); // every rank has read what it needs: the buffers may be released
		return MDBCU_OK;
	}
	if (!ctx->nccl_comm)
		return mdb_fail(ctx, MDBCU_EERROR, "distributed plan without mdbcu_comm_init");
	NCCL_TRY(ctx, g_nccl.GroupStart());
	for (int s = 0; s < nsides; s++)
		for (int r = 0; r < W; r++) {
			const uint64_t lo = first_of(r), cnt = first_of(r + 1) - lo;
			if (!cnt)
				continue;
			uint32_t *p = buf + (size_t)s * n + lo;
			NCCL_TRY(ctx, g_nccl.Reduce(p, p, cnt, NCCL_UINT32, NCCL_SUM, r, (ncclComm_t)ctx->nccl_comm, ctx->stream));
		}
	NCCL_TRY(ctx, g_nccl.GroupEnd());
	return MDBCU_OK;
}
