// mdb_comm.cu - multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-process and has no communication layer; the only exchange on
// the path is the key-partitioned shuffle of join sides.  NCCL is resolved with dlopen
// at mdbcu_comm_init time so the single-GPU path carries no link-time dependency on libnccl and a
// process that already loaded torch's bundled libnccl.so.2 keeps using that copy.
#include "mdb_common.cuh"

#include <dlfcn.h>
#include <string.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
#define NCCL_UINT8 1
#define NCCL_UINT64 5

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
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl(mdbcu_ctx *ctx)
{
	if (g_nccl.handle)
		return MDBCU_OK;
	const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
	void *h = nullptr;
	for (int i = 0; names[i] && !h; i++)
		h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
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
}

int mdb_comm_allgather_u64(mdbcu_ctx *ctx, const uint64_t *send, uint64_t *recv, size_t count)
{
	if (!ctx->nccl_comm)
		return mdb_fail(ctx, MDBCU_EERROR, "distributed plan without mdbcu_comm_init");
	NCCL_TRY(ctx, g_nccl.AllGather(send, recv, count, NCCL_UINT64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	return MDBCU_OK;
}

// ---- NCCL is plumbing here (statistics, IPC handles); the join's data moves through the arena below

int mdb_comm_allgather_bytes(mdbcu_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank)
{
	if (!ctx->nccl_comm)
		return mdb_fail(ctx, MDBCU_EERROR, "distributed plan without mdbcu_comm_init");
	NCCL_TRY(ctx, g_nccl.AllGather(send, recv, bytes_per_rank, NCCL_UINT8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	return MDBCU_OK;
}

// ---- peer-mapped exchange arena: one cudaMalloc'd block per rank, mapped into every other rank's address
// space with CUDA IPC, so a kernel can store straight into the owner GPU's memory over NVLink / NVSwitch.
// COLLECTIVE: every rank calls it with the same size; the mapping is cached until a larger one is needed.

#define MDB_ARENA_HEADER 256 // bytes in front of every arena: one barrier flag word per source rank

static void arena_release(mdbcu_ctx *ctx)
{
	for (int r = 0; r < MDB_MAX_RANKS; r++) {
		if (ctx->arena_peer[r] && r != ctx->rank)
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
		if (W > 1) {
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
struct ArenaBarrierArgs {
	uint32_t *peer_hdr[MDB_MAX_RANKS];
	int world, self;
	uint32_t epoch;
};

__global__ void k_arena_barrier(ArenaBarrierArgs a, const uint32_t *__restrict__ err, uint32_t *__restrict__ all)
{
	const int r = threadIdx.x;
	if (r >= a.world)
		return;
	const uint32_t mine = (a.epoch << 4) | (*err & 0xfu);
	__threadfence_system(); // everything this GPU wrote into peer memory before this kernel is visible first
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_hdr[r] + a.self), "r"(mine) : "memory");
	const uint32_t *slot = a.peer_hdr[a.self] + r;
	uint32_t v;
	do {
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
	} while ((v >> 4) < a.epoch);
	all[r] = v & 0xfu;
}

int mdb_comm_arena_barrier(mdbcu_ctx *ctx, const uint32_t *d_err, uint32_t *d_all)
{
	if (!ctx->arena_local)
		return mdb_fail(ctx, MDBCU_EERROR, "arena barrier without an arena");
	ArenaBarrierArgs a;
	memset(&a, 0, sizeof(a));
	for (int r = 0; r < ctx->world; r++)
		a.peer_hdr[r] = (uint32_t*)ctx->arena_peer[r];
	a.world = ctx->world;
	a.self = ctx->rank;
	a.epoch = ++ctx->arena_epoch;
	MDB_LAUNCH(ctx, k_arena_barrier, 1, 32, 0, a, d_err, d_all);
	return MDBCU_OK;
}

