// mdb_storage.cu - context lifetime and the columnar device mirror of the reference's row storage.
//
// K0 `k_unpack_pages` replaces the per-row AoS reads the reference does in cpy_cols
// (src/engine/executor_select.c:340-400): it de-interleaves uploaded 4 KiB page images
// (include/primitive/datablock.h:9-13, rows laid out by table_insert_row, src/primitive/row.c:26)
// into per-column 8-byte cell arrays + present/live bitmaps and gathers zone-map statistics
// (min / max / NULL count) on the way.
#include "mdb_common.cuh"

#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <new>

static thread_local std::string g_init_error;

void mdb_set_global_error(const char *msg)
{
	g_init_error = msg ? msg : "";
}

int mdb_fail(mdbcu_ctx *ctx, int code, const char *fmt, ...)
{
	char buf[1024]; // same size as query_output_error.message, include/engine/query.h:30-32
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx)
		ctx->err = buf;
	else
		g_init_error = buf;
	return code;
}

extern "C" const char *mdbcu_version(void)
{
	return "midoridb_cuda 0.1 (sm_100a)";
}

extern "C" const char *mdbcu_last_error(mdbcu_ctx *ctx)
{
	return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" int mdbcu_init(int device, mdbcu_ctx **out)
{
	int count = 0;
	cudaError_t e;

	if (!out)
		return mdb_fail(nullptr, MDBCU_EERROR, "mdbcu_init: out is NULL");
	*out = nullptr;

	e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return mdb_fail(nullptr, MDBCU_ECUDA, "mdbcu_init: no CUDA device (%s); there is no CPU fallback",
				e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
	if (device < 0 || device >= count)
		return mdb_fail(nullptr, MDBCU_EERROR, "mdbcu_init: device %d out of range (0..%d)", device, count - 1);

	mdbcu_ctx *ctx = new (std::nothrow) mdbcu_ctx();
	if (!ctx)
		return mdb_fail(nullptr, MDBCU_ENOMEM, "mdbcu_init: out of host memory");
	ctx->device = device;

	cudaDeviceProp prop;
	if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
		delete ctx;
		return mdb_fail(nullptr, MDBCU_ECUDA, "mdbcu_init: %s", cudaGetErrorString(e));
	}
	if (prop.major < 10) {
		delete ctx;
		return mdb_fail(nullptr, MDBCU_ECUDA, "mdbcu_init: device %d is sm_%d%d; this library is built for sm_100a only",
				device, prop.major, prop.minor);
	}
	ctx->num_sms = prop.multiProcessorCount;

	if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
			(e = cudaMallocHost((void**)&ctx->h_scalar, 64 * sizeof(uint64_t))) != cudaSuccess ||
			(e = cudaMalloc((void**)&ctx->d_scalar, 64 * sizeof(uint64_t))) != cudaSuccess) {
		delete ctx;
		return mdb_fail(nullptr, MDBCU_ECUDA, "mdbcu_init: %s", cudaGetErrorString(e));
	}

	// keep freed query temporaries cached in the stream-ordered pool: no cudaMalloc inside timed queries
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
		uint64_t threshold = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
	}

	*out = ctx;
	return MDBCU_OK;
}

extern "C" void mdbcu_shutdown(mdbcu_ctx *ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	mdb_comm_arena_destroy(ctx);
	mdb_comm_destroy(ctx);
	cudaFreeHost(ctx->h_scalar);
	cudaFree(ctx->d_scalar);
	cudaFree(ctx->scratch);
	if (ctx->side_stream) {
		cudaStreamDestroy(ctx->side_stream);
		cudaEventDestroy(ctx->side_ev[0]);
		cudaEventDestroy(ctx->side_ev[1]);
	}
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" int mdbcu_device_sync(mdbcu_ctx *ctx)
{
	if (!ctx)
		return MDBCU_EERROR;
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return MDBCU_OK;
}

int mdb_read_u64(mdbcu_ctx *ctx, const uint64_t *d_ptr, uint64_t *out)
{
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalar, d_ptr, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	*out = ctx->h_scalar[0];
	return MDBCU_OK;
}

// ------------------------------------------------------------------------------------------- tables

struct ColMeta {
	int64_t *data;
	uint32_t *present;
	int32_t row_off;
	int32_t width;
	int32_t is_int; // min/max statistics are kept for integer-like columns
	int32_t _pad;
};

// per-table device statistics block: [0] live rows, then per column {min, max, nulls}
struct TableStatsDev {
	unsigned long long live;
	long long cmin[MDBCU_MAX_COLUMNS];
	long long cmax[MDBCU_MAX_COLUMNS];
	unsigned long long nulls[MDBCU_MAX_COLUMNS];
};

struct TableExtra {
	ColMeta *d_meta = nullptr;
	TableStatsDev *d_stats = nullptr;
};

// TableExtra lives right behind the public struct so other translation units need not know it
struct TableImpl : mdbcu_table {
	TableExtra x;
};

static TableImpl *impl(mdbcu_table *t)
{
	return static_cast<TableImpl*>(t);
}

static bool type_is_intlike(int type)
{
	return type == MDBCU_CT_INTEGER || type == MDBCU_CT_TINYINT || type == MDBCU_CT_DATE || type == MDBCU_CT_DATETIME;
}

static int upload_meta(TableImpl *t)
{
	std::vector<ColMeta> m(t->ncols);
	for (int c = 0; c < t->ncols; c++) {
		m[c].data = t->cols[c].data;
		m[c].present = t->cols[c].present;
		m[c].row_off = t->cols[c].row_off;
		m[c].width = t->cols[c].width;
		m[c].is_int = type_is_intlike(t->cols[c].type);
		m[c]._pad = 0;
	}
	CUDA_TRY(t->ctx, cudaMemcpyAsync(t->x.d_meta, m.data(), sizeof(ColMeta) * t->ncols, cudaMemcpyHostToDevice,
			t->ctx->stream));
	// m goes out of scope: the copy from pageable memory is complete on return of cudaMemcpyAsync
	return MDBCU_OK;
}

__global__ void k_init_stats(TableStatsDev *s)
{
	int c = threadIdx.x;
	if (c == 0)
		s->live = 0;
	if (c < MDBCU_MAX_COLUMNS) {
		s->cmin[c] = INT64_MAX;
		s->cmax[c] = INT64_MIN;
		s->nulls[c] = 0;
	}
}

extern "C" int mdbcu_table_create(mdbcu_ctx *ctx, const char *name, int ncols, const int32_t *col_types,
		mdbcu_table **out)
{
	if (!ctx)
		return MDBCU_EERROR;
	if (!out || !col_types || ncols <= 0 || ncols > MDBCU_MAX_COLUMNS)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_create: bad arguments");
	cudaSetDevice(ctx->device);

	TableImpl *t = new (std::nothrow) TableImpl();
	if (!t)
		return mdb_fail(ctx, MDBCU_ENOMEM, "out of host memory");
	t->ctx = ctx;
	t->name = name ? name : "";
	t->ncols = ncols;
	t->cols.resize(ncols);
	int off = 0;
	for (int c = 0; c < ncols; c++) {
		if (col_types[c] < MDBCU_CT_VARCHAR || col_types[c] > MDBCU_CT_DATETIME) {
			delete t;
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_create: unknown column type %d", col_types[c]);
		}
		t->cols[c].type = col_types[c];
		t->cols[c].width = mdb_col_width(col_types[c]);
		t->cols[c].row_off = off;
		off += t->cols[c].width;
	}
	t->row_size = MDBCU_ROW_HEADER + off;                  // table_calc_row_size, row.c:21
	t->rows_per_page = (MDBCU_PAGE_SIZE - 1) / t->row_size; // insert rolls over when off+len >= 4096, row.c:38

	int rc = mdb_alloc(ctx, &t->x.d_meta, ncols);
	if (rc == MDBCU_OK)
		rc = mdb_alloc(ctx, &t->x.d_stats, 1);
	if (rc != MDBCU_OK) {
		delete t;
		return rc;
	}
	MDB_LAUNCH(ctx, k_init_stats, 1, MDBCU_MAX_COLUMNS, 0, t->x.d_stats);
	*out = t;
	return MDBCU_OK;
}

extern "C" void mdbcu_table_drop(mdbcu_table *tt)
{
	if (!tt)
		return;
	TableImpl *t = impl(tt);
	cudaSetDevice(t->ctx->device);
	for (auto &c : t->cols) {
		mdb_free(t->ctx, c.data);
		mdb_free(t->ctx, c.present);
	}
	mdb_free(t->ctx, t->live);
	mdb_free(t->ctx, t->x.d_meta);
	mdb_free(t->ctx, t->x.d_stats);
	delete t;
}

static size_t bitmap_words(uint64_t rows)
{
	return (size_t)((rows + 31) / 32) + 4; // slack so neighbouring-word atomics never run off the end
}

int mdb_table_reserve(mdbcu_table *tt, uint64_t rows)
{
	TableImpl *t = impl(tt);
	mdbcu_ctx *ctx = t->ctx;

	if (rows <= t->cap)
		return MDBCU_OK;
	if (rows >= (1ull << 32) - 64)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "tables are limited to 2^32-64 device rows (row ids are 32-bit)");
	uint64_t ncap = std::max<uint64_t>(rows, t->cap + t->cap / 2);
	ncap = (ncap + 1023) & ~1023ull;

	size_t old_words = t->cap ? bitmap_words(t->cap) : 0, new_words = bitmap_words(ncap);

	// every new buffer is allocated and filled BEFORE any pointer of the table changes: a failure half-way (out of memory)
	// frees what was allocated here and leaves the table exactly as it was
	std::vector<int64_t*> nd(t->cols.size(), nullptr);
	std::vector<uint32_t*> np(t->cols.size(), nullptr);
	uint32_t *nl = nullptr;
	auto grow = [&]() -> int {
		for (size_t i = 0; i < t->cols.size(); i++) {
			MDB_TRY(mdb_alloc(ctx, &nd[i], ncap));
			MDB_TRY(mdb_alloc(ctx, &np[i], new_words));
			CUDA_TRY(ctx, cudaMemsetAsync(np[i], 0, new_words * sizeof(uint32_t), ctx->stream));
			if (t->n_slots) {
				CUDA_TRY(ctx, cudaMemcpyAsync(nd[i], t->cols[i].data, t->n_slots * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
				CUDA_TRY(ctx, cudaMemcpyAsync(np[i], t->cols[i].present, old_words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
			}
		}
		MDB_TRY(mdb_alloc(ctx, &nl, new_words));
		CUDA_TRY(ctx, cudaMemsetAsync(nl, 0, new_words * sizeof(uint32_t), ctx->stream));
		if (t->n_slots)
			CUDA_TRY(ctx, cudaMemcpyAsync(nl, t->live, old_words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
		return MDBCU_OK;
	};
	const int grc = grow();
	if (grc != MDBCU_OK) {
		for (size_t i = 0; i < t->cols.size(); i++) {
			mdb_free(ctx, nd[i]);
			mdb_free(ctx, np[i]);
		}
		mdb_free(ctx, nl);
		return grc;
	}
	for (size_t i = 0; i < t->cols.size(); i++) {
		mdb_free(ctx, t->cols[i].data);
		mdb_free(ctx, t->cols[i].present);
		t->cols[i].data = nd[i];
		t->cols[i].present = np[i];
	}
	mdb_free(ctx, t->live);
	t->live = nl;
	t->cap = ncap;
	return upload_meta(t);
}

// pull the device statistics block and fold it into the host-side column descriptors
static int sync_stats(TableImpl *t)
{
	mdbcu_ctx *ctx = t->ctx;
	static_assert(sizeof(TableStatsDev) % 8 == 0, "stats block must be 8-byte granular");
	std::vector<uint64_t> raw(sizeof(TableStatsDev) / 8);

	CUDA_TRY(ctx, cudaMemcpyAsync(raw.data(), t->x.d_stats, sizeof(TableStatsDev), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	const TableStatsDev *s = reinterpret_cast<const TableStatsDev*>(raw.data());
	t->all_live = (s->live == t->n_slots);
	for (int c = 0; c < t->ncols; c++) {
		DevColumn &col = t->cols[c];
		col.has_nulls = s->nulls[c] != 0;
		if (type_is_intlike(col.type)) {
			col.stats_ok = true;
			col.imin = s->cmin[c];
			col.imax = s->cmax[c];
		}
	}
	return MDBCU_OK;
}

// ----------------------------------------------------------------------------------- K0 unpack

#define UNPACK_THREADS 192 // >= 4096/25 slots (one TINYINT column is the narrowest row)

__device__ static inline int64_t smem_read_cell(const unsigned char *p, int width)
{
	if (width == 1)
		return (int64_t)p[0];
	if ((((uintptr_t)p) & 7) == 0)
		return *(const int64_t*)p;
	uint64_t v = 0;
#pragma unroll
	for (int b = 0; b < 8; b++)
		v |= (uint64_t)p[b] << (8 * b);
	return (int64_t)v;
}

__device__ static inline void bitmap_store_range(uint32_t *bm, uint64_t first_bit, uint32_t bits, uint32_t mask)
{
	// write `bits` (valid where `mask`) at bit offset first_bit; ranges of different warps are disjoint
	uint64_t w = first_bit >> 5;
	uint32_t sh = (uint32_t)(first_bit & 31);
	uint32_t lo_mask = mask << sh, lo_bits = bits << sh;
	if (lo_mask) {
		atomicAnd(&bm[w], ~lo_mask);
		if (lo_bits)
			atomicOr(&bm[w], lo_bits);
	}
	if (sh) {
		uint32_t hi_mask = mask >> (32 - sh), hi_bits = bits >> (32 - sh);
		if (hi_mask) {
			atomicAnd(&bm[w + 1], ~hi_mask);
			if (hi_bits)
				atomicOr(&bm[w + 1], hi_bits);
		}
	}
}

__global__ void __launch_bounds__(UNPACK_THREADS)
k_unpack_pages(const unsigned char *__restrict__ pages, uint64_t n_pages, uint64_t first_page, int row_size,
		int rows_per_page, int ncols, const ColMeta *__restrict__ meta, uint32_t *__restrict__ live,
		TableStatsDev *__restrict__ stats, int count_stats)
{
	__shared__ __align__(16) unsigned char page[MDBCU_PAGE_SIZE];
	__shared__ int first_empty;
	__shared__ long long s_min, s_max;
	__shared__ unsigned int s_nulls, s_live;

	const int slots = MDBCU_PAGE_SIZE / row_size; // the executor's loop bound, executor_select.c:1099
	const int s = threadIdx.x;
	const int lane = threadIdx.x & 31;

	for (uint64_t p = blockIdx.x; p < n_pages; p += gridDim.x) {
		const int4 *src = reinterpret_cast<const int4*>(pages + p * MDBCU_PAGE_SIZE);
		for (int i = threadIdx.x; i < MDBCU_PAGE_SIZE / 16; i += blockDim.x)
			reinterpret_cast<int4*>(page)[i] = mdb_ldg_stream(src + i);
		if (threadIdx.x == 0) {
			first_empty = slots;
			s_live = 0;
		}
		__syncthreads();

		const unsigned char *row = page + (size_t)s * row_size;
		bool in_page = s < slots;
		if (in_page && row[0]) // flags.empty: the executor stops at the first empty slot
			atomicMin(&first_empty, s);
		__syncthreads();

		bool is_live = in_page && s < rows_per_page && s < first_empty && !row[1] /* flags.deleted */;
		uint64_t base = (first_page + p) * (uint64_t)rows_per_page;
		uint32_t warp_first = (uint32_t)(s - lane);
		uint32_t range_mask = 0;
		if ((int)warp_first < rows_per_page) {
			int nvalid = min(32, rows_per_page - (int)warp_first);
			range_mask = nvalid == 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
		}

		uint32_t live_bits = __ballot_sync(0xffffffffu, is_live);
		if (lane == 0 && range_mask) {
			bitmap_store_range(live, base + warp_first, live_bits, range_mask);
			if (count_stats)
				atomicAdd(&s_live, __popc(live_bits));
		}

		for (int c = 0; c < ncols; c++) {
			ColMeta m = meta[c];
			if (threadIdx.x == 0) {
				s_min = INT64_MAX;
				s_max = INT64_MIN;
				s_nulls = 0;
			}
			__syncthreads();
			bool isnull = false;
			int64_t v = 0;
			if (is_live) {
				// bit_test(row->null_bitmap, c), src/lib/bit.c:3
				isnull = (row[MDBCU_NULL_BITMAP_OFF + (c >> 3)] >> (c & 7)) & 1;
				if (!isnull)
					v = smem_read_cell(row + MDBCU_ROW_HEADER + m.row_off, m.width);
			}
			bool present = is_live && !isnull;
			if (s < rows_per_page)
				m.data[base + s] = v;
			uint32_t pbits = __ballot_sync(0xffffffffu, present);
			if (lane == 0 && range_mask)
				bitmap_store_range(m.present, base + warp_first, pbits, range_mask);
			if (count_stats) {
				if (m.is_int) {
					long long lo = present ? v : INT64_MAX, hi = present ? v : INT64_MIN;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) {
						lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
						hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
					}
					if (lane == 0 && pbits) {
						atomicMin(&s_min, lo);
						atomicMax(&s_max, hi);
					}
				}
				uint32_t nbits = __ballot_sync(0xffffffffu, is_live && isnull);
				if (lane == 0 && nbits)
					atomicAdd(&s_nulls, __popc(nbits));
				__syncthreads();
				if (threadIdx.x == 0) {
					if (m.is_int && s_min <= s_max) {
						atomicMin(&stats->cmin[c], s_min);
						atomicMax(&stats->cmax[c], s_max);
					}
					if (s_nulls)
						atomicAdd(&stats->nulls[c], (unsigned long long)s_nulls);
				}
			}
			__syncthreads();
		}
		if (count_stats && threadIdx.x == 0 && s_live)
			atomicAdd(&stats->live, (unsigned long long)s_live);
		__syncthreads();
	}
}

// recount live rows / NULLs after tombstones or page reloads (bitmaps are the source of truth)
__global__ void k_recount(const uint32_t *__restrict__ live, const ColMeta *__restrict__ meta, int ncols, uint64_t n_slots,
		TableStatsDev *__restrict__ stats)
{
	uint64_t words = (n_slots + 31) / 32;
	unsigned long long nlive = 0;
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < words; w += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t valid = (w == words - 1 && (n_slots & 31)) ? ((1u << (n_slots & 31)) - 1u) : 0xffffffffu;
		uint32_t l = live[w] & valid;
		nlive += __popc(l);
		for (int c = 0; c < ncols; c++) {
			uint32_t nulls = l & ~meta[c].present[w];
			if (nulls)
				atomicAdd(&stats->nulls[c], (unsigned long long)__popc(nulls));
		}
	}
	for (int o = 16; o > 0; o >>= 1)
		nlive += __shfl_xor_sync(0xffffffffu, nlive, o);
	if ((threadIdx.x & 31) == 0 && nlive)
		atomicAdd(&stats->live, nlive);
}

__global__ void k_reset_counts(TableStatsDev *s, int ncols)
{
	int c = threadIdx.x;
	if (c == 0)
		s->live = 0;
	if (c < ncols)
		s->nulls[c] = 0;
}

static int recount(TableImpl *t)
{
	mdbcu_ctx *ctx = t->ctx;
	MDB_LAUNCH(ctx, k_reset_counts, 1, MDBCU_MAX_COLUMNS, 0, t->x.d_stats, t->ncols);
	if (t->n_slots) {
		int grid = (int)std::min<uint64_t>(mdb_div_up((t->n_slots + 31) / 32, 256), (uint64_t)ctx->num_sms * 8);
		MDB_LAUNCH(ctx, k_recount, grid, 256, 0, t->live, t->x.d_meta, t->ncols, t->n_slots, t->x.d_stats);
	}
	CUDA_CHECK_LAUNCH(ctx);
	return sync_stats(t);
}

#define STAGE_PAGES 16384 // 64 MiB staging chunks

// upload pages [0, n_pages) (fetched through `fetch`) to device pages starting at first_page and unpack them
template <typename Fetch>
static int upload_and_unpack(TableImpl *t, size_t n_pages, uint64_t first_page, bool count_stats, Fetch fetch)
{
	mdbcu_ctx *ctx = t->ctx;
	if (n_pages == 0)
		return MDBCU_OK;

	size_t chunk = std::min<size_t>(n_pages, STAGE_PAGES);
	unsigned char *d_stage[2] = {nullptr, nullptr};
	cudaEvent_t done[2] = {nullptr, nullptr};
	DevTemp tmp(ctx);
	MDB_TRY(tmp.alloc(&d_stage[0], chunk * MDBCU_PAGE_SIZE));
	if (n_pages > chunk)
		MDB_TRY(tmp.alloc(&d_stage[1], chunk * MDBCU_PAGE_SIZE));
	int rc = MDBCU_OK;

	for (size_t p0 = 0, it = 0; p0 < n_pages; p0 += chunk, it++) {
		size_t n = std::min(chunk, n_pages - p0);
		int b = (int)(it & 1);
		if (done[b]) {
			// the previous unpack reading this staging buffer is ordered before us on the same stream
		}
		rc = fetch(p0, n, d_stage[b]);
		if (rc != MDBCU_OK)
			break;
		int grid = (int)std::min<size_t>(n, (size_t)ctx->num_sms * 10);
		MDB_LAUNCH(ctx, k_unpack_pages, grid, UNPACK_THREADS, 0, d_stage[b], (uint64_t)n, first_page + p0,
				(int)t->row_size, (int)t->rows_per_page, t->ncols, t->x.d_meta, t->live, t->x.d_stats,
				count_stats ? 1 : 0);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) {
			rc = mdb_fail(ctx, MDBCU_ECUDA, "k_unpack_pages launch failed: %s", cudaGetErrorString(e));
			break;
		}
	}
	return rc;
}

static int append_pages_common(TableImpl *t, size_t n_pages, const void *contig, size_t stride, const void *const *ptrs)
{
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	if (n_pages == 0)
		return MDBCU_OK;
	if (!t->paged && t->n_slots)
		return mdb_fail(ctx, MDBCU_EERROR, "table '%s' was bulk-loaded; page appends need a page-mirrored table",
				t->name.c_str());
	t->paged = true;

	uint64_t first_page = t->n_pages;
	uint64_t new_slots = (first_page + n_pages) * t->rows_per_page;
	MDB_TRY(mdb_table_reserve(t, new_slots));
	t->n_slots = new_slots;
	t->n_pages = first_page + n_pages;

	unsigned char *h_stage = nullptr;
	if (ptrs) {
		size_t chunk = std::min<size_t>(n_pages, STAGE_PAGES);
		if (cudaMallocHost((void**)&h_stage, chunk * MDBCU_PAGE_SIZE) != cudaSuccess)
			return mdb_fail(ctx, MDBCU_ENOMEM, "cannot allocate pinned staging memory");
	}

	int rc = upload_and_unpack(t, n_pages, first_page, true, [&](size_t p0, size_t n, unsigned char *dst) -> int {
		if (ptrs) {
			// the reference mallocs every datablock separately (datablock.c:17): gather into pinned staging.
			// The staging buffer is reused, so wait for the previous chunk's DMA before overwriting it.
			CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
			for (size_t i = 0; i < n; i++)
				memcpy(h_stage + i * MDBCU_PAGE_SIZE, ptrs[p0 + i], MDBCU_PAGE_SIZE);
			CUDA_TRY(ctx, cudaMemcpyAsync(dst, h_stage, n * MDBCU_PAGE_SIZE, cudaMemcpyHostToDevice, ctx->stream));
		} else if (stride == MDBCU_PAGE_SIZE) {
			CUDA_TRY(ctx, cudaMemcpyAsync(dst, (const unsigned char*)contig + p0 * stride, n * MDBCU_PAGE_SIZE,
					cudaMemcpyHostToDevice, ctx->stream));
		} else {
			CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, MDBCU_PAGE_SIZE, (const unsigned char*)contig + p0 * stride, stride,
					MDBCU_PAGE_SIZE, n, cudaMemcpyHostToDevice, ctx->stream));
		}
		return MDBCU_OK;
	});
	if (rc == MDBCU_OK)
		rc = sync_stats(t);
	else
		cudaStreamSynchronize(ctx->stream);
	if (h_stage)
		cudaFreeHost(h_stage);
	return rc;
}

extern "C" int mdbcu_table_append_pages(mdbcu_table *t, const void *pages, size_t n_pages, size_t page_stride)
{
	if (!t)
		return MDBCU_EERROR;
	t->version++; // cached per-column facts (sortedness) are stale from here on
	if ((!pages && n_pages) || page_stride < MDBCU_PAGE_SIZE)
		return mdb_fail(t->ctx, MDBCU_EERROR, "mdbcu_table_append_pages: bad arguments");
	return append_pages_common(impl(t), n_pages, pages, page_stride, nullptr);
}

extern "C" int mdbcu_table_append_page_ptrs(mdbcu_table *t, const void *const *page_ptrs, size_t n_pages)
{
	if (!t)
		return MDBCU_EERROR;
	t->version++; // cached per-column facts (sortedness) are stale from here on
	if (!page_ptrs && n_pages)
		return mdb_fail(t->ctx, MDBCU_EERROR, "mdbcu_table_append_page_ptrs: bad arguments");
	return append_pages_common(impl(t), n_pages, nullptr, 0, page_ptrs);
}

extern "C" int mdbcu_table_reload_pages(mdbcu_table *tt, size_t first_page, const void *const *page_ptrs, size_t n_pages)
{
	if (!tt)
		return MDBCU_EERROR;
	tt->version++; // cached per-column facts (sortedness) are stale from here on
	TableImpl *t = impl(tt);
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	if (!t->paged || first_page + n_pages > t->n_pages || (!page_ptrs && n_pages))
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_reload_pages: page range outside the mirror");
	if (n_pages == 0)
		return MDBCU_OK;

	unsigned char *h_stage = nullptr;
	size_t chunk = std::min<size_t>(n_pages, STAGE_PAGES);
	if (cudaMallocHost((void**)&h_stage, chunk * MDBCU_PAGE_SIZE) != cudaSuccess)
		return mdb_fail(ctx, MDBCU_ENOMEM, "cannot allocate pinned staging memory");
	// min/max only ever widen here (still conservative bounds); counts are recomputed from the bitmaps
	int rc = upload_and_unpack(t, n_pages, first_page, true, [&](size_t p0, size_t n, unsigned char *dst) -> int {
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		for (size_t i = 0; i < n; i++)
			memcpy(h_stage + i * MDBCU_PAGE_SIZE, page_ptrs[p0 + i], MDBCU_PAGE_SIZE);
		CUDA_TRY(ctx, cudaMemcpyAsync(dst, h_stage, n * MDBCU_PAGE_SIZE, cudaMemcpyHostToDevice, ctx->stream));
		return MDBCU_OK;
	});
	if (rc == MDBCU_OK)
		rc = recount(t);
	else
		cudaStreamSynchronize(ctx->stream);
	cudaFreeHost(h_stage);
	return rc;
}

__global__ void k_tombstone(const uint64_t *__restrict__ page_idx, const uint32_t *__restrict__ slot_idx, size_t n,
		uint64_t rows_per_page, uint64_t n_slots, uint32_t *__restrict__ live, const ColMeta *__restrict__ meta, int ncols)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	uint64_t r = page_idx[i] * rows_per_page + slot_idx[i];
	if (slot_idx[i] >= rows_per_page || r >= n_slots)
		return;
	uint32_t m = ~(1u << (r & 31));
	atomicAnd(&live[r >> 5], m); // flags.deleted = true, row.c:137
	for (int c = 0; c < ncols; c++)
		atomicAnd(&meta[c].present[r >> 5], m);
}

extern "C" int mdbcu_table_tombstone(mdbcu_table *tt, const uint64_t *page_idx, const uint32_t *slot_idx, size_t n)
{
	if (!tt)
		return MDBCU_EERROR;
	tt->version++; // cached per-column facts (sortedness) are stale from here on
	TableImpl *t = impl(tt);
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	if (n == 0)
		return MDBCU_OK;
	if (!page_idx || !slot_idx)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_tombstone: bad arguments");
	// bulk-loaded tables address rows as page 0, slot = row index
	uint64_t rpp = t->paged ? t->rows_per_page : (1ull << 32);
	DevTemp tmp(ctx);
	uint64_t *d_p;
	uint32_t *d_s;
	MDB_TRY(tmp.alloc(&d_p, n));
	MDB_TRY(tmp.alloc(&d_s, n));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_p, page_idx, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_s, slot_idx, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
	MDB_LAUNCH(ctx, k_tombstone, (unsigned)mdb_div_up(n, 256), 256, 0, d_p, d_s, n, rpp, t->n_slots, t->live, t->x.d_meta,
			t->ncols);
	CUDA_CHECK_LAUNCH(ctx);
	return recount(t);
}

// ----------------------------------------------------------------------------------- bulk columnar load

// finish one bulk-loaded column: present bitmap from byte flags (or all present), zero NULL cells, statistics
__global__ void k_finish_column(int64_t *__restrict__ data, uint32_t *__restrict__ present, const uint8_t *__restrict__ nulls,
		uint64_t first, uint64_t n, int is_int, int c, TableStatsDev *__restrict__ stats)
{
	__shared__ long long s_min, s_max;
	__shared__ unsigned int s_nulls;
	if (threadIdx.x == 0) {
		s_min = INT64_MAX;
		s_max = INT64_MIN;
		s_nulls = 0;
	}
	__syncthreads();
	long long lo = INT64_MAX, hi = INT64_MIN;
	unsigned int nn = 0;
	// every warp iteration covers one aligned 32-row group so the ballot maps to one bitmap word (first % 32 == 0)
	uint64_t groups = (n + 31) / 32;
	uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	int lane = threadIdx.x & 31;
	for (uint64_t g = warp; g < groups; g += nwarps) {
		uint64_t i = g * 32 + lane;
		bool in = i < n;
		bool isnull = in && nulls && nulls[i];
		int64_t v = 0;
		if (in && !isnull) {
			v = data[first + i];
			lo = min(lo, (long long)v);
			hi = max(hi, (long long)v);
		}
		if (isnull) {
			data[first + i] = 0;
			nn++;
		}
		uint32_t bits = __ballot_sync(0xffffffffu, in && !isnull);
		if (lane == 0)
			present[(first >> 5) + g] = bits;
	}
	for (int o = 16; o > 0; o >>= 1) {
		lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
		hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
		nn += __shfl_xor_sync(0xffffffffu, nn, o);
	}
	if (lane == 0) {
		if (is_int && lo <= hi) {
			atomicMin(&s_min, lo);
			atomicMax(&s_max, hi);
		}
		if (nn)
			atomicAdd(&s_nulls, nn);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (is_int && s_min <= s_max) {
			atomicMin(&stats->cmin[c], s_min);
			atomicMax(&stats->cmax[c], s_max);
		}
		if (s_nulls)
			atomicAdd(&stats->nulls[c], (unsigned long long)s_nulls);
	}
}

__global__ void k_fill_bitmap(uint32_t *__restrict__ bm, uint64_t first, uint64_t n)
{
	// set bits [first, first+n), first % 32 == 0
	uint64_t words = (n + 31) / 32;
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < words; w += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t v = (w == words - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
		bm[(first >> 5) + w] = v;
	}
}

__global__ void k_add_live(TableStatsDev *s, unsigned long long n)
{
	s->live += n;
}

static int bulk_prepare(TableImpl *t, uint64_t n_rows, uint64_t *first_out)
{
	mdbcu_ctx *ctx = t->ctx;
	if (t->paged)
		return mdb_fail(ctx, MDBCU_EERROR, "table '%s' mirrors pages; bulk loads need a separate table", t->name.c_str());
	if (t->n_slots & 31)
		return mdb_fail(ctx, MDBCU_EUNSUPPORTED, "bulk loads must append at a multiple of 32 rows (have %llu)",
				(unsigned long long)t->n_slots);
	*first_out = t->n_slots;
	MDB_TRY(mdb_table_reserve(t, t->n_slots + n_rows));
	return MDBCU_OK;
}

extern "C" int mdbcu_table_append_columns(mdbcu_table *tt, size_t n_rows, const void *const *col_data,
		const uint8_t *const *col_nulls)
{
	if (!tt)
		return MDBCU_EERROR;
	tt->version++; // cached per-column facts (sortedness) are stale from here on
	TableImpl *t = impl(tt);
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	if (n_rows == 0)
		return MDBCU_OK;
	if (!col_data)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_append_columns: bad arguments");
	uint64_t first = 0;
	MDB_TRY(bulk_prepare(t, n_rows, &first));

	DevTemp tmp(ctx);
	int grid = (int)std::min<uint64_t>(mdb_div_up(n_rows, 256), (uint64_t)ctx->num_sms * 8);
	for (int c = 0; c < t->ncols; c++) {
		DevColumn &col = t->cols[c];
		uint8_t *d_nulls = nullptr;
		if (!col_data[c])
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_append_columns: column %d has no data", c);
		// (cudaMemcpyDefault: the arrays may be host or device memory - mdb_dist_group.cu appends device-resident partials)
		CUDA_TRY(ctx, cudaMemcpyAsync(col.data + first, col_data[c], n_rows * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
		if (col_nulls && col_nulls[c]) {
			MDB_TRY(tmp.alloc(&d_nulls, n_rows));
			CUDA_TRY(ctx, cudaMemcpyAsync(d_nulls, col_nulls[c], n_rows, cudaMemcpyDefault, ctx->stream));
		}
		MDB_LAUNCH(ctx, k_finish_column, grid, 256, 0, col.data, col.present, d_nulls, first, (uint64_t)n_rows,
				type_is_intlike(col.type) ? 1 : 0, c, t->x.d_stats);
	}
	MDB_LAUNCH(ctx, k_fill_bitmap, grid, 256, 0, t->live, first, (uint64_t)n_rows);
	MDB_LAUNCH(ctx, k_add_live, 1, 1, 0, t->x.d_stats, (unsigned long long)n_rows);
	CUDA_CHECK_LAUNCH(ctx);
	t->n_slots = first + n_rows;
	return sync_stats(t);
}

// ----------------------------------------------------------------------------------- synthetic generator

__device__ static inline uint64_t gen_hash(uint64_t seed, uint64_t idx)
{
	return mdb_mix64(seed * 0x9e3779b97f4a7c15ULL + mdb_mix64(idx + 0x632be59bd9b4e019ULL));
}

// bijection on k-bit integers (k >= 1): multiply by odd constants and xorshift, both invertible mod 2^k
__host__ __device__ static inline uint64_t gen_permute(uint64_t x, int k, uint64_t seed)
{
	uint64_t mask = k >= 64 ? ~0ull : ((1ull << k) - 1);
	int sh = k > 1 ? (k + 1) / 2 : 1;
	uint64_t a = (mdb_mix64(seed) | 1ull), b = (mdb_mix64(seed + 1) | 1ull), cst = mdb_mix64(seed + 2);
	x = (x * a + cst) & mask;
	x ^= x >> sh;
	x = (x * b) & mask;
	x ^= x >> sh;
	x = (x * 0x9e3779b97f4a7c15ULL) & mask;
	x ^= x >> sh;
	return x & mask;
}

__global__ void k_generate(int64_t *__restrict__ data, uint64_t first, uint64_t n, uint64_t row_offset, mdbcu_gen_spec spec,
		int perm_bits)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t g = row_offset + i;
		uint64_t h = gen_hash(spec.seed, g);
		int64_t v = 0;
		uint64_t range = (uint64_t)(spec.hi - spec.lo) + 1ull; // 0 means the full 2^64 range
		switch (spec.kind) {
		case MDBCU_GEN_UNIFORM_INT:
			v = spec.lo + (int64_t)(range ? __umul64hi(h, range) : h);
			break;
		case MDBCU_GEN_UNIFORM_DBL: {
			double d = (double)(h >> 11) * (1.0 / 9007199254740992.0);
			v = __double_as_longlong(d);
			break;
		}
		case MDBCU_GEN_PERMUTATION:
			v = spec.lo + (int64_t)gen_permute(g & (range - 1), perm_bits, spec.seed);
			break;
		case MDBCU_GEN_ZIPF: {
			// bounded-Pareto inverse CDF: rank in [1, N], P(rank) ~ rank^-s ; s = spec.param > 1
			double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
			double N = (double)range, s1 = 1.0 - spec.param;
			double x = pow((pow(N + 1.0, s1) - 1.0) * u + 1.0, 1.0 / s1);
			uint64_t rank = (uint64_t)x;
			if (rank < 1) rank = 1;
			if (rank > range) rank = range;
			v = spec.lo + (int64_t)(rank - 1);
			break;
		}
		case MDBCU_GEN_SEQUENCE:
			v = spec.lo + (int64_t)g;
			break;
		}
		data[first + i] = v;
	}
}

__global__ void k_generate_nulls(uint8_t *__restrict__ nulls, uint64_t n, uint64_t row_offset, uint64_t seed, int permille)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t h = gen_hash(seed ^ 0x5bd1e995u, row_offset + i);
		nulls[i] = (h % 1000u) < (uint64_t)permille;
	}
}

extern "C" int mdbcu_table_generate(mdbcu_table *tt, uint64_t n_rows, uint64_t row_offset, const struct mdbcu_gen_spec *specs)
{
	if (!tt)
		return MDBCU_EERROR;
	tt->version++; // cached per-column facts (sortedness) are stale from here on
	TableImpl *t = impl(tt);
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	if (n_rows == 0)
		return MDBCU_OK;
	if (!specs)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_generate: bad arguments");
	uint64_t first = 0;
	MDB_TRY(bulk_prepare(t, n_rows, &first));

	DevTemp tmp(ctx);
	int grid = (int)std::min<uint64_t>(mdb_div_up(n_rows, 256), (uint64_t)ctx->num_sms * 16);
	for (int c = 0; c < t->ncols; c++) {
		DevColumn &col = t->cols[c];
		mdbcu_gen_spec sp = specs[c];
		int perm_bits = 0;
		if (sp.kind < MDBCU_GEN_UNIFORM_INT || sp.kind > MDBCU_GEN_SEQUENCE)
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_generate: unknown generator %d", sp.kind);
		if (sp.kind != MDBCU_GEN_UNIFORM_DBL && sp.kind != MDBCU_GEN_SEQUENCE && sp.hi < sp.lo)
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_generate: hi < lo for column %d", c);
		if (sp.kind == MDBCU_GEN_PERMUTATION) {
			uint64_t range = (uint64_t)(sp.hi - sp.lo) + 1ull;
			if (range == 0 || (range & (range - 1)))
				return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_generate: permutation range must be a power of two");
			while ((1ull << perm_bits) < range)
				perm_bits++;
			if (perm_bits == 0)
				perm_bits = 1;
		}
		if (sp.kind == MDBCU_GEN_ZIPF && !(sp.param > 1.0))
			return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_generate: Zipf exponent must be > 1");
		MDB_LAUNCH(ctx, k_generate, grid, 256, 0, col.data, first, n_rows, row_offset, sp, perm_bits);
		uint8_t *d_nulls = nullptr;
		if (sp.null_permille > 0) {
			MDB_TRY(tmp.alloc(&d_nulls, n_rows));
			MDB_LAUNCH(ctx, k_generate_nulls, grid, 256, 0, d_nulls, n_rows, row_offset, sp.seed, sp.null_permille);
		}
		MDB_LAUNCH(ctx, k_finish_column, grid, 256, 0, col.data, col.present, d_nulls, first, n_rows,
				type_is_intlike(col.type) ? 1 : 0, c, t->x.d_stats);
	}
	MDB_LAUNCH(ctx, k_fill_bitmap, grid, 256, 0, t->live, first, n_rows);
	MDB_LAUNCH(ctx, k_add_live, 1, 1, 0, t->x.d_stats, (unsigned long long)n_rows);
	CUDA_CHECK_LAUNCH(ctx);
	t->n_slots = first + n_rows;
	return sync_stats(t);
}

int mdb_table_refresh_stats(mdbcu_table *t, int col)
{
	(void)col;
	return sync_stats(impl(t));
}

// collective: all-gather every rank's {min, max, has-data} per column and fold them into global bounds
extern "C" int mdbcu_table_sync_stats(mdbcu_table *tt)
{
	if (!tt)
		return MDBCU_EERROR;
	TableImpl *t = impl(tt);
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	const int W = ctx->world;
	if (W == 1 && !mdb_comm_ready(ctx)) {
		for (auto &c : t->cols) {
			c.gstats_ok = c.stats_ok;
			c.gmin = c.imin;
			c.gmax = c.imax;
		}
		t->global_slots = t->n_slots;
		return MDBCU_OK;
	}
	const size_t per_rank = (size_t)3 * t->ncols + 1;
	std::vector<int64_t> mine(per_rank), all(per_rank * W);
	for (int c = 0; c < t->ncols; c++) {
		mine[3 * c] = t->cols[c].imin;
		mine[3 * c + 1] = t->cols[c].imax;
		mine[3 * c + 2] = t->cols[c].stats_ok ? 1 : 0;
	}
	mine[per_rank - 1] = (int64_t)t->n_slots;
	DevTemp tmp(ctx);
	int64_t *d_mine, *d_all;
	MDB_TRY(tmp.alloc(&d_mine, per_rank));
	MDB_TRY(tmp.alloc(&d_all, per_rank * W));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_mine, mine.data(), per_rank * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
	MDB_TRY(mdb_comm_allgather_bytes(ctx, d_mine, d_all, per_rank * sizeof(int64_t)));
	CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), d_all, per_rank * W * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	t->global_slots = 0;
	for (int r = 0; r < W; r++)
		t->global_slots += (uint64_t)all[(size_t)r * per_rank + per_rank - 1];
	for (int c = 0; c < t->ncols; c++) {
		DevColumn &col = t->cols[c];
		col.gstats_ok = true;
		col.gmin = INT64_MAX;
		col.gmax = INT64_MIN;
		for (int r = 0; r < W; r++) {
			const int64_t *p = &all[(size_t)r * per_rank + 3 * c];
			if (!p[2])
				col.gstats_ok = false;
			col.gmin = std::min(col.gmin, p[0]);
			col.gmax = std::max(col.gmax, p[1]);
		}
	}
	return MDBCU_OK;
}

// ----------------------------------------------------------------------------------- inspection

extern "C" uint64_t mdbcu_table_slots(const mdbcu_table *t)
{
	return t ? t->n_slots : 0;
}

__global__ void k_popcount(const uint32_t *__restrict__ bm, uint64_t n_bits, unsigned long long *out)
{
	uint64_t words = (n_bits + 31) / 32;
	unsigned long long acc = 0;
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < words; w += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t valid = (w == words - 1 && (n_bits & 31)) ? ((1u << (n_bits & 31)) - 1u) : 0xffffffffu;
		acc += __popc(bm[w] & valid);
	}
	for (int o = 16; o > 0; o >>= 1)
		acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc)
		atomicAdd(out, acc);
}

extern "C" uint64_t mdbcu_table_live_rows(mdbcu_table *t)
{
	if (!t || !t->n_slots)
		return 0;
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	if (t->all_live)
		return t->n_slots;
	cudaMemsetAsync(ctx->d_scalar, 0, sizeof(uint64_t), ctx->stream);
	int grid = (int)std::min<uint64_t>(mdb_div_up((t->n_slots + 31) / 32, 256), (uint64_t)ctx->num_sms * 8);
	MDB_LAUNCH(ctx, k_popcount, grid, 256, 0, t->live, t->n_slots, (unsigned long long*)ctx->d_scalar);
	uint64_t v = 0;
	if (mdb_read_u64(ctx, ctx->d_scalar, &v) != MDBCU_OK)
		return 0;
	return v;
}

extern "C" int mdbcu_table_column_device_ptr(mdbcu_table *t, int col, const void **cells, uint64_t *n_slots)
{
	if (!t || !cells)
		return MDBCU_EERROR;
	if (col < 0 || col >= t->ncols)
		return mdb_fail(t->ctx, MDBCU_EERROR, "mdbcu_table_column_device_ptr: no column %d", col);
	cudaSetDevice(t->ctx->device);
	cudaStreamSynchronize(t->ctx->stream); // loads issued so far are complete: any stream may read the column
	*cells = t->cols[col].data;
	if (n_slots)
		*n_slots = t->n_slots;
	return MDBCU_OK;
}

__global__ void k_expand_bits(const uint32_t *__restrict__ bm, uint64_t first, uint64_t n, uint8_t *__restrict__ out)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = mdb_bit(bm, first + i);
}

extern "C" int mdbcu_table_read_column(mdbcu_table *t, int col, uint64_t first, uint64_t n, void *cells, uint8_t *valid)
{
	if (!t)
		return MDBCU_EERROR;
	mdbcu_ctx *ctx = t->ctx;
	cudaSetDevice(ctx->device);
	if (col < 0 || col >= t->ncols || first + n > t->n_slots)
		return mdb_fail(ctx, MDBCU_EERROR, "mdbcu_table_read_column: out of range");
	if (n == 0)
		return MDBCU_OK;
	if (cells)
		CUDA_TRY(ctx, cudaMemcpyAsync(cells, t->cols[col].data + first, n * sizeof(int64_t), cudaMemcpyDeviceToHost,
				ctx->stream));
	if (valid) {
		DevTemp tmp(ctx);
		uint8_t *d;
		MDB_TRY(tmp.alloc(&d, n));
		int grid = (int)std::min<uint64_t>(mdb_div_up(n, 256), (uint64_t)ctx->num_sms * 8);
		MDB_LAUNCH(ctx, k_expand_bits, grid, 256, 0, t->cols[col].present, first, n, d);
		CUDA_TRY(ctx, cudaMemcpyAsync(valid, d, n, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	}
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return MDBCU_OK;
}
