// mdb_radix_types.cuh - constants and kernel-parameter structs of the radix join (mdb_radix.cu and its pass headers)
#pragma once
#include <stdint.h>

#define RJ_MAX_PART 4096           // partitions (12 radix bits)
#define RJ_MAX_SHIFT 16            // remainder bits: 2-byte remainders
#define RJ_CAP 20                  // staging slots per partition in shared memory
#define RJ_FLUSH 16                // a partition is flushed when 16 remainders (= one 32-byte sector) are staged
#define RJ_CHUNK 256               // remainders per chunk (512 bytes = one warp-wide 128-bit load)
#define RJ_BLOCKS_PER_CHUNK (RJ_CHUNK / RJ_FLUSH)
#define RJ_P1_THREADS 1024
#define RJ_OVF_CAP 512             // keys per round that may find their staging row full and wait one round
#define RJ_NONE 0xffffffffu

#define RJ_ERR_POOL 1u             // chunk pool exhausted
#define RJ_ERR_COUNTER 2u          // a packed counter wrapped (too many equal keys for the counter width)
#define RJ_ERR_SKEW 4u             // more than RJ_OVF_CAP keys per round hit full staging rows

// one run of <= RJ_CHUNK remainders: 16-byte aligned offset into a remainder buffer, valid entries
struct RJDesc {
	uint32_t off16; // in units of 16 bytes (8 remainders)
	uint32_t ne;
};

#define RJ_MAX_RANKS 8

// where the chunks of the partitions owned by one rank are written: this GPU's own arrays, or - in a
// multi-GPU plan - the owner's arena mapped over NVLink (CUDA IPC), so pass 1 IS the exchange
struct RJTarget {
	uint16_t *pool;            // pool_chunks * RJ_CHUNK remainders
	uint32_t *pool_next;       // allocation cursor
	uint16_t *chunk_part;      // partition of each chunk
	uint16_t *chunk_entries;   // valid remainders in each chunk
	uint32_t *dir_cnt;         // chunks per partition
};

struct RJSide {
	const int64_t *keys;
	const uint32_t *present;
	uint64_t n;
	int all_in_range;          // every key of the column lies in [kmin, kmin + range): no per-key range test
	uint32_t hints;            // RJ_HINT_* cache-hint switches of pass 1
	int world, self;           // owner ranks; index of this GPU in dst[]
	uint32_t pool_chunks;      // capacity of every target's pool
	uint32_t id_batch, id_low; // chunk ids a CTA reserves per owner at a time / refill threshold
	RJTarget dst[RJ_MAX_RANKS];
	uint16_t *pool;            // pass 2 reads remainders from here (dst[self].pool unless an NCCL exchange staged them)
	RJDesc *dir;               // (offset, entries) of this GPU's chunks grouped by partition
	uint64_t *dir_off;         // exclusive offsets into dir
	uint32_t *dir_fill;
};

struct RJParams {
	long long kmin;
	unsigned long long range;  // keys in [kmin, kmin + range) can match
	int shift;                 // remainder bits
	uint32_t mask;             // (1 << shift) - 1
	int nparts;
	int part_first, part_end;  // pass 2 handles partitions [part_first, part_end) (all of them on one GPU)
	uint32_t *error_flag;
};

