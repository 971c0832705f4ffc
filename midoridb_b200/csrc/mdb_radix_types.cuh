// mdb_radix_types.cuh - constants and kernel-parameter structs of the radix join (mdb_radix.cu and its pass headers)
#pragma once
#include <stdint.h>

#define RJ_MAX_PART 4096           // partitions (12 radix bits)
#define RJ_MAX_SHIFT 16            // remainder bits: 2-byte remainders
#define RJ_CAP 20                  // staging slots per partition in shared memory
#define RJ_FLUSH 16                // a partition is flushed when 16 remainders (= one 32-byte sector) are staged
#ifndef RJ_SPLIT
#define RJ_SPLIT 1                 // pass-1 CTAs per SM: CTA b stages only the partitions p with p % RJ_SPLIT == b % RJ_SPLIT
#endif                             // (measured with 2: every tile is decoded twice, 1.26 ms against 0.85 ms per launch)
#define RJ_P1_THREADS (1024 / RJ_SPLIT)
#define RJ_ROWS (RJ_MAX_PART / RJ_SPLIT) // staging rows per CTA
#define RJ_NONE 0xffffffffu

// The key range is cut into nparts = ceil(range / width) <= 4096 partitions of `width` = ceil(range / 4096) key values
// each - NOT into power-of-two blocks: a range a little above a power of two would otherwise use only half of the 4096
// staging rows of pass 1, every row would receive twice the keys per round and overflow ten times as often
// (profiles/r02: 2049 partitions, flags "stream full" + "clustered keys" on plain uniform keys).
#define RJ_MAX_RANKS 8
#define RJ_CUR_STRIDE 2            // 32-bit words per partition in the cursor array: [main cursor, tail cursor].  (One 128-byte
                                   // line per partition was measured 5% SLOWER: the atomics like their few hot L2 lines.)

#define RJ_ERR_STREAM 1u           // a partition's stream is full (its keys are far more frequent than the average)
#define RJ_ERR_COUNTER 2u          // a packed counter wrapped (too many equal keys for the counter width)
#define RJ_ERR_SKEW 4u             // more than RJ_SPILL_CAP keys of one round found their staging row full (clustered keys)

// One side of the join on this GPU.  Pass 1 appends the 2-byte remainders of partition p to p's stream:
//   main  stream + p * cap,       cursor[p] entries      whole 32-byte sectors, appended by every CTA (position =
//                                                         one global atomic per sector: consecutive sectors of a
//                                                         128-byte line are written within about a microsecond by
//                                                         different CTAs and merge in L2 on their way to DRAM)
//   tail  tail + p * tail_cap,    tail_cursor[p] entries  32-byte TAIL SECTORS [up to 15 remainders, count in entry 15]: one per
//                                                         key that found its staging row full (count 1), and one per staging
//                                                         row that is left partially filled at the end of pass 1
struct RJSide {
	const int64_t *keys;
	const uint32_t *present;
	uint64_t n;
	int all_in_range;          // every key of the column lies in [kmin, kmin + range): no per-key range test
	uint32_t hints;            // RJ_HINT_* cache-hint switches of pass 1
	uint16_t *stream;
	uint32_t *cursor;          // cursor[p * RJ_CUR_STRIDE]
	uint16_t *tail;
	uint32_t *tail_cursor;     // tail_cursor[p * RJ_CUR_STRIDE] (same line as the cursor of p)
	uint32_t cap, tail_cap;    // entries per partition (multiples of 16)
};

// What pass 2 reads for one side: for every source rank its streams of the partitions this GPU owns.  Single-GPU
// plans have one source (the local streams); in multi-GPU plans source o is rank o's arena, read over NVLink (peer-mapped
// memory, same layout on every rank), source `self` being this GPU's own.
struct RJRuns {
	int nsrc;                  // 0: the side is a sorted column, see sorted_keys
	int self;                  // multi-GPU plans: index of this rank among the sources
	unsigned long long *pulled_bytes; // statistics: bytes of remainders read from OTHER ranks (nullptr: not counted)
	const int64_t *sorted_keys; // sorted column: partition p = rows [sorted_bnd[p], sorted_bnd[p + 1]) (mdb_radix_sorted.cuh)
	const uint64_t *sorted_bnd;
	uint32_t cap, tail_cap;
	const uint16_t *stream[RJ_MAX_RANKS];
	const uint16_t *tail[RJ_MAX_RANKS];
	const uint32_t *cursor[RJ_MAX_RANKS];
	const uint32_t *tail_cursor[RJ_MAX_RANKS];
	uint32_t first[RJ_MAX_RANKS];
	uint32_t cur_stride[RJ_MAX_RANKS]; // words between two partitions' cursors (RJ_CUR_STRIDE for the local streams, 1 in arena slots)
};

struct RJParams {
	long long kmin;
	unsigned long long range;  // keys in [kmin, kmin + range) can match
	uint32_t width;            // key values per partition (2 .. 65536): partition = (key - kmin) / width, remainder = the rest
	uint32_t magic;            // floor(2^32 / width): __umulhi(d, magic) is d / width or one less (rj_pack corrects it)
	int nparts;
	int part_first, part_end;  // pass 2 handles partitions [part_first, part_end) (all of them on one GPU)
	uint32_t *error_flag;
	const uint32_t *peer_flags; // multi-GPU: every rank's error flags after the exchange (pass 2 does nothing if any is set)
	int n_peer_flags;
	int plain_emit;            // pass 2: always take the compiler-generated emit loop (MDBCU_P2_PLAIN_EMIT=1, A/B measurements)
};
