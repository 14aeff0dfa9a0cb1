// vg_internal.h -- host-visible launch interface between vg_capi.cpp and vg_kernels.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vg {

constexpr int kMaxWorld = 16;     // ranks of a vg_comm (one GPU each)
struct PeerPtrs {                 // the same symmetric-arena object on every rank, indexed by rank
    void* p[kMaxWorld];
};

struct IndexView {
    uint64_t* slots;                // nbuckets * 4 slots of [canonical k-mer:56 | count:8]; empty = ~0
    uint32_t nbuckets;              // buckets of THIS table
    uint32_t k;
    uint64_t mask;                  // 2^(2k) - 1
    // A sharded index is one table of nb_total buckets cut into `world` equal runs, one per GPU: this
    // table holds global buckets [b_base, b_base + nbuckets).  Unsharded: nb_total == nbuckets, b_base == 0.
    uint32_t nb_total;
    uint32_t b_base;
    // Partitioned probing keeps the read-coverage counters OUTSIDE the table, densely, in slot order: entry
    // rank_base[b] + s of cvec belongs to slot s of bucket b (buckets fill front to back, so the occupied slots of
    // a bucket are 0 .. occ-1 and rank_base is the exclusive prefix sum of the occupancies).  The table is then
    // read-only while counting, a sample's result is cvec itself (no gather over the keys, no sweep to zero the
    // counts: one memset), and the sweep writes nothing back into the slices it pulled into L2.
    // nullptr (direct probing of a tiny table): the count lives in the low byte of each slot.
    uint32_t* rank_base;
    uint8_t* cvec;
};

// Presence pre-filter: a word-blocked Bloom filter over the index keys (both bits of a key sit in
// one 32-bit word, so a query is one 4-byte load).  Small enough to stay L2-resident (an L2
// persisting access-policy window pins it), it lets the scatter drop most k-mers that are not in
// the index before they cost a key write and a probe.  No false negatives, so results are unchanged
// (the reference has no such filter on the read path: SURVEY F1).
struct PrefilterView {
    const uint32_t* words;          // nullptr: disabled
    uint32_t nwords;
    uint32_t span;                  // read positions one lookup speaks for: 4 (L2-resident filter) or 8 (filter in DRAM)
};

struct CountStats {                 // lives in device memory
    unsigned long long positions;   // emitted k-mer positions (what kmer_sketch_fastq tests against the map)
    unsigned long long hits;        // positions whose k-mer is in the index
    unsigned long long keys;        // partitioned path: k-mers that passed the pre-filter and went through the key lists
};

struct InsertReport {               // lives in device memory
    unsigned long long duplicates;
    unsigned long long failed;
};

// Partitioned probing ("accumulate, then probe one L2-sized slice of the table at a time").
// Keys of a round are scattered by table slice (partition = home bucket >> shift); each slice
// is then streamed into L2 once and probed by all of its keys, so index probes and counter
// updates hit L2 instead of paying one random DRAM access each.
struct PartView {
    uint64_t* keybuf;               // this GPU's key lists: P_local slices x world sources x cap canonical k-mers
    unsigned long long* cursor;     // P fill counts of the lists THIS GPU writes (global slice numbering)
    uint32_t P;                     // slices of the whole (global) table = P_local * world
    uint32_t shift;                 // slice = global bucket >> shift
    uint64_t cap;                   // capacity of one key list; keys beyond it are probed directly
    uint32_t* ctr;                  // 2 x (4 << shift2) u32 side counters: hits of the slice being probed
    // Two-level scatter (tables of many slices).  P / shift / cap / keybuf / cursor above then describe COARSE
    // partitions of 2^sub_bits slices each, so that a CTA tile's k-mers always fall into a few dozen fat bins whatever
    // the size of the table; the sweep takes one coarse list at a time and re-scatters it (rescatter_kernel) into the
    // lists of its slices -- cap2 keys each at keybuf2, fill counts at cursor2 -- right before probing them.
    // sub_bits == 0: one level, a partition is a slice (shift2 == shift).
    uint32_t sub_bits;
    uint32_t shift2;                // slice = bucket >> shift2; shift == shift2 + sub_bits
    uint64_t cap2;
    uint64_t* keybuf2;
    unsigned long long* cursor2;
    // Sharded index (world > 1): slice p belongs to GPU p / P_local, and the scatter stores its keys
    // straight into that GPU's key list for source `rank` over NVLink (peer memory mapped with CUDA
    // IPC) -- the all-to-all of k-mers is the scatter's own copy-out.  Unsharded: world == 1,
    // P_local == P, peer_keybuf[0] == keybuf, peer_slots[0] == the table.
    uint32_t world, rank, P_local;
    unsigned long long* incount;    // world x P_local: how many keys each source put into each of MY lists
    uint64_t* peer_keybuf[kMaxWorld];
    uint64_t* peer_slots[kMaxWorld];           // for the rare direct probe of a key whose list is full
    uint32_t* peer_rank_base[kMaxWorld];       //   ... and that table's rank_base / cvec (see IndexView)
    uint8_t* peer_cvec[kMaxWorld];
    unsigned long long* peer_incount[kMaxWorld];
};
constexpr uint32_t kMaxPartitions = 1024;
constexpr uint32_t kMaxSubBits = 6;

// On-device FASTQ parsing: per-file and per-block state (device memory) and the scratch of one context.
struct FastqFileState {
    unsigned int bad;               // a block failed the four-line check: it and every later block are not counted
    unsigned int first_bad_block;
    unsigned int blocks_ok;
    unsigned int pad;
    unsigned long long read_bases;  // sum of seq.l over the counted blocks (mReadBase)
};
struct FastqBlockState {
    unsigned int nlines, bad;
    unsigned long long read_bases;
};
struct FastqScratch {
    uint32_t* tile_count;           // newlines per 4 KiB tile
    uint32_t* tile_base;            // lines that start before each tile
    uint32_t* nlpos;                // block offset of the newline that ends each line
    FastqBlockState* blk;
    uint32_t max_tiles, max_lines;
};
// raw block (record-aligned four-line FASTQ text, 16-byte aligned, len < 2^32) -> d_masked (only the sequence
// lines survive, everything else is '\n'), the format check and the per-file bookkeeping; all on stream s.
cudaError_t launch_fastq_block(const uint8_t* d_raw, uint32_t len, uint8_t* d_masked, const FastqScratch& sc,
                               FastqFileState* d_file, uint32_t block_no, cudaStream_t s);

// the bookkeeping of a chunk the host stripped: file->read_bases += bases (and blocks_ok) unless the file is already bad
cudaError_t launch_fastq_strip_commit(FastqFileState* d_file, unsigned long long bases, bool whole_block, cudaStream_t s);

struct CbfView {
    uint8_t* cells;                 // m saturating u8 counters
    uint64_t m;
    uint64_t magic_hi, magic_lo;    // fastmod constant for % m
    uint32_t num_hashes;            // <= 16
    uint32_t seeds[16];             // already truncated to 32 bit, as the reference does at the call
};

int sm_count(int device);

cudaError_t launch_table_fill_empty(uint64_t* slots, uint64_t nslots, cudaStream_t s);
// in place: hash (the reference's key >> 8) -> canonical k-mer (hash64 is invertible)
cudaError_t launch_unhash(uint64_t* d_key56, uint64_t n, uint64_t mask, cudaStream_t s);
// device-resident caller keys -> canonical k-mers (validated; *d_bad must start at ~0)
cudaError_t launch_keys_to_key56(const uint64_t* d_keys, uint64_t n, uint32_t k, uint64_t mask, uint64_t* d_key56,
                                 unsigned long long* d_bad, cudaStream_t s);
cudaError_t launch_insert(const IndexView& ix, const uint64_t* d_key56, uint64_t n, InsertReport* d_rep,
                          cudaStream_t s);
// Sharded build: of the n canonical k-mers at d_key56 (caller positions first_idx + i), append those whose
// home bucket lies in this table to d_own (+ their caller positions to d_own_idx) at *d_n_own.
// d_own == nullptr only counts them.
cudaError_t launch_select_owned(const IndexView& ix, const uint64_t* d_key56, uint64_t n, uint64_t first_idx,
                                uint64_t* d_own, uint64_t* d_own_idx, unsigned long long* d_n_own, cudaStream_t s);
cudaError_t launch_clear_counts(const IndexView& ix, cudaStream_t s);
// rank_base[b] = occupied slots in buckets [0, b); *d_total = occupied slots of the whole table.
// d_block_sums: scratch of (nbuckets + 1023) / 1024 + 1 u32.
cudaError_t launch_rank_scan(const IndexView& ix, uint32_t* d_block_sums, unsigned long long* d_total, cudaStream_t s);
// perm[i] = slot-order position of key i (0xffffffff: the key is not in this table)
cudaError_t launch_slot_perm(const IndexView& ix, const uint64_t* d_key56, uint64_t n, uint32_t* d_perm, cudaStream_t s);
// out[d_idx ? d_idx[i] : i] = cvec[perm[i]] (0 where perm[i] == 0xffffffff): counts in the caller's key order
cudaError_t launch_gather_counts(const uint8_t* cvec, const uint32_t* d_perm, const uint64_t* d_idx, uint64_t n, void* d_out,
                                 int out_elem_bytes, cudaStream_t s);
// out[perm[i]] = in[i]: per-key bytes (the histogram's subset flags) into slot order
cudaError_t launch_scatter_bytes(const uint8_t* d_in, const uint32_t* d_perm, uint64_t n, uint8_t* d_out, cudaStream_t s);
// d_skip (optional): device flag, non-zero = count nothing (see Chunk::skip)
cudaError_t launch_count(const IndexView& ix, const uint8_t* d_bases, uint64_t nbytes, CountStats* d_stats,
                         int ctas_per_sm, int nsm, cudaStream_t s, const unsigned int* d_skip = nullptr);
// Scatter the k-mers ending in tiles [first_tile, first_tile + ntiles) of the chunk (d_bases, nbytes)
// into the partition buffers; launch_probe_partitions then probes every partition and resets them.
cudaError_t launch_scatter(const IndexView& ix, const PartView& pv, const PrefilterView& pf, const uint8_t* d_bases,
                           uint64_t nbytes, int64_t first_tile, int64_t ntiles, CountStats* d_stats, int nsm,
                           cudaStream_t s, const unsigned int* d_skip = nullptr);
cudaError_t launch_prefilter_build(uint32_t* words, uint32_t nwords, const uint64_t* d_key56, uint64_t n, uint32_t k, uint32_t span,
                                   cudaStream_t s);
// h_slice_rank (host, slices + 1 entries, may be NULL): cvec position of the first slot of each slice of this table
cudaError_t launch_probe_partitions(const IndexView& ix, const PartView& pv, const uint32_t* h_slice_rank, CountStats* d_stats,
                                    int nsm, cudaStream_t s);
// *d_total = sum of the P fill counts at `cursor` (each clipped to cap): the keys of the round about to be swept
cudaError_t launch_sum_cursors(const unsigned long long* cursor, uint32_t P, uint64_t cap, unsigned long long* d_total, CountStats* d_stats,
                               cudaStream_t s);
uint64_t sweep_launches(const IndexView& ix, const PartView& pv);
int64_t chunk_tiles(const uint8_t* d_bases, uint64_t nbytes);
// d_idx == nullptr: out[i] = count of d_key56[i]; else out[d_idx[i]] = ... (a sharded index's own keys)
cudaError_t launch_extract(const IndexView& ix, const uint64_t* d_key56, const uint64_t* d_idx, uint64_t n, void* d_out,
                           int out_elem_bytes, cudaStream_t s);
// ---- peer-memory collectives (vg_comm) ----
// Device-side barrier: every rank stores `epoch` into slot `rank` of every peer's flag array, then waits
// until all `world` slots of its own array have reached it.  Gives up after timeout_ns (a dead peer must
// not hang the GPU) and raises *d_timeout.
cudaError_t launch_peer_barrier(const PeerPtrs& peer_flags, int world, int rank, unsigned long long epoch,
                                unsigned long long timeout_ns, unsigned int* d_timeout, cudaStream_t s);
// Tell every owner how many keys this rank put into each of its lists (cursor -> peer incount), re-arm the cursors.
cudaError_t launch_publish_counts(const PartView& pv, cudaStream_t s);
// out[i] = min(255, sum over ranks of peer_counts[r][i]) for i in [0, n): the count reduce over NVLink.
cudaError_t launch_combine_counts(const PeerPtrs& peer_counts, int world, uint64_t n, uint8_t* d_out, int nsm,
                                  cudaStream_t s);
// Replica group (same slot order on every rank): vecs.p[r] = rank r's count vector, nbytes long (padded to 16), cut into
// `world` segments of seg_bytes (a multiple of 16).  reduce: this rank's segment = min(255, sum over ranks), in place;
// gather: every other segment from its owner.  A barrier belongs before, between and after.
cudaError_t launch_reduce_segment(const PeerPtrs& vecs, int world, int rank, uint64_t seg_bytes, uint64_t nbytes, int nsm, cudaStream_t s);
cudaError_t launch_gather_segments(const PeerPtrs& vecs, int world, int rank, uint64_t seg_bytes, uint64_t nbytes, int nsm, cudaStream_t s);
cudaError_t launch_histogram(const uint8_t* d_counts, const uint8_t* d_flags, uint64_t n, unsigned long long* d_hist,
                             cudaStream_t s);
cudaError_t launch_positions(uint32_t k, const uint8_t* d_bases, uint64_t nbytes, uint64_t* d_out,
                             cudaStream_t s);

// d_bases: 16-byte aligned buffer holding one sequence; bytes [0, hi) are resident.  Adds every
// k-mer that ENDS in [own_from, hi); own_from must be a multiple of the 4 KiB CTA tile.
cudaError_t launch_cbf_add(const CbfView& cbf, uint32_t k, const uint8_t* d_bases, uint64_t hi, uint64_t own_from,
                           unsigned long long* d_added, int nsm, cudaStream_t s);
cudaError_t launch_cbf_query(const CbfView& cbf, const uint64_t* d_keys, uint64_t n, uint8_t* d_count,
                             uint8_t* d_find, cudaStream_t s);

cudaError_t launch_random_sectors(const uint64_t* table, uint32_t nbuckets, uint32_t rounds, int grid,
                                  unsigned long long* sink, cudaStream_t s);
int probe_batch();

}  // namespace vg
