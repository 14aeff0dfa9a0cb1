/* vgb200.h -- C ABI of libvgb200.so: varigraph's read k-mer counting path on B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference has no plugin ABI; its seam is the class pair
 * FastqKmer / FastqKmerKernel (and BloomFilter / BloomFilterKernel on the construct side).
 * Each entry point below names the reference interface it stands in for (file:line relative
 * to the reference tree).  varigraph_b200/host/ holds the C++ subclasses that keep the
 * reference's signatures and call these functions; INTEGRATION.md shows the wiring.
 *
 * Conventions: plain C types only; every function returns VG_OK (0) or a negative VG_E_*
 * code and leaves a message for vg_last_error() (thread-local); the caller owns all host
 * buffers; a handle serves one sample at a time, distinct handles may be used concurrently
 * from different threads.  There is no CPU fallback: without a CUDA device every call that
 * needs one fails with VG_E_CUDA.
 */
#ifndef VGB200_H
#define VGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_OK 0
#define VG_E_INVALID (-1) /* bad argument (k out of 5..28, key whose low byte != k, NULL, ...) */
#define VG_E_CUDA (-2)    /* CUDA runtime error; text in vg_last_error() */
#define VG_E_NOMEM (-3)
#define VG_E_IO (-4)      /* FASTQ file missing / unreadable */
#define VG_E_STATE (-5)   /* call out of order (submit before begin, ...) */

typedef struct vg_ctx vg_ctx;     /* one CUDA device: streams, staging rings */
typedef struct vg_index vg_index; /* device-resident graph k-mer index + read-coverage counters */
typedef struct vg_cbf vg_cbf;     /* device-resident counting Bloom filter */
typedef struct vg_comm vg_comm;   /* one rank of a group of GPUs that map each other's memory (NVLink) */

const char* vg_last_error(void);
int vg_version(void);

/* ---- context ---------------------------------------------------------------------------
 * Replaces cudaSetDevice(config.gpu) in main.cu:221,444.  buffer_mb is the reference's
 * --buffer (include/varigraph.cuh:19-28): the size of one staged chunk of read bases. */
int vg_ctx_create(int device, int buffer_mb, vg_ctx** out);
int vg_ctx_destroy(vg_ctx* ctx);
int vg_ctx_device(const vg_ctx* ctx);
int vg_ctx_synchronize(vg_ctx* ctx);
/* Run the context's kernels on the caller's stream (a cudaStream_t; NULL restores the context's
 * own) -- the cublasSetStream convention, so a host that owns its streams can order and time them. */
int vg_ctx_set_stream(vg_ctx* ctx, void* cuda_stream);

/* ---- the index -------------------------------------------------------------------------
 * Device twin of ConstructIndex::mGraphKmerHashHapStrMap (include/construct_index.hpp:140):
 * keys are the map's keys verbatim, `hash64(canonical k-mer) << 8 | k` (src/kmer.cpp:138).
 * Built once per load() (src/varigraph.cpp:65-83).  load_factor in (0, 0.9], 0 = default. */
int vg_index_create(vg_ctx* ctx, const uint64_t* keys, uint64_t n, uint32_t k, double load_factor,
                    vg_index** out);
/* The same from keys that already sit in the memory of ctx's GPU (a graph loaded or generated there): no
 * host copy of the key array is ever made -- 16 GB per process at human scale (2e9 k-mers). */
int vg_index_create_device(vg_ctx* ctx, const uint64_t* dev_keys, uint64_t n, uint32_t k, double load_factor,
                           vg_index** out);
int vg_index_destroy(vg_index* ix);
uint64_t vg_index_size(const vg_index* ix);          /* n */
uint64_t vg_index_table_bytes(const vg_index* ix);   /* bytes of the slot table in HBM */
uint32_t vg_index_partitions(const vg_index* ix);    /* partitions the scatter bins by (table slices, or coarse groups of them), 0 = direct */
uint32_t vg_index_slices(const vg_index* ix);        /* L2-sized table slices of this GPU's table the sweep probes, 0 = direct */
uint64_t vg_index_launches(const vg_index* ix);      /* kernels launched for this index so far */
uint64_t vg_index_duplicates(const vg_index* ix);    /* keys given more than once at create (they share a slot) */
uint64_t vg_count_h2d_bytes(const vg_index* ix);     /* bytes copied host -> device since the last vg_count_begin */
/* Diagnostic: with timing on, every scatter launch and every sweep over the table slices is bracketed by CUDA
 * events on the context stream and waited for (so it serialises the host: not for production runs);
 * vg_index_timing returns the device milliseconds and launch counts accumulated since vg_index_set_timing. */
int vg_index_set_timing(vg_index* ix, int on);
int vg_index_timing(const vg_index* ix, double* scatter_ms, double* sweep_ms, uint64_t* scatter_launches,
                    uint64_t* sweeps);

/* ---- one sample's count phase ----------------------------------------------------------
 * Together these replace FastqKmer::build_fastq_index (src/fastq_kmer.cpp:41-187) /
 * FastqKmerKernel::build_fastq_index_kernel (src/fastq_kmer.cu:20-270).
 * Post-condition (src/fastq_kmer.cpp:132-138): c[i] == min(255, #emitted read k-mers equal
 * to keys[i]).  vg_count_begin also stands in for ConstructIndex::reset()'s zeroing of c
 * (include/construct_index.hpp:326-329). */
int vg_count_begin(vg_index* ix);

/* A staged chunk: read sequences (ASCII, any case, N allowed) each followed by '\n' -- the
 * reference GPU path's 'N'-separated buffer (src/fastq_kmer.cu:171-176) with a separator that
 * also resets the encoder registers as a new kmer_sketch_fastq call does.  Reads must not be
 * split across calls.  host_bases may be pageable or pinned; the copy is pipelined against
 * the kernel through the context's staging ring.  Returns once the bytes are consumed from
 * host_bases (not necessarily counted yet). */
int vg_count_submit(vg_index* ix, const char* host_bases, uint64_t nbytes);

/* Same, for bases already in device memory; enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = the context's compute stream).  Asynchronous. */
int vg_count_submit_device(vg_index* ix, const void* dev_bases, uint64_t nbytes, void* cuda_stream);

/* FASTQ/FASTA files (plain or gzip), kseq semantics (include/kseq.h:192-232): the whole of
 * FastqKmer::fastq_file_open for each path.  threads = inflate/parse workers (files are
 * processed concurrently).  *read_bases accumulates mReadBase (src/fastq_kmer.cpp:105). */
int vg_count_files(vg_index* ix, const char* const* paths, int npaths, int threads, uint64_t* read_bases);
/* The same for ONE sample counted on several GPUs of this process: ixs[g] = a replica of the index on GPU g
 * (vg_index_replicate over a vg_comm_create_local group), each begun with vg_count_begin.  One set of workers reads
 * the files; the staged chunks are dealt to the GPUs as they come, so the reads are sharded over the GPUs whatever the
 * number of files.  Afterwards every rank calls vg_count_allreduce_slots (concurrently). */
int vg_count_files_multi(vg_index* const* ixs, int nix, const char* const* paths, int npaths, int threads,
                         uint64_t* read_bases);
/* Plain (not gzip) four-line FASTQ is cut into record-aligned blocks that many workers handle at once: by default they
 * strip it to its sequences on the host (vector scan of the memory-mapped file, every record checked against kseq's
 * rules; VG_FASTQ_ROAD=device ships the raw text and does the same on the GPU).  Anything else -- and a file from its
 * first irregular record on -- goes through the host kseq reader, so results never depend on the road taken.  This
 * counts the blocks accepted without a flaw so far (diagnostic; VG_FASTQ_ROAD=kseq disables the block roads). */
uint64_t vg_index_fastq_blocks(const vg_index* ix);
/* Host-only diagnostic (no GPU needed): the byte offset at which the raw road would cut `path` at or after
 * `at` -- the first line start whose line begins with '@' and whose next-but-one line begins with '+' within
 * `window` bytes; -1 if there is none, -2 if the file cannot be read. */
int64_t vg_fastq_record_boundary(const char* path, uint64_t at, uint64_t window);

/* Host-only diagnostic (no GPU needed): what the host workers make of one block of plain four-line FASTQ text that
 * starts at a record boundary -- "sequence\n" records at out (room for nbytes / 2 + 256 bytes), the number of bytes
 * written as the result, *bases = their seq.l sum, *bad_at = offset of the first record that is not what kseq reads as a
 * four-line record (-1: none; nothing from there on is written; the caller re-reads from there with the kseq reader).
 * last != 0: the block ends at the end of the file, where the final newline may be missing. */
int64_t vg_fastq_strip_block(const char* text, uint64_t nbytes, int last, uint8_t* out, uint64_t* bases, int64_t* bad_at);

/* Host-only diagnostic (no GPU needed): the parallel inflater behind vg_count_files' gzip road (replaces gzread of
 * /root/reference/include/kseq.h:242 + src/fastq_kmer.cpp:74 for .gz input).  Inflates the gzip file at path -- one or
 * many members -- with `threads` workers, `chunk_bytes` compressed bytes per worker and round; *out (release it with
 * vg_gunzip_free) holds the *out_len bytes zlib would return.  0 ok, -1 cannot read the file, -2 not gzip / corrupt. */
int vg_gunzip_parallel(const char* path, int threads, uint64_t chunk_bytes, uint8_t** out, uint64_t* out_len);
void vg_gunzip_free(uint8_t* p);
/* Host-only diagnostic: the CRC-32 the inflater checks gzip members with (carry-less multiply where the CPU has it);
 * equals zlib's crc32(crc, buf, len). */
uint32_t vg_crc32(uint32_t crc, const uint8_t* buf, uint64_t len);

/* Enqueues whatever counting work is still deferred (the partitioned path accumulates k-mers of a
 * round before probing); asynchronous.  vg_count_end / _stats / _extract_device imply it. */
int vg_count_flush(vg_index* ix);

/* Waits for the sample's kernels, then writes c in the key order given at create.
 * c_out (n bytes, host), positions, hits may each be NULL. */
int vg_count_end(vg_index* ix, uint8_t* c_out, uint64_t* positions, uint64_t* hits);

/* The same result without the gather into key order.  An index of 8 MB or more keeps its counters in ONE dense
 * u8 vector, one entry per distinct k-mer, in the order the k-mers sit in the device table ("slot order"); the
 * counting kernels accumulate straight into it, so a sample's result exists the moment its last kernel ends.
 *   vg_index_slots      m, the length of that vector (<= n; == n when the keys are distinct)
 *   vg_index_slot_perm  perm[i] = position of keys[i] in it (n entries, fixed for the life of the index;
 *                       0xffffffff for a key no read can produce, whose count is always 0).  The host side
 *                       permutes its pointers into the map once per graph (what DeviceGraphIndex does with
 *                       mGraphKmerHashHapStrMap's entries) and afterwards writes c[perm[i]] back per sample.
 *   vg_count_end_slots  vg_count_end, with c in slot order (m bytes)
 *   vg_count_slots_device  implies vg_count_flush; *dev_counts = the vector on the device (m bytes, valid until
 *                       the next vg_count_begin, ordered after the sample's kernels on the context stream)
 * (a tiny index that is probed directly reports slot order == key order).  Not for sharded indexes. */
uint64_t vg_index_slots(const vg_index* ix);
int vg_index_slot_perm(vg_index* ix, uint32_t* perm_out);
int vg_count_end_slots(vg_index* ix, uint8_t* c_slots_out, uint64_t* positions, uint64_t* hits);
int vg_count_slots_device(vg_index* ix, const uint8_t** dev_counts);

/* Device-side result for a multi-GPU reduce: counts in key order as u8 (elem_bytes 1) or u32
 * (elem_bytes 4, what ncclAllReduce(sum) wants) into dev_out, on `cuda_stream`. */
int vg_count_extract_device(vg_index* ix, void* dev_out, int elem_bytes, void* cuda_stream);
int vg_count_stats(vg_index* ix, uint64_t* positions, uint64_t* hits);
/* Diagnostic: k-mers of the sample that passed the presence pre-filter and went through the key lists of the partitioned
 * path, as of the last vg_count_stats / vg_count_end (0 for a directly probed index). */
uint64_t vg_count_keys(const vg_index* ix);

/* Count consumers on the device: hist[v] = number of index entries whose count is v,
 * over all entries or over the subset marked by vg_index_set_flags (flags: n bytes in key order,
 * non-zero = counted; NULL clears the subset).  With flags = "f <= 1 and homozygous in some sample"
 * this is the histogram Varigraph::get_hom_kmer builds by walking the host map
 * (src/varigraph.cpp:253-296).  Call after the sample's reads are submitted; implies a flush. */
int vg_index_set_flags(vg_index* ix, const uint8_t* flags);
int vg_count_histogram(vg_index* ix, uint64_t* hist256);

/* Test / tooling hook: out[p] = key of the k-mer ENDING at byte p of the chunk, or ~0 when
 * the reference encoder emits nothing there (src/kmer.cpp:126-146).  Device buffers. */
int vg_encode_positions_device(vg_ctx* ctx, const void* dev_bases, uint64_t nbytes, uint32_t k,
                               uint64_t* dev_keys_out, void* cuda_stream);
int vg_encode_positions(vg_ctx* ctx, const char* host_bases, uint64_t nbytes, uint32_t k,
                        uint64_t* host_keys_out);

/* ---- counting Bloom filter (construct side) --------------------------------------------
 * Device twin of BloomFilter / BloomFilterKernel (include/counting_bloom_filter.hpp:29-96,
 * include/counting_bloom_filter.cuh:34-86).  m and num_hashes are the host class's _size and
 * _numHashes (src/counting_bloom_filter.cpp:70-77); seeds are its _seeds (:80-87) -- the host
 * draws them, the device receives them, so both sides build the same filter. */
int vg_cbf_create(vg_ctx* ctx, uint64_t m, uint32_t num_hashes, const uint64_t* seeds, vg_cbf** out);
int vg_cbf_destroy(vg_cbf* cbf);
/* kmer_sketch_bf over one chromosome (src/kmer.cpp:20-52, caller src/construct_index.cpp:161-166;
 * GPU caller src/construct_index.cu:65-88). *added gets the number of k-mers added (may be NULL). */
int vg_cbf_add_sequence(vg_cbf* cbf, const char* host_seq, uint64_t len, uint32_t k, uint64_t* added);
/* BloomFilterKernel::copyFilterDToHost (include/counting_bloom_filter.cuh:69-82): m bytes. */
int vg_cbf_download(vg_cbf* cbf, uint8_t* host_filter);
/* BloomFilter::count / find for a batch (src/counting_bloom_filter.cpp:40-67); either out may be NULL. */
int vg_cbf_query(vg_cbf* cbf, const uint64_t* host_keys, uint64_t n, uint8_t* count_out, uint8_t* find_out);

/* ---- several GPUs (one process per GPU) ---------------------------------------------------
 * The reference counts on one device (main.cu:221, src/fastq_kmer.cu); BASELINE.json's north star
 * shards the reads over the GPUs of a box.  Peers reach each other's memory directly over NVLink
 * (CUDA IPC mappings): no host and no NCCL in the data path.  The caller only carries the 64-byte
 * handles between the processes (MPI, torch.distributed, a file ...):
 *     vg_comm_create -> vg_comm_handle -> [all-gather the handles] -> vg_comm_connect.
 * arena_bytes: device memory this rank exposes to its peers (count vectors; for a sharded index also
 * its part of the table and the key lists it receives).  Calls marked COLLECTIVE must be made by every
 * rank, in the same order; a rank whose peer does not arrive within VG_BARRIER_TIMEOUT_MS (20 s)
 * gives up and the next host-visible result call returns VG_E_STATE. */
#define VG_COMM_HANDLE_BYTES 64
int vg_comm_create(vg_ctx* ctx, int rank, int world, uint64_t arena_bytes, vg_comm** out);
int vg_comm_handle(const vg_comm* comm, void* handle_out /* VG_COMM_HANDLE_BYTES */);
int vg_comm_connect(vg_comm* comm, const void* handles /* world x VG_COMM_HANDLE_BYTES, rank order */);
/* The same group with all ranks in ONE process (a host program that drives several GPUs itself, as the drop-in binary
 * does for --gpu 0,1,...): out[r] becomes rank r on ctxs[r]'s GPU, the arenas are reached by ordinary peer access, no
 * handles travel.  COLLECTIVE calls on such a group are made by every rank concurrently, one host thread per rank. */
int vg_comm_create_local(vg_ctx* const* ctxs, int world, uint64_t arena_bytes, vg_comm** out /* world entries */);
int vg_comm_destroy(vg_comm* comm);
int vg_comm_rank(const vg_comm* comm);
int vg_comm_world(const vg_comm* comm);
uint64_t vg_comm_launches(const vg_comm* comm);  /* barrier / combine kernels launched so far */
int vg_comm_barrier(vg_comm* comm); /* COLLECTIVE; device-side, enqueued on the context stream */
int vg_comm_check(vg_comm* comm);   /* waits for the context stream; VG_E_STATE if a barrier timed out */

/* Replicated index, reads sharded over the ranks (the default layout): COLLECTIVE end of a sample.
 * c = min(255, sum over ranks of the per-rank counts) -- exact, counts being saturating sums -- with
 * every rank reading its peers' u8 count vectors over NVLink.  Implies vg_count_flush.  The n counts
 * land in dev_out (device, may be NULL) and / or c_out (host, may be NULL); with c_out the call
 * returns when they are there, otherwise it is asynchronous on the context stream. */
int vg_count_allreduce(vg_comm* comm, vg_index* ix, uint8_t* c_out, void* dev_out);

/* Replica group: the index is built ONCE and copied, bit for bit, to the other ranks over NVLink (COLLECTIVE).  Rank
 * `root` passes an index it built with vg_index_create / _device (8 MB or more, i.e. partitioned) and gets the same handle
 * back in *out; every other rank passes NULL and gets a replica.  No rank but the root ever holds the key array, all ranks
 * share one slot order, and the count vectors move into the group's arena (arena_bytes >= n + 1 MB), so that
 * vg_count_allreduce_slots can combine them where they lie: a reduce-scatter and an all-gather over peer memory,
 * 2 (world-1)/world bytes per k-mer and rank on the wire, result = every rank's vector in slot order
 * (vg_index_slot_perm maps it to the keys; identical on all ranks).  vg_count_allreduce on such an index does the same
 * and then gathers into key order.  c_slots_out (host, vg_index_slots bytes) and dev_counts may each be NULL; with
 * c_slots_out the call returns when the counts are there, otherwise it is asynchronous on the context stream. */
int vg_index_replicate(vg_comm* comm, int root, vg_index* root_ix, vg_index** out);
int vg_count_allreduce_slots(vg_comm* comm, vg_index* ix, uint8_t* c_slots_out, const uint8_t** dev_counts);

/* Sharded index (the index does not fit one GPU): ONE open-addressing table cut into `world` runs of
 * buckets, one per rank.  COLLECTIVE; every rank passes the same keys in the same order and keeps those
 * whose home bucket it owns.  Each rank then submits ITS reads with vg_count_submit / _submit_device /
 * _files as usual: the scatter kernel stores every k-mer straight into the key list of the GPU that owns
 * its table slice (the all-to-all of k-mers is the kernel's own copy-out), and each GPU probes its
 * slices.  Differences from a plain index:
 *   - a round holds at most round_bytes of submitted bases per rank (0 = 256 MB); vg_count_room() says
 *     how much still fits, a submit beyond it fails with VG_E_STATE;
 *   - vg_count_begin, vg_count_flush (end of a round) and vg_count_end (last round + result) are COLLECTIVE;
 *   - vg_count_end gives every rank the counts of all n keys (c_out), its own positions, and the hits
 *     found in its part of the table; sums over ranks equal the single-GPU figures;
 *   - vg_count_extract_device / vg_count_histogram are not available. */
int vg_index_create_sharded(vg_comm* comm, const uint64_t* keys, uint64_t n, uint32_t k, double load_factor,
                            uint64_t round_bytes, vg_index** out);
/* The same from keys resident in the memory of this rank's GPU (a graph generated or loaded there): no rank holds or
 * stages the key array in host memory (16 GB per process at human scale). */
int vg_index_create_sharded_device(vg_comm* comm, const uint64_t* dev_keys, uint64_t n, uint32_t k, double load_factor,
                                   uint64_t round_bytes, vg_index** out);
uint64_t vg_index_own_keys(const vg_index* ix);  /* keys in this rank's part of the table */
uint64_t vg_count_room(const vg_index* ix);      /* bytes of bases the current round still takes; ~0 if unlimited */

/* Diagnostic: measured throughput of uniform random 32-byte sector gathers over a table of
 * table_bytes (>> L2) with the same load shape as the index probe: the denominator of the
 * "fraction of random-access HBM peak" figure.  No reference counterpart. */
int vg_probe_random_sectors(vg_ctx* ctx, uint64_t table_bytes, uint32_t rounds, double* gbytes_per_s,
                            double* sectors_per_s);

/* Pinned host memory for callers that want zero-copy staging (cudaHostAlloc / cudaFreeHost). */
int vg_host_alloc(void** out, uint64_t nbytes);
int vg_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* VGB200_H */
