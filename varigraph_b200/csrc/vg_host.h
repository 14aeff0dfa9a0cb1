// vg_host.h -- host-side objects behind the opaque handles of include/vgb200.h.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "vg_internal.h"

#define VG_CU(expr)                                                                                \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return vg::fail(e__ == cudaErrorMemoryAllocation ? -3 /* VG_E_NOMEM */ : -2 /* VG_E_CUDA */, "%s: %s", #expr, \
                            cudaGetErrorString(e__));                                              \
    } while (0)

namespace vg {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// One pinned-host / device buffer pair of the staging ring.
struct StageSlot {
    uint8_t* d_buf = nullptr;
    uint8_t* h_pin = nullptr;
    cudaEvent_t copied = nullptr;  // H2D of this slot finished
    cudaEvent_t done = nullptr;    // the count kernel that read d_buf finished
    bool busy = false;
};

struct DeviceMisc {  // small device-resident scalars of one index
    CountStats stats;
    InsertReport report;
};

constexpr int kCtaThreadsHost = 256;                // == kCtaThreads in vg_device.cuh
constexpr uint64_t kTilePieceBytes = 64ull << 20;  // multiple of the 4 KiB CTA tile

int fail(int code, const char* fmt, ...);

}  // namespace vg

struct vg_index;
struct vg_comm;

struct vg_ctx {
    int device = 0;
    int nsm = 148;
    int ctas_per_sm = 0;  // 0: ask the occupancy calculator
    size_t chunk_bytes = 64u << 20;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t compute_stream = nullptr;
    cudaStream_t own_compute_stream = nullptr;
    std::vector<vg::StageSlot> ring;
    int next_slot = 0;
    uint8_t* d_masked = nullptr;         // on-device FASTQ parsing: a block with only its sequence lines left
    vg::FastqScratch fq{};
    bool has_l2_window = false;          // L2 persisting window over the presence pre-filter
    cudaAccessPolicyWindow l2_window{};
};

// Host-side state of the partitioned probing path of one index.
struct PartState {
    bool enabled = false;
    vg::PartView view{};
    uint64_t round_keys = 0;   // keys a round may accumulate before it must be probed
    uint64_t pending = 0;      // upper bound of keys scattered since the last probe pass
    uint64_t slack = 0;        // extra keys per list on top of 1.25 x the even share
    bool may_grow = false;     // rounds double (up to 4 G bases) when a sample needs more than one
    vg::PrefilterView filter{nullptr, 0, 4};
    uint32_t* d_filter = nullptr;
    // How long a round may get.  The lists are sized for round_keys KEYS, but `pending` counts scattered text, and the
    // pre-filter keeps only part of it: every sweep reports the keys its round really held (a device sum of the fill
    // counts, copied back asynchronously), and later rounds take pending up to round_keys / (that share x 1.25).
    // An underestimate is harmless: keys that find their list full are probed at once, exactly.
    double key_share = 1.0;              // keys per byte of scattered text, last observed
    unsigned long long* d_round_keys = nullptr;
    unsigned long long* h_round_keys = nullptr;  // pinned
    cudaEvent_t ev_round = nullptr;
    uint64_t round_pending = 0;          // the `pending` that h_round_keys belongs to (0: nothing in flight)
    std::vector<uint32_t> slice_rank;  // cvec position of the first slot of every table slice (+ the total): what to prefetch
    // diagnostic (vg_index_set_timing): CUDA events around every scatter launch and every sweep, accumulated
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms_scatter = 0, ms_sweep = 0;
    uint64_t n_scatter = 0, n_sweep = 0;
};

// One rank of a group of GPUs whose processes map each other's memory (CUDA IPC, NVLink).  Everything a
// peer must reach lives in ONE device allocation per rank, the arena; all ranks carve it up with the same
// sequence of sizes, so an object sits at the same offset everywhere and a peer's copy is
// peer_base[r] + offset (a symmetric heap).
// The ranks of a vg_comm_create_local group share this: their barrier is events + a rendezvous of the host threads
// (each rank records an event on its stream, everybody meets, each stream then waits for every peer's event), so no
// kernel ever spins on the device -- a spinning kernel and a device-wide synchronisation in a sibling thread of the same
// process (cudaFree, cudaDeviceSetLimit ...) would wait for each other.
struct LocalGroup {
    std::mutex mu;
    std::condition_variable cv;
    int world = 0, arrived = 0;
    unsigned long long generation = 0;
    std::vector<cudaEvent_t> ev;  // world x 2 (alternating by barrier parity)
    std::vector<int> device;
};

struct vg_comm {
    vg_ctx* ctx = nullptr;
    std::shared_ptr<LocalGroup> group;   // in-process group only
    int rank = 0, world = 1;
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0, arena_used = 0;
    uint8_t* peer_base[vg::kMaxWorld] = {};
    bool connected = false;
    bool local = false;                  // all ranks live in this process (vg_comm_create_local): plain peer pointers, no CUDA IPC
    unsigned long long epoch = 0;        // barriers enqueued so far (same on every rank)
    unsigned int* d_timeout = nullptr;   // raised by a barrier that gave up on a peer
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
    size_t reduce_off = 0, reduce_bytes = 0;  // symmetric scratch of vg_count_allreduce
    uint8_t* d_reduced = nullptr;
    size_t reduced_bytes = 0;
    uint64_t launches = 0;
};

struct vg_index {
    vg_ctx* ctx = nullptr;
    PartState part;
    uint64_t n = 0;
    vg_comm* comm = nullptr;       // sharded index: the group it is spread over
    bool sharded = false;
    uint64_t n_own = 0;            // sharded: keys whose home bucket lies in this GPU's table
    uint64_t* d_idx = nullptr;     // sharded: caller position of each own key
    size_t counts_off = 0;         // sharded: arena offset of d_counts (peers read it in the combine)
    uint8_t* d_combined = nullptr; // sharded: counts of all n keys after the combine
    vg::IndexView view{};
    uint64_t* d_key56 = nullptr;   // direct probing only: the canonical k-mer of each key, in the order given at create
    uint8_t* d_counts = nullptr;   // scratch for vg_count_end (counts in key order)
    uint8_t* d_flags = nullptr;    // optional per-entry subset for vg_count_histogram (slot order when partitioned)
    // partitioned probing: the counters live in view.cvec, in slot order (see IndexView)
    uint32_t* d_perm = nullptr;    // slot-order position of each key (sharded: of each own key)
    uint64_t m_slots = 0;          // occupied slots of this table == entries of view.cvec
    vg_comm* replica_of = nullptr; // member of a replica group (vg_index_replicate): same table on every rank, cvec in the arena
    size_t cvec_off = 0;           //   arena offset of view.cvec
    unsigned long long* d_hist = nullptr;
    vg::DeviceMisc* d_misc = nullptr;
    uint64_t duplicates = 0;
    uint64_t launches = 0;
    uint64_t fastq_blocks = 0;     // raw FASTQ blocks parsed, checked and counted on the device so far
    uint64_t last_keys = 0;        // CountStats::keys as of the last vg_count_stats / _end
    uint64_t h2d_bytes = 0;        // bytes copied host -> device for this index since the last vg_count_begin
    bool counting = false;
    bool foreign_streams = false;
};

struct vg_cbf {
    vg_ctx* ctx = nullptr;
    vg::CbfView view{};
    unsigned long long* d_added = nullptr;
    uint8_t* d_seq = nullptr;
    uint64_t d_seq_cap = 0;
};

namespace vg {
struct PartGeometry {
    uint32_t shift2;    // buckets per slice = 2^shift2
    uint32_t sub_bits;  // slices per coarse partition = 2^sub_bits (0: the scatter bins by slice)
    uint32_t P_local;   // partitions of the scatter per GPU
};
bool part_geometry(uint64_t nbuckets, uint32_t world, PartGeometry& g);
uint64_t sweep_launches(const IndexView& ix, const PartView& pv);  // kernels one sweep launches
cudaError_t part_alloc_lists(vg_index* ix);
void part_free_lists(vg_index* ix);
int fetch_slice_ranks(vg_index* ix);  // fills part.slice_rank from rank_base (after the rank scan)
// size and span of the presence pre-filter of an index of n keys (bytes == 0: none)
void prefilter_plan(uint64_t n, uint32_t k, uint64_t& bytes, uint32_t& span);
void pin_in_l2(vg_ctx* c, void* ptr, size_t bytes);   // L2 persisting window over the presence pre-filter
cudaError_t counts_in_key_order(vg_index* ix, void* d_out, int elem_bytes, cudaStream_t s);
int sharded_flush(vg_index* ix, cudaStream_t s);      // vg_comm.cpp: publish, barrier, sweep, barrier
int sharded_end(vg_index* ix, uint8_t* c_out);        // vg_comm.cpp: extract own keys, combine over NVLink
int enqueue_piece(vg_index* ix, int slot, const char* src, uint64_t len, const unsigned int* d_skip = nullptr);
// raw four-line FASTQ text in ring slot `slot`'s pinned buffer: copy, parse and check on the device, count
int enqueue_raw_piece(vg_index* ix, int slot, uint64_t len, FastqFileState* d_file, uint32_t block_no);
int commit_stripped(vg_index* ix, FastqFileState* d_file, uint64_t bases, bool whole_block);
int ctx_ensure_fastq(vg_ctx* c);
int count_files(vg_index* ix, const char* const* paths, int npaths, int threads, uint64_t* read_bases);
}  // namespace vg
