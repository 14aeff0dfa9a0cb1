// vg_host.h -- host-side objects behind the opaque handles of include/vgb200.h.
#pragma once
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "vg_internal.h"

namespace vg {

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
    bool has_l2_window = false;          // L2 persisting window over the presence pre-filter
    cudaAccessPolicyWindow l2_window{};
};

// Host-side state of the partitioned probing path of one index.
struct PartState {
    bool enabled = false;
    vg::PartView view{};
    uint64_t round_keys = 0;   // keys a round may accumulate before it must be probed
    uint64_t pending = 0;      // upper bound of keys scattered since the last probe pass
    vg::PrefilterView filter{nullptr, 0};
    uint32_t* d_filter = nullptr;
};

struct vg_index {
    vg_ctx* ctx = nullptr;
    PartState part;
    uint64_t n = 0;
    vg::IndexView view{};
    uint64_t* d_key56 = nullptr;   // key order given at create: the canonical k-mer of each key
    uint8_t* d_counts = nullptr;   // scratch for vg_count_end
    uint8_t* d_flags = nullptr;    // optional per-entry subset for vg_count_histogram
    unsigned long long* d_hist = nullptr;
    vg::DeviceMisc* d_misc = nullptr;
    uint64_t duplicates = 0;
    uint64_t launches = 0;
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
int enqueue_piece(vg_index* ix, int slot, const char* src, uint64_t len);
int count_files(vg_index* ix, const char* const* paths, int npaths, int threads, uint64_t* read_bases);
}  // namespace vg
