// vg_capi.cpp -- the extern "C" surface declared in include/vgb200.h.
// Host-side plumbing only: device memory, streams, the pinned staging ring.  All arithmetic of
// the path lives in vg_kernels.cu.  No CPU fallback exists: every compute entry point needs a
// CUDA device and fails with VG_E_CUDA otherwise.
#include "../../include/vgb200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vg_host.h"
#include "vg_internal.h"

namespace {
thread_local std::string g_err;
}

namespace vg {
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
}  // namespace vg

using vg::fail;

#define CU VG_CU
using vg::DeviceGuard;

// ---- partitioned probing: set-up, scatter, flush ---------------------------------------------
// Used for every table of 8 MB or more whose slice count stays within what one CTA tile can scatter
// well (<= 1024): direct probing of a table much larger than L2 pays one random DRAM access per
// k-mer, and even an L2-resident table is probed faster through the sweep than with per-hit CAS.
// VG_PARTITION=0 forces direct probing, =1 forces partitioning; VG_SLICE_BYTES / VG_ROUND_KEYS /
// VG_PART_SLACK tune it (the tests use them to drive tiny tables through every branch).
void vg::pin_in_l2(vg_ctx* c, void* ptr, size_t bytes) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
        const size_t want = std::min<size_t>(bytes, (size_t)prop.persistingL2CacheMaxSize);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = ptr;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(c->compute_stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
        c->l2_window = av.accessPolicyWindow;
        c->has_l2_window = true;
    }
}

// Slices and coarse partitions of a table of `nbuckets` buckets per GPU, `world` GPUs (the same on every rank).
// false: the table is to be probed directly (no workable partitioning).
bool vg::part_geometry(uint64_t nbuckets, uint32_t world, PartGeometry& g) {
    const uint64_t table_bytes = 32ull * nbuckets;
    uint64_t slice_bytes = 32ull << 20;
    const char* sb = getenv("VG_SLICE_BYTES");
    if (sb && strtoull(sb, nullptr, 10) >= 64) {
        slice_bytes = strtoull(sb, nullptr, 10);
    } else {
        const uint64_t floor_bytes = world > 1 ? 64 : (1ull << 20);
        while (slice_bytes > floor_bytes && table_bytes / slice_bytes < 3) slice_bytes >>= 1;
    }
    uint32_t shift2 = 0;
    while ((32ull << (shift2 + 1)) <= slice_bytes) ++shift2;  // buckets per slice = 2^shift2
    const uint64_t nslices = (nbuckets + (1ull << shift2) - 1) >> shift2;
    uint64_t two_from = 96;
    if (const char* e = getenv("VG_TWO_LEVEL_FROM")) two_from = strtoull(e, nullptr, 10);
    uint32_t sub_bits = 0;
    if (nslices * world > two_from) {
        const uint64_t coarse_max = std::max<uint64_t>(64 / world, 2);  // per GPU; all GPUs together: a CTA tile's bins
        while (sub_bits < vg::kMaxSubBits && ((nslices + (1ull << sub_bits) - 1) >> sub_bits) > coarse_max) ++sub_bits;
        if (sub_bits == 0) sub_bits = 1;
    }
    uint64_t P_local = (nslices + (1ull << sub_bits) - 1) >> sub_bits;
    if (world > 1 && sub_bits == 0 && P_local < 3) P_local = 3;  // a sharded table is sized by its slices: the sweep needs three
    if (world > 1 && sub_bits && P_local < 2) P_local = 2;
    if (P_local * world > vg::kMaxPartitions) return false;
    if (world == 1 && nslices < 3) return false;  // the sweep retires slice p-1 while slice p is probed
    g.shift2 = shift2;
    g.sub_bits = sub_bits;
    g.P_local = (uint32_t)P_local;
    return true;
}

// Presence pre-filter.  Up to 128 M keys: 4 bits per key (measured best on B200: 30 MB for the chr20 index beats both
// 8 bits per key and none), at most 64 MB, pinned in L2, one lookup per 4 read positions.  Larger indexes: 8 bits per key
// in plain HBM, one lookup per 8 positions -- a random DRAM sector per lookup, but at human variant density it keeps
// 70 % of the k-mers out of the key lists, the re-scatter and the sweep (42 -> 64 G k-mers/s on the 1.2e9-key index).
// VG_PREFILTER=0 disables it, VG_PREFILTER_BYTES / VG_PREFILTER_SPAN (4 | 8) override.  Keyed by sub-words of the k-mers.
void vg::prefilter_plan(uint64_t n, uint32_t k, uint64_t& bytes, uint32_t& span) {
    bytes = 0;
    span = 4;
    const char* pe = getenv("VG_PREFILTER");
    if ((pe && atoi(pe) == 0) || n == 0 || k < 8) return;
    if (n / 2 <= (64ull << 20)) {
        bytes = n / 2;
    } else {
        bytes = std::min<uint64_t>(n, 0x7fffffffull * 4);
        span = 8;
    }
    if (const char* fb = getenv("VG_PREFILTER_BYTES")) bytes = strtoull(fb, nullptr, 10);
    if (const char* fs = getenv("VG_PREFILTER_SPAN")) span = atoi(fs) == 8 ? 8 : 4;
    if (k < 12) span = 4;  // the filter's words are (k - span + 1)-mers
}

int vg::fetch_slice_ranks(vg_index* ix) {
    PartState& ps = ix->part;
    const uint64_t per = 1ull << ps.view.shift2;
    const uint64_t nslices = (ix->view.nbuckets + per - 1) / per;
    ps.slice_rank.assign((size_t)nslices + 1, 0);
    CU(cudaMemcpy2D(ps.slice_rank.data(), sizeof(uint32_t), ix->view.rank_base, per * sizeof(uint32_t), sizeof(uint32_t), (size_t)nslices,
                    cudaMemcpyDeviceToHost));
    ps.slice_rank[(size_t)nslices] = (uint32_t)ix->m_slots;
    return VG_OK;
}

// The buffers of the partitioned path whose sizes follow from ps.view (P, shift2, sub_bits, cap, cap2): key lists, fill
// counts, side counters, and the feedback on how full a round's lists get.
cudaError_t vg::part_alloc_lists(vg_index* ix) {
    PartState& ps = ix->part;
    cudaStream_t s = ix->ctx->compute_stream;
    const uint64_t P = ps.view.P;
    const uint32_t shift2 = ps.view.shift2, sub_bits = ps.view.sub_bits;
    cudaError_t e = cudaMalloc((void**)&ps.view.keybuf, P * ps.view.cap * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ps.view.cursor, P * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ps.view.ctr, ((size_t)8 << shift2) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(ps.view.ctr, 0, ((size_t)8 << shift2) * sizeof(uint32_t), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(ps.view.cursor, 0, P * sizeof(unsigned long long), s);
    if (e == cudaSuccess && sub_bits) e = cudaMalloc((void**)&ps.view.keybuf2, (ps.view.cap2 << sub_bits) * sizeof(uint64_t));
    if (e == cudaSuccess && sub_bits) e = cudaMalloc((void**)&ps.view.cursor2, sizeof(unsigned long long) << sub_bits);
    if (e != cudaSuccess) return e;
    if (cudaMalloc((void**)&ps.d_round_keys, sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc((void**)&ps.h_round_keys, sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&ps.ev_round, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();  // no feedback on the rounds' fill: they stay as long as the lists are
        cudaFree(ps.d_round_keys);
        ps.d_round_keys = nullptr;
    }
    return cudaSuccess;
}
void vg::part_free_lists(vg_index* ix) {
    PartState& ps = ix->part;
    cudaFree(ps.view.keybuf);
    cudaFree(ps.view.cursor);
    cudaFree(ps.view.ctr);
    cudaFree(ps.view.keybuf2);
    cudaFree(ps.view.cursor2);
    ps.view.keybuf = ps.view.keybuf2 = nullptr;
    ps.view.cursor = ps.view.cursor2 = nullptr;
    ps.view.ctr = nullptr;
}

static int part_setup(vg_index* ix) {
    vg_ctx* c = ix->ctx;
    const char* env = getenv("VG_PARTITION");
    const int force = env ? atoi(env) : -1;
    if (force == 0) return VG_OK;
    const uint64_t table_bytes = 32ull * ix->view.nbuckets;
    if (force != 1 && table_bytes < (8ull << 20)) return VG_OK;   // tiny table: direct probing
    // Slices of 32 MB (L2-resident with room to spare); smaller ones for small tables so that at least three
    // slices exist -- measured on B200, the scatter + sweep pipeline (pre-filter, fire-and-forget side counters)
    // also beats direct probing with CAS when the whole table fits L2 (68 vs 52 G k-mers/s on a 96 MB table).
    // A CTA tile of 4 Ki positions scatters well into a few dozen bins, not into hundreds: beyond
    // VG_TWO_LEVEL_FROM slices (default 96) the scatter bins by coarse partitions of 2^sub_bits slices (at most 64
    // of them) and the sweep re-scatters each coarse list into its slices' lists before probing them.
    vg::PartGeometry geo;
    if (!vg::part_geometry(ix->view.nbuckets, 1, geo)) return VG_OK;  // direct probing
    const uint32_t shift2 = geo.shift2, sub_bits = geo.sub_bits, shift = shift2 + sub_bits;
    const uint64_t P = geo.P_local;
    // A round = the bases scattered before one sweep of the table: 1 G to start with (10 bytes of key list
    // per base); part_grow doubles it when a sample turns out to be longer (VG_ROUND_KEYS pins it).
    uint64_t round_keys = 1024ull << 20, slack = 65536;
    ix->part.may_grow = true;
    if (const char* e = getenv("VG_ROUND_KEYS")) {
        round_keys = strtoull(e, nullptr, 10) >= 4096 ? strtoull(e, nullptr, 10) : round_keys;
        ix->part.may_grow = false;
    }
    if (const char* e = getenv("VG_PART_SLACK")) slack = strtoull(e, nullptr, 10);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    while (round_keys > (8u << 20) && round_keys * 10 > free_b / 4) round_keys >>= 1;  // 10 B/key, <= 1/4 of free HBM
    round_keys &= ~4095ull;
    PartState& ps = ix->part;
    ps.view.P = (uint32_t)P;
    ps.view.shift = shift;
    ps.view.shift2 = shift2;
    ps.view.sub_bits = sub_bits;
    ps.view.cap = (round_keys / P) * 5 / 4 + slack;
    ps.view.cap2 = sub_bits ? (ps.view.cap >> sub_bits) * 5 / 4 + slack : 0;
    ps.slack = slack;
    ps.view.world = 1;
    ps.view.rank = 0;
    ps.view.P_local = (uint32_t)P;
    ps.round_keys = round_keys;
    cudaError_t e = vg::part_alloc_lists(ix);
    if (e != cudaSuccess) {  // not enough memory for the key buffers: fall back to direct probing
        vg::part_free_lists(ix);
        ps = PartState();
        cudaGetLastError();
        return VG_OK;
    }
    {   // slot order: rank_base (exclusive scan of the bucket occupancies), the count vector, key -> slot
        const uint32_t nblocks = (uint32_t)((ix->view.nbuckets + 1023ull) / 1024);
        uint32_t* d_sums = nullptr;
        unsigned long long* d_total = nullptr;
        unsigned long long total = 0;
        e = cudaMalloc((void**)&ix->view.rank_base, (size_t)ix->view.nbuckets * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_sums, ((size_t)nblocks + 1) * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_total, sizeof(unsigned long long));
        if (e == cudaSuccess) e = vg::launch_rank_scan(ix->view, d_sums, d_total, c->compute_stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, c->compute_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute_stream);
        cudaFree(d_sums);
        cudaFree(d_total);
        ix->m_slots = total;
        if (e == cudaSuccess) e = cudaMalloc((void**)&ix->view.cvec, (size_t)((total + 31) & ~15ull));
        if (e == cudaSuccess) e = cudaMemsetAsync(ix->view.cvec, 0, (size_t)((total + 31) & ~15ull), c->compute_stream);
        if (e == cudaSuccess) e = cudaMalloc((void**)&ix->d_perm, std::max<uint64_t>(ix->n, 1) * sizeof(uint32_t));
        if (e == cudaSuccess) e = vg::launch_slot_perm(ix->view, ix->d_key56, ix->n, ix->d_perm, c->compute_stream);
        if (e != cudaSuccess) {
            vg::part_free_lists(ix);
            cudaFree(ix->view.rank_base);
            cudaFree(ix->view.cvec);
            cudaFree(ix->d_perm);
            ix->view.rank_base = nullptr;
            ix->view.cvec = nullptr;
            ix->d_perm = nullptr;
            ix->m_slots = 0;
            ps = PartState();
            cudaGetLastError();
            return VG_OK;
        }
    }
    ps.enabled = true;
    {
        int rc = vg::fetch_slice_ranks(ix);
        if (rc) return rc;
    }
    {
        uint64_t bytes = 0;
        uint32_t span = 4;
        vg::prefilter_plan(ix->n, ix->view.k, bytes, span);
        if (bytes >= 64) {
            const uint32_t nwords = (uint32_t)std::min<uint64_t>(bytes / 4, 0x7fffffffull);
            if (cudaMalloc((void**)&ps.d_filter, (size_t)nwords * 4) == cudaSuccess) {
                CU(cudaMemsetAsync(ps.d_filter, 0, (size_t)nwords * 4, c->compute_stream));
                CU(vg::launch_prefilter_build(ps.d_filter, nwords, ix->d_key56, ix->n, ix->view.k, span, c->compute_stream));
                CU(cudaStreamSynchronize(c->compute_stream));
                ps.filter.words = ps.d_filter;
                ps.filter.nwords = nwords;
                ps.filter.span = span;
                if ((size_t)nwords * 4 <= (64ull << 20)) vg::pin_in_l2(c, ps.d_filter, (size_t)nwords * 4);
            } else {
                cudaGetLastError();
            }
        }
    }
    // the keys themselves are no longer needed: every later lookup by key goes through d_perm
    CU(cudaStreamSynchronize(c->compute_stream));
    cudaFree(ix->d_key56);
    ix->d_key56 = nullptr;
    return VG_OK;
}

// vg_index_set_timing: bracket a phase with events and add its device time up (synchronises: diagnostic only)
static void phase_begin(PartState& ps, cudaStream_t s) {
    if (ps.timing) cudaEventRecord(ps.ev0, s);
}
static void phase_end(PartState& ps, cudaStream_t s, double& acc, uint64_t& n) {
    if (!ps.timing) return;
    float ms = 0;
    if (cudaEventRecord(ps.ev1, s) == cudaSuccess && cudaEventSynchronize(ps.ev1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ps.ev0, ps.ev1) == cudaSuccess) {
        acc += ms;
        n += 1;
    }
    cudaGetLastError();
}

static int part_flush(vg_index* ix, cudaStream_t s) {
    PartState& ps = ix->part;
    if (ix->sharded) return VG_OK;  // a sharded round ends only in the collective calls (vg_count_flush / _end)
    if (!ps.enabled || ps.pending == 0) return VG_OK;
    if (ps.d_round_keys) {  // how many keys did this much text leave? (read back later, without waiting)
        CU(vg::launch_sum_cursors(ps.view.cursor, ps.view.P, ps.view.cap, ps.d_round_keys, &ix->d_misc->stats, s));
        ix->launches += 1;
        if (ps.round_pending == 0) {
            CU(cudaMemcpyAsync(ps.h_round_keys, ps.d_round_keys, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            CU(cudaEventRecord(ps.ev_round, s));
            ps.round_pending = ps.pending;
        }
    }
    if (getenv("VG_ROUND_DEBUG"))
        fprintf(stderr, "[vg round] sweep after %.3f G bytes of text; lists for %.3f G keys (%u x %llu), key share %.3f, last round %llu keys of %llu bytes\n",
                ps.pending * 1e-9, ps.round_keys * 1e-9, ps.view.P, (unsigned long long)ps.view.cap, ps.key_share,
                ps.h_round_keys ? *ps.h_round_keys : 0ull, (unsigned long long)ps.round_pending);
    phase_begin(ps, s);
    CU(vg::launch_probe_partitions(ix->view, ps.view, ps.slice_rank.empty() ? nullptr : ps.slice_rank.data(), &ix->d_misc->stats,
                                   ix->ctx->nsm, s));
    phase_end(ps, s, ps.ms_sweep, ps.n_sweep);
    ix->launches += vg::sweep_launches(ix->view, ps.view);
    ps.pending = 0;
    return VG_OK;
}

// A sample did not fit one round.  Every sweep costs a pass over the whole table plus ~2 P launches (about
// 1.1 ms on the chr20 index), so the next rounds are made twice as long -- up to 4 G bases, and only while
// the key lists stay within a quarter of the free HBM.  Called right after a sweep: the lists are empty.
static void part_grow(vg_index* ix, cudaStream_t s) {
    PartState& ps = ix->part;
    if (!ps.may_grow || ix->sharded || ps.round_keys >= (4096ull << 20)) return;
    const uint64_t P = ps.view.P, round2 = ps.round_keys * 2, cap2 = (round2 / P) * 5 / 4 + ps.slack;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || (cap2 - ps.view.cap) * P * sizeof(uint64_t) > free_b / 4) {
        cudaGetLastError();
        ps.may_grow = false;
        return;
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) return;
    uint64_t* bigger = nullptr;
    if (cudaMalloc((void**)&bigger, P * cap2 * sizeof(uint64_t)) != cudaSuccess) {
        cudaGetLastError();
        ps.may_grow = false;
        return;
    }
    if (ps.view.sub_bits) {  // the slices' own lists grow with the coarse ones
        const uint64_t sub_cap = (cap2 >> ps.view.sub_bits) * 5 / 4 + ps.slack;
        uint64_t* bigger2 = nullptr;
        if (cudaMalloc((void**)&bigger2, (sub_cap << ps.view.sub_bits) * sizeof(uint64_t)) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(bigger);
            ps.may_grow = false;
            return;
        }
        cudaFree(ps.view.keybuf2);
        ps.view.keybuf2 = bigger2;
        ps.view.cap2 = sub_cap;
    }
    cudaFree(ps.view.keybuf);
    ps.view.keybuf = bigger;
    ps.view.cap = cap2;
    ps.round_keys = round2;
}

// Text a round may take (see PartState::key_share).
static uint64_t round_limit(PartState& ps) {
    if (ps.round_pending && cudaEventQuery(ps.ev_round) == cudaSuccess) {
        if (ps.round_pending >= (64u << 20))  // short rounds say little
            ps.key_share = std::min(1.0, std::max(0.05, (double)*ps.h_round_keys / (double)ps.round_pending));
        ps.round_pending = 0;
    }
    cudaGetLastError();
    static const bool adapt = [] { const char* e = getenv("VG_ADAPTIVE_ROUNDS"); return !(e && atoi(e) == 0); }();
    if (!adapt || !ps.may_grow) return ps.round_keys;
    const double lim = (double)ps.round_keys / std::min(1.0, ps.key_share * 1.25);
    return (uint64_t)std::min(lim, 16.0 * 1073741824.0) & ~4095ull;
}

// Count every k-mer of a device-resident chunk on stream s (direct or partitioned).
static int count_device_chunk(vg_index* ix, const uint8_t* d_bases, uint64_t nbytes, cudaStream_t s,
                              const unsigned int* d_skip = nullptr) {
    vg_ctx* c = ix->ctx;
    PartState& ps = ix->part;
    if (!ps.enabled) {
        CU(vg::launch_count(ix->view, d_bases, nbytes, &ix->d_misc->stats, c->ctas_per_sm, c->nsm, s, d_skip));
        ix->launches += 1;
        return VG_OK;
    }
    const int64_t tile_bytes = 4096;
    const int64_t T = vg::chunk_tiles(d_bases, nbytes);
    int64_t t = 0;
    while (t < T) {
        const uint64_t limit = ix->sharded ? ps.round_keys : round_limit(ps);
        int64_t room = limit > ps.pending ? (int64_t)((limit - ps.pending) / tile_bytes) : 0;
        if (ix->sharded && room < T - t)
            return fail(VG_E_STATE, "sharded index: the round is full (%llu of %llu bytes); call vg_count_flush on every rank",
                        (unsigned long long)ps.pending, (unsigned long long)ps.round_keys);
        if (room < 64 && ps.pending) {
            int rc = part_flush(ix, s);
            if (rc) return rc;
            part_grow(ix, s);
            room = (int64_t)(round_limit(ps) / tile_bytes);
        }
        const int64_t nt = std::min<int64_t>(T - t, room);
        phase_begin(ps, s);
        CU(vg::launch_scatter(ix->view, ps.view, ps.filter, d_bases, nbytes, t, nt, &ix->d_misc->stats, c->nsm, s, d_skip));
        phase_end(ps, s, ps.ms_scatter, ps.n_scatter);
        ix->launches += 1;
        ps.pending += (uint64_t)nt * tile_bytes;
        t += nt;
    }
    return VG_OK;
}

// Scratch of the on-device FASTQ parser, sized for one staging chunk.
int vg::ctx_ensure_fastq(vg_ctx* c) {
    if (c->d_masked) return VG_OK;
    vg::FastqScratch& f = c->fq;
    f.max_tiles = (uint32_t)(c->chunk_bytes / 4096 + 2);
    f.max_lines = (uint32_t)(c->chunk_bytes / 8 + 64);  // four-line FASTQ has lines of ~100 bytes; beyond this: host parser
    // fastq_mask_kernel writes whole 4 KiB tiles: up to one tile beyond the block's last byte
    CU(cudaMalloc((void**)&c->d_masked, c->chunk_bytes + 8192));
    CU(cudaMalloc((void**)&f.tile_count, (size_t)f.max_tiles * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&f.tile_base, (size_t)f.max_tiles * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&f.nlpos, (size_t)f.max_lines * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&f.blk, sizeof(vg::FastqBlockState)));
    return VG_OK;
}

int vg::enqueue_raw_piece(vg_index* ix, int si, uint64_t len, vg::FastqFileState* d_file, uint32_t block_no) {
    vg_ctx* c = ix->ctx;
    vg::StageSlot& sl = c->ring[(size_t)si];
    CU(cudaMemcpyAsync(sl.d_buf, sl.h_pin, len, cudaMemcpyHostToDevice, c->copy_stream));
    ix->h2d_bytes += len;
    CU(cudaEventRecord(sl.copied, c->copy_stream));
    CU(cudaStreamWaitEvent(c->compute_stream, sl.copied, 0));
    CU(vg::launch_fastq_block(sl.d_buf, (uint32_t)len, c->d_masked, c->fq, d_file, block_no, c->compute_stream));
    ix->launches += 5;
    return vg::enqueue_piece(ix, si, nullptr, len, &d_file->bad);
}

int vg::commit_stripped(vg_index* ix, vg::FastqFileState* d_file, uint64_t bases, bool whole_block) {
    CU(vg::launch_fastq_strip_commit(d_file, bases, whole_block, ix->ctx->compute_stream));
    ix->launches += 1;
    return VG_OK;
}

// Enqueue one staged piece that already sits in ring slot `si`'s pinned buffer (or at `src`); src == nullptr:
// the piece is already on the device in ctx->d_masked (and is skipped if *d_skip turns out non-zero).
int vg::enqueue_piece(vg_index* ix, int si, const char* src, uint64_t len, const unsigned int* d_skip) {
    vg_ctx* c = ix->ctx;
    vg::StageSlot& sl = c->ring[(size_t)si];
    if (src) {
        ix->h2d_bytes += len;
        CU(cudaMemcpyAsync(sl.d_buf, src, len, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(sl.copied, c->copy_stream));
        CU(cudaStreamWaitEvent(c->compute_stream, sl.copied, 0));
    }
    int rc = count_device_chunk(ix, src ? sl.d_buf : c->d_masked, len, c->compute_stream, d_skip);
    if (rc) return rc;
    CU(cudaEventRecord(sl.done, c->compute_stream));
    // Staged input arrives at PCIe speed, slower than the kernels: sweep every 384 M k-mers (tunable:
    // VG_STAGED_ROUND_KEYS) so the sweeps hide under the copies and only a short one is left after
    // the last piece.
    static const uint64_t staged_round = [] {
        const char* e = getenv("VG_STAGED_ROUND_KEYS");
        const uint64_t v = e ? strtoull(e, nullptr, 10) : 0;
        return v >= 4096 ? v : (384ull << 20);  // measured flat between 256 M and 1 G, worse below
    }();
    if (ix->part.enabled && !ix->sharded && ix->part.pending >= std::min<uint64_t>(round_limit(ix->part), staged_round)) {
        rc = part_flush(ix, c->compute_stream);
        if (rc) return rc;
    }
    sl.busy = true;
    return VG_OK;
}

// Counts in the key order given at create, on the device: a gather through d_perm out of the count vector
// (partitioned) or one probe per key (direct probing of a tiny table, counts in the slots).
cudaError_t vg::counts_in_key_order(vg_index* ix, void* d_out, int elem_bytes, cudaStream_t s) {
    if (ix->view.cvec) return vg::launch_gather_counts(ix->view.cvec, ix->d_perm, nullptr, ix->n, d_out, elem_bytes, s);
    return vg::launch_extract(ix->view, ix->d_key56, nullptr, ix->n, d_out, elem_bytes, s);
}

extern "C" {

const char* vg_last_error(void) { return g_err.c_str(); }
int vg_version(void) { return 100; }

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
int vg_ctx_create(int device, int buffer_mb, vg_ctx** out) {
    if (!out) return fail(VG_E_INVALID, "vg_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(VG_E_CUDA, "no CUDA device available (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(VG_E_INVALID, "device %d out of range [0,%d)", device, ndev);
    if (buffer_mb <= 0) buffer_mb = 64;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(VG_E_CUDA, "device %d is sm_%d%d; libvgb200 is built for sm_100a only", device, prop.major,
                    prop.minor);
    if (const char* e = getenv("VG_L2_FETCH")) {  // tuning knob: L2 fetch granularity hint (32/64/128 bytes)
        int gran = atoi(e);
        if (gran == 32 || gran == 64 || gran == 128) CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran));
    }
    vg_ctx* c = new vg_ctx();
    c->device = device;
    if (const char* e = getenv("VG_CTAS_PER_SM")) c->ctas_per_sm = atoi(e);
    c->nsm = vg::sm_count(device);
    c->chunk_bytes = (size_t)buffer_mb << 20;
    cudaError_t se = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&c->compute_stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        delete c;
        return fail(VG_E_CUDA, "vg_ctx_create: %s", cudaGetErrorString(se));
    }
    c->own_compute_stream = c->compute_stream;
    *out = c;
    return VG_OK;
}

// Device buffers of the staging ring; the pinned twins are only allocated for callers that hand
// over pageable memory (a pinned source is DMA-ed from where it lies).
static int ctx_ensure_ring(vg_ctx* c, int nslots, bool need_pinned) {
    while ((int)c->ring.size() < nslots) {
        vg::StageSlot s;
        CU(cudaMalloc((void**)&s.d_buf, c->chunk_bytes + 256));
        CU(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        c->ring.push_back(s);
    }
    if (need_pinned)
        for (auto& s : c->ring)
            if (!s.h_pin) CU(cudaHostAlloc((void**)&s.h_pin, c->chunk_bytes + 256, cudaHostAllocDefault));
    return VG_OK;
}

int vg_ctx_destroy(vg_ctx* c) {
    if (!c) return VG_OK;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    for (auto& s : c->ring) {
        cudaFree(s.d_buf);
        if (s.h_pin) cudaFreeHost(s.h_pin);
        cudaEventDestroy(s.copied);
        cudaEventDestroy(s.done);
    }
    cudaFree(c->d_masked);
    cudaFree(c->fq.tile_count);
    cudaFree(c->fq.tile_base);
    cudaFree(c->fq.nlpos);
    cudaFree(c->fq.blk);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->own_compute_stream);
    delete c;
    return VG_OK;
}

int vg_ctx_device(const vg_ctx* c) { return c ? c->device : -1; }

int vg_ctx_set_stream(vg_ctx* c, void* cuda_stream) {
    if (!c) return fail(VG_E_INVALID, "ctx is NULL");
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->compute_stream));
    c->compute_stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_compute_stream;
    if (c->has_l2_window) {  // the caller's stream inherits the pre-filter's L2 persisting window
        cudaStreamAttrValue av{};
        av.accessPolicyWindow = c->l2_window;
        cudaStreamSetAttribute(c->compute_stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    return VG_OK;
}

int vg_probe_random_sectors(vg_ctx* c, uint64_t table_bytes, uint32_t rounds, double* gbytes_per_s,
                            double* sectors_per_s) {
    if (!c) return fail(VG_E_INVALID, "ctx is NULL");
    if (table_bytes < (1u << 20) || rounds == 0) return fail(VG_E_INVALID, "table too small or rounds == 0");
    DeviceGuard g(c->device);
    uint64_t nb = table_bytes / 32;
    if (nb >= 0xffffffffull) nb = 0xfffffffeull;
    uint64_t* table = nullptr;
    unsigned long long* sink = nullptr;
    CU(cudaMalloc((void**)&table, nb * 32));
    cudaError_t e = cudaMalloc((void**)&sink, 8);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t s = c->compute_stream;
    int grid = c->nsm * 8;
    float ms = 0;
    if (e == cudaSuccess) e = cudaMemsetAsync(table, 0x5a, nb * 32, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(sink, 0, 8, s);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = vg::launch_random_sectors(table, (uint32_t)nb, rounds / 4 + 1, grid, sink, s);  // warm-up
    if (e == cudaSuccess) e = cudaEventRecord(e0, s);
    if (e == cudaSuccess) e = vg::launch_random_sectors(table, (uint32_t)nb, rounds, grid, sink, s);
    if (e == cudaSuccess) e = cudaEventRecord(e1, s);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(table);
    cudaFree(sink);
    if (e != cudaSuccess) return fail(VG_E_CUDA, "vg_probe_random_sectors: %s", cudaGetErrorString(e));
    double loads = (double)grid * vg::kCtaThreadsHost * (double)rounds * vg::probe_batch();
    if (sectors_per_s) *sectors_per_s = loads / (ms * 1e-3);
    if (gbytes_per_s) *gbytes_per_s = loads * 32.0 / (ms * 1e-3) / 1e9;
    return VG_OK;
}

int vg_ctx_synchronize(vg_ctx* c) {
    if (!c) return fail(VG_E_INVALID, "ctx is NULL");
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->compute_stream));
    return VG_OK;
}

// ---------------------------------------------------------------------------
// index
// ---------------------------------------------------------------------------
// keys: host memory, or (keys_on_device) device memory of ctx's GPU
static int index_create(vg_ctx* c, const uint64_t* keys, bool keys_on_device, uint64_t n, uint32_t k, double load_factor,
                        vg_index** out) {
    if (!c || !out || (!keys && n)) return fail(VG_E_INVALID, "vg_index_create: NULL argument");
    *out = nullptr;
    if (k < 1 || k > 28) return fail(VG_E_INVALID, "k=%u outside 1..28 (reference asserts k<=28, src/kmer.cpp:124)", k);
    if (load_factor <= 0) load_factor = 0.3;
    if (load_factor > 0.9) return fail(VG_E_INVALID, "load_factor %.3f > 0.9", load_factor);
    DeviceGuard g(c->device);
    uint64_t nb64 = (uint64_t)((double)n / (4.0 * load_factor)) + 1;
    if (nb64 < 64) nb64 = 64;
    if (nb64 >= 0xffffffffull || n >= 0xffffffffull)
        return fail(VG_E_INVALID, "index of %llu keys needs too many buckets", (unsigned long long)n);

    vg_index* ix = new vg_index();
    ix->ctx = c;
    ix->n = n;
    ix->view.k = k;
    ix->view.mask = (1ULL << (2 * k)) - 1;
    ix->view.nbuckets = (uint32_t)nb64;
    ix->view.nb_total = (uint32_t)nb64;
    ix->view.b_base = 0;
    auto bail = [&](int code) {
        vg_index_destroy(ix);
        return code;
    };
#define CUB(expr)                                                                      \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess)                                                        \
            return bail(fail(e__ == cudaErrorMemoryAllocation ? VG_E_NOMEM : VG_E_CUDA, "%s: %s", #expr, \
                             cudaGetErrorString(e__)));                                \
    } while (0)
    uint64_t nslots = 4ull * ix->view.nbuckets;
    CUB(cudaMalloc((void**)&ix->view.slots, nslots * sizeof(uint64_t)));
    CUB(cudaMalloc((void**)&ix->d_key56, std::max<uint64_t>(n, 1) * sizeof(uint64_t)));
    CUB(cudaMalloc((void**)&ix->d_counts, std::max<uint64_t>(n, 4)));
    CUB(cudaMalloc((void**)&ix->d_misc, sizeof(vg::DeviceMisc)));
    CUB(cudaMemset(ix->d_misc, 0, sizeof(vg::DeviceMisc)));
    cudaStream_t s = c->compute_stream;
    CUB(vg::launch_table_fill_empty(ix->view.slots, nslots, s));

    if (keys_on_device) {  // validated and un-hashed where they lie
        unsigned long long* d_bad = nullptr;
        unsigned long long bad = ~0ull;
        CUB(cudaMalloc((void**)&d_bad, sizeof bad));
        cudaError_t e = cudaMemcpyAsync(d_bad, &bad, sizeof bad, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = vg::launch_keys_to_key56(keys, n, k, ix->view.mask, ix->d_key56, d_bad, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(d_bad);
        CUB(e);
        if (bad != ~0ull)
            return bail(fail(VG_E_INVALID, "keys[%llu]: low byte is not k=%u or the hash exceeds 2k bits (src/kmer.cpp:138)", bad, k));
    } else {
        // keys -> key56, staged through a bounded host buffer
        const uint64_t piece = 1ull << 22;
        std::vector<uint64_t> tmp((size_t)std::min<uint64_t>(piece, std::max<uint64_t>(n, 1)));
        for (uint64_t off = 0; off < n; off += piece) {
            uint64_t m = std::min<uint64_t>(piece, n - off);
            for (uint64_t i = 0; i < m; ++i) {
                uint64_t key = keys[off + i];
                if ((key & 0xffu) != k)
                    return bail(fail(VG_E_INVALID, "keys[%llu]=0x%llx: low byte is not k=%u (src/kmer.cpp:138)",
                                     (unsigned long long)(off + i), (unsigned long long)key, k));
                uint64_t h = key >> 8;
                if (h > ix->view.mask)
                    return bail(fail(VG_E_INVALID, "keys[%llu]: hash exceeds 2k bits", (unsigned long long)(off + i)));
                tmp[(size_t)i] = h;
            }
            CUB(cudaMemcpyAsync(ix->d_key56 + off, tmp.data(), m * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
            CUB(cudaStreamSynchronize(s));
        }
        CUB(vg::launch_unhash(ix->d_key56, n, ix->view.mask, s));
    }
    CUB(vg::launch_insert(ix->view, ix->d_key56, n, &ix->d_misc->report, s));
    vg::DeviceMisc misc;
    CUB(cudaMemcpyAsync(&misc, ix->d_misc, sizeof misc, cudaMemcpyDeviceToHost, s));
    CUB(cudaStreamSynchronize(s));
    if (misc.report.failed) return bail(fail(VG_E_NOMEM, "index build: %llu keys found no slot", misc.report.failed));
    ix->duplicates = misc.report.duplicates;
#undef CUB
    int prc = part_setup(ix);
    if (prc != VG_OK) return bail(prc);
    *out = ix;
    return VG_OK;
}

int vg_index_create(vg_ctx* c, const uint64_t* keys, uint64_t n, uint32_t k, double load_factor, vg_index** out) {
    return index_create(c, keys, false, n, k, load_factor, out);
}
int vg_index_create_device(vg_ctx* c, const uint64_t* dev_keys, uint64_t n, uint32_t k, double load_factor, vg_index** out) {
    return index_create(c, dev_keys, true, n, k, load_factor, out);
}

int vg_index_destroy(vg_index* ix) {
    if (!ix) return VG_OK;
    DeviceGuard g(ix->ctx->device);
    cudaDeviceSynchronize();
    if (!ix->sharded) {  // a sharded index keeps these in its group's arena
        cudaFree(ix->view.slots);
        cudaFree(ix->d_counts);
        cudaFree(ix->part.view.keybuf);
        cudaFree(ix->view.rank_base);
        if (!ix->replica_of) cudaFree(ix->view.cvec);  // a replica group keeps the count vectors in its arena
    }
    cudaFree(ix->d_perm);
    cudaFree(ix->d_key56);
    cudaFree(ix->d_idx);
    cudaFree(ix->d_combined);
    cudaFree(ix->d_flags);
    cudaFree(ix->d_hist);
    cudaFree(ix->d_misc);
    cudaFree(ix->part.view.cursor);
    cudaFree(ix->part.view.ctr);
    cudaFree(ix->part.view.keybuf2);
    cudaFree(ix->part.view.cursor2);
    cudaFree(ix->part.d_filter);
    cudaFree(ix->part.d_round_keys);
    if (ix->part.h_round_keys) cudaFreeHost(ix->part.h_round_keys);
    if (ix->part.ev_round) cudaEventDestroy(ix->part.ev_round);
    if (ix->part.ev0) cudaEventDestroy(ix->part.ev0);
    if (ix->part.ev1) cudaEventDestroy(ix->part.ev1);
    delete ix;
    return VG_OK;
}

uint64_t vg_index_size(const vg_index* ix) { return ix ? ix->n : 0; }
uint64_t vg_index_duplicates(const vg_index* ix) { return ix ? ix->duplicates : 0; }
uint64_t vg_count_h2d_bytes(const vg_index* ix) { return ix ? ix->h2d_bytes : 0; }

int vg_index_set_timing(vg_index* ix, int on) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    PartState& ps = ix->part;
    DeviceGuard g(ix->ctx->device);
    if (on && !ps.ev0) {
        CU(cudaEventCreate(&ps.ev0));
        CU(cudaEventCreate(&ps.ev1));
    }
    ps.timing = on != 0;
    ps.ms_scatter = ps.ms_sweep = 0;
    ps.n_scatter = ps.n_sweep = 0;
    return VG_OK;
}
int vg_index_timing(const vg_index* ix, double* scatter_ms, double* sweep_ms, uint64_t* scatter_launches, uint64_t* sweeps) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    if (scatter_ms) *scatter_ms = ix->part.ms_scatter;
    if (sweep_ms) *sweep_ms = ix->part.ms_sweep;
    if (scatter_launches) *scatter_launches = ix->part.n_scatter;
    if (sweeps) *sweeps = ix->part.n_sweep;
    return VG_OK;
}
uint32_t vg_index_partitions(const vg_index* ix) { return ix && ix->part.enabled ? ix->part.view.P : 0; }
uint32_t vg_index_slices(const vg_index* ix) {
    if (!ix || !ix->part.enabled) return 0;
    return (uint32_t)(((uint64_t)ix->view.nbuckets + ((1ull << ix->part.view.shift2) - 1)) >> ix->part.view.shift2);
}
uint64_t vg_index_launches(const vg_index* ix) { return ix ? ix->launches : 0; }
uint64_t vg_index_table_bytes(const vg_index* ix) { return ix ? 32ull * ix->view.nbuckets : 0; }

// ---------------------------------------------------------------------------
// count phase
// ---------------------------------------------------------------------------
int vg_count_begin(vg_index* ix) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    if (ix->view.cvec) CU(cudaMemsetAsync(ix->view.cvec, 0, (size_t)ix->m_slots, c->compute_stream));  // the table itself is read-only
    else CU(vg::launch_clear_counts(ix->view, c->compute_stream));
    CU(cudaMemsetAsync(&ix->d_misc->stats, 0, sizeof(vg::CountStats), c->compute_stream));
    ix->h2d_bytes = 0;
    if (ix->part.enabled) {
        CU(cudaMemsetAsync(ix->part.view.cursor, 0, ix->part.view.P * sizeof(unsigned long long), c->compute_stream));
        ix->part.pending = 0;
    }
    if (ix->sharded) {  // no peer may probe this table (a key whose list is full) before it is cleared
        int rc = vg_comm_barrier(ix->comm);
        if (rc) return rc;
    }
    ix->counting = true;
    return VG_OK;
}

int vg_count_submit_device(vg_index* ix, const void* dev_bases, uint64_t nbytes, void* cuda_stream) {
    if (!ix || (!dev_bases && nbytes)) return fail(VG_E_INVALID, "vg_count_submit_device: NULL argument");
    if (!ix->counting) return fail(VG_E_STATE, "vg_count_submit_device before vg_count_begin");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->compute_stream;
    if (ix->part.enabled && s != c->compute_stream)
        return fail(VG_E_INVALID, "a partitioned index counts on the context stream only: use vg_ctx_set_stream");
    if (cuda_stream && s != c->compute_stream) {  // order after vg_count_begin's clears on the context stream
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU(cudaEventRecord(ev, c->compute_stream));
        CU(cudaStreamWaitEvent(s, ev, 0));
        CU(cudaEventDestroy(ev));
        ix->foreign_streams = true;
    }
    return count_device_chunk(ix, (const uint8_t*)dev_bases, nbytes, s);
}

int vg_count_submit(vg_index* ix, const char* host_bases, uint64_t nbytes) {
    if (!ix || (!host_bases && nbytes)) return fail(VG_E_INVALID, "vg_count_submit: NULL argument");
    if (!ix->counting) return fail(VG_E_STATE, "vg_count_submit before vg_count_begin");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, host_bases) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    // A deep ring lets the copy engine run ahead while a probe sweep occupies the compute stream.
    const int want = pinned ? (int)std::min<uint64_t>(12, std::max<uint64_t>(3, (768ull << 20) / c->chunk_bytes)) : 3;
    int rc = ctx_ensure_ring(c, std::max<int>(want, (int)c->ring.size()), !pinned);
    if (rc) return rc;
    uint64_t off = 0;
    int last = -1;
    while (off < nbytes) {
        uint64_t len = std::min<uint64_t>(c->chunk_bytes, nbytes - off);
        if (off + len < nbytes) {  // cut at a read boundary
            const void* nl = memrchr(host_bases + off, '\n', (size_t)len);
            if (!nl) return fail(VG_E_INVALID, "a read is longer than the %zu-byte staging buffer", c->chunk_bytes);
            len = (uint64_t)((const char*)nl - (host_bases + off)) + 1;
        }
        int si = c->next_slot;
        c->next_slot = (c->next_slot + 1) % (int)c->ring.size();
        vg::StageSlot& sl = c->ring[(size_t)si];
        if (sl.busy) {
            CU(cudaEventSynchronize(sl.done));
            sl.busy = false;
        }
        const char* src = host_bases + off;
        if (!pinned) {
            memcpy(sl.h_pin, src, (size_t)len);
            src = (const char*)sl.h_pin;
        }
        rc = vg::enqueue_piece(ix, si, src, len);
        if (rc) return rc;
        off += len;
        last = si;
    }
    // "returns once the bytes are consumed from host_bases": a pinned source is DMA-ed from where it lies, so the
    // last copy must have left the caller's buffer before the caller may refill it (pageable sources were copied
    // into the ring's own pinned twins already)
    if (pinned && last >= 0) CU(cudaEventSynchronize(c->ring[(size_t)last].copied));
    return VG_OK;
}

int vg_count_flush(vg_index* ix) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    DeviceGuard g(ix->ctx->device);
    if (ix->sharded) {
        CU(cudaStreamSynchronize(ix->ctx->copy_stream));
        return vg::sharded_flush(ix, ix->ctx->compute_stream);
    }
    return part_flush(ix, ix->ctx->compute_stream);
}

uint64_t vg_count_room(const vg_index* ix) {
    if (!ix || !ix->sharded) return ~0ull;
    // every staged piece of a submit may end in a partly filled 4 KiB tile, which counts as a whole one
    const uint64_t used = ix->part.pending + 4096 * (ix->part.round_keys / ix->ctx->chunk_bytes + 2);
    return used >= ix->part.round_keys ? 0 : ix->part.round_keys - used;
}

int vg_count_stats(vg_index* ix, uint64_t* positions, uint64_t* hits) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    if (ix->foreign_streams) CU(cudaDeviceSynchronize());
    CU(cudaStreamSynchronize(c->copy_stream));
    {
        int frc = part_flush(ix, c->compute_stream);
        if (frc) return frc;
    }
    CU(cudaStreamSynchronize(c->compute_stream));
    vg::CountStats st;
    CU(cudaMemcpy(&st, &ix->d_misc->stats, sizeof st, cudaMemcpyDeviceToHost));
    if (positions) *positions = st.positions;
    if (hits) *hits = st.hits;
    ix->last_keys = st.keys;
    return VG_OK;
}
uint64_t vg_count_keys(const vg_index* ix) { return ix ? ix->last_keys : 0; }

int vg_count_extract_device(vg_index* ix, void* dev_out, int elem_bytes, void* cuda_stream) {
    if (!ix || !dev_out) return fail(VG_E_INVALID, "vg_count_extract_device: NULL argument");
    if (elem_bytes != 1 && elem_bytes != 4) return fail(VG_E_INVALID, "elem_bytes must be 1 or 4");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->compute_stream;
    {
        int frc = part_flush(ix, c->compute_stream);
        if (frc) return frc;
    }
    if (cuda_stream) {  // counting kernels submitted through the context must finish first
        CU(cudaStreamSynchronize(c->copy_stream));
        CU(cudaStreamSynchronize(c->compute_stream));
    }
    if (ix->sharded) return fail(VG_E_STATE, "sharded index: the counts of all keys come from vg_count_end");
    CU(vg::counts_in_key_order(ix, dev_out, elem_bytes, s));
    return VG_OK;
}

int vg_index_set_flags(vg_index* ix, const uint8_t* flags) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    if (!flags) {
        CU(cudaStreamSynchronize(c->compute_stream));
        cudaFree(ix->d_flags);
        ix->d_flags = nullptr;
        return VG_OK;
    }
    if (ix->sharded) return fail(VG_E_STATE, "sharded index: take the histogram of vg_count_end's counts");
    if (!ix->d_flags) CU(cudaMalloc((void**)&ix->d_flags, std::max<uint64_t>(std::max(ix->n, ix->m_slots), 4)));
    if (ix->view.cvec) {  // keep the flags in slot order, next to the counts they select
        uint8_t* d_tmp = nullptr;
        CU(cudaMalloc((void**)&d_tmp, std::max<uint64_t>(ix->n, 4)));
        cudaError_t e = cudaMemcpyAsync(d_tmp, flags, ix->n, cudaMemcpyHostToDevice, c->compute_stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(ix->d_flags, 0, (size_t)ix->m_slots, c->compute_stream);
        if (e == cudaSuccess) e = vg::launch_scatter_bytes(d_tmp, ix->d_perm, ix->n, ix->d_flags, c->compute_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute_stream);
        cudaFree(d_tmp);
        if (e != cudaSuccess) return fail(VG_E_CUDA, "vg_index_set_flags: %s", cudaGetErrorString(e));
        return VG_OK;
    }
    CU(cudaMemcpyAsync(ix->d_flags, flags, ix->n, cudaMemcpyHostToDevice, c->compute_stream));
    CU(cudaStreamSynchronize(c->compute_stream));
    return VG_OK;
}

int vg_count_histogram(vg_index* ix, uint64_t* hist256) {
    if (!ix || !hist256) return fail(VG_E_INVALID, "vg_count_histogram: NULL argument");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->copy_stream));
    int rc = part_flush(ix, c->compute_stream);
    if (rc) return rc;
    if (ix->sharded) return fail(VG_E_STATE, "sharded index: take the histogram of vg_count_end's counts");
    if (!ix->d_hist) CU(cudaMalloc((void**)&ix->d_hist, 256 * sizeof(unsigned long long)));
    if (ix->view.cvec) {  // straight over the count vector: one entry per distinct k-mer of the index
        CU(vg::launch_histogram(ix->view.cvec, ix->d_flags, ix->m_slots, ix->d_hist, c->compute_stream));
    } else {
        CU(vg::launch_extract(ix->view, ix->d_key56, nullptr, ix->n, ix->d_counts, 1, c->compute_stream));
        CU(vg::launch_histogram(ix->d_counts, ix->d_flags, ix->n, ix->d_hist, c->compute_stream));
    }
    unsigned long long h[256];
    CU(cudaMemcpyAsync(h, ix->d_hist, sizeof h, cudaMemcpyDeviceToHost, c->compute_stream));
    CU(cudaStreamSynchronize(c->compute_stream));
    for (int i = 0; i < 256; ++i) hist256[i] = h[i];
    return VG_OK;
}

int vg_count_end(vg_index* ix, uint8_t* c_out, uint64_t* positions, uint64_t* hits) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    if (!ix->counting) return fail(VG_E_STATE, "vg_count_end before vg_count_begin");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    if (ix->sharded) {  // collective: last round, then every rank gets the counts of all keys
        CU(cudaStreamSynchronize(c->copy_stream));
        int frc = vg::sharded_flush(ix, c->compute_stream);
        if (frc == VG_OK) frc = vg::sharded_end(ix, c_out);
        if (frc) return frc;
    }
    int rc = vg_count_stats(ix, positions, hits);
    if (rc) return rc;
    for (auto& sl : c->ring) sl.busy = false;
    if (c_out && ix->n && !ix->sharded) {
        CU(vg::counts_in_key_order(ix, ix->d_counts, 1, c->compute_stream));
        CU(cudaMemcpyAsync(c_out, ix->d_counts, ix->n, cudaMemcpyDeviceToHost, c->compute_stream));
        CU(cudaStreamSynchronize(c->compute_stream));
    }
    ix->counting = false;
    ix->foreign_streams = false;
    return VG_OK;
}

// ---- the result in slot order --------------------------------------------------------------------------
uint64_t vg_index_slots(const vg_index* ix) { return ix ? (ix->view.cvec ? ix->m_slots : ix->n) : 0; }

int vg_index_slot_perm(vg_index* ix, uint32_t* perm_out) {
    if (!ix || (!perm_out && ix->n)) return fail(VG_E_INVALID, "vg_index_slot_perm: NULL argument");
    if (ix->sharded) return fail(VG_E_STATE, "sharded index: counts come in key order from vg_count_end");
    if (!ix->view.cvec) {  // direct probing of a tiny table: slot order is key order
        for (uint64_t i = 0; i < ix->n; ++i) perm_out[i] = (uint32_t)i;
        return VG_OK;
    }
    DeviceGuard g(ix->ctx->device);
    CU(cudaMemcpy(perm_out, ix->d_perm, ix->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return VG_OK;
}

int vg_count_slots_device(vg_index* ix, const uint8_t** dev_counts) {
    if (!ix || !dev_counts) return fail(VG_E_INVALID, "vg_count_slots_device: NULL argument");
    if (ix->sharded) return fail(VG_E_STATE, "sharded index: the counts of all keys come from vg_count_end");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    int rc = part_flush(ix, c->compute_stream);
    if (rc) return rc;
    if (ix->view.cvec) {
        *dev_counts = ix->view.cvec;
    } else {
        CU(vg::launch_extract(ix->view, ix->d_key56, nullptr, ix->n, ix->d_counts, 1, c->compute_stream));
        *dev_counts = ix->d_counts;
    }
    return VG_OK;
}

int vg_count_end_slots(vg_index* ix, uint8_t* c_slots_out, uint64_t* positions, uint64_t* hits) {
    if (!ix) return fail(VG_E_INVALID, "index is NULL");
    if (!ix->counting) return fail(VG_E_STATE, "vg_count_end_slots before vg_count_begin");
    if (ix->sharded) return fail(VG_E_STATE, "sharded index: use vg_count_end");
    vg_ctx* c = ix->ctx;
    DeviceGuard g(c->device);
    int rc = vg_count_stats(ix, positions, hits);
    if (rc) return rc;
    for (auto& sl : c->ring) sl.busy = false;
    if (c_slots_out && vg_index_slots(ix)) {
        const uint8_t* src = nullptr;
        rc = vg_count_slots_device(ix, &src);
        if (rc) return rc;
        CU(cudaMemcpyAsync(c_slots_out, src, vg_index_slots(ix), cudaMemcpyDeviceToHost, c->compute_stream));
        CU(cudaStreamSynchronize(c->compute_stream));
    }
    ix->counting = false;
    ix->foreign_streams = false;
    return VG_OK;
}

// ---------------------------------------------------------------------------
// per-position keys
// ---------------------------------------------------------------------------
int vg_encode_positions_device(vg_ctx* c, const void* dev_bases, uint64_t nbytes, uint32_t k, uint64_t* dev_keys_out,
                               void* cuda_stream) {
    if (!c || (!dev_bases && nbytes) || (!dev_keys_out && nbytes)) return fail(VG_E_INVALID, "NULL argument");
    if (k < 1 || k > 28) return fail(VG_E_INVALID, "k=%u outside 1..28", k);
    DeviceGuard g(c->device);
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->compute_stream;
    CU(vg::launch_positions(k, (const uint8_t*)dev_bases, nbytes, dev_keys_out, s));
    return VG_OK;
}

int vg_encode_positions(vg_ctx* c, const char* host_bases, uint64_t nbytes, uint32_t k, uint64_t* host_keys_out) {
    if (!c || (!host_bases && nbytes) || (!host_keys_out && nbytes)) return fail(VG_E_INVALID, "NULL argument");
    if (k < 1 || k > 28) return fail(VG_E_INVALID, "k=%u outside 1..28", k);
    if (nbytes == 0) return VG_OK;
    DeviceGuard g(c->device);
    uint8_t* d_b = nullptr;
    uint64_t* d_k = nullptr;
    CU(cudaMalloc((void**)&d_b, nbytes + 64));
    cudaError_t e = cudaMalloc((void**)&d_k, nbytes * sizeof(uint64_t));
    if (e != cudaSuccess) {
        cudaFree(d_b);
        return fail(VG_E_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    int rc = VG_OK;
    cudaStream_t s = c->compute_stream;
    if ((e = cudaMemcpyAsync(d_b, host_bases, nbytes, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = vg::launch_positions(k, d_b, nbytes, d_k, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(host_keys_out, d_k, nbytes * sizeof(uint64_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s)) != cudaSuccess)
        rc = fail(VG_E_CUDA, "vg_encode_positions: %s", cudaGetErrorString(e));
    cudaFree(d_b);
    cudaFree(d_k);
    return rc;
}

// ---------------------------------------------------------------------------
// counting Bloom filter
// ---------------------------------------------------------------------------
int vg_cbf_create(vg_ctx* c, uint64_t m, uint32_t num_hashes, const uint64_t* seeds, vg_cbf** out) {
    if (!c || !out || !seeds) return fail(VG_E_INVALID, "vg_cbf_create: NULL argument");
    *out = nullptr;
    if (m == 0) return fail(VG_E_INVALID, "filter size 0");
    if (num_hashes == 0 || num_hashes > 16) return fail(VG_E_INVALID, "num_hashes=%u outside 1..16", num_hashes);
    DeviceGuard g(c->device);
    vg_cbf* f = new vg_cbf();
    f->ctx = c;
    f->view.m = m;
    f->view.num_hashes = num_hashes;
    for (uint32_t i = 0; i < num_hashes; ++i) f->view.seeds[i] = (uint32_t)seeds[i];  // counting_bloom_filter.cpp:91
    unsigned __int128 M = (~(unsigned __int128)0) / m + 1;
    f->view.magic_hi = (uint64_t)(M >> 64);
    f->view.magic_lo = (uint64_t)M;
    uint64_t alloc = (m + 3) & ~3ULL;
    cudaError_t e = cudaMalloc((void**)&f->view.cells, alloc);
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_added, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(f->view.cells, 0, alloc, c->compute_stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->d_added, 0, sizeof(unsigned long long), c->compute_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute_stream);
    if (e != cudaSuccess) {
        vg_cbf_destroy(f);
        return fail(e == cudaErrorMemoryAllocation ? VG_E_NOMEM : VG_E_CUDA, "vg_cbf_create: %s", cudaGetErrorString(e));
    }
    *out = f;
    return VG_OK;
}

int vg_cbf_destroy(vg_cbf* f) {
    if (!f) return VG_OK;
    DeviceGuard g(f->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(f->view.cells);
    cudaFree(f->d_added);
    cudaFree(f->d_seq);
    delete f;
    return VG_OK;
}

int vg_cbf_add_sequence(vg_cbf* f, const char* host_seq, uint64_t len, uint32_t k, uint64_t* added) {
    if (!f || (!host_seq && len)) return fail(VG_E_INVALID, "vg_cbf_add_sequence: NULL argument");
    if (k < 1 || k > 28) return fail(VG_E_INVALID, "k=%u outside 1..28", k);
    if (added) *added = 0;
    if (len == 0) return VG_OK;
    vg_ctx* c = f->ctx;
    DeviceGuard g(c->device);
    if (f->d_seq_cap < len + 256) {
        CU(cudaStreamSynchronize(c->compute_stream));
        cudaFree(f->d_seq);
        f->d_seq = nullptr;
        f->d_seq_cap = 0;
        CU(cudaMalloc((void**)&f->d_seq, len + 256));
        f->d_seq_cap = len + 256;
    }
    CU(cudaMemsetAsync(f->d_added, 0, sizeof(unsigned long long), c->compute_stream));
    // The chromosome is one contiguous device buffer; it is uploaded in tile-aligned pieces so
    // the fill of piece i overlaps the copy of piece i+1.  A piece's k-mers may start in the
    // previous piece: the kernel simply looks back into bytes that are already resident.
    const uint64_t piece = (uint64_t)vg::kTilePieceBytes;
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    int rc = VG_OK;
    for (uint64_t off = 0; off < len && rc == VG_OK; off += piece) {
        uint64_t m = std::min<uint64_t>(piece, len - off);
        cudaError_t e = cudaMemcpyAsync(f->d_seq + off, host_seq + off, m, cudaMemcpyHostToDevice, c->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev, c->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->compute_stream, ev, 0);
        if (e == cudaSuccess)
            e = vg::launch_cbf_add(f->view, k, f->d_seq, off + m, off, f->d_added, c->nsm, c->compute_stream);
        if (e != cudaSuccess) rc = fail(VG_E_CUDA, "vg_cbf_add_sequence: %s", cudaGetErrorString(e));
    }
    cudaEventDestroy(ev);
    if (rc) return rc;
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, f->d_added, sizeof n, cudaMemcpyDeviceToHost, c->compute_stream));
    CU(cudaStreamSynchronize(c->compute_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    if (added) *added = n;
    return VG_OK;
}

int vg_cbf_download(vg_cbf* f, uint8_t* host_filter) {
    if (!f || !host_filter) return fail(VG_E_INVALID, "vg_cbf_download: NULL argument");
    vg_ctx* c = f->ctx;
    DeviceGuard g(c->device);
    CU(cudaMemcpyAsync(host_filter, f->view.cells, f->view.m, cudaMemcpyDeviceToHost, c->compute_stream));
    CU(cudaStreamSynchronize(c->compute_stream));
    return VG_OK;
}

int vg_cbf_query(vg_cbf* f, const uint64_t* host_keys, uint64_t n, uint8_t* count_out, uint8_t* find_out) {
    if (!f || (!host_keys && n)) return fail(VG_E_INVALID, "vg_cbf_query: NULL argument");
    if (n == 0) return VG_OK;
    vg_ctx* c = f->ctx;
    DeviceGuard g(c->device);
    uint64_t* d_k = nullptr;
    uint8_t* d_o = nullptr;
    CU(cudaMalloc((void**)&d_k, n * sizeof(uint64_t)));
    cudaError_t e = cudaMalloc((void**)&d_o, 2 * n);
    cudaStream_t s = c->compute_stream;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_k, host_keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = vg::launch_cbf_query(f->view, d_k, n, d_o, d_o + n, s);
    if (e == cudaSuccess && count_out) e = cudaMemcpyAsync(count_out, d_o, n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && find_out) e = cudaMemcpyAsync(find_out, d_o + n, n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_k);
    cudaFree(d_o);
    if (e != cudaSuccess) return fail(VG_E_CUDA, "vg_cbf_query: %s", cudaGetErrorString(e));
    return VG_OK;
}

int vg_host_alloc(void** out, uint64_t nbytes) {
    if (!out) return fail(VG_E_INVALID, "out is NULL");
    CU(cudaHostAlloc(out, nbytes ? nbytes : 1, cudaHostAllocDefault));
    return VG_OK;
}
int vg_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return VG_OK;
}

}  // extern "C"
