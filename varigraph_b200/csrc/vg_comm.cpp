// vg_comm.cpp -- the multi-GPU half of include/vgb200.h: one process per GPU, peers reach each other's
// memory directly over NVLink / NVSwitch (CUDA IPC mappings), no NCCL and no host in the data path.
//
//   vg_comm                  rank, world, a symmetric arena, a device-side barrier over peer flags
//   vg_count_allreduce       replicated index, reads sharded: min(255, sum of the ranks' u8 counts),
//                            each rank reading its peers' count vectors (1 byte per key on the wire)
//   vg_index_create_sharded  the index is ONE open-addressing table cut into `world` runs of buckets,
//                            one per GPU (index > HBM variant); the scatter kernel's copy-out
//                            stores every k-mer into the key list of the GPU that owns its table slice,
//                            so the all-to-all of k-mers is fused into the kernel that produces them.
//
// The reference has no multi-GPU path (src/fastq_kmer.cu runs one device, main.cu:221); this is the
// BASELINE.json north-star's "hash-partitioned shard with an all-to-all of k-mers ... count arrays
// reduced over NVLink".
#include "../../include/vgb200.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vg_host.h"
#include "vg_internal.h"

using vg::DeviceGuard;
using vg::fail;
#define CU VG_CU

static_assert(sizeof(cudaIpcMemHandle_t) == VG_COMM_HANDLE_BYTES, "VG_COMM_HANDLE_BYTES must match cudaIpcMemHandle_t");

namespace {

constexpr size_t kArenaAlign = 256;
constexpr size_t kFlagsBytes = 256;  // kMaxWorld x u64, at arena offset 0

// Bump allocation out of the symmetric arena: every rank must make the same calls in the same order.
void* arena_alloc(vg_comm* cm, size_t bytes, size_t* off_out) {
    const size_t off = (cm->arena_used + kArenaAlign - 1) & ~(kArenaAlign - 1);
    if (off + bytes > cm->arena_bytes) return nullptr;
    cm->arena_used = off + bytes;
    if (off_out) *off_out = off;
    return cm->arena + off;
}

vg::PeerPtrs peers_at(const vg_comm* cm, size_t off) {
    vg::PeerPtrs pp{};
    for (int r = 0; r < cm->world; ++r) pp.p[r] = cm->peer_base[r] + off;
    return pp;
}

int barrier_on(vg_comm* cm, cudaStream_t s) {
    if (cm->world == 1) return VG_OK;
    cm->epoch += 1;
    if (cm->group) {  // ranks of one process: events and a rendezvous of their host threads (see LocalGroup)
        LocalGroup& g = *cm->group;
        const size_t par = (size_t)(cm->epoch & 1);
        CU(cudaEventRecord(g.ev[(size_t)cm->rank * 2 + par], s));
        {
            std::unique_lock<std::mutex> lk(g.mu);
            const unsigned long long gen = g.generation;
            if (++g.arrived == g.world) {
                g.arrived = 0;
                g.generation += 1;
                g.cv.notify_all();
            } else if (!g.cv.wait_for(lk, std::chrono::nanoseconds(cm->timeout_ns), [&] { return g.generation != gen; })) {
                g.arrived -= 1;
                return fail(VG_E_STATE, "rank %d: a peer did not reach a barrier within %.1f s", cm->rank, cm->timeout_ns * 1e-9);
            }
        }
        for (int r = 0; r < cm->world; ++r)
            if (r != cm->rank) CU(cudaStreamWaitEvent(s, g.ev[(size_t)r * 2 + par], 0));
        return VG_OK;
    }
    CU(vg::launch_peer_barrier(peers_at(cm, 0), cm->world, cm->rank, cm->epoch, cm->timeout_ns, cm->d_timeout, s));
    cm->launches += 1;
    return VG_OK;
}

// After a stream synchronize: did any barrier give up on a peer?
int check_peers(vg_comm* cm) {
    if (cm->world == 1) return VG_OK;
    unsigned int t = 0;
    CU(cudaMemcpy(&t, cm->d_timeout, sizeof t, cudaMemcpyDeviceToHost));
    if (t) return fail(VG_E_STATE, "rank %d: a peer did not reach a barrier within %.1f s; results are incomplete", cm->rank,
                       cm->timeout_ns * 1e-9);
    return VG_OK;
}

// extract-to-symmetric-buffer done on every rank -> out = min(255, sum over ranks), on stream s.
int combine_from_peers(vg_comm* cm, size_t counts_off, uint64_t n, uint8_t* d_out, cudaStream_t s) {
    int rc = barrier_on(cm, s);  // every rank's vector is complete
    if (rc) return rc;
    CU(vg::launch_combine_counts(peers_at(cm, counts_off), cm->world, n, d_out, cm->ctx->nsm, s));
    cm->launches += 1;
    return barrier_on(cm, s);    // nobody overwrites its vector while a peer still reads it
}

}  // namespace

extern "C" {
static int allreduce_slots(vg_comm* cm, vg_index* ix, cudaStream_t s);


int vg_comm_create(vg_ctx* c, int rank, int world, uint64_t arena_bytes, vg_comm** out) {
    if (!c || !out) return fail(VG_E_INVALID, "vg_comm_create: NULL argument");
    *out = nullptr;
    if (world < 1 || world > vg::kMaxWorld || rank < 0 || rank >= world)
        return fail(VG_E_INVALID, "rank %d / world %d outside 0 <= rank < world <= %d", rank, world, vg::kMaxWorld);
    if (arena_bytes < (1u << 20)) arena_bytes = 1u << 20;
    DeviceGuard g(c->device);
    vg_comm* cm = new vg_comm();
    cm->ctx = c;
    cm->rank = rank;
    cm->world = world;
    cm->arena_bytes = (size_t)arena_bytes;
    if (const char* e = getenv("VG_BARRIER_TIMEOUT_MS")) {
        const unsigned long long ms = strtoull(e, nullptr, 10);
        if (ms) cm->timeout_ns = ms * 1000000ull;
    }
    cudaError_t e = cudaMalloc((void**)&cm->arena, cm->arena_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&cm->d_timeout, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(cm->arena, 0, kFlagsBytes);
    if (e == cudaSuccess) e = cudaMemset(cm->d_timeout, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(cm->arena);
        cudaFree(cm->d_timeout);
        delete cm;
        return fail(e == cudaErrorMemoryAllocation ? VG_E_NOMEM : VG_E_CUDA, "vg_comm_create (%llu-byte arena): %s",
                    (unsigned long long)arena_bytes, cudaGetErrorString(e));
    }
    cm->arena_used = kFlagsBytes;
    cm->peer_base[rank] = cm->arena;
    cm->connected = world == 1;
    *out = cm;
    return VG_OK;
}

// All ranks in one process (the C++ host driving several GPUs): the arenas are reached through ordinary peer access.
// Collective calls on such a group must still be made by every rank, CONCURRENTLY (one host thread per rank): they
// wait for each other on the device.
int vg_comm_create_local(vg_ctx* const* ctxs, int world, uint64_t arena_bytes, vg_comm** out) {
    if (!ctxs || !out) return fail(VG_E_INVALID, "vg_comm_create_local: NULL argument");
    if (world < 1 || world > vg::kMaxWorld) return fail(VG_E_INVALID, "world %d outside 1..%d", world, vg::kMaxWorld);
    for (int r = 0; r < world; ++r) {
        out[r] = nullptr;
        if (!ctxs[r]) return fail(VG_E_INVALID, "vg_comm_create_local: ctxs[%d] is NULL", r);
        for (int q = 0; q < r; ++q) {
            if (ctxs[q] == ctxs[r]) return fail(VG_E_INVALID, "vg_comm_create_local: one context listed twice");
            // two ranks on one device work (the tests of a 1-GPU box do that) but make no sense in production
            if (ctxs[q]->device == ctxs[r]->device && !getenv("VG_ALLOW_SAME_DEVICE"))
                return fail(VG_E_INVALID, "vg_comm_create_local: device %d listed twice", ctxs[r]->device);
        }
    }
    int rc = VG_OK;
    for (int r = 0; r < world && rc == VG_OK; ++r) rc = vg_comm_create(ctxs[r], r, world, arena_bytes, &out[r]);
    for (int r = 0; r < world && rc == VG_OK; ++r) {
        DeviceGuard g(ctxs[r]->device);
        for (int q = 0; q < world && rc == VG_OK; ++q) {
            if (q == r) continue;
            if (ctxs[r]->device != ctxs[q]->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device);
                if (!can) {
                    rc = fail(VG_E_CUDA, "GPU %d cannot reach GPU %d's memory", ctxs[r]->device, ctxs[q]->device);
                    break;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail(VG_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
                cudaGetLastError();
            }
            out[r]->peer_base[q] = out[q]->arena;
        }
        if (rc == VG_OK) {
            out[r]->connected = true;
            out[r]->local = true;
        }
    }
    if (rc == VG_OK && world > 1) {
        auto group = std::make_shared<LocalGroup>();
        group->world = world;
        group->ev.assign((size_t)world * 2, nullptr);
        for (int r = 0; r < world && rc == VG_OK; ++r) {
            DeviceGuard g(ctxs[r]->device);
            for (int p = 0; p < 2; ++p)
                if (cudaEventCreateWithFlags(&group->ev[(size_t)r * 2 + p], cudaEventDisableTiming) != cudaSuccess)
                    rc = fail(VG_E_CUDA, "vg_comm_create_local: cudaEventCreate failed");
            group->device.push_back(ctxs[r]->device);
        }
        for (int r = 0; r < world; ++r) out[r]->group = group;
    }
    if (rc != VG_OK)
        for (int r = 0; r < world; ++r) {
            if (out[r]) {
                for (int q = 0; q < world; ++q)
                    if (q != r) out[r]->peer_base[q] = nullptr;
                out[r]->local = true;  // nothing to unmap
                vg_comm_destroy(out[r]);
                out[r] = nullptr;
            }
        }
    return rc;
}

int vg_comm_handle(const vg_comm* cm, void* handle_out) {
    if (!cm || !handle_out) return fail(VG_E_INVALID, "vg_comm_handle: NULL argument");
    DeviceGuard g(cm->ctx->device);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, cm->arena));
    memcpy(handle_out, &h, sizeof h);
    return VG_OK;
}

int vg_comm_connect(vg_comm* cm, const void* handles) {
    if (!cm || !handles) return fail(VG_E_INVALID, "vg_comm_connect: NULL argument");
    if (cm->connected) return cm->world == 1 ? VG_OK : fail(VG_E_STATE, "vg_comm_connect: already connected");
    DeviceGuard g(cm->ctx->device);
    for (int r = 0; r < cm->world; ++r) {
        if (r == cm->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * VG_COMM_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (q != cm->rank && cm->peer_base[q]) {
                    cudaIpcCloseMemHandle(cm->peer_base[q]);
                    cm->peer_base[q] = nullptr;
                }
            return fail(VG_E_CUDA, "rank %d cannot map rank %d's arena (cudaIpcOpenMemHandle: %s)", cm->rank, r,
                        cudaGetErrorString(e));
        }
        cm->peer_base[r] = (uint8_t*)p;
    }
    cm->connected = true;
    return VG_OK;
}

int vg_comm_barrier(vg_comm* cm) {
    if (!cm || !cm->connected) return fail(VG_E_STATE, "vg_comm_barrier: not connected");
    DeviceGuard g(cm->ctx->device);
    return barrier_on(cm, cm->ctx->compute_stream);
}

int vg_comm_rank(const vg_comm* cm) { return cm ? cm->rank : -1; }
int vg_comm_world(const vg_comm* cm) { return cm ? cm->world : 0; }
uint64_t vg_comm_launches(const vg_comm* cm) { return cm ? cm->launches : 0; }

int vg_comm_destroy(vg_comm* cm) {
    if (!cm) return VG_OK;
    DeviceGuard g(cm->ctx->device);
    cudaDeviceSynchronize();
    if (cm->group) {  // this rank's two events go with it
        for (int p = 0; p < 2; ++p) {
            cudaEvent_t& e = cm->group->ev[(size_t)cm->rank * 2 + p];
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
    }
    if (!cm->local)
        for (int r = 0; r < cm->world; ++r)
            if (r != cm->rank && cm->peer_base[r]) cudaIpcCloseMemHandle(cm->peer_base[r]);
    cudaFree(cm->arena);
    cudaFree(cm->d_timeout);
    cudaFree(cm->d_reduced);
    cudaGetLastError();
    delete cm;
    return VG_OK;
}

// Replicated index, reads sharded over ranks: every rank extracts its u8 counts into the arena, then
// sums all ranks' vectors with a clamp at 255 -- exact, counts being saturating sums:
// min(255, sum_r min(255, c_r)) == min(255, total).  Collective; the result lands in dev_out (n bytes
// on the device, may be NULL) and / or c_out (host, may be NULL).
int vg_count_allreduce(vg_comm* cm, vg_index* ix, uint8_t* c_out, void* dev_out) {
    if (!cm || !ix) return fail(VG_E_INVALID, "vg_count_allreduce: NULL argument");
    if (!cm->connected) return fail(VG_E_STATE, "vg_count_allreduce: vg_comm_connect first");
    if (ix->sharded) return fail(VG_E_STATE, "vg_count_allreduce is for replicated indexes; a sharded one combines in vg_count_end");
    if (ix->ctx != cm->ctx) return fail(VG_E_INVALID, "index and comm live on different contexts");
    vg_ctx* c = cm->ctx;
    DeviceGuard g(c->device);
    const uint64_t n = ix->n;
    const size_t need = (size_t)((n + 15) & ~15ull) + 16;
    const bool replica = ix->replica_of == cm;
    if (!replica && cm->reduce_bytes < need) {  // every rank's key-order vector goes through the arena
        if (cm->reduce_bytes) return fail(VG_E_STATE, "vg_count_allreduce: one index size per comm (arena scratch is %zu bytes)", cm->reduce_bytes);
        if (!arena_alloc(cm, need, &cm->reduce_off)) return fail(VG_E_NOMEM, "arena too small for a %zu-byte count vector", need);
        cm->reduce_bytes = need;
    }
    if (!dev_out && cm->reduced_bytes < need) {
        cudaFree(cm->d_reduced);
        cm->d_reduced = nullptr;
        cm->reduced_bytes = 0;
        CU(cudaMalloc((void**)&cm->d_reduced, need));
        cm->reduced_bytes = need;
    }
    cudaStream_t s = c->compute_stream;
    int rc = vg_count_flush(ix);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->copy_stream));
    uint8_t* out = dev_out ? (uint8_t*)dev_out : cm->d_reduced;
    if (replica) {  // same slot order everywhere: reduce where the counts lie, then one gather into key order
        if (cm->world > 1 && (rc = allreduce_slots(cm, ix, s)) != VG_OK) return rc;
        CU(vg::counts_in_key_order(ix, out, 1, s));
        ix->launches += 1;
    } else {
        CU(vg::counts_in_key_order(ix, cm->arena + cm->reduce_off, 1, s));
        ix->launches += 1;
        rc = combine_from_peers(cm, cm->reduce_off, n, out, s);
        if (rc) return rc;
    }
    if (c_out) {
        CU(cudaMemcpyAsync(c_out, out, n, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return check_peers(cm);
    }
    return VG_OK;
}

// ---------------------------------------------------------------------------
// replica group: one rank builds the index, the others receive it over NVLink
// ---------------------------------------------------------------------------
namespace {
struct ReplicaMail {  // what the root tells the others, through its arena
    uint64_t magic, n, m_slots, duplicates, cap, cap2, round_keys, slack;
    uint32_t k, nbuckets, P, shift, shift2, sub_bits, filter_nwords, filter_span, may_grow, pad;
    cudaIpcMemHandle_t h_slots, h_rank_base, h_perm, h_filter;
    const void *p_slots, *p_rank_base, *p_perm, *p_filter;  // the same buffers as plain pointers (ranks of one process)
};
constexpr uint64_t kMailMagic = 0x76676232303072ULL;

// one buffer of the root's index -> this rank's copy of it (peer-to-peer over NVLink through a CUDA IPC mapping)
int pull_buffer(const cudaIpcMemHandle_t& h, const void* local_src, void* dst, size_t bytes, cudaStream_t s) {
    void* src = const_cast<void*>(local_src);
    if (!local_src) CU(cudaIpcOpenMemHandle(&src, h, cudaIpcMemLazyEnablePeerAccess));
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (!local_src) cudaIpcCloseMemHandle(src);
    if (e != cudaSuccess) return fail(VG_E_CUDA, "replica copy of %zu bytes: %s", bytes, cudaGetErrorString(e));
    return VG_OK;
}
}  // namespace

int vg_index_replicate(vg_comm* cm, int root, vg_index* root_ix, vg_index** out) {
    if (!cm || !out) return fail(VG_E_INVALID, "vg_index_replicate: NULL argument");
    *out = nullptr;
    if (!cm->connected) return fail(VG_E_STATE, "vg_index_replicate: vg_comm_connect first");
    if (root < 0 || root >= cm->world) return fail(VG_E_INVALID, "root %d outside [0,%d)", root, cm->world);
    const bool is_root = cm->rank == root;
    if (is_root && (!root_ix || root_ix->ctx != cm->ctx || root_ix->sharded || !root_ix->part.enabled || root_ix->replica_of))
        return fail(VG_E_INVALID, "vg_index_replicate: the root passes a partitioned, unsharded index of the comm's context");
    vg_ctx* c = cm->ctx;
    DeviceGuard g(c->device);
    cudaStream_t s = c->compute_stream;
    size_t mail_off = 0;
    uint8_t* d_mail = (uint8_t*)arena_alloc(cm, sizeof(ReplicaMail), &mail_off);
    if (!d_mail) return fail(VG_E_NOMEM, "arena too small for the replica mailbox");
    ReplicaMail mail{};
    if (is_root) {
        const PartState& ps = root_ix->part;
        mail.magic = kMailMagic;
        mail.n = root_ix->n;
        mail.m_slots = root_ix->m_slots;
        mail.duplicates = root_ix->duplicates;
        mail.cap = ps.view.cap;
        mail.cap2 = ps.view.cap2;
        mail.round_keys = ps.round_keys;
        mail.slack = ps.slack;
        mail.k = root_ix->view.k;
        mail.nbuckets = root_ix->view.nbuckets;
        mail.P = ps.view.P;
        mail.shift = ps.view.shift;
        mail.shift2 = ps.view.shift2;
        mail.sub_bits = ps.view.sub_bits;
        mail.filter_nwords = ps.filter.words ? ps.filter.nwords : 0;
        mail.filter_span = ps.filter.span;
        mail.may_grow = ps.may_grow ? 1 : 0;
        if (cm->local) {
            mail.p_slots = root_ix->view.slots;
            mail.p_rank_base = root_ix->view.rank_base;
            mail.p_perm = root_ix->d_perm;
            mail.p_filter = ps.d_filter;
        } else {
            CU(cudaIpcGetMemHandle(&mail.h_slots, root_ix->view.slots));
            CU(cudaIpcGetMemHandle(&mail.h_rank_base, root_ix->view.rank_base));
            CU(cudaIpcGetMemHandle(&mail.h_perm, root_ix->d_perm));
            if (mail.filter_nwords) CU(cudaIpcGetMemHandle(&mail.h_filter, ps.d_filter));
        }
        CU(cudaMemcpyAsync(d_mail, &mail, sizeof mail, cudaMemcpyHostToDevice, s));
    }
    int rc = barrier_on(cm, s);  // the mail is in the root's arena
    if (rc) return rc;
    if (!is_root) {
        CU(cudaMemcpyAsync(&mail, cm->peer_base[root] + mail_off, sizeof mail, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        rc = check_peers(cm);
        if (rc) return rc;
        if (mail.magic != kMailMagic) return fail(VG_E_STATE, "rank %d: no replica mail from rank %d", cm->rank, root);
    }
    // the count vectors of a replica group live in the arena, at the same offset on every rank: the reduce reads them
    const size_t cvec_bytes = (size_t)((mail.m_slots + 31) & ~15ull);
    size_t cvec_off = 0;
    uint8_t* cvec = (uint8_t*)arena_alloc(cm, cvec_bytes, &cvec_off);
    if (!cvec) return fail(VG_E_NOMEM, "arena of %zu bytes too small for a %zu-byte count vector", cm->arena_bytes, cvec_bytes);
    CU(cudaMemsetAsync(cvec, 0, cvec_bytes, s));
    vg_index* ix = root_ix;
    if (is_root) {
        CU(cudaStreamSynchronize(s));
        cudaFree(ix->view.cvec);
    } else {
        ix = new vg_index();
        ix->ctx = c;
        ix->n = mail.n;
        ix->m_slots = mail.m_slots;
        ix->duplicates = mail.duplicates;
        ix->view.k = mail.k;
        ix->view.mask = (1ULL << (2 * mail.k)) - 1;
        ix->view.nbuckets = ix->view.nb_total = mail.nbuckets;
        ix->view.b_base = 0;
        PartState& ps = ix->part;
        ps.view.P = ps.view.P_local = mail.P;
        ps.view.shift = mail.shift;
        ps.view.shift2 = mail.shift2;
        ps.view.sub_bits = mail.sub_bits;
        ps.view.cap = mail.cap;
        ps.view.cap2 = mail.cap2;
        ps.view.world = 1;
        ps.view.rank = 0;
        ps.round_keys = mail.round_keys;
        ps.slack = mail.slack;
        ps.may_grow = mail.may_grow != 0;
        auto bail = [&](int code) {
            ix->view.cvec = nullptr;
            vg_index_destroy(ix);
            return code;
        };
        cudaError_t e = cudaMalloc((void**)&ix->view.slots, (size_t)mail.nbuckets * 32);
        if (e == cudaSuccess) e = cudaMalloc((void**)&ix->view.rank_base, (size_t)mail.nbuckets * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&ix->d_perm, std::max<uint64_t>(mail.n, 1) * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&ix->d_counts, std::max<uint64_t>(mail.n, 4));
        if (e == cudaSuccess) e = cudaMalloc((void**)&ix->d_misc, sizeof(vg::DeviceMisc));
        if (e == cudaSuccess) e = cudaMemsetAsync(ix->d_misc, 0, sizeof(vg::DeviceMisc), s);
        if (e == cudaSuccess && mail.filter_nwords) e = cudaMalloc((void**)&ps.d_filter, (size_t)mail.filter_nwords * 4);
        if (e == cudaSuccess) e = vg::part_alloc_lists(ix);
        if (e != cudaSuccess)
            return bail(fail(e == cudaErrorMemoryAllocation ? VG_E_NOMEM : VG_E_CUDA, "replica of %llu k-mers: %s",
                             (unsigned long long)mail.n, cudaGetErrorString(e)));
        if ((rc = pull_buffer(mail.h_slots, cm->local ? mail.p_slots : nullptr, ix->view.slots, (size_t)mail.nbuckets * 32, s)) != VG_OK) return bail(rc);
        if ((rc = pull_buffer(mail.h_rank_base, cm->local ? mail.p_rank_base : nullptr, ix->view.rank_base, (size_t)mail.nbuckets * sizeof(uint32_t), s)) != VG_OK) return bail(rc);
        if ((rc = pull_buffer(mail.h_perm, cm->local ? mail.p_perm : nullptr, ix->d_perm, (size_t)mail.n * sizeof(uint32_t), s)) != VG_OK) return bail(rc);
        if (mail.filter_nwords) {
            if ((rc = pull_buffer(mail.h_filter, cm->local ? mail.p_filter : nullptr, ps.d_filter, (size_t)mail.filter_nwords * 4, s)) != VG_OK) return bail(rc);
            ps.filter.words = ps.d_filter;
            ps.filter.nwords = mail.filter_nwords;
            ps.filter.span = mail.filter_span;
            if ((size_t)mail.filter_nwords * 4 <= (64ull << 20)) vg::pin_in_l2(c, ps.d_filter, (size_t)mail.filter_nwords * 4);
        }
        ps.enabled = true;
        ix->view.cvec = cvec;  // fetch_slice_ranks reads rank_base only
        if ((rc = vg::fetch_slice_ranks(ix)) != VG_OK) return bail(rc);
    }
    ix->view.cvec = cvec;
    ix->cvec_off = cvec_off;
    ix->replica_of = cm;
    rc = barrier_on(cm, s);  // the root keeps its buffers untouched until everybody has its copy
    if (rc == VG_OK) {
        CU(cudaStreamSynchronize(s));
        rc = check_peers(cm);
    }
    if (rc) {
        if (!is_root) {
            ix->view.cvec = nullptr;
            vg_index_destroy(ix);
        }
        return rc;
    }
    *out = ix;
    return VG_OK;
}

// Slot-order count reduce of a replica group (COLLECTIVE): reduce-scatter + all-gather over peer memory, in place in
// the ranks' count vectors; afterwards every rank's vector holds min(255, sum over ranks).
static int allreduce_slots(vg_comm* cm, vg_index* ix, cudaStream_t s) {
    const uint64_t nbytes = (ix->m_slots + 15) & ~15ull;
    const uint64_t seg = ((nbytes / cm->world + 15) & ~15ull) ? ((nbytes / cm->world + 15) & ~15ull) : 16;
    const vg::PeerPtrs vecs = peers_at(cm, ix->cvec_off);
    int rc = barrier_on(cm, s);  // every rank has finished counting
    if (rc) return rc;
    CU(vg::launch_reduce_segment(vecs, cm->world, cm->rank, seg, nbytes, cm->ctx->nsm, s));
    rc = barrier_on(cm, s);      // every segment is reduced
    if (rc) return rc;
    CU(vg::launch_gather_segments(vecs, cm->world, cm->rank, seg, nbytes, cm->ctx->nsm, s));
    cm->launches += 2;
    return barrier_on(cm, s);    // nobody zeroes its vector for the next sample while a peer still reads it
}

int vg_count_allreduce_slots(vg_comm* cm, vg_index* ix, uint8_t* c_slots_out, const uint8_t** dev_counts) {
    if (!cm || !ix) return fail(VG_E_INVALID, "vg_count_allreduce_slots: NULL argument");
    if (ix->replica_of != cm) return fail(VG_E_STATE, "vg_count_allreduce_slots: the index is not a replica of this comm's group (vg_index_replicate)");
    vg_ctx* c = cm->ctx;
    DeviceGuard g(c->device);
    cudaStream_t s = c->compute_stream;
    int rc = vg_count_flush(ix);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->copy_stream));
    if (cm->world > 1 && (rc = allreduce_slots(cm, ix, s)) != VG_OK) return rc;
    if (dev_counts) *dev_counts = ix->view.cvec;
    if (c_slots_out) {
        CU(cudaMemcpyAsync(c_slots_out, ix->view.cvec, ix->m_slots, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return check_peers(cm);
    }
    return VG_OK;
}

int vg_comm_check(vg_comm* cm) {
    if (!cm) return fail(VG_E_INVALID, "comm is NULL");
    DeviceGuard g(cm->ctx->device);
    CU(cudaStreamSynchronize(cm->ctx->compute_stream));
    return check_peers(cm);
}

// ---------------------------------------------------------------------------
// sharded index
// ---------------------------------------------------------------------------
static int index_create_sharded(vg_comm* cm, const uint64_t* keys, bool keys_on_device, uint64_t n, uint32_t k, double load_factor,
                                uint64_t round_bytes, vg_index** out);
int vg_index_create_sharded(vg_comm* cm, const uint64_t* keys, uint64_t n, uint32_t k, double load_factor,
                            uint64_t round_bytes, vg_index** out) {
    return index_create_sharded(cm, keys, false, n, k, load_factor, round_bytes, out);
}
int vg_index_create_sharded_device(vg_comm* cm, const uint64_t* dev_keys, uint64_t n, uint32_t k, double load_factor,
                                   uint64_t round_bytes, vg_index** out) {
    return index_create_sharded(cm, dev_keys, true, n, k, load_factor, round_bytes, out);
}
// keys: host memory, staged through a bounded buffer twice -- or (keys_on_device) memory of this rank's GPU, read in place
static int index_create_sharded(vg_comm* cm, const uint64_t* keys, bool keys_on_device, uint64_t n, uint32_t k, double load_factor,
                                uint64_t round_bytes, vg_index** out) {
    if (!cm || !out || (!keys && n)) return fail(VG_E_INVALID, "vg_index_create_sharded: NULL argument");
    *out = nullptr;
    if (!cm->connected) return fail(VG_E_STATE, "vg_index_create_sharded: vg_comm_connect first");
    if (k < 1 || k > 28) return fail(VG_E_INVALID, "k=%u outside 1..28 (reference asserts k<=28, src/kmer.cpp:124)", k);
    if (load_factor <= 0) load_factor = 0.3;
    if (load_factor > 0.9) return fail(VG_E_INVALID, "load_factor %.3f > 0.9", load_factor);
    vg_ctx* c = cm->ctx;
    DeviceGuard g(c->device);
    const uint32_t W = (uint32_t)cm->world;

    // ---- geometry: the same on every rank (depends on n, world and the tuning variables only) ----
    uint64_t nb_min = (uint64_t)((double)n / W / (4.0 * load_factor)) + 1;
    if (nb_min < 64) nb_min = 64;
    vg::PartGeometry geo;  // a sharded table is sized by its slices; small ones shrink them until every GPU has three
    if (!vg::part_geometry(nb_min, W, geo))
        return fail(VG_E_INVALID, "index of %llu keys cannot be partitioned over %u GPUs", (unsigned long long)n, W);
    const uint32_t shift2 = geo.shift2, sub_bits = geo.sub_bits, shift = shift2 + sub_bits;
    const uint64_t P_local = geo.P_local;
    const uint64_t nb_local = P_local << shift, nb_total = nb_local * W;
    if (nb_total >= 0xffffffffull) return fail(VG_E_INVALID, "index of %llu keys needs too many buckets", (unsigned long long)n);
    if (round_bytes == 0) round_bytes = 256ull << 20;
    round_bytes = (round_bytes + 4095) & ~4095ull;
    uint64_t slack = 65536;
    if (const char* e = getenv("VG_PART_SLACK")) slack = strtoull(e, nullptr, 10);
    const uint64_t P = P_local * W;
    const uint64_t cap = (round_bytes / P) * 5 / 4 + slack;
    const uint64_t cap2 = sub_bits ? ((cap * W) >> sub_bits) * 5 / 4 + slack : 0;  // a coarse partition receives from W sources

    vg_index* ix = new vg_index();
    ix->ctx = c;
    ix->comm = cm;
    ix->sharded = true;
    ix->n = n;
    ix->view.k = k;
    ix->view.mask = (1ULL << (2 * k)) - 1;
    ix->view.nbuckets = (uint32_t)nb_local;
    ix->view.nb_total = (uint32_t)nb_total;
    ix->view.b_base = (uint32_t)(nb_local * cm->rank);
    auto bail = [&](int code) {
        vg_index_destroy(ix);
        return code;
    };
#define CUB(expr)                                                                                            \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return bail(fail(e__ == cudaErrorMemoryAllocation ? VG_E_NOMEM : VG_E_CUDA, "%s: %s", #expr,     \
                             cudaGetErrorString(e__)));                                                      \
    } while (0)
    // ---- symmetric objects (same order, same sizes on every rank) ----
    size_t slots_off = 0, keybuf_off = 0, incount_off = 0, rank_off = 0, cvec_off = 0;
    const size_t counts_bytes = (size_t)((n + 15) & ~15ull) + 16;
    // this rank's count vector (slot order) and rank_base: peers reach them for the rare direct probe of a key
    // whose list is full, so they live in the arena too; no rank owns more keys than min(n, slots of its table)
    const size_t cvec_bytes = (size_t)((std::min<uint64_t>(n, 4 * nb_local) + 31) & ~15ull);
    PartState& ps = ix->part;
    ix->view.slots = (uint64_t*)arena_alloc(cm, nb_local * 32, &slots_off);
    ix->d_counts = (uint8_t*)arena_alloc(cm, counts_bytes, &ix->counts_off);
    ps.view.keybuf = (uint64_t*)arena_alloc(cm, P * cap * sizeof(uint64_t), &keybuf_off);
    ps.view.incount = (unsigned long long*)arena_alloc(cm, P * sizeof(unsigned long long), &incount_off);
    ix->view.rank_base = (uint32_t*)arena_alloc(cm, nb_local * sizeof(uint32_t), &rank_off);
    ix->view.cvec = (uint8_t*)arena_alloc(cm, cvec_bytes, &cvec_off);
    if (!ix->view.slots || !ix->d_counts || !ps.view.keybuf || !ps.view.incount || !ix->view.rank_base || !ix->view.cvec) {
        ix->view.slots = nullptr;
        return bail(fail(VG_E_NOMEM, "arena of %zu bytes too small: table %llu + counts %zu + key lists %llu bytes per rank",
                         cm->arena_bytes, (unsigned long long)(nb_local * 32), counts_bytes, (unsigned long long)(P * cap * 8)));
    }
    ps.view.P = (uint32_t)P;
    ps.view.shift = shift;
    ps.view.shift2 = shift2;
    ps.view.sub_bits = sub_bits;
    ps.view.cap = cap;
    ps.view.cap2 = cap2;
    ps.view.world = W;
    ps.view.rank = (uint32_t)cm->rank;
    ps.view.P_local = (uint32_t)P_local;
    for (uint32_t r = 0; r < W; ++r) {
        ps.view.peer_keybuf[r] = (uint64_t*)(cm->peer_base[r] + keybuf_off);
        ps.view.peer_slots[r] = (uint64_t*)(cm->peer_base[r] + slots_off);
        ps.view.peer_incount[r] = (unsigned long long*)(cm->peer_base[r] + incount_off);
        ps.view.peer_rank_base[r] = (uint32_t*)(cm->peer_base[r] + rank_off);
        ps.view.peer_cvec[r] = (uint8_t*)(cm->peer_base[r] + cvec_off);
    }
    ps.round_keys = round_bytes;
    cudaStream_t s = c->compute_stream;
    CUB(cudaMalloc((void**)&ps.view.cursor, P * sizeof(unsigned long long)));
    CUB(cudaMalloc((void**)&ps.view.ctr, ((size_t)8 << shift2) * sizeof(uint32_t)));
    if (sub_bits) {
        CUB(cudaMalloc((void**)&ps.view.keybuf2, (cap2 << sub_bits) * sizeof(uint64_t)));
        CUB(cudaMalloc((void**)&ps.view.cursor2, sizeof(unsigned long long) << sub_bits));
    }
    CUB(cudaMalloc((void**)&ix->d_combined, counts_bytes));
    CUB(cudaMalloc((void**)&ix->d_misc, sizeof(vg::DeviceMisc)));
    CUB(cudaMemsetAsync(ix->d_misc, 0, sizeof(vg::DeviceMisc), s));
    CUB(cudaMemsetAsync(ps.view.ctr, 0, ((size_t)8 << shift2) * sizeof(uint32_t), s));
    CUB(cudaMemsetAsync(ps.view.cursor, 0, P * sizeof(unsigned long long), s));
    CUB(cudaMemsetAsync(ps.view.incount, 0, P * sizeof(unsigned long long), s));
    CUB(cudaMemsetAsync(ix->view.cvec, 0, cvec_bytes, s));
    CUB(vg::launch_table_fill_empty(ix->view.slots, 4ull * nb_local, s));

    // ---- presence pre-filter over ALL keys (every rank filters its own reads before the exchange) ----
    uint32_t nwords = 0, fspan = 4;
    {
        uint64_t bytes = 0;
        vg::prefilter_plan(n, k, bytes, fspan);
        if (bytes >= 64) {
            nwords = (uint32_t)std::min<uint64_t>(bytes / 4, 0x7fffffffull);
            CUB(cudaMalloc((void**)&ps.d_filter, (size_t)nwords * 4));
            CUB(cudaMemsetAsync(ps.d_filter, 0, (size_t)nwords * 4, s));
        }
    }

    // ---- keys: pass 0 counts this rank's own keys, pass 1 keeps them (with their caller positions) ----
    uint64_t piece = 1ull << 22;
    if (const char* e = getenv("VG_SHARD_PIECE")) piece = std::max<uint64_t>(64, strtoull(e, nullptr, 10));  // (tests: many pieces)
    std::vector<uint64_t> tmp(keys_on_device ? 0 : (size_t)std::min<uint64_t>(piece, std::max<uint64_t>(n, 1)));
    uint64_t* d_piece = nullptr;
    unsigned long long* d_n_own = nullptr;
    unsigned long long* d_bad = nullptr;
    CUB(cudaMalloc((void**)&d_piece, (size_t)std::min<uint64_t>(piece, std::max<uint64_t>(n, 1)) * sizeof(uint64_t)));
    auto bail2 = [&](int code) {
        cudaFree(d_piece);
        cudaFree(d_n_own);
        cudaFree(d_bad);
        return bail(code);
    };
#define CUB2(expr)                                                                                           \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return bail2(fail(e__ == cudaErrorMemoryAllocation ? VG_E_NOMEM : VG_E_CUDA, "%s: %s", #expr,    \
                              cudaGetErrorString(e__)));                                                     \
    } while (0)
    CUB2(cudaMalloc((void**)&d_n_own, sizeof(unsigned long long)));
    if (keys_on_device) {
        unsigned long long none = ~0ull;
        CUB2(cudaMalloc((void**)&d_bad, sizeof(unsigned long long)));
        CUB2(cudaMemcpyAsync(d_bad, &none, sizeof none, cudaMemcpyHostToDevice, s));
    }
    for (int pass = 0; pass < 2; ++pass) {
        CUB2(cudaMemsetAsync(d_n_own, 0, sizeof(unsigned long long), s));
        for (uint64_t off = 0; off < n; off += piece) {
            const uint64_t m = std::min<uint64_t>(piece, n - off);
            if (keys_on_device) {  // validated and un-hashed on the device, piece by piece
                CUB2(vg::launch_keys_to_key56(keys + off, m, k, ix->view.mask, d_piece, d_bad, s));
                CUB2(vg::launch_select_owned(ix->view, d_piece, m, off, pass ? ix->d_key56 : nullptr, pass ? ix->d_idx : nullptr,
                                             d_n_own, s));
                if (pass == 1 && nwords) CUB2(vg::launch_prefilter_build(ps.d_filter, nwords, d_piece, m, k, fspan, s));
                continue;
            }
            for (uint64_t i = 0; i < m; ++i) {
                const uint64_t key = keys[off + i];
                if (pass == 0) {
                    if ((key & 0xffu) != k)
                        return bail2(fail(VG_E_INVALID, "keys[%llu]=0x%llx: low byte is not k=%u (src/kmer.cpp:138)",
                                          (unsigned long long)(off + i), (unsigned long long)key, k));
                    if ((key >> 8) > ix->view.mask)
                        return bail2(fail(VG_E_INVALID, "keys[%llu]: hash exceeds 2k bits", (unsigned long long)(off + i)));
                }
                tmp[(size_t)i] = key >> 8;
            }
            CUB2(cudaMemcpyAsync(d_piece, tmp.data(), m * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
            CUB2(vg::launch_unhash(d_piece, m, ix->view.mask, s));
            CUB2(vg::launch_select_owned(ix->view, d_piece, m, off, pass ? ix->d_key56 : nullptr, pass ? ix->d_idx : nullptr,
                                         d_n_own, s));
            if (pass == 1 && nwords) CUB2(vg::launch_prefilter_build(ps.d_filter, nwords, d_piece, m, k, fspan, s));
            CUB2(cudaStreamSynchronize(s));  // tmp is reused
        }
        if (pass == 0) {
            unsigned long long no = 0;
            CUB2(cudaStreamSynchronize(s));
            if (keys_on_device) {
                unsigned long long bad = ~0ull;
                CUB2(cudaMemcpy(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost));
                if (bad != ~0ull)
                    return bail2(fail(VG_E_INVALID, "keys[%llu]: low byte is not k=%u or the hash exceeds 2k bits (src/kmer.cpp:138)", bad, k));
            }
            CUB2(cudaMemcpy(&no, d_n_own, sizeof no, cudaMemcpyDeviceToHost));
            ix->n_own = no;
            CUB2(cudaMalloc((void**)&ix->d_key56, std::max<uint64_t>(no, 1) * sizeof(uint64_t)));
            CUB2(cudaMalloc((void**)&ix->d_idx, std::max<uint64_t>(no, 1) * sizeof(uint64_t)));
        }
    }
    CUB2(cudaStreamSynchronize(s));
    cudaFree(d_piece);
    cudaFree(d_n_own);
    cudaFree(d_bad);
    d_piece = nullptr;
    d_n_own = nullptr;
    d_bad = nullptr;
    if (ix->n_own > 3.6 * nb_local)
        return bail(fail(VG_E_NOMEM, "rank %d owns %llu keys for %llu slots", cm->rank, (unsigned long long)ix->n_own,
                         (unsigned long long)(4 * nb_local)));
    CUB(vg::launch_insert(ix->view, ix->d_key56, ix->n_own, &ix->d_misc->report, s));
    vg::DeviceMisc misc;
    CUB(cudaMemcpyAsync(&misc, ix->d_misc, sizeof misc, cudaMemcpyDeviceToHost, s));
    CUB(cudaStreamSynchronize(s));
    if (misc.report.failed) return bail(fail(VG_E_NOMEM, "index build: %llu keys found no slot", misc.report.failed));
    ix->duplicates = misc.report.duplicates;
    {   // slot order of this rank's table; from here on its own keys are only needed as positions in it
        const uint32_t nblocks = (uint32_t)((nb_local + 1023ull) / 1024);
        uint32_t* d_sums = nullptr;
        unsigned long long* d_total = nullptr;
        unsigned long long total = 0;
        CUB(cudaMalloc((void**)&d_sums, ((size_t)nblocks + 1) * sizeof(uint32_t)));
        cudaError_t e = cudaMalloc((void**)&d_total, sizeof(unsigned long long));
        if (e == cudaSuccess) e = vg::launch_rank_scan(ix->view, d_sums, d_total, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMalloc((void**)&ix->d_perm, std::max<uint64_t>(ix->n_own, 1) * sizeof(uint32_t));
        if (e == cudaSuccess) e = vg::launch_slot_perm(ix->view, ix->d_key56, ix->n_own, ix->d_perm, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(d_sums);
        cudaFree(d_total);
        CUB(e);
        ix->m_slots = total;
        cudaFree(ix->d_key56);
        ix->d_key56 = nullptr;
        int frc = vg::fetch_slice_ranks(ix);
        if (frc) return bail(frc);
    }
    if (nwords) {
        ps.filter.words = ps.d_filter;
        ps.filter.nwords = nwords;
        ps.filter.span = fspan;  // the scatter must ask with the word length the filter was built with
        if ((size_t)nwords * 4 <= (64ull << 20)) vg::pin_in_l2(c, ps.d_filter, (size_t)nwords * 4);
    }
    ps.enabled = true;
#undef CUB
#undef CUB2
    // nobody scatters into (or probes) a peer's table before that peer has built it
    int rc = barrier_on(cm, s);
    if (rc) return bail(rc);
    *out = ix;
    return VG_OK;
}

uint64_t vg_index_own_keys(const vg_index* ix) { return ix ? (ix->sharded ? ix->n_own : ix->n) : 0; }

}  // extern "C"

// End of a round on a sharded index (collective): tell the owners how many keys they received, wait
// for every rank's scatter, sweep this GPU's slices, wait again so that the next round's scatter
// cannot overwrite lists a peer is still probing.
int vg::sharded_flush(vg_index* ix, cudaStream_t s) {
    vg_comm* cm = ix->comm;
    PartState& ps = ix->part;
    if (cm->world > 1) {
        CU(vg::launch_publish_counts(ps.view, s));
        int rc = barrier_on(cm, s);
        if (rc) return rc;
        CU(vg::launch_probe_partitions(ix->view, ps.view, ps.slice_rank.empty() ? nullptr : ps.slice_rank.data(), &ix->d_misc->stats,
                                       ix->ctx->nsm, s));
        ix->launches += vg::sweep_launches(ix->view, ps.view) + 1;
        rc = barrier_on(cm, s);
        if (rc) return rc;
    } else {  // a group of one: the key lists and cursors are local
        CU(vg::launch_probe_partitions(ix->view, ps.view, ps.slice_rank.empty() ? nullptr : ps.slice_rank.data(), &ix->d_misc->stats,
                                       ix->ctx->nsm, s));
        ix->launches += vg::sweep_launches(ix->view, ps.view);
    }
    ps.pending = 0;
    return VG_OK;
}

// After the last round: counts of this GPU's own keys at their caller positions (zero elsewhere), then
// the sum over ranks -- every key is owned by exactly one rank -- gives every rank all n counts.
int vg::sharded_end(vg_index* ix, uint8_t* c_out) {
    vg_comm* cm = ix->comm;
    vg_ctx* c = ix->ctx;
    cudaStream_t s = c->compute_stream;
    const uint64_t n = ix->n;
    CU(cudaMemsetAsync(ix->d_counts, 0, (size_t)((n + 15) & ~15ull), s));
    CU(vg::launch_gather_counts(ix->view.cvec, ix->d_perm, ix->d_idx, ix->n_own, ix->d_counts, 1, s));
    ix->launches += 1;
    int rc = combine_from_peers(cm, ix->counts_off, n, ix->d_combined, s);
    if (rc) return rc;
    if (c_out && n) CU(cudaMemcpyAsync(c_out, ix->d_combined, n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return check_peers(cm);
}
