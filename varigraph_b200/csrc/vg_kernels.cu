// vg_kernels.cu -- the CUDA kernels of the read k-mer counting path, sm_100a only.
//
//   K1  rolling canonical k-mer encoder      (vg_device.cuh: encode_keys_odd / encode_keys_any)
//   K2  one-sector open-addressing index probe (probe_and_count below)
//   K3  saturating u8 counter accumulation with warp-aggregated CAS
//   K4  counting-Bloom-filter fill for `construct` (cbf_add_kernel)
//   K5  batched CBF count/find (cbf_query_kernel)
// K1-K3 are fused in count_kernel: nothing but the final counters ever leaves the SM.
//
// Replaces (does not port) src/kmer.cu:39-69, src/fastq_kmer.cu:99-162 (sort + reduce_by_key +
// host map probes) and src/counting_bloom_filter.cu:5-104 of the reference; results follow the
// reference CPU path src/kmer.cpp:110-149 + src/fastq_kmer.cpp:126-141.
#include <algorithm>
#include <cstdlib>

#include "vg_device.cuh"
#include "vg_internal.h"

namespace vg {

// exclusive prefix sum over the CTA's threads (blockDim.x == kCtaThreads); total in `total`
__device__ __forceinline__ uint32_t cta_exclusive_sum(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int i = 0; i < kCtaThreads / 32; ++i) {
        const uint32_t ws = warp_sums[i];
        if ((uint32_t)i < warp) before += ws;
        all += ws;
    }
    total = all;
    __syncthreads();
    return before + inc - v;
}

// ---------------------------------------------------------------------------
// index build
// ---------------------------------------------------------------------------
__global__ void fill_empty_kernel(uint64_t* slots, uint64_t nslots) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nslots; i += stride) slots[i] = kSlotEmpty;
}

__global__ void unhash_kernel(uint64_t* key56, uint64_t n, uint64_t mask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key56[i] = hash64_inv(key56[i], mask);
}

// keys as the caller holds them (hash64(canonical k-mer) << 8 | k, src/kmer.cpp:138) -> canonical k-mers; *bad = the
// lowest position whose key is malformed (low byte != k, or a hash beyond 2k bits), ~0 if none
__global__ void keys_to_key56_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t k, uint64_t mask, uint64_t* key56,
                                     unsigned long long* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i], h = key >> 8;
    if ((key & 0xffu) != k || h > mask) {
        atomicMin(bad, (unsigned long long)i);
        key56[i] = kKey56Max;
        return;
    }
    key56[i] = hash64_inv(h, mask);
}

// One thread per key.  Claims the first empty slot in probe order with a 64-bit CAS; because
// slots are never freed, "an empty slot ends the search" holds for every later lookup.
__global__ void insert_kernel(IndexView ix, const uint64_t* __restrict__ key56, uint64_t n, InsertReport* rep) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = key56[i];
    if (key == kKey56Max) return;  // not a canonical k-mer (min(fwd, rev) is never all ones): cannot be hit
    uint64_t want = key << 8;
    uint32_t b = bucket_of(key, ix.nb_total) - ix.b_base;  // a sharded build only passes this table's own keys
    for (uint32_t tries = 0; tries < ix.nbuckets; ++tries) {
        uint64_t* base = ix.slots + 4ull * b;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            uint64_t cur = base[s];
            if (cur == kSlotEmpty) cur = atomicCAS((unsigned long long*)(base + s), kSlotEmpty, want);
            if (cur == kSlotEmpty) return;  // claimed
            if ((cur >> 8) == key) {
                atomicAdd(&rep->duplicates, 1ull);
                return;
            }
        }
        b = (b + 1 == ix.nbuckets) ? 0 : b + 1;
    }
    atomicAdd(&rep->failed, 1ull);
}

// Sharded build: keep the keys whose home bucket lies in this GPU's run of the global table.
__global__ void select_owned_kernel(IndexView ix, const uint64_t* __restrict__ key56, uint64_t n, uint64_t first_idx,
                                    uint64_t* own, uint64_t* own_idx, unsigned long long* n_own) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = key56[i];
    if (key == kKey56Max) return;  // in nobody's table: its count is 0 everywhere
    if (bucket_of(key, ix.nb_total) - ix.b_base >= ix.nbuckets) return;
    const unsigned long long at = atomicAdd(n_own, 1ull);
    if (own) {
        own[at] = key;
        own_idx[at] = first_idx + i;
    }
}

__global__ void clear_counts_kernel(uint64_t* slots, uint64_t nslots) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nslots; i += stride) {
        uint64_t v = slots[i];
        if (v != kSlotEmpty && (v & 0xffu)) slots[i] = v & ~0xffULL;
    }
}

__device__ __forceinline__ int match_slot(const uint64_t (&v)[4], uint64_t key, bool& saw_empty, uint64_t& seen);

// ---- slot order (partitioned path): rank_base, the key -> slot permutation, gathers through it --------------
constexpr int kScanChunk = kCtaThreads * 4;  // buckets per CTA of the occupancy scan
__device__ __forceinline__ uint32_t bucket_occupancy(const uint64_t* slots, uint64_t b) {
    uint64_t v[4];
    ld_bucket(slots + 4 * b, v);
    return (v[0] != kSlotEmpty) + (v[1] != kSlotEmpty) + (v[2] != kSlotEmpty) + (v[3] != kSlotEmpty);
}
__global__ void __launch_bounds__(kCtaThreads) occ_block_sums_kernel(IndexView ix, uint32_t* block_sums) {
    __shared__ uint32_t ws[kCtaThreads / 32];
    const uint64_t b0 = (uint64_t)blockIdx.x * kScanChunk + threadIdx.x * 4;
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (b0 + j < ix.nbuckets) v += bucket_occupancy(ix.slots, b0 + j);
    uint32_t total;
    cta_exclusive_sum(v, ws, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
// one CTA: exclusive scan of the block sums in place; the grand total goes to *total
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(uint32_t* block_sums, uint32_t nblocks, unsigned long long* total) {
    __shared__ unsigned long long part[1024];
    const uint32_t per = (nblocks + 1023) / 1024, t0 = threadIdx.x * per;
    unsigned long long sum = 0;
    for (uint32_t i = 0; i < per && t0 + i < nblocks; ++i) sum += block_sums[t0 + i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const unsigned long long v = part[i];
            part[i] = run;
            run += v;
        }
        *total = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = 0; i < per && t0 + i < nblocks; ++i) {
        const uint32_t v = block_sums[t0 + i];
        block_sums[t0 + i] = (uint32_t)run;  // the caller checked that the total fits 32 bits
        run += v;
    }
}
__global__ void __launch_bounds__(kCtaThreads) rank_base_kernel(IndexView ix, const uint32_t* block_sums) {
    __shared__ uint32_t ws[kCtaThreads / 32];
    const uint64_t b0 = (uint64_t)blockIdx.x * kScanChunk + threadIdx.x * 4;
    uint32_t occ[4], v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        occ[j] = b0 + j < ix.nbuckets ? bucket_occupancy(ix.slots, b0 + j) : 0u;
        v += occ[j];
    }
    uint32_t total;
    uint32_t run = block_sums[blockIdx.x] + cta_exclusive_sum(v, ws, total);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (b0 + j < ix.nbuckets) ix.rank_base[b0 + j] = run;
        run += occ[j];
    }
}

__global__ void slot_perm_kernel(IndexView ix, const uint64_t* __restrict__ key56, uint64_t n, uint32_t* perm) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = key56[i];
    uint32_t r = 0xffffffffu;  // not a canonical k-mer / not in this table: its count is always 0
    if (key != kKey56Max) {
        uint32_t b = bucket_of(key, ix.nb_total) - ix.b_base;
        if (b < ix.nbuckets) {
            for (uint32_t tries = 0; tries < ix.nbuckets; ++tries) {
                uint64_t v[4];
                ld_bucket(ix.slots + 4ull * b, v);
                bool saw_empty;
                uint64_t seen = 0;
                const int hs = match_slot(v, key, saw_empty, seen);
                if (hs >= 0) { r = ix.rank_base[b] + (uint32_t)hs; break; }
                if (saw_empty) break;
                b = (b + 1 == ix.nbuckets) ? 0 : b + 1;
            }
        }
    }
    perm[i] = r;
}

template <typename OutT>
__global__ void gather_counts_kernel(const uint8_t* __restrict__ cvec, const uint32_t* __restrict__ perm,
                                     const uint64_t* __restrict__ idx, uint64_t n, OutT* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = perm[i];
    out[idx ? idx[i] : i] = (OutT)(r == 0xffffffffu ? 0u : (uint32_t)cvec[r]);
}

__global__ void scatter_bytes_kernel(const uint8_t* __restrict__ in, const uint32_t* __restrict__ perm, uint64_t n, uint8_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = perm[i];
    if (r != 0xffffffffu && in[i]) out[r] = in[i];  // duplicates of a key share a slot: any non-zero flag marks it
}

// ---------------------------------------------------------------------------
// probe + count
// ---------------------------------------------------------------------------
// c = min(255, c + n) on the low byte of a slot; the key bits can never be touched.
__device__ __forceinline__ void slot_sat_add(uint64_t* p, uint64_t seen, uint32_t n) {
    uint64_t old = seen;
    for (;;) {
        uint32_t c = (uint32_t)(old & 0xffu);
        if (c == 255u) return;  // saturation is absorbing
        uint32_t add = min(n, 255u - c);
        uint64_t prev = atomicCAS((unsigned long long*)p, old, old + add);
        if (prev == old) return;
        old = prev;
    }
}

// The same on the dense count vector of the partitioned path (IndexView::cvec): byte r, by a 32-bit CAS on
// the word that holds it.  Rare paths only (a key whose list is full, a hit that spilled over a slice edge).
// kSystem: the vector lives on a peer GPU (mapped over NVLink); .gpu scope is not atomic across devices.
template <bool kSystem>
__device__ __forceinline__ void cvec_sat_add(uint8_t* cvec, uint64_t r, uint32_t n) {
    uint32_t* w = (uint32_t*)(cvec + (r & ~3ULL));
    const uint32_t sh = (uint32_t)(r & 3u) * 8u;
    uint32_t old = *(volatile uint32_t*)w;
    for (;;) {
        const uint32_t c = (old >> sh) & 0xffu;
        if (c == 255u) return;
        const uint32_t nw = old + (min(n, 255u - c) << sh);
        const uint32_t prev = kSystem ? atomicCAS_system(w, old, nw) : atomicCAS(w, old, nw);
        if (prev == old) return;
        old = prev;
    }
}

// Slot match for a lookup: key must not be kKey56Max (callers peel that one off), so an empty slot
// (all ones) never matches.  `seen` gets the matching slot's value.
__device__ __forceinline__ int match_slot(const uint64_t (&v)[4], uint64_t key, bool& saw_empty, uint64_t& seen) {
    int hs = -1;
    saw_empty = false;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if ((v[s] >> 8) == key) { hs = s; seen = v[s]; }
        if (v[s] == kSlotEmpty) saw_empty = true;
    }
    return hs;
}

constexpr int kProbeBatch = 8;  // sector loads in flight per lane in the microbenchmark and default kernel

template <int N, typename T>
__device__ __forceinline__ T pick(const T (&a)[N], uint32_t i) {  // a[i] without local memory
    T r = a[0];
#pragma unroll
    for (int j = 1; j < N; ++j) r = (i == (uint32_t)j) ? a[j] : r;
    return r;
}

// All 32 lanes call this together (ballot / match inside).
// K1 hands over kBatch keys and their emit mask; K2 probes one 32-byte bucket per key with kBatch
// sector loads in flight per lane; K3 adds to the 8-bit counter in the matched slot.
//  * A key whose home bucket is full and holds no match may have spilled into the next bucket.
//    That is rare per key but near-certain per warp, so those probes are finished in ONE
//    out-of-line loop per batch instead of one warp-wide loop per position.
//  * The CAS of a position is issued as soon as the batch is matched; its result is only checked
//    after all of them are in flight, so the round trips overlap instead of serialising.
// st[b]: bit 31 hit, bits 16-23 count seen, bits 8-9 slot in bucket; later bit 30 CAS issued and
// bits 0-7 the amount to add.
template <int kBatch>
__device__ __forceinline__ void probe_and_count(const IndexView& ix, const uint64_t (&keys)[kBatch], uint32_t emit,
                                                uint32_t& n_hit) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t bk[kBatch], st[kBatch];
    uint64_t prev[kBatch];
    const uint32_t havem = emit & ((1u << kBatch) - 1);
    uint32_t pend = 0;  // positions whose search continues in the next bucket
    {
        uint64_t v[kBatch][4];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            bk[b] = bucket_of(keys[b], ix.nb_total) - ix.b_base;
            if ((havem >> b) & 1u) ld_bucket(ix.slots + 4ull * bk[b], v[b]);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const uint64_t key = keys[b];
            const bool have = (havem >> b) & 1u;
            const uint32_t want_hi = (uint32_t)(key >> 24);
            const uint32_t want_lo = (uint32_t)(key << 8);
            // slot == key<<8 | count  <=>  high words equal and low words differ only in the count byte
            const bool m0 = (uint32_t)(v[b][0] >> 32) == want_hi && (((uint32_t)v[b][0] ^ want_lo) < 256u);
            const bool m1 = (uint32_t)(v[b][1] >> 32) == want_hi && (((uint32_t)v[b][1] ^ want_lo) < 256u);
            const bool m2 = (uint32_t)(v[b][2] >> 32) == want_hi && (((uint32_t)v[b][2] ^ want_lo) < 256u);
            const bool m3 = (uint32_t)(v[b][3] >> 32) == want_hi && (((uint32_t)v[b][3] ^ want_lo) < 256u);
            // insertion fills a bucket front to back, so "bucket not full" == "last slot empty"
            const bool full = v[b][3] != kSlotEmpty;
            const bool hit = have && (m0 || m1 || m2 || m3);
            const uint32_t hs = m1 ? 1u : (m2 ? 2u : (m3 ? 3u : 0u));
            const uint32_t lo = m1 ? (uint32_t)v[b][1] : (m2 ? (uint32_t)v[b][2] : (m3 ? (uint32_t)v[b][3] : (uint32_t)v[b][0]));
            st[b] = hit ? (0x80000000u | ((lo & 0xffu) << 16) | (hs << 8)) : 0u;
            if (have && !hit && full) pend |= 1u << b;
        }
    }
    while (__any_sync(kFullMask, pend != 0)) {
        if (pend) {
            const uint32_t b = (uint32_t)__ffs(pend) - 1;
            const uint64_t key = pick<kBatch>(keys, b);
            uint32_t nb = pick<kBatch>(bk, b) + 1;
            if (nb == ix.nbuckets) nb = 0;
            uint64_t w[4];
            ld_bucket(ix.slots + 4ull * nb, w);
            bool saw_empty;
            uint64_t sv = 0;
            const int h2 = match_slot(w, key, saw_empty, sv);
            const uint32_t ns = h2 >= 0 ? (0x80000000u | (((uint32_t)sv & 0xffu) << 16) | ((uint32_t)h2 << 8)) : 0u;
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                if (b == (uint32_t)j) {
                    bk[j] = nb;
                    st[j] = ns;
                }
            }
            if (h2 >= 0 || saw_empty) pend &= pend - 1;  // resolved; otherwise keep walking
        }
    }
    // K3: warp-aggregated saturating add; one CAS per distinct slot per position step
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
        const bool hit = st[b] >> 31;
        const uint32_t hm = __ballot_sync(kFullMask, hit);
        if (hit) {
            n_hit += 1;
            const uint32_t hs = (st[b] >> 8) & 3u, cnt = (st[b] >> 16) & 0xffu;
            const uint64_t slot = 4ull * bk[b] + hs;
            const uint32_t peers = (hm & (hm - 1)) ? __match_any_sync(hm, slot) : hm;
            st[b] &= ~0x400000ffu;
            if ((uint32_t)(__ffs(peers) - 1) == lane && cnt != 255u) {
                const uint32_t add = min((uint32_t)__popc(peers), 255u - cnt);
                const uint64_t seen = (keys[b] << 8) | cnt;
                prev[b] = atomicCAS((unsigned long long*)(ix.slots + slot), seen, seen + add);
                st[b] |= 0x40000000u | (uint32_t)__popc(peers);
            }
        }
    }
    // verify; a lost race (or a key this lane hit twice in the batch) retries here
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
        if (st[b] & 0x40000000u) {
            const uint64_t seen = (keys[b] << 8) | ((st[b] >> 16) & 0xffu);
            if (prev[b] != seen)
                slot_sat_add(ix.slots + 4ull * bk[b] + ((st[b] >> 8) & 3u), prev[b], st[b] & 0xffu);
        }
    }
}

// kEnc: which encoder -- kEncAny the byte-wise state machine (any k), kEncOdd / kEncEven the position-parallel window
// encoders (vg_device.cuh)
template <int kEnc, int kBatch>
__global__ void __launch_bounds__(kCtaThreads, kBatch >= 8 ? 2 : 3)
count_kernel(IndexView ix, Chunk c, int64_t ntiles, CountStats* stats) {
    constexpr bool kOdd = kEnc != kEncAny;  // a window encoder
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ unsigned long long blk[2];
    if (c.skip && *c.skip) return;
    lut_init(lut, kEnc == kEncEven);
    if (threadIdx.x < 2) blk[threadIdx.x] = 0;
    __syncthreads();
    KmerParams kp{ix.k, ix.mask};
    uint32_t n_pos = 0, n_hit = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t off = t * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        if (kOdd) {
            typename EncoderOf<kEnc>::type enc;
            enc.init(c, off, kp, lut);
            // rolled on purpose: one probe batch of registers, and a loop body that stays in the I-cache
#pragma unroll 1
            for (int part = 0; part < 16 / kBatch; ++part) {
                uint64_t keys[kBatch];
                uint32_t emit = enc.template next<kBatch, false>(kp, keys);
                n_pos += __popc(emit);
                probe_and_count<kBatch>(ix, keys, emit, n_hit);
            }
        } else {
            uint64_t k16[16];
            uint32_t emit = encode_keys_any<false>(c, off, kp, lut, k16);
            n_pos += __popc(emit);
#pragma unroll
            for (int part = 0; part < 16 / kBatch; ++part) {
                uint64_t keys[kBatch];
#pragma unroll
                for (int j = 0; j < kBatch; ++j) keys[j] = k16[part * kBatch + j];
                probe_and_count<kBatch>(ix, keys, (emit >> (part * kBatch)) & ((1u << kBatch) - 1), n_hit);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        n_pos += __shfl_xor_sync(kFullMask, n_pos, d);
        n_hit += __shfl_xor_sync(kFullMask, n_hit, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&blk[0], (unsigned long long)n_pos);
        atomicAdd(&blk[1], (unsigned long long)n_hit);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&stats->positions, blk[0]);
        atomicAdd(&stats->hits, blk[1]);
    }
}

// ---------------------------------------------------------------------------
// partitioned probing: scatter by table slice, then probe slice by slice out of L2
// ---------------------------------------------------------------------------
// Shared memory of the scatter (dynamic, sized by the launcher):
//   bins[P * stride]     the tile's k-mers, one fixed-capacity bin per table slice; the stride is odd
//                        (in 8-byte words) so that equal ranks of different bins fall into different banks
//   base[P]              where the tile's run starts in each slice's key list (after reservation)
//   hist[P], cnt[P], fit[P]   k-mers per slice: seen / in the bin / fitting the slice's key list
// plus the nt4 LUT.
struct ScatterCfg {
    uint32_t cap;     // keys per bin; the rare overflow goes to global memory key by key
    uint32_t stride;  // cap | 1
};
// c_magic[d] = 2^32 / d + 1, so that x / d == __umulhi(x, c_magic[d]) for x < 2^32 / d: the copy-out
// walks bins x ranks with the tile's largest bin as the row length, whatever that turns out to be.
constexpr uint32_t kMaxBinCap = 4095;
__constant__ uint32_t c_magic[kMaxBinCap + 1];

// Rare path of the scatter: the key's partition buffer is full (a very skewed round, e.g. thousands
// of identical reads).  Probe it right here, one key at a time, so no key is ever dropped and no
// worst-case overflow buffer has to exist.  Kept out of line: it must not cost the hot path registers.
// `slots` / `b` are the table that owns the key and its home bucket there: for a sharded index that may be
// a peer GPU's table (64-bit CAS works over NVLink; no sweep runs while keys are being scattered).
struct TableRef {  // a table a key may be probed in directly: this GPU's, or (sharded index) a peer's over NVLink
    uint64_t* slots;
    uint32_t* rank_base;
    uint8_t* cvec;
    bool remote;
};
__device__ __noinline__ void probe_one_direct(TableRef t, uint32_t nbuckets, uint32_t b, uint64_t key, CountStats* stats) {
    for (uint32_t tries = 0; tries < nbuckets; ++tries) {
        uint64_t v[4];
        ld_bucket(t.slots + 4ull * b, v);
        bool saw_empty;
        uint64_t seen = 0;
        const int hs = match_slot(v, key, saw_empty, seen);
        if (hs >= 0) {
            const uint64_t r = (uint64_t)t.rank_base[b] + (uint32_t)hs;
            if (t.remote) cvec_sat_add<true>(t.cvec, r, 1u);
            else cvec_sat_add<false>(t.cvec, r, 1u);
            atomicAdd(&stats->hits, 1ull);
            return;
        }
        if (saw_empty) return;
        b = (b + 1 == nbuckets) ? 0 : b + 1;
    }
}
// The table that owns `key` and the key's home bucket there.
__device__ __forceinline__ void owner_table(const IndexView& ix, const PartView& pv, uint64_t key, TableRef& t, uint32_t& b) {
    const uint32_t bg = bucket_of(key, ix.nb_total);
    t = TableRef{ix.slots, ix.rank_base, ix.cvec, false};
    b = bg;
    if (pv.world > 1) {
        const uint32_t o = bg / ix.nbuckets;
        t = TableRef{pv.peer_slots[o], pv.peer_rank_base[o], pv.peer_cvec[o], o != pv.rank};
        b = bg - o * ix.nbuckets;
    }
}
__device__ __forceinline__ void probe_one_direct(const IndexView& ix, const PartView& pv, uint64_t key, CountStats* stats) {
    TableRef t;
    uint32_t b;
    owner_table(ix, pv, key, t, b);
    probe_one_direct(t, ix.nbuckets, b, key, stats);
}
// Where slice p's keys from this GPU go: the slice owner's key list for source `rank`.
__device__ __forceinline__ uint64_t* list_of(const PartView& pv, uint32_t p) {
    if (pv.world > 1) {
        const uint32_t o = p / pv.P_local, pl = p - o * pv.P_local;
        return pv.peer_keybuf[o] + ((uint64_t)pl * pv.world + pv.rank) * pv.cap;
    }
    return pv.keybuf + (uint64_t)p * pv.cap;
}

// Every (k - span + 1)-mer of every index k-mer, canonical (see shared_smer).
__global__ void prefilter_build_kernel(uint32_t* words, uint32_t nwords, const uint64_t* __restrict__ key56, uint64_t n,
                                       uint32_t k, uint32_t span) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = key56[i];
    if (key == kKey56Max) return;
    const uint32_t ks = k - (span - 1);
    const uint64_t smask = (1ULL << (2 * ks)) - 1;
    for (uint32_t h = 0; h < span; ++h) {
        const uint64_t sub = (key >> (2 * h)) & smask, rc = revcomp2k(sub, ks);
        uint32_t w, bits;
        prefilter_slot(sub < rc ? sub : rc, nwords, w, bits);
        atomicOr(words + w, prefilter_mask(bits));
    }
}

__device__ __forceinline__ void st_shared_u64(uint32_t addr, uint64_t v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_shared_u64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

// A key that finds its tile bin full: reserve one place in the slice's key list right away.
__device__ __noinline__ void scatter_one_global(unsigned long long* cursor_p, uint64_t* list, uint64_t cap, TableRef t,
                                                uint32_t nbuckets, uint32_t b, uint64_t key, CountStats* stats) {
    const unsigned long long pos = atomicAdd(cursor_p, 1ull);
    if (pos < cap) list[pos] = key;
    else probe_one_direct(t, nbuckets, b, key, stats);
}

// ---- bulk async copies (TMA, cp.async.bulk) with mbarrier completion: the tile load of scatter_kernel<.., kTma> ----
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on mbar
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mbar) : "memory");
}

// K1 of the partitioned path.  Per 4 KiB CTA tile, eight positions per lane at a time: encode + hash,
// drop what the presence pre-filter rules out, and append each survivor to its table slice's bin in
// shared memory (rank = shared-memory atomic).  After a barrier the bins are copied out: one global
// reservation per slice and tile (a thread per slice), a second barrier, then a thread per bin slot
// -- a tile's run for a slice is contiguous in the slice's key list, so the 8-byte stores coalesce
// run by run whatever the number of slices (a warp per bin wastes lanes once bins get short).
// kTma: the tile's text (plus the 32 bytes in front of it) is staged in shared memory by a bulk async copy that one
// thread issues a whole tile ahead (two buffers, one mbarrier each): the load latency never sits in front of the
// encoder and the LSU only sees shared-memory loads.
// kK: the k-mer length as a compile-time constant (27, the reference's default) or 0 = whatever the index says; with a
// constant k the validity windows, the window offsets and the masks of the encoder fold into immediates.
template <int kEnc, bool kTma, int kSpan, int kK>
__global__ void __launch_bounds__(kCtaThreads, 4)
scatter_kernel(IndexView ix, PartView pv, PrefilterView pf, ScatterCfg cfg, Chunk c, int64_t first_tile, int64_t ntiles,
               CountStats* stats) {
    constexpr bool kOdd = kEnc != kEncAny;  // a window encoder (odd k, or even k with the palindrome rule)
    extern __shared__ __align__(16) uint64_t smem_bins[];  // indexed directly: STS / LDS with the base folded in
    const uint32_t P = pv.P, cap = cfg.cap;
    unsigned long long* base_s = reinterpret_cast<unsigned long long*>(smem_bins + (size_t)P * cfg.stride);
    uint32_t* hist = reinterpret_cast<uint32_t*>(base_s + P);
    uint32_t* cnt_s = hist + P;
    uint32_t* fit_s = cnt_s + P;
    __shared__ __align__(16) uint8_t lut[kLutBytes];  // static: its address is a constant, one add less per base looked up
    __shared__ unsigned long long blk_pos;
    __shared__ uint32_t max_cnt;
    if (c.skip && *c.skip) return;
    lut_init(lut, kEnc == kEncEven);
    if (threadIdx.x == 0) blk_pos = 0, max_cnt = 0;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const KmerParams kp{kK ? (uint32_t)kK : ix.k, kK ? ((1ULL << (2 * (kK ? kK : 1))) - 1) : ix.mask};
    const uint32_t lane = threadIdx.x & 31;
    // 32-bit shared-window address of the bins, made opaque: left to itself the compiler re-derives it
    // (S2UR SR_CgaCtaId + three uniform ops) in front of every single bin store
    uint32_t bins_s = (uint32_t)__cvta_generic_to_shared(smem_bins);
    asm volatile("" : "+r"(bins_s));
    uint32_t stride = cfg.stride;  // likewise: otherwise a constant-bank load sits between every rank and its store
    asm volatile("" : "+r"(stride));
    const uint32_t base_a = bins_s + ((P * stride) << 3);  // base_s[], same address space, same reason
    // the three constants of a bin insert, likewise pinned: left in the parameter bank they are re-loaded (LDCU) for
    // every single read position
    // (an empty asm does not stop ptxas from re-loading a kernel parameter; a zero that comes out of shared memory does)
    const uint32_t zero_r = max_cnt;  // 0: set above, read after the barrier
    const uint32_t nb_total_r = ix.nb_total + zero_r, shift_r = pv.shift + zero_r, cap_r = cap + zero_r;
    stride += zero_r;
    const uint32_t hist_a = (uint32_t)__cvta_generic_to_shared(hist) + zero_r;
    uint32_t n_pos = 0;
    constexpr int kStage = kTileBytes + 32;
    __shared__ __align__(16) uint8_t tile_s[kTma ? 2 * kStage : 16];
    __shared__ __align__(8) uint64_t tile_bar[2];
    const uint32_t tile_a = (uint32_t)__cvta_generic_to_shared(tile_s), bar_a = (uint32_t)__cvta_generic_to_shared(tile_bar);
    // stage tile `t` (grid-stride numbering) into buffer `b`: bytes [start, start + kStage) of the chunk, clipped to the
    // 16-byte aligned extent of the caller's buffer
    auto stage = [&](int64_t t, uint32_t b) {
        const int64_t start = (first_tile + t) * kTileBytes - 32;
        const int64_t s0 = start < 0 ? 0 : start, e0 = min(start + kStage, (c.hi + 15) & ~(int64_t)15);
        if (e0 > s0) {
            mbar_expect_tx(bar_a + 8 * b, (uint32_t)(e0 - s0));
            bulk_g2s(tile_a + b * kStage + (uint32_t)(s0 - start), c.al + s0, (uint32_t)(e0 - s0), bar_a + 8 * b);
        } else {
            mbar_arrive(bar_a + 8 * b);
        }
    };
    if (kTma && kOdd) {
        if (threadIdx.x == 0) {
            mbar_init(bar_a, 1);
            mbar_init(bar_a + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if ((int64_t)blockIdx.x < ntiles) stage(blockIdx.x, 0);
        }
        __syncthreads();
    }
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int64_t off = (first_tile + t) * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        if (kTma && kOdd) {
            // every thread is past the previous tile's barriers, so nobody still reads the other buffer
            if (threadIdx.x == 0 && t + gridDim.x < ntiles) stage(t + gridDim.x, (it + 1) & 1);
            mbar_wait(bar_a + 8 * (it & 1), (it >> 1) & 1);
        }
        // ---- encode, filter, bin -----------------------------------------------------------------
        constexpr int kGroups = 8 / kSpan;
        constexpr uint32_t kGroupMask = (1u << kSpan) - 1u;
        auto bin8 = [&](const uint64_t (&keys)[8], const uint64_t (&pairs)[kGroups], uint32_t emit) {
            if (pf.words) {  // presence pre-filter, one lookup per kSpan positions: never a false negative
                uint32_t fw[kGroups], fb[kGroups];
#pragma unroll
                for (int q = 0; q < kGroups; ++q) {
                    uint32_t w;
                    prefilter_slot(pairs[q], pf.nwords, w, fb[q]);
                    fw[q] = ((emit >> (kSpan * q)) & kGroupMask) ? __ldg(pf.words + w) : 0u;
                }
#pragma unroll
                for (int q = 0; q < kGroups; ++q) {
                    const uint32_t m = prefilter_mask(fb[q]);
                    if ((fw[q] & m) != m) emit &= ~(kGroupMask << (kSpan * q));
                }
            }
            uint32_t over = 0;   // keys whose bin is full (rare): handled after the hot loop
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if ((emit >> j) & 1u) {
                    const uint32_t p = __umulhi(key_mix_hi(keys[j]), nb_total_r) >> shift_r;
                    uint32_t r;
                    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(hist_a + (p << 2)) : "memory");
                    if (r < cap_r) st_shared_u64(bins_s + ((p * stride + r) << 3), keys[j]);
                    else over |= 1u << j;
                }
            }
            if (over) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if ((over >> j) & 1u) {
                        TableRef tr;
                        uint32_t b;
                        owner_table(ix, pv, keys[j], tr, b);
                        const uint32_t p = bucket_of(keys[j], ix.nb_total) >> pv.shift;
                        scatter_one_global(&pv.cursor[p], list_of(pv, p), pv.cap, tr, ix.nbuckets, b, keys[j], stats);
                    }
            }
        };
        if (kOdd) {
            typename EncoderOf<kEnc>::type enc;
            if (kTma) enc.init(c, SharedText{tile_a + (it & 1) * kStage, (first_tile + t) * kTileBytes - 32}, off, kp, lut);
            else enc.init(c, off, kp, lut);
            uint64_t keys[8], pairs[kGroups];
            uint32_t emit = enc.template next<8, false, true, kSpan>(kp, keys, pairs);
            n_pos += __popc(emit);
            bin8(keys, pairs, emit);
            emit = enc.template next<8, false, true, kSpan>(kp, keys, pairs);
            n_pos += __popc(emit);
            bin8(keys, pairs, emit);
        } else {
            uint64_t k16[16], p8[2 * kGroups], keys[8], pairs[kGroups];
            const uint32_t emit = encode_keys_any<false, kSpan>(c, off, kp, lut, k16, p8);
            n_pos += __popc(emit);
#pragma unroll
            for (int j = 0; j < 8; ++j) keys[j] = k16[j];
#pragma unroll
            for (int q = 0; q < kGroups; ++q) pairs[q] = p8[q];
            bin8(keys, pairs, emit & 0xffu);
#pragma unroll
            for (int j = 0; j < 8; ++j) keys[j] = k16[8 + j];
#pragma unroll
            for (int q = 0; q < kGroups; ++q) pairs[q] = p8[kGroups + q];
            bin8(keys, pairs, emit >> 8);
        }
        __syncthreads();
        // ---- copy-out: reserve (one global atomic per slice and tile, all slices at once) ... ----------
        for (uint32_t p = threadIdx.x; p < P; p += blockDim.x) {
            const uint32_t n = min(hist[p], cap);
            hist[p] = 0;  // re-armed for the next tile
            unsigned long long b = 0;
            if (n) b = atomicAdd(&pv.cursor[p], (unsigned long long)n);
            cnt_s[p] = n;
            if (n) atomicMax(&max_cnt, n);
            fit_s[p] = b >= pv.cap ? 0u : (uint32_t)min((unsigned long long)n, pv.cap - b);
            base_s[p] = (unsigned long long)(list_of(pv, p) + b);
        }
        __syncthreads();
        // ---- ... then one bin slot per thread and step; a bin's keys are contiguous on both sides ----
        const uint32_t row = max_cnt, nslots = P * row, magic = c_magic[row];
        for (uint32_t q = threadIdx.x; q < nslots; q += blockDim.x) {
            const uint32_t p = __umulhi(q, magic), i = q - p * row;
            if (i < cnt_s[p]) {
                const uint64_t key = ld_shared_u64(bins_s + ((p * stride + i) << 3));
                if (i < fit_s[p]) reinterpret_cast<uint64_t*>(ld_shared_u64(base_a + (p << 3)))[i] = key;
                else probe_one_direct(ix, pv, key, stats);  // slice list full (a very skewed round): still exact
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) max_cnt = 0;  // read again only after the next tile's two barriers
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_pos += __shfl_xor_sync(kFullMask, n_pos, d);
    if (lane == 0 && n_pos) atomicAdd(&blk_pos, (unsigned long long)n_pos);
    __syncthreads();
    if (threadIdx.x == 0 && blk_pos) atomicAdd(&stats->positions, blk_pos);
}

__device__ __forceinline__ uint64_t ld_key_stream(const uint64_t* p) {
    uint64_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// Second level of the scatter (PartView::sub_bits > 0): one coarse partition's key list(s) -> the key lists of its
// Q table slices.  The binning of scatter_kernel with keys in place of text: 4096 keys per CTA tile (a thread takes
// every 256th, so the 8-byte loads coalesce), bins in shared memory, one reservation per slice and tile, copy-out a
// thread per bin slot.  Q <= 64, so the bins are always fat.  kMulti (sharded index): the coarse partition has `nsub`
// lists, one per source GPU, pv.cap keys apart, their fill counts `count_stride` apart.
template <bool kMulti>
__global__ void __launch_bounds__(kCtaThreads, 4)
rescatter_kernel(IndexView ix, PartView pv, ScatterCfg cfg, const uint64_t* __restrict__ lists, const unsigned long long* count_ptr,
                 uint32_t nsub, uint32_t count_stride, uint32_t slice0, uint32_t Q, CountStats* stats) {
    extern __shared__ __align__(16) uint64_t smem_bins[];
    const uint32_t cap = cfg.cap;
    unsigned long long* base_s = reinterpret_cast<unsigned long long*>(smem_bins + (size_t)Q * cfg.stride);
    uint32_t* hist = reinterpret_cast<uint32_t*>(base_s + Q);
    uint32_t* cnt_s = hist + Q;
    uint32_t* fit_s = cnt_s + Q;
    __shared__ uint32_t max_cnt;
    if (threadIdx.x == 0) max_cnt = 0;
    for (uint32_t i = threadIdx.x; i < Q; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    uint32_t bins_s = (uint32_t)__cvta_generic_to_shared(smem_bins);
    asm volatile("" : "+r"(bins_s));
    uint32_t stride = cfg.stride;
    asm volatile("" : "+r"(stride));
    const uint32_t base_a = bins_s + ((Q * stride) << 3);
    const TableRef mine{ix.slots, ix.rank_base, ix.cvec, false};
    constexpr uint32_t kTileKeys = kCtaThreads * 16;
#pragma unroll 1
    for (uint32_t sub = 0; sub < (kMulti ? nsub : 1u); ++sub) {
        const uint64_t* list = kMulti ? lists + (uint64_t)sub * pv.cap : lists;
        const uint64_t n = min((uint64_t)count_ptr[kMulti ? (size_t)sub * count_stride : 0], pv.cap);
        const uint64_t ntiles = (n + kTileKeys - 1) / kTileKeys;
        for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                uint64_t keys[8];
                uint32_t emit = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint64_t idx = t * kTileKeys + (uint64_t)(half * 8 + j) * kCtaThreads + threadIdx.x;
                    keys[j] = 0;
                    if (idx < n) {
                        keys[j] = ld_key_stream(list + idx);
                        emit |= 1u << j;
                    }
                }
                uint32_t over = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if ((emit >> j) & 1u) {
                        const uint32_t q = ((bucket_of(keys[j], ix.nb_total) - ix.b_base) >> pv.shift2) - slice0;
                        if (q < Q) {
                            const uint32_t r = atomicAdd(&hist[q], 1u);
                            if (r < cap) st_shared_u64(bins_s + ((q * stride + r) << 3), keys[j]);
                            else over |= 1u << j;
                        } else {
                            over |= 1u << j;  // cannot happen with a consistent geometry; still counted exactly
                        }
                    }
                }
                if (over) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((over >> j) & 1u) {
                            const uint32_t b = bucket_of(keys[j], ix.nb_total) - ix.b_base;
                            const uint32_t q = (b >> pv.shift2) - slice0;
                            if (q < Q) scatter_one_global(&pv.cursor2[q], pv.keybuf2 + (uint64_t)q * pv.cap2, pv.cap2, mine, ix.nbuckets, b, keys[j], stats);
                            else probe_one_direct(mine, ix.nbuckets, b, keys[j], stats);
                        }
                }
            }
            __syncthreads();
            for (uint32_t q = threadIdx.x; q < Q; q += blockDim.x) {
                const uint32_t m = min(hist[q], cap);
                hist[q] = 0;
                unsigned long long b = 0;
                if (m) b = atomicAdd(&pv.cursor2[q], (unsigned long long)m);
                cnt_s[q] = m;
                if (m) atomicMax(&max_cnt, m);
                fit_s[q] = b >= pv.cap2 ? 0u : (uint32_t)min((unsigned long long)m, pv.cap2 - b);
                base_s[q] = (unsigned long long)(pv.keybuf2 + (uint64_t)q * pv.cap2 + b);
            }
            __syncthreads();
            const uint32_t row = max_cnt, nslots = Q * row, magic = c_magic[row];
            for (uint32_t s = threadIdx.x; s < nslots; s += blockDim.x) {
                const uint32_t q = __umulhi(s, magic), i = s - q * row;
                if (i < cnt_s[q]) {
                    const uint64_t key = ld_shared_u64(bins_s + ((q * stride + i) << 3));
                    if (i < fit_s[q]) reinterpret_cast<uint64_t*>(ld_shared_u64(base_a + (q << 3)))[i] = key;
                    else probe_one_direct(mine, ix.nbuckets, bucket_of(key, ix.nb_total) - ix.b_base, key, stats);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) max_cnt = 0;
        }
    }
}

// K2 + K3 of the partitioned path.  All probes of a key list fall into one L2-resident slice of the
// table [b0, b1), so a hit is recorded with a fire-and-forget 32-bit reduction into the slice's
// side counters (also L2-resident) instead of a CAS round trip; retire_slice folds them into the
// saturating 8-bit slot counters once the slice has been probed by every key of the round.
// A hit outside the slice (a key that spilled over the slice edge) takes the CAS path.
template <int kBatch>
__device__ __forceinline__ void probe_and_red(const IndexView& ix, const uint64_t (&keys)[kBatch], uint32_t emit,
                                              uint32_t b0, uint32_t b1, uint32_t* ctr, uint32_t& n_hit) {
    uint32_t bk[kBatch], st[kBatch];  // st: bit 31 hit, bits 0-1 slot
    uint32_t pend = 0;                // positions whose search continues in the next bucket
    // A scattered 256-bit load costs 1.5 L1TEX cycles per lane, a 128-bit one 1.0 (tools/gather_bench.cu), and
    // at load factor 0.3 most keys sit in the first two slots of their bucket (buckets fill front to back):
    // look at slots 0-1 first, fetch slots 2-3 only where slot 1 is taken and nothing matched yet.
    uint32_t more = 0;                // positions that need the second half of their bucket
    auto half = [&](const uint64_t (&v)[2], uint64_t key, uint32_t first_slot, bool& second_taken) {
        const uint32_t want_hi = (uint32_t)(key >> 24), want_lo = (uint32_t)(key << 8);
        const bool m0 = (uint32_t)(v[0] >> 32) == want_hi && (uint32_t)v[0] == want_lo;  // slot == key << 8: the low
        const bool m1 = (uint32_t)(v[1] >> 32) == want_hi && (uint32_t)v[1] == want_lo;  // byte of a slot stays 0 here
        second_taken = v[1] != kSlotEmpty;
        return (m0 || m1) ? (0x80000000u | (first_slot + (m1 ? 1u : 0u))) : 0u;
    };
    {
        uint64_t v[kBatch][2];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            bk[b] = bucket_of(keys[b], ix.nb_total) - ix.b_base;
            v[b][0] = v[b][1] = kSlotEmpty;
            if ((emit >> b) & 1u) {
                const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(ix.slots + 4ull * bk[b]);
                v[b][0] = t.x;
                v[b][1] = t.y;
            }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            bool taken;
            st[b] = half(v[b], keys[b], 0u, taken);
            if (((emit >> b) & 1u) && !st[b] && taken) more |= 1u << b;
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            v[b][0] = v[b][1] = kSlotEmpty;
            if ((more >> b) & 1u) {
                const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(ix.slots + 4ull * bk[b] + 2);
                v[b][0] = t.x;
                v[b][1] = t.y;
            }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            if ((more >> b) & 1u) {
                bool full;
                st[b] = half(v[b], keys[b], 2u, full);
                if (!st[b] && full) pend |= 1u << b;
            }
        }
    }
    while (__any_sync(kFullMask, pend != 0)) {
        if (pend) {
            const uint32_t b = (uint32_t)__ffs(pend) - 1;
            const uint64_t key = pick<kBatch>(keys, b);
            uint32_t nb = pick<kBatch>(bk, b) + 1;
            if (nb == ix.nbuckets) nb = 0;
            uint64_t w[4];
            ld_bucket(ix.slots + 4ull * nb, w);
            bool saw_empty;
            uint64_t sv = 0;
            const int h2 = match_slot(w, key, saw_empty, sv);
            const uint32_t ns = h2 >= 0 ? (0x80000000u | (uint32_t)h2) : 0u;
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                if (b == (uint32_t)j) {
                    bk[j] = nb;
                    st[j] = ns;
                }
            }
            if (h2 >= 0 || saw_empty) pend &= pend - 1;
        }
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
        if (st[b] >> 31) {
            n_hit += 1;
            const uint32_t hs = st[b] & 3u;
            if (bk[b] >= b0 && bk[b] < b1) atomicAdd(ctr + (size_t)(bk[b] - b0) * 4 + hs, 1u);  // result unused: a RED, no round trip
            else cvec_sat_add<false>(ix.cvec, (uint64_t)ix.rank_base[bk[b]] + hs, 1u);           // spilled over the slice edge
        }
    }
}

// Folds the side counters of slice [b0, b1) into the count vector: c = min(255, c + hits) at rank_base[b] + slot.
// One bucket (16 bytes of counters) per thread and step, kRetireBatch buckets in flight per thread: the three
// dependent accesses (counters -> rank_base -> counts) are what this pass costs, so they are issued batch-wise and
// the launch that probed the slice has already pulled its rank_base and cvec lines into L2.  Neighbouring buckets own
// neighbouring bytes of cvec.  Plain byte stores are safe: nothing else touches these bytes while the slice is retired
// (the sweep needs >= 3 slices, so a key spilling over the table's end never lands in the slice being retired, and the
// 32-bit CAS of such a spill elsewhere never changes a byte it does not own).
constexpr int kRetireBatch = 4;
__device__ __forceinline__ void retire_slice(const IndexView& ix, uint32_t b0, uint32_t b1, uint32_t* ctr) {
    const uint64_t nb = b1 - b0;
    const uint32_t* rb = ix.rank_base + b0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nb; i0 += kRetireBatch * stride) {
        uint4 c[kRetireBatch];
        uint32_t base[kRetireBatch], w0[kRetireBatch], w1[kRetireBatch];
#pragma unroll
        for (int u = 0; u < kRetireBatch; ++u) {
            const uint64_t i = i0 + u * stride;
            c[u] = i < nb ? *reinterpret_cast<const uint4*>(ctr + 4 * i) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kRetireBatch; ++u) base[u] = (c[u].x | c[u].y | c[u].z | c[u].w) ? rb[i0 + u * stride] : 0u;
#pragma unroll
        for (int u = 0; u < kRetireBatch; ++u) {  // the bucket's four counts: two aligned words (cvec is padded)
            w0[u] = w1[u] = 0;
            if (c[u].x | c[u].y | c[u].z | c[u].w) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(ix.cvec + (base[u] & ~3u));
                w0[u] = wp[0];
                w1[u] = wp[1];
            }
        }
#pragma unroll
        for (int u = 0; u < kRetireBatch; ++u) {
            if (c[u].x | c[u].y | c[u].z | c[u].w) {
                const uint32_t old4 = __funnelshift_r(w0[u], w1[u], (base[u] & 3u) * 8u);
                uint8_t* cv = ix.cvec + base[u];
                const uint32_t cs[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (cs[j]) cv[j] = (uint8_t)min(255u, ((old4 >> (8 * j)) & 0xffu) + min(cs[j], 255u));
                *reinterpret_cast<uint4*>(ctr + 4 * (i0 + u * stride)) = make_uint4(0, 0, 0, 0);
            }
        }
    }
}

// One step of the partition sweep: retire the previous slice [r0, r1) (its side counters are in
// ctr_prev), pull slice [b0, b1) into L2 with sequential prefetches, then probe the slice's key
// list into ctr_cur.  Either part may be empty.  kMulti (sharded index): the slice has `nsub` key lists, one
// per source GPU, `cap` keys apart, their fill counts `count_stride` apart.
template <int kBatch, bool kMulti>
__global__ void __launch_bounds__(kCtaThreads, kBatch >= 8 ? 2 : 4)
probe_slice_kernel(IndexView ix, const uint64_t* __restrict__ lists, const unsigned long long* count_ptr, uint32_t nsub,
                   uint32_t count_stride, uint64_t cap, uint32_t b0, uint32_t b1, uint32_t* ctr_cur, uint32_t r0,
                   uint32_t r1, uint32_t* ctr_prev, uint32_t f0, uint32_t f1, uint32_t c0, uint32_t c1, CountStats* stats) {
    __shared__ unsigned long long blk_hit;
    if (threadIdx.x == 0) blk_hit = 0;
    __syncthreads();
    {   // buckets [f0, f1) -> L2: this slice, or (prefetch-ahead) the next one, whose probes then never wait for DRAM;
        // and what the NEXT launch needs to retire the slice probed here: its rank_base entries and its counts [c0, c1)
        const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const uint64_t gsz = (uint64_t)gridDim.x * blockDim.x;
        const char* tbl = (const char*)ix.slots;
        for (uint64_t line = (uint64_t)f0 * 32 / 128 + gtid; line * 128 < (uint64_t)f1 * 32; line += gsz)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(tbl + line * 128));
        const char* rbp = (const char*)ix.rank_base;
        for (uint64_t line = (uint64_t)b0 * 4 / 128 + gtid; line * 128 < (uint64_t)b1 * 4; line += gsz)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rbp + line * 128));
        const char* cvp = (const char*)ix.cvec;
        for (uint64_t line = (uint64_t)c0 / 128 + gtid; line * 128 < (uint64_t)c1; line += gsz)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(cvp + line * 128));
    }
    if (r1 > r0) retire_slice(ix, r0, r1, ctr_prev);  // independent of the probes below: no barrier in between
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp_gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    constexpr uint64_t kPerWarp = 32ull * kBatch;
    uint32_t n_hit = 0;
#pragma unroll 1
    for (uint32_t sub = 0; sub < (kMulti ? nsub : 1u); ++sub) {
    const uint64_t* list = kMulti ? lists + (uint64_t)sub * cap : lists;
    const uint64_t n = b1 > b0 ? min((uint64_t)count_ptr[kMulti ? (size_t)sub * count_stride : 0], cap) : 0;
    uint64_t keys[kBatch], next[kBatch];
    auto fetch = [&](uint64_t base, uint64_t (&dst)[kBatch]) {
        uint32_t e = 0;
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const uint64_t idx = base + (uint64_t)b * 32 + lane;
            dst[b] = 0;
            if (idx < n) {
                dst[b] = ld_key_stream(list + idx);
                e |= 1u << b;
            }
        }
        return e;
    };
    uint64_t base = warp_gid * kPerWarp;
    uint32_t emit = base < n ? fetch(base, keys) : 0u;
    for (; base < n; base += nwarps * kPerWarp) {
        // the next batch of keys streams in from HBM while this one is probed out of L2
        const uint64_t nbase = base + nwarps * kPerWarp;
        const uint32_t nemit = nbase < n ? fetch(nbase, next) : 0u;
        probe_and_red<kBatch>(ix, keys, emit, b0, b1, ctr_cur, n_hit);
#pragma unroll
        for (int b = 0; b < kBatch; ++b) keys[b] = next[b];
        emit = nemit;
    }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_hit += __shfl_xor_sync(kFullMask, n_hit, d);
    if (lane == 0 && n_hit) atomicAdd(&blk_hit, (unsigned long long)n_hit);
    __syncthreads();
    if (threadIdx.x == 0 && blk_hit) atomicAdd(&stats->hits, blk_hit);
}

// ---------------------------------------------------------------------------
// extraction: counts in the key order given at index creation
// ---------------------------------------------------------------------------
template <typename OutT>
__global__ void extract_kernel(IndexView ix, const uint64_t* __restrict__ key56, const uint64_t* __restrict__ idx, uint64_t n,
                               OutT* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = key56[i];
    uint32_t c = 0;
    if (key == kKey56Max) {
        c = 0;  // not a canonical k-mer: no read position can produce it
    } else {
        uint32_t b = bucket_of(key, ix.nb_total) - ix.b_base;
        for (uint32_t tries = 0; tries < ix.nbuckets; ++tries) {
            uint64_t v[4];
            ld_bucket(ix.slots + 4ull * b, v);
            bool saw_empty;
            uint64_t seen = 0;
            int hs = match_slot(v, key, saw_empty, seen);
            if (hs >= 0) { c = (uint32_t)(seen & 0xffu); break; }
            if (saw_empty) break;
            b = (b + 1 == ix.nbuckets) ? 0 : b + 1;
        }
    }
    out[idx ? idx[i] : i] = (OutT)c;
}

// ---------------------------------------------------------------------------
// peer-memory collectives of vg_comm (one process per GPU; peers mapped with CUDA IPC over NVLink)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void peer_barrier_kernel(PeerPtrs flags, int world, int rank, unsigned long long epoch,
                                    unsigned long long timeout_ns, unsigned int* timed_out) {
    const int t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    unsigned long long* theirs = (unsigned long long*)flags.p[t] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* mine = (const unsigned long long*)flags.p[rank] + t;
    const unsigned long long t0 = global_ns();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (global_ns() - t0 > timeout_ns) {  // a dead peer must not hang this GPU
            atomicExch(timed_out, 1u);
            break;
        }
        __nanosleep(256);
    }
}

__global__ void publish_counts_kernel(PartView pv) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pv.P) return;
    const uint32_t o = p / pv.P_local, pl = p - o * pv.P_local;
    const unsigned long long n = pv.cursor[p];
    pv.peer_incount[o][(size_t)pv.rank * pv.P_local + pl] = n < pv.cap ? n : pv.cap;
    pv.cursor[p] = 0;  // re-armed for the next round
}

// out = min(255, sum over ranks) per byte, 16 bytes per thread and step: even and odd bytes are summed in
// 16-bit lanes (world <= 16 ranks x 255 fits), clamped, and packed again.
__global__ void combine_counts_kernel(PeerPtrs src, int world, uint64_t n, uint8_t* out) {
    const uint64_t nvec = n / 16;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint32_t ev[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0};
        for (int r = 0; r < world; ++r) {
            const uint4 v = reinterpret_cast<const uint4*>(src.p[r])[i];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ev[j] += w[j] & 0x00ff00ffu;
                od[j] += (w[j] >> 8) & 0x00ff00ffu;
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __vminu2(ev[j], 0x00ff00ffu) | (__vminu2(od[j], 0x00ff00ffu) << 8);
        reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 15)) {
        const uint64_t i = nvec * 16 + threadIdx.x;
        uint32_t sum = 0;
        for (int r = 0; r < world; ++r) sum += reinterpret_cast<const uint8_t*>(src.p[r])[i];
        out[i] = (uint8_t)min(sum, 255u);
    }
}

// Count reduce of a replica group, in place and in slot order.  Phase 1 (reduce-scatter): this rank owns bytes
// [seg0, seg1) of the vector; it sums every rank's copy of them (its own included), clamps at 255 and writes the
// result back into its own copy -- peers only ever read the OTHER segments of this copy meanwhile.  Phase 2
// (all-gather, after a barrier): it fetches every other segment from the rank that owns it.  Per rank 2 (W-1)/W bytes
// per entry cross NVLink, against W bytes per entry when every rank reads all vectors.
__global__ void reduce_segment_kernel(PeerPtrs src, int world, int rank, uint64_t seg0, uint64_t seg1) {
    const uint64_t v0 = seg0 / 16, v1 = (seg1 + 15) / 16;  // segments are cut at multiples of 16 bytes (the vectors are padded)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += stride) {
        uint32_t ev[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0};
        for (int r = 0; r < world; ++r) {
            const uint4 v = reinterpret_cast<const uint4*>(src.p[r])[i];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ev[j] += w[j] & 0x00ff00ffu;
                od[j] += (w[j] >> 8) & 0x00ff00ffu;
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __vminu2(ev[j], 0x00ff00ffu) | (__vminu2(od[j], 0x00ff00ffu) << 8);
        reinterpret_cast<uint4*>(src.p[rank])[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
__global__ void gather_segments_kernel(PeerPtrs src, int world, int rank, uint64_t seg_bytes, uint64_t nbytes) {
    const uint64_t nvec = (nbytes + 15) / 16, seg_vec = seg_bytes / 16;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint4* mine = reinterpret_cast<uint4*>(src.p[rank]);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const int owner = (int)min((uint64_t)(world - 1), i / seg_vec);
        if (owner != rank) mine[i] = reinterpret_cast<const uint4*>(src.p[owner])[i];
    }
}

// ---------------------------------------------------------------------------
// count consumers: 256-bin histogram of c over a (static, per-graph) subset of the index entries --
// the device half of Varigraph::get_hom_kmer (src/varigraph.cpp:253-296), SURVEY 8f N1.
// ---------------------------------------------------------------------------
__global__ void histogram_kernel(const uint8_t* __restrict__ counts, const uint8_t* __restrict__ flags, uint64_t n,
                                 unsigned long long* hist) {
    __shared__ unsigned int sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (!flags || flags[i]) atomicAdd(&sh[counts[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// ---------------------------------------------------------------------------
// per-position keys (test hook + synthetic index construction for bench.py)
// out[p] = (hash << 8 | k) for the k-mer ENDING at byte p, or ~0.
// ---------------------------------------------------------------------------
template <int kEnc>
__global__ void __launch_bounds__(kCtaThreads) positions_kernel(KmerParams kp, Chunk c, int64_t ntiles, uint64_t* out) {
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    lut_init(lut, kEnc == kEncEven);
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t off = t * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        uint64_t keys[16];
        uint32_t emit = kEnc == kEncOdd ? encode_keys_odd(c, off, kp, lut, keys)
                        : kEnc == kEncEven ? encode_keys_even(c, off, kp, lut, keys)
                                           : encode_keys_any(c, off, kp, lut, keys);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int64_t p = off + j;
            if (p >= c.lo && p < c.hi)
                out[p - c.lo] = ((emit >> j) & 1u) ? ((keys[j] << 8) | kp.k) : kNoKmer;
        }
    }
}

// ---------------------------------------------------------------------------
// counting Bloom filter (construct side)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cell_sat_inc(uint8_t* cells, uint64_t pos) {
    uint32_t* w = (uint32_t*)(cells + (pos & ~3ULL));
    uint32_t sh = (uint32_t)(pos & 3u) * 8u;
    uint32_t old = *w;
    for (;;) {
        if (((old >> sh) & 0xffu) == 255u) return;
        uint32_t prev = atomicCAS(w, old, old + (1u << sh));
        if (prev == old) return;
        old = prev;
    }
}

// Even k through the window encoder as well: the byte-wise state machine walks back to the last newline, and a chromosome
// has none -- every lane inside a multi-Mb run of N walked the whole run (quadratic); now only the few lanes at the
// end of such a run do, once.
template <int kEnc>
__global__ void __launch_bounds__(kCtaThreads) cbf_add_kernel(CbfView cbf, KmerParams kp, Chunk c, int64_t first_tile,
                                                            int64_t ntiles, unsigned long long* added) {
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    lut_init(lut, kEnc == kEncEven);
    FastMod64 fm{cbf.magic_hi, cbf.magic_lo, cbf.m};
    uint32_t n = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t off = (first_tile + t) * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        uint64_t keys[16];
        uint32_t emit = kEnc == kEncOdd ? encode_keys_odd(c, off, kp, lut, keys)
                        : kEnc == kEncEven ? encode_keys_even(c, off, kp, lut, keys)
                                           : encode_keys_any(c, off, kp, lut, keys);
        n += __popc(emit);
        // (Loading a k-mer's seven cells together before one CAS each was tried: 2.03 vs 2.76 G k-mers/s -- the rolled
        // loop it needs keeps fewer k-mers in flight per lane than this unrolled one.)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (!((emit >> j) & 1u)) continue;
            uint64_t k1 = murmur3_k1((keys[j] << 8) | kp.k);
            for (uint32_t h = 0; h < cbf.num_hashes; ++h)
                cell_sat_inc(cbf.cells, fastmod64(murmur3_sum_from_k1(k1, cbf.seeds[h]), fm));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(kFullMask, n, d);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(added, (unsigned long long)n);
}

__global__ void cbf_query_kernel(CbfView cbf, const uint64_t* __restrict__ keys, uint64_t n, uint8_t* count,
                                 uint8_t* find) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FastMod64 fm{cbf.magic_hi, cbf.magic_lo, cbf.m};
    uint64_t k1 = murmur3_k1(keys[i]);
    uint32_t lo = 255;
    for (uint32_t h = 0; h < cbf.num_hashes; ++h) {
        uint32_t v = cbf.cells[fastmod64(murmur3_sum_from_k1(k1, cbf.seeds[h]), fm)];
        lo = min(lo, v);
    }
    if (count) count[i] = (uint8_t)lo;
    if (find) find[i] = lo != 0 ? 1 : 0;
}

// ---------------------------------------------------------------------------
// roofline denominator: uniform random 32-byte sector gathers over a table >> L2
// (the "random-access HBM peak" BASELINE.json's metric names).  Same load
// instruction and the same number of loads in flight per lane as probe_and_count.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads) random_sector_kernel(const uint64_t* table, uint32_t nbuckets,
                                                                  uint32_t rounds, unsigned long long* sink) {
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    uint64_t acc = 0;
    for (uint32_t r = 0; r < rounds; ++r) {
        uint64_t v[kProbeBatch][4];
#pragma unroll
        for (int b = 0; b < kProbeBatch; ++b) {
            x ^= x >> 12; x ^= x << 25; x ^= x >> 27;  // xorshift64*
            uint32_t bk = __umulhi((uint32_t)((x * 0x2545F4914F6CDD1DULL) >> 32), nbuckets);
            ld_bucket(table + 4ull * bk, v[b]);
        }
#pragma unroll
        for (int b = 0; b < kProbeBatch; ++b) acc ^= v[b][0] ^ v[b][1] ^ v[b][2] ^ v[b][3];
    }
    if (acc == 0x0123456789abcdefULL) atomicAdd(sink, 1ull);
}

// ---------------------------------------------------------------------------
// on-device FASTQ parsing: the host ships raw text of plain four-line FASTQ, cut at
// record boundaries; these kernels find the lines, check EVERY record against what kseq would accept
// as a four-line record (include/kseq.h:192-232), blank everything but the sequence lines and sum
// seq.l (mReadBase, src/fastq_kmer.cpp:105).  A block that fails the check is not counted (Chunk::skip),
// nor is any later block of its file: the host re-parses from there with the kseq reader, so the result
// is the reference's for any input.
// ---------------------------------------------------------------------------
// bit j of the result: byte j of the 16-byte segment at `off` is '\n' (bytes at or beyond len: never)
__device__ __forceinline__ uint32_t seg_newlines(const uint8_t* raw, uint32_t off, uint32_t len, uint4& w) {
    w = make_uint4(0, 0, 0, 0);
    if (off >= len) return 0;
    w = *reinterpret_cast<const uint4*>(raw + off);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        m |= ((((__vcmpeq4(ws[i], 0x0a0a0a0au) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * i);
    if (len - off < 16) m &= (1u << (len - off)) - 1u;
    return m;
}

__global__ void __launch_bounds__(kCtaThreads) fastq_count_lines_kernel(const uint8_t* raw, uint32_t len, uint32_t* tile_count) {
    __shared__ uint32_t ws[kCtaThreads / 32];
    uint4 w;
    const uint32_t m = seg_newlines(raw, blockIdx.x * kTileBytes + threadIdx.x * kSegBytes, len, w);
    uint32_t total;
    cta_exclusive_sum(__popc(m), ws, total);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}

// One CTA: tile_base[t] = lines that start before tile t; blk = {lines in the block, bad flag, read bases} reset.
__global__ void __launch_bounds__(1024) fastq_scan_tiles_kernel(const uint32_t* tile_count, uint32_t ntiles, uint32_t* tile_base,
                                                              uint32_t max_lines, FastqBlockState* blk) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (ntiles + 1023) / 1024, t0 = threadIdx.x * per;
    uint32_t sum = 0;
    for (uint32_t i = 0; i < per && t0 + i < ntiles; ++i) sum += tile_count[t0 + i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint32_t v = part[i];
            part[i] = run;
            run += v;
        }
        blk->nlines = run;
        blk->bad = run > max_lines ? 1u : 0u;  // more lines than nlpos holds: not FASTQ as we know it
        blk->read_bases = 0;
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t i = 0; i < per && t0 + i < ntiles; ++i) {
        tile_base[t0 + i] = run;
        run += tile_count[t0 + i];
    }
}

// masked[i] = raw[i] on sequence lines (line index % 4 == 1), '\n' elsewhere; nlpos[line] = where it ends.
__global__ void __launch_bounds__(kCtaThreads) fastq_mask_kernel(const uint8_t* raw, uint32_t len, const uint32_t* tile_base,
                                                               uint8_t* masked, uint32_t* nlpos, uint32_t max_lines,
                                                               FastqBlockState* blk) {
    __shared__ uint32_t ws[kCtaThreads / 32];
    const uint32_t off = blockIdx.x * kTileBytes + threadIdx.x * kSegBytes;
    uint4 w;
    const uint32_t m = seg_newlines(raw, off, len, w);
    uint32_t total;
    uint32_t line = tile_base[blockIdx.x] + cta_exclusive_sum(__popc(m), ws, total);
    const uint32_t in[4] = {w.x, w.y, w.z, w.w};
    uint32_t out[4];
    bool nul = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = 4 * i + j;
            const uint32_t ch = (in[i] >> (8 * j)) & 0xffu;
            const bool live = off + b < len;
            nul |= live && ch == 0;
            o |= ((live && (line & 3u) == 1u) ? ch : (uint32_t)'\n') << (8 * j);
            if ((m >> b) & 1u) {
                if (line < max_lines) nlpos[line] = off + b;
                ++line;
            }
        }
        out[i] = o;
    }
    *reinterpret_cast<uint4*>(masked + off) = make_uint4(out[0], out[1], out[2], out[3]);
    if (nul) atomicOr(&blk->bad, 1u);  // the reference cuts a sequence at a NUL byte: leave that to the host parser
}

// One thread per line: is every record what kseq reads as header / sequence / '+' / quality of equal length?
__global__ void fastq_validate_kernel(const uint8_t* raw, const uint32_t* nlpos, uint32_t max_lines, FastqBlockState* blk) {
    const uint32_t nlines = blk->nlines;
    if (nlines > max_lines) return;  // already marked bad
    uint32_t bases = 0;
    bool bad = false;
    const uint32_t L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L < nlines) {
        auto line_len = [&](uint32_t l, uint32_t& start) {  // kseq's getline: one trailing CR goes, unless the line is just "\r"
            const uint32_t p = nlpos[l];
            start = l ? nlpos[l - 1] + 1 : 0;
            const uint32_t n = p - start;
            return (n > 1 && raw[p - 1] == '\r') ? n - 1 : n;
        };
        uint32_t s;
        const uint32_t n = line_len(L, s);
        const uint32_t raw_n = nlpos[L] - s;
        const uint8_t c0 = raw_n ? raw[s] : 0;
        switch (L & 3u) {
            case 0: bad = c0 != '@'; break;
            case 1: bad = raw_n && (c0 == '@' || c0 == '+' || c0 == '>'); bases = n; break;
            case 2: bad = c0 != '+'; break;
            default: {
                uint32_t s2;
                bad = n != line_len(L - 2, s2);
            }
        }
    }
    if (L == 0 && (nlines & 3u)) bad = true;  // the block does not end on a record boundary
    if (__any_sync(kFullMask, bad) && (threadIdx.x & 31) == 0) atomicOr(&blk->bad, 1u);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bases += __shfl_xor_sync(kFullMask, bases, d);
    if ((threadIdx.x & 31) == 0 && bases) atomicAdd(&blk->read_bases, (unsigned long long)bases);
}

__global__ void fastq_commit_kernel(const FastqBlockState* blk, FastqFileState* file, uint32_t block_no) {
    if (blk->bad || file->bad) {
        if (!file->bad) {
            file->bad = 1;
            file->first_bad_block = block_no;
        }
    } else {
        file->read_bases += blk->read_bases;
        file->blocks_ok += 1;
    }
}

// A chunk a HOST worker stripped (hybrid road): its bases count only while no earlier block of the file was refused
// on the device -- the chunk's own count kernel was skipped by the same flag (Chunk::skip), in the same stream order.
__global__ void fastq_strip_commit_kernel(FastqFileState* file, unsigned long long bases, unsigned int whole_block) {
    if (file->bad) return;
    file->read_bases += bases;
    file->blocks_ok += whole_block;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline Chunk make_chunk(const uint8_t* p, uint64_t nbytes, const unsigned int* skip = nullptr) {
    uintptr_t a = (uintptr_t)p;
    uintptr_t al = a & ~(uintptr_t)15;
    Chunk c;
    c.al = (const uint8_t*)al;
    c.lo = (int64_t)(a - al);
    c.hi = c.lo + (int64_t)nbytes;
    c.skip = skip;
    return c;
}
static inline int64_t tiles_for(const Chunk& c) { return (c.hi + kTileBytes - 1) / kTileBytes; }
static inline unsigned grid_1d(uint64_t n, unsigned block, unsigned cap) {
    uint64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    return (unsigned)(g > cap ? cap : g);
}

// Tuning knob (not a fallback: every variant is the same CUDA path): VG_COUNT_BATCH=4|8 picks how
// many sector loads each lane keeps in flight (and with it registers / CTAs per SM).
int count_variant() {
    static int v = [] {
        const char* e = getenv("VG_COUNT_BATCH");
        int x = e ? atoi(e) : 4;
        return x == 8 ? 8 : 4;
    }();
    return v;
}

// A/B knob: VG_EVEN_WINDOW=0 runs even k through the byte-wise state machine (round 1's only even-k path) instead of
// the window encoder with the palindrome rule.  Same results either way.
static bool even_window() {
#ifdef VG_ROLLING_ENCODER
    return false;
#else
    static bool v = [] {
        const char* e = getenv("VG_EVEN_WINDOW");
        return !(e && atoi(e) == 0);
    }();
    return v;
#endif
}

int sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    return n;
}

cudaError_t launch_table_fill_empty(uint64_t* slots, uint64_t nslots, cudaStream_t s) {
    fill_empty_kernel<<<grid_1d(nslots, 256, 148 * 16), 256, 0, s>>>(slots, nslots);
    return cudaGetLastError();
}

cudaError_t launch_unhash(uint64_t* d_key56, uint64_t n, uint64_t mask, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    unhash_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_key56, n, mask);
    return cudaGetLastError();
}

cudaError_t launch_keys_to_key56(const uint64_t* d_keys, uint64_t n, uint32_t k, uint64_t mask, uint64_t* d_key56,
                                 unsigned long long* d_bad, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    keys_to_key56_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_keys, n, k, mask, d_key56, d_bad);
    return cudaGetLastError();
}

cudaError_t launch_insert(const IndexView& ix, const uint64_t* d_key56, uint64_t n, InsertReport* d_rep,
                          cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    uint64_t g = (n + 255) / 256;
    insert_kernel<<<(unsigned)g, 256, 0, s>>>(ix, d_key56, n, d_rep);
    return cudaGetLastError();
}

cudaError_t launch_select_owned(const IndexView& ix, const uint64_t* d_key56, uint64_t n, uint64_t first_idx,
                                uint64_t* d_own, uint64_t* d_own_idx, unsigned long long* d_n_own, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    select_owned_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ix, d_key56, n, first_idx, d_own, d_own_idx, d_n_own);
    return cudaGetLastError();
}

cudaError_t launch_peer_barrier(const PeerPtrs& flags, int world, int rank, unsigned long long epoch,
                                unsigned long long timeout_ns, unsigned int* d_timeout, cudaStream_t s) {
    peer_barrier_kernel<<<1, 32, 0, s>>>(flags, world, rank, epoch, timeout_ns, d_timeout);
    return cudaGetLastError();
}

cudaError_t launch_publish_counts(const PartView& pv, cudaStream_t s) {
    publish_counts_kernel<<<(pv.P + 255) / 256, 256, 0, s>>>(pv);
    return cudaGetLastError();
}

cudaError_t launch_combine_counts(const PeerPtrs& counts, int world, uint64_t n, uint8_t* d_out, int nsm, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    combine_counts_kernel<<<grid_1d(n / 16 + 1, 256, (unsigned)nsm * 8), 256, 0, s>>>(counts, world, n, d_out);
    return cudaGetLastError();
}

cudaError_t launch_reduce_segment(const PeerPtrs& vecs, int world, int rank, uint64_t seg_bytes, uint64_t nbytes, int nsm, cudaStream_t s) {
    const uint64_t seg0 = std::min<uint64_t>(nbytes, seg_bytes * rank);
    const uint64_t seg1 = rank == world - 1 ? nbytes : std::min<uint64_t>(nbytes, seg_bytes * (rank + 1));
    if (seg1 <= seg0) return cudaSuccess;
    reduce_segment_kernel<<<grid_1d((seg1 - seg0) / 16 + 1, 256, (unsigned)nsm * 8), 256, 0, s>>>(vecs, world, rank, seg0, seg1);
    return cudaGetLastError();
}
cudaError_t launch_gather_segments(const PeerPtrs& vecs, int world, int rank, uint64_t seg_bytes, uint64_t nbytes, int nsm, cudaStream_t s) {
    if (nbytes == 0) return cudaSuccess;
    gather_segments_kernel<<<grid_1d(nbytes / 16 + 1, 256, (unsigned)nsm * 8), 256, 0, s>>>(vecs, world, rank, seg_bytes, nbytes);
    return cudaGetLastError();
}

cudaError_t launch_clear_counts(const IndexView& ix, cudaStream_t s) {
    uint64_t nslots = 4ull * ix.nbuckets;
    clear_counts_kernel<<<grid_1d(nslots, 256, 148 * 16), 256, 0, s>>>(ix.slots, nslots);
    return cudaGetLastError();
}

cudaError_t launch_rank_scan(const IndexView& ix, uint32_t* d_block_sums, unsigned long long* d_total, cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)(((uint64_t)ix.nbuckets + kScanChunk - 1) / kScanChunk);
    occ_block_sums_kernel<<<nblocks, kCtaThreads, 0, s>>>(ix, d_block_sums);
    scan_block_sums_kernel<<<1, 1024, 0, s>>>(d_block_sums, nblocks, d_total);
    rank_base_kernel<<<nblocks, kCtaThreads, 0, s>>>(ix, d_block_sums);
    return cudaGetLastError();
}

cudaError_t launch_slot_perm(const IndexView& ix, const uint64_t* d_key56, uint64_t n, uint32_t* d_perm, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    slot_perm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ix, d_key56, n, d_perm);
    return cudaGetLastError();
}

cudaError_t launch_gather_counts(const uint8_t* cvec, const uint32_t* d_perm, const uint64_t* d_idx, uint64_t n, void* d_out,
                                 int out_elem_bytes, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned g = (unsigned)((n + 255) / 256);
    if (out_elem_bytes == 1) gather_counts_kernel<uint8_t><<<g, 256, 0, s>>>(cvec, d_perm, d_idx, n, (uint8_t*)d_out);
    else if (out_elem_bytes == 4) gather_counts_kernel<uint32_t><<<g, 256, 0, s>>>(cvec, d_perm, d_idx, n, (uint32_t*)d_out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_scatter_bytes(const uint8_t* d_in, const uint32_t* d_perm, uint64_t n, uint8_t* d_out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    scatter_bytes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_in, d_perm, n, d_out);
    return cudaGetLastError();
}

cudaError_t launch_count(const IndexView& ix, const uint8_t* d_bases, uint64_t nbytes, CountStats* d_stats,
                         int ctas_per_sm, int nsm, cudaStream_t s, const unsigned int* d_skip) {
    if (nbytes == 0) return cudaSuccess;
    Chunk c = make_chunk(d_bases, nbytes, d_skip);
    int64_t ntiles = tiles_for(c);
    // persistent grid: exactly the CTAs that are resident at once, each striding over the tiles
    using KernelT = void (*)(IndexView, Chunk, int64_t, CountStats*);
    KernelT kern = (ix.k & 1) ? (count_variant() == 4 ? (KernelT)count_kernel<kEncOdd, 4> : (KernelT)count_kernel<kEncOdd, 8>)
                   : even_window() ? (KernelT)count_kernel<kEncEven, 4>
                                   : (KernelT)count_kernel<kEncAny, 4>;
    int occ = 0;
    if (ctas_per_sm <= 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCtaThreads, 0) != cudaSuccess || occ < 1) occ = 2;
    } else {
        occ = ctas_per_sm;
    }
    int64_t grid = (int64_t)nsm * occ;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, kCtaThreads, 0, s>>>(ix, c, ntiles, d_stats);
    return cudaGetLastError();
}

int64_t chunk_tiles(const uint8_t* d_bases, uint64_t nbytes) { return tiles_for(make_chunk(d_bases, nbytes)); }

cudaError_t launch_prefilter_build(uint32_t* words, uint32_t nwords, const uint64_t* d_key56, uint64_t n, uint32_t k, uint32_t span,
                                   cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    prefilter_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(words, nwords, d_key56, n, k, span);
    return cudaGetLastError();
}

static cudaError_t ensure_magic();
cudaError_t launch_scatter(const IndexView& ix, const PartView& pv, const PrefilterView& pf, const uint8_t* d_bases,
                           uint64_t nbytes, int64_t first_tile, int64_t ntiles, CountStats* d_stats, int nsm,
                           cudaStream_t s, const unsigned int* d_skip) {
    if (ntiles <= 0) return cudaSuccess;
    Chunk c = make_chunk(d_bases, nbytes, d_skip);
    using KernelT = void (*)(IndexView, PartView, PrefilterView, ScatterCfg, Chunk, int64_t, int64_t, CountStats*);
    // VG_SCATTER_TMA=1: the tile load as a bulk async copy into shared memory, a tile ahead (A/B knob; odd k)
    const char* tma_env = getenv("VG_SCATTER_TMA");
    const bool tma = tma_env && atoi(tma_env) != 0 && (ix.k & 1);
    const bool s8 = pf.words && pf.span == 8;
    KernelT kern;
    if (!(ix.k & 1) && even_window()) kern = s8 ? (KernelT)scatter_kernel<kEncEven, false, 8, 0> : (KernelT)scatter_kernel<kEncEven, false, 4, 0>;
    else if (!(ix.k & 1)) kern = s8 ? (KernelT)scatter_kernel<kEncAny, false, 8, 0> : (KernelT)scatter_kernel<kEncAny, false, 4, 0>;
    else if (tma) kern = s8 ? (KernelT)scatter_kernel<kEncOdd, true, 8, 0> : (KernelT)scatter_kernel<kEncOdd, true, 4, 0>;
    else if (ix.k == 27) kern = s8 ? (KernelT)scatter_kernel<kEncOdd, false, 8, 27> : (KernelT)scatter_kernel<kEncOdd, false, 4, 27>;
    else kern = s8 ? (KernelT)scatter_kernel<kEncOdd, false, 8, 0> : (KernelT)scatter_kernel<kEncOdd, false, 4, 0>;
    // Bin capacity: ~1.8x the expected k-mers per slice and tile (the pre-filter passes roughly half),
    // shrunk to a shared-memory budget that lets four CTAs share an SM -- or two, when there are so many
    // slices that four would leave bins of a handful of keys.  Overflowing keys take the key-by-key
    // path, so this only tunes speed.
    static const double capx = [] { const char* e = getenv("VG_SCATTER_CAPX"); return e ? atof(e) : 1.8; }();
    static const size_t budget_kb = [] { const char* e = getenv("VG_SCATTER_SMEM_KB"); return (size_t)(e ? atoi(e) : 52); }();
    const size_t budget = (budget_kb - (tma ? 9 : 0)) * 1024;  // the staged tiles take 8.3 KB of a CTA's share
    const double expect = (pf.words ? 0.6 : 1.0) * kTileBytes / (double)pv.P;
    const uint32_t want = (uint32_t)(expect * capx) + 8;
    auto bytes = [&](uint32_t cp) { return (size_t)pv.P * ((size_t)(cp | 1u) * 8 + 20) + 16; };
    {
        cudaError_t e = ensure_magic();
        if (e != cudaSuccess) return e;
    }
    uint32_t cap = want > kMaxBinCap ? kMaxBinCap : want;
    while (cap > 4 && bytes(cap) > budget) --cap;
    if (cap < want && cap < 24) {  // many slices: two CTAs per SM with deeper bins
        cap = want > kMaxBinCap ? kMaxBinCap : want;
        while (cap > 4 && bytes(cap) > 2 * budget) --cap;
    }
    ScatterCfg cfg{cap, cap | 1u};
    const size_t smem = bytes(cap);
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCtaThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    int64_t grid = (int64_t)nsm * occ;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, kCtaThreads, smem, s>>>(ix, pv, pf, cfg, c, first_tile, ntiles, d_stats);
    return cudaGetLastError();
}

static cudaError_t ensure_magic() {  // c_magic lives in constant memory, which is per device
    static bool magic_on[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& ready = magic_on[dev & 63];
    if (ready) return cudaSuccess;
    static uint32_t m[kMaxBinCap + 1];
    m[0] = 0;
    for (uint32_t d = 1; d <= kMaxBinCap; ++d) m[d] = (uint32_t)((1ull << 32) / d) + 1u;
    cudaError_t e = cudaMemcpyToSymbol(c_magic, m, sizeof m);
    if (e == cudaSuccess) ready = true;
    return e;
}

cudaError_t launch_probe_partitions(const IndexView& ix, const PartView& pv, const uint32_t* h_slice_rank, CountStats* d_stats,
                                    int nsm, cudaStream_t s) {
    const bool b4 = count_variant() == 4;
    int occ = 0;
    auto slice = [&](uint32_t p, uint32_t& b0, uint32_t& b1) {  // local slice p of THIS table
        b0 = p << pv.shift2;
        const uint64_t e = ((uint64_t)(p + 1)) << pv.shift2;
        b1 = (uint32_t)(e > ix.nbuckets ? ix.nbuckets : e);
    };
    const size_t ctr_elems = (size_t)4 << pv.shift2;
    const bool sharded = pv.world > 1;
    const uint32_t nsub = sharded ? pv.world : 1u;
    // sweep: launch p probes slice p and retires slice p-1; one extra launch retires the last slice.
    // A sharded index has one key list per source GPU and slice: the kMulti kernel walks them all.
    auto launch = [&](auto kern, const uint64_t* lists, const unsigned long long* cnt, uint32_t nlists, uint64_t cap, uint32_t b0,
                      uint32_t b1, uint32_t* cur, uint32_t r0, uint32_t r1, uint32_t* prv, uint32_t f0, uint32_t f1, uint32_t c0,
                      uint32_t c1) {
        if (!occ && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCtaThreads, 0) != cudaSuccess || occ < 1)) occ = 2;
        kern<<<(unsigned)(nsm * occ), kCtaThreads, 0, s>>>(ix, lists, cnt, nlists, pv.P_local, cap, b0, b1, cur, r0, r1, prv, f0, f1, c0, c1,
                                                           d_stats);
    };
    const uint32_t nslices = (uint32_t)(((uint64_t)ix.nbuckets + ((1ull << pv.shift2) - 1)) >> pv.shift2);
    // VG_PREFETCH_AHEAD=1: launch p pulls slice p+1 into L2 while it probes slice p (the first launch of a group
    // pulls its own slice as well); default: every launch pulls the slice it probes.
    const char* ahead_env = getenv("VG_PREFETCH_AHEAD");
    const bool ahead = ahead_env && atoi(ahead_env) != 0;
    auto probe_step = [&](uint32_t p, const uint64_t* lists, const unsigned long long* cnt, bool multi, uint64_t cap, bool first_of_group) {
        uint32_t b0 = 0, b1 = 0, r0 = 0, r1 = 0, f0 = 0, f1 = 0;
        if (p < nslices) slice(p, b0, b1);
        if (p > 0) slice(p - 1, r0, r1);
        if (!ahead) {
            f0 = b0, f1 = b1;
        } else if (p < nslices) {
            uint32_t n0 = b1, n1 = b1;
            if (p + 1 < nslices) slice(p + 1, n0, n1);
            f0 = first_of_group ? b0 : n0;
            f1 = n1;
        }
        uint32_t* cur = pv.ctr + (size_t)(p & 1) * ctr_elems;
        uint32_t* prv = pv.ctr + (size_t)((p + 1) & 1) * ctr_elems;
        const uint32_t c0 = (h_slice_rank && p < nslices) ? h_slice_rank[p] : 0u, c1 = (h_slice_rank && p < nslices) ? h_slice_rank[p + 1] : 0u;
        if (multi) {
            if (b4) launch(probe_slice_kernel<4, true>, lists, cnt, nsub, cap, b0, b1, cur, r0, r1, prv, f0, f1, c0, c1);
            else launch(probe_slice_kernel<8, true>, lists, cnt, nsub, cap, b0, b1, cur, r0, r1, prv, f0, f1, c0, c1);
        } else {
            if (b4) launch(probe_slice_kernel<4, false>, lists, cnt, 1u, cap, b0, b1, cur, r0, r1, prv, f0, f1, c0, c1);
            else launch(probe_slice_kernel<8, false>, lists, cnt, 1u, cap, b0, b1, cur, r0, r1, prv, f0, f1, c0, c1);
        }
    };
    if (pv.sub_bits == 0) {
        for (uint32_t p = 0; p <= pv.P_local; ++p) {
            const uint32_t q = p < pv.P_local ? p : 0;
            probe_step(p, pv.keybuf + (uint64_t)q * nsub * pv.cap, (sharded ? pv.incount : pv.cursor) + q, sharded, pv.cap, p == 0);
        }
    } else {
        // two levels: coarse list c -> the lists of its Q slices (rescatter), then those slices are probed
        cudaError_t e = ensure_magic();
        if (e != cudaSuccess) return e;
        const uint32_t Q = 1u << pv.sub_bits;
        const uint32_t want = (uint32_t)(1.6 * (kCtaThreads * 16) / Q) + 8;
        auto bytes = [&](uint32_t cp) { return (size_t)Q * ((size_t)(cp | 1u) * 8 + 20) + 16; };
        uint32_t cap = want > kMaxBinCap ? kMaxBinCap : want;
        while (cap > 4 && bytes(cap) > (size_t)52 * 1024) --cap;
        const ScatterCfg cfg{cap, cap | 1u};
        const size_t smem = bytes(cap);
        using RK = void (*)(IndexView, PartView, ScatterCfg, const uint64_t*, const unsigned long long*, uint32_t, uint32_t, uint32_t,
                            uint32_t, CountStats*);
        RK rk = sharded ? (RK)rescatter_kernel<true> : (RK)rescatter_kernel<false>;
        e = cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int rocc = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rocc, rk, kCtaThreads, smem) != cudaSuccess || rocc < 1) rocc = 1;
        uint32_t p = 0;
        for (uint32_t c = 0; c < pv.P_local && p < nslices; ++c) {
            e = cudaMemsetAsync(pv.cursor2, 0, Q * sizeof(unsigned long long), s);
            if (e != cudaSuccess) return e;
            rk<<<(unsigned)(nsm * rocc), kCtaThreads, smem, s>>>(ix, pv, cfg, pv.keybuf + (uint64_t)c * nsub * pv.cap,
                                                               (sharded ? pv.incount : pv.cursor) + c, nsub, pv.P_local, c << pv.sub_bits, Q,
                                                               d_stats);
            for (uint32_t q = 0; q < Q && p < nslices; ++q, ++p)
                probe_step(p, pv.keybuf2 + (uint64_t)q * pv.cap2, pv.cursor2 + q, false, pv.cap2, q == 0);
        }
        probe_step(nslices, pv.keybuf2, pv.cursor2, false, pv.cap2, false);  // retires the last slice, probes nothing
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || sharded) return e;  // sharded: publish_counts re-armed the cursors already
    return cudaMemsetAsync(pv.cursor, 0, pv.P * sizeof(unsigned long long), s);
}

__global__ void sum_cursors_kernel(const unsigned long long* cursor, uint32_t P, uint64_t cap, unsigned long long* total, CountStats* stats) {
    __shared__ unsigned long long sh[32];
    unsigned long long v = 0;
    for (uint32_t p = threadIdx.x; p < P; p += blockDim.x) v += min((unsigned long long)cap, cursor[p]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w) t += sh[w];
        *total = t;
        stats->keys += t;
    }
}
cudaError_t launch_sum_cursors(const unsigned long long* cursor, uint32_t P, uint64_t cap, unsigned long long* d_total, CountStats* d_stats,
                               cudaStream_t s) {
    sum_cursors_kernel<<<1, 256, 0, s>>>(cursor, P, cap, d_total, d_stats);
    return cudaGetLastError();
}

uint64_t sweep_launches(const IndexView& ix, const PartView& pv) {
    const uint64_t nslices = ((uint64_t)ix.nbuckets + ((1ull << pv.shift2) - 1)) >> pv.shift2;
    return nslices + 1 + (pv.sub_bits ? pv.P_local : 0);
}

cudaError_t launch_extract(const IndexView& ix, const uint64_t* d_key56, const uint64_t* d_idx, uint64_t n, void* d_out,
                           int out_elem_bytes, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    unsigned g = (unsigned)((n + 255) / 256);
    if (out_elem_bytes == 1) extract_kernel<uint8_t><<<g, 256, 0, s>>>(ix, d_key56, d_idx, n, (uint8_t*)d_out);
    else if (out_elem_bytes == 4) extract_kernel<uint32_t><<<g, 256, 0, s>>>(ix, d_key56, d_idx, n, (uint32_t*)d_out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_histogram(const uint8_t* d_counts, const uint8_t* d_flags, uint64_t n, unsigned long long* d_hist,
                             cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), s);
    if (e != cudaSuccess || n == 0) return e;
    histogram_kernel<<<grid_1d(n, 256, 148 * 8), 256, 0, s>>>(d_counts, d_flags, n, d_hist);
    return cudaGetLastError();
}

cudaError_t launch_fastq_block(const uint8_t* d_raw, uint32_t len, uint8_t* d_masked, const FastqScratch& sc,
                               FastqFileState* d_file, uint32_t block_no, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    const uint32_t ntiles = (len + kTileBytes - 1) / kTileBytes;
    if (ntiles > sc.max_tiles) return cudaErrorInvalidValue;
    fastq_count_lines_kernel<<<ntiles, kCtaThreads, 0, s>>>(d_raw, len, sc.tile_count);
    fastq_scan_tiles_kernel<<<1, 1024, 0, s>>>(sc.tile_count, ntiles, sc.tile_base, sc.max_lines, sc.blk);
    fastq_mask_kernel<<<ntiles, kCtaThreads, 0, s>>>(d_raw, len, sc.tile_base, d_masked, sc.nlpos, sc.max_lines, sc.blk);
    // one thread per possible line; the kernel reads the true number from blk
    const uint32_t lines_cap = len < sc.max_lines ? len : sc.max_lines;
    fastq_validate_kernel<<<(lines_cap + 255) / 256, 256, 0, s>>>(d_raw, sc.nlpos, sc.max_lines, sc.blk);
    fastq_commit_kernel<<<1, 1, 0, s>>>(sc.blk, d_file, block_no);
    return cudaGetLastError();
}

cudaError_t launch_fastq_strip_commit(FastqFileState* d_file, unsigned long long bases, bool whole_block, cudaStream_t s) {
    fastq_strip_commit_kernel<<<1, 1, 0, s>>>(d_file, bases, whole_block ? 1u : 0u);
    return cudaGetLastError();
}

cudaError_t launch_positions(uint32_t k, const uint8_t* d_bases, uint64_t nbytes, uint64_t* d_out,
                             cudaStream_t s) {
    if (nbytes == 0) return cudaSuccess;
    Chunk c = make_chunk(d_bases, nbytes);
    int64_t ntiles = tiles_for(c);
    KmerParams kp{k, (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1)};
    int64_t grid = ntiles < 148 * 8 ? ntiles : 148 * 8;
    if (k & 1) positions_kernel<kEncOdd><<<(unsigned)grid, kCtaThreads, 0, s>>>(kp, c, ntiles, d_out);
    else if (even_window()) positions_kernel<kEncEven><<<(unsigned)grid, kCtaThreads, 0, s>>>(kp, c, ntiles, d_out);
    else positions_kernel<kEncAny><<<(unsigned)grid, kCtaThreads, 0, s>>>(kp, c, ntiles, d_out);
    return cudaGetLastError();
}

cudaError_t launch_cbf_add(const CbfView& cbf, uint32_t k, const uint8_t* d_bases, uint64_t hi, uint64_t own_from,
                           unsigned long long* d_added, int nsm, cudaStream_t s) {
    if (hi <= own_from) return cudaSuccess;
    Chunk c = make_chunk(d_bases, hi);  // d_bases is 16-byte aligned by contract (cudaMalloc)
    int64_t first_tile = (int64_t)(own_from / kTileBytes);
    int64_t ntiles = tiles_for(c) - first_tile;
    KmerParams kp{k, (1ULL << (2 * k)) - 1};
    int64_t grid = (int64_t)nsm * 8;
    if (grid > ntiles) grid = ntiles;
    if (k & 1) cbf_add_kernel<kEncOdd><<<(unsigned)grid, kCtaThreads, 0, s>>>(cbf, kp, c, first_tile, ntiles, d_added);
    else if (even_window()) cbf_add_kernel<kEncEven><<<(unsigned)grid, kCtaThreads, 0, s>>>(cbf, kp, c, first_tile, ntiles, d_added);
    else cbf_add_kernel<kEncAny><<<(unsigned)grid, kCtaThreads, 0, s>>>(cbf, kp, c, first_tile, ntiles, d_added);
    return cudaGetLastError();
}

cudaError_t launch_random_sectors(const uint64_t* table, uint32_t nbuckets, uint32_t rounds, int grid,
                                  unsigned long long* sink, cudaStream_t s) {
    random_sector_kernel<<<(unsigned)grid, kCtaThreads, 0, s>>>(table, nbuckets, rounds, sink);
    return cudaGetLastError();
}
int probe_batch() { return kProbeBatch; }

cudaError_t launch_cbf_query(const CbfView& cbf, const uint64_t* d_keys, uint64_t n, uint8_t* d_count,
                             uint8_t* d_find, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    cbf_query_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cbf, d_keys, n, d_count, d_find);
    return cudaGetLastError();
}

}  // namespace vg
