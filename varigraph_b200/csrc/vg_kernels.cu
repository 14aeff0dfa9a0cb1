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
#include <cstdlib>

#include "vg_device.cuh"
#include "vg_internal.h"

namespace vg {

// ---------------------------------------------------------------------------
// index build
// ---------------------------------------------------------------------------
__global__ void fill_empty_kernel(uint64_t* slots, uint64_t nslots) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nslots; i += stride) slots[i] = kSlotEmpty;
}

// One thread per key.  Claims the first empty slot in probe order with a 64-bit CAS; because
// slots are never freed, "an empty slot ends the search" holds for every later lookup.
__global__ void insert_kernel(IndexView ix, const uint64_t* __restrict__ key56, uint64_t n, InsertReport* rep) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = key56[i];
    if (key == kKey56Max) return;  // lives in ix.special
    uint64_t want = key << 8;
    uint32_t b = bucket_of(key, ix.nbuckets);
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

__global__ void clear_counts_kernel(uint64_t* slots, uint64_t nslots) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nslots; i += stride) {
        uint64_t v = slots[i];
        if (v != kSlotEmpty && (v & 0xffu)) slots[i] = v & ~0xffULL;
    }
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

// All 32 lanes call this together (ballot / match inside).
// K1 hands over 8 keys and their emit mask; K2 probes one 32-byte bucket per key with 8 sector
// loads in flight per lane; K3 adds to the 8-bit counter in the matched slot.  The CAS of a
// position is issued as soon as its slot is matched but its result is only checked after the
// whole batch, so the round trips overlap instead of serialising.
// meta[b]: bit 31 CAS issued, bits 16-23 count seen, bits 8-9 slot in bucket, bits 0-7 amount.
template <bool kK28, int kBatch>
__device__ __forceinline__ void probe_and_count(const IndexView& ix, const uint64_t (&keys)[kBatch], uint32_t emit,
                                                uint32_t& n_hit) {
    constexpr int kProbeBatch = kBatch;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t bk[kProbeBatch], meta[kProbeBatch];
    uint64_t prev[kProbeBatch];
    uint32_t havem = emit & ((1u << kProbeBatch) - 1);
    if (kK28) {
#pragma unroll
        for (int b = 0; b < kProbeBatch; ++b) {
            if (((havem >> b) & 1u) && keys[b] == kKey56Max) {  // the hash no slot can hold
                havem &= ~(1u << b);
                if (ix.has_special) {
                    atomicAdd(ix.special, 1ull);
                    n_hit += 1;
                }
            }
        }
    }
    {
        uint64_t v[kProbeBatch][4];
#pragma unroll
        for (int b = 0; b < kProbeBatch; ++b) {
            bk[b] = bucket_of(keys[b], ix.nbuckets);
            if ((havem >> b) & 1u) ld_bucket(ix.slots + 4ull * bk[b], v[b]);
        }
#pragma unroll
        for (int b = 0; b < kProbeBatch; ++b) {
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
            const bool not_full = v[b][3] == kSlotEmpty;
            bool hit = have && (m0 || m1 || m2 || m3);
            uint32_t hs = m1 ? 1u : (m2 ? 2u : (m3 ? 3u : 0u));
            uint32_t cnt = (m1 ? (uint32_t)v[b][1] : (m2 ? (uint32_t)v[b][2] : (m3 ? (uint32_t)v[b][3] : (uint32_t)v[b][0]))) & 0xffu;
            bool more = have && !hit && !not_full;  // bucket full: the key may have spilled over
            while (__any_sync(kFullMask, more)) {
                if (more) {
                    bk[b] = (bk[b] + 1 == ix.nbuckets) ? 0 : bk[b] + 1;
                    uint64_t w[4];
                    ld_bucket(ix.slots + 4ull * bk[b], w);
                    bool se;
                    uint64_t sv = 0;
                    int h2 = match_slot(w, key, se, sv);
                    if (h2 >= 0) { hit = true; hs = (uint32_t)h2; cnt = (uint32_t)sv & 0xffu; }
                    more = h2 < 0 && !se;
                }
            }
            meta[b] = 0;
            const uint32_t hm = __ballot_sync(kFullMask, hit);
            if (hit) {
                n_hit += 1;
                const uint64_t slot = 4ull * bk[b] + hs;
                const uint32_t peers = (hm & (hm - 1)) ? __match_any_sync(hm, slot) : hm;  // warp-aggregate
                if ((uint32_t)(__ffs(peers) - 1) == lane && cnt != 255u)
                    meta[b] = 0x80000000u | (cnt << 16) | (hs << 8) | (uint32_t)__popc(peers);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < kProbeBatch; ++b) {
        if (meta[b] & 0x80000000u) {
            const uint32_t cnt = (meta[b] >> 16) & 0xffu;
            const uint64_t seen = (keys[b] << 8) | cnt;
            const uint32_t add = min(meta[b] & 0xffu, 255u - cnt);
            prev[b] = atomicCAS((unsigned long long*)(ix.slots + 4ull * bk[b] + ((meta[b] >> 8) & 3u)), seen, seen + add);
        }
    }
    // verify; a lost race (or a key this lane hit twice in the batch) retries here
#pragma unroll
    for (int b = 0; b < kProbeBatch; ++b) {
        if (meta[b] & 0x80000000u) {
            const uint64_t seen = (keys[b] << 8) | ((meta[b] >> 16) & 0xffu);
            if (prev[b] != seen)
                slot_sat_add(ix.slots + 4ull * bk[b] + ((meta[b] >> 8) & 3u), prev[b], meta[b] & 0xffu);
        }
    }
}

template <bool kOdd, bool kK28, int kBatch>
__global__ void __launch_bounds__(kCtaThreads, kBatch >= 8 ? 2 : 3)
count_kernel(IndexView ix, Chunk c, int64_t ntiles, CountStats* stats) {
    __shared__ uint8_t lut[256];
    __shared__ unsigned long long blk[2];
    lut_init(lut);
    if (threadIdx.x < 2) blk[threadIdx.x] = 0;
    __syncthreads();
    KmerParams kp{ix.k, ix.mask};
    uint32_t n_pos = 0, n_hit = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t off = t * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        if (kOdd) {
            OddEncoder enc;
            enc.init(c, off, kp, lut);
            // rolled on purpose: one probe batch of registers, and a loop body that stays in the I-cache
#pragma unroll 1
            for (int part = 0; part < 16 / kBatch; ++part) {
                uint64_t keys[kBatch];
                uint32_t emit = enc.next<kBatch>(kp, keys);
                n_pos += __popc(emit);
                probe_and_count<kK28, kBatch>(ix, keys, emit, n_hit);
            }
        } else {
            uint64_t k16[16];
            uint32_t emit = encode_keys_any(c, off, kp, lut, k16);
            n_pos += __popc(emit);
#pragma unroll
            for (int part = 0; part < 16 / kBatch; ++part) {
                uint64_t keys[kBatch];
#pragma unroll
                for (int j = 0; j < kBatch; ++j) keys[j] = k16[part * kBatch + j];
                probe_and_count<kK28, kBatch>(ix, keys, (emit >> (part * kBatch)) & ((1u << kBatch) - 1), n_hit);
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
// extraction: counts in the key order given at index creation
// ---------------------------------------------------------------------------
template <typename OutT>
__global__ void extract_kernel(IndexView ix, const uint64_t* __restrict__ key56, uint64_t n, OutT* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = key56[i];
    uint32_t c = 0;
    if (key == kKey56Max) {
        unsigned long long s = *ix.special;
        c = s > 255ull ? 255u : (uint32_t)s;
    } else {
        uint32_t b = bucket_of(key, ix.nbuckets);
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
    out[i] = (OutT)c;
}

// ---------------------------------------------------------------------------
// per-position keys (test hook + synthetic index construction for bench.py)
// out[p] = (hash << 8 | k) for the k-mer ENDING at byte p, or ~0.
// ---------------------------------------------------------------------------
template <bool kOdd>
__global__ void __launch_bounds__(kCtaThreads) positions_kernel(KmerParams kp, Chunk c, int64_t ntiles, uint64_t* out) {
    __shared__ uint8_t lut[256];
    lut_init(lut);
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t off = t * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        uint64_t keys[16];
        uint32_t emit = kOdd ? encode_keys_odd(c, off, kp, lut, keys) : encode_keys_any(c, off, kp, lut, keys);
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

template <bool kOdd>
__global__ void __launch_bounds__(kCtaThreads) cbf_add_kernel(CbfView cbf, KmerParams kp, Chunk c, int64_t first_tile,
                                                            int64_t ntiles, unsigned long long* added) {
    __shared__ uint8_t lut[256];
    lut_init(lut);
    FastMod64 fm{cbf.magic_hi, cbf.magic_lo, cbf.m};
    uint32_t n = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t off = (first_tile + t) * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        uint64_t keys[16];
        uint32_t emit = kOdd ? encode_keys_odd(c, off, kp, lut, keys) : encode_keys_any(c, off, kp, lut, keys);
        n += __popc(emit);
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
// launchers
// ---------------------------------------------------------------------------
static inline Chunk make_chunk(const uint8_t* p, uint64_t nbytes) {
    uintptr_t a = (uintptr_t)p;
    uintptr_t al = a & ~(uintptr_t)15;
    Chunk c;
    c.al = (const uint8_t*)al;
    c.lo = (int64_t)(a - al);
    c.hi = c.lo + (int64_t)nbytes;
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
        int x = e ? atoi(e) : 8;
        return x == 4 ? 4 : 8;
    }();
    return v;
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

cudaError_t launch_insert(const IndexView& ix, const uint64_t* d_key56, uint64_t n, InsertReport* d_rep,
                          cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    uint64_t g = (n + 255) / 256;
    insert_kernel<<<(unsigned)g, 256, 0, s>>>(ix, d_key56, n, d_rep);
    return cudaGetLastError();
}

cudaError_t launch_clear_counts(const IndexView& ix, cudaStream_t s) {
    uint64_t nslots = 4ull * ix.nbuckets;
    clear_counts_kernel<<<grid_1d(nslots, 256, 148 * 16), 256, 0, s>>>(ix.slots, nslots);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(ix.special, 0, sizeof(unsigned long long), s);
}

cudaError_t launch_count(const IndexView& ix, const uint8_t* d_bases, uint64_t nbytes, CountStats* d_stats,
                         int ctas_per_sm, int nsm, cudaStream_t s) {
    if (nbytes == 0) return cudaSuccess;
    Chunk c = make_chunk(d_bases, nbytes);
    int64_t ntiles = tiles_for(c);
    // persistent grid: exactly the CTAs that are resident at once, each striding over the tiles
    using KernelT = void (*)(IndexView, Chunk, int64_t, CountStats*);
    KernelT kern = (ix.k & 1) ? (count_variant() == 4 ? (KernelT)count_kernel<true, false, 4> : (KernelT)count_kernel<true, false, 8>)
                   : (ix.k == 28 ? (KernelT)count_kernel<false, true, 4> : (KernelT)count_kernel<false, false, 4>);
    int occ = 0;
    if (ctas_per_sm <= 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCtaThreads, 0) != cudaSuccess || occ < 1) occ = 2;
    } else {
        occ = ctas_per_sm;
    }
    int64_t grid = (int64_t)nsm * occ;
    if (grid > ntiles) grid = ntiles;
    if (ix.k & 1) {
        if (count_variant() == 4) count_kernel<true, false, 4><<<(unsigned)grid, kCtaThreads, 0, s>>>(ix, c, ntiles, d_stats);
        else count_kernel<true, false, 8><<<(unsigned)grid, kCtaThreads, 0, s>>>(ix, c, ntiles, d_stats);
    } else if (ix.k == 28) {
        count_kernel<false, true, 4><<<(unsigned)grid, kCtaThreads, 0, s>>>(ix, c, ntiles, d_stats);
    } else {
        count_kernel<false, false, 4><<<(unsigned)grid, kCtaThreads, 0, s>>>(ix, c, ntiles, d_stats);
    }
    return cudaGetLastError();
}

cudaError_t launch_extract(const IndexView& ix, const uint64_t* d_key56, uint64_t n, void* d_out,
                           int out_elem_bytes, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    unsigned g = (unsigned)((n + 255) / 256);
    if (out_elem_bytes == 1) extract_kernel<uint8_t><<<g, 256, 0, s>>>(ix, d_key56, n, (uint8_t*)d_out);
    else if (out_elem_bytes == 4) extract_kernel<uint32_t><<<g, 256, 0, s>>>(ix, d_key56, n, (uint32_t*)d_out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_positions(uint32_t k, const uint8_t* d_bases, uint64_t nbytes, uint64_t* d_out,
                             cudaStream_t s) {
    if (nbytes == 0) return cudaSuccess;
    Chunk c = make_chunk(d_bases, nbytes);
    int64_t ntiles = tiles_for(c);
    KmerParams kp{k, (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1)};
    int64_t grid = ntiles < 148 * 8 ? ntiles : 148 * 8;
    if (k & 1) positions_kernel<true><<<(unsigned)grid, kCtaThreads, 0, s>>>(kp, c, ntiles, d_out);
    else positions_kernel<false><<<(unsigned)grid, kCtaThreads, 0, s>>>(kp, c, ntiles, d_out);
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
    if (k & 1) cbf_add_kernel<true><<<(unsigned)grid, kCtaThreads, 0, s>>>(cbf, kp, c, first_tile, ntiles, d_added);
    else cbf_add_kernel<false><<<(unsigned)grid, kCtaThreads, 0, s>>>(cbf, kp, c, first_tile, ntiles, d_added);
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
