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

// key must not be kKey56Max (callers peel that one off), so an empty slot never matches.
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

constexpr int kProbeBatch = 8;  // independent 32-byte sector loads in flight per lane

// All 32 lanes call this together (ballot / match inside).
__device__ __forceinline__ void probe_and_count(const IndexView& ix, const uint64_t (&keys)[16],
                                                uint32_t& n_pos, uint32_t& n_hit) {
#pragma unroll
    for (int g = 0; g < 16; g += kProbeBatch) {
        uint64_t v[kProbeBatch][4];
        uint32_t bk[kProbeBatch];
#pragma unroll
        for (int b = 0; b < kProbeBatch; ++b) {
            uint64_t key = keys[g + b];
            bool have = key != kNoKmer && key != kKey56Max;
            bk[b] = bucket_of(key, ix.nbuckets);
            if (have) ld_bucket(ix.slots + 4ull * bk[b], v[b]);
        }
#pragma unroll
        for (int b = 0; b < kProbeBatch; ++b) {
            uint64_t key = keys[g + b];
            bool emitted = key != kNoKmer;
            bool have = emitted && key != kKey56Max;
            n_pos += emitted ? 1u : 0u;
            if (emitted && !have && ix.has_special) {  // k == 28 corner: the all-ones hash
                atomicAdd(ix.special, 1ull);
                n_hit += 1;
            }
            int hs = -1;
            bool saw_empty = false;
            uint64_t seen = 0;
            if (have) hs = match_slot(v[b], key, saw_empty, seen);
            bool more = have && hs < 0 && !saw_empty;  // bucket full, key may have spilled over
            while (__any_sync(kFullMask, more)) {
                if (more) {
                    bk[b] = (bk[b] + 1 == ix.nbuckets) ? 0 : bk[b] + 1;
                    ld_bucket(ix.slots + 4ull * bk[b], v[b]);
                    hs = match_slot(v[b], key, saw_empty, seen);
                    more = hs < 0 && !saw_empty;
                }
            }
            bool hit = hs >= 0;
            uint64_t slot = 4ull * bk[b] + (uint32_t)(hit ? hs : 0);
            uint32_t hm = __ballot_sync(kFullMask, hit);
            if (hit) {
                n_hit += 1;
                uint32_t peers = (hm & (hm - 1)) ? __match_any_sync(hm, slot) : hm;
                if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31))
                    slot_sat_add(ix.slots + slot, seen, (uint32_t)__popc(peers));
            }
        }
    }
}

template <bool kOdd>
__global__ void __launch_bounds__(kCtaThreads) count_kernel(IndexView ix, Chunk c, int64_t ntiles, CountStats* stats) {
    __shared__ uint8_t lut[256];
    __shared__ unsigned long long blk[2];
    lut_init(lut);
    if (threadIdx.x < 2) blk[threadIdx.x] = 0;
    __syncthreads();
    KmerParams kp{ix.k, ix.mask};
    uint32_t n_pos = 0, n_hit = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t off = t * kTileBytes + (int64_t)threadIdx.x * kSegBytes;
        uint64_t keys[16];
        if (kOdd) encode_keys_odd(c, off, kp, lut, keys);
        else encode_keys_any(c, off, kp, lut, keys);
        probe_and_count(ix, keys, n_pos, n_hit);
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
        if (kOdd) encode_keys_odd(c, off, kp, lut, keys);
        else encode_keys_any(c, off, kp, lut, keys);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int64_t p = off + j;
            if (p >= c.lo && p < c.hi)
                out[p - c.lo] = keys[j] == kNoKmer ? kNoKmer : ((keys[j] << 8) | kp.k);
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
        if (kOdd) encode_keys_odd(c, off, kp, lut, keys);
        else encode_keys_any(c, off, kp, lut, keys);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (keys[j] == kNoKmer) continue;
            ++n;
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
    int64_t grid = (int64_t)nsm * ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    if (ix.k & 1) count_kernel<true><<<(unsigned)grid, kCtaThreads, 0, s>>>(ix, c, ntiles, d_stats);
    else count_kernel<false><<<(unsigned)grid, kCtaThreads, 0, s>>>(ix, c, ntiles, d_stats);
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

cudaError_t launch_cbf_query(const CbfView& cbf, const uint64_t* d_keys, uint64_t n, uint8_t* d_count,
                             uint8_t* d_find, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    cbf_query_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cbf, d_keys, n, d_count, d_find);
    return cudaGetLastError();
}

}  // namespace vg
