// vg_device.cuh -- device-side primitives of the read k-mer counting path (sm_100a).
//
// What each piece reproduces (reference file:line, /root/reference):
//   nt4 LUT            include/seq_nt4_table.hpp:5-22   (byte -> 0..3, else ambiguous)
//   hash64             include/hash64.hpp:5-14
//   rolling encoder    src/kmer.cpp:126-146             (fwd/rev registers, run length, fwd==rev skip)
//   murmur3_x64_128    src/MurmurHash3.cpp:255-332 for len == 8, summed as
//                      src/counting_bloom_filter.cpp:90-98
// The layout of the index (one 32-byte sector = 4 slots of [hash:56 | count:8]) is ours.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vg {

constexpr uint64_t kNoKmer = ~0ULL;                // "this position emits nothing"
constexpr uint64_t kSlotEmpty = ~0ULL;             // never equals a live slot (see kKey56Max)
constexpr uint64_t kKey56Max = (1ULL << 56) - 1;   // only reachable for k == 28: kept outside the table
constexpr uint32_t kFullMask = 0xffffffffu;
constexpr int kSegBytes = 16;                      // bytes per lane per step (one 128-bit load)
constexpr int kCtaThreads = 256;
constexpr int kTileBytes = kCtaThreads * kSegBytes;  // 4 KiB of bases per CTA step

// LUT entry: bits 0-1 code, bit 2 valid, bit 3 newline (hard read boundary).
__device__ __forceinline__ uint8_t nt4_entry(uint32_t b) {
    uint32_t code = 0, valid = 0;
    uint32_t up = b & 0xDFu;
    if (b < 4) { code = b; valid = 1; }
    else if (up == 'A') { code = 0; valid = 1; }
    else if (up == 'C') { code = 1; valid = 1; }
    else if (up == 'G') { code = 2; valid = 1; }
    else if (up == 'T' || up == 'U') { code = 3; valid = 1; }
    uint32_t nl = (b == '\n') ? 1u : 0u;
    return (uint8_t)(code | (valid << 2) | (nl << 3));
}

// Shared-memory tables of a CTA: lut[0..255] the byte table above, then four 16-bit tables, one per byte
// position j of a 32-bit word of text (j = 0 is the first base): code << 2(3-j) | valid << (8 + 3-j).  OR-ing
// the four lookups of a word yields its four 2-bit codes (first base highest) in bits 0-7 and its four
// validity bits in bits 8-11, already in place -- no per-base shifting.  hard_flags (the even-k window encoder):
// bits 12-15 likewise say which of the four bytes are newlines.
constexpr int kLutBytes = 256 + 4 * 256 * 2;
__device__ __forceinline__ void lut_init(uint8_t* lut, bool hard_flags = false) {
    uint16_t* w4 = reinterpret_cast<uint16_t*>(lut + 256);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint32_t e = nt4_entry((uint32_t)i);
        lut[i] = (uint8_t)e;
        const uint32_t nl = hard_flags ? (e >> 3) & 1u : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            w4[j * 256 + i] = (uint16_t)(((e & 3u) << (2 * (3 - j))) | (((e >> 2) & 1u) << (8 + 3 - j)) | (nl << (12 + 3 - j)));
    }
    __syncthreads();
}

// The shift-add chains of the reference are products by 2^21-1, 265, 21 and 2^31+1 (mod 2^64, then
// masked): written as multiplies they cost two IMADs each instead of four shifts and adds.
__device__ __forceinline__ uint64_t hash64(uint64_t x, uint64_t m) {
    x = (x * 0x1FFFFFULL - 1ULL) & m;  // ~x + (x << 21)
    x ^= x >> 24;
    x = (x * 265ULL) & m;              // x + (x << 3) + (x << 8)
    x ^= x >> 14;
    x = (x * 21ULL) & m;               // x + (x << 2) + (x << 4)
    x ^= x >> 28;
    x = (x * 0x80000001ULL) & m;       // x + (x << 31)
    return x;
}

// Same function for 2k >= 32 (k >= 16): the mask's low word is all ones, so only high words are masked.
__device__ __forceinline__ uint64_t hash64_wide(uint64_t x, uint32_t mask_hi) {
    const uint64_t m = ((uint64_t)mask_hi << 32) | 0xffffffffu;
    x = (x * 0x1FFFFFULL - 1ULL) & m;
    x ^= x >> 24;
    x = (x * 265ULL) & m;
    x ^= x >> 14;
    x = (x * 21ULL) & m;
    x ^= x >> 28;
    x = (x * 0x80000001ULL) & m;
    return x;
}

// hash64 is a bijection on [0, 2^2k): the inverse lets the device index be keyed by the canonical
// k-mer itself, so the per-position hash (26 instructions) is paid once per index key at build time
// instead of once per read position.  Matching canonical k-mers == matching their hashes.
__host__ __device__ constexpr uint64_t inv_odd64(uint64_t a) {  // a^-1 mod 2^64 by Newton's iteration
    uint64_t x = a;
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}
__device__ __forceinline__ uint64_t unxorshift(uint64_t y, int s) {
    uint64_t x = y;
#pragma unroll
    for (int i = 0; i < 4; ++i) x = y ^ (x >> s);  // enough for 56-bit values and s >= 14
    return x;
}
__device__ __forceinline__ uint64_t hash64_inv(uint64_t y, uint64_t m) {
    uint64_t x = (y * inv_odd64(0x80000001ULL)) & m;
    x = unxorshift(x, 28);
    x = (x * inv_odd64(21ULL)) & m;
    x = unxorshift(x, 14);
    x = (x * inv_odd64(265ULL)) & m;
    x = unxorshift(x, 24);
    return ((x + 1ULL) * inv_odd64(0x1FFFFFULL)) & m;
}

// Reverse complement of the k bases held in the low 2k bits of f (oldest base highest).
__device__ __forceinline__ uint64_t revcomp2k(uint64_t f, uint32_t k) {
    uint64_t y = __brevll(~f);
    y = ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
    return y >> (64 - 2 * k);
}

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t fmix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}
// k1 is the seed-independent half of the work: mix the 8-byte key once, reuse for all seeds.
__device__ __forceinline__ uint64_t murmur3_k1(uint64_t key) {
    uint64_t k1 = key * 0x87c37b91114253d5ULL;
    return rotl64(k1, 31) * 0x4cf5ad432745937fULL;
}
__device__ __forceinline__ uint64_t murmur3_sum_from_k1(uint64_t k1, uint32_t seed) {
    uint64_t h1 = seed, h2 = seed;
    h1 ^= k1;
    h1 ^= 8;
    h2 ^= 8;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    h1 += h2;
    h2 += h1;
    return h1 + h2;
}

// a % d for a fixed d, exact for all 64-bit a (Lemire fastmod with a 128-bit magic M = floor((2^128-1)/d)+1).
struct FastMod64 {
    uint64_t m_hi, m_lo, d;
};
__device__ __forceinline__ uint64_t fastmod64(uint64_t a, const FastMod64& f) {
    // lowbits = (M * a) mod 2^128
    uint64_t lb_lo = f.m_lo * a;
    uint64_t lb_hi = __umul64hi(f.m_lo, a) + f.m_hi * a;
    // result = (lowbits * d) >> 128
    uint64_t t_hi = __umul64hi(lb_lo, f.d);
    uint64_t u_lo = lb_hi * f.d;
    uint64_t u_hi = __umul64hi(lb_hi, f.d);
    uint64_t s = t_hi + u_lo;
    return u_hi + (s < t_hi ? 1ULL : 0ULL);
}

// ---- 256-bit sector load: one index bucket ---------------------------------
__device__ __forceinline__ void ld_bucket(const uint64_t* p, uint64_t (&v)[4]) {
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3])
                 : "l"(p));
}

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Home bucket of a key (a canonical k-mer): high word of the Fibonacci product, range-reduced without a division.
// Fibonacci multiply: the product's high word (bucket, filter word) depends on every bit of the k-mer, the top
// of its low word (filter bits) on the 16 most recent bases.
__device__ __forceinline__ uint64_t key_mix(uint64_t key56) { return key56 * 0x9E3779B97F4A7C15ULL; }
// high word of key56 * 0x9E3779B97F4A7C15 in three multiply-adds (the compiler's 64-bit product takes four and an add)
__device__ __forceinline__ uint32_t key_mix_hi(uint64_t key56) {
    const uint32_t lo = (uint32_t)key56, hi = (uint32_t)(key56 >> 32);
    uint32_t t = __umulhi(lo, 0x7F4A7C15u);
    t = lo * 0x9E3779B9u + t;
    return hi * 0x7F4A7C15u + t;
}
__device__ __forceinline__ uint32_t bucket_of(uint64_t key56, uint32_t nbuckets) {
    return __umulhi(key_mix_hi(key56), nbuckets);
}

// ---- presence pre-filter (word-blocked Bloom, 2 bits per entry in one 32-bit word) -----------
// word: which 32-bit word; bits: the entry's two bit positions (see prefilter_mask).  The word comes from the
// high half of the Fibonacci product, the bits from the top of its low half.
__device__ __forceinline__ void prefilter_slot(uint64_t entry, uint32_t nwords, uint32_t& word, uint32_t& bits) {
    const uint64_t mixed = key_mix(entry);
    word = __umulhi((uint32_t)(mixed >> 32), nwords);
    bits = (uint32_t)mixed >> 22;
}
// two bits of the word: position s = bits & 31 and s + d (mod 32), d = bits >> 5 (d == 0: just one)
__device__ __forceinline__ uint32_t prefilter_mask(uint32_t bits) {
    const uint32_t pair = 1u | (1u << (bits >> 5));
    return __funnelshift_l(pair, pair, bits & 31u);
}

// The presence pre-filter is keyed by canonical (k-d)-mers, d = kFilterDrop: every one of the d+1 such words
// inside every index k-mer.  The k-mers ending at read positions i ... i+d all contain the (k-d)-mer ending
// at i, so ONE filter lookup can rule out all d+1 of them -- and the scatter kernel is bound by exactly these
// gathers (one L1TEX wavefront each).  With fwd / rev the encoder's registers at position i: that word is the
// low 2(k-d) bits of fwd, and its reverse complement is rev without its d lowest bases.  Only the k-d most
// recent bases enter, so the value is right whenever any of the d+1 k-mers is emitted.
// kSpan = d + 1, the read positions one filter lookup speaks for (4 or 8): 4 where the filter is L2-resident (a lookup
// is then one L1TEX wavefront), 8 for the filter of a huge index, which lives in DRAM, where every lookup saved is a random
// sector not fetched (measured on the human-scale index: 58.9 vs 53.1 G k-mers/s).
template <int kSpan>
__device__ __forceinline__ uint64_t shared_smer(uint64_t fwd, uint64_t rev, uint64_t kmask) {
    const uint64_t a = fwd & (kmask >> (2 * (kSpan - 1))), b = rev >> (2 * (kSpan - 1));
    return a < b ? a : b;
}

// ---- the view of a staged chunk --------------------------------------------
// `al` is the 16-byte aligned-down base; live bytes are [lo, hi) relative to it.
// Everything outside behaves like '\n'.
struct Chunk {
    const uint8_t* al;
    int64_t lo, hi;
    const unsigned int* skip;  // optional device flag: non-zero = do not count this chunk (a FASTQ block that
                               // failed the on-device format check; the host re-parses it with kseq semantics)
};

struct KmerParams {
    uint32_t k;
    uint64_t mask;
};

// Where a kernel reads the chunk's text from: global memory as it lies (streaming 128-bit loads), or a copy of the
// CTA's tile that a bulk async copy (TMA) staged in shared memory (byte `off0` of the chunk sits at shared address saddr).
struct GlobalText {
    const uint8_t* al;
    __device__ __forceinline__ uint4 ld16(int64_t off) const { return ld_stream16(al + off); }
    __device__ __forceinline__ uint32_t ld4(int64_t off) const {
        uint32_t r;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(al + off));
        return r;
    }
};
struct SharedText {
    uint32_t saddr;
    int64_t off0;
    __device__ __forceinline__ uint4 ld16(int64_t off) const {
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr + (uint32_t)(off - off0)));
        return r;
    }
    __device__ __forceinline__ uint32_t ld4(int64_t off) const {
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(saddr + (uint32_t)(off - off0)));
        return r;
    }
};

// One 32-bit word of text -> bits 0-7: its four 2-bit codes (first base highest), bits 8-11: their validity
// (first base in bit 11).  Out-of-range bytes are invalid.
template <class Src>
__device__ __forceinline__ uint32_t encode_word(const Chunk& c, const Src& src, int64_t off, const uint8_t* lut) {
    if (off + 4 <= c.lo || off >= c.hi || off < 0) return 0;
    const uint32_t t = src.ld4(off);
    const uint16_t* w4 = reinterpret_cast<const uint16_t*>(lut + 256);
    uint32_t v = (uint32_t)w4[t & 0xffu] | w4[256 + ((t >> 8) & 0xffu)] | w4[512 + ((t >> 16) & 0xffu)] | w4[768 + (t >> 24)];
    if (off < c.lo || off + 4 > c.hi) {
        uint32_t keep = 0xffu;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (off + j >= c.lo && off + j < c.hi) keep |= 1u << (11 - j);
        v &= keep;
    }
    return v;
}

// The same for a word that lies wholly inside the chunk, without the table where it is not needed: the code of A C G T
// (either case) is ((b >> 1) ^ (b >> 2)) & 3, done on the four bytes at once; a byte permute turns the codes back into
// the letters they stand for, and a word that equals that (case folded) holds four valid bases.  Any other word -- it
// holds an N, a newline, a U, one of the bytes 0-3 the reference's table also accepts, ... -- takes the table.
// 13 instructions per word instead of ~50 (text is almost all ACGT) -- and yet SLOWER in the scatter kernel than the
// table (7.50 vs 7.22 ms on the chr20 workload): the lookups ride on the otherwise idle shared-memory pipe, the extra
// integer work does not.  Kept for the record as an A/B build (-DVG_ARITH_ENCODER); the table is the default.
__device__ __forceinline__ uint32_t encode_word_in_range(uint32_t w, const uint8_t* lut) {
    const uint32_t t = ((w >> 1) ^ (w >> 2)) & 0x03030303u;        // code of byte i in bits 8i, 8i+1
    const uint32_t u = (t | (t >> 4)) & 0x00330033u;               // codes of bytes 0,1 in nibbles 0,1; of bytes 2,3 in nibbles 4,5
    const uint32_t letters = __byte_perm(0x54474341u, 0u, __byte_perm(u, 0u, 0x4420));  // selector nibble i = code of byte i
    if (((w & 0xDFDFDFDFu) ^ letters) == 0u) return ((t * 0x40100401u) >> 24) | 0xf00u;
    const uint16_t* w4 = reinterpret_cast<const uint16_t*>(lut + 256);
    return (uint32_t)w4[w & 0xffu] | w4[256 + ((w >> 8) & 0xffu)] | w4[512 + ((w >> 16) & 0xffu)] | w4[768 + (w >> 24)];
}

// Encode one 16-byte segment to (2-bit packed, first base in the top bits; validity mask, first
// base in bit 15).  Out-of-range bytes are invalid.
template <class Src>
__device__ __forceinline__ void encode_seg(const Chunk& c, const Src& src, int64_t off, const uint8_t* lut,
                                           uint32_t& packed, uint32_t& vmask) {
    packed = 0;
    vmask = 0;
    if (off + kSegBytes <= c.lo || off >= c.hi || off < 0) return;
    uint4 w = src.ld16(off);
    uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#ifndef VG_ARITH_ENCODER
        const uint32_t t = ws[i];
        const uint16_t* w4 = reinterpret_cast<const uint16_t*>(lut + 256);
        const uint32_t v = (uint32_t)w4[t & 0xffu] | w4[256 + ((t >> 8) & 0xffu)] | w4[512 + ((t >> 16) & 0xffu)] |
                           w4[768 + (t >> 24)];
#else
        const uint32_t v = encode_word_in_range(ws[i], lut);
#endif
        packed = (packed << 8) | (v & 0xffu);
        vmask = (vmask << 4) | (v >> 8);
    }
    if (off < c.lo || off + kSegBytes > c.hi) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (off + j >= c.lo && off + j < c.hi) keep |= 1u << (15 - j);
        vmask &= keep;
    }
}

// The same two with the newline flags of a table built with hard_flags: encode_word_h returns them in bits 12-15,
// encode_seg_h as a third mask.  Bytes outside the chunk are hard boundaries.
template <class Src>
__device__ __forceinline__ uint32_t encode_word_h(const Chunk& c, const Src& src, int64_t off, const uint8_t* lut) {
    if (off + 4 <= c.lo || off >= c.hi || off < 0) return 0xf000u;
    const uint32_t t = src.ld4(off);
    const uint16_t* w4 = reinterpret_cast<const uint16_t*>(lut + 256);
    uint32_t v = (uint32_t)w4[t & 0xffu] | w4[256 + ((t >> 8) & 0xffu)] | w4[512 + ((t >> 16) & 0xffu)] | w4[768 + (t >> 24)];
    if (off < c.lo || off + 4 > c.hi) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (off + j < c.lo || off + j >= c.hi) v = (v & ~(0x1100u << (3 - j))) | (0x1000u << (3 - j));
    }
    return v;
}
template <class Src>
__device__ __forceinline__ void encode_seg_h(const Chunk& c, const Src& src, int64_t off, const uint8_t* lut,
                                             uint32_t& packed, uint32_t& vmask, uint32_t& hmask) {
    packed = 0;
    vmask = 0;
    hmask = 0xffffu;
    if (off + kSegBytes <= c.lo || off >= c.hi || off < 0) return;
    uint4 w = src.ld16(off);
    uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    hmask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t t = ws[i];
        const uint16_t* w4 = reinterpret_cast<const uint16_t*>(lut + 256);
        const uint32_t v = (uint32_t)w4[t & 0xffu] | w4[256 + ((t >> 8) & 0xffu)] | w4[512 + ((t >> 16) & 0xffu)] |
                           w4[768 + (t >> 24)];
        packed = (packed << 8) | (v & 0xffu);
        vmask = (vmask << 4) | ((v >> 8) & 0xfu);
        hmask = (hmask << 4) | (v >> 12);
    }
    if (off < c.lo || off + kSegBytes > c.hi) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (off + j >= c.lo && off + j < c.hi) keep |= 1u << (15 - j);
        vmask &= keep;
        hmask = (hmask & keep) | (~keep & 0xffffu);
    }
}

// ---- odd k: warp-cooperative position-parallel encoder ----------------------
// For odd k, fwd == rev cannot happen, so "emit at i" == "the k bytes ending at i
// are all valid".  Each lane owns 16 consecutive positions and borrows the 32 preceding bases
// from its two left neighbours by shuffle (lanes 0/1 re-read them from memory).
// Warp-cooperative set-up, then per-lane rolling a few positions at a time so that only one probe
// batch of keys is live at once (the sector loads in flight per lane are the register budget).
// reverse complement of 16 bases held in a word (first base in the top bits)
__device__ __forceinline__ uint32_t revcomp16(uint32_t x) {
    const uint32_t y = __brev(~x);
    return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}

#ifdef VG_ROLLING_ENCODER
struct OddEncoder {  // rolling registers (A/B build: -DVG_ROLLING_ENCODER)
    uint64_t fwd, rev;
    uint32_t p0;     // own 16 bases, 2 bits each, first base in the top bits
    uint32_t all_k;  // bit (15 - j): the k bytes ending at own position j are all valid

    __device__ __forceinline__ void init(const Chunk& c, int64_t off, const KmerParams& kp, const uint8_t* lut) {
        init(c, GlobalText{c.al}, off, kp, lut);
    }
    template <class Src>
    __device__ __forceinline__ void init(const Chunk& c, const Src& src, int64_t off, const KmerParams& kp, const uint8_t* lut) {
        const int lane = threadIdx.x & 31;
        uint32_t v0;
        encode_seg(c, src, off, lut, p0, v0);
        // The 32 bases in front of the warp's text (what lanes 0 and 1 lack a left neighbour for): lanes 0-7 encode one
        // 32-bit word each and everybody collects the eight results -- a few dozen instructions, where two lanes
        // running the whole 16-byte encoder would cost every warp two full segments' worth of issue slots.
        uint32_t wv = 0;
        if (lane < 8) wv = encode_word(c, src, off - 16 * lane - 32 + 4 * lane, lut);
        uint32_t x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = __shfl_sync(kFullMask, wv, q);
        // segment (warp_off - 32) from words 0-3, segment (warp_off - 16) from words 4-7
        const uint32_t ex0 = ((x[0] & 0xffu) << 24) | ((x[1] & 0xffu) << 16) | ((x[2] & 0xffu) << 8) | (x[3] & 0xffu);
        const uint32_t exv0 = ((x[0] >> 8) << 12) | ((x[1] >> 8) << 8) | ((x[2] >> 8) << 4) | (x[3] >> 8);
        const uint32_t ex1 = ((x[4] & 0xffu) << 24) | ((x[5] & 0xffu) << 16) | ((x[6] & 0xffu) << 8) | (x[7] & 0xffu);
        const uint32_t exv1 = ((x[4] >> 8) << 12) | ((x[5] >> 8) << 8) | ((x[6] >> 8) << 4) | (x[7] >> 8);
        uint32_t p1 = __shfl_up_sync(kFullMask, p0, 1), v1 = __shfl_up_sync(kFullMask, v0, 1);
        uint32_t p2 = __shfl_up_sync(kFullMask, p0, 2), v2 = __shfl_up_sync(kFullMask, v0, 2);
        if (lane == 0) { p1 = ex1; v1 = exv1; p2 = ex0; v2 = exv0; }
        if (lane == 1) { p2 = ex1; v2 = exv1; }
        // V: validity of the 48 bases in view, first base in bit 47.  f(n)[i] = bits i..i+n-1 all
        // set = "the n bytes ending at the base of bit i are valid"; f(a+b) = f(a) & (f(b) >> a),
        // so f(k) falls out of the binary decomposition of k.
        const uint64_t V = ((uint64_t)v2 << 32) | ((uint64_t)v1 << 16) | v0;
        uint64_t pw = V, acc = ~0ULL;
        uint32_t have = 0;
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
            if (kp.k & (1u << bit)) {
                acc &= pw >> have;
                have += 1u << bit;
            }
            pw &= pw >> (1u << bit);
        }
        all_k = (uint32_t)acc & 0xffffu;
        fwd = (((uint64_t)p2 << 32) | p1) & kp.mask;
        rev = revcomp2k(fwd, kp.k);
    }

    // Consumes the next N own positions: keys[j] = the canonical k-mer ending there (kHashed: its
    // hash64, i.e. the reference's key >> 8); returns the N-bit emit mask (bit j: the reference encoder
    // emits; keys[j] is meaningless where it does not).  With kPairs, pairs[q] = the pre-filter entry that
    // speaks for positions q * kSpan ... (q + 1) * kSpan - 1 (see shared_smer).
    template <int N, bool kHashed = true, bool kPairs = false, int kSpan = 4>
    __device__ __forceinline__ uint32_t next(const KmerParams& kp, uint64_t (&keys)[N], uint64_t* pairs = nullptr) {
        const uint32_t top = 2 * (kp.k - 1);
        uint32_t emit = 0;
        if (kp.k >= 17) {  // uniform: 2(k-1) >= 32, so the incoming complement only touches the high word
            const uint32_t mask_hi = (uint32_t)(kp.mask >> 32), tsh = top - 32;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const uint32_t cb = p0 >> 30;
                p0 <<= 2;
                fwd = ((fwd << 2) | cb) & kp.mask;
                const uint32_t rhi = (uint32_t)(rev >> 32), rlo = (uint32_t)rev;
                rev = ((uint64_t)((rhi >> 2) | ((3u ^ cb) << tsh)) << 32) | __funnelshift_r(rlo, rhi, 2);
                const uint64_t canon = fwd < rev ? fwd : rev;
                keys[j] = kHashed ? hash64_wide(canon, mask_hi) : canon;
                if (kPairs && j % kSpan == 0) pairs[j / kSpan] = shared_smer<kSpan>(fwd, rev, kp.mask);
                emit |= ((all_k >> 15) & 1u) << j;
                all_k <<= 1;
            }
            return emit;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            uint64_t cb = p0 >> 30;
            p0 <<= 2;
            fwd = ((fwd << 2) | cb) & kp.mask;
            rev = (rev >> 2) | ((3ULL ^ cb) << top);
            const uint64_t canon = fwd < rev ? fwd : rev;
            keys[j] = kHashed ? hash64(canon, kp.mask) : canon;
            if (kPairs && j % kSpan == 0) pairs[j / kSpan] = shared_smer<kSpan>(fwd, rev, kp.mask);
            emit |= ((all_k >> 15) & 1u) << j;
            all_k <<= 1;
        }
        return emit;
    }
};

#else
// Window extraction instead of rolling registers: the lane's 48 bases in view are three words (p2 p1 p0), the k-mer
// that ends at own position j is a 2k-bit window of them -- two funnel shifts and a mask -- and its reverse complement
// is the mirrored window of the reverse-complemented words (computed once per segment, pre-shifted by the part of
// the offset that depends on k).  Every position is independent of the others: no serial chain through fwd / rev,
// and roughly half the instructions of the rolling update.
struct OddEncoder {
    uint32_t p0, p1, p2;  // own 16 bases and the 32 in front of them, 2 bits each, first base in the top bits
    uint32_t l0, l1, l2;  // revcomp(p2 p1 p0) >> (66 - 2k): the window of position j starts at bit 2j
    uint32_t all_k;       // bit (15 - j): the k bytes ending at own position j are all valid
    uint32_t pos;         // own positions consumed so far

    __device__ __forceinline__ void init(const Chunk& c, int64_t off, const KmerParams& kp, const uint8_t* lut) {
        init(c, GlobalText{c.al}, off, kp, lut);
    }
    template <class Src>
    __device__ __forceinline__ void init(const Chunk& c, const Src& src, int64_t off, const KmerParams& kp, const uint8_t* lut) {
        const int lane = threadIdx.x & 31;
        uint32_t v0;
        encode_seg(c, src, off, lut, p0, v0);
        // The 32 bases in front of the warp's text (what lanes 0 and 1 lack a left neighbour for): lanes 0-7 encode one
        // 32-bit word each and everybody collects the eight results -- a few dozen instructions, where two lanes
        // running the whole 16-byte encoder would cost every warp two full segments' worth of issue slots.
        uint32_t wv = 0;
        if (lane < 8) wv = encode_word(c, src, off - 16 * lane - 32 + 4 * lane, lut);
        uint32_t x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = __shfl_sync(kFullMask, wv, q);
        // segment (warp_off - 32) from words 0-3, segment (warp_off - 16) from words 4-7
        const uint32_t ex0 = ((x[0] & 0xffu) << 24) | ((x[1] & 0xffu) << 16) | ((x[2] & 0xffu) << 8) | (x[3] & 0xffu);
        const uint32_t exv0 = ((x[0] >> 8) << 12) | ((x[1] >> 8) << 8) | ((x[2] >> 8) << 4) | (x[3] >> 8);
        const uint32_t ex1 = ((x[4] & 0xffu) << 24) | ((x[5] & 0xffu) << 16) | ((x[6] & 0xffu) << 8) | (x[7] & 0xffu);
        const uint32_t exv1 = ((x[4] >> 8) << 12) | ((x[5] >> 8) << 8) | ((x[6] >> 8) << 4) | (x[7] >> 8);
        uint32_t v1 = __shfl_up_sync(kFullMask, v0, 1), v2 = __shfl_up_sync(kFullMask, v0, 2);
        p1 = __shfl_up_sync(kFullMask, p0, 1);
        p2 = __shfl_up_sync(kFullMask, p0, 2);
        if (lane == 0) { p1 = ex1; v1 = exv1; p2 = ex0; v2 = exv0; }
        if (lane == 1) { p2 = ex1; v2 = exv1; }
        // V: validity of the 48 bases in view, first base in bit 47.  f(n)[i] = bits i..i+n-1 all
        // set = "the n bytes ending at the base of bit i are valid"; f(a+b) = f(a) & (f(b) >> a),
        // so f(k) falls out of the binary decomposition of k.
        const uint64_t V = ((uint64_t)v2 << 32) | ((uint64_t)v1 << 16) | v0;
        uint64_t pw = V, acc = ~0ULL;
        uint32_t have = 0;
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
            if (kp.k & (1u << bit)) {
                acc &= pw >> have;
                have += 1u << bit;
            }
            pw &= pw >> (1u << bit);
        }
        all_k = (uint32_t)acc & 0xffffu;
        // reverse complement of the 96 bits in view, top word first: rc(p0) rc(p1) rc(p2); shifted right by 66 - 2k
        const uint32_t rt = revcomp16(p0), rm = revcomp16(p1), rb = revcomp16(p2);
        const uint32_t base = 66u - 2u * kp.k;  // 10 .. 64 for k = 28 .. 1 (uniform)
        if (base < 32u) {
            l0 = __funnelshift_r(rb, rm, base);
            l1 = __funnelshift_r(rm, rt, base);
            l2 = rt >> base;
        } else if (base < 64u) {
            l0 = __funnelshift_r(rm, rt, base - 32u);
            l1 = rt >> (base - 32u);
            l2 = 0;
        } else {
            l0 = rt;
            l1 = l2 = 0;
        }
        pos = 0;
    }

    // Consumes the next N own positions: keys[j] = the canonical k-mer ending there (kHashed: its
    // hash64, i.e. the reference's key >> 8); returns the N-bit emit mask (bit j: the reference encoder
    // emits; keys[j] is meaningless where it does not).  With kPairs, pairs[q] = the pre-filter entry that
    // speaks for positions q * kSpan ... (q + 1) * kSpan - 1 (see shared_smer).
    template <int N, bool kHashed = true, bool kPairs = false, int kSpan = 4>
    __device__ __forceinline__ uint32_t next(const KmerParams& kp, uint64_t (&keys)[N], uint64_t* pairs = nullptr) {
        const uint32_t mask_lo = (uint32_t)kp.mask, mask_hi = (uint32_t)(kp.mask >> 32);
        uint32_t emit = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const uint32_t q = pos + j;        // own position 0..15 (a constant wherever the caller's loop is unrolled)
            const uint32_t sf = 30u - 2u * q;  // the forward window ends 2(15 - q) bits above the bottom of p0
            const uint32_t sr = 2u * q;
            const uint64_t fwd = ((uint64_t)(__funnelshift_r(p1, p2, sf) & mask_hi) << 32) | (__funnelshift_r(p0, p1, sf) & mask_lo);
            const uint64_t rev = ((uint64_t)(__funnelshift_r(l1, l2, sr) & mask_hi) << 32) | (__funnelshift_r(l0, l1, sr) & mask_lo);
            const uint64_t canon = fwd < rev ? fwd : rev;
            keys[j] = kHashed ? hash64(canon, kp.mask) : canon;
            if (kPairs && j % kSpan == 0) pairs[j / kSpan] = shared_smer<kSpan>(fwd, rev, kp.mask);
            emit |= ((all_k >> (15u - q)) & 1u) << j;
        }
        pos += N;
        return emit;
    }
};
#endif

// Whole segment at once (used where register pressure does not matter).
__device__ __forceinline__ uint32_t encode_keys_odd(const Chunk& c, int64_t off, const KmerParams& kp,
                                                    const uint8_t* lut, uint64_t (&keys)[16]) {
    OddEncoder enc;
    enc.init(c, off, kp, lut);
    return enc.next<16>(kp, keys);
}

// ---- any k: exact state machine with a look-back to a synchronisation point --
// Needed for even k, where fwd == rev (a reverse-complement palindrome) skips a position
// without advancing the run length, and the registers survive ambiguous bases (src/kmer.cpp:
// 131-146, note `else l = 0, kmer_span = 0;` leaves kmer[] alone).  The state at the start of a
// lane's segment is recovered exactly by scanning back to either a hard boundary ('\n' / chunk
// edge: state zero) or 2k consecutive valid bytes none of whose k+1 clean windows is a
// palindrome (then the run length is >= k whatever came before), and rolling forward from there.
struct RollState {
    uint64_t fwd, rev;
    uint32_t run;  // saturates at k: only run >= k is ever observed
};

template <bool kHashed = true>
__device__ __forceinline__ bool roll_push(RollState& s, uint32_t e, const KmerParams& kp, uint64_t& key) {
    if (!(e & 4u)) {
        s.run = 0;
        return false;
    }
    uint64_t cb = e & 3u;
    s.fwd = ((s.fwd << 2) | cb) & kp.mask;
    s.rev = (s.rev >> 2) | ((3ULL ^ cb) << (2 * (kp.k - 1)));
    if (s.fwd == s.rev) return false;
    if (s.run < kp.k) s.run += 1;
    if (s.run < kp.k) return false;
    const uint64_t canon = s.fwd < s.rev ? s.fwd : s.rev;
    key = kHashed ? hash64(canon, kp.mask) : canon;
    return true;
}

__device__ __forceinline__ uint32_t chunk_entry(const Chunk& c, int64_t pos, const uint8_t* lut) {
    if (pos < c.lo || pos >= c.hi) return 8u;  // newline
    return lut[c.al[pos]];
}

// pairs (optional, 16 / kSpan entries): pairs[q] = shared_smer of the registers after byte q * kSpan
// of the segment -- right whenever one of the kSpan k-mers ending from there on is emitted (then the
// k - kSpan + 1 most recent valid bases are the bytes ending there).
template <bool kHashed = true, int kSpan = 4>
__device__ inline uint32_t encode_keys_any(const Chunk& c, int64_t off, const KmerParams& kp,
                                           const uint8_t* lut, uint64_t (&keys)[16], uint64_t* pairs = nullptr) {
#pragma unroll
    for (int j = 0; j < 16; ++j) keys[j] = 0;
    if (off >= c.hi || off + kSegBytes <= c.lo) return 0;
    const int64_t need = 2 * (int64_t)kp.k;
    RollState st;
    int64_t scan_from = off;  // exclusive upper end of the region still to scan backwards
    for (;;) {
        // scan back for a hard boundary or `need` consecutive valid bytes
        int64_t p = scan_from, cnt = 0, start = -1;
        bool hard = false;
        while (true) {
            if (p <= c.lo) { hard = true; start = c.lo; break; }
            uint32_t e = chunk_entry(c, p - 1, lut);
            if (e & 8u) { hard = true; start = p; break; }
            cnt = (e & 4u) ? cnt + 1 : 0;
            --p;
            if (cnt == need) { start = p; break; }
            if (!(e & 4u) && (p & 7) == 0) {
                // inside a run of N (a chromosome's centromere is megabases of them): eight at a time
                while (p - 8 >= c.lo) {
                    const uint64_t w = *reinterpret_cast<const uint64_t*>(c.al + p - 8);
                    if (w != 0x4E4E4E4E4E4E4E4EULL && w != 0x6E6E6E6E6E6E6E6EULL) break;
                    p -= 8;
                }
            }
        }
        st.fwd = st.rev = 0;
        st.run = 0;
        if (hard) {
            uint64_t key;
            for (int64_t q = start; q < off; ++q) roll_push(st, chunk_entry(c, q, lut), kp, key);
            break;
        }
        // candidate: roll through the 2k valid bytes; the last k+1 windows must not be palindromes
        bool clean = true;
        for (int64_t q = start; q < start + need; ++q) {
            uint32_t e = chunk_entry(c, q, lut);
            uint64_t cb = e & 3u;
            st.fwd = ((st.fwd << 2) | cb) & kp.mask;
            st.rev = (st.rev >> 2) | ((3ULL ^ cb) << (2 * (kp.k - 1)));
            if (q - start >= (int64_t)kp.k - 1 && st.fwd == st.rev) clean = false;
        }
        if (clean) {
            st.run = kp.k;
            uint64_t key;
            for (int64_t q = start + need; q < off; ++q) roll_push(st, chunk_entry(c, q, lut), kp, key);
            break;
        }
        scan_from = start + need - 1;  // look for an earlier synchronisation point
    }
    uint32_t emit = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        uint32_t e = chunk_entry(c, off + j, lut);
        if (e & 8u) {  // hard boundary: a new read starts after it
            st.fwd = st.rev = 0;
            st.run = 0;
        } else {
            uint64_t key;
            if (roll_push<kHashed>(st, e, kp, key)) {
                keys[j] = key;
                emit |= 1u << j;
            }
        }
        if (pairs && j % kSpan == 0) pairs[j / kSpan] = shared_smer<kSpan>(st.fwd, st.rev, kp.mask);
    }
    return emit;
}


// ---- even k: the window encoder with the reference's palindrome rule ------------------------
// For even k a window can equal its own reverse complement; the reference then skips the position WITHOUT advancing its
// run length (src/kmer.cpp:131-146), and an ambiguous base clears the run length but not the registers.  What that
// state machine emits, position by position:
//   * After a hard boundary (newline, chunk edge: registers zeroed) the registers cannot be equal while they fill up
//     (fwd pads with A, rev pads with the complement of T), and a palindrome at a later position t skips t alone: by
//     position i there have been i - b bases since the boundary b and at most i - b - k skipped ones, so the run length
//     is >= k whenever the k bytes ending at i are valid.  Hence emit(i) = all_k(i) && !palindrome(i): no look-back at all.
//   * After an ambiguous base the stale registers CAN be equal during the next k - 1 bases, each time delaying the first
//     emission by one; only then do palindromes further on matter as well.
// So a lane takes the closed form unless (a) a position fewer than k bases behind an ambiguous byte has registers --
// the bases on both sides of that byte, spliced -- that read the same on both strands, and that position lies in the
// lane's own segment or the four in front of it (stale_palindrome: the byte is cut out of the lane's 96-bit view and the
// windows are tested slot-parallel; only run in warps that have such a byte in sight), or (b) a clean palindrome lies
// in its own segment or the two in front of it.  Then it runs the exact state machine (exact_emit_mask) for its own 16
// positions.  With neither, the 2k - 1 bytes ending at an emitting position are valid or cut by a newline, and k clean
// non-palindromic windows in a row saturate the run length whatever came before.
// The rule itself is pinned without a GPU (tests/test_oracle.py::test_even_k_closed_form_rule), the kernels against the
// state machine in tests/test_gpu_parity.py (text dense in palindromes, tandem repeats, N and newlines).

// exact emit mask of the 16 positions of a segment (bit 15 - j = position j): encode_keys_any without the keys
__device__ __noinline__ uint32_t exact_emit_mask(const Chunk& c, int64_t off, const KmerParams& kp, const uint8_t* lut) {
    uint64_t keys[16];
    const uint32_t e = encode_keys_any<false, 16>(c, off, kp, lut, keys, nullptr);
    return __brev(e) >> 16;
}

#ifndef VG_ROLLING_ENCODER
// palindromes among the windows ending at the 16 own positions: bit 2(15 - j) of the result (a 2-bit slot per position)
// is set iff the k bases ending at own position j read the same on both strands.  Slot-parallel: pair d of a window is
// (base[t - d], base[t - k + 1 + d]); the pairs of all 16 windows are two shifted views of the 96 bits in sight, and a
// pair is complementary iff its XOR is 3.  The warp stops as soon as no lane has a candidate left (4^-d of them survive
// d pairs).
__device__ __forceinline__ uint32_t palindrome_slots(uint32_t p2, uint32_t p1, uint32_t p0, uint32_t k) {
    uint32_t acc = 0x55555555u;
    const uint32_t half = k >> 1;
    for (uint32_t d = 0; d < half; ++d) {
        const uint32_t a = __funnelshift_r(p0, p1, 2u * d);  // slot j: base[t_j - d]; 2d < 32
        const uint32_t sh = 2u * (k - 1u - d);               // 2 .. 54
        const uint32_t b = sh < 32u ? __funnelshift_r(p0, p1, sh) : __funnelshift_r(p1, p2, sh - 32u);
        const uint32_t x = a ^ b;
        acc &= x & (x >> 1);
        if ((d & 1u) && !__any_sync(kFullMask, acc != 0u)) break;
    }
    return acc;
}

// bit i of the result: the k bytes ending at the byte of bit i are all valid (V: one bit per byte, earlier bytes higher)
__device__ __forceinline__ uint64_t all_valid_k(uint64_t V, uint32_t k) {
    uint64_t pw = V, acc = ~0ULL;
    uint32_t have = 0;
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
        if (k & (1u << bit)) {
            acc &= pw >> have;
            have += 1u << bit;
        }
        pw &= pw >> (1u << bit);
    }
    return acc;
}

// palindrome_slots for one lane on its own (no votes), only for the positions in `want` (bit 15 - j = own position j)
__device__ __forceinline__ bool palindrome_among(uint32_t p2, uint32_t p1, uint32_t p0, uint32_t k, uint32_t want) {
    uint32_t acc = want;  // spread to one 2-bit slot per position
    acc = (acc | (acc << 8)) & 0x00ff00ffu;
    acc = (acc | (acc << 4)) & 0x0f0f0f0fu;
    acc = (acc | (acc << 2)) & 0x33333333u;
    acc = (acc | (acc << 1)) & 0x55555555u;
    const uint32_t half = k >> 1;
    for (uint32_t d = 0; d < half && acc; ++d) {
        const uint32_t a = __funnelshift_r(p0, p1, 2u * d);
        const uint32_t sh = 2u * (k - 1u - d);
        const uint32_t b = sh < 32u ? __funnelshift_r(p0, p1, sh) : __funnelshift_r(p1, p2, sh - 32u);
        const uint32_t x = a ^ b;
        acc &= x & (x >> 1);
    }
    return acc != 0u;
}

// The registers after an ambiguous byte: it is skipped, so until k bases have followed it they hold the bases on both
// sides of it spliced together -- and may read the same on both strands.  True iff that happens (or cannot be ruled
// out from the 48 bases in view) at one of the 16 positions of the view's last segment.  V / S: validity and
// ambiguity of the 48 bytes, first byte in bit 47.  Only called where S != 0.
// Per ambiguous byte: cut it out of the view (bases, V and S alike) and look for palindromes among the windows of the
// spliced text that end at own positions fewer than k bases after it.
__device__ __noinline__ bool stale_palindrome(uint32_t p2, uint32_t p1, uint32_t p0, uint64_t V, uint64_t S, uint32_t k) {
    const uint32_t warm = (uint32_t)V & ~(uint32_t)all_valid_k(V, k) & 0xffffu;  // own, valid, fewer than k bases into the run
    if (!warm) return false;
    uint64_t soft = S & 0xffffffffffffULL;
    while (soft) {
        const uint32_t bb = 63u - (uint32_t)__clzll((long long)soft);
        soft &= ~(1ULL << bb);
        if (bb == 0u) continue;
        const uint64_t below = (1ULL << bb) - 1ULL;
        const uint64_t inv_below = ~V & below;
        const uint32_t stop = inv_below ? 64u - (uint32_t)__clzll((long long)inv_below) : 0u;  // the run after it ends above this bit
        const uint32_t lo = max(stop, bb >= k - 1u ? bb - (k - 1u) : 0u);
        const uint32_t mine = (uint32_t)(below & ~((1ULL << lo) - 1ULL)) & warm;  // own positions 1 .. k-1 bases after the byte
        if (!mine) continue;
        // the view without the byte: everything above it moves down one place
        const uint64_t Vs = ((V >> (bb + 1u)) << bb) | (V & below);
        const uint64_t Ss = ((S >> (bb + 1u)) << bb) | (S & below) | (1ULL << 47);  // what moves in at the top is unknown
        const uint32_t y0 = __funnelshift_r(p0, p1, 2u), y1 = __funnelshift_r(p1, p2, 2u), y2 = p2 >> 2;
        uint32_t m0, m1, m2;  // 96-bit mask of the bases below the byte
        const uint32_t b2 = 2u * bb;
        if (b2 >= 64u) { m0 = m1 = ~0u; m2 = (1u << (b2 - 64u)) - 1u; }
        else if (b2 >= 32u) { m0 = ~0u; m1 = (1u << (b2 - 32u)) - 1u; m2 = 0u; }
        else { m0 = (1u << b2) - 1u; m1 = m2 = 0u; }
        const uint32_t q0 = (p0 & m0) | (y0 & ~m0), q1 = (p1 & m1) | (y1 & ~m1), q2 = (p2 & m2) | (y2 & ~m2);
        const uint32_t clean = (uint32_t)all_valid_k(Vs, k) & mine;
        if (clean && palindrome_among(q2, q1, q0, k, clean)) return true;
        if (mine & ~clean) {
            // another invalid byte fewer than k bases in front of the cut: a newline means registers still filling up
            // (never equal); a second ambiguous byte is not worked out here
            const uint64_t inv_above = ~Vs >> bb;  // bit 0: the byte that moved next to the run
            const uint32_t a = bb + (uint32_t)__ffsll((long long)inv_above) - 1u;  // inv_above != 0: bit 47 of Vs is clear
            if ((Ss >> a) & 1ULL) return true;
        }
    }
    return false;
}

struct EvenEncoder : OddEncoder {
    __device__ __forceinline__ void init(const Chunk& c, int64_t off, const KmerParams& kp, const uint8_t* lut) {
        init(c, GlobalText{c.al}, off, kp, lut);
    }
    // lut: built with hard_flags
    template <class Src>
    __device__ __forceinline__ void init(const Chunk& c, const Src& src, int64_t off, const KmerParams& kp, const uint8_t* lut) {
        const int lane = threadIdx.x & 31;
        uint32_t v0, h0;
        encode_seg_h(c, src, off, lut, p0, v0, h0);
        // The 64 bases in front of the warp's text: lanes 0-15 encode a word each, lanes 0 4 8 12 assemble a segment
        // each, everybody collects the four segments (ex[3] is the one right in front of the warp's text).
        uint32_t wv = 0xf000u;
        if (lane < 16) wv = encode_word_h(c, src, off - 16 * lane - 64 + 4 * lane, lut);
        const uint32_t w1 = __shfl_down_sync(kFullMask, wv, 1), w2 = __shfl_down_sync(kFullMask, wv, 2),
                       w3 = __shfl_down_sync(kFullMask, wv, 3);
        const uint32_t segp = ((wv & 0xffu) << 24) | ((w1 & 0xffu) << 16) | ((w2 & 0xffu) << 8) | (w3 & 0xffu);
        const uint32_t segv = (((wv >> 8) & 0xfu) << 12) | (((w1 >> 8) & 0xfu) << 8) | (((w2 >> 8) & 0xfu) << 4) | ((w3 >> 8) & 0xfu);
        const uint32_t segh = ((wv >> 12) << 12) | ((w1 >> 12) << 8) | ((w2 >> 12) << 4) | (w3 >> 12);
        uint32_t ex[4], exv[4], exs[4];  // bases, validity, ambiguous (neither valid nor newline)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            ex[m] = __shfl_sync(kFullMask, segp, 4 * m);
            const uint32_t vh = __shfl_sync(kFullMask, segv | (segh << 16), 4 * m);
            exv[m] = vh & 0xffffu;
            exs[m] = ~(vh | (vh >> 16)) & 0xffffu;
        }
        uint32_t v1 = __shfl_up_sync(kFullMask, v0, 1), v2 = __shfl_up_sync(kFullMask, v0, 2);
        p1 = __shfl_up_sync(kFullMask, p0, 1);
        p2 = __shfl_up_sync(kFullMask, p0, 2);
        if (lane == 0) { p1 = ex[3]; v1 = exv[3]; p2 = ex[2]; v2 = exv[2]; }
        if (lane == 1) { p2 = ex[3]; v2 = exv[3]; }
        const uint64_t V = ((uint64_t)v2 << 32) | ((uint64_t)v1 << 16) | v0;
        uint64_t pw = V, acc = ~0ULL;
        uint32_t have = 0;
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
            if (kp.k & (1u << bit)) {
                acc &= pw >> have;
                have += 1u << bit;
            }
            pw &= pw >> (1u << bit);
        }
        all_k = (uint32_t)acc & 0xffffu;
        // palindromes at own positions (only where the window is clean), spread to one bit per position
        uint32_t pal = palindrome_slots(p2, p1, p0, kp.k);
        if (pal) {
            pal = (pal | (pal >> 1)) & 0x33333333u;
            pal = (pal | (pal >> 2)) & 0x0f0f0f0fu;
            pal = (pal | (pal >> 4)) & 0x00ff00ffu;
            pal = (pal | (pal >> 8)) & 0x0000ffffu;
            pal &= all_k;
        }
        // palindromes among the k - 1 windows that end in front of the warp's text: lane q tests the one ending q + 1 bases
        // before it (its window is clean iff the k bytes are valid)
        bool lead_pal = false;
        {
            const uint64_t hi = ((uint64_t)ex[0] << 32) | ex[1], lo = ((uint64_t)ex[2] << 32) | ex[3];
            const uint64_t lv = ((uint64_t)exv[0] << 48) | ((uint64_t)exv[1] << 32) | ((uint64_t)exv[2] << 16) | exv[3];
            const uint32_t q = (uint32_t)lane;
            const uint64_t win = (q ? (lo >> (2u * q)) | (hi << (64u - 2u * q)) : lo) & kp.mask;
            const uint64_t ones = (1ULL << kp.k) - 1ULL;
            lead_pal = q + 1u < kp.k && ((lv >> q) & ones) == ones && win == revcomp2k(win, kp.k);
        }
        // Who must run the state machine instead: lanes within reach of a position where the registers, spliced over an
        // ambiguous byte, read the same on both strands (reach: the 2k - 2 positions after it, i.e. the lane's own segment
        // and the four in front of it; the lead-in segments count as lanes -4 .. -1), and lanes with a palindrome in
        // their own segment or the two in front of it.
        const uint32_t sf0 = ~(v0 | h0) & 0xffffu;
        const uint32_t soft_lanes = __ballot_sync(kFullMask, sf0 != 0u);
        const uint32_t pal_lanes = __ballot_sync(kFullMask, pal != 0u);
        const uint32_t lead_pal_any = __ballot_sync(kFullMask, lead_pal);
        const uint32_t lead_soft = (exs[0] ? 1u : 0u) | (exs[1] ? 2u : 0u) | (exs[2] ? 4u : 0u) | (exs[3] ? 8u : 0u);
        all_k &= ~pal;
        if (soft_lanes | pal_lanes | lead_pal_any | lead_soft) {  // rare: nothing of the kind in most warps' sight
            uint32_t stale_lanes = 0, lead_stale = 0;
            if (soft_lanes | lead_soft) {
                uint32_t sf1 = __shfl_up_sync(kFullMask, sf0, 1), sf2 = __shfl_up_sync(kFullMask, sf0, 2);
                if (lane == 0) { sf1 = exs[3]; sf2 = exs[2]; }
                if (lane == 1) sf2 = exs[3];
                const uint64_t S = ((uint64_t)sf2 << 32) | ((uint64_t)sf1 << 16) | sf0;
                bool stale = false;
                if (S) stale = stale_palindrome(p2, p1, p0, V, S, kp.k);
                stale_lanes = __ballot_sync(kFullMask, stale);
                if (lead_soft) {
                    // The same for the four lead-in segments, by lanes 0-3.  Their views reach 32 bases further back:
                    // lanes 0-7 fetch those now (two more segments, assembled at lanes 0 and 4).
                    uint32_t xw = 0xf000u;
                    if (lane < 8) xw = encode_word_h(c, src, off - 16 * lane - 96 + 4 * lane, lut);
                    const uint32_t x1 = __shfl_down_sync(kFullMask, xw, 1), x2 = __shfl_down_sync(kFullMask, xw, 2),
                                   x3 = __shfl_down_sync(kFullMask, xw, 3);
                    const uint32_t xp = ((xw & 0xffu) << 24) | ((x1 & 0xffu) << 16) | ((x2 & 0xffu) << 8) | (x3 & 0xffu);
                    const uint32_t xv = (((xw >> 8) & 0xfu) << 12) | (((x1 >> 8) & 0xfu) << 8) | (((x2 >> 8) & 0xfu) << 4) | ((x3 >> 8) & 0xfu);
                    const uint32_t xh = ((xw >> 12) << 12) | ((x1 >> 12) << 8) | ((x2 >> 12) << 4) | (x3 >> 12);
                    uint32_t fx[6], fv[6], fs[6];  // segments -2 .. 3 in front of the warp's text
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        fx[m] = __shfl_sync(kFullMask, xp, 4 * m);
                        const uint32_t vh = __shfl_sync(kFullMask, xv | (xh << 16), 4 * m);
                        fv[m] = vh & 0xffffu;
                        fs[m] = ~(vh | (vh >> 16)) & 0xffffu;
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m) { fx[m + 2] = ex[m]; fv[m + 2] = exv[m]; fs[m + 2] = exs[m]; }
                    // lane m < 4 takes lead-in segment m: view = segments m, m + 1, m + 2 of the six (selects, not
                    // indexing: the arrays live in registers)
                    auto at = [&](const uint32_t (&a)[6], int i) {
                        return i == 0 ? a[0] : i == 1 ? a[1] : i == 2 ? a[2] : i == 3 ? a[3] : i == 4 ? a[4] : a[5];
                    };
                    bool ls = false;
                    if (lane < 4) {
                        const uint64_t LV = ((uint64_t)at(fv, lane) << 32) | ((uint64_t)at(fv, lane + 1) << 16) | at(fv, lane + 2);
                        const uint64_t LS = ((uint64_t)at(fs, lane) << 32) | ((uint64_t)at(fs, lane + 1) << 16) | at(fs, lane + 2);
                        if (LS) ls = stale_palindrome(at(fx, lane), at(fx, lane + 1), at(fx, lane + 2), LV, LS, kp.k);
                    }
                    lead_stale = __ballot_sync(kFullMask, ls) & 0xfu;
                }
            }
            const uint64_t SL = ((uint64_t)stale_lanes << 4) | lead_stale;             // bit (l + 4): lane l
            const uint64_t PL = ((uint64_t)pal_lanes << 2) | (lead_pal_any ? 3u : 0u);  // bit (l + 2): lane l
            bool slow = (((SL >> lane) & 0x1fu) | ((PL >> lane) & 0x7u)) != 0;
#ifdef VG_EVEN_SKIP_SLOW  // timing experiment only: wrong counts near such positions
            slow = false;
#endif
            if (slow) all_k = exact_emit_mask(c, off, kp, lut);
        }
        const uint32_t rt = revcomp16(p0), rm = revcomp16(p1), rb = revcomp16(p2);
        const uint32_t base = 66u - 2u * kp.k;
        if (base < 32u) {
            l0 = __funnelshift_r(rb, rm, base);
            l1 = __funnelshift_r(rm, rt, base);
            l2 = rt >> base;
        } else if (base < 64u) {
            l0 = __funnelshift_r(rm, rt, base - 32u);
            l1 = rt >> (base - 32u);
            l2 = 0;
        } else {
            l0 = rt;
            l1 = l2 = 0;
        }
        pos = 0;
    }
};
#endif

#ifndef VG_ROLLING_ENCODER
// Whole segment at once, even k (lut built with hard_flags).
__device__ __forceinline__ uint32_t encode_keys_even(const Chunk& c, int64_t off, const KmerParams& kp,
                                                     const uint8_t* lut, uint64_t (&keys)[16]) {
    EvenEncoder enc;
    enc.init(c, off, kp, lut);
    return enc.next<16>(kp, keys);
}
#else
__device__ __forceinline__ uint32_t encode_keys_even(const Chunk& c, int64_t off, const KmerParams& kp,
                                                     const uint8_t* lut, uint64_t (&keys)[16]) {
    return encode_keys_any(c, off, kp, lut, keys);
}
#endif

// which encoder a kernel instantiation runs
constexpr int kEncAny = 0, kEncOdd = 1, kEncEven = 2;
template <int kEnc> struct EncoderOf { using type = OddEncoder; };
#ifndef VG_ROLLING_ENCODER
template <> struct EncoderOf<kEncEven> { using type = EvenEncoder; };
#endif

}  // namespace vg
