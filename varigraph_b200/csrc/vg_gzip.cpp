// vg_gzip.cpp -- parallel inflate of a gzip file that is held in memory (mmap), for the FASTQ feeder.
//
// The reference reads .gz input through zlib's gzread on one thread per file (include/kseq.h:242 `KSEQ_INIT(gzFile,
// gzread)`, src/fastq_kmer.cpp:74 `gzopen`), ~0.2 G bases/s -- three orders of magnitude below what the count kernels
// take.  A DEFLATE stream has no index, but it can still be cut (the two-pass scheme of pugz / rapidgzip):
//   1. the compressed bytes are cut into chunks; for every chunk but the first a worker SEARCHES the first bit position
//      at or after the chunk's first byte where a dynamic-Huffman block can start (header fields in range, code-length
//      code complete, literal/length and distance codes complete, the first symbols decode to text);
//   2. every worker inflates from its start until it arrives EXACTLY at the start of a later chunk.  It does not know
//      the 32 KiB of text in front of its first byte, so it writes 16-bit symbols: a literal byte, or a marker "the byte
//      at offset o of the unknown window"; copies out of its own output carry markers along;
//   3. in file order, the last 32 KiB of every chunk are resolved against the window of its predecessor (a table lookup
//      per symbol); then all chunks are resolved to bytes in parallel, and the CRC-32 of every gzip member is checked
//      (per-chunk CRCs folded with crc32_combine).
// Nothing depends on the search being right: a start that no predecessor arrives at is discarded (the predecessor simply
// keeps inflating), a chunk with no start found is inflated by its predecessor, and a CRC or ISIZE mismatch is an
// error.  Multi-member files (bgzip, concatenated gzip) are followed across member boundaries; after one, the window
// is known to be empty and no markers are produced.  Output bytes are identical to zlib's (tests/test_capi_cpu.py
// compares against Python's zlib on single-member, multi-member, stored-block, fixed-block and binary inputs).
#include "vg_gzip.h"

#include <immintrin.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace vg {
namespace gz {
namespace {

constexpr int kLitBits = 10, kDistBits = 9;
constexpr uint32_t kInvalid = 0, kLiteral = 1, kLength = 2, kEob = 3, kSub = 4, kDist = 5;
constexpr int kLitTabSize = (1 << kLitBits) + 288 * 32, kDistTabSize = (1 << kDistBits) + 32 * 64;

inline uint32_t entry(uint32_t kind, uint32_t bits, uint32_t extra, uint32_t val) { return bits | (kind << 5) | (extra << 8) | (val << 16); }
inline uint32_t e_bits(uint32_t e) { return e & 31u; }
inline uint32_t e_kind(uint32_t e) { return (e >> 5) & 7u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 15u; }
inline uint32_t e_val(uint32_t e) { return e >> 16; }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Bits {
    const uint8_t* data;
    uint64_t size;   // bytes
    uint64_t pos;    // next byte to load (may run past size: zeros are shifted in, `overrun` says so)
    uint64_t buf;
    int cnt;
    void seek(uint64_t bit) {
        pos = bit >> 3;
        buf = 0;
        cnt = 0;
        refill();
        const int skip = (int)(bit & 7u);
        buf >>= skip;
        cnt -= skip;
    }
    inline void refill() {
        if (pos + 8 <= size) {
            uint64_t w;
            memcpy(&w, data + pos, 8);
            buf |= w << cnt;
            pos += (uint64_t)((63 - cnt) >> 3);
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (pos < size) buf |= (uint64_t)data[pos] << cnt;
                ++pos;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    inline void drop(int n) {
        buf >>= n;
        cnt -= n;
    }
    inline uint32_t take(int n) {
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
    uint64_t bitpos() const { return pos * 8 - (uint64_t)cnt; }
    bool overrun() const { return bitpos() > size * 8; }
};

inline uint32_t reverse_bits(uint32_t c, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((c >> i) & 1u) << (n - 1 - i);
    return r;
}

// Canonical Huffman code -> two-level lookup table indexed by the next bits of the stream (LSB first).
// Returns false for an over-subscribed code, or an incomplete one unless `allow_incomplete` and it is the single
// 1-bit code zlib accepts (inftrees.c: "incomplete set" unless max == 1) or has no symbols at all.
template <class MakeEntry>
bool build_table(const uint8_t* lens, int n, int primary, uint32_t* tab, int tab_size, bool allow_incomplete, MakeEntry make) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    int max_len = 0;
    for (int l = 1; l < 16; ++l)
        if (count[l]) max_len = l;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    if (left > 0 && !(allow_incomplete && max_len <= 1)) return false;
    memset(tab, 0, sizeof(uint32_t) * (size_t)(1 << primary));
    if (max_len == 0) return true;
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l < 16; ++l) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
    }
    uint8_t sub_max[1 << kLitBits];
    bool any_long = max_len > primary;
    if (any_long) memset(sub_max, 0, (size_t)(1 << primary));
    uint32_t codes[288];
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = reverse_bits(next_code[l]++, l);
        codes[s] = r;
        if (l <= primary) {
            const uint32_t e = make(s, l);
            for (uint32_t i = r; i < (1u << primary); i += 1u << l) tab[i] = e;
        } else {
            uint8_t& m = sub_max[r & ((1u << primary) - 1u)];
            if (l > m) m = (uint8_t)l;
        }
    }
    if (!any_long) return true;
    int cur = 1 << primary;
    for (uint32_t low = 0; low < (1u << primary); ++low) {
        if (!sub_max[low]) continue;
        const int sb = sub_max[low] - primary;
        if (cur + (1 << sb) > tab_size) return false;
        memset(tab + cur, 0, sizeof(uint32_t) * (size_t)(1 << sb));
        tab[low] = entry(kSub, (uint32_t)primary, (uint32_t)sb, (uint32_t)cur);
        cur += 1 << sb;
    }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l <= primary) continue;
        const uint32_t r = codes[s], low = r & ((1u << primary) - 1u);
        const uint32_t sub = tab[low];
        const uint32_t off = e_val(sub), sb = e_extra(sub);
        const uint32_t e = make(s, l - primary);
        for (uint32_t i = r >> primary; i < (1u << sb); i += 1u << (l - primary)) tab[off + i] = e;
    }
    return true;
}

inline uint32_t lit_entry(int s, int bits) {
    if (s < 256) return entry(kLiteral, (uint32_t)bits, 0, (uint32_t)s);
    if (s == 256) return entry(kEob, (uint32_t)bits, 0, 0);
    if (s > 285) return entry(kInvalid, (uint32_t)bits, 0, 0);
    return entry(kLength, (uint32_t)bits, kLenExtra[s - 257], kLenBase[s - 257]);
}
inline uint32_t dist_entry(int s, int bits) {
    if (s > 29) return entry(kInvalid, (uint32_t)bits, 0, 0);
    return entry(kDist, (uint32_t)bits, kDistExtra[s], kDistBase[s]);
}

struct Tables {
    uint32_t lit[kLitTabSize];
    uint32_t dist[kDistTabSize];
};

void fixed_lengths(uint8_t* ll, uint8_t* dl) {
    for (int i = 0; i < 144; ++i) ll[i] = 8;
    for (int i = 144; i < 256; ++i) ll[i] = 9;
    for (int i = 256; i < 280; ++i) ll[i] = 7;
    for (int i = 280; i < 288; ++i) ll[i] = 8;
    for (int i = 0; i < 32; ++i) dl[i] = 5;
}

// Header of a dynamic block (the 3 block bits already consumed): reads the code lengths and builds the tables.
// strict: the search's idea of plausible (complete literal/length code; zlib itself also takes a lone 1-bit code).
bool read_dynamic_header(Bits& b, Tables& t, bool strict) {
    b.refill();
    const uint32_t hlit = b.take(5) + 257, hdist = b.take(5) + 1, hclen = b.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t cl[19] = {0};
    for (uint32_t i = 0; i < hclen; ++i) {
        if (b.cnt < 3) b.refill();
        cl[kClOrder[i]] = (uint8_t)b.take(3);
    }
    uint32_t pre[128];
    {   // the code-length code: 7 bits at most, one level
        int count[8] = {0};
        for (int i = 0; i < 19; ++i) count[cl[i]]++;
        count[0] = 0;
        int left = 1;
        for (int l = 1; l < 8; ++l) {
            left <<= 1;
            left -= count[l];
            if (left < 0) return false;
        }
        if (left > 0) return false;  // zlib: an incomplete code-length code is always an error
        uint32_t next_code[8], code = 0;
        for (int l = 1; l < 8; ++l) {
            code = (code + (uint32_t)count[l - 1]) << 1;
            next_code[l] = code;
        }
        memset(pre, 0, sizeof(pre));
        for (int s = 0; s < 19; ++s) {
            const int l = cl[s];
            if (!l) continue;
            const uint32_t r = reverse_bits(next_code[l]++, l);
            for (uint32_t i = r; i < 128; i += 1u << l) pre[i] = (uint32_t)l | ((uint32_t)s << 8);
        }
    }
    uint8_t lens[320];
    uint32_t n = 0;
    const uint32_t total = hlit + hdist;
    while (n < total) {
        b.refill();
        const uint32_t e = pre[b.peek(7)];
        if (!e) return false;
        b.drop((int)(e & 0xffu));
        const uint32_t s = e >> 8;
        if (s < 16) {
            lens[n++] = (uint8_t)s;
        } else {
            uint32_t rep, val = 0;
            if (s == 16) {
                if (n == 0) return false;
                val = lens[n - 1];
                rep = 3 + b.take(2);
            } else if (s == 17) {
                rep = 3 + b.take(3);
            } else {
                rep = 11 + b.take(7);
            }
            if (n + rep > total) return false;
            while (rep--) lens[n++] = (uint8_t)val;
        }
    }
    if (b.overrun()) return false;
    if (lens[256] == 0) return false;  // no end-of-block code
    if (!build_table(lens, (int)hlit, kLitBits, t.lit, kLitTabSize, !strict, lit_entry)) return false;
    if (!build_table(lens + hlit, (int)hdist, kDistBits, t.dist, kDistTabSize, true, dist_entry)) return false;
    return true;
}

inline bool text_byte(uint32_t c) { return (c >= 32 && c < 127) || c == '\n' || c == '\r' || c == '\t'; }

// Can a dynamic block start at this bit?  Cheap tests first; then the header, then a trial decode of the first symbols.
bool plausible_block_start(const uint8_t* data, uint64_t size, uint64_t bit, Tables& t) {
    {   // BTYPE == 2, HLIT <= 29, HDIST <= 29 straight from the bytes
        const uint64_t byte = bit >> 3;
        if (byte + 4 > size) return false;
        uint32_t w;
        memcpy(&w, data + byte, 4);
        w >>= (bit & 7u);
        if (((w >> 1) & 3u) != 2u) return false;
        if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) return false;
        // the code-length code must be complete (Kraft sum exactly 1): 3 bits per length straight from the bytes; of
        // the candidates that get this far, some 99 % stop here, before any table is built
        if (byte + 12 > size) return false;
        uint64_t lo, hi = 0;
        memcpy(&lo, data + byte, 8);
        memcpy(&hi, data + byte + 8, 4);
        const unsigned sh = (unsigned)(bit & 7u) + 13u;  // 3 block bits + HLIT 5 + HDIST 5
        const uint64_t bits = (lo >> sh) | (hi << (64u - sh));  // >= 61 valid bits: HCLEN (4) + up to 19 x 3
        const uint32_t hclen = (uint32_t)(bits & 15u) + 4u;
        uint32_t kraft = 0;  // in units of 2^-7
        uint64_t rest = bits >> 4;
        for (uint32_t i = 0; i < hclen; ++i, rest >>= 3) {
            const uint32_t l = (uint32_t)(rest & 7u);
            if (l) kraft += 128u >> l;
        }
        if (kraft != 128u) return false;
    }
    Bits b{data, size, 0, 0, 0};
    b.seek(bit);
    b.drop(3);
    if (!read_dynamic_header(b, t, true)) return false;
    for (int n = 0; n < 4096; ++n) {
        b.refill();
        uint32_t e = t.lit[b.peek(kLitBits)];
        if (e_kind(e) == kSub) {
            b.drop(kLitBits);
            e = t.lit[e_val(e) + b.peek((int)e_extra(e))];
        }
        b.drop((int)e_bits(e));
        switch (e_kind(e)) {
            case kLiteral:
                if (!text_byte(e_val(e))) return false;
                break;
            case kLength: {
                b.drop((int)e_extra(e));
                uint32_t d = t.dist[b.peek(kDistBits)];
                if (e_kind(d) == kSub) {
                    b.drop(kDistBits);
                    d = t.dist[e_val(d) + b.peek((int)e_extra(d))];
                }
                if (e_kind(d) != kDist) return false;
                b.drop((int)e_bits(d));
                b.refill();
                b.drop((int)e_extra(d));
                break;
            }
            case kEob: {
                if (b.overrun()) return false;
                b.refill();
                const uint32_t hdr = b.peek(3);
                return (hdr >> 1) != 3u;  // the block after it must have a legal type (or be the end of the member)
            }
            default:
                return false;
        }
        if (b.overrun()) return false;
    }
    return true;
}

// gzip member header at byte `at` (RFC 1952): returns the byte offset of the deflate data, 0 if there is no header there
uint64_t parse_member_header(const uint8_t* d, uint64_t size, uint64_t at) {
    if (at + 18 > size || d[at] != 0x1f || d[at + 1] != 0x8b || d[at + 2] != 8) return 0;
    const uint8_t flg = d[at + 3];
    uint64_t p = at + 10;
    if (flg & 4) {  // FEXTRA
        if (p + 2 > size) return 0;
        p += 2 + ((uint64_t)d[p] | ((uint64_t)d[p + 1] << 8));
    }
    if (flg & 8) {  // FNAME
        while (p < size && d[p]) ++p;
        ++p;
    }
    if (flg & 16) {  // FCOMMENT
        while (p < size && d[p]) ++p;
        ++p;
    }
    if (flg & 2) p += 2;  // FHCRC
    return p < size ? p : 0;
}

// CRC-32 (the gzip polynomial, bit-reflected) by carry-less multiplication: four 128-bit lanes folded 64 bytes at a time,
// then down to 128, 64 and (Barrett) 32 bits -- V. Gopal et al., "Fast CRC computation for generic polynomials using
// PCLMULQDQ", Intel 2009; the constants are the paper's for this polynomial (x^(512+32), x^(512-32), x^(128+32),
// x^(128-32), x^64 mod P, then P and its inverse).  ~10x zlib's table-driven loop, which otherwise costs as much as the
// inflate itself.  len: a multiple of 16, >= 64.  crc: the raw register (zlib's value, inverted).
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(const uint8_t* buf, uint64_t len, uint32_t crc) {
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124);
    const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
    __m128i x1 = _mm_loadu_si128((const __m128i*)(buf + 0)), x2 = _mm_loadu_si128((const __m128i*)(buf + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i*)(buf + 32)), x4 = _mm_loadu_si128((const __m128i*)(buf + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    buf += 64;
    len -= 64;
    while (len >= 64) {
        const __m128i a1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), a2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        const __m128i a3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), a4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
        x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
        x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, a1), _mm_loadu_si128((const __m128i*)(buf + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, a2), _mm_loadu_si128((const __m128i*)(buf + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, a3), _mm_loadu_si128((const __m128i*)(buf + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, a4), _mm_loadu_si128((const __m128i*)(buf + 48)));
        buf += 64;
        len -= 64;
    }
#define VG_CRC_FOLD(into)                                            \
    do {                                                             \
        const __m128i a_ = _mm_clmulepi64_si128(x1, k3k4, 0x00);     \
        x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);                   \
        x1 = _mm_xor_si128(_mm_xor_si128(x1, (into)), a_);           \
    } while (0)
    VG_CRC_FOLD(x2);
    VG_CRC_FOLD(x3);
    VG_CRC_FOLD(x4);
    while (len >= 16) {
        VG_CRC_FOLD(_mm_loadu_si128((const __m128i*)buf));
        buf += 16;
        len -= 16;
    }
#undef VG_CRC_FOLD
    const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
    t = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, mask32);
    x1 = _mm_xor_si128(_mm_clmulepi64_si128(x1, k5, 0x00), t);
    t = _mm_and_si128(x1, mask32);
    t = _mm_clmulepi64_si128(t, poly, 0x10);
    t = _mm_and_si128(t, mask32);
    t = _mm_clmulepi64_si128(t, poly, 0x00);
    x1 = _mm_xor_si128(x1, t);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}

}  // namespace
// zlib's crc32(crc, buf, len) for any length
uint32_t crc32_fast(uint32_t crc, const uint8_t* buf, uint64_t len) {
    static const bool clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (clmul && len >= 64) {
        const uint64_t body = len & ~15ull;
        crc = ~crc32_clmul(buf, body, ~crc);
        buf += body;
        len -= body;
    }
    while (len) {  // zlib takes 32-bit lengths
        const uint64_t step = std::min<uint64_t>(len, 1u << 30);
        crc = (uint32_t)crc32(crc, buf, (uInt)step);
        buf += step;
        len -= step;
    }
    return crc;
}
namespace {

struct MemberEnd {
    uint64_t out_pos;  // symbols of this chunk in front of the member's end
    uint32_t crc, isize;
};

struct Chunk {
    uint64_t first_byte = 0;          // the chunk's share of the file starts here ...
    uint64_t start_bit = ~0ull;       // ... and this is where a block starts (searched, or known for the first one)
    bool fresh_member = false;        // start_bit is the first block of a member: nothing can be referenced in front of it
    // filled by the worker
    uint16_t* sym = nullptr;
    uint64_t nsym = 0, cap = 0;
    uint64_t end_bit = 0;             // where it stopped: the start of chunk `next`, or the end of the data
    int next = -1;                    // index of the chunk it arrived at (-1: end of batch / of file)
    bool eof = false;                 // the last member ended and nothing that looks like another follows
    int err = 0;
    int64_t last_member_start = -1;   // symbols in front of the first block of the last member that began inside (-1: none)
    uint32_t min_marker = 32768;      // lowest window offset referenced
    std::vector<MemberEnd> ends;
    // filled by the chain pass
    bool in_chain = false;
    uint64_t out_off = 0;             // where its bytes go in the round's output
    uint8_t window[32768];            // the 32 KiB in front of it, resolved (index 32767 = the byte right before it)
    std::vector<uint32_t> seg_crc;    // CRC-32 of its bytes between member ends
    ~Chunk() { free(sym); }
    void reset() {  // for the next round; the symbol buffer (and the pages behind it) are kept
        start_bit = ~0ull;
        fresh_member = false;
        nsym = 0;
        end_bit = 0;
        next = -1;
        eof = false;
        err = 0;
        last_member_start = -1;
        min_marker = 32768;
        ends.clear();
        in_chain = false;
        out_off = 0;
        seg_crc.clear();
    }
};

inline bool grow(Chunk& c, uint64_t need) {
    if (c.nsym + need <= c.cap) return true;
    uint64_t ncap = std::max<uint64_t>(c.cap * 2, c.nsym + need + (1u << 20));
    uint16_t* p = (uint16_t*)realloc(c.sym, ncap * sizeof(uint16_t));
    if (!p) return false;
    c.sym = p;
    c.cap = ncap;
    return true;
}

enum { kErrNone = 0, kErrCorrupt = 1, kErrMemory = 2 };

// Inflate from c.start_bit until a block boundary that is the start of a later chunk (targets: ascending start bits of the
// chunks after this one, ~0 where none was found), or -- past the last target -- the first boundary at or after
// `stop_byte`, or the end of the data.
void inflate_chunk(const uint8_t* data, uint64_t size, Chunk& c, const std::vector<Chunk*>& later, uint64_t stop_byte, Tables& t) {
    Bits b{data, size, 0, 0, 0};
    b.seek(c.start_bit);
    size_t ti = 0;
    int64_t floor = c.fresh_member ? 0 : -32768;  // lowest symbol index a copy may reach
    if (c.fresh_member) c.last_member_start = 0;
    bool fixed_built = false;
    Tables* fixed = nullptr;
    auto fail = [&](int e) { c.err = e; };
    for (;;) {
        // ---- at a block boundary ----
        const uint64_t here = b.bitpos();
        while (ti < later.size() && (later[ti]->start_bit == ~0ull || later[ti]->start_bit < here)) ++ti;  // starts nobody arrives at
        if (ti < later.size() && later[ti]->start_bit == here && c.nsym > 0) {
            c.end_bit = here;
            c.next = (int)ti;  // index into `later`
            break;
        }
        if (ti >= later.size() && (here >> 3) >= stop_byte && c.nsym > 0) {
            c.end_bit = here;
            break;
        }
        b.refill();
        if (b.overrun()) { fail(kErrCorrupt); break; }
        const uint32_t hdr = b.take(3);
        const bool final_block = hdr & 1u;
        const uint32_t type = hdr >> 1;
        if (type == 3) { fail(kErrCorrupt); break; }
        if (type == 0) {  // stored
            b.drop(b.cnt & 7);
            b.refill();
            const uint32_t len = b.take(16), nlen = b.take(16);
            if ((len ^ 0xffffu) != nlen) { fail(kErrCorrupt); break; }
            uint64_t at = b.bitpos() >> 3;
            if (at + len > size) { fail(kErrCorrupt); break; }
            if (!grow(c, len)) { fail(kErrMemory); break; }
            for (uint32_t i = 0; i < len; ++i) c.sym[c.nsym + i] = data[at + i];
            c.nsym += len;
            b.seek((at + len) * 8);
        } else {
            Tables* tt = &t;
            if (type == 1) {
                if (!fixed_built) {
                    fixed = new Tables;
                    uint8_t ll[288], dl[32];
                    fixed_lengths(ll, dl);
                    build_table(ll, 288, kLitBits, fixed->lit, kLitTabSize, false, lit_entry);
                    build_table(dl, 32, kDistBits, fixed->dist, kDistTabSize, true, dist_entry);
                    fixed_built = true;
                }
                tt = fixed;
            } else if (!read_dynamic_header(b, t, false)) {
                fail(kErrCorrupt);
                break;
            }
            const uint32_t* lt = tt->lit;
            const uint32_t* dt = tt->dist;
            // the hot loop works on locals (the chunk's fields are written back behind it)
            uint16_t* sym = c.sym;
            uint64_t n = c.nsym, cap = c.cap;
            uint32_t min_marker = c.min_marker;
            const uint64_t pos_limit = size + 16;
            bool done = false;
            while (!done) {
                if (n + 258 + 16 > cap) {
                    c.nsym = n;
                    if (!grow(c, 258 + 16)) { fail(kErrMemory); break; }
                    sym = c.sym;
                    cap = c.cap;
                }
                b.refill();
                if (b.pos > pos_limit) { fail(kErrCorrupt); break; }  // reading zeros past the end of a truncated file
                uint32_t e = lt[b.peek(kLitBits)];
                if (e_kind(e) == kLiteral) {  // up to three literals from one refill (>= 56 bits: 3 x 10 fit)
                    b.drop((int)e_bits(e));
                    sym[n++] = (uint16_t)e_val(e);
                    e = lt[b.peek(kLitBits)];
                    if (e_kind(e) != kLiteral) continue;
                    b.drop((int)e_bits(e));
                    sym[n++] = (uint16_t)e_val(e);
                    e = lt[b.peek(kLitBits)];
                    if (e_kind(e) != kLiteral) continue;
                    b.drop((int)e_bits(e));
                    sym[n++] = (uint16_t)e_val(e);
                    continue;
                }
                if (e_kind(e) == kSub) {
                    b.drop(kLitBits);
                    e = lt[e_val(e) + b.peek((int)e_extra(e))];
                }
                b.drop((int)e_bits(e));
                const uint32_t kind = e_kind(e);
                if (kind == kLiteral) {
                    sym[n++] = (uint16_t)e_val(e);
                    continue;
                }
                if (kind == kLength) {
                    const uint32_t len = e_val(e) + b.take((int)e_extra(e));
                    uint32_t d = dt[b.peek(kDistBits)];
                    if (e_kind(d) == kSub) {
                        b.drop(kDistBits);
                        d = dt[e_val(d) + b.peek((int)e_extra(d))];
                    }
                    if (e_kind(d) != kDist) { fail(kErrCorrupt); break; }
                    b.drop((int)e_bits(d));
                    if (b.cnt < 13) b.refill();
                    const uint32_t dist = e_val(d) + b.take((int)e_extra(d));
                    const int64_t src = (int64_t)n - (int64_t)dist;
                    if (src < floor) { fail(kErrCorrupt); break; }
                    uint16_t* out = sym + n;
                    if (src >= 0) {
                        const uint16_t* in = sym + src;
                        if (dist >= 8) {  // eight symbols at a time; may write up to 7 past the match (the buffer has the room)
                            for (uint32_t i = 0; i < len; i += 8) memcpy(out + i, in + i, 16);
                        } else {
                            for (uint32_t i = 0; i < len; ++i) out[i] = in[i];
                        }
                    } else {
                        if ((uint32_t)(32768 + src) < min_marker) min_marker = (uint32_t)(32768 + src);
                        for (uint32_t i = 0; i < len; ++i) {
                            const int64_t sp = src + (int64_t)i;
                            out[i] = sp >= 0 ? sym[sp] : (uint16_t)(0x8000u | (uint32_t)(32768 + sp));
                        }
                    }
                    n += len;
                    continue;
                }
                if (kind == kEob) { done = true; continue; }
                fail(kErrCorrupt);
                break;
            }
            c.nsym = n;
            c.min_marker = min_marker;
            if (c.err) break;
        }
        if (b.overrun()) { fail(kErrCorrupt); break; }
        if (final_block) {  // end of a gzip member: trailer, then maybe another member
            const uint64_t at = (b.bitpos() + 7) >> 3;
            if (at + 8 > size) { fail(kErrCorrupt); break; }
            uint32_t crc, isize;
            memcpy(&crc, data + at, 4);
            memcpy(&isize, data + at + 4, 4);
            c.ends.push_back({c.nsym, crc, isize});
            const uint64_t next_data = parse_member_header(data, size, at + 8);
            if (!next_data) {  // end of file, or trailing garbage (zlib's gzread ignores it as well)
                c.end_bit = size * 8;
                c.eof = true;
                break;
            }
            b.seek(next_data * 8);
            floor = (int64_t)c.nsym;
            c.last_member_start = (int64_t)c.nsym;
        }
    }
    delete fixed;
}

}  // namespace

// n 16-bit symbols -> bytes through the table above.  Most of a FASTQ chunk's symbols are literals (markers survive
// where text was copied, copy after copy, from the unknown window: the read names' common prefix), so 64 symbols with
// no marker among them are narrowed with two vector instructions; the others take the table, branch-free either way
// (the per-symbol "marker?" branch this replaces mispredicted its way to 0.4 GB/s).
__attribute__((target("avx512bw"))) static void resolve_symbols_avx512(const uint16_t* sy, uint64_t n, const uint8_t* tab, uint8_t* o) {
    uint64_t j = 0;
    for (; j + 64 <= n; j += 64) {
        const __m512i a = _mm512_loadu_si512((const void*)(sy + j)), b = _mm512_loadu_si512((const void*)(sy + j + 32));
        if (_mm512_movepi16_mask(_mm512_or_si512(a, b)) == 0) {
            _mm256_storeu_si256((__m256i*)(o + j), _mm512_cvtepi16_epi8(a));
            _mm256_storeu_si256((__m256i*)(o + j + 32), _mm512_cvtepi16_epi8(b));
        } else {
            for (int i = 0; i < 64; ++i) o[j + i] = tab[sy[j + i]];
        }
    }
    for (; j < n; ++j) o[j] = tab[sy[j]];
}
static void resolve_symbols(const uint16_t* sy, uint64_t n, const uint8_t* tab, uint8_t* o) {
    static const bool avx512 = __builtin_cpu_supports("avx512bw");
    if (avx512) return resolve_symbols_avx512(sy, n, tab, o);
    uint64_t j = 0;
    for (; j + 8 <= n; j += 8)
        for (int i = 0; i < 8; ++i) o[j + i] = tab[sy[j + i]];
    for (; j < n; ++j) o[j] = tab[sy[j]];
}

struct Stream::Impl {
    const uint8_t* data;
    uint64_t size;
    int threads;
    uint64_t chunk_bytes;
    uint64_t next_bit;                // where the next round starts (a block boundary)
    bool fresh = true;                // ... and whether that is the first block of a member
    bool eof = false;
    uint8_t window[32768];            // the last 32 KiB of output so far
    uint64_t window_valid = 0;        // bytes of it that belong to the current member (what a copy may reach), <= 32768
    uint32_t run_crc = 0;             // CRC-32 / length of the current member so far
    uint64_t run_len = 0;
    std::string error;
    Scratch* scratch = nullptr;       // the chunk pool lives here (the caller's, or our own)
    bool own_scratch = false;
    ~Impl() {
        if (own_scratch) delete scratch;
    }
};

// The workers' symbol buffers, reused round after round and -- when the caller keeps the Scratch -- file after file:
// fresh pages cost more than the inflate itself.
struct Scratch::Pool {
    std::vector<Chunk*> chunks;
    ~Pool() {
        for (Chunk* c : chunks) delete c;
    }
};
Scratch::Scratch() : pool_(new Pool) {}
Scratch::~Scratch() { delete pool_; }

Stream::Stream(const uint8_t* data, uint64_t size, int threads, uint64_t chunk_bytes, Scratch* scratch) : impl_(new Impl) {
    impl_->scratch = scratch ? scratch : new Scratch;
    impl_->own_scratch = scratch == nullptr;
    impl_->data = data;
    impl_->size = size;
    impl_->threads = std::max(1, threads);
    impl_->chunk_bytes = std::max<uint64_t>(chunk_bytes, 1024);
    const uint64_t d = parse_member_header(data, size, 0);
    if (!d) {
        impl_->error = "not a gzip file";
        impl_->eof = true;
        impl_->next_bit = 0;
    } else {
        impl_->next_bit = d * 8;
    }
    memset(impl_->window, 0, sizeof(impl_->window));
}
Stream::~Stream() { delete impl_; }
bool Stream::eof() const { return impl_->eof; }
const std::string& Stream::error() const { return impl_->error; }

// One round: the next `threads * per_thread` chunks of compressed bytes, inflated in parallel; their text is appended to
// `out`.  false on error (error() says what).
bool Stream::next(Buffer& out, int per_thread) {
    Impl& s = *impl_;
    if (!s.error.empty()) return false;
    if (s.eof) return true;
    const int nchunks = std::max(1, s.threads * std::max(1, per_thread));
    const uint64_t first_byte = s.next_bit >> 3;
    std::vector<Chunk*> chunks;
    for (int i = 0; i < nchunks; ++i) {
        const uint64_t fb = first_byte + (uint64_t)i * s.chunk_bytes;
        if (i > 0 && fb >= s.size) break;
        std::vector<Chunk*>& pool = s.scratch->pool_->chunks;
        if ((int)pool.size() <= i) pool.push_back(new Chunk);
        Chunk* c = pool[(size_t)i];
        c->reset();
        c->first_byte = fb;
        chunks.push_back(c);
    }
    const uint64_t stop_byte = first_byte + (uint64_t)chunks.size() * s.chunk_bytes;
    chunks[0]->start_bit = s.next_bit;
    chunks[0]->fresh_member = s.fresh;
    const int n = (int)chunks.size();
    const int nthreads = std::min(s.threads, n);
    auto parallel = [&](auto fn) {
        std::atomic<int> next{0};
        std::vector<std::thread> pool;
        auto body = [&] {
            Tables* t = new Tables;
            for (int i; (i = next.fetch_add(1)) < n;) fn(i, *t);
            delete t;
        };
        for (int w = 1; w < nthreads; ++w) pool.emplace_back(body);
        body();
        for (auto& th : pool) th.join();
    };
    const bool dbg = getenv("VG_GZ_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t0 = now();
    // 1. where can the chunks after the first start?
    parallel([&](int i, Tables& t) {
        if (i == 0) return;
        Chunk& c = *chunks[(size_t)i];
        const uint64_t lo = c.first_byte * 8, hi = std::min(s.size, c.first_byte + s.chunk_bytes) * 8;
        for (uint64_t bit = lo; bit < hi; ++bit)
            if (plausible_block_start(s.data, s.size, bit, t)) {
                c.start_bit = bit;
                break;
            }
    });
    const auto t1 = now();
    // 2. inflate
    parallel([&](int i, Tables& t) {
        Chunk& c = *chunks[(size_t)i];
        if (c.start_bit == ~0ull) return;
        std::vector<Chunk*> later(chunks.begin() + i + 1, chunks.end());
        inflate_chunk(s.data, s.size, c, later, stop_byte, t);
        if (c.next >= 0) c.next += i + 1;
    });
    const auto t2 = now();
    // 3. follow the chain from the first chunk: windows, output offsets
    bool ok = true;
    std::vector<int> chain;
    uint64_t total = 0;
    {
        std::vector<uint8_t> win(s.window, s.window + 32768), nw(32768);
        uint64_t valid = s.window_valid;
        for (int i = 0; i >= 0;) {
            Chunk& c = *chunks[(size_t)i];
            if (c.err) {
                s.error = c.err == kErrMemory ? "out of memory while inflating" : "corrupt deflate data";
                ok = false;
                break;
            }
            if (c.min_marker < 32768 - valid) {  // a copy reaches in front of the member's first byte
                s.error = "corrupt deflate data (distance too far back)";
                ok = false;
                break;
            }
            c.in_chain = true;
            c.out_off = total;
            memcpy(c.window, win.data(), 32768);
            chain.push_back(i);
            total += c.nsym;
            // the window behind it: its last 32 KiB, resolved
            const uint64_t take = std::min<uint64_t>(c.nsym, 32768);
            if (take < 32768) memcpy(nw.data(), win.data() + take, 32768 - take);
            for (uint64_t k = 0; k < take; ++k) {
                const uint16_t v = c.sym[c.nsym - take + k];
                nw[32768 - take + k] = (v & 0x8000u) ? win[v & 0x7fffu] : (uint8_t)v;
            }
            win.swap(nw);
            valid = c.last_member_start >= 0 ? std::min<uint64_t>(32768, c.nsym - (uint64_t)c.last_member_start)
                                             : std::min<uint64_t>(32768, valid + c.nsym);
            if (c.next >= 0) {
                i = c.next;
            } else {
                memcpy(s.window, win.data(), 32768);
                s.window_valid = valid;
                s.next_bit = c.end_bit;
                s.fresh = c.last_member_start == (int64_t)c.nsym;
                s.eof = c.eof;
                i = -1;
            }
        }
    }
    const auto t3 = now();
    // 4. resolve every chunk of the chain into place, CRC by member segment
    if (ok) {
        const uint64_t base = out.size;
        if (!out.reserve(base + total)) {
            s.error = "out of memory while inflating";
            return false;
        }
        out.size = base + total;
        uint8_t* dst = out.data + base;
        const int nc = (int)chain.size();
        std::atomic<int> next{0};
        auto body = [&] {
            std::vector<uint8_t> tab(65536);
            for (int i = 0; i < 256; ++i) tab[(size_t)i] = (uint8_t)i;
            for (int k; (k = next.fetch_add(1)) < nc;) {
                Chunk& c = *chunks[(size_t)chain[(size_t)k]];
                uint8_t* o = dst + c.out_off;
                const uint16_t* sy = c.sym;
                // symbol -> byte as ONE table: a literal maps to itself, marker 0x8000 | o to byte o of the chunk's window
                memcpy(tab.data() + 0x8000, c.window, 32768);
                // tile by tile, so that the CRC reads the bytes while they are still in the cache
                uint64_t from = 0;
                for (size_t m = 0; m <= c.ends.size(); ++m) {
                    const uint64_t to = m < c.ends.size() ? c.ends[m].out_pos : c.nsym;
                    uint32_t crc = 0;
                    for (uint64_t p = from; p < to;) {
                        const uint64_t e = std::min<uint64_t>(to, p + (64u << 10));
                        resolve_symbols(sy + p, e - p, tab.data(), o + p);
                        crc = crc32_fast(crc, o + p, e - p);
                        p = e;
                    }
                    c.seg_crc.push_back(crc);
                    from = to;
                }
            }
        };
        std::vector<std::thread> pool;
        for (int w = 1; w < std::min(s.threads, nc); ++w) pool.emplace_back(body);
        body();
        for (auto& th : pool) th.join();
        for (int k = 0; k < nc && ok; ++k) {
            Chunk& c = *chunks[(size_t)chain[(size_t)k]];
            uint64_t from = 0;
            for (size_t m = 0; m <= c.ends.size(); ++m) {
                const uint64_t to = m < c.ends.size() ? c.ends[m].out_pos : c.nsym;
                s.run_crc = (uint32_t)crc32_combine(s.run_crc, c.seg_crc[m], (z_off_t)(to - from));
                s.run_len += to - from;
                if (m < c.ends.size()) {
                    if (s.run_crc != c.ends[m].crc || (uint32_t)s.run_len != c.ends[m].isize) {
                        s.error = "gzip member fails its CRC-32 / length check";
                        ok = false;
                        break;
                    }
                    s.run_crc = 0;
                    s.run_len = 0;
                }
                from = to;
            }
        }
        if (!ok) out.size = base;
    }
    if (dbg)
        fprintf(stderr, "[vg_gzip] round of %d chunks (%zu in chain): search %.1f ms, inflate %.1f ms, chain %.1f ms, resolve + crc %.1f ms, %.1f MB out\n",
                n, chain.size(), ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, now()), total / 1e6);
    return ok;
}

}  // namespace gz
}  // namespace vg

// Host-only test hook (needs no GPU): inflate a whole gzip file with `threads` workers and chunks of `chunk_bytes`
// compressed bytes.  *out is malloc'ed (vg_gunzip_free).  0 ok, -1 cannot read, -2 inflate error.
extern "C" int vg_gunzip_parallel(const char* path, int threads, uint64_t chunk_bytes, uint8_t** out, uint64_t* out_len) {
    if (!path || !out || !out_len) return -1;
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    std::vector<uint8_t> in;
    uint8_t tmp[1 << 16];
    for (size_t r; (r = fread(tmp, 1, sizeof(tmp), f)) > 0;) in.insert(in.end(), tmp, tmp + r);
    fclose(f);
    vg::gz::Stream st(in.data(), in.size(), threads, chunk_bytes);
    vg::gz::Buffer text;
    while (!st.eof())
        if (!st.next(text, 2)) break;
    if (!st.error().empty()) {
        if (getenv("VG_GZ_DEBUG")) fprintf(stderr, "vg_gunzip_parallel: %s\n", st.error().c_str());
        return -2;
    }
    if (!text.data && !text.reserve(1)) return -2;
    *out = text.data;  // handed over
    *out_len = text.size;
    text.data = nullptr;
    return 0;
}
extern "C" void vg_gunzip_free(uint8_t* p) { free(p); }

// Host-only test hook: the CRC-32 the inflater checks members with (carry-less multiply where the CPU has it); must equal zlib's.
extern "C" uint32_t vg_crc32(uint32_t crc, const uint8_t* buf, uint64_t len) { return vg::gz::crc32_fast(crc, buf, len); }
