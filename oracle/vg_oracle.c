/* vg_oracle.c -- plain-C restatement of varigraph's read k-mer counting path
 * (and of the construct-side counting Bloom filter that shares its encoder).
 *
 * TEST INFRASTRUCTURE ONLY -- see vg_oracle.h.  Parity status: PINNED against
 * the unmodified reference (oracle/_ref/libvgref.so) and tests/golden/.
 *
 * Reference behaviour each function follows (paths relative to /root/reference):
 *   vgo_nt4                  include/seq_nt4_table.hpp:5-22
 *   vgo_hash64               include/hash64.hpp:5-14
 *   roll_* / vgo_sketch      src/kmer.cpp:110-149  (same loop at :20-52, :65-97, :162-198)
 *   vgo_murmur3_x64_128_sum  src/MurmurHash3.cpp:255-332 restricted to len == 8,
 *                            summed as src/counting_bloom_filter.cpp:90-98 does
 *   vgo_cbf_size/num_hashes  src/counting_bloom_filter.cpp:70-77
 *   vgo_cbf_add/count/find   src/counting_bloom_filter.cpp:28-67
 *   vgo_cbf_fill             src/kmer.cpp:20-52 called per chromosome by src/construct_index.cpp:161-166
 *   vgo_count_*              src/fastq_kmer.cpp:97-141 + :314-332 (c = min(255, c + 1) per emitted key present)
 *   vgo_fastq_to_lines       include/kseq.h:192-232 over an in-memory text (natural EOF; the
 *                            16 KiB refill quirk at exact multiples of the buffer size is not modelled)
 */
#include "vg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ---- alphabet ------------------------------------------------------------ */
uint8_t vgo_nt4(uint8_t b) {
    switch (b) {
        case 0: case 'A': case 'a': return 0;
        case 1: case 'C': case 'c': return 1;
        case 2: case 'G': case 'g': return 2;
        case 3: case 'T': case 't': case 'U': case 'u': return 3;
        default: return 4;
    }
}

/* ---- hash64: invertible mix, masked to 2k bits after every add ----------- */
uint64_t vgo_hash64(uint64_t x, uint64_t m) {
    x = (~x + (x << 21)) & m;
    x ^= x >> 24;
    x = (x + (x << 3) + (x << 8)) & m;
    x ^= x >> 14;
    x = (x + (x << 2) + (x << 4)) & m;
    x ^= x >> 28;
    x = (x + (x << 31)) & m;
    return x;
}

/* ---- rolling canonical k-mer state machine -------------------------------
 * One instance per read.  fwd/rev are NOT cleared by an ambiguous base (only
 * the run length is), and a base whose fwd == rev neither counts towards the
 * run nor emits: both quirks matter for even k. */
typedef struct {
    uint64_t fwd, rev, mask;
    uint32_t k, run, top_shift;
} roll_t;

static void roll_init(roll_t* r, uint32_t k) {
    r->fwd = r->rev = 0;
    r->run = 0;
    r->k = k;
    r->mask = (1ULL << (2 * k)) - 1;
    r->top_shift = 2 * (k - 1);
}

/* Returns 1 and sets *key when this byte completes an emitted k-mer. */
static int roll_push(roll_t* r, uint8_t byte, uint64_t* key) {
    uint8_t c = vgo_nt4(byte);
    if (c > 3) {
        r->run = 0;
        return 0;
    }
    r->fwd = ((r->fwd << 2) | c) & r->mask;
    r->rev = (r->rev >> 2) | ((uint64_t)(3 ^ c) << r->top_shift);
    if (r->fwd == r->rev) return 0;
    r->run += 1;
    if (r->run < r->k) return 0;
    uint64_t canon = r->fwd < r->rev ? r->fwd : r->rev;
    *key = (vgo_hash64(canon, r->mask) << 8) | r->k;
    return 1;
}

int64_t vgo_sketch(const char* seq, int64_t len, uint32_t k, uint64_t* out, int64_t cap) {
    roll_t r;
    roll_init(&r, k);
    int64_t n = 0;
    for (int64_t i = 0; i < len; ++i) {
        uint64_t key;
        if (roll_push(&r, (uint8_t)seq[i], &key)) {
            if (n < cap) out[n] = key;
            ++n;
        }
    }
    return n;
}

/* Per-byte view of a staged chunk ('\n' separates reads): out[i] = key of the k-mer the encoder
 * emits when it consumes byte i, or ~0.  Returns the number of emitted keys. */
int64_t vgo_positions(const char* buf, int64_t nbytes, uint32_t k, uint64_t* out) {
    roll_t r;
    roll_init(&r, k);
    int64_t n = 0;
    for (int64_t i = 0; i < nbytes; ++i) {
        uint64_t key;
        out[i] = ~0ULL;
        if (buf[i] == '\n') {
            roll_init(&r, k); /* a new read: fresh registers (kmer_sketch_fastq is called per read) */
            continue;
        }
        if (roll_push(&r, (uint8_t)buf[i], &key)) {
            out[i] = key;
            ++n;
        }
    }
    return n;
}

/* ---- MurmurHash3 x64_128 for one 8-byte key ------------------------------ */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}
uint64_t vgo_murmur3_x64_128_sum(uint64_t key, uint32_t seed) {
    uint64_t h1 = seed, h2 = seed;
    /* len 8: no 16-byte body block; the tail holds the whole key in k1 */
    uint64_t k1 = key * 0x87c37b91114253d5ULL;
    k1 = rotl64(k1, 31) * 0x4cf5ad432745937fULL;
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

/* ---- counting Bloom filter ----------------------------------------------- */
uint64_t vgo_cbf_size(uint64_t n, double p) {
    return (uint64_t)ceil(((double)n * log(p)) / log(1.0 / pow(2.0, log(2.0))));
}
uint32_t vgo_cbf_num_hashes(uint64_t n, uint64_t m) {
    return (uint32_t)round((double)m * log(2.0) / (double)n);
}
static inline uint64_t cbf_cell(uint64_t m, uint64_t seed, uint64_t key) {
    return vgo_murmur3_x64_128_sum(key, (uint32_t)seed) % m; /* seed truncated to 32 bit */
}
void vgo_cbf_add(uint8_t* f, uint64_t m, const uint64_t* seeds, uint32_t nh, uint64_t key) {
    for (uint32_t i = 0; i < nh; ++i) {
        uint64_t p = cbf_cell(m, seeds[i], key);
        if (f[p] != 255) f[p] += 1;
    }
}
uint8_t vgo_cbf_count(const uint8_t* f, uint64_t m, const uint64_t* seeds, uint32_t nh, uint64_t key) {
    uint8_t lo = 255;
    for (uint32_t i = 0; i < nh; ++i) {
        uint8_t v = f[cbf_cell(m, seeds[i], key)];
        if (v < lo) lo = v;
    }
    return lo;
}
int vgo_cbf_find(const uint8_t* f, uint64_t m, const uint64_t* seeds, uint32_t nh, uint64_t key) {
    for (uint32_t i = 0; i < nh; ++i)
        if (f[cbf_cell(m, seeds[i], key)] == 0) return 0;
    return 1;
}
uint64_t vgo_cbf_fill(uint8_t* f, uint64_t m, const uint64_t* seeds, uint32_t nh, const char* seq,
                      int64_t len, uint32_t k) {
    roll_t r;
    roll_init(&r, k);
    uint64_t n = 0;
    for (int64_t i = 0; i < len; ++i) {
        uint64_t key;
        if (roll_push(&r, (uint8_t)seq[i], &key)) {
            vgo_cbf_add(f, m, seeds, nh, key);
            ++n;
        }
    }
    return n;
}

/* ---- the graph k-mer index (stand-in for mGraphKmerHashHapStrMap) -------- */
typedef struct {
    uint64_t cap; /* power of two */
    uint64_t n;
    uint64_t* key;
    int64_t* id;
} oidx_t;

static inline uint64_t oidx_home(const oidx_t* t, uint64_t key) {
    return (key * 0x9E3779B97F4A7C15ULL) >> 20 & (t->cap - 1);
}

void* vgo_index_create(const uint64_t* keys, uint64_t n) {
    oidx_t* t = (oidx_t*)calloc(1, sizeof(oidx_t));
    t->cap = 16;
    while (t->cap < 2 * n + 1) t->cap <<= 1;
    t->n = n;
    t->key = (uint64_t*)malloc(t->cap * sizeof(uint64_t));
    t->id = (int64_t*)malloc(t->cap * sizeof(int64_t));
    for (uint64_t i = 0; i < t->cap; ++i) t->id[i] = -1;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t s = oidx_home(t, keys[i]);
        while (t->id[s] >= 0 && t->key[s] != keys[i]) s = (s + 1) & (t->cap - 1);
        if (t->id[s] < 0) { /* duplicates keep their first id */
            t->key[s] = keys[i];
            t->id[s] = (int64_t)i;
        }
    }
    return t;
}
void vgo_index_destroy(void* idx) {
    oidx_t* t = (oidx_t*)idx;
    if (!t) return;
    free(t->key);
    free(t->id);
    free(t);
}
int64_t vgo_index_find(const void* idx, uint64_t key) {
    const oidx_t* t = (const oidx_t*)idx;
    uint64_t s = oidx_home(t, key);
    while (t->id[s] >= 0) {
        if (t->key[s] == key) return t->id[s];
        s = (s + 1) & (t->cap - 1);
    }
    return -1;
}

/* One read.  counts[] is u8 per index entry (key order given at create),
 * saturating at 255.  Returns the number of emitted k-mer positions. */
uint64_t vgo_count_seq(const void* idx, const char* seq, int64_t len, uint32_t k, uint8_t* counts,
                       uint64_t* hits) {
    roll_t r;
    roll_init(&r, k);
    uint64_t emitted = 0, h = 0;
    for (int64_t i = 0; i < len; ++i) {
        uint64_t key;
        if (!roll_push(&r, (uint8_t)seq[i], &key)) continue;
        ++emitted;
        int64_t id = vgo_index_find(idx, key);
        if (id < 0) continue;
        ++h;
        if (counts[id] != 255) counts[id] += 1;
    }
    if (hits) *hits += h;
    return emitted;
}

/* A staged chunk: reads separated by '\n' (a trailing '\n' is optional).
 * Each line is an independent read, as in FastqKmer::fastq_file_open. */
uint64_t vgo_count_lines(const void* idx, const char* buf, int64_t nbytes, uint32_t k, uint8_t* counts,
                         uint64_t* hits, uint64_t* nreads) {
    uint64_t emitted = 0, reads = 0;
    int64_t start = 0;
    for (int64_t i = 0; i <= nbytes; ++i) {
        if (i == nbytes || buf[i] == '\n') {
            if (i > start) {
                emitted += vgo_count_seq(idx, buf + start, i - start, k, counts, hits);
                ++reads;
            }
            start = i + 1;
        }
    }
    if (nreads) *nreads = reads;
    return emitted;
}

/* ---- kseq-style FASTA/FASTQ reader over memory ---------------------------
 * Writes each record's sequence followed by '\n' into out (up to cap bytes;
 * the return value is the size needed).  *read_bases sums the sequence
 * lengths as mReadBase does (src/fastq_kmer.cpp:105).  *status: -1 = clean
 * EOF, -2 = stopped at a record whose quality is missing or of another
 * length (the reference's `while (kseq_read(ks) >= 0)` stops there too,
 * without counting that record). */
typedef struct { const unsigned char* p; int64_t n, i; } mstream_t;
typedef struct { char* s; int64_t l, m; } mstr_t;

static int ms_getc(mstream_t* s) { return s->i < s->n ? s->p[s->i++] : -1; }
static void mstr_put(mstr_t* d, const unsigned char* src, int64_t len) {
    if (d->l + len + 1 > d->m) {
        d->m = 2 * (d->l + len + 1);
        d->s = (char*)realloc(d->s, (size_t)d->m);
    }
    if (len > 0) memcpy(d->s + d->l, src, (size_t)len);
    d->l += len;
}
/* ks_getuntil2(ks, KS_SEP_LINE, str, 0, append): -1 when already at EOF. */
static int64_t ms_getline(mstream_t* s, mstr_t* d, int append) {
    if (!append) d->l = 0;
    if (s->i >= s->n) return -1;
    int64_t j = s->i;
    while (j < s->n && s->p[j] != '\n') ++j;
    mstr_put(d, s->p + s->i, j - s->i);
    s->i = j < s->n ? j + 1 : j;
    if (d->l > 1 && d->s[d->l - 1] == '\r') d->l -= 1;
    return d->l;
}

/* One kseq_read(): >= 0 sequence length, -1 EOF, -2 bad quality. */
static int64_t ms_read_record(mstream_t* s, int* last_char, mstr_t* seq, mstr_t* qual, mstr_t* scratch) {
    int c;
    if (*last_char == 0) {
        while ((c = ms_getc(s)) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        *last_char = c;
    }
    seq->l = qual->l = 0;
    if (s->i >= s->n) return -1; /* header char was the last byte */
    c = 0;
    while (s->i < s->n) { /* name: up to the first whitespace */
        int ch = s->p[s->i++];
        if (isspace(ch)) { c = ch; break; }
    }
    if (c != '\n') ms_getline(s, scratch, 0); /* comment */
    while ((c = ms_getc(s)) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        unsigned char b = (unsigned char)c;
        mstr_put(seq, &b, 1);
        ms_getline(s, seq, 1);
    }
    if (c == '>' || c == '@') *last_char = c;
    if (c != '+') return seq->l;
    while ((c = ms_getc(s)) != -1 && c != '\n') {}
    if (c == -1) return -2;
    while (ms_getline(s, qual, 1) >= 0 && qual->l < seq->l) {}
    *last_char = 0;
    if (seq->l != qual->l) return -2;
    return seq->l;
}

int64_t vgo_fastq_to_lines(const char* text, int64_t n, char* out, int64_t cap, uint64_t* nreads,
                           uint64_t* read_bases, int* status) {
    mstream_t s = {(const unsigned char*)text, n, 0};
    mstr_t seq = {0, 0, 0}, qual = {0, 0, 0}, scratch = {0, 0, 0};
    int64_t w = 0, r;
    uint64_t reads = 0, bases = 0;
    int last_char = 0;
    while ((r = ms_read_record(&s, &last_char, &seq, &qual, &scratch)) >= 0) {
        /* the reference builds a std::string from a C string: it stops at the first NUL */
        int64_t use = 0;
        while (use < seq.l && seq.s[use] != '\0') ++use;
        for (int64_t t = 0; t < use; ++t)
            if (w + t < cap) out[w + t] = seq.s[t];
        w += use;
        if (w < cap) out[w] = '\n';
        w += 1;
        bases += (uint64_t)seq.l;
        reads += 1;
    }
    free(seq.s);
    free(qual.s);
    free(scratch.s);
    if (nreads) *nreads = reads;
    if (read_bases) *read_bases = bases;
    if (status) *status = (int)r;
    return w;
}
