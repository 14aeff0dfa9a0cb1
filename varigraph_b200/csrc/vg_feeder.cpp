// vg_feeder.cpp -- FASTQ/FASTA (plain or gzip) -> pinned staging ring -> count kernels.
//
// Stands in for FastqKmer::fastq_file_open (src/fastq_kmer.cpp:65-187): the reference inflates and
// parses on its main thread (kseq over gzread, 16 KiB buffer), copies every read into a
// std::string, upper-cases it and batches `threads*100` reads per pool task.  Here each file
// gets its own inflate/parse worker that writes "sequence\n" records straight into a pinned
// chunk; the calling thread only issues cudaMemcpyAsync + kernel launches, so copy, kernel and
// parsing overlap and R1/R2 are read concurrently.
//
// Plain (not gzip) four-line FASTQ takes a shorter road, cut into record-aligned blocks that many
// workers handle at once.  Two ways to take a block, and by default (VG_FASTQ_ROAD=hybrid) both at once, because one
// is bound by the host cores and the other by PCIe: VG_STRIP_SHARE (default 0.65) of the blocks are stripped, the rest
// go to the device as raw text, in one submission order, so the pipeline runs at the sum of the two rates.
//   strip            the workers scan the memory-mapped file with vector compares (64 bytes per step), check every
//                    record against what kseq reads as a four-line record, and copy only the sequences into the
//                    pinned chunks: 0.49 bytes cross PCIe per byte of FASTQ, and the page cache is read in place
//                    (a pread() of the same pages runs at a third of the speed of a scan over the mapping);
//   device           the workers copy raw text from the mapping into the pinned chunks and the GPU finds the lines, checks
//                    the records and blanks everything but the sequences (fastq_*_kernel in vg_kernels.cu).
// Whatever fails the check -- multi-line records, FASTA, a truncated tail, NUL bytes -- is not counted from the
// offending record (strip) or block (device) on and is re-read with the kseq reader, so the result is the
// reference's for any input.  VG_FASTQ_ROAD=kseq (or VG_RAW_FASTQ=0) sends every file through the kseq reader.
//
// What counts as a read follows kseq (include/kseq.h:192-232) exactly: header at '@' or '>',
// sequence = every line up to one starting with '+', '>' or '@', one trailing CR stripped per
// line, quality must match the sequence length or the file stops there (return -2) without
// counting that record; mReadBase sums seq.l (src/fastq_kmer.cpp:105).  The reference builds a
// std::string from the C string, so a sequence is cut at its first NUL byte.
#include <fcntl.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vgb200.h"
#include "vg_gzip.h"
#include "vg_host.h"

namespace {

class KseqReader {
   public:
    explicit KseqReader(gzFile f) : f_(f), buf_(1u << 20) {}

    // >= 0: sequence length (seq filled); -1: end of file; -2: truncated / mismatched quality
    int64_t next(std::string& seq) {
        int c;
        if (last_char_ == 0) {
            while ((c = getc()) != -1 && c != '>' && c != '@') {}
            if (c == -1) return -1;
            last_char_ = c;
        }
        seq.clear();
        qual_.clear();
        if (!fill()) return -1;  // header char was the last byte of the file
        c = skip_name();
        if (c != '\n') skip_line();
        while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            seq.push_back((char)c);
            getline(seq);
        }
        if (c == '>' || c == '@') last_char_ = c;
        if (c != '+') return (int64_t)seq.size();
        while ((c = getc()) != -1 && c != '\n') {}
        if (c == -1) return -2;
        while (getline(qual_) >= 0 && qual_.size() < seq.size()) {}
        last_char_ = 0;
        if (qual_.size() != seq.size()) return -2;
        return (int64_t)seq.size();
    }

   private:
    bool fill() {  // true when at least one byte is available
        if (begin_ < end_) return true;
        if (eof_) return false;
        int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
        begin_ = 0;
        end_ = n > 0 ? (size_t)n : 0;
        if (n < (int)buf_.size()) eof_ = true;
        return end_ > 0;
    }
    int getc() { return fill() ? buf_[begin_++] : -1; }
    int skip_name() {  // consume up to and including the first whitespace; returns it (0 at EOF)
        while (fill()) {
            while (begin_ < end_) {
                int ch = buf_[begin_++];
                if (isspace(ch)) return ch;
            }
        }
        return 0;
    }
    void skip_line() {
        while (fill()) {
            const void* nl = memchr(buf_.data() + begin_, '\n', end_ - begin_);
            if (nl) {
                begin_ = (size_t)((const unsigned char*)nl - buf_.data()) + 1;
                return;
            }
            begin_ = end_;
        }
    }
    // append the rest of the current line; -1 when already at EOF (nothing appended, no CR strip)
    int64_t getline(std::string& s) {
        if (!fill()) return -1;
        for (;;) {
            const unsigned char* p = buf_.data() + begin_;
            const void* nl = memchr(p, '\n', end_ - begin_);
            size_t seg = nl ? (size_t)((const unsigned char*)nl - p) : end_ - begin_;
            s.append((const char*)p, seg);
            begin_ += seg + (nl ? 1 : 0);
            if (nl || !fill()) break;
        }
        if (s.size() > 1 && s.back() == '\r') s.pop_back();
        return (int64_t)s.size();
    }

    gzFile f_;
    std::vector<unsigned char> buf_;
    size_t begin_ = 0, end_ = 0;
    bool eof_ = false;
    int last_char_ = 0;
    std::string qual_;
};

// The consumers of one feeder: one index per GPU (one for vg_count_files).  A staging slot belongs to one of them; its
// id is sink * kSlotsPerSink + slot, and a chunk goes to the GPU whose slot a worker happened to fill.
constexpr int kSlotsPerSink = 4096;
struct Sinks {
    std::vector<vg_index*> ix;
    vg_ctx* ctx(int id) const { return ix[(size_t)(id / kSlotsPerSink)]->ctx; }
    vg_index* index(int id) const { return ix[(size_t)(id / kSlotsPerSink)]; }
    vg::StageSlot& slot(int id) const { return ctx(id)->ring[(size_t)(id % kSlotsPerSink)]; }
    size_t chunk_bytes() const { return ix[0]->ctx->chunk_bytes; }
};

struct Filled {
    int slot;
    uint64_t len;
    int64_t raw_item;  // >= 0: a block of a plain FASTQ file, item number in submission order; -1: "sequence\n" records
    uint64_t bad_at = ~0ull;  // strip road: file offset of the first record that is not plain four-line FASTQ
    uint64_t bases = 0;       // strip road: sum of seq.l over the records in front of it
};

struct KseqItem {  // a file (or its tail from a record boundary on) for the kseq reader
    std::string path;
    uint64_t offset;
};

struct RawFile {  // a plain four-line FASTQ file shipped as raw text
    std::string path;
    int fd = -1;
    uint64_t size = 0;
    std::vector<uint64_t> cut;           // record boundaries: block b = [cut[b], cut[b + 1])
    uint64_t tail_from = ~0ull;          // no boundary found beyond this one: the rest goes to the kseq reader
    const char* map = nullptr;           // strip road: the file, memory-mapped
    uint64_t bad_from = ~0ull;           // strip road: first irregular record found so far (submission order)
    bool borrowed = false;               // map is somebody else's memory (inflated gzip text): never unmapped here
    bool ends_at_eof = true;             // the text ends where the input ends (false: a window of inflated text, more follows)
};

struct RawItem {
    int file;
    uint32_t block;
    uint64_t start, end;
    bool last;  // ends at EOF: make sure the text ends with a newline, as kseq treats EOF
    bool strip; // stripped to its sequences by a host worker (else: shipped as raw text, parsed on the device)
};

struct Feeder {
    std::mutex mu;
    std::condition_variable cv_free, cv_ready;
    std::deque<int> free_q;
    std::deque<Filled> ready_q;
    int workers_left = 0;
    bool abort = false;
    int err = VG_OK;
    std::string err_msg;
    std::atomic<int> next_file{0};
    std::atomic<size_t> next_raw{0};
    std::atomic<uint64_t> read_bases{0};
    std::atomic<uint64_t> busy_ns{0};   // time the block workers spent working (VG_FEEDER_DEBUG)
    std::atomic<uint64_t> wait_ns{0};   // ... and waiting for a free staging slot

    int take_free() {
        std::unique_lock<std::mutex> lk(mu);
        if (free_q.empty() && !abort) {
            const auto t0 = std::chrono::steady_clock::now();
            cv_free.wait(lk, [&] { return abort || !free_q.empty(); });
            wait_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
        }
        if (abort) return -1;
        int s = free_q.front();
        free_q.pop_front();
        return s;
    }
    void give_free(int s) {
        {
            std::lock_guard<std::mutex> lk(mu);
            free_q.push_back(s);
        }
        cv_free.notify_one();
    }
    void push_ready(int s, uint64_t len, int64_t raw_item = -1) {
        {
            std::lock_guard<std::mutex> lk(mu);
            ready_q.push_back({s, len, raw_item});
        }
        cv_ready.notify_one();
    }
    void set_error(int code, const std::string& msg) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (err == VG_OK) {
                err = code;
                err_msg = msg;
            }
            abort = true;
        }
        cv_free.notify_all();
        cv_ready.notify_all();
    }
    void worker_done() {
        {
            std::lock_guard<std::mutex> lk(mu);
            workers_left -= 1;
        }
        cv_ready.notify_all();
    }
};

void kseq_worker(Feeder* fd, const Sinks* sinks, const std::vector<KseqItem>* items) {
    std::string seq;
    int slot = -1;
    uint64_t w = 0, bases = 0;
    const uint64_t cap = sinks->chunk_bytes();
    for (;;) {
        int fi = fd->next_file.fetch_add(1);
        if (fi >= (int)items->size()) break;
        const KseqItem& it = (*items)[(size_t)fi];
        gzFile gz = gzopen(it.path.c_str(), "rb");
        if (!gz) {
            fd->set_error(VG_E_IO, "'" + it.path + "': No such file or directory.");
            break;
        }
        gzbuffer(gz, 1u << 20);
        if (it.offset) gzseek(gz, (z_off_t)it.offset, SEEK_SET);
        KseqReader rd(gz);
        bool stop = false;
        while (rd.next(seq) >= 0) {
            bases += seq.size();
            uint64_t use = strnlen(seq.data(), seq.size());
            if (use == 0) continue;
            if (use + 1 > cap) {
                fd->set_error(VG_E_INVALID, "a read is longer than the staging buffer (raise --buffer)");
                stop = true;
                break;
            }
            if (slot >= 0 && w + use + 1 > cap) {
                fd->push_ready(slot, w);
                slot = -1;
            }
            if (slot < 0) {
                slot = fd->take_free();
                if (slot < 0) { stop = true; break; }
                w = 0;
            }
            uint8_t* dst = sinks->slot(slot).h_pin + w;
            memcpy(dst, seq.data(), (size_t)use);
            dst[use] = '\n';
            w += use + 1;
        }
        gzclose(gz);
        if (stop) break;
    }
    if (slot >= 0) {
        if (w > 0) fd->push_ready(slot, w);
        else fd->give_free(slot);
    }
    fd->read_bases.fetch_add(bases);
    fd->worker_done();
}

// ---- strip road: four-line FASTQ text -> "sequence\n" records, on the host -------------------------------
struct StripState {
    const char* rec;      // start of the current record
    const char* nl[4];    // the newlines that end its lines
    int li = 0;
    const char* nul_hi = nullptr;  // the last NUL byte seen so far (the scan runs ahead of the records by up to a vector)
    uint8_t* o;
    uint64_t bases = 0;
    const char* end;
};
// One record whose four line ends are known.  Accepted only if kseq (include/kseq.h:192-232) reads exactly these four
// lines as one record: '@' header, a sequence line that does not start like a header / separator, a '+' line, a
// quality line of the sequence's length (one trailing CR per line dropped, as ks_getuntil2 does), no NUL in the
// sequence (the reference's std::string would end there).
static inline __attribute__((always_inline)) bool strip_record(StripState& st) {
    const char* seq = st.nl[0] + 1;
    const char* plus = st.nl[1] + 1;
    const char* q = st.nl[2] + 1;
    const size_t sraw = (size_t)(st.nl[1] - seq), qraw = (size_t)(st.nl[3] - q);
    const size_t sl = (sraw > 1 && seq[sraw - 1] == '\r') ? sraw - 1 : sraw;
    const size_t ql = (qraw > 1 && q[qraw - 1] == '\r') ? qraw - 1 : qraw;
    const char s0 = sraw ? seq[0] : 'A';
    const bool ok = (st.rec[0] == '@') & (st.nl[0] != st.rec) & (st.nl[2] != plus) & (sl == ql) & (s0 != '@') & (s0 != '+') & (s0 != '>');
    if (!ok || plus[0] != '+') return false;
    if (st.nul_hi >= seq && memchr(seq, 0, sl)) return false;
    if (sl) {
        if (sl <= 192 && seq + 192 <= st.end) memcpy(st.o, seq, 192);  // three whole vectors (the chunk has slack behind it)
        else memcpy(st.o, seq, sl);
        st.o[sl] = '\n';
        st.o += sl + 1;
    }
    st.bases += sl;
    st.rec = st.nl[3] + 1;
    return true;
}
#define VG_STRIP_STEP(mask_nl, mask_zero, width)                      \
    while (p + (width) <= end) {                                      \
        uint64_t m = (mask_nl);                                       \
        const uint64_t mz = (mask_zero);                              \
        if (mz) st.nul_hi = p + (63 - __builtin_clzll(mz));           \
        while (m) {                                                   \
            st.nl[st.li] = p + __builtin_ctzll(m);                    \
            m &= m - 1;                                               \
            if (++st.li == 4) {                                       \
                st.li = 0;                                            \
                if (!strip_record(st)) return p = nullptr, false;     \
            }                                                         \
        }                                                             \
        p += (width);                                                 \
    }
__attribute__((target("avx512bw"))) static bool strip_scan_avx512(StripState& st, const char*& p, const char* end) {
    const __m512i nlv = _mm512_set1_epi8('\n'), zero = _mm512_setzero_si512();
    VG_STRIP_STEP(_mm512_cmpeq_epi8_mask(_mm512_loadu_si512((const void*)p), nlv),
                  _mm512_cmpeq_epi8_mask(_mm512_loadu_si512((const void*)p), zero), 64)
    return true;
}
__attribute__((target("avx2"))) static bool strip_scan_avx2(StripState& st, const char*& p, const char* end) {
    const __m256i nlv = _mm256_set1_epi8('\n'), zero = _mm256_setzero_si256();
    VG_STRIP_STEP((uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)p), nlv)),
                  (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)p), zero)), 32)
    return true;
}
#undef VG_STRIP_STEP
// [p0, end) starts at a record boundary -> "sequence\n" records at out; returns bytes written.  *bad = the first
// record that is not plain four-line FASTQ (nullptr: the whole block was).  `last`: the block ends at EOF, where the
// final newline may be missing.
static uint64_t strip_block(const char* p0, const char* end, bool last, uint8_t* out, uint64_t& bases, const char*& bad) {
    static const int isa = __builtin_cpu_supports("avx512bw") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    StripState st;
    st.rec = p0;
    st.o = out;
    st.end = end;
    const char* p = p0;
    bool ok = true;
    if (isa == 2) ok = strip_scan_avx512(st, p, end);
    else if (isa == 1) ok = strip_scan_avx2(st, p, end);
    for (; ok && p < end; ++p) {
        if (*p == 0) st.nul_hi = p;
        if (*p == '\n') {
            st.nl[st.li] = p;
            if (++st.li == 4) {
                st.li = 0;
                ok = strip_record(st);
            }
        }
    }
    if (ok && st.rec < end) {  // text behind the last complete record
        if (last && st.li == 3 && st.nl[2] + 1 < end) {
            st.nl[3] = end;
            st.li = 0;
            ok = strip_record(st);
        } else {
            ok = false;
        }
    }
    bases = st.bases;
    bad = ok ? nullptr : st.rec;
    return (uint64_t)(st.o - out);
}

// A slot first, then the next block: the lowest outstanding block always has a buffer, so the in-order
// submission of the calling thread cannot starve.
void block_worker(Feeder* fd, const Sinks* sinks, const std::vector<RawFile>* files, const std::vector<RawItem>* items) {
    for (;;) {
        int slot = fd->take_free();
        if (slot < 0) break;
        const size_t idx = fd->next_raw.fetch_add(1);
        if (idx >= items->size()) {
            fd->give_free(slot);
            break;
        }
        const RawItem& it = (*items)[idx];
        const RawFile& f = (*files)[(size_t)it.file];
        uint8_t* dst = sinks->slot(slot).h_pin;
        Filled out{slot, 0, (int64_t)idx};
        const auto t0 = std::chrono::steady_clock::now();
        if (it.strip) {
            const char* bad = nullptr;
            out.len = strip_block(f.map + it.start, f.map + it.end, it.last, dst, out.bases, bad);
            if (bad) out.bad_at = (uint64_t)(bad - f.map);
        } else {  // raw text for the device parser (a copy out of the mapping runs at 2.5x the speed of a pread)
            out.len = it.end - it.start;
            memcpy(dst, f.map + it.start, (size_t)out.len);
            if (it.last && out.len && dst[out.len - 1] != '\n') dst[out.len++] = '\n';
        }
        {   // give the block's pages back right away, here, in parallel: tearing down the whole mapping at the end costs
            // the calling thread ~20 ms per GB (the kseq fallback re-opens the file, it does not need the mapping)
            const uint64_t page = 4096, a = (it.start + page - 1) & ~(page - 1), b = it.end & ~(page - 1);
            if (b > a && !f.borrowed) munmap((void*)(f.map + a), (size_t)(b - a));
        }
        fd->busy_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
        {
            std::lock_guard<std::mutex> lk(fd->mu);
            fd->ready_q.push_back(out);
        }
        fd->cv_ready.notify_one();
    }
    fd->worker_done();
}

// First record start at or after `at`: a line that begins with '@' whose next-but-one line begins with '+'.
// (A quality line may begin with '@' too, but then the line two further on is a sequence, which never
// begins with '+' in the four-line format the device goes on to verify record by record.)  -1: none in the window.
// `text` = the n bytes of the file from offset `from` = at - 1 on
int64_t record_boundary_in(const char* text, uint64_t n, uint64_t from) {
    for (uint64_t i = 1; i < n; ++i) {
        if (text[i - 1] != '\n' || text[i] != '@') continue;
        const char* e1 = (const char*)memchr(text + i, '\n', (size_t)(n - i));
        if (!e1) return -1;
        const uint64_t l2 = (uint64_t)(e1 - text) + 1;
        const char* e2 = l2 < n ? (const char*)memchr(text + l2, '\n', (size_t)(n - l2)) : nullptr;
        if (!e2) return -1;
        const uint64_t l3 = (uint64_t)(e2 - text) + 1;
        if (l3 >= n) return -1;
        if (text[l3] == '+') return (int64_t)(from + i);
    }
    return -1;
}
// map (optional): the file, memory-mapped -- the search then touches the few hundred bytes it needs instead of
// reading the whole window
int64_t record_boundary(int fd, const char* map, uint64_t at, uint64_t size, uint64_t window, std::vector<char>& buf) {
    if (at == 0) return 0;
    if (at >= size) return (int64_t)size;
    const uint64_t from = at - 1, n = std::min<uint64_t>(window + 1, size - from);
    if (map) return record_boundary_in(map + from, n, from);
    buf.resize((size_t)n);
    uint64_t got = 0;
    while (got < n) {
        ssize_t r = pread(fd, buf.data() + got, (size_t)(n - got), (off_t)(from + got));
        if (r <= 0) return -1;
        got += (uint64_t)r;
    }
    return record_boundary_in(buf.data(), n, from);
}

enum Road { kKseq, kDevice, kStrip, kHybrid };
Road fastq_road() {
    const char* e = getenv("VG_RAW_FASTQ");
    if (e && atoi(e) == 0) return kKseq;
    const char* r = getenv("VG_FASTQ_ROAD");
    if (r && !strcmp(r, "kseq")) return kKseq;
    if (r && !strcmp(r, "device")) return kDevice;
    if (r && !strcmp(r, "strip")) return kStrip;
    return kHybrid;
}
// share of a file's blocks the host workers strip (the rest is parsed on the device)
double strip_share(bool multi) {
    const Road r = fastq_road();
    if (multi || r == kStrip) return 1.0;  // several GPUs: host only (the device road's per-file state lives on ONE device)
    if (r == kDevice) return 0.0;
    const char* e = getenv("VG_STRIP_SHARE");
    const double v = e ? atof(e) : 0.65;
    return v < 0 ? 0 : (v > 1 ? 1 : v);
}
bool raw_enabled() { return fastq_road() != kKseq; }

// Cut a mapped text into record-aligned blocks of at most a staging chunk each, `share` of them for the strip road.
void cut_blocks(RawFile& f, vg_ctx* ctx, double share, int file_index, std::vector<RawItem>& items) {
    const uint64_t window = std::min<uint64_t>(1u << 20, ctx->chunk_bytes / 4);
    const uint64_t step = ctx->chunk_bytes - window - 64;
    std::vector<char> buf;
    f.cut.push_back(0);
    while (f.cut.back() < f.size) {
        const uint64_t target = f.cut.back() + step;
        if (target >= f.size) {
            f.cut.push_back(f.size);
            break;
        }
        const int64_t b = record_boundary(f.fd, f.map, target, f.size, window, buf);
        if (b < 0) {  // records longer than the window, or not four-line FASTQ: the host parser takes over here
            f.tail_from = f.cut.back();
            break;
        }
        f.cut.push_back((uint64_t)b);
    }
    for (size_t b = 0; b + 1 < f.cut.size(); ++b) {
        const bool strip = (uint64_t)((double)(b + 1) * share) > (uint64_t)((double)b * share);  // evenly spread over the file
        items.push_back({file_index, (uint32_t)b, f.cut[b], f.cut[b + 1], f.cut[b + 1] == f.size && f.ends_at_eof, strip});
    }
}

// gzip input takes the parallel inflater (vg_gzip.cpp) unless VG_GZ_PARALLEL=0 (then, as in round 1: zlib on one thread per file)
bool gz_parallel_enabled() {
    const char* e = getenv("VG_GZ_PARALLEL");
    return !(e && atoi(e) == 0) && raw_enabled();
}

// Route one path: plain text starting with '@' -> raw blocks (as far as record boundaries can be found),
// anything else (gzip, FASTA, leading junk) -> the kseq reader.  false: cannot open.
bool plan_file(const char* path, vg_ctx* ctx, bool multi, std::vector<RawFile>& raws, std::vector<RawItem>& items,
               std::vector<KseqItem>& kseqs, std::vector<std::string>& gzs) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    unsigned char magic[2] = {0, 0};
    struct stat st;
    const bool regular = raw_enabled() && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 &&
                         pread(fd, magic, 2, 0) == 2 && (uint64_t)st.st_size < (1ull << 62) &&
                         ctx->chunk_bytes >= (64u << 10) && ctx->chunk_bytes < (1ull << 32);
    if (regular && magic[0] == 0x1f && magic[1] == 0x8b && gz_parallel_enabled()) {
        close(fd);
        gzs.push_back(path);
        return true;
    }
    const bool plain = regular && magic[0] == '@';
    if (!plain) {
        close(fd);
        kseqs.push_back({path, 0});
        return true;
    }
    RawFile f;
    f.path = path;
    f.fd = fd;
    f.size = (uint64_t)st.st_size;
    {
        void* m = mmap(nullptr, (size_t)f.size, PROT_READ, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) {
            close(fd);
            kseqs.push_back({path, 0});
            return true;
        }
        madvise(m, (size_t)f.size, MADV_SEQUENTIAL);
        f.map = (const char*)m;
    }
    cut_blocks(f, ctx, strip_share(multi), (int)raws.size(), items);
    raws.push_back(std::move(f));
    return true;
}

// One pass of workers over the given work; the calling thread copies and launches.  Raw blocks are submitted
// strictly in item order (the per-file "bad from here on" flag relies on stream order).
int run_feeder(const Sinks& sinks, const std::vector<KseqItem>& kseqs, std::vector<RawFile>& raws,
               const std::vector<RawItem>& items, vg::FastqFileState* d_files, int threads, uint64_t* read_bases) {
    const int nsink = (int)sinks.ix.size();
    bool any_strip = false, any_raw = false;
    for (const auto& it : items) (it.strip ? any_strip : any_raw) = true;
    int nk = std::min<int>(threads, (int)kseqs.size());
    int nr = items.empty() ? 0 : std::max(1, std::min(std::min(threads - nk, 64), (int)items.size()));
    const int nworkers = nk + nr;
    if (nworkers == 0) return VG_OK;
    // a slot stays taken from the moment a worker starts filling it until the kernel that read its device twin is done, and
    // blocks are submitted in file order: two slots per worker keep the workers busy while finished chunks wait their turn
    const int nslots = std::max(3, (2 * nworkers + 4 + nsink - 1) / nsink);
    // ring slots (pinned + device pairs); allocation is done here on the calling thread
    for (vg_index* ix : sinks.ix) {
        vg_ctx* ctx = ix->ctx;
        vg::DeviceGuard g(ctx->device);
        while ((int)ctx->ring.size() < nslots) {
            vg::StageSlot s;
            cudaError_t e = cudaMalloc((void**)&s.d_buf, ctx->chunk_bytes + 256);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
            if (e != cudaSuccess) return vg::fail(VG_E_NOMEM, "staging ring: %s", cudaGetErrorString(e));
            ctx->ring.push_back(s);
        }
        for (auto& s : ctx->ring) {  // the feeder's workers fill the pinned twins directly
            if (!s.h_pin && cudaHostAlloc((void**)&s.h_pin, ctx->chunk_bytes + 256, cudaHostAllocDefault) != cudaSuccess)
                return vg::fail(VG_E_NOMEM, "pinned staging buffer: %s", cudaGetErrorString(cudaGetLastError()));
        }
        for (auto& sl : ctx->ring) {
            if (sl.busy) {
                cudaEventSynchronize(sl.done);
                sl.busy = false;
            }
        }
        if (any_raw) {
            int rc = vg::ctx_ensure_fastq(ctx);
            if (rc) return rc;
        }
    }
    Feeder fd;
    const auto t_start = std::chrono::steady_clock::now();
    {   // free slots, the GPUs' interleaved: consecutive chunks go to different GPUs
        size_t most = 0;
        for (vg_index* ix : sinks.ix) most = std::max(most, ix->ctx->ring.size());
        for (size_t i = 0; i < most; ++i)
            for (int c = 0; c < nsink; ++c)
                if (i < sinks.ix[(size_t)c]->ctx->ring.size()) fd.free_q.push_back(c * kSlotsPerSink + (int)i);
    }
    fd.workers_left = nworkers;
    std::vector<std::thread> pool;
    for (int i = 0; i < nk; ++i) pool.emplace_back(kseq_worker, &fd, &sinks, &kseqs);
    for (int i = 0; i < nr; ++i) pool.emplace_back(block_worker, &fd, &sinks, &raws, &items);
    uint64_t strip_bases = 0;
    auto enqueue = [&](int id, uint64_t len, const unsigned int* d_skip = nullptr) {  // a staged chunk of "sequence\n" records -> its GPU
        vg::DeviceGuard g(sinks.ctx(id)->device);
        return vg::enqueue_piece(sinks.index(id), id % kSlotsPerSink, (const char*)sinks.slot(id).h_pin, len, d_skip);
    };

    std::deque<int> inflight;
    std::map<int64_t, Filled> raw_ready;  // raw blocks that arrived ahead of their turn
    int64_t next_raw_submit = 0;
    int rc = VG_OK;
    for (;;) {
        for (auto it = inflight.begin(); it != inflight.end();) {  // per GPU in order; the GPUs finish independently
            if (cudaEventQuery(sinks.slot(*it).done) == cudaSuccess) {
                sinks.slot(*it).busy = false;
                fd.give_free(*it);
                it = inflight.erase(it);
            } else {
                ++it;
            }
        }
        cudaGetLastError();
        Filled f{-1, 0, -1};
        bool finished = false;
        {
            std::unique_lock<std::mutex> lk(fd.mu);
            if (fd.ready_q.empty() && fd.workers_left > 0 && !fd.abort) {
                if (inflight.empty()) fd.cv_ready.wait(lk, [&] { return !fd.ready_q.empty() || fd.workers_left == 0 || fd.abort; });
                else fd.cv_ready.wait_for(lk, std::chrono::microseconds(200));
            }
            if (!fd.ready_q.empty()) {
                f = fd.ready_q.front();
                fd.ready_q.pop_front();
            } else if (fd.workers_left == 0 || fd.abort) {
                finished = true;
            }
        }
        if (f.slot >= 0 && f.raw_item < 0) {
            rc = enqueue(f.slot, f.len);
            if (rc == VG_OK) inflight.push_back(f.slot);
        } else if (f.slot >= 0) {
            raw_ready[f.raw_item] = f;
            for (auto it = raw_ready.find(next_raw_submit); rc == VG_OK && it != raw_ready.end();
                 it = raw_ready.find(next_raw_submit)) {
                const RawItem& item = items[(size_t)next_raw_submit];
                const Filled g = it->second;
                raw_ready.erase(it);
                ++next_raw_submit;
                // in file order: nothing behind the first irregular record a HOST worker found is ever submitted; what a
                // DEVICE check refuses raises the file's flag on the device, and everything submitted after it -- raw block
                // or stripped chunk -- is skipped there (stream order), bases included
                RawFile& rf = raws[(size_t)item.file];
                const bool skip = rf.bad_from != ~0ull;
                if (item.strip) {
                    if (!skip && g.bad_at != ~0ull) rf.bad_from = g.bad_at;
                    if (!skip && !d_files) {  // host-only bookkeeping (no device road in this call)
                        strip_bases += g.bases;
                        if (g.bad_at == ~0ull) sinks.index(g.slot)->fastq_blocks += 1;
                    }
                    if (skip || (g.len == 0 && g.bases == 0)) {
                        fd.give_free(g.slot);
                        continue;
                    }
                    vg::FastqFileState* df = d_files ? d_files + item.file : nullptr;
                    if (g.len) rc = enqueue(g.slot, g.len, df ? &df->bad : nullptr);
                    if (rc == VG_OK && df) rc = vg::commit_stripped(sinks.ix[0], df, g.bases, g.bad_at == ~0ull);
                    if (rc == VG_OK && g.len) inflight.push_back(g.slot);
                    else fd.give_free(g.slot);
                    continue;
                }
                if (skip || g.len == 0) {
                    fd.give_free(g.slot);
                    continue;
                }
                rc = vg::enqueue_raw_piece(sinks.ix[0], g.slot, g.len, d_files + item.file, item.block);  // device road: one GPU
                if (rc == VG_OK) inflight.push_back(g.slot);
            }
        }
        if (rc != VG_OK) {
            fd.set_error(rc, vg_last_error());
            finished = true;
        }
        if (finished) break;
    }
    for (auto& t : pool) t.join();
    if (getenv("VG_FEEDER_DEBUG"))
        fprintf(stderr, "[vg_feeder] %d kseq + %d block workers (%s road), %zu blocks, workers busy %.1f ms / waiting for slots %.1f ms in total, wall %.1f ms\n", nk, nr,
                any_strip && any_raw ? "strip + device" : (any_strip ? "strip" : "device"), items.size(), fd.busy_ns.load() * 1e-6, fd.wait_ns.load() * 1e-6,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    if (read_bases) *read_bases += fd.read_bases.load() + strip_bases;
    if (fd.err != VG_OK) return vg::fail(fd.err, "%s", fd.err_msg.c_str());
    return VG_OK;
}

}  // namespace

// gzip files: windows of inflated text (vg_gzip.cpp, all workers), each cut at its last record boundary and run through
// the strip road like a mapped plain file; the cut-off tail is carried into the next window.  A helper thread inflates
// the next window (of this file or the next) while the calling thread counts the current one.  The moment anything is
// off -- the text is not four-line FASTQ, a record fails the strip road's check, the inflater reports an error -- the
// rest of that file, from the uncompressed offset reached, goes to the kseq reader over zlib (`kseqs`), which is the
// reference's own road (src/fastq_kmer.cpp:74-141): same reads counted for any input.
namespace {
struct GzWindow {
    vg::gz::Buffer* text = nullptr;  // [0, use) is this window; null: the producer is done
    uint64_t use = 0;
    uint64_t done = 0;               // uncompressed offset of text->data[0]
    bool eof = false;                // the file ends with this window
    bool good = true;                // false: nothing of this window is to be counted; zlib takes over at `done`
    int file = -1;
};
struct GzPipe {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<GzWindow> ready;
    std::deque<vg::gz::Buffer*> free_bufs;
    std::vector<char> stopped;       // per file: the consumer gave up on the fast road
    bool abort = false;
    double inflate_s = 0;
    vg::gz::Buffer* take() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return abort || !free_bufs.empty(); });
        if (abort) return nullptr;
        vg::gz::Buffer* b = free_bufs.front();
        free_bufs.pop_front();
        return b;
    }
    void give(vg::gz::Buffer* b) {
        {
            std::lock_guard<std::mutex> lk(mu);
            free_bufs.push_back(b);
        }
        cv.notify_all();
    }
    void push(const GzWindow& w) {
        {
            std::lock_guard<std::mutex> lk(mu);
            ready.push_back(w);
        }
        cv.notify_all();
    }
    bool is_stopped(int f) {
        std::lock_guard<std::mutex> lk(mu);
        return abort || stopped[(size_t)f] != 0;
    }
};
// working memory kept from call to call (per calling thread): the inflated windows and the workers' symbol buffers --
// mapping and faulting in ~1 GB afresh per file costs as much as inflating it
struct GzArena {
    vg::gz::Scratch scratch;
    vg::gz::Buffer bufs[2];
};

void gz_produce(GzPipe* pipe, GzArena* arena, const std::vector<std::string>* paths, int threads, uint64_t chunk, uint64_t window_bytes,
                uint64_t boundary_back, bool debug);
void gz_producer(GzPipe* pipe, GzArena* arena, const std::vector<std::string>* paths, int threads, uint64_t chunk, uint64_t window_bytes,
                 uint64_t boundary_back, bool debug) {
    gz_produce(pipe, arena, paths, threads, chunk, window_bytes, boundary_back, debug);
    pipe->push(GzWindow{});  // whatever happened: the consumer waits for this
}
void gz_produce(GzPipe* pipe, GzArena* arena, const std::vector<std::string>* paths, int threads, uint64_t chunk, uint64_t window_bytes,
                uint64_t boundary_back, bool debug) {
    // chunks per worker and round: several, so that the workers' loads even out (a chunk of text-like input takes 2-3x as
    // long as one of low-entropy qualities) and a file's last round is not a short one
    const char* pe = getenv("VG_GZ_PER_THREAD");
    const int per_thread = pe && atoi(pe) > 0 ? atoi(pe) : 4;
    for (int fi = 0; fi < (int)paths->size(); ++fi) {
        const std::string& path = (*paths)[(size_t)fi];
        int fd = open(path.c_str(), O_RDONLY);
        struct stat st;
        void* m = MAP_FAILED;
        if (fd >= 0 && fstat(fd, &st) == 0) m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
        if (fd >= 0) close(fd);
        GzWindow w;
        w.file = fi;
        if (m == MAP_FAILED) {  // the kseq road reports what is wrong with it
            w.good = false;
            w.text = nullptr;
            vg::gz::Buffer* b = pipe->take();
            if (!b) return;
            b->size = 0;
            w.text = b;
            pipe->push(w);
            continue;
        }
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        {
            vg::gz::Stream stream((const uint8_t*)m, (uint64_t)st.st_size, threads, chunk, &arena->scratch);
            vg::gz::Buffer* cur = pipe->take();
            if (!cur) {
                munmap(m, (size_t)st.st_size);
                return;
            }
            cur->size = 0;
            uint64_t done = 0;
            bool first = true;
            for (;;) {
                if (pipe->is_stopped(fi)) {
                    pipe->give(cur);
                    break;
                }
                const auto t0 = std::chrono::steady_clock::now();
                bool good = true;
                while (good && !stream.eof() && cur->size < window_bytes) good = stream.next(*cur, per_thread);
                pipe->inflate_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                const bool eof = good && stream.eof();
                if (first && cur->size && cur->data[0] != '@') good = false;  // FASTA, or not sequence data at all
                first = false;
                uint64_t use = cur->size;
                if (good && !eof) {  // cut at the first record boundary inside the last stretch of the window
                    const uint64_t back = std::min<uint64_t>(cur->size > 1 ? cur->size - 1 : 0, boundary_back);
                    const uint64_t from = cur->size - back;
                    const int64_t b = back ? record_boundary_in((const char*)cur->data + from - 1, back + 1, from - 1) : -1;
                    if (b <= 0) good = false;
                    else use = (uint64_t)b;
                }
                if (!good && debug && !stream.error().empty())
                    fprintf(stderr, "[vg_feeder] %s: %s; the rest goes through zlib\n", path.c_str(), stream.error().c_str());
                vg::gz::Buffer* nxt = nullptr;
                if (good && !eof) {  // the tail moves to the front of the next window
                    nxt = pipe->take();
                    if (!nxt || !nxt->reserve(cur->size - use + 1)) {
                        if (nxt) pipe->give(nxt);
                        nxt = nullptr;
                        good = false;
                    } else {
                        memcpy(nxt->data, cur->data + use, (size_t)(cur->size - use));
                        nxt->size = cur->size - use;
                    }
                }
                w.text = cur;
                w.use = use;
                w.done = done;
                w.eof = eof;
                w.good = good;
                pipe->push(w);
                if (!good || eof) break;
                done += use;
                cur = nxt;
            }
        }
        munmap(m, (size_t)st.st_size);
    }
}
}  // namespace

static int count_gz_files(const Sinks& sinks, const std::vector<std::string>& paths, int threads, uint64_t* read_bases,
                          std::vector<KseqItem>& kseqs) {
    vg_ctx* ctx = sinks.ix[0]->ctx;
    const char* e = getenv("VG_GZ_CHUNK");
    const uint64_t chunk = e ? strtoull(e, nullptr, 10) : (1ull << 20);
    const char* wmb = getenv("VG_GZ_WINDOW_MB");
    const uint64_t window_bytes = (wmb ? strtoull(wmb, nullptr, 10) : 512ull) << 20;
    const bool debug = getenv("VG_FEEDER_DEBUG") != nullptr;
    static thread_local GzArena arena;
    GzPipe pipe;
    pipe.stopped.assign(paths.size(), 0);
    for (auto& b : arena.bufs) pipe.free_bufs.push_back(&b);
    std::thread producer(gz_producer, &pipe, &arena, &paths, threads, chunk, window_bytes,
                         std::min<uint64_t>(1u << 20, ctx->chunk_bytes / 4), debug);
    double t_count = 0;
    int rc = VG_OK;
    for (;;) {
        GzWindow w;
        {
            std::unique_lock<std::mutex> lk(pipe.mu);
            pipe.cv.wait(lk, [&] { return !pipe.ready.empty(); });
            w = pipe.ready.front();
            pipe.ready.pop_front();
        }
        if (!w.text) break;
        const std::string& path = paths[(size_t)w.file];
        char& stopped = pipe.stopped[(size_t)w.file];
        if (rc == VG_OK && !stopped) {
            if (!w.good) {  // text->data[0] on (uncompressed offset `done`) has not been counted
                kseqs.push_back({path, w.done});
                std::lock_guard<std::mutex> lk(pipe.mu);
                stopped = 1;
            } else if (w.use) {
                const auto t1 = std::chrono::steady_clock::now();
                std::vector<RawFile> raws(1);
                std::vector<RawItem> items;
                RawFile& f = raws[0];
                f.path = path;
                f.size = w.use;
                f.map = (const char*)w.text->data;
                f.borrowed = true;
                f.ends_at_eof = w.eof;
                cut_blocks(f, ctx, 1.0, 0, items);
                rc = run_feeder(sinks, {}, raws, items, nullptr, threads, read_bases);
                t_count += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
                const uint64_t from = std::min(f.tail_from, f.bad_from);
                if (rc == VG_OK && from != ~0ull && from < w.use) {  // an irregular record: exact semantics from there on
                    kseqs.push_back({path, w.done + from});
                    std::lock_guard<std::mutex> lk(pipe.mu);
                    stopped = 1;
                }
                if (rc != VG_OK) {
                    std::lock_guard<std::mutex> lk(pipe.mu);
                    pipe.abort = true;
                }
            }
        }
        pipe.give(w.text);
    }
    producer.join();
    if (debug) fprintf(stderr, "[vg_feeder] %zu gzip file(s): inflate %.1f ms (helper thread), count %.1f ms\n", paths.size(), pipe.inflate_s * 1e3, t_count * 1e3);
    return rc;
}

// ixs: one index per GPU, all of them counting; the chunks of the files are dealt to them as they come.
static int count_files_multi(const std::vector<vg_index*>& ixs, const char* const* paths, int npaths, int threads, uint64_t* read_bases) {
    Sinks sinks{ixs};
    const auto t_plan = std::chrono::steady_clock::now();
    vg_index* ix = ixs[0];
    vg_ctx* ctx = ix->ctx;
    const bool multi = ixs.size() > 1;
    if (threads < 1) threads = 1;
    std::vector<KseqItem> kseqs;
    std::vector<RawFile> raws;
    std::vector<RawItem> items;
    auto close_all = [&] {
        for (auto& f : raws) {
            if (f.map && !f.borrowed) munmap((void*)f.map, (size_t)f.size);
            if (f.fd >= 0) close(f.fd);
            f.map = nullptr;
            f.fd = -1;
        }
    };
    std::vector<std::string> gzs;
    for (int i = 0; i < npaths; ++i) {
        if (!plan_file(paths[i], ctx, multi, raws, items, kseqs, gzs)) {
            close_all();
            return vg::fail(VG_E_IO, "'%s': No such file or directory.", paths[i]);
        }
    }
    // gzip files first, one after the other, each inflated by all the workers; whatever the fast road cannot take of
    // them joins the kseq list below
    if (!gzs.empty()) {
        int rc = count_gz_files(sinks, gzs, threads, read_bases, kseqs);
        if (rc != VG_OK) {
            close_all();
            return rc;
        }
    }
    bool any_raw = false;
    for (const auto& it : items) any_raw |= !it.strip;
    if (getenv("VG_FEEDER_DEBUG"))
        fprintf(stderr, "[vg_feeder] %d file(s) planned in %.1f ms\n", npaths,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan).count());
    vg::FastqFileState* d_files = nullptr;
    if (any_raw) {  // per-file state on the device: the raw blocks' verdicts, and the bases of what was counted
        cudaError_t e = cudaMalloc((void**)&d_files, raws.size() * sizeof(vg::FastqFileState));
        if (e == cudaSuccess) e = cudaMemsetAsync(d_files, 0, raws.size() * sizeof(vg::FastqFileState), ctx->compute_stream);
        if (e != cudaSuccess) {
            close_all();
            cudaFree(d_files);
            return vg::fail(VG_E_NOMEM, "FASTQ file states: %s", cudaGetErrorString(e));
        }
    }
    int rc = run_feeder(sinks, kseqs, raws, items, d_files, threads, read_bases);
    close_all();
    // What was refused (from the first record / block that is not plain four-line FASTQ on) and what could
    // not be cut into blocks goes through the kseq reader now.
    std::vector<KseqItem> again;
    if (rc == VG_OK && !d_files) {
        for (auto& f : raws) {
            const uint64_t from = std::min(f.tail_from, f.bad_from);
            if (from != ~0ull && from < f.size) again.push_back({f.path, from});
        }
    } else if (rc == VG_OK) {
        std::vector<vg::FastqFileState> st(raws.size());
        cudaError_t e = cudaStreamSynchronize(ctx->compute_stream);
        if (e == cudaSuccess) e = cudaMemcpy(st.data(), d_files, st.size() * sizeof(vg::FastqFileState), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = vg::fail(VG_E_CUDA, "FASTQ file states: %s", cudaGetErrorString(e));
        for (size_t i = 0; rc == VG_OK && i < raws.size(); ++i) {
            if (read_bases) *read_bases += st[i].read_bases;
            ix->fastq_blocks += st[i].blocks_ok;
            uint64_t from = std::min(raws[i].tail_from, raws[i].bad_from);  // what the host found ...
            if (st[i].bad) from = std::min(from, raws[i].cut[st[i].first_bad_block]);  // ... and what the device refused
            if (from != ~0ull && from < raws[i].size) again.push_back({raws[i].path, from});
        }
    }
    cudaFree(d_files);
    std::vector<RawFile> none;
    if (rc == VG_OK && !again.empty()) rc = run_feeder(sinks, again, none, {}, nullptr, threads, read_bases);
    return rc;
}

int vg::count_files(vg_index* ix, const char* const* paths, int npaths, int threads, uint64_t* read_bases) {
    return count_files_multi({ix}, paths, npaths, threads, read_bases);
}

extern "C" uint64_t vg_index_fastq_blocks(const vg_index* ix) { return ix ? ix->fastq_blocks : 0; }

// Host-only test hook (needs no GPU): where the raw road would cut `path` at or after byte `at`.
extern "C" int64_t vg_fastq_record_boundary(const char* path, uint64_t at, uint64_t window) {
    if (!path) return -2;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -2;
    struct stat st;
    int64_t r = -2;
    if (fstat(fd, &st) == 0) {
        std::vector<char> buf;
        r = record_boundary(fd, nullptr, at, (uint64_t)st.st_size, window, buf);
    }
    close(fd);
    return r;
}

// Host-only test hook (needs no GPU): the strip road's scanner over one block of text that starts at a record boundary.
extern "C" int64_t vg_fastq_strip_block(const char* text, uint64_t nbytes, int last, uint8_t* out, uint64_t* bases, int64_t* bad_at) {
    if ((!text && nbytes) || !out) return -1;
    uint64_t b = 0;
    const char* bad = nullptr;
    const uint64_t w = strip_block(text, text + nbytes, last != 0, out, b, bad);
    if (bases) *bases = b;
    if (bad_at) *bad_at = bad ? (int64_t)(bad - text) : -1;
    return (int64_t)w;
}

extern "C" int vg_count_files_multi(vg_index* const* ixs, int nix, const char* const* paths, int npaths, int threads,
                                    uint64_t* read_bases) {
    if (!ixs || nix <= 0 || !paths || npaths <= 0) return vg::fail(VG_E_INVALID, "Parameter error: -f");
    std::vector<vg_index*> v;
    for (int i = 0; i < nix; ++i) {
        if (!ixs[i]) return vg::fail(VG_E_INVALID, "vg_count_files_multi: ixs[%d] is NULL", i);
        if (!ixs[i]->counting) return vg::fail(VG_E_STATE, "vg_count_files_multi before vg_count_begin (index %d)", i);
        if (ixs[i]->sharded) return vg::fail(VG_E_STATE, "vg_count_files_multi is for replicas of an index, one per GPU");
        if (ixs[i]->ctx->chunk_bytes != ixs[0]->ctx->chunk_bytes) return vg::fail(VG_E_INVALID, "vg_count_files_multi: contexts with different --buffer");
        for (int j = 0; j < i; ++j)
            if (ixs[j]->ctx == ixs[i]->ctx) return vg::fail(VG_E_INVALID, "vg_count_files_multi: one context listed twice");
        v.push_back(ixs[i]);
    }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(v[0]->ctx->device);
    int rc = count_files_multi(v, paths, npaths, threads, read_bases);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

extern "C" int vg_count_files(vg_index* ix, const char* const* paths, int npaths, int threads, uint64_t* read_bases) {
    if (!ix || !paths || npaths <= 0) return vg::fail(VG_E_INVALID, "Parameter error: -f");
    if (!ix->counting) return vg::fail(VG_E_STATE, "vg_count_files before vg_count_begin");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ix->ctx->device);
    int rc = vg::count_files(ix, paths, npaths, threads, read_bases);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}
