// vg_feeder.cpp -- FASTQ/FASTA (plain or gzip) -> pinned staging ring -> count kernels.
//
// Stands in for FastqKmer::fastq_file_open (src/fastq_kmer.cpp:65-187): the reference inflates and
// parses on its main thread (kseq over gzread, 16 KiB buffer), copies every read into a
// std::string, upper-cases it and batches `threads*100` reads per pool task.  Here each file
// gets its own inflate/parse worker that writes "sequence\n" records straight into a pinned
// chunk; the calling thread only issues cudaMemcpyAsync + kernel launches, so copy, kernel and
// parsing overlap and R1/R2 are read concurrently.
//
// Plain (not gzip) four-line FASTQ takes a shorter road: the workers only pread() raw text
// into the pinned chunks, cut at record boundaries; the GPU finds the lines, checks every record against
// kseq's rules and blanks everything but the sequences (fastq_*_kernel in vg_kernels.cu).  Whatever fails
// that check -- multi-line records, FASTA, a truncated tail, NUL bytes -- is left uncounted on the device
// from the offending block on and re-read here with the kseq reader, so the result is the reference's
// for any input.  VG_RAW_FASTQ=0 sends every file through the kseq reader.
//
// What counts as a read follows kseq (include/kseq.h:192-232) exactly: header at '@' or '>',
// sequence = every line up to one starting with '+', '>' or '@', one trailing CR stripped per
// line, quality must match the sequence length or the file stops there (return -2) without
// counting that record; mReadBase sums seq.l (src/fastq_kmer.cpp:105).  The reference builds a
// std::string from the C string, so a sequence is cut at its first NUL byte.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vgb200.h"
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

struct Filled {
    int slot;
    uint64_t len;
    int64_t raw_item;  // >= 0: raw FASTQ text, item number in submission order; -1: "sequence\n" records
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
};

struct RawItem {
    int file;
    uint32_t block;
    uint64_t start, end;
    bool last;  // ends at EOF: make sure the text ends with a newline, as kseq treats EOF
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

    int take_free() {
        std::unique_lock<std::mutex> lk(mu);
        cv_free.wait(lk, [&] { return abort || !free_q.empty(); });
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

void kseq_worker(Feeder* fd, vg_ctx* ctx, const std::vector<KseqItem>* items) {
    std::string seq;
    int slot = -1;
    uint64_t w = 0, bases = 0;
    const uint64_t cap = ctx->chunk_bytes;
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
            uint8_t* dst = ctx->ring[(size_t)slot].h_pin + w;
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

// A slot first, then the next block: the lowest outstanding block always has a buffer, so the in-order
// submission of the calling thread cannot starve.
void raw_worker(Feeder* fd, vg_ctx* ctx, const std::vector<RawFile>* files, const std::vector<RawItem>* items) {
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
        uint8_t* dst = ctx->ring[(size_t)slot].h_pin;
        uint64_t len = it.end - it.start, got = 0;
        while (got < len) {
            ssize_t r = pread(f.fd, dst + got, (size_t)(len - got), (off_t)(it.start + got));
            if (r <= 0) break;
            got += (uint64_t)r;
        }
        if (got != len) {
            fd->set_error(VG_E_IO, "'" + f.path + "': read error");
            break;
        }
        if (it.last && len && dst[len - 1] != '\n') dst[len++] = '\n';
        fd->push_ready(slot, len, (int64_t)idx);
    }
    fd->worker_done();
}

// First record start at or after `at`: a line that begins with '@' whose next-but-one line begins with '+'.
// (A quality line may begin with '@' too, but then the line two further on is a sequence, which never
// begins with '+' in the four-line format the device goes on to verify record by record.)  -1: none in the window.
int64_t record_boundary(int fd, uint64_t at, uint64_t size, uint64_t window, std::vector<char>& buf) {
    if (at == 0) return 0;
    if (at >= size) return (int64_t)size;
    const uint64_t from = at - 1, n = std::min<uint64_t>(window + 1, size - from);
    buf.resize((size_t)n);
    uint64_t got = 0;
    while (got < n) {
        ssize_t r = pread(fd, buf.data() + got, (size_t)(n - got), (off_t)(from + got));
        if (r <= 0) return -1;
        got += (uint64_t)r;
    }
    for (uint64_t i = 1; i < n; ++i) {
        if (buf[i - 1] != '\n' || buf[i] != '@') continue;
        const char* e1 = (const char*)memchr(buf.data() + i, '\n', (size_t)(n - i));
        if (!e1) return -1;
        const uint64_t l2 = (uint64_t)(e1 - buf.data()) + 1;
        const char* e2 = l2 < n ? (const char*)memchr(buf.data() + l2, '\n', (size_t)(n - l2)) : nullptr;
        if (!e2) return -1;
        const uint64_t l3 = (uint64_t)(e2 - buf.data()) + 1;
        if (l3 >= n) return -1;
        if (buf[l3] == '+') return (int64_t)(from + i);
    }
    return -1;
}

bool raw_enabled() {
    const char* e = getenv("VG_RAW_FASTQ");
    return !(e && atoi(e) == 0);
}

// Route one path: plain text starting with '@' -> raw blocks (as far as record boundaries can be found),
// anything else (gzip, FASTA, leading junk) -> the kseq reader.  false: cannot open.
bool plan_file(const char* path, vg_ctx* ctx, std::vector<RawFile>& raws, std::vector<RawItem>& items,
               std::vector<KseqItem>& kseqs) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    unsigned char magic[2] = {0, 0};
    struct stat st;
    const bool plain = raw_enabled() && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 &&
                       pread(fd, magic, 2, 0) >= 1 && magic[0] == '@' && (uint64_t)st.st_size < (1ull << 62) &&
                       ctx->chunk_bytes >= (64u << 10) && ctx->chunk_bytes < (1ull << 32);
    if (!plain) {
        close(fd);
        kseqs.push_back({path, 0});
        return true;
    }
    RawFile f;
    f.path = path;
    f.fd = fd;
    f.size = (uint64_t)st.st_size;
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
        const int64_t b = record_boundary(fd, target, f.size, window, buf);
        if (b < 0) {  // records longer than the window, or not four-line FASTQ: the host parser takes over here
            f.tail_from = f.cut.back();
            break;
        }
        f.cut.push_back((uint64_t)b);
    }
    const int fi = (int)raws.size();
    for (size_t b = 0; b + 1 < f.cut.size(); ++b)
        items.push_back({fi, (uint32_t)b, f.cut[b], f.cut[b + 1], f.cut[b + 1] == f.size});
    raws.push_back(std::move(f));
    return true;
}

// One pass of workers over the given work; the calling thread copies and launches.  Raw blocks are submitted
// strictly in item order (the per-file "bad from here on" flag relies on stream order).
int run_feeder(vg_index* ix, const std::vector<KseqItem>& kseqs, const std::vector<RawFile>& raws,
               const std::vector<RawItem>& items, vg::FastqFileState* d_files, int threads, uint64_t* read_bases) {
    vg_ctx* ctx = ix->ctx;
    int nk = std::min<int>(threads, (int)kseqs.size());
    int nr = items.empty() ? 0 : std::max(1, std::min(std::min(threads - nk, 16), (int)items.size()));
    const int nworkers = nk + nr;
    if (nworkers == 0) return VG_OK;
    const int nslots = std::max(3, nworkers + 2);
    // ring slots (pinned + device pairs); allocation is done here on the calling thread
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
    if (nr) {
        int rc = vg::ctx_ensure_fastq(ctx);
        if (rc) return rc;
    }
    Feeder fd;
    for (int i = 0; i < (int)ctx->ring.size(); ++i) fd.free_q.push_back(i);
    fd.workers_left = nworkers;
    std::vector<std::thread> pool;
    for (int i = 0; i < nk; ++i) pool.emplace_back(kseq_worker, &fd, ctx, &kseqs);
    for (int i = 0; i < nr; ++i) pool.emplace_back(raw_worker, &fd, ctx, &raws, &items);

    std::deque<int> inflight;
    std::map<int64_t, Filled> raw_ready;  // raw blocks that arrived ahead of their turn
    int64_t next_raw_submit = 0;
    int rc = VG_OK;
    for (;;) {
        while (!inflight.empty() && cudaEventQuery(ctx->ring[(size_t)inflight.front()].done) == cudaSuccess) {
            ctx->ring[(size_t)inflight.front()].busy = false;
            fd.give_free(inflight.front());
            inflight.pop_front();
        }
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
            rc = vg::enqueue_piece(ix, f.slot, (const char*)ctx->ring[(size_t)f.slot].h_pin, f.len);
            if (rc == VG_OK) inflight.push_back(f.slot);
        } else if (f.slot >= 0) {
            raw_ready[f.raw_item] = f;
            for (auto it = raw_ready.find(next_raw_submit); rc == VG_OK && it != raw_ready.end();
                 it = raw_ready.find(next_raw_submit)) {
                const RawItem& item = items[(size_t)next_raw_submit];
                const Filled g = it->second;
                raw_ready.erase(it);
                ++next_raw_submit;
                if (g.len == 0) {
                    fd.give_free(g.slot);
                    continue;
                }
                rc = vg::enqueue_raw_piece(ix, g.slot, g.len, d_files + item.file, item.block);
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
    if (read_bases) *read_bases += fd.read_bases.load();
    if (fd.err != VG_OK) return vg::fail(fd.err, "%s", fd.err_msg.c_str());
    return VG_OK;
}

}  // namespace

int vg::count_files(vg_index* ix, const char* const* paths, int npaths, int threads, uint64_t* read_bases) {
    vg_ctx* ctx = ix->ctx;
    if (threads < 1) threads = 1;
    std::vector<KseqItem> kseqs;
    std::vector<RawFile> raws;
    std::vector<RawItem> items;
    auto close_all = [&] {
        for (auto& f : raws)
            if (f.fd >= 0) close(f.fd);
    };
    for (int i = 0; i < npaths; ++i) {
        if (!plan_file(paths[i], ctx, raws, items, kseqs)) {
            close_all();
            return vg::fail(VG_E_IO, "'%s': No such file or directory.", paths[i]);
        }
    }
    vg::FastqFileState* d_files = nullptr;
    if (!raws.empty()) {
        cudaError_t e = cudaMalloc((void**)&d_files, raws.size() * sizeof(vg::FastqFileState));
        if (e == cudaSuccess) e = cudaMemsetAsync(d_files, 0, raws.size() * sizeof(vg::FastqFileState), ctx->compute_stream);
        if (e != cudaSuccess) {
            close_all();
            cudaFree(d_files);
            return vg::fail(VG_E_NOMEM, "FASTQ file states: %s", cudaGetErrorString(e));
        }
    }
    int rc = run_feeder(ix, kseqs, raws, items, d_files, threads, read_bases);
    close_all();
    // What the device refused (from the first block that is not plain four-line FASTQ on) and what could
    // not be cut into blocks goes through the kseq reader now.
    std::vector<KseqItem> again;
    if (rc == VG_OK && !raws.empty()) {
        std::vector<vg::FastqFileState> st(raws.size());
        cudaError_t e = cudaStreamSynchronize(ctx->compute_stream);
        if (e == cudaSuccess) e = cudaMemcpy(st.data(), d_files, st.size() * sizeof(vg::FastqFileState), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = vg::fail(VG_E_CUDA, "FASTQ file states: %s", cudaGetErrorString(e));
        for (size_t i = 0; rc == VG_OK && i < raws.size(); ++i) {
            if (read_bases) *read_bases += st[i].read_bases;
            ix->fastq_blocks += st[i].blocks_ok;
            uint64_t from = raws[i].tail_from;
            if (st[i].bad) from = raws[i].cut[st[i].first_bad_block];
            if (from != ~0ull && from < raws[i].size) again.push_back({raws[i].path, from});
        }
    }
    cudaFree(d_files);
    if (rc == VG_OK && !again.empty()) rc = run_feeder(ix, again, {}, {}, nullptr, threads, read_bases);
    return rc;
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
        r = record_boundary(fd, at, (uint64_t)st.st_size, window, buf);
    }
    close(fd);
    return r;
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
