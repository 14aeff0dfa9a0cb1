// vg_feeder.cpp -- FASTQ/FASTA (plain or gzip) -> pinned staging ring -> count kernels.
//
// Stands in for FastqKmer::fastq_file_open (src/fastq_kmer.cpp:65-187): the reference inflates and
// parses on its main thread (kseq over gzread, 16 KiB buffer), copies every read into a
// std::string, upper-cases it and batches `threads*100` reads per pool task.  Here each file
// gets its own inflate/parse worker that writes "sequence\n" records straight into a pinned
// chunk; the calling thread only issues cudaMemcpyAsync + kernel launches, so copy, kernel and
// parsing overlap and R1/R2 are read concurrently.
//
// What counts as a read follows kseq (include/kseq.h:192-232) exactly: header at '@' or '>',
// sequence = every line up to one starting with '+', '>' or '@', one trailing CR stripped per
// line, quality must match the sequence length or the file stops there (return -2) without
// counting that record; mReadBase sums seq.l (src/fastq_kmer.cpp:105).  The reference builds a
// std::string from the C string, so a sequence is cut at its first NUL byte.
#include <zlib.h>

#include <atomic>
#include <cctype>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
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
    void push_ready(int s, uint64_t len) {
        {
            std::lock_guard<std::mutex> lk(mu);
            ready_q.push_back({s, len});
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
};

void worker(Feeder* fd, vg_ctx* ctx, const char* const* paths, int npaths) {
    std::string seq;
    int slot = -1;
    uint64_t w = 0, bases = 0;
    const uint64_t cap = ctx->chunk_bytes;
    for (;;) {
        int fi = fd->next_file.fetch_add(1);
        if (fi >= npaths) break;
        gzFile gz = gzopen(paths[fi], "rb");
        if (!gz) {
            fd->set_error(VG_E_IO, std::string("'") + paths[fi] + "': No such file or directory.");
            break;
        }
        gzbuffer(gz, 1u << 20);
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
    {
        std::lock_guard<std::mutex> lk(fd->mu);
        fd->workers_left -= 1;
    }
    fd->cv_ready.notify_all();
}

}  // namespace

int vg::count_files(vg_index* ix, const char* const* paths, int npaths, int threads, uint64_t* read_bases) {
    vg_ctx* ctx = ix->ctx;
    int nworkers = threads < 1 ? 1 : threads;
    if (nworkers > npaths) nworkers = npaths;
    const int nslots = nworkers + 2 < 3 ? 3 : nworkers + 2;
    // ring slots (pinned + device pairs); allocation is done here on the calling thread
    while ((int)ctx->ring.size() < nslots) {
        vg::StageSlot s;
        cudaError_t e = cudaMalloc((void**)&s.d_buf, ctx->chunk_bytes + 256);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        if (e != cudaSuccess) return vg::fail(VG_E_NOMEM, "staging ring: %s", cudaGetErrorString(e));
        ctx->ring.push_back(s);
    }
    for (auto& s : ctx->ring) {  // the feeder's workers parse straight into the pinned twins
        if (!s.h_pin && cudaHostAlloc((void**)&s.h_pin, ctx->chunk_bytes + 256, cudaHostAllocDefault) != cudaSuccess)
            return vg::fail(VG_E_NOMEM, "pinned staging buffer: %s", cudaGetErrorString(cudaGetLastError()));
    }
    for (auto& sl : ctx->ring) {
        if (sl.busy) {
            cudaEventSynchronize(sl.done);
            sl.busy = false;
        }
    }
    Feeder fd;
    for (int i = 0; i < (int)ctx->ring.size(); ++i) fd.free_q.push_back(i);
    fd.workers_left = nworkers;
    std::vector<std::thread> pool;
    for (int i = 0; i < nworkers; ++i) pool.emplace_back(worker, &fd, ctx, paths, npaths);

    std::deque<int> inflight;
    int rc = VG_OK;
    for (;;) {
        while (!inflight.empty() && cudaEventQuery(ctx->ring[(size_t)inflight.front()].done) == cudaSuccess) {
            ctx->ring[(size_t)inflight.front()].busy = false;
            fd.give_free(inflight.front());
            inflight.pop_front();
        }
        Filled f{-1, 0};
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
        if (f.slot >= 0) {
            rc = vg::enqueue_piece(ix, f.slot, (const char*)ctx->ring[(size_t)f.slot].h_pin, f.len);
            if (rc != VG_OK) {
                fd.set_error(rc, vg_last_error());
                finished = true;
            } else {
                inflight.push_back(f.slot);
            }
        }
        if (finished) break;
    }
    for (auto& t : pool) t.join();
    if (read_bases) *read_bases += fd.read_bases.load();
    if (fd.err != VG_OK) return vg::fail(fd.err, "%s", fd.err_msg.c_str());
    return VG_OK;
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
