// fastq_kmer_b200.hpp -- drop-in for the reference's include/fastq_kmer.cuh + src/fastq_kmer.cu.
//
// Same class name, constructor and method as the reference GPU build
// (include/fastq_kmer.cuh:10-49): FastqKmerKernel(map&, files, k, threads, buffer) and
// build_fastq_index_kernel(); public result mReadBase (include/fastq_kmer.hpp:42).
// Post-condition (src/fastq_kmer.cpp:132-138): every entry's c == min(255, occurrences of its key
// among the emitted k-mers of all files); f and BitVec untouched; nothing inserted or removed.
//
// Instead of sort + reduce_by_key + host map probes per chunk (src/fastq_kmer.cu:99-162) the map's
// keys live in a device index built once per graph (DeviceGraphIndex below) and the whole count
// phase runs behind the C ABI (include/vgb200.h); only the u8 count vector returns to the host.
//
// --gpu takes a list (main.cu:141-143 takes one id): the index is built on the first GPU and replicated to the others
// over NVLink (vg_index_replicate), and either ONE sample's reads are dealt over all of them and the counts combined
// in place (count_sample), or -- with at least as many samples as GPUs -- every GPU counts a sample of its own while
// the host genotypes the previous ones (count_sample_on; BASELINE config 5).
#pragma once
#include <cstdint>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "fastq_kmer.hpp"  // reference header: FastqKmer, kmerCovFreBitVec
#include "vgb200.h"
#include "vgb200_check.hpp"

namespace vgb200 {

inline std::vector<int>& selected_gpus() {  // set by the CLI's --gpu (main.cu:141-143)
    static std::vector<int> gpus{0};
    return gpus;
}

// One device index per host map, built on first use and reused for every sample: the reference
// constructs a FastqKmer per sample (src/varigraph.cpp:185-198) but the map lives for the process.
class DeviceGraphIndex {
public:
    static DeviceGraphIndex& get(std::unordered_map<uint64_t, kmerCovFreBitVec>& map, uint32_t k, int buffer_mb) {
        static DeviceGraphIndex inst;
        if (inst.map_ != &map || inst.size_ != map.size() || inst.k_ != k) inst.rebuild(map, k, selected_gpus(), buffer_mb);
        return inst;
    }
    size_t ngpus() const { return ix_.size(); }
    vg_index* index(size_t g = 0) const { return ix_[g]; }
    size_t size() const { return size_; }          // map entries, in the iteration order captured at build time
    size_t slots() const { return cptr_.size(); }  // entries of the device's count vector
    bool has_flags = false;  // vg_index_set_flags done for this graph (see VarigraphKernel)

    // c[s] belongs to the map entry whose k-mer sits at position s of the device's slot-order count vector
    // (vg_index_slot_perm, captured once per graph): the per-sample result needs no reordering on the device
    void write_back(const std::vector<uint8_t>& c, uint32_t threads) {
        const size_t m = cptr_.size();
        const uint32_t nt = std::max<uint32_t>(1, std::min<uint32_t>(threads, 64));
        std::vector<std::thread> pool;
        for (uint32_t t = 0; t < nt; ++t)
            pool.emplace_back([&, t] {
                for (size_t s = m * t / nt, e = m * (t + 1) / nt; s < e; ++s)
                    if (cptr_[s]) *cptr_[s] = c[s];
            });
        for (auto& th : pool) th.join();
    }
    // the map entries in the order the device index knows them
    template <typename Fn>
    void for_each_entry(std::unordered_map<uint64_t, kmerCovFreBitVec>& map, Fn fn) {
        size_t i = 0;
        for (auto& kv : map) fn(i++, kv.second);
    }
    void set_flags(const std::vector<uint8_t>& flags) {
        for (vg_index* ix : ix_) VGB200_CHECK(vg_index_set_flags(ix, flags.data()));
        has_flags = true;
    }

    // One sample over ALL GPUs: the chunks of its files are dealt to the replicas as they come, the counts are combined
    // in slot order over NVLink.  hist256 (optional): the device histogram over the flagged entries.
    void count_sample(const std::vector<std::string>& files, uint32_t threads, uint64_t& readBase, std::vector<uint8_t>& c,
                      uint64_t* hist256 = nullptr) {
        std::vector<const char*> paths;
        for (const auto& f : files) paths.push_back(f.c_str());
        c.resize(slots());
        if (ix_.size() == 1) {
            VGB200_CHECK(vg_count_begin(ix_[0]));
            VGB200_CHECK(vg_count_files(ix_[0], paths.data(), (int)paths.size(), (int)threads, &readBase));
            if (hist256) VGB200_CHECK(vg_count_histogram(ix_[0], hist256));
            VGB200_CHECK(vg_count_end_slots(ix_[0], c.data(), nullptr, nullptr));
            return;
        }
        for (vg_index* ix : ix_) VGB200_CHECK(vg_count_begin(ix));
        VGB200_CHECK(vg_count_files_multi(ix_.data(), (int)ix_.size(), paths.data(), (int)paths.size(), (int)threads, &readBase));
        on_every_rank([&](size_t g) {  // collective: one host thread per rank
            VGB200_CHECK(vg_count_allreduce_slots(comm_[g], ix_[g], g == 0 ? c.data() : nullptr, nullptr));
        });
        if (hist256) VGB200_CHECK(vg_count_histogram(ix_[0], hist256));  // every replica now holds the combined counts
        for (vg_index* ix : ix_) VGB200_CHECK(vg_count_end(ix, nullptr, nullptr, nullptr));
    }
    // One sample on GPU g alone (samples dealt over the GPUs); safe to call for different g from different threads.
    void count_sample_on(size_t g, const std::vector<std::string>& files, uint32_t threads, uint64_t& readBase,
                         std::vector<uint8_t>& c, uint64_t* hist256) {
        std::vector<const char*> paths;
        for (const auto& f : files) paths.push_back(f.c_str());
        c.resize(slots());
        VGB200_CHECK(vg_count_begin(ix_[g]));
        VGB200_CHECK(vg_count_files(ix_[g], paths.data(), (int)paths.size(), (int)threads, &readBase));
        if (hist256) VGB200_CHECK(vg_count_histogram(ix_[g], hist256));
        VGB200_CHECK(vg_count_end_slots(ix_[g], c.data(), nullptr, nullptr));
    }
    ~DeviceGraphIndex() { release(); }

private:
    template <typename Fn>
    void on_every_rank(Fn fn) {
        std::vector<std::thread> pool;
        for (size_t g = 0; g < ix_.size(); ++g) pool.emplace_back([&, g] { fn(g); });
        for (auto& th : pool) th.join();
    }
    void release() {
        for (size_t g = ix_.size(); g-- > 0;)
            if (ix_[g]) vg_index_destroy(ix_[g]);
        for (vg_comm* cm : comm_)
            if (cm) vg_comm_destroy(cm);
        for (vg_ctx* c : ctx_)
            if (c) vg_ctx_destroy(c);
        ix_.clear();
        comm_.clear();
        ctx_.clear();
    }
    void rebuild(std::unordered_map<uint64_t, kmerCovFreBitVec>& map, uint32_t k, const std::vector<int>& gpus, int buffer_mb) {
        release();
        std::cerr << "[" << __func__ << "::" << getTime() << "] " << "Building the device k-mer index ("
                  << map.size() << " k-mers) on GPU " << gpus[0];
        if (gpus.size() > 1) std::cerr << ", replicas on " << gpus.size() - 1 << " more";
        std::cerr << " ...\n";
        std::vector<uint64_t> keys;
        keys.reserve(map.size());
        for (auto& kv : map) keys.push_back(kv.first);
        ctx_.assign(gpus.size(), nullptr);
        ix_.assign(gpus.size(), nullptr);
        for (size_t g = 0; g < gpus.size(); ++g) VGB200_CHECK(vg_ctx_create(gpus[g], buffer_mb, &ctx_[g]));
        VGB200_CHECK(vg_index_create(ctx_[0], keys.data(), keys.size(), k, 0.0, &ix_[0]));
        std::vector<uint32_t> perm(keys.size());
        VGB200_CHECK(vg_index_slot_perm(ix_[0], perm.data()));
        std::vector<uint64_t>().swap(keys);
        cptr_.assign(vg_index_slots(ix_[0]), nullptr);
        size_t i = 0;
        for (auto& kv : map) {
            const uint32_t s = perm[i++];
            if (s != 0xffffffffu) cptr_[s] = &kv.second.c;  // else: a key no read can produce; its c stays 0
        }
        if (gpus.size() > 1) {
            if (vg_index_partitions(ix_[0]) == 0) {
                std::cerr << "[" << __func__ << "::" << getTime() << "] " << "the index is too small to be worth more than one GPU; using GPU "
                          << gpus[0] << " only\n";
                for (size_t g = 1; g < gpus.size(); ++g) vg_ctx_destroy(ctx_[g]);
                ctx_.resize(1);
                ix_.resize(1);
            } else {
                comm_.assign(gpus.size(), nullptr);
                VGB200_CHECK(vg_comm_create_local(ctx_.data(), (int)gpus.size(), vg_index_slots(ix_[0]) + (4u << 20), comm_.data()));
                on_every_rank([&](size_t g) { VGB200_CHECK(vg_index_replicate(comm_[g], 0, g == 0 ? ix_[0] : nullptr, &ix_[g])); });
            }
        }
        map_ = &map;
        size_ = map.size();
        k_ = k;
        has_flags = false;
    }
    const void* map_ = nullptr;
    size_t size_ = 0;
    uint32_t k_ = 0;
    std::vector<vg_ctx*> ctx_;
    std::vector<vg_comm*> comm_;
    std::vector<vg_index*> ix_;
    std::vector<uint8_t*> cptr_;
};

}  // namespace vgb200

class FastqKmerKernel : public FastqKmer {
public:
    int buffer_ = 500;  // Buffer size (MB), as include/fastq_kmer.cuh:12

    FastqKmerKernel(
        unordered_map<uint64_t, kmerCovFreBitVec>& GraphKmerHashHapStrMap,
        const vector<string>& fastqFileNameVec,
        const uint32_t& kmerLen,
        const uint32_t& threads,
        const int buffer
    ) : FastqKmer(GraphKmerHashHapStrMap, fastqFileNameVec, kmerLen, threads) {
        buffer_ = buffer;
    }
    ~FastqKmerKernel() {}

    // src/fastq_kmer.cu:20-31 + :43-270 (fastq_file_open_kernel per file).  hist256 (optional): the sample's
    // coverage histogram over the entries flagged with DeviceGraphIndex::set_flags, taken on the device.
    void build_fastq_index_kernel(uint64_t* hist256 = nullptr) {
        if (fastqFileNameVec_.empty()) {
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Parameter error: -f\n";
            exit(1);
        }
        auto& dev = vgb200::DeviceGraphIndex::get(GraphKmerHashHapStrMap_, kmerLen_, buffer_);
        for (const auto& f : fastqFileNameVec_)
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Collecting kmers from read on GPU: " << f << endl;
        uint64_t readBase = 0;
        vector<uint8_t> c;
        dev.count_sample(fastqFileNameVec_, threads_, readBase, c, hist256);
        mReadBase += readBase;
        dev.write_back(c, threads_);
        malloc_trim(0);
    }
};
