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

// One device index per host map, built on first use and reused for every sample: the reference
// constructs a FastqKmer per sample (src/varigraph.cpp:185-198) but the map lives for the process.
class DeviceGraphIndex {
public:
    static DeviceGraphIndex& get(std::unordered_map<uint64_t, kmerCovFreBitVec>& map, uint32_t k, int gpu, int buffer_mb) {
        static DeviceGraphIndex inst;
        if (inst.map_ != &map || inst.size_ != map.size() || inst.k_ != k) inst.rebuild(map, k, gpu, buffer_mb);
        return inst;
    }
    vg_index* index() const { return ix_; }
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
    size_t size() const { return size_; }          // map entries, in the iteration order captured at build time
    size_t slots() const { return cptr_.size(); }  // entries of the device's count vector
    bool has_flags = false;  // vg_index_set_flags done for this graph (see VarigraphKernel)
    // the map entries in the order the device index knows them
    template <typename Fn>
    void for_each_entry(std::unordered_map<uint64_t, kmerCovFreBitVec>& map, Fn fn) {
        size_t i = 0;
        for (auto& kv : map) fn(i++, kv.second);
    }
    ~DeviceGraphIndex() { release(); }

private:
    void release() {
        if (ix_) vg_index_destroy(ix_);
        if (ctx_) vg_ctx_destroy(ctx_);
        ix_ = nullptr;
        ctx_ = nullptr;
    }
    void rebuild(std::unordered_map<uint64_t, kmerCovFreBitVec>& map, uint32_t k, int gpu, int buffer_mb) {
        release();
        std::cerr << "[" << __func__ << "::" << getTime() << "] " << "Building the device k-mer index ("
                  << map.size() << " k-mers) on GPU " << gpu << " ...\n";
        std::vector<uint64_t> keys;
        keys.reserve(map.size());
        for (auto& kv : map) keys.push_back(kv.first);
        VGB200_CHECK(vg_ctx_create(gpu, buffer_mb, &ctx_));
        VGB200_CHECK(vg_index_create(ctx_, keys.data(), keys.size(), k, 0.0, &ix_));
        std::vector<uint32_t> perm(keys.size());
        VGB200_CHECK(vg_index_slot_perm(ix_, perm.data()));
        std::vector<uint64_t>().swap(keys);
        cptr_.assign(vg_index_slots(ix_), nullptr);
        size_t i = 0;
        for (auto& kv : map) {
            const uint32_t s = perm[i++];
            if (s != 0xffffffffu) cptr_[s] = &kv.second.c;  // else: a key no read can produce; its c stays 0
        }
        map_ = &map;
        size_ = map.size();
        k_ = k;
        has_flags = false;
    }
    const void* map_ = nullptr;
    size_t size_ = 0;
    uint32_t k_ = 0;
    vg_ctx* ctx_ = nullptr;
    vg_index* ix_ = nullptr;
    std::vector<uint8_t*> cptr_;
};

inline int& selected_gpu() {  // set by the CLI's --gpu (main.cu:141-143)
    static int gpu = 0;
    return gpu;
}

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

    // src/fastq_kmer.cu:20-31 + :43-270 (fastq_file_open_kernel per file)
    void build_fastq_index_kernel() {
        if (fastqFileNameVec_.empty()) {
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Parameter error: -f\n";
            exit(1);
        }
        auto& dev = vgb200::DeviceGraphIndex::get(GraphKmerHashHapStrMap_, kmerLen_, vgb200::selected_gpu(), buffer_);
        vector<const char*> paths;
        for (const auto& f : fastqFileNameVec_) {
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Collecting kmers from read on GPU: " << f << endl;
            paths.push_back(f.c_str());
        }
        VGB200_CHECK(vg_count_begin(dev.index()));
        uint64_t readBase = 0;
        VGB200_CHECK(vg_count_files(dev.index(), paths.data(), (int)paths.size(), (int)threads_, &readBase));
        vector<uint8_t> c(dev.slots());
        VGB200_CHECK(vg_count_end_slots(dev.index(), c.data(), nullptr, nullptr));
        mReadBase += readBase;
        dev.write_back(c, threads_);
        malloc_trim(0);
    }
};
