// counting_bloom_filter_b200.hpp -- drop-in for include/counting_bloom_filter.cuh +
// src/counting_bloom_filter.cu: BloomFilterKernel keeps the host class (sizing, seeds, _filter,
// count/find: src/counting_bloom_filter.cpp:28-98) and adds a device twin fed through the C ABI.
// The host draws the seeds (random_device, :80-87); the device receives them, so the downloaded
// filter is byte-identical to what BloomFilter::add would have built with those seeds.
#pragma once
#include <string>
#include <vector>

#include "counting_bloom_filter.hpp"  // reference header: BloomFilter
#include "vgb200.h"
#include "vgb200_check.hpp"

class BloomFilterKernel : public BloomFilter {
public:
    BloomFilterKernel() : BloomFilter() {}
    BloomFilterKernel(uint64_t size, double errorRate, int gpu = 0, int buffer_mb = 100) : BloomFilter(size, errorRate) {
        VGB200_CHECK(vg_ctx_create(gpu, buffer_mb, &ctx_));
        VGB200_CHECK(vg_cbf_create(ctx_, _size, _numHashes, _seeds.data(), &cbf_));
    }
    ~BloomFilterKernel() { release_device(); }

    // The device twin (about 9.6 bytes per genome base) is only needed until the filter has been downloaded.
    void release_device() {
        if (cbf_) vg_cbf_destroy(cbf_);
        if (ctx_) vg_ctx_destroy(ctx_);
        cbf_ = nullptr;
        ctx_ = nullptr;
    }

    // kmer_sketch_bf (src/kmer.cpp:20-52) for one chromosome, fused with the filter update on the
    // device; replaces the kmer_sketch_kernel + add_kernel pair of src/construct_index.cu:69-84.
    uint64_t add_sequence_kernel(const std::string& sequence, uint32_t kmerLen) {
        uint64_t added = 0;
        VGB200_CHECK(vg_cbf_add_sequence(cbf_, sequence.data(), sequence.size(), kmerLen, &added));
        return added;
    }

    // include/counting_bloom_filter.cuh:69-82: afterwards the inherited count()/find() work on _filter
    void copyFilterDToHost() {
        VGB200_CHECK(vg_cbf_download(cbf_, _filter));
        cerr << "[" << __func__ << "::" << getTime() << "] " << "Counting Bloom Filter copied from device to host ...\n";
    }

    // batched BloomFilter::count / find (src/counting_bloom_filter.cpp:40-67) on the device filter
    void count_kernel(const std::vector<uint64_t>& kmers, std::vector<uint8_t>& counts) {
        counts.resize(kmers.size());
        VGB200_CHECK(vg_cbf_query(cbf_, kmers.data(), kmers.size(), counts.data(), nullptr));
    }
    void find_kernel(const std::vector<uint64_t>& kmers, std::vector<uint8_t>& found) {
        found.resize(kmers.size());
        VGB200_CHECK(vg_cbf_query(cbf_, kmers.data(), kmers.size(), nullptr, found.data()));
    }

private:
    vg_ctx* ctx_ = nullptr;
    vg_cbf* cbf_ = nullptr;
};
