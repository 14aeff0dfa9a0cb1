// construct_index_b200.hpp -- drop-in for include/construct_index.cuh + src/construct_index.cu.
// Only make_mbf differs from the CPU class: the counting Bloom filter is filled on the device
// (src/construct_index.cu:39-106) and copied back, after which the reference's own index()
// (src/construct_index.cpp:592-700) runs unmodified against mbf -- index_kernel is, as in the
// reference, the host code path (src/construct_index.cu:116-300 logs "on CPU").
#pragma once
#include "construct_index.hpp"  // reference header
#include "counting_bloom_filter_b200.hpp"

class ConstructIndexKernel : public ConstructIndex {
public:
    BloomFilterKernel* mbfD = nullptr;  // owned through the base class pointer `mbf`
    int buffer_ = 100;
    int gpu_ = 0;

    ConstructIndexKernel(
        const string& refFileName, const string& vcfFileName, const string& inputGraphFileName,
        const string& outputGraphFileName, const bool& fastMode, const bool& useUniqueKmers,
        const uint32_t& kmerLen, const uint32_t& vcfPloidy, const bool& debug, const uint32_t& threads,
        const int buffer, const int gpu = 0
    ) : ConstructIndex(refFileName, vcfFileName, inputGraphFileName, outputGraphFileName, fastMode, useUniqueKmers,
                       kmerLen, vcfPloidy, debug, threads), buffer_(buffer), gpu_(gpu) {}

    void make_mbf_kernel() {
        const uint64_t n = mGenomeSize - mKmerLen + 1;  // as ConstructIndex::make_mbf sizes it (p = 0.01)
        mbfD = new BloomFilterKernel(n, 0.01, gpu_, buffer_);
        mbf = mbfD;  // ConstructIndex::index() / clear_mbf() use and free it through the base pointer
        cerr << "[" << __func__ << "::" << getTime() << "] " << "Counting reference k-mers into the Bloom filter on GPU "
             << gpu_ << " (" << mbf->get_size() << " cells, " << mbf->get_num() << " hashes) ...\n";
        uint64_t added = 0;
        for (const auto& [chromosome, sequence] : mFastaSeqMap) added += mbfD->add_sequence_kernel(sequence, mKmerLen);
        mbfD->copyFilterDToHost();
        mbfD->release_device();  // index() only needs the host copy (inherited count / find)
        cerr << "[" << __func__ << "::" << getTime() << "] " << added << " k-mers added; filter usage rate "
             << fixed << setprecision(2) << mbf->get_cap() << defaultfloat << setprecision(6) << "\n\n";
        malloc_trim(0);
    }

    // ConstructIndex::clear_mbf (src/construct_index.cpp:53-60) deletes `mbf` through BloomFilter*, whose destructor
    // is not virtual: delete the filter as what it is, then let the base class do the rest.
    void clear_mbf_kernel() {
        delete mbfD;
        mbfD = nullptr;
        mbf = nullptr;
        clear_mbf();
    }
    ~ConstructIndexKernel() {
        if (mbfD) {
            delete mbfD;
            mbfD = nullptr;
            mbf = nullptr;
        }
    }

    void index_kernel() { index(); }
};
