// varigraph_b200.hpp -- drop-in for include/varigraph.cuh + src/varigraph.cu: the orchestrator
// subclass whose only job is to call the *Kernel classes at the two seams of the hot path
// (src/varigraph.cu:14-60 construct_kernel, :62-118 fastq_genotype_kernel / kmer_read_kernel).
// Everything else -- load(), parse_sample_config(), cal_ave_cov_kmer(), genotype() -- is the
// reference's host code, inherited unmodified.
#pragma once
#include "construct_index_b200.hpp"
#include "fastq_kmer_b200.hpp"
#include "varigraph.hpp"  // reference header

class VarigraphKernelConfig : public VarigraphConfig {
public:
    int gpu;     // GPU ID
    int buffer;  // staged chunk size in MB
    VarigraphKernelConfig() : VarigraphConfig(), gpu(0), buffer(100) {}

    void logKernelConfig() const {
        cerr << "[" << __func__ << "::" << getTime() << "] " << "Selected GPU ID: " << gpu << endl;
        cerr << "[" << __func__ << "::" << getTime() << "] " << "GPU buffer size: " << buffer << " MB" << endl;
    }
};

class VarigraphKernel : public Varigraph {
public:
    int buffer_ = 100;
    int gpu_ = 0;

    VarigraphKernel(const VarigraphKernelConfig& config) : Varigraph(config), buffer_(config.buffer), gpu_(config.gpu) {
        vgb200::selected_gpu() = config.gpu;
    }

    // src/varigraph.cu:14-60: same step order as Varigraph::construct, make_mbf on the device
    void construct_kernel() {
        ConstructIndexKernel* ci = new ConstructIndexKernel(refFileName_, vcfFileName_, inputGraphFileName_, outputGraphFileName_,
                                                           fastMode_, useUniqueKmers_, kmerLen_, vcfPloidy_, debug_, threads_,
                                                           buffer_, gpu_);
        ConstructIndexClassPtr_ = ci;  // freed by ~Varigraph
        ci->build_fasta_index();
        ci->make_mbf_kernel();
        ci->construct();
        ci->make_QRmap();
        ci->index_kernel();
        ci->save_index();
        ci->clear_mbf_kernel();
        cerr << "[" << __func__ << "::" << getTime() << "] " << "graph: " << ci->mGraphBaseNum << " bases, "
             << ci->mGraphKmerHashHapStrMap.size() << " k-mers, " << ci->mHapMap.size() << " haplotypes\n\n";
    }

    // src/varigraph.cu:62-86
    void fastq_genotype_kernel() {
        ConstructIndexClassPtr_->graph2node();
        for (const auto& [sampleName, fastqFileNameVec] : sampleConfigTupleVec_) {
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Processing sample: " << sampleName << endl << endl;
            kmer_read_kernel(fastqFileNameVec);
            genotype(sampleName);
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Sample: " << sampleName << " has been processed." << endl << endl << endl;
            ConstructIndexClassPtr_->reset();
        }
    }

    // Varigraph::cal_ave_cov_kmer (src/varigraph.cpp:220-246) with get_hom_kmer's walk over the host map
    // (:253-296) replaced by a device histogram: which entries qualify ("f <= 1 and some sample carries
    // the k-mer on all of its haplotypes") depends on the graph only, so it is computed once per graph
    // and kept on the device as a flag per entry; per sample only 256 numbers come back.
    void cal_ave_cov_kmer_kernel() {
        auto& graphMap = ConstructIndexClassPtr_->mGraphKmerHashHapStrMap;
        auto& dev = vgb200::DeviceGraphIndex::get(graphMap, kmerLen_, gpu_, buffer_);
        if (!dev.has_flags) {
            const auto& hapIdxQRmap = ConstructIndexClassPtr_->mHapIdxQRmap;
            const size_t hapNum = ConstructIndexClassPtr_->mHapNum;
            vector<uint8_t> flags(dev.size(), 0);
            dev.for_each_entry(graphMap, [&](size_t i, const kmerCovFreBitVec& e) {
                if (e.f > 1) return;
                uint32_t inSample = 0, carried = 0;
                for (size_t h = 1; h < hapNum; h++) {
                    const auto& qr = hapIdxQRmap.at(h);
                    if (construct_index::get_bit(e.BitVec[get<0>(qr)], get<1>(qr)) > 0) carried++;
                    if (++inSample == vcfPloidy_) {
                        if (carried == vcfPloidy_) { flags[i] = 1; break; }
                        inSample = carried = 0;
                    }
                }
            });
            VGB200_CHECK(vg_index_set_flags(dev.index(), flags.data()));
            dev.has_flags = true;
        }
        uint64_t hist[256];
        VGB200_CHECK(vg_count_histogram(dev.index(), hist));
        map<uint8_t, uint64_t> kmerCovFreMap;  // map<coverage, frequency>, c == 0 skipped as the reference does
        for (int c = 1; c < 256; c++)
            if (hist[c]) kmerCovFreMap[(uint8_t)c] = hist[c];
        uint8_t maxCoverage, homCoverage;
        tie(maxCoverage, homCoverage) = get_hom_kmer_c(kmerCovFreMap);
        if (useDepth_) homCoverage = ReadDepth_ * 0.8;
        cal_hap_kmer_cov(homCoverage);
        kmer_histogram(maxCoverage, homCoverage, kmerCovFreMap);
    }

    // src/varigraph.cu:93-118
    void kmer_read_kernel(vector<string> fastqFileNameVec) {
        FastqKmerKernel FastqKmerKernelClass(ConstructIndexClassPtr_->mGraphKmerHashHapStrMap, fastqFileNameVec, kmerLen_, threads_, buffer_);
        FastqKmerKernelClass.build_fastq_index_kernel();
        ReadDepth_ = FastqKmerKernelClass.mReadBase / (float)ConstructIndexClassPtr_->mGenomeSize;
        cal_ave_cov_kmer_kernel();
        cerr << "[" << __func__ << "::" << getTime() << "] " << fixed << setprecision(2) << "sequenced "
             << FastqKmerKernelClass.mReadBase / 1e9 << " Gb, depth " << ReadDepth_ << ", haplotype k-mer coverage "
             << hapKmerCoverage_ << defaultfloat << setprecision(6) << "\n\n";
    }
};
