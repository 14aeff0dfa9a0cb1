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
        ci->clear_mbf();
        cerr << endl;
        cerr << "           - " << "Total number of bases in the Genome Graph: " << ci->mGraphBaseNum << endl;
        cerr << "           - " << "Total number of k-mers present in the Genome Graph: " << ci->mGraphKmerHashHapStrMap.size() << endl;
        cerr << "           - " << "Total number of haplotypes present in the Genome Graph: " << ci->mHapMap.size() << endl << endl << endl;
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

    // src/varigraph.cu:93-118
    void kmer_read_kernel(vector<string> fastqFileNameVec) {
        FastqKmerKernel FastqKmerKernelClass(ConstructIndexClassPtr_->mGraphKmerHashHapStrMap, fastqFileNameVec, kmerLen_, threads_, buffer_);
        FastqKmerKernelClass.build_fastq_index_kernel();
        ReadDepth_ = FastqKmerKernelClass.mReadBase / (float)ConstructIndexClassPtr_->mGenomeSize;
        cal_ave_cov_kmer();
        cerr << endl;
        cerr << fixed << setprecision(2);
        cerr << "           - " << "Size of the sequenced data: " << FastqKmerKernelClass.mReadBase / 1e9 << " Gb" << endl;
        cerr << "           - " << "Depth of the sequenced data: " << ReadDepth_ << endl;
        cerr << "           - " << "Coverage of haplotype k-mers: " << hapKmerCoverage_ << endl << endl << endl;
        cerr << defaultfloat << setprecision(6);
    }
};
