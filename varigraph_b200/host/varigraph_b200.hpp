// varigraph_b200.hpp -- drop-in for include/varigraph.cuh + src/varigraph.cu: the orchestrator
// subclass whose only job is to call the *Kernel classes at the two seams of the hot path
// (src/varigraph.cu:14-60 construct_kernel, :62-118 fastq_genotype_kernel / kmer_read_kernel).
// Everything else -- load(), parse_sample_config(), cal_ave_cov_kmer(), genotype() -- is the
// reference's host code, inherited unmodified.
#pragma once
#include "construct_index_b200.hpp"
#include "fastq_kmer_b200.hpp"
#include <condition_variable>
#include <mutex>

#include "varigraph.hpp"  // reference header

class VarigraphKernelConfig : public VarigraphConfig {
public:
    int gpu;           // GPU ID (the first of --gpu's list)
    vector<int> gpus;  // --gpu a,b,c: the index is replicated over these GPUs
    int buffer;        // staged chunk size in MB
    VarigraphKernelConfig() : VarigraphConfig(), gpu(0), gpus{0}, buffer(100) {}

    void logKernelConfig() const {
        cerr << "[" << __func__ << "::" << getTime() << "] " << "Selected GPU ID: ";
        for (size_t i = 0; i < gpus.size(); i++) cerr << (i ? "," : "") << gpus[i];
        cerr << endl;
        cerr << "[" << __func__ << "::" << getTime() << "] " << "GPU buffer size: " << buffer << " MB" << endl;
    }
};

class VarigraphKernel : public Varigraph {
public:
    int buffer_ = 100;
    int gpu_ = 0;

    VarigraphKernel(const VarigraphKernelConfig& config) : Varigraph(config), buffer_(config.buffer), gpu_(config.gpu) {
        vgb200::selected_gpus() = config.gpus.empty() ? vector<int>{config.gpu} : config.gpus;
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

    // src/varigraph.cu:62-86.  With several GPUs and at least as many samples (BASELINE config 5) the samples are dealt
    // over the GPUs: each GPU counts a sample of its own on its replica of the index while this thread genotypes the
    // previous ones in the order of the sample list (the host map's c and the HMM state exist once, so genotyping
    // stays sequential).  Otherwise every sample's reads are spread over all the GPUs.
    void fastq_genotype_kernel() {
        ConstructIndexClassPtr_->graph2node();
        auto& graphMap = ConstructIndexClassPtr_->mGraphKmerHashHapStrMap;
        auto& dev = vgb200::DeviceGraphIndex::get(graphMap, kmerLen_, buffer_);
        ensure_flags(dev);
        const size_t G = dev.ngpus(), S = sampleConfigTupleVec_.size();
        if (G > 1)
            cerr << "[" << __func__ << "::" << getTime() << "] " << S << " sample(s), " << G << " GPUs: "
                 << (S >= G ? "the samples are dealt over the GPUs" : "every sample's reads are spread over all GPUs") << endl << endl;
        if (G > 1 && S >= G) {
            struct Result {
                vector<uint8_t> c;
                uint64_t readBase = 0;
                uint64_t hist[256];
                bool ready = false;
            };
            vector<Result> res(S);
            std::mutex mu;
            std::condition_variable cv;
            size_t consumed = 0;  // samples the host is done with
            vector<std::thread> workers;
            const uint32_t tpg = std::max<uint32_t>(1, threads_ / (uint32_t)G);
            for (size_t g = 0; g < G; g++)
                workers.emplace_back([&, g] {
                    for (size_t s = g; s < S; s += G) {
                        {   // at most one finished sample per GPU waits for the host
                            std::unique_lock<std::mutex> lk(mu);
                            cv.wait(lk, [&] { return s < consumed + G; });
                        }
                        Result& r = res[s];
                        dev.count_sample_on(g, get<1>(sampleConfigTupleVec_[s]), tpg, r.readBase, r.c, r.hist);
                        {
                            std::lock_guard<std::mutex> lk(mu);
                            r.ready = true;
                        }
                        cv.notify_all();
                    }
                });
            for (size_t s = 0; s < S; s++) {
                const auto& [sampleName, fastqFileNameVec] = sampleConfigTupleVec_[s];
                cerr << "[" << __func__ << "::" << getTime() << "] " << "Processing sample: " << sampleName << " (counted on GPU "
                     << vgb200::selected_gpus()[s % G] << ")" << endl << endl;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return res[s].ready; });
                }
                dev.write_back(res[s].c, threads_);
                vector<uint8_t>().swap(res[s].c);
                after_counting(res[s].readBase, res[s].hist);
                genotype(sampleName);
                cerr << "[" << __func__ << "::" << getTime() << "] " << "Sample: " << sampleName << " has been processed." << endl << endl << endl;
                ConstructIndexClassPtr_->reset();
                {
                    std::lock_guard<std::mutex> lk(mu);
                    consumed = s + 1;
                }
                cv.notify_all();
            }
            for (auto& w : workers) w.join();
            return;
        }
        for (const auto& [sampleName, fastqFileNameVec] : sampleConfigTupleVec_) {
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Processing sample: " << sampleName << endl << endl;
            kmer_read_kernel(fastqFileNameVec);
            genotype(sampleName);
            cerr << "[" << __func__ << "::" << getTime() << "] " << "Sample: " << sampleName << " has been processed." << endl << endl << endl;
            ConstructIndexClassPtr_->reset();
        }
    }

    // Which entries qualify for Varigraph::get_hom_kmer's histogram ("f <= 1 and some sample carries the k-mer on all of
    // its haplotypes", src/varigraph.cpp:253-296) depends on the graph only: computed once per graph and kept on the
    // device(s) as a flag per entry; per sample only 256 numbers come back.
    void ensure_flags(vgb200::DeviceGraphIndex& dev) {
        if (dev.has_flags) return;
        auto& graphMap = ConstructIndexClassPtr_->mGraphKmerHashHapStrMap;
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
        dev.set_flags(flags);
    }

    // Varigraph::cal_ave_cov_kmer (src/varigraph.cpp:220-246) with get_hom_kmer's walk over the host map replaced by the
    // device histogram that came back with the sample's counts.
    void after_counting(uint64_t readBase, const uint64_t* hist) {
        ReadDepth_ = readBase / (float)ConstructIndexClassPtr_->mGenomeSize;
        map<uint8_t, uint64_t> kmerCovFreMap;  // map<coverage, frequency>, c == 0 skipped as the reference does
        for (int c = 1; c < 256; c++)
            if (hist[c]) kmerCovFreMap[(uint8_t)c] = hist[c];
        uint8_t maxCoverage, homCoverage;
        tie(maxCoverage, homCoverage) = get_hom_kmer_c(kmerCovFreMap);
        if (useDepth_) homCoverage = ReadDepth_ * 0.8;
        cal_hap_kmer_cov(homCoverage);
        kmer_histogram(maxCoverage, homCoverage, kmerCovFreMap);
        cerr << "[" << __func__ << "::" << getTime() << "] " << fixed << setprecision(2) << "sequenced "
             << readBase / 1e9 << " Gb, depth " << ReadDepth_ << ", haplotype k-mer coverage "
             << hapKmerCoverage_ << defaultfloat << setprecision(6) << "\n\n";
    }

    // src/varigraph.cu:93-118
    void kmer_read_kernel(vector<string> fastqFileNameVec) {
        auto& dev = vgb200::DeviceGraphIndex::get(ConstructIndexClassPtr_->mGraphKmerHashHapStrMap, kmerLen_, buffer_);
        ensure_flags(dev);
        FastqKmerKernel FastqKmerKernelClass(ConstructIndexClassPtr_->mGraphKmerHashHapStrMap, fastqFileNameVec, kmerLen_, threads_, buffer_);
        uint64_t hist[256];
        FastqKmerKernelClass.build_fastq_index_kernel(hist);
        after_counting(FastqKmerKernelClass.mReadBase, hist);
    }
};
