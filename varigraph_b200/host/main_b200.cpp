// main_b200.cpp -- `varigraph construct` / `varigraph genotype --gpu --buffer`: the reference's
// GPU-build CLI (main.cu:32-56,74-242,274-471) over the B200 classes in this directory.
// Built only where the reference tree is present (oracle/Makefile target `integrated`): this file
// and the headers beside it are ours; ConstructIndex, GENOTYPE, HaplotypeSelect, SAVE ... are the
// reference's own sources, compiled where they lie.
#include <getopt.h>

#include <cstdio>
#include <iostream>
#include <string>

#include "sys.hpp"  // reference header: realtime(), cputime(), peakrss()
#include "varigraph_b200.hpp"

using namespace std;

static const char* kVersion = "1.0.8-b200";

static void usage(const char* prog) {
    cerr << "Usage: " << prog << " construct -r FILE -v FILE [--save-graph FILE] [--vcf-ploidy INT] [-k INT] [--fast]\n"
         << "                 [--use-unique-kmers] [--gpu INT[,INT...]] [--buffer MB] [-D] [-t INT]\n"
         << "       " << prog << " genotype --load-graph FILE -s FILE [-g hom|het] [--sample-ploidy INT] [-n INT]\n"
         << "                 [--granularity FLOAT] [-m fre|rec] [--sv] [--min-support FLOAT] [--use-depth]\n"
         << "                 [--gpu INT[,INT...]] [--buffer MB] [-D] [-t INT]\n"
         << "Options and defaults are those of varigraph v1.0.8 (GPU build); k-mer counting runs on a B200.\n";
}

static int fail(const char* prog, const string& what) {
    cerr << "[main::" << getTime() << "] Parameter error: " << what << "\n\n";
    usage(prog);
    return 1;
}

enum { OPT_GRAPH = 1000, OPT_VCF_PLOIDY, OPT_FAST, OPT_UNIQUE, OPT_SAMPLE_PLOIDY, OPT_GRAN, OPT_SV, OPT_MINSUP, OPT_DEPTH, OPT_GPU, OPT_BUFFER };

int main(int argc, char** argv) {
    if (argc < 2) { usage(argv[0]); return 1; }
    const string cmd = argv[1];
    if (cmd == "-h" || cmd == "--help") { usage(argv[0]); return 0; }
    if (cmd != "construct" && cmd != "genotype") return fail(argv[0], "unknown subcommand '" + cmd + "'");
    const bool construct = cmd == "construct";
    const double t0 = realtime();
    VarigraphKernelConfig cfg;
    static const option opts[] = {
        {"reference", required_argument, 0, 'r'}, {"vcf", required_argument, 0, 'v'},
        {"save-graph", required_argument, 0, OPT_GRAPH}, {"load-graph", required_argument, 0, OPT_GRAPH},
        {"vcf-ploidy", required_argument, 0, OPT_VCF_PLOIDY}, {"kmer", required_argument, 0, 'k'},
        {"fast", no_argument, 0, OPT_FAST}, {"use-unique-kmers", no_argument, 0, OPT_UNIQUE},
        {"samples", required_argument, 0, 's'}, {"sample", required_argument, 0, 's'},
        {"genotype", required_argument, 0, 'g'}, {"sample-ploidy", required_argument, 0, OPT_SAMPLE_PLOIDY},
        {"number", required_argument, 0, 'n'}, {"granularity", required_argument, 0, OPT_GRAN},
        {"mode", required_argument, 0, 'm'}, {"sv", no_argument, 0, OPT_SV},
        {"min-support", required_argument, 0, OPT_MINSUP}, {"use-depth", no_argument, 0, OPT_DEPTH},
        {"gpu", required_argument, 0, OPT_GPU}, {"buffer", required_argument, 0, OPT_BUFFER},
        {"debug", no_argument, 0, 'D'}, {"threads", required_argument, 0, 't'}, {"help", no_argument, 0, 'h'},
        {0, 0, 0, 0}};
    optind = 2;
    for (int c; (c = getopt_long(argc, argv, "r:v:k:s:g:n:m:Dt:h", opts, nullptr)) != -1;) {
        switch (c) {
            case 'r': cfg.refFileName = optarg; break;
            case 'v': cfg.vcfFileName = optarg; break;
            case OPT_GRAPH: (construct ? cfg.outputGraphFileName : cfg.inputGraphFileName) = optarg; break;
            case OPT_VCF_PLOIDY: cfg.vcfPloidy = max(stoi(optarg), 2); break;
            case 'k': cfg.kmerLen = max(stoi(optarg), 5); break;
            case OPT_FAST: cfg.fastMode = true; break;
            case OPT_UNIQUE: cfg.useUniqueKmers = true; break;
            case 's': cfg.samplesConfigFileName = optarg; break;
            case 'g': cfg.sampleType = optarg; break;
            case OPT_SAMPLE_PLOIDY: cfg.samplePloidy = max(stoi(optarg), 2); break;
            case 'n': cfg.haploidNum = stoull(optarg); break;
            case OPT_GRAN: cfg.chrLenThread = stof(optarg) * 1e6; break;
            case 'm': cfg.transitionProType = optarg; break;
            case OPT_SV: cfg.svGenotypeBool = true; break;
            case OPT_MINSUP: cfg.minSupportingGQ = stof(optarg); break;
            case OPT_DEPTH: cfg.useDepth = true; break;
            case OPT_GPU: {  // one id (main.cu:141-143) or a list: 0,1,2,3
                cfg.gpus.clear();
                string tok;
                for (const char* q = optarg;; ++q) {
                    if (*q == ',' || *q == 0) {
                        if (tok.empty() || tok.find_first_not_of("0123456789") != string::npos) return fail(argv[0], "--gpu");
                        cfg.gpus.push_back(stoi(tok));
                        tok.clear();
                        if (*q == 0) break;
                    } else {
                        tok += *q;
                    }
                }
                cfg.gpu = cfg.gpus[0];
                break;
            }
            case OPT_BUFFER: cfg.buffer = stoi(optarg); break;
            case 'D': cfg.debug = true; break;
            case 't': cfg.threads = max(stoi(optarg), 1); break;
            default: usage(argv[0]); return c == 'h' ? 0 : 1;
        }
    }
    if (cfg.gpu < 0) return fail(argv[0], "--gpu");
    if (cfg.buffer <= 0) return fail(argv[0], "--buffer must be greater than 0");
    if (construct) {
        if (cfg.refFileName.empty()) return fail(argv[0], "-r");
        if (cfg.vcfFileName.empty()) return fail(argv[0], "-v");
        if (cfg.vcfPloidy == 0 || cfg.vcfPloidy > 8) return fail(argv[0], "--vcf-ploidy must be between 2 and 8");
        if (cfg.kmerLen == 0 || cfg.kmerLen > 28) return fail(argv[0], "-k must be at most 28");
    } else {
        if (cfg.samplesConfigFileName.empty()) return fail(argv[0], "-s");
        if (cfg.sampleType != "hom" && cfg.sampleType != "het") return fail(argv[0], "-g must be hom or het");
        if (cfg.samplePloidy == 0 || cfg.samplePloidy > 8) return fail(argv[0], "--sample-ploidy must be between 2 and 8");
        if (cfg.haploidNum == 0) return fail(argv[0], "-n must be greater than 0");
        if (cfg.chrLenThread < 1) return fail(argv[0], "--granularity");
        if (cfg.transitionProType != "fre" && cfg.transitionProType != "rec") return fail(argv[0], "-m must be fre or rec");
    }
    cerr << "[main::" << getTime() << "] You are now running varigraph (v" << kVersion << ").\n\n\n";
    if (construct) cfg.logConstructionConfig(); else cfg.logGenotypeConfig();
    cfg.logKernelConfig();

    VarigraphKernel vg(cfg);
    if (construct) {
        vg.construct_kernel();
    } else {
        vg.parse_sample_config();
        vg.load();
        vg.fastq_genotype_kernel();
    }
    cerr << "[main::" << getTime() << "] Done ...\n\n\n";
    fprintf(stderr, "[varigraph::main] Real time: %.3f sec; CPU: %.3f sec; Peak RSS: %.3f GB\n", realtime() - t0, cputime(),
            peakrss() / 1024.0 / 1024.0 / 1024.0);
    return 0;
}
