// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-ABI window onto the UNMODIFIED reference implementation.  This file is
// ours; it #includes the reference's headers and is linked against the
// reference's own src/*.cpp where they lie under $(REF) (see oracle/Makefile).
// Output: oracle/_ref/libvgref.so.  Used by tests/ and gen_golden.py to pin
// oracle/vg_oracle.c, and by bench.py's reference arm to time the
// reference's own CPU count phase.
//
// Every entry point calls one reference function directly:
//   ref_hash64               -> include/hash64.hpp:5-14
//   ref_nt4                  -> include/seq_nt4_table.hpp:5-22
//   ref_murmur3_x64_128_sum  -> src/MurmurHash3.cpp:255-332 as used by
//                               src/counting_bloom_filter.cpp:90-98
//   ref_sketch               -> src/kmer.cpp:110-149 (kmer_sketch_fastq) with a
//                               map holding exactly the keys kmer_sketch_genotype
//                               (src/kmer.cpp:162-198) emits for the same string
//   ref_cbf_*                -> src/counting_bloom_filter.cpp:28-98 (seeds injected
//                               through a subclass: _seeds/_filter are protected)
//   ref_cbf_fill             -> src/kmer.cpp:20-52 (kmer_sketch_bf)
//   ref_graph_*              -> src/construct_index.cpp:911 (load_index)
//   ref_count_files          -> src/fastq_kmer.cpp:41-187 (build_fastq_index)
//   ref_map_*                -> the same FastqKmer::build_fastq_index over a caller-supplied key set
#include <chrono>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "construct_index.hpp"
#include "counting_bloom_filter.hpp"
#include "fastq_kmer.hpp"
#include "hash64.hpp"
#include "kmer.hpp"
#include "seq_nt4_table.hpp"
#include "MurmurHash3.h"

namespace {

struct SeededBloom : public BloomFilter {
    SeededBloom(uint64_t n, double p, const uint64_t* seeds, uint32_t nseeds) : BloomFilter(n, p) {
        if (seeds != nullptr) {
            _seeds.assign(seeds, seeds + nseeds);
            _numHashes = nseeds;
        }
    }
    uint8_t* raw() { return _filter; }
    const std::vector<uint64_t>& seeds() const { return _seeds; }
};

struct GraphHolder {
    std::string ref, vcf, in, out;
    ConstructIndex* ci = nullptr;
    std::vector<uint64_t> keys;  // iteration order snapshot
};

}  // namespace

extern "C" {

uint64_t ref_hash64(uint64_t key, uint64_t mask) { return hash64(key, mask); }

int ref_nt4(int byte) { return seq_nt4_table[(uint8_t)byte]; }

uint64_t ref_murmur3_x64_128_sum(uint64_t key, uint32_t seed) {
    uint64_t out[2];
    MurmurHash3_x64_128(&key, (int)sizeof(key), seed, out);
    return out[0] + out[1];
}

// Ordered, with multiplicity: every key the reference encoder emits for `seq`.
// Returns the count; writes at most `cap` keys.
int64_t ref_sketch(const char* seq, int64_t len, uint32_t k, uint64_t* out, int64_t cap) {
    std::string s(seq, (size_t)len);
    if (s.empty()) return 0;
    std::unordered_set<uint64_t> all = kmerBit::kmer_sketch_genotype(s, k);
    std::unordered_map<uint64_t, kmerCovFreBitVec> m;
    for (uint64_t h : all) m[h];
    std::vector<uint64_t> v = kmerBit::kmer_sketch_fastq(s, k, m);
    int64_t n = (int64_t)v.size();
    for (int64_t i = 0; i < n && i < cap; ++i) out[i] = v[(size_t)i];
    return n;
}

// ---- counting Bloom filter ------------------------------------------------
void* ref_cbf_create(uint64_t n, double p, const uint64_t* seeds, uint32_t nseeds) {
    return new SeededBloom(n, p, seeds, nseeds);
}
void ref_cbf_destroy(void* h) { delete static_cast<SeededBloom*>(h); }
uint64_t ref_cbf_size(void* h) { return static_cast<SeededBloom*>(h)->get_size(); }
uint32_t ref_cbf_num_hashes(void* h) { return static_cast<SeededBloom*>(h)->get_num(); }
void ref_cbf_seeds(void* h, uint64_t* out) {
    const auto& s = static_cast<SeededBloom*>(h)->seeds();
    for (size_t i = 0; i < s.size(); ++i) out[i] = s[i];
}
void ref_cbf_add(void* h, uint64_t key) { static_cast<SeededBloom*>(h)->add(key); }
int ref_cbf_find(void* h, uint64_t key) { return static_cast<SeededBloom*>(h)->find(key) ? 1 : 0; }
int ref_cbf_count(void* h, uint64_t key) { return static_cast<SeededBloom*>(h)->count(key); }
const uint8_t* ref_cbf_filter(void* h) { return static_cast<SeededBloom*>(h)->raw(); }
// Fill from a sequence exactly as ConstructIndex::make_mbf does per chromosome.
void ref_cbf_fill(void* h, const char* seq, int64_t len, uint32_t k) {
    std::string s(seq, (size_t)len);
    if (s.empty()) return;
    kmerBit::kmer_sketch_bf(s, k, static_cast<SeededBloom*>(h));
}

// ---- graph index + count phase -------------------------------------------
void* ref_graph_load(const char* graph_bin, uint32_t threads) {
    GraphHolder* g = new GraphHolder();
    g->in = graph_bin;
    bool fast = false, uniq = false, debug = false;
    uint32_t k = 27, ploidy = 2;
    g->ci = new ConstructIndex(g->ref, g->vcf, g->in, g->out, fast, uniq, k, ploidy, debug, threads);
    g->ci->load_index();
    g->keys.reserve(g->ci->mGraphKmerHashHapStrMap.size());
    for (const auto& kv : g->ci->mGraphKmerHashHapStrMap) g->keys.push_back(kv.first);
    return g;
}
void ref_graph_destroy(void* h) {
    GraphHolder* g = static_cast<GraphHolder*>(h);
    delete g->ci;
    delete g;
}
uint64_t ref_graph_num_kmers(void* h) { return static_cast<GraphHolder*>(h)->keys.size(); }
uint32_t ref_graph_kmer_len(void* h) { return static_cast<GraphHolder*>(h)->ci->mKmerLen; }
void ref_graph_keys(void* h, uint64_t* out) {
    GraphHolder* g = static_cast<GraphHolder*>(h);
    std::memcpy(out, g->keys.data(), g->keys.size() * sizeof(uint64_t));
}
void ref_graph_reset(void* h) { static_cast<GraphHolder*>(h)->ci->reset(); }

// Runs the reference's CPU count phase over `files`; writes c for every key in
// the order ref_graph_keys() reports, returns wall seconds of
// build_fastq_index() alone.
double ref_count_files(void* h, const char** files, int nfiles, uint32_t threads, uint8_t* c_out,
                       uint64_t* read_bases) {
    GraphHolder* g = static_cast<GraphHolder*>(h);
    std::vector<std::string> fv;
    for (int i = 0; i < nfiles; ++i) fv.emplace_back(files[i]);
    FastqKmer fk(g->ci->mGraphKmerHashHapStrMap, fv, g->ci->mKmerLen, threads);
    auto t0 = std::chrono::steady_clock::now();
    fk.build_fastq_index();
    auto t1 = std::chrono::steady_clock::now();
    if (read_bases) *read_bases = fk.mReadBase;
    if (c_out) {
        for (size_t i = 0; i < g->keys.size(); ++i)
            c_out[i] = g->ci->mGraphKmerHashHapStrMap.find(g->keys[i])->second.c;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- the count phase over an arbitrary key set (bench.py's reference arm) -------------------
// The synthetic bench index is not the output of `construct`, so there is no graph.bin to load;
// FastqKmer only needs the map (include/fastq_kmer.hpp:57-62), so build it directly.
struct MapHolder {
    std::unordered_map<uint64_t, kmerCovFreBitVec> map;
    std::vector<uint64_t> keys;
    uint32_t k;
};
void* ref_map_create(const uint64_t* keys, uint64_t n, uint32_t k) {
    MapHolder* m = new MapHolder();
    m->k = k;
    m->keys.assign(keys, keys + n);
    m->map.reserve(n);
    for (uint64_t i = 0; i < n; ++i) m->map[keys[i]];
    return m;
}
void ref_map_destroy(void* h) { delete static_cast<MapHolder*>(h); }
void ref_map_reset(void* h) {
    for (auto& kv : static_cast<MapHolder*>(h)->map) kv.second.reset();  // as ConstructIndex::reset()
}
double ref_map_count_files(void* h, const char** files, int nfiles, uint32_t threads, uint8_t* c_out,
                           uint64_t* read_bases) {
    MapHolder* m = static_cast<MapHolder*>(h);
    std::vector<std::string> fv;
    for (int i = 0; i < nfiles; ++i) fv.emplace_back(files[i]);
    FastqKmer fk(m->map, fv, m->k, threads);
    auto t0 = std::chrono::steady_clock::now();
    fk.build_fastq_index();
    auto t1 = std::chrono::steady_clock::now();
    if (read_bases) *read_bases = fk.mReadBase;
    if (c_out)
        for (size_t i = 0; i < m->keys.size(); ++i) c_out[i] = m->map.find(m->keys[i])->second.c;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
