/* vg_oracle.h -- CPU restatement of varigraph's read k-mer counting path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under varigraph_b200/ may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / reference legs use it, as the checker.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle.py)
 * against the unmodified reference compiled from /root/reference into
 * oracle/_ref/libvgref.so, and against the fixtures under tests/golden/ that
 * oracle/gen_golden.py produced by calling that library.
 */
#ifndef VG_ORACLE_H
#define VG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

uint8_t  vgo_nt4(uint8_t byte);
uint64_t vgo_hash64(uint64_t key, uint64_t mask);
int64_t  vgo_sketch(const char* seq, int64_t len, uint32_t k, uint64_t* out, int64_t cap);
int64_t  vgo_positions(const char* buf, int64_t nbytes, uint32_t k, uint64_t* out);

uint64_t vgo_murmur3_x64_128_sum(uint64_t key, uint32_t seed);
uint64_t vgo_cbf_size(uint64_t n, double p);
uint32_t vgo_cbf_num_hashes(uint64_t n, uint64_t m);
void     vgo_cbf_add(uint8_t* filter, uint64_t m, const uint64_t* seeds, uint32_t nh, uint64_t key);
uint8_t  vgo_cbf_count(const uint8_t* filter, uint64_t m, const uint64_t* seeds, uint32_t nh, uint64_t key);
int      vgo_cbf_find(const uint8_t* filter, uint64_t m, const uint64_t* seeds, uint32_t nh, uint64_t key);
uint64_t vgo_cbf_fill(uint8_t* filter, uint64_t m, const uint64_t* seeds, uint32_t nh, const char* seq,
                      int64_t len, uint32_t k);

void*    vgo_index_create(const uint64_t* keys, uint64_t n);
void     vgo_index_destroy(void* idx);
int64_t  vgo_index_find(const void* idx, uint64_t key);
uint64_t vgo_count_seq(const void* idx, const char* seq, int64_t len, uint32_t k, uint8_t* counts,
                       uint64_t* hits);
uint64_t vgo_count_lines(const void* idx, const char* buf, int64_t nbytes, uint32_t k, uint8_t* counts,
                         uint64_t* hits, uint64_t* nreads);

int64_t  vgo_fastq_to_lines(const char* text, int64_t n, char* out, int64_t cap, uint64_t* nreads,
                            uint64_t* read_bases, int* status);

#ifdef __cplusplus
}
#endif
#endif
