// vg_gzip.h -- parallel inflate of a gzip file held in memory (see vg_gzip.cpp).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace vg {
namespace gz {

uint32_t crc32_fast(uint32_t crc, const uint8_t* buf, uint64_t len);  // == zlib's crc32

// a growable byte buffer that does not zero what it grows by (std::vector::resize does)
struct Buffer {
    uint8_t* data = nullptr;
    uint64_t size = 0, cap = 0;
    Buffer() = default;
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;
    ~Buffer() { free(data); }
    bool reserve(uint64_t n) {
        if (n <= cap) return true;
        const uint64_t ncap = n + n / 4 + 4096;
        uint8_t* p = (uint8_t*)realloc(data, ncap);
        if (!p) return false;
        data = p;
        cap = ncap;
        return true;
    }
};

// Working memory of the inflater that is worth keeping from one file to the next.
class Scratch {
   public:
    Scratch();
    ~Scratch();
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;

   private:
    friend class Stream;
    struct Pool;
    Pool* pool_;
};

class Stream {
   public:
    // data / size: the whole .gz file (it must stay mapped); chunk_bytes: compressed bytes per worker and round;
    // scratch (optional): working memory to reuse, not shared with another live Stream
    Stream(const uint8_t* data, uint64_t size, int threads, uint64_t chunk_bytes, Scratch* scratch = nullptr);
    ~Stream();
    Stream(const Stream&) = delete;
    Stream& operator=(const Stream&) = delete;
    // Inflates the next threads * per_thread chunks and appends their text to `out`.  false: error() says why; what
    // earlier rounds returned is good (it is exactly what zlib returns for those bytes).
    bool next(Buffer& out, int per_thread);
    bool eof() const;
    const std::string& error() const;

   private:
    struct Impl;
    Impl* impl_;
};

}  // namespace gz
}  // namespace vg
