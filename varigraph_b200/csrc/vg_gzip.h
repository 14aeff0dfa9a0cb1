// vg_gzip.h -- parallel inflate of a gzip file held in memory (see vg_gzip.cpp).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace vg {
namespace gz {

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

class Stream {
   public:
    // data / size: the whole .gz file (it must stay mapped); chunk_bytes: compressed bytes per worker and round
    Stream(const uint8_t* data, uint64_t size, int threads, uint64_t chunk_bytes);
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
