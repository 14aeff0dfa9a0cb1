// Microbenchmark: random 4-byte / 16-byte gathers out of an L2-resident buffer through the LSU path
// (ld.global.nc) versus the texture path (tex1Dfetch), and random 32-byte bucket loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t rnd(uint64_t& x) {
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    return (uint32_t)((x * 0x2545F4914F6CDD1DULL) >> 32);
}

template <int MODE>  // 0: ldg u32, 1: tex u32, 2: ldg uint4, 3: tex uint4, 4: ldg 32 B (2 x uint4 adjacent), 5: tex 32 B, 6: 256-bit ld, 7: pairs
__global__ void __launch_bounds__(256) gather(const uint32_t* buf, cudaTextureObject_t t32, cudaTextureObject_t t128,
                                              uint32_t nwords, int rounds, unsigned long long* sink) {
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 12345;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        uint32_t v[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const uint32_t w = __umulhi(rnd(x), nwords);
            if (MODE == 0) v[b] = __ldg(buf + w);
            else if (MODE == 1) v[b] = tex1Dfetch<uint32_t>(t32, (int)w);
            else if (MODE == 2) { uint4 q = __ldg(reinterpret_cast<const uint4*>(buf) + (w >> 2)); v[b] = q.x ^ q.y ^ q.z ^ q.w; }
            else if (MODE == 3) { uint4 q = tex1Dfetch<uint4>(t128, (int)(w >> 2)); v[b] = q.x ^ q.y ^ q.z ^ q.w; }
            else if (MODE == 4) { const uint4* p = reinterpret_cast<const uint4*>(buf) + ((w >> 3) << 1); uint4 q = __ldg(p), s = __ldg(p + 1); v[b] = q.x ^ q.w ^ s.y ^ s.z; }
            else if (MODE == 5) { int i = (int)((w >> 3) << 1); uint4 q = tex1Dfetch<uint4>(t128, i), s = tex1Dfetch<uint4>(t128, i + 1); v[b] = q.x ^ q.w ^ s.y ^ s.z; }
            else if (MODE == 6) {  // one 256-bit load per lane, as the index probe does
                unsigned long long a0, a1, a2, a3;
                asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(buf + ((w >> 3) << 3)));
                v[b] = (uint32_t)(a0 ^ a1 ^ a2 ^ a3);
            } else {  // MODE 7: lane pairs share a 32-byte bucket, 16 bytes each
                const uint32_t wp = __shfl_sync(0xffffffffu, w, threadIdx.x & 30);
                uint4 q = __ldg(reinterpret_cast<const uint4*>(buf) + ((wp >> 3) << 1) + (threadIdx.x & 1));
                v[b] = q.x ^ q.y ^ q.z ^ q.w;
            }
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) acc ^= v[b];
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 30;
    const uint32_t nwords = (uint32_t)(mb << 20) / 4;
    uint32_t* buf; unsigned long long* sink;
    CK(cudaMalloc(&buf, (size_t)nwords * 4)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(buf, 0x5a, (size_t)nwords * 4)); CK(cudaMemset(sink, 0, 8));
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = buf; rd.res.linear.sizeInBytes = (size_t)nwords * 4;
    cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t32 = 0, t128 = 0;
    rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>(); CK(cudaCreateTextureObject(&t32, &rd, &td, nullptr));
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>(); CK(cudaCreateTextureObject(&t128, &rd, &td, nullptr));
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * 8, rounds = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[8] = {"ldg u32", "tex u32", "ldg 16B", "tex 16B", "ldg 32B (2x16)", "tex 32B (2x16)", "ld.v4.u64 32B",
                            "lane pairs 2x16B"};
    for (int mode = 0; mode < 8; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: gather<0><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
                case 1: gather<1><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
                case 2: gather<2><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
                case 3: gather<3><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
                case 4: gather<4><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
                case 5: gather<5><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
                case 6: gather<6><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
                default: gather<7><<<grid, 256>>>(buf, t32, t128, nwords, rounds, sink); break;
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double n = (double)grid * 256 * rounds * 8;
        printf("%-16s %3zu MB buffer: %7.1f G lane-gathers/s (%.3f ms)\n", names[mode], mb, n / (best * 1e-3) / 1e9, best);
    }
    return 0;
}
