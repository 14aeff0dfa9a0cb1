// Microbenchmark: two ways to give every key its rank within its bin (47 bins, ~half the lanes active),
// as the scatter kernel's binning does: (a) one shared-memory atomicAdd per key on CTA-wide counters,
// (b) warp-private counters, __match_any_sync groups, rank = counter + position in group, leader stores.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rank_bench rank_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rnd(uint64_t& x) {
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    return (uint32_t)((x * 0x2545F4914F6CDD1DULL) >> 32);
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) rank_kernel(int rounds, uint32_t P, unsigned long long* sink) {
    __shared__ uint32_t hist[8 * 64];
    for (int i = threadIdx.x; i < 8 * 64; i += 256) hist[i] = 0;
    __syncthreads();
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 777;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        const uint32_t v = rnd(x);
        const uint32_t p = __umulhi(v, P);
        const bool act = (v & 1u) != 0;  // ~half the lanes hold a surviving k-mer
        if (MODE == 0) {
            if (act) acc += atomicAdd(&hist[p], 1u);
        } else if (MODE == 1) {
            const uint32_t am = __ballot_sync(0xffffffffu, act);
            if (act) {
                const uint32_t m = __match_any_sync(am, p);
                uint32_t* h = &hist[warp * 64 + p];
                const uint32_t r0 = *h;
                const uint32_t pos = __popc(m & ((1u << lane) - 1u));
                __syncwarp(am);
                if (pos == 0) *h = r0 + __popc(m);
                acc += r0 + pos;
            }
        } else {  // MODE 2: the ballots / match alone, no shared memory
            const uint32_t am = __ballot_sync(0xffffffffu, act);
            if (act) acc += __popc(__match_any_sync(am, p) & ((1u << lane) - 1u));
        }
    }
    if (acc == 0x12345u) atomicAdd(sink, 1ull);
}

int main() {
    unsigned long long* sink; cudaMalloc(&sink, 8);
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * 4, rounds = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[3] = {"ATOMS per key (CTA-wide)", "match_any + warp-private counters", "ballot + match_any only"};
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) rank_kernel<0><<<grid, 256>>>(rounds, 47, sink);
            else if (mode == 1) rank_kernel<1><<<grid, 256>>>(rounds, 47, sink);
            else rank_kernel<2><<<grid, 256>>>(rounds, 47, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double steps = (double)grid * 8 * rounds;  // warp-steps
        printf("%-36s %.3f ms, %.2f cycles per warp-step per SM, %.1f G keys/s\n", names[mode], best,
               best * 1e-3 * 1.965e9 * nsm / steps, steps * 16 / (best * 1e-3) / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
