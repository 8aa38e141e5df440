// Yardstick only (NOT part of libeast_b200.so): cub::DeviceRadixSort::SortPairs on the same problem shape as one
// round-0 sort of the global prefix-doubling build (BASELINE configs[2]: 1.83e8 pairs of 64-bit key + 32-bit value,
// 55 key bits), timed with CUDA events.  The product's own sort is k_rs_onesweep (csrc/radix_sort.cuh).
// build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o cub_sort_pairs cub_sort_pairs.cu
// run:    ./cub_sort_pairs [n = 182776850] [key bits = 55]
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__global__ void fill(uint64_t *k, uint32_t *v, size_t n, int bits) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9e3779b97f4a7c15ull;
        x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 32;
        k[i] = bits >= 64 ? x : (x & ((1ull << bits) - 1ull));
        v[i] = (uint32_t)i;
    }
}

int main(int argc, char **argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 182776850ull;
    const int bits = argc > 2 ? atoi(argv[2]) : 55;
    uint64_t *ka, *kb; uint32_t *va, *vb;
    cudaMalloc(&ka, n * 8); cudaMalloc(&kb, n * 8); cudaMalloc(&va, n * 4); cudaMalloc(&vb, n * 4);
    void *tmp = nullptr; size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint64_t> dk(ka, kb); cub::DoubleBuffer<uint32_t> dv(va, vb);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int64_t)n, 0, bits);
    cudaMalloc(&tmp, tmp_bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, sum = 0.f;
    const int reps = 5;
    for (int r = 0; r < reps + 1; ++r) {
        fill<<<148 * 8, 256>>>(ka, va, n, bits);
        cub::DoubleBuffer<uint64_t> k2(ka, kb); cub::DoubleBuffer<uint32_t> v2(va, vb);
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k2, v2, (int64_t)n, 0, bits);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) { best = ms < best ? ms : best; sum += ms; }
    }
    const int passes = (bits + 7) / 8;
    printf("{\"what\": \"cub::DeviceRadixSort::SortPairs u64 keys + u32 values\", \"n\": %zu, \"key_bits\": %d, \"ms_best\": %.3f, "
           "\"ms_mean\": %.3f, \"passes_of_8_bits\": %d, \"GBps_at_24B_per_pair_and_pass\": %.1f, \"cuda_error\": \"%s\"}\n",
           n, bits, best, sum / reps, passes, 24.0 * n * passes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
