// Measurement aid (not part of the product): the rate at which a B200 serves RANDOM small reads from HBM -- the
// access pattern of the index probes (key buckets, entry sides, filter blocks, unitig words) once the index is far
// larger than L2. Every thread issues `mlp` independent loads of `bytes` (32 or 64) at random aligned addresses of a
// `gb`-GB buffer per round. Reports accesses/s, the useful GB/s, and the DRAM GB/s if every access costs a 64-byte
// fetch (what ncu's dram__sectors_read shows for the mapping kernels: twice the sectors asked for).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_random tools/hbm_random.cu && ./hbm_random
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int BYTES, int MLP>
__global__ void k_random(const uint4* buf, uint64_t n_units, int rounds, uint64_t* sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0, s = mix(t + 1);
    for (int r = 0; r < rounds; ++r) {
        uint4 v[MLP * (BYTES / 16)];
#pragma unroll
        for (int j = 0; j < MLP; ++j) {
            s = mix(s + 0x9E3779B97F4A7C15ull);
            const uint64_t u = (uint64_t)(((unsigned __int128)s * n_units) >> 64);
            const uint4* p = buf + u * (BYTES / 16);
#pragma unroll
            for (int q = 0; q < BYTES / 16; ++q) v[j * (BYTES / 16) + q] = __ldg(p + q);
        }
#pragma unroll
        for (int j = 0; j < MLP * (BYTES / 16); ++j) acc += v[j].x ^ v[j].w;
    }
    if (acc == 0x1234567) *sink = acc;
}

template <int BYTES, int MLP>
void run(const uint4* buf, uint64_t bytes_total, uint64_t* sink, int warps_per_sm) {
    const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads, rounds = 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_random<BYTES, MLP><<<blocks, threads>>>(buf, bytes_total / BYTES, 4, sink);
    cudaEventRecord(a);
    k_random<BYTES, MLP><<<blocks, threads>>>(buf, bytes_total / BYTES, rounds, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double acc = (double)blocks * threads * rounds * MLP;
    printf("%2d B x mlp %d, %2d warps/SM: %7.2f G accesses/s, useful %6.0f GB/s, as 64-byte fetches %6.0f GB/s\n", BYTES, MLP,
           warps_per_sm, acc / ms / 1e6, acc * BYTES / ms / 1e6, acc * 64 / ms / 1e6);
}

// Second form: `hbm_random sweep` -- footprints from 8 MB to 12 GB under each L2 fetch granularity the driver
// accepts (cudaLimitMaxL2FetchGranularity 32 / 64 / 128): where random reads stop being L2 hits, and whether the
// granularity changes the rate HBM serves them at.
static void sweep() {
    uint4* buf;
    uint64_t* sink;
    if (cudaMalloc(&buf, 12ull << 30) != cudaSuccess) return;
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, 12ull << 30);
    for (size_t gran : {64, 32, 128}) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("L2 fetch granularity asked %zu (%s), in force %zu\n", gran, cudaGetErrorName(e), got);
        for (uint64_t mb : {8, 16, 24, 32, 48, 64, 96, 128, 256, 1024, 12288}) {
            printf("  footprint %5llu MB: ", (unsigned long long)mb);
            run<32, 4>(buf, mb << 20, sink, 32);
        }
        printf("  footprint 12288 MB, 64 B: ");
        run<64, 4>(buf, 12288ull << 20, sink, 32);
    }
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 's') {
        sweep();
        return 0;
    }
    const uint64_t gb = argc > 1 ? atoll(argv[1]) : 12;
    const uint64_t bytes_total = gb << 30;
    uint4* buf;
    uint64_t* sink;
    if (cudaMalloc(&buf, bytes_total) != cudaSuccess) return 1;
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes_total);
    printf("buffer %llu GB\n", (unsigned long long)gb);
    for (int w : {16, 32, 64}) {
        run<32, 1>(buf, bytes_total, sink, w);
        run<32, 4>(buf, bytes_total, sink, w);
        run<64, 1>(buf, bytes_total, sink, w);
        run<64, 4>(buf, bytes_total, sink, w);
    }
    return 0;
}
