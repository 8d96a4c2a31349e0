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

// Third form: `hbm_random mix` -- the seed scan's pattern at 3 Gbp: NS random 8-byte reads of a SMALL region (the
// filter) for every NB random 32-byte reads of a 12 GB region (the key buckets), with and without L2 eviction
// policies on either stream. If L2 kept the small region the rate would be near NS/128 + NB/35 ns per round.
__device__ __forceinline__ uint64_t ld8_hint(const void* p, uint64_t pol) {
    uint64_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld32_hint(const void* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
// HS / HB: 0 = plain __ldg, 1 = evict_last, 2 = evict_first, 3 = no_allocate
template <int NS, int NB, int HS, int HB, int WR = 0>
__global__ void k_mix(const uint64_t* small_, uint64_t n_small, const uint4* big, uint64_t n_big, int rounds, uint64_t* sink,
                      uint4* wr = nullptr, const uint4* rd = nullptr) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0, s = mix(t + 1);
    uint64_t pl, pf, pn;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pn));
    for (int r = 0; r < rounds; ++r) {
        uint64_t a[NS];
        uint4 b[NB + 1];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            s = mix(s + 0x9E3779B97F4A7C15ull);
            const uint64_t* p = small_ + (uint64_t)(((unsigned __int128)s * n_small) >> 64);
            a[j] = HS == 0 ? __ldg((const unsigned long long*)p) : ld8_hint(p, HS == 1 ? pl : (HS == 2 ? pf : pn));
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            s = mix(s + 0x9E3779B97F4A7C15ull);
            const uint4* p = big + 2 * (uint64_t)(((unsigned __int128)s * n_big) >> 64);
            b[j] = HB == 0 ? __ldg(p) : ld32_hint(p, HB == 1 ? pl : (HB == 2 ? pf : pn));
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) acc += a[j];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc += b[j].x ^ b[j].w;
        if (WR) {  // the scan's streams: 64 bytes written and 64 bytes read once per round and thread
            const uint64_t at = ((uint64_t)r * gridDim.x * blockDim.x + t) * 4;
            uint4 v = make_uint4((unsigned)acc, 1, 2, 3);
            if (WR & 2) {
                const uint4 x = __ldcs(rd + at), y = __ldcs(rd + at + 2);
                v.y ^= x.x ^ y.w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (WR & 4) __stcs(wr + at + q, v);
                else wr[at + q] = v;
            }
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
template <int NS, int NB, int HS, int HB, int WR>
static void run_mix_wr(const uint4* buf, uint64_t small_mb, uint64_t* sink) {
    const int threads = 256, blocks = 148 * 32 * 32 / threads, rounds = 64;
    const uint64_t* sm = (const uint64_t*)buf;
    const uint4* big = buf + (3ull << 30) / 16;  // [1 GB, 2 GB): written stream, [2 GB, 3 GB): read stream
    const uint64_t n_big = (9ull << 30) / 32, n_small = (small_mb << 20) / 8;
    uint4* wr = (uint4*)buf + (1ull << 30) / 16;
    const uint4* rd = buf + (2ull << 30) / 16;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_mix<NS, NB, HS, HB, WR><<<blocks, threads>>>(sm, n_small, big, n_big, 8, sink, wr, rd);
    cudaEventRecord(a);
    k_mix<NS, NB, HS, HB, WR><<<blocks, threads>>>(sm, n_small, big, n_big, rounds, sink, wr, rd);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double rdn = (double)blocks * threads * rounds;
    printf("small %3llu MB, %d small : %d big, streams %d (1 = 64 B written, 3 = and 64 B read per round): %6.2f G rounds/s\n",
           (unsigned long long)small_mb, NS, NB, WR, rdn / ms / 1e6);
}
template <int NS, int NB, int HS, int HB>
static void run_mix(const uint4* buf, uint64_t small_mb, uint64_t* sink) {
    const int threads = 256, blocks = 148 * 32 * 32 / threads, rounds = 64;
    const uint64_t* sm = (const uint64_t*)buf;  // the first small_mb MB; the big region is everything behind 1 GB
    const uint4* big = buf + (1ull << 30) / 16;
    const uint64_t n_big = (11ull << 30) / 32, n_small = (small_mb << 20) / 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_mix<NS, NB, HS, HB><<<blocks, threads>>>(sm, n_small, big, n_big, 8, sink);
    cudaEventRecord(a);
    k_mix<NS, NB, HS, HB><<<blocks, threads>>>(sm, n_small, big, n_big, rounds, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double rd = (double)blocks * threads * rounds;
    const double ideal = NS / 128.0 + NB / 33.0, none = (NS + NB) / 33.0;  // ns per round per 1e9
    printf("small %3llu MB, %d small : %d big, hints %d/%d: %6.2f G rounds/s (small resident: %5.2f, nothing resident: %5.2f)\n",
           (unsigned long long)small_mb, NS, NB, HS, HB, rd / ms / 1e6, 1.0 / ideal, 1.0 / none);
}
static void mixes() {
    uint4* buf;
    uint64_t* sink;
    if (cudaMalloc(&buf, 12ull << 30) != cudaSuccess) return;
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, 12ull << 30);
    for (uint64_t mb : {16, 32, 64}) {
        run_mix<4, 1, 0, 0>(buf, mb, sink);
        run_mix<4, 1, 1, 0>(buf, mb, sink);
        run_mix<4, 1, 0, 2>(buf, mb, sink);
        run_mix<4, 1, 1, 2>(buf, mb, sink);
        run_mix<4, 1, 1, 3>(buf, mb, sink);
        run_mix<4, 4, 0, 0>(buf, mb, sink);
        run_mix<4, 4, 1, 2>(buf, mb, sink);
        run_mix<4, 0, 0, 0>(buf, mb, sink);
        run_mix_wr<4, 1, 0, 0, 1>(buf, mb, sink);
        run_mix_wr<4, 1, 0, 0, 3>(buf, mb, sink);
        run_mix_wr<4, 0, 0, 0, 3>(buf, mb, sink);
        run_mix_wr<4, 1, 0, 0, 7>(buf, mb, sink);  // 7: the written stream as st.global.cs
    }
}

// Fourth form: `hbm_random v256` -- the same random 32-byte sector read as one 256-bit load (ld.global.nc.v4.u64,
// SASS LDG.E.ENL2.256.CONSTANT: what the key-bucket and entry-side loads use) against two 128-bit loads. Run it under
// `ncu --metrics dram__bytes_read.sum` to see what each form makes HBM deliver per access.
template <int FORM>
__global__ void k_v256(const uint4* buf, uint64_t n_units, int rounds, uint64_t* sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0, s = mix(t + 1);
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s = mix(s + 0x9E3779B97F4A7C15ull);
            const uint4* p = buf + 2 * (uint64_t)(((unsigned __int128)s * n_units) >> 64);
            if (FORM == 0) {
                const uint4 a = __ldg(p), b = __ldg(p + 1);
                acc += a.x ^ b.w;
            } else if (FORM == 1) {
                uint64_t a, b, c, d;
                asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
                acc += a ^ d;
            } else {
                uint64_t a, b, c, d;
                asm volatile("ld.global.nc.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
                acc += a ^ d;
            }
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
template <int FORM>
static void run_v256(const uint4* buf, uint64_t bytes_total, uint64_t* sink) {
    const int threads = 256, blocks = 148 * 32 * 32 / threads, rounds = 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k_v256<FORM><<<blocks, threads>>>(buf, bytes_total / 32, rounds, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double acc = (double)blocks * threads * rounds * 4;
    printf("form %d (%s): %.0f accesses, %7.2f G accesses/s\n", FORM,
           FORM == 0 ? "2 x 128-bit" : (FORM == 1 ? "256-bit" : "256-bit, L2::64B"), acc, acc / ms / 1e6);
}
static void v256() {
    uint4* buf;
    uint64_t* sink;
    if (cudaMalloc(&buf, 12ull << 30) != cudaSuccess) return;
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, 12ull << 30);
    for (int i = 0; i < 2; ++i) {
        run_v256<0>(buf, 12ull << 30, sink);
        run_v256<1>(buf, 12ull << 30, sink);
        run_v256<2>(buf, 12ull << 30, sink);
    }
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 'v') {
        v256();
        return 0;
    }
    if (argc > 1 && argv[1][0] == 's') {
        sweep();
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'm') {
        mixes();
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
