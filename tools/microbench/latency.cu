// Development microbenchmark: dependent-load latency on this GPU for (a) a pointer chase by one warp per SM over a
// working set of `mb` MiB (L2-resident or not), (b) the same with every warp of the grid chasing the SAME chain
// (the access pattern of haplotype walks that move through the index as a pack).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__global__ void chase(const uint32_t* __restrict__ next, uint32_t start_stride, int steps, uint32_t* out, long long* cycles) {
    uint32_t p = (blockIdx.x * start_stride) ;
    // every lane follows the same chain (broadcast loads), like the warp-mode walk
    long long t0 = clock64();
    for (int i = 0; i < steps; i++) p = __ldg(next + p);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = p; cycles[blockIdx.x] = t1 - t0; }
}

__global__ void chase_wide(const uint4* __restrict__ next, uint32_t start_stride, int steps, uint32_t* out, long long* cycles) {
    uint32_t p = (blockIdx.x * start_stride);
    long long t0 = clock64();
    for (int i = 0; i < steps; i++) {
        // five independent loads per step, the address of the next step depends on the last one issued
        uint4 a = __ldg(next + p), b = __ldg(next + (p ^ 1)), c = __ldg(next + (p ^ 2)), d = __ldg(next + (p ^ 3)), e = __ldg(next + (p ^ 4));
        p = (e.x + ((a.y ^ b.y ^ c.y ^ d.y) & 0)) ;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = p; cycles[blockIdx.x] = t1 - t0; }
}

// 256-bit loads (ld.global.nc.v8.u32) on 32-byte elements: one per step, or three in parallel plus two 128-bit ones
__device__ __forceinline__ void ld256(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
__global__ void chase256(const uint4* __restrict__ next, uint32_t start_stride, int steps, int mode, uint32_t* out, long long* cycles) {
    uint32_t p = (blockIdx.x * start_stride) & ~1u;
    long long t0 = clock64();
    for (int i = 0; i < steps; i++) {
        uint32_t a[8];
        ld256(next + p, a);
        if (mode == 0) { p = a[0] & ~1u; }
        else if (mode == 1) { p = a[4] & ~1u; }                       // consume the upper half first
        else {
            uint32_t b[8], c[8];
            ld256(next + (p ^ 2), b); ld256(next + (p ^ 4), c);
            uint4 d = __ldg(next + (p ^ 6)), e = __ldg(next + (p ^ 8));
            p = (a[0] + ((b[1] ^ c[1] ^ d.y ^ e.y) & 0)) & ~1u;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = p; cycles[blockIdx.x] = t1 - t0; }
}

int main(int argc, char** argv) {
    double mbf = argc > 1 ? atof(argv[1]) : 32; size_t mb = (size_t)mbf;
    int blocks = argc > 2 ? atoi(argv[2]) : 148;
    int same = argc > 3 ? atoi(argv[3]) : 0;   // 1: all blocks start at the same element
    int steps = 20000;
    size_t n = (size_t)(mbf * 1024 * 1024) / 16;
    std::vector<uint32_t> perm(n);
    for (size_t i = 0; i < n; i++) perm[i] = i;
    std::mt19937_64 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    std::vector<uint4> h(n);
    for (size_t i = 0; i < n; i++) { h[perm[i]] = make_uint4(perm[(i + 1) % n], 0, 0, 0); }
    uint4* d; cudaMalloc(&d, n * 16); cudaMemcpy(d, h.data(), n * 16, cudaMemcpyHostToDevice);
    uint32_t* out; long long* cyc; cudaMalloc(&out, blocks * 4); cudaMalloc(&cyc, blocks * 8);
    std::vector<long long> hc(blocks);
    for (int rep = 0; rep < 3; rep++) {
        chase_wide<<<blocks, 32>>>(d, same ? 0 : (uint32_t)(n / blocks), steps, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(hc.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto c : hc) avg += c; avg /= blocks;
        printf("mb=%zu blocks=%d same=%d rep=%d wide: %.1f cycles/step\n", mb, blocks, same, rep, avg / steps);
    }
    {
        // 32-byte elements: element 2j holds the index (even) of the next element in .x and in word 4
        std::vector<uint4> h2(n);
        size_t m = n / 2;
        std::vector<uint32_t> perm2(m);
        for (size_t i = 0; i < m; i++) perm2[i] = i;
        std::shuffle(perm2.begin(), perm2.end(), rng);
        for (size_t i = 0; i < m; i++) { uint32_t nx = 2 * perm2[(i + 1) % m]; h2[2 * perm2[i]] = make_uint4(nx, 0, 0, 0); h2[2 * perm2[i] + 1] = make_uint4(nx, 0, 0, 0); }
        cudaMemcpy(d, h2.data(), n * 16, cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 3; mode++) {
            for (int rep = 0; rep < 3; rep++) {
                chase256<<<blocks, 32>>>(d, same ? 0 : (uint32_t)(n / blocks), steps, mode, out, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(hc.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (auto c : hc) avg += c; avg /= blocks;
            printf("mb=%zu blocks=%d same=%d rep=2 v8 mode %d: %.1f cycles/step\n", mb, blocks, same, mode, avg / steps);
        }
    }
    // 4-byte chase over the same buffer viewed as u32 with stride 4 (only .x used)
    std::vector<uint32_t> h32(n * 4, 0);
    for (size_t i = 0; i < n; i++) h32[4 * i] = 4 * h[i].x;
    cudaMemcpy(d, h32.data(), n * 16, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; rep++) {
        chase<<<blocks, 32>>>((const uint32_t*)d, same ? 0 : (uint32_t)(4 * (n / blocks)), steps, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(hc.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto c : hc) avg += c; avg /= blocks;
        printf("mb=%zu blocks=%d same=%d rep=%d single: %.1f cycles/step\n", mb, blocks, same, rep, avg / steps);
    }
    return 0;
}
