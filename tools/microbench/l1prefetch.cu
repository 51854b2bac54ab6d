// Development microbenchmark: can a dependent load be turned into an L1 hit by prefetching its (known-in-advance)
// address `dist` steps earlier? One warp per block chases a random chain over `mb` MiB; the addresses of the chain
// are also available in order[] (read sequentially), the way a look-ahead over a deterministic skeleton would know
// them. mode 0: no prefetch; 1: prefetch.global.L1; 2: prefetch.global.L2; 3: plain load whose result is discarded.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__global__ void chase(const uint4* __restrict__ next, const uint32_t* __restrict__ order, int steps, int mode, int dist,
                      uint32_t* out, long long* cycles) {
    const uint32_t* ord = order + (size_t)blockIdx.x * (steps + 64);
    uint32_t p = ord[0], sink = 0;
    long long t0 = clock64();
    for (int i = 0; i < steps; i++) {
        const uint32_t ahead = __ldg(ord + i + dist);
        if (mode == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(next + ahead));
        else if (mode == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(next + ahead));
        else if (mode == 3) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(next + ahead)); sink ^= v & 0; }
        const uint4 v = __ldg(next + p);
        p = v.x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = p + sink; cycles[blockIdx.x] = t1 - t0; }
}

// Same chase, with the look-ahead done by a real load issued by lane 1 only (its result feeds a dummy that is
// consumed after the loop), dist steps ahead.
template <int DIST>
__global__ void chase_ld(const uint4* __restrict__ next, const uint32_t* __restrict__ order, int steps, uint32_t* out, long long* cycles) {
    const uint32_t* ord = order + (size_t)blockIdx.x * (steps + 64);
    uint32_t p = ord[0], acc = 0;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < steps; i++) {
        if (DIST > 0) {
            const uint32_t ahead = __ldg(ord + i + DIST);
            uint32_t v;
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(next + ahead));
            acc += v;
        }
        uint32_t x;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(next + p));
        p = x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = p + acc; cycles[blockIdx.x] = t1 - t0; }
}

// Look-ahead by real loads whose results are consumed 8 iterations later (rotating registers, unrolled by 8).
template <int DIST>
__global__ void chase_ring(const uint4* __restrict__ next, const uint32_t* __restrict__ order, int steps, uint32_t* out, long long* cycles) {
    const uint32_t* ord = order + (size_t)blockIdx.x * (steps + 64);
    uint32_t p = ord[0], acc = 0;
    uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < steps; i += 8) {
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            acc += v[j];
            const uint32_t ahead = __ldg(ord + i + j + DIST);
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v[j]) : "l"(next + ahead));
            uint32_t x;
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(next + p));
            p = x;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = p + acc; cycles[blockIdx.x] = t1 - t0; }
}

int main(int argc, char** argv) {
    size_t mb = argc > 1 ? atol(argv[1]) : 16;
    int blocks = argc > 2 ? atoi(argv[2]) : 148;
    int steps = 10000;
    size_t n = mb * 1024 * 1024 / 16;
    std::vector<uint32_t> perm(n);
    for (size_t i = 0; i < n; i++) perm[i] = i;
    std::mt19937_64 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    std::vector<uint4> h(n);
    for (size_t i = 0; i < n; i++) h[perm[i]] = make_uint4(perm[(i + 1) % n], 0, 0, 0);
    std::vector<uint32_t> order((size_t)blocks * (steps + 64));
    for (int b = 0; b < blocks; b++) {
        size_t at = (size_t)b * (n / blocks);
        for (int i = 0; i < steps + 64; i++) order[(size_t)b * (steps + 64) + i] = perm[(at + i) % n];
    }
    uint4* d; cudaMalloc(&d, n * 16); cudaMemcpy(d, h.data(), n * 16, cudaMemcpyHostToDevice);
    uint32_t* dord; cudaMalloc(&dord, order.size() * 4); cudaMemcpy(dord, order.data(), order.size() * 4, cudaMemcpyHostToDevice);
    uint32_t* out; long long* cyc; cudaMalloc(&out, blocks * 4); cudaMalloc(&cyc, blocks * 8);
    std::vector<long long> hc(blocks);
    for (int mode = 0; mode < 4; mode++) for (int dist : {1, 2, 4}) {
        for (int rep = 0; rep < 3; rep++) { chase<<<blocks, 32>>>(d, dord, steps, mode, dist, out, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(hc.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto c : hc) avg += c; avg /= blocks;
        printf("mb=%zu blocks=%d mode=%d dist=%d: %.1f cycles/step\n", mb, blocks, mode, dist, avg / steps);
    }
    for (int v = 0; v < 4; v++) {
        for (int rep = 0; rep < 3; rep++) {
            if (v == 0) chase_ld<0><<<blocks, 32>>>(d, dord, steps, out, cyc);
            if (v == 1) chase_ld<1><<<blocks, 32>>>(d, dord, steps, out, cyc);
            if (v == 2) chase_ld<2><<<blocks, 32>>>(d, dord, steps, out, cyc);
            if (v == 3) chase_ld<8><<<blocks, 32>>>(d, dord, steps, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(hc.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto c : hc) avg += c; avg /= blocks;
        printf("mb=%zu blocks=%d chase_ld variant %d (dist 0/1/2/8): %.1f cycles/step\n", mb, blocks, v, avg / steps);
    }
    for (int v = 0; v < 3; v++) {
        for (int rep = 0; rep < 3; rep++) {
            if (v == 0) chase_ring<1><<<blocks, 32>>>(d, dord, steps, out, cyc);
            if (v == 1) chase_ring<2><<<blocks, 32>>>(d, dord, steps, out, cyc);
            if (v == 2) chase_ring<4><<<blocks, 32>>>(d, dord, steps, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(hc.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto c : hc) avg += c; avg /= blocks;
        printf("mb=%zu blocks=%d chase_ring variant %d (dist 1/2/4): %.1f cycles/step\n", mb, blocks, v, avg / steps);
    }
    return 0;
}
