// Development microbenchmark for the one-pass placement of the window search (find_window.cu: k_window_place_direct):
// what bounds "read the first node of every pattern row, take a slot of its window with a returning atomic, store the
// query index there" -- the isolated sector reads, the atomics, or the latency of the dependent chain.
// Usage: place [log2 queries, default 26] [bytes per node, 8 or 4]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}

template <class T>
__global__ void k_init(T* patterns, size_t n, size_t k, uint32_t records) {
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x)
        patterns[q * k] = (T)(2 + mix64(q) % records);
}

// A: the shipped kernel's loop
template <class T>
__global__ void __launch_bounds__(256) k_place(const T* __restrict__ patterns, size_t n, size_t k, uint32_t wshift, uint32_t cap,
                                               uint32_t* __restrict__ filled, uint32_t* __restrict__ slots, uint32_t* __restrict__ deferred, uint32_t* counters) {
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const uint64_t node = __ldg(patterns + q * k);
        const uint32_t b = (uint32_t)((node - 2) >> wshift);
        const uint32_t at = atomicAdd(filled + b, 1u);
        if (at < cap) slots[(size_t)b * cap + at] = (uint32_t)q;
        else deferred[atomicAdd(&counters[1], 1u)] = (uint32_t)q;
    }
}

// B: the loads alone
template <class T>
__global__ void __launch_bounds__(256) k_loads(const T* __restrict__ patterns, size_t n, size_t k, uint32_t* out) {
    uint64_t acc = 0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) acc ^= __ldg(patterns + q * k);
    if (acc == 0x123456789ull) out[0] = 1;
}

// C: the atomics and the stores alone (window from a hash of the query index)
__global__ void __launch_bounds__(256) k_atomics(size_t n, uint32_t windows, uint32_t cap, uint32_t* __restrict__ filled, uint32_t* __restrict__ slots, int store) {
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(mix64(q) % windows);
        const uint32_t at = atomicAdd(filled + b, 1u);
        if (store && at < cap) slots[(size_t)b * cap + at] = (uint32_t)q;
        if (!store && at == 0xFFFFFFFFu) slots[0] = 1;
    }
}

// D: U queries per thread: all loads, then all atomics, then all stores
template <class T, int U>
__global__ void __launch_bounds__(256) k_place_unrolled(const T* __restrict__ patterns, size_t n, size_t k, uint32_t wshift, uint32_t cap,
                                                        uint32_t* __restrict__ filled, uint32_t* __restrict__ slots, uint32_t* __restrict__ deferred, uint32_t* counters) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q0 < n; q0 += stride * U) {
        uint32_t b[U], at[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t q = q0 + u * stride;
            b[u] = q < n ? (uint32_t)((__ldg(patterns + q * k) - 2) >> wshift) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int u = 0; u < U; u++) at[u] = b[u] != 0xFFFFFFFFu ? atomicAdd(filled + b[u], 1u) : 0u;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t q = q0 + u * stride;
            if (b[u] == 0xFFFFFFFFu) continue;
            if (at[u] < cap) slots[(size_t)b[u] * cap + at[u]] = (uint32_t)q;
            else deferred[atomicAdd(&counters[1], 1u)] = (uint32_t)q;
        }
    }
}

// E: a warp takes 32 consecutive rows; the aggregated variant merges the atomics of lanes with the same window (rare for
// random batches, common for crowded ones)
template <class T>
__global__ void __launch_bounds__(256) k_place_match(const T* __restrict__ patterns, size_t n, size_t k, uint32_t wshift, uint32_t cap,
                                                     uint32_t* __restrict__ filled, uint32_t* __restrict__ slots, uint32_t* __restrict__ deferred, uint32_t* counters) {
    const uint32_t lane = threadIdx.x & 31u;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < ((n + 31) & ~(size_t)31); q += (size_t)gridDim.x * blockDim.x) {
        const bool live = q < n;
        const uint32_t b = live ? (uint32_t)((__ldg(patterns + q * k) - 2) >> wshift) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, b);
        const uint32_t leader = __ffs(peers) - 1u, rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t at = 0;
        if (live && lane == leader) at = atomicAdd(filled + b, (uint32_t)__popc(peers));
        at = __shfl_sync(0xFFFFFFFFu, at, leader) + rank;
        if (!live) continue;
        if (at < cap) slots[(size_t)b * cap + at] = (uint32_t)q;
        else deferred[atomicAdd(&counters[1], 1u)] = (uint32_t)q;
    }
}

template <class T>
int run(int lg) {
    const size_t n = (size_t)1 << lg, k = 32;
    const uint32_t records = 20000001u, wshift = 9, windows = ((records - 1) >> wshift) + 1;
    const uint32_t cap = (uint32_t)(2 * ((n + windows - 1) / windows));
    T* patterns; uint32_t *filled, *slots, *deferred, *counters;
    CK(cudaMalloc(&patterns, n * k * sizeof(T)));
    CK(cudaMalloc(&filled, windows * 4)); CK(cudaMalloc(&slots, (size_t)cap * windows * 4)); CK(cudaMalloc(&deferred, n * 4)); CK(cudaMalloc(&counters, 8));
    CK(cudaMemset(patterns, 0, n * k * sizeof(T)));
    k_init<T><<<148 * 8, 256>>>(patterns, n, k, records);
    CK(cudaDeviceSynchronize());
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    auto timed = [&](const char* name, auto launch) {
        float best = 1e9f, sum = 0;
        for (int it = 0; it < 6; it++) {
            cudaMemset(filled, 0, windows * 4); cudaMemset(counters, 0, 8);
            cudaEventRecord(t0); launch(); cudaEventRecord(t1); cudaEventSynchronize(t1);
            float ms; cudaEventElapsedTime(&ms, t0, t1);
            if (it > 0) { best = ms < best ? ms : best; sum += ms; }
        }
        uint32_t c[2]; cudaMemcpy(c, counters, 8, cudaMemcpyDeviceToHost);
        printf("%-44s nodes of %zu bytes  best %.3f ms  mean %.3f ms  deferred %u  (%s)\n", name, sizeof(T), best, sum / 5, c[1], cudaGetErrorString(cudaGetLastError()));
    };
    for (int ctas_per_sm : {8, 4, 16}) {
        const unsigned grid = 148 * ctas_per_sm;
        char name[96];
        snprintf(name, sizeof name, "A shipped loop, %d CTAs/SM", ctas_per_sm);
        timed(name, [&] { k_place<T><<<grid, 256>>>(patterns, n, k, wshift, cap, filled, slots, deferred, counters); });
    }
    const unsigned grid = 148 * 8;
    timed("B loads alone", [&] { k_loads<T><<<grid, 256>>>(patterns, n, k, counters); });
    timed("C atomics alone", [&] { k_atomics<<<grid, 256>>>(n, windows, cap, filled, slots, 0); });
    timed("C atomics + stores", [&] { k_atomics<<<grid, 256>>>(n, windows, cap, filled, slots, 1); });
    timed("D 2 queries per thread", [&] { k_place_unrolled<T, 2><<<grid, 256>>>(patterns, n, k, wshift, cap, filled, slots, deferred, counters); });
    timed("D 4 queries per thread", [&] { k_place_unrolled<T, 4><<<grid, 256>>>(patterns, n, k, wshift, cap, filled, slots, deferred, counters); });
    timed("D 8 queries per thread", [&] { k_place_unrolled<T, 8><<<grid, 256>>>(patterns, n, k, wshift, cap, filled, slots, deferred, counters); });
    timed("D 4 queries per thread, 4 CTAs/SM", [&] { k_place_unrolled<T, 4><<<148 * 4, 256>>>(patterns, n, k, wshift, cap, filled, slots, deferred, counters); });
    timed("E match_any aggregation", [&] { k_place_match<T><<<grid, 256>>>(patterns, n, k, wshift, cap, filled, slots, deferred, counters); });
    cudaFree(patterns); cudaFree(filled); cudaFree(slots); cudaFree(deferred); cudaFree(counters);
    return 0;
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 26;
    const int bytes = argc > 2 ? atoi(argv[2]) : 8;
    return bytes == 4 ? run<uint32_t>(lg) : run<uint64_t>(lg);
}
