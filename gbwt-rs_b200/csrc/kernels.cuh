// kernels.cuh -- the CUDA kernels of the hot path (sm_100a). One thread owns one query (or one path) and
// runs the one-lane code of record_scan.cuh; thousands of independent queries per SM hide the HBM latency
// of the random record fetches. Grids are persistent: gridDim = SMs x resident CTAs, grid-stride loops.
//
//   K1  k_find / k_extend / k_find_extend(_ragged)   GBWT::find, extend          (src/gbwt.rs:269-304)
//   K2  k_bd_find / k_bd_extend / k_bd_search        bd_find, extend_forward/backward (src/gbwt.rs:311-384)
//   K3  k_start / k_forward / k_backward / k_sequence_lengths / k_extract   (src/gbwt.rs:213-261, 557-568)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "record_scan.cuh"

namespace gbwt_b200 {

constexpr int BLOCK_THREADS = 256;

#define GBWT_GRID_STRIDE(i, n) \
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < (n); i += static_cast<size_t>(gridDim.x) * blockDim.x)

// 24-byte / 48-byte results are written with 8-byte stores; neighbouring threads write neighbouring
// records, so every warp store covers whole sectors.
__device__ __forceinline__ void store_state(gbwt_b200_state* out, const gbwt_b200_state& s) { *out = s; }

__global__ void __launch_bounds__(BLOCK_THREADS) k_find(IndexView ix, const uint64_t* __restrict__ nodes, size_t n,
                                                         gbwt_b200_state* __restrict__ out) {
    GBWT_GRID_STRIDE(q, n) {
        gbwt_b200_state st;
        gbwt_find(ix, __ldg(nodes + q), st);
        store_state(out + q, st);
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_extend(IndexView ix, const gbwt_b200_state* states,
                                                           const uint64_t* __restrict__ nodes, size_t n,
                                                           gbwt_b200_state* out) {
    GBWT_GRID_STRIDE(q, n) {
        const gbwt_b200_state in = states[q];  // out may alias states
        gbwt_b200_state st;
        gbwt_extend(ix, in, __ldg(nodes + q), st);
        store_state(out + q, st);
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_find_extend(IndexView ix, const uint64_t* __restrict__ patterns,
                                                                size_t n, size_t k, gbwt_b200_state* __restrict__ out) {
    GBWT_GRID_STRIDE(q, n) {
        gbwt_b200_state st;
        query_find_extend(ix, patterns + q * k, k, st);
        store_state(out + q, st);
    }
}

// `base` = offsets[0] of the chunk that `nodes` starts at (host entry points upload node chunks).
__global__ void __launch_bounds__(BLOCK_THREADS) k_find_extend_ragged(IndexView ix, const uint64_t* __restrict__ nodes,
                                                                       const uint64_t* __restrict__ offsets, uint64_t base,
                                                                       size_t n, gbwt_b200_state* __restrict__ out) {
    GBWT_GRID_STRIDE(q, n) {
        const uint64_t lo = __ldg(offsets + q), hi = __ldg(offsets + q + 1);
        gbwt_b200_state st;
        query_find_extend(ix, nodes + (lo - base), hi > lo ? hi - lo : 0, st);
        store_state(out + q, st);
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_bd_find(IndexView ix, const uint64_t* __restrict__ nodes, size_t n,
                                                            gbwt_b200_bdstate* __restrict__ out) {
    GBWT_GRID_STRIDE(q, n) {
        gbwt_b200_bdstate st;
        gbwt_bd_find(ix, __ldg(nodes + q), st);
        out[q] = st;
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_bd_extend(IndexView ix, const gbwt_b200_bdstate* states,
                                                              const uint64_t* __restrict__ nodes, size_t n, int backward,
                                                              gbwt_b200_bdstate* out) {
    GBWT_GRID_STRIDE(q, n) {
        const gbwt_b200_bdstate in = states[q];
        gbwt_b200_bdstate st;
        if (backward) gbwt_extend_backward(ix, in, __ldg(nodes + q), st);
        else gbwt_extend_forward(ix, in, __ldg(nodes + q), st);
        out[q] = st;
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_bd_search(IndexView ix, const uint64_t* __restrict__ nodes,
                                                              const uint64_t* __restrict__ offsets, uint64_t base,
                                                              const uint64_t* __restrict__ first,
                                                              const uint64_t* __restrict__ start,
                                                              const uint64_t* __restrict__ end, size_t n,
                                                              gbwt_b200_bdstate* __restrict__ out) {
    GBWT_GRID_STRIDE(q, n) {
        const uint64_t lo = __ldg(offsets + q), hi = __ldg(offsets + q + 1);
        gbwt_b200_bdstate st;
        query_bd_search(ix, nodes + (lo - base), hi > lo ? hi - lo : 0, __ldg(first + q), __ldg(start + q), __ldg(end + q), st);
        out[q] = st;
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_start(IndexView ix, const uint64_t* __restrict__ ids, size_t n,
                                                          gbwt_b200_pos* __restrict__ out) {
    GBWT_GRID_STRIDE(q, n) {
        gbwt_b200_pos p;
        gbwt_start(ix, __ldg(ids + q), p);
        out[q] = p;
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_forward(IndexView ix, const gbwt_b200_pos* positions, size_t n,
                                                            gbwt_b200_pos* out) {
    GBWT_GRID_STRIDE(q, n) {
        const gbwt_b200_pos in = positions[q];
        gbwt_b200_pos p;
        gbwt_forward(ix, in, p);
        out[q] = p;
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_backward(IndexView ix, const gbwt_b200_pos* positions, size_t n,
                                                             gbwt_b200_pos* out) {
    GBWT_GRID_STRIDE(q, n) {
        const gbwt_b200_pos in = positions[q];
        gbwt_b200_pos p;
        gbwt_backward(ix, in, p);
        out[q] = p;
    }
}

// K3. A path walk is a dependent chain of LF steps (src/gbwt.rs:557-568): one thread per sequence, all
// sequences of the batch in flight at once. Paths of a pangenome move through the same records at about the
// same time, so after the first chain has pulled a record into L2 the others hit there.
// `nodes == nullptr` only counts (GBWT::sequence(id).count()).
__global__ void __launch_bounds__(64) k_extract(IndexView ix, const uint64_t* __restrict__ ids, size_t m,
                                                 const uint64_t* __restrict__ out_offsets, uint64_t base,
                                                 uint64_t* __restrict__ nodes, uint64_t* __restrict__ lengths) {
    GBWT_GRID_STRIDE(i, m) {
        uint64_t* dst = nullptr;
        uint64_t cap = 0;
        if (nodes != nullptr) {
            const uint64_t lo = __ldg(out_offsets + i), hi = __ldg(out_offsets + i + 1);
            dst = nodes + (lo - base);
            cap = hi > lo ? hi - lo : 0;
        }
        const uint64_t len = walk_sequence(ix, __ldg(ids + i), dst, cap);
        if (lengths != nullptr) lengths[i] = len;
    }
}

}  // namespace gbwt_b200
