// gbwt_synth_cuda.cu -- BENCHMARK INPUT GENERATOR on the device (not part of the product, not the oracle).
//
// Writes the query patterns of SURVEY.md 8(d) straight into HBM so that bench.py's device-timed region
// starts with its inputs resident, without staging 1 G x 32 x 8 bytes through the host. Bit-identical to
// synth_patterns() in gbwt_synth.c (tests/test_gpu_synth.py compares them).
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__device__ inline uint64_t fwd_node(uint64_t seed, uint64_t S, uint64_t h, uint64_t p) {
    const uint64_t s = p >> 1;
    if ((p & 1) == 0) return 2 * (3 * s + 1);
    return 2 * (3 * s + 2 + (mix64(seed + h * S + s) >> 63));
}

__device__ inline uint64_t seq_node(uint64_t seed, uint64_t S, uint64_t seq, uint64_t p) {
    const uint64_t h = seq >> 1;
    if ((seq & 1) == 0) return fwd_node(seed, S, h, p);
    return fwd_node(seed, S, h, 2 * S - p) ^ 1;
}

// One thread per pattern node: consecutive threads write consecutive words.
__global__ void k_patterns(uint64_t S, uint64_t H, uint64_t seed, uint64_t seed_q, uint64_t q0, uint64_t n, uint64_t k,
                           uint64_t* __restrict__ out) {
    const uint64_t span = 2 * S + 1 - (k - 1);
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n * k;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t q = q0 + i / k, j = i % k;
        const uint64_t h = mix64(seed_q + 3 * q) % H;
        const uint64_t o = mix64(seed_q + 3 * q + 1) & 1;
        const uint64_t t = mix64(seed_q + 3 * q + 2) % span;
        out[i] = seq_node(seed, S, 2 * h + o, t + j);
    }
}

}  // namespace

extern "C" int synth_patterns_device(uint64_t S, uint64_t H, uint64_t seed, uint64_t seed_q, uint64_t q0, uint64_t n,
                                     uint64_t k, uint64_t* d_out, void* stream) {
    if (n == 0 || k == 0) return 0;
    const uint64_t total = n * k;
    const unsigned blocks = static_cast<unsigned>(total / 256 + 1 < 148u * 64u ? total / 256 + 1 : 148u * 64u);
    k_patterns<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, H, seed, seed_q, q0, n, k, d_out);
    return static_cast<int>(cudaGetLastError());
}
