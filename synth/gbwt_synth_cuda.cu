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

// The allele model of gbwt_synth.c (model_t): thr == 0 is the legacy model (two equally likely alleles, 3 ids per site).
struct Model { uint64_t seed, S; uint32_t thr, tri_mod, stride; };

__device__ inline uint32_t allele(const Model& m, uint64_t h, uint64_t s) {
    const uint64_t r = mix64(m.seed + h * m.S + s);
    if (m.thr == 0) return static_cast<uint32_t>(r >> 63);
    const uint64_t u = r >> 32;
    if (u < m.thr) return 1;
    if (u < 2 * static_cast<uint64_t>(m.thr) && m.tri_mod != 0 && mix64(m.seed * 31 + 0x5BD1E995ULL + s) % m.tri_mod == 0) return 2;
    return 0;
}

__device__ inline uint64_t fwd_node(const Model& m, uint64_t h, uint64_t p) {
    const uint64_t s = p >> 1;
    if ((p & 1) == 0) return 2 * (m.stride * s + 1);
    return 2 * (m.stride * s + 2 + allele(m, h, s));
}

__device__ inline uint64_t seq_node(const Model& m, uint64_t seq, uint64_t p) {
    const uint64_t h = seq >> 1;
    if ((seq & 1) == 0) return fwd_node(m, h, p);
    return fwd_node(m, h, 2 * m.S - p) ^ 1;
}

// One thread per pattern node: consecutive threads write consecutive words.
__global__ void k_patterns(Model m, uint64_t H, uint64_t seed_q, uint64_t q0, uint64_t n, uint64_t k,
                           uint64_t* __restrict__ out) {
    const uint64_t span = 2 * m.S + 1 - (k - 1);
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n * k;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t q = q0 + i / k, j = i % k;
        const uint64_t h = mix64(seed_q + 3 * q) % H;
        const uint64_t o = mix64(seed_q + 3 * q + 1) & 1;
        const uint64_t t = mix64(seed_q + 3 * q + 2) % span;
        out[i] = seq_node(m, 2 * h + o, t + j);
    }
}

}  // namespace

extern "C" int synth_patterns_device_v(uint64_t S, uint64_t H, uint64_t seed, uint64_t alt_ppm, uint64_t tri_mod, uint64_t seed_q,
                                       uint64_t q0, uint64_t n, uint64_t k, uint64_t* d_out, void* stream) {
    if (n == 0 || k == 0) return 0;
    Model m;
    m.seed = seed; m.S = S;
    m.thr = alt_ppm == 0 ? 0u : static_cast<uint32_t>((alt_ppm * 4294967296.0) / 1e6);  // as make_model() in gbwt_synth.c
    if (alt_ppm != 0 && m.thr == 0) m.thr = 1;
    m.tri_mod = alt_ppm == 0 ? 0u : static_cast<uint32_t>(tri_mod);
    m.stride = alt_ppm == 0 ? 3u : 4u;
    const uint64_t total = n * k;
    const unsigned blocks = static_cast<unsigned>(total / 256 + 1 < 148u * 64u ? total / 256 + 1 : 148u * 64u);
    k_patterns<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(m, H, seed_q, q0, n, k, d_out);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int synth_patterns_device(uint64_t S, uint64_t H, uint64_t seed, uint64_t seed_q, uint64_t q0, uint64_t n,
                                     uint64_t k, uint64_t* d_out, void* stream) {
    return synth_patterns_device_v(S, H, seed, 0, 0, seed_q, q0, n, k, d_out, stream);
}
