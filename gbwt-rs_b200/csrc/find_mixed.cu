// find_mixed.cu -- the lean find/extend loop (find_lean.cuh) for indexes that also hold run-length bodies or records
// of outdegree > 2: real pangenome indexes, where those records are a small minority. A step on one of them is the
// general edge lookup + run scan, in an out-of-line device function so that the loop keeps its registers.
#include "find_mixed.h"

#include "find_lean.cuh"

namespace gbwt_b200 {

__device__ __noinline__ void follow_other_record(const Unit16* bodies, const Edge* edges, uint32_t d0, uint32_t d1, uint32_t d2,
                                                 uint32_t d3, uint32_t d4, uint32_t d5, uint32_t d6, uint32_t d7, uint32_t next,
                                                 uint32_t* range) {
    Desc d;
    d.a.x = d0; d.a.y = d1; d.a.z = d2; d.a.w = d3; d.b.x = d4; d.b.y = d5; d.b.z = d6; d.b.w = d7;
    IndexView view;  // follow_body only looks at the bodies and the edge lists
    view.desc = nullptr; view.bodies = bodies; view.edges = edges; view.endmarker = nullptr;
    view.records = 0; view.offset = 0; view.alphabet_size = 0; view.sequences = 0; view.endmarker_len = 0;
    view.bidirectional = 0; view.skips = nullptr; view.edges_valid = 1;
    // follow_body for a record that is neither single-edge nor dense (the caller handles those): edge lookup, then the
    // run scan up to the end of the range. (The 256-bit inline-asm load of the dense path must not appear in an
    // out-of-line function: that is what crashes ptxas 12.9.)
    uint32_t start = range[0], end = range[1];
    range[0] = range[1] = 0;
    uint32_t rank = 0, edge_offset = 0;
    FlipSet fs;
    fs.lt = 0; fs.extra = NO_SYMBOL;
    if (!find_edge<false>(view, d, next, rank, edge_offset, fs)) return;
    const uint32_t total = d.total_len();
    if (start > total) start = total;
    if (end > total) end = total;
    Ranks r;
    r.at_start = r.at_end = r.flipped = 0;
    rank_runs_inline<false>(view, d, rank, fs, start, end, r);
    if (r.at_start >= r.at_end) return;
    range[0] = edge_offset + r.at_start; range[1] = edge_offset + r.at_end;
}

namespace {

__global__ void __launch_bounds__(BLOCK_THREADS, 5) k_find_extend_lean_mixed(IndexView ix, const uint64_t* __restrict__ patterns,
                                                                              const uint32_t* __restrict__ perm, size_t n, size_t k,
                                                                              gbwt_b200_state* __restrict__ out) {
    GBWT_FOR_EACH_QUERY(q, n, perm) {
        gbwt_b200_state st;
        ChunkReader rd(patterns + q * k, static_cast<uint32_t>(k));
        query_find_extend_lean<true>(ix, rd, static_cast<uint32_t>(k), st);
        store_state(out + q, st);
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS, 5) k_find_extend_ragged_lean_mixed(IndexView ix, const uint64_t* __restrict__ nodes,
                                                                                     const uint64_t* __restrict__ offsets, uint64_t base,
                                                                                     const uint32_t* __restrict__ perm, size_t n,
                                                                                     gbwt_b200_state* __restrict__ out) {
    GBWT_FOR_EACH_QUERY(q, n, perm) {
        const uint64_t lo = __ldg(offsets + q), hi = __ldg(offsets + q + 1);
        const uint64_t len = hi > lo ? hi - lo : 0;
        gbwt_b200_state st;
        set_none(st);
        if (len <= 0xFFFFFFFFull) {
            ChunkReader rd(nodes + (lo - base), static_cast<uint32_t>(len));
            query_find_extend_lean<true>(ix, rd, static_cast<uint32_t>(len), st);
        }
        store_state(out + q, st);
    }
}

}  // namespace

void launch_find_extend_lean_mixed(const IndexView& ix, const uint64_t* patterns, const uint32_t* perm, size_t n, size_t k,
                                   gbwt_b200_state* out, unsigned grid, cudaStream_t stream) {
    k_find_extend_lean_mixed<<<grid, BLOCK_THREADS, 0, stream>>>(ix, patterns, perm, n, k, out);
}

void launch_find_extend_ragged_lean_mixed(const IndexView& ix, const uint64_t* nodes, const uint64_t* offsets, uint64_t base,
                                          const uint32_t* perm, size_t n, gbwt_b200_state* out, unsigned grid, cudaStream_t stream) {
    k_find_extend_ragged_lean_mixed<<<grid, BLOCK_THREADS, 0, stream>>>(ix, nodes, offsets, base, perm, n, out);
}

}  // namespace gbwt_b200
