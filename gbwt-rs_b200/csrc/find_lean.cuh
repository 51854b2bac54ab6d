// find_lean.cuh -- the lean find/extend loop and what it needs (pattern reader, work assignment), shared by kernels.cuh
// and find_mixed.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "record_scan.cuh"

namespace gbwt_b200 {

constexpr int BLOCK_THREADS = 256;

// Work assignment of the search kernels: item i of the batch, or, with a permutation (locality schedule), one
// contiguous span of the sorted batch per CTA, so that the queries a CTA runs one after the other come from the
// same and then the neighbouring buckets.
#define GBWT_FOR_EACH_QUERY(q, n, perm)                                                                              \
    const size_t _stride = (perm) != nullptr ? blockDim.x : static_cast<size_t>(gridDim.x) * blockDim.x;            \
    size_t _span = ((n) + gridDim.x - 1) / gridDim.x;                                                                \
    _span = (_span + blockDim.x - 1) / blockDim.x * blockDim.x;                                                      \
    const size_t _begin = (perm) != nullptr ? static_cast<size_t>(blockIdx.x) * _span : static_cast<size_t>(blockIdx.x) * blockDim.x; \
    const size_t _end = (perm) != nullptr ? (_begin + _span < (n) ? _begin + _span : (n)) : (n);                      \
    for (size_t _i = _begin + threadIdx.x, q = 0; _i < _end && ((q = (perm) != nullptr ? (perm)[_i] : _i), true); _i += _stride)

// 24-byte / 48-byte results are written with 8-byte stores; neighbouring threads write neighbouring
// records, so every warp store covers whole sectors.
__device__ __forceinline__ void store_state(gbwt_b200_state* out, const gbwt_b200_state& s) { *out = s; }

// Reads the pattern one 32-byte sector (four nodes) at a time with one 256-bit load, keeps the nodes as 32-bit
// values plus a "does not fit 32 bits" mask (see PlainReader), and always has the NEXT sector in flight: the
// pattern rows of a bucketed batch are scattered over HBM, so this is the one load of the loop that pays full
// DRAM latency, and nothing depends on it for four steps.
struct ChunkReader {
    const uint64_t* p;
    uint32_t k, base;
    uint32_t c0, c1, c2, c3, bad;
    uint64_t n0, n1, n2, n3;  // the sector after `base`, requested one chunk early
    bool vec;
    __device__ __forceinline__ void fetch(uint32_t b, uint64_t& v0, uint64_t& v1, uint64_t& v2, uint64_t& v3) const {
        v0 = v1 = v2 = v3 = 0;
        if (b >= k) return;
        if (vec && k - b >= 4) {
            asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v0), "=l"(v1), "=l"(v2), "=l"(v3) : "l"(p + b));
        } else {
            v0 = __ldg(p + b);
            if (b + 1 < k) v1 = __ldg(p + b + 1);
            if (b + 2 < k) v2 = __ldg(p + b + 2);
            if (b + 3 < k) v3 = __ldg(p + b + 3);
        }
    }
    __device__ __forceinline__ ChunkReader(const uint64_t* pattern, uint32_t len)
        : p(pattern), k(len), base(0xFFFFFFFFu), c0(0), c1(0), c2(0), c3(0), bad(0), vec((reinterpret_cast<uintptr_t>(pattern) & 31) == 0) {
        fetch(0, n0, n1, n2, n3);
    }
    // Nodes are read in increasing order of i (the search loop), so a new chunk is always the prefetched one.
    __device__ __forceinline__ bool node(uint32_t i, uint32_t& out) {
        const uint32_t b = i & ~3u;
        if (b != base) {
            base = b;
            c0 = static_cast<uint32_t>(n0); c1 = static_cast<uint32_t>(n1);
            c2 = static_cast<uint32_t>(n2); c3 = static_cast<uint32_t>(n3);
            bad = ((n0 >> 32) != 0 ? 1u : 0u) | ((n1 >> 32) != 0 ? 2u : 0u) | ((n2 >> 32) != 0 ? 4u : 0u) | ((n3 >> 32) != 0 ? 8u : 0u);
            fetch(b + 4, n0, n1, n2, n3);
        }
        const uint32_t j = i & 3u;
        out = j == 0 ? c0 : (j == 1 ? c1 : (j == 2 ? c2 : c3));
        return ((bad >> j) & 1u) == 0;
    }
};

// rank1 of position r < 192 inside one dense block and the bit at r (layout.h: {ones_before, c0 | c1 << 8, 192 bits}).
__device__ __forceinline__ uint32_t dense_block_rank_lean(const Quad& lo, const Quad& hi, uint32_t r, uint32_t& bit) {
    const uint32_t j = r >> 6, p = r & 63u;
    const uint32_t w_lo = j == 0 ? lo.z : (j == 1 ? hi.x : hi.z);
    const uint32_t w_hi = j == 0 ? lo.w : (j == 1 ? hi.y : hi.w);
    const uint64_t w = (static_cast<uint64_t>(w_hi) << 32) | w_lo;
    const uint32_t sub = ((lo.y << 8) >> (8 * j)) & 0xFFu;  // 0, c0, c1
    bit = static_cast<uint32_t>(w >> p) & 1u;
    return lo.x + sub + static_cast<uint32_t>(__popcll(w & ((1ull << p) - 1ull)));
}

// GBWT::find + extends (src/gbwt.rs:269-304 over Record::follow, src/bwt.rs:595-616) on an index with validated edge
// targets (IndexView::edges_valid) whose records are mostly SINGLE or DENSE2 (what the dense policy makes of a
// pangenome GBWT): same results as query_find_extend_rounds<true>, written for the instruction count. ncu put the general loop at 165 instructions
// per pattern node with the issue slots as the bound (53 % busy, 0.97 eligible warps per cycle); here a pattern node
// must EQUAL an edge target of the current record to go on, so it needs no range checks of its own, the descriptor
// is loaded without the empty-record preamble, and the two ranks of a dense step share one block whenever the
// range lies inside it. Same rounds as the general loop (single-edge records, then one record with a body), so the
// lanes of a warp meet at the rank step.
// A step on a record the lean loop does not handle itself (a run-length body, outdegree > 2): the general code, out of
// line so that it does not weigh on the registers of the loop. Defined in find_mixed.cu, the only translation unit
// that instantiates MIXED = true. Gets the two arrays it needs by value; range[0..1] = start, end on the way in and out, start >= end = None.
__device__ void follow_other_record(const Unit16* bodies, const Edge* edges, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3,
                                    uint32_t d4, uint32_t d5, uint32_t d6, uint32_t d7, uint32_t next, uint32_t* range);

// MIXED: the index also has other records (run-length bodies, outdegree > 2); a step on one of them is the general one.
template <bool MIXED, class Reader>
__device__ __forceinline__ void query_find_extend_lean(const IndexView& ix, Reader& rd, uint32_t k, gbwt_b200_state& out) {
    set_none(out);
    if (k == 0) return;
    const RecordDesc* const descs = ix.desc;
    const Unit16* const bodies = ix.bodies;
    const uint32_t base = static_cast<uint32_t>(ix.offset), records = static_cast<uint32_t>(ix.records);
    uint32_t x;
    // GBWT::find on pattern node 0 (src/gbwt.rs:269-281): a node of the alphabet with a non-empty record
    if (records == 0 || !rd.node(0, x) || x - base - 1u >= records - 1u) return;
    Desc d;
    load_sector(reinterpret_cast<const Unit16*>(descs + (x - base)), d.a, d.b);
    uint32_t start = 0, end = d.total_len(), node = x;
    if (end == 0) return;
    uint32_t i = 1;
    while (i < k) {
        // The single-edge loop has ONE exit (failures leave through a flag): the lanes of the warp reconverge behind
        // it and take the rank step together. With returns inside, the compiler merges the two loops and lanes that
        // started on different kinds of record stay out of phase for the whole pattern (14.8 of 32 lanes active).
        bool dead = false;
        uint32_t fmt = d.fmt();
        while (i < k && fmt == FMT_SINGLE) {
            // every position maps to edge 0 (follow_single); the endmarker (node 0) is below first_node
            if (!rd.node(i, x) || x != d.node0() || x == 0) { dead = true; break; }
            const uint32_t total = d.total_len();
            start = d.offset0() + (start < total ? start : total);
            end = d.offset0() + (end < total ? end : total);
            if (start >= end) { dead = true; break; }
            node = x;
            i++;
            if (i < k) {
                load_sector(reinterpret_cast<const Unit16*>(descs + (x - base)), d.a, d.b);
                fmt = d.fmt();
            }
        }
        if (dead) return;
        if (i >= k) break;
        if (fmt == FMT_EMPTY || !rd.node(i, x) || x == 0) return;  // EMPTY: BWT::record() is None
        if (fmt != FMT_DENSE2) {
            // the few records of a pangenome index that are neither single-edge nor dense (outdegree > 2, long runs)
            if constexpr (MIXED) {
                uint32_t range[2] = {start, end};
                follow_other_record(bodies, ix.edges, d.a.x, d.a.y, d.a.z, d.a.w, d.b.x, d.b.y, d.b.z, d.b.w, x, range);
                if (range[0] >= range[1]) return;
                start = range[0]; end = range[1];
                node = x;
                if (++i >= k) break;
                load_sector(reinterpret_cast<const Unit16*>(descs + (x - base)), d.a, d.b);
                continue;
            } else {
                return;
            }
        }
        uint32_t symbol, edge_offset;
        if (x == d.node0()) { symbol = 0; edge_offset = d.offset0(); }
        else if (x == d.node1()) { symbol = 1; edge_offset = d.offset1(); }
        else return;
        const uint32_t total = d.total_len();
        const uint32_t s = start < total ? start : total, e = end < total ? end : total;
        if (s >= e) return;
        // rank1(s) from the block of s; rank1(e) = rank1(e - 1) + bit(e - 1) from the block of e - 1 (e >= 1)
        const uint32_t blk_s = __umulhi(s, 0xAAAAAAABu) >> 7, blk_e = __umulhi(e - 1u, 0xAAAAAAABu) >> 7;
        const Unit16* body = bodies + d.body();
        Quad lo, hi;
        load_sector(body + 2u * blk_s, lo, hi);
        uint32_t bit;
        const uint32_t ones_s = dense_block_rank_lean(lo, hi, s - blk_s * DENSE_BITS, bit);
        uint32_t ones_e;
        if (blk_e == blk_s) {
            ones_e = dense_block_rank_lean(lo, hi, e - 1u - blk_e * DENSE_BITS, bit) + bit;
        } else {
            Quad lo2, hi2;
            load_sector(body + 2u * blk_e, lo2, hi2);
            ones_e = dense_block_rank_lean(lo2, hi2, e - 1u - blk_e * DENSE_BITS, bit) + bit;
        }
        const uint32_t rs = symbol ? ones_s : s - ones_s, re = symbol ? ones_e : e - ones_e;
        if (rs >= re) return;
        start = edge_offset + rs; end = edge_offset + re;
        node = x;
        if (++i >= k) break;
        load_sector(reinterpret_cast<const Unit16*>(descs + (x - base)), d.a, d.b);
    }
    out.node = node; out.start = start; out.end = end;
}

}  // namespace gbwt_b200
