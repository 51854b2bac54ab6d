// find_window.cu -- GBWT::find + extends (src/gbwt.rs:269-304 over Record::follow, src/bwt.rs:595-616) with the
// records in shared memory.
//
// After the locality sort the queries of one bucket all start inside one window of consecutive records, and a
// pattern stays close to its first record (node identifiers follow the graph's topological order). So a CTA takes a
// window, stages the window plus a margin either side -- descriptors, two-hop shortcuts and the contiguous range of
// bodies, three cp.async.bulk copies completing on one mbarrier -- and its threads resolve the window's queries from
// shared memory: a dependent step costs a shared-memory round trip (~30 cycles) instead of an L2 / HBM one
// (300 / 800), the index is read from HBM once per batch in large sequential pieces, and the only scattered global
// accesses left are each query's own pattern row and its 24-byte result.
//
// Per step: the record u of the current node, the pattern node x1 (which must equal one of u's edge targets), and --
// the two-hop shortcut of layout.h -- if the successor over that edge is a single-edge record whose only target is the
// pattern node after x1, both nodes are consumed at once without looking at the successor's record. A bubble
// (SNP / indel site) is therefore ONE step: descriptor + shortcut + one or two 64-bit words of the bitvector.
//
// Anything the window cannot answer exactly -- a record outside the staged range, a body that did not fit, a record
// that is neither single-edge nor dense -- puts the query on a deferred list, and the general kernel finishes the
// list afterwards. Results are therefore always those of the general kernel; the window only makes the common case fast.
#include "find_window.h"

#include <algorithm>
#include <cstdlib>

#include "find_lean.cuh"

namespace gbwt_b200 {

namespace {

// ---- shared-memory accessors (32-bit shared-space addresses: no generic-pointer arithmetic in the loop) --------------

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }

// ---- pattern rows -----------------------------------------------------------------------------------------------------
// A warp takes 32 queries at a time and reads their pattern rows TOGETHER, one segment of 16 nodes per round: the lanes
// that share a row read consecutive 32-byte sectors of it, so a load instruction moves whole 128-byte lines (64-bit
// nodes: 4 lanes per row, 8 rows per instruction, 4 instructions per segment) instead of 32 isolated sectors. (The
// first versions had every thread read its own row sector by sector: one L1 wavefront per 32 bytes, 3.9 M of the
// 12.2 M data-stage wavefronts per SM, and the consumer waited for DRAM eight times per query -- 21 % of the stall
// samples in profiles/r2_find_window_v2_ncu.txt.) Each node is narrowed on the spot to a 16-bit index into the staged
// window (NODE_OUTSIDE if it is not in there) and stored in the warp's tile, node-major: tile[t][row], so that in the
// search loop a lane reads node t of its own row with one conflict-free 16-bit load however t moves.
constexpr uint32_t SEGMENT = 16;            // pattern nodes per round
constexpr uint32_t TILE_PITCH = 34;         // 16-bit slots per tile line: 32 rows + 2 (the lanes that share a row store 4 lines apart: 16 distinct banks)
constexpr uint32_t TILE_LINE = TILE_PITCH * 2u;  // bytes from node t to node t + 1 of a row
constexpr uint32_t TILE_BYTES = (SEGMENT + 1u) * TILE_LINE;  // one line more than a segment: the line after the last node always reads NODE_OUTSIDE
constexpr uint32_t NODE_OUTSIDE = 0xFFFFu;  // not a record of the staged window (or node 0, the endmarker)

// Window index of an edge target (node 0, the endmarker, is never a record to go on with).
__device__ __forceinline__ uint32_t window_index(uint32_t node, uint32_t origin, uint32_t count) {
    const uint32_t rel = node - origin;
    return node != 0 && rel < count ? rel : NODE_OUTSIDE;
}
// Window index of a pattern node. No test for node 0: if record 0 is staged at all (origin == 0) it is the endmarker's,
// which window_find defers and which no table entry names as a target, so a pattern node 0 never matches anything.
__device__ __forceinline__ uint32_t pattern_index(uint64_t node, uint32_t origin, uint32_t count) {
    const uint64_t rel = node - origin;
    return rel < count ? static_cast<uint32_t>(rel) : NODE_OUTSIDE;
}
__device__ __forceinline__ uint32_t pattern_index(uint32_t node, uint32_t origin, uint32_t count) {
    const uint32_t rel = node - origin;
    return rel < count ? rel : NODE_OUTSIDE;
}

// (16-bit shared accesses on 32-bit registers: the load zero-extends, the store writes the low half)
__device__ __forceinline__ void sts16(uint32_t a, uint32_t x) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// One 32-byte sector of pattern nodes, bypassing L1 (every byte is used once).
__device__ __forceinline__ void load_nodes(const uint64_t* p, uint64_t (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
__device__ __forceinline__ void load_nodes(const uint32_t* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}

// Fills the warp's tile with nodes [seg, seg + n_seg) of its 32 rows. `row` is this lane's own row (nullptr: no query).
// Sector loads when every row segment starts on a 32-byte boundary and n_seg is a whole number of sectors, else node
// by node.
template <class T>
__device__ __forceinline__ void load_tile(const T* row, uint32_t seg, uint32_t n_seg, bool sectors, uint32_t tile, uint32_t origin,
                                          uint32_t count, uint32_t lane) {
    constexpr uint32_t PER_SECTOR = 32u / sizeof(T);          // nodes per 32-byte sector: 4 or 8
    constexpr uint32_t MAX_LANES_PER_ROW = SEGMENT / PER_SECTOR;  // 4 or 2
    const unsigned long long mine = reinterpret_cast<unsigned long long>(row);
    if (sectors) {
        const uint32_t per_row = n_seg / PER_SECTOR;  // sectors per row segment: 1 .. MAX_LANES_PER_ROW (whole lanes only when a power of two)
        // lanes_per_row = MAX_LANES_PER_ROW always (a shorter segment leaves the extra lanes idle: only the last round of a pattern)
        const uint32_t sub = lane % MAX_LANES_PER_ROW, first = lane / MAX_LANES_PER_ROW;
        constexpr uint32_t ROWS_PER_LOAD = 32u / MAX_LANES_PER_ROW;  // 8 or 16
        T v[32u / ROWS_PER_LOAD][PER_SECTOR];
        bool have[32u / ROWS_PER_LOAD];
#pragma unroll
        for (uint32_t j = 0; j < 32u / ROWS_PER_LOAD; j++) {
            const uint32_t r = j * ROWS_PER_LOAD + first;
            const T* p = reinterpret_cast<const T*>(__shfl_sync(0xFFFFFFFFu, mine, r));
            have[j] = p != nullptr && sub < per_row;
            if (have[j]) load_nodes(p + seg + sub * PER_SECTOR, v[j]);
        }
#pragma unroll
        for (uint32_t j = 0; j < 32u / ROWS_PER_LOAD; j++) {
            const uint32_t r = j * ROWS_PER_LOAD + first;
            if (have[j]) {
#pragma unroll
                for (uint32_t t = 0; t < PER_SECTOR; t++)
                    sts16(tile + ((sub * PER_SECTOR + t) * TILE_PITCH + r) * 2u, pattern_index(v[j][t], origin, count));
            }
        }
    } else {
        for (uint32_t e = lane; e < 32u * n_seg; e += 32u) {
            const uint32_t r = e / n_seg, t = e - r * n_seg;
            const T* p = reinterpret_cast<const T*>(__shfl_sync(0xFFFFFFFFu, mine, r));
            if (p != nullptr) sts16(tile + (t * TILE_PITCH + r) * 2u, pattern_index(__ldg(p + seg + t), origin, count));
        }
    }
}

// Asks L2 for a pattern row ahead of its use: its first two 128-byte lines (a length-32 pattern is two lines of 64-bit
// nodes, one of 32-bit nodes).
template <class T>
__device__ __forceinline__ void prefetch_row(const T* row, uint32_t k) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
    if (k * sizeof(T) > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(row) + 128));
}

// First node of pattern row q as a 64-bit value (the sort key is computed from it).
__device__ __forceinline__ uint64_t first_node(const uint64_t* patterns, size_t q, size_t k) { return __ldg(patterns + q * k); }
__device__ __forceinline__ uint64_t first_node(const uint32_t* patterns, size_t q, size_t k) { return __ldg(patterns + q * k); }

// ---- the staged window -----------------------------------------------------------------------------------------------
// What a CTA decodes into shared memory for the records [lo, lo + count) -- once per window, used by ~3 queries per
// record times 16 steps. Arranged for the loop, not like the global layout, and in 16-bit fields: the data stage of
// L1 was the busiest unit of the first versions (78 % with 32-bit patterns), half of its shared-memory wavefronts
// being bank conflicts of 8- and 16-byte loads at random records, so every table entry is as narrow as it can be.
// Nodes are indexes into the window (NODE_OUTSIDE when the node is not staged); a record with a field that does not
// fit 16 bits is marked KIND_DEFER.
//   rec[r]   16 B  {target0 | target1 << 16, total_len | kind << 16, shortcut0, shortcut1}: kind = index of the record's
//                  first rank entry for a dense record, or KIND_SINGLE / KIND_EMPTY / KIND_DEFER (run-length body,
//                  outdegree > 2, body not staged); shortcut = the two-hop shortcut of the edge {landing node | offset << 16}
//                  (layout.h, IndexView::skips). One 16-byte load per step.
//   offs[r]   4 B  {offset0 | offset1 << 16}: only read when a step cannot take the shortcut
//   ranks[]   4 B  {ones before | 16 bits << 16}: rank at position p of a record is ONE 4-byte load at kind + p / 16
constexpr uint32_t KIND_SINGLE = 0xFFFFu, KIND_EMPTY = 0xFFFEu, KIND_DEFER = 0xFFFDu;  // anything below: dense
constexpr uint32_t RECORD_BYTES = 20;  // rec + offs

struct Staged {
    uint32_t rec, offs, ranks;  // shared-space addresses
    uint32_t lo, count;         // staged records [lo, lo + count)
};

enum : uint32_t { QUERY_ACTIVE = 0, QUERY_FOUND = 1, QUERY_NONE = 2, QUERY_DEFER = 3 };

// The state of one query between rounds.
struct WindowQuery {
    uint32_t idx;         // window index of the current node
    uint32_t start, end;  // the range in its record
    uint32_t i;           // next pattern position
    uint32_t status;
};

// GBWT::find on the first pattern node (window index x, NODE_OUTSIDE if it is not staged).
__device__ __forceinline__ void window_find(const Staged& st, uint32_t x, uint32_t k, WindowQuery& q) {
    q.idx = x; q.start = 0; q.end = 0; q.i = 1; q.status = QUERY_ACTIVE;
    if (x == NODE_OUTSIDE || x + st.lo == 0) { q.status = QUERY_DEFER; return; }  // (record 0 is the endmarker: find() is None; let the general code say so)
    const uint2 h = lds64(st.rec + 16u * x);
    const uint32_t kind = h.y >> 16;
    q.end = h.y & 0xFFFFu;
    if (kind == KIND_DEFER) q.status = QUERY_DEFER;
    else if (kind == KIND_EMPTY || q.end == 0) q.status = QUERY_NONE;  // GBWT::find: no record
    else if (k == 1) q.status = QUERY_FOUND;
}

// Extends an active query through the pattern nodes [q.i, seg + n_seg) that the warp's tile holds (tile line t = node
// seg + t of every row, the line after the last node reads NODE_OUTSIDE; `slot` = this lane's column). QUERY_FOUND:
// (idx, start, end) is the reference's SearchState; QUERY_NONE: the reference returns None; QUERY_DEFER: the window
// could not decide and the general kernel redoes the query from the start. The loop has ONE exit (every failure
// breaks out with its status), so the lanes of a warp reconverge after every step instead of carrying a stack of
// divergent returns. Per step: the two pattern nodes, one 16-byte record entry, one or two rank entries.
__device__ __forceinline__ void window_extend(const Staged& st, uint32_t slot, uint32_t seg, uint32_t n_seg, uint32_t k, WindowQuery& q) {
    uint32_t idx = q.idx, start = q.start, end = q.end, status = QUERY_ACTIVE;
    uint32_t pat = slot + (q.i - seg) * TILE_LINE;
    const uint32_t pat_end = slot + n_seg * TILE_LINE;
    uint4 h = lds128(st.rec + 16u * idx);
    while (pat < pat_end) {
        const uint32_t x1 = lds16(pat), x2 = lds16(pat + TILE_LINE);
        const uint32_t total = h.y & 0xFFFFu, kind = h.y >> 16;
        const uint32_t s = start < total ? start : total, e = end < total ? end : total;
        if (x1 == NODE_OUTSIDE) { status = QUERY_DEFER; break; }
        if (s >= e) { status = QUERY_NONE; break; }  // Record::follow on an empty range
        const uint32_t b = x1 == (h.x >> 16) ? 1u : 0u;
        uint32_t rs = s, re = e;
        if (kind < KIND_DEFER) {
            // dense record: rank1(s), and rank1(e) = ones up to and including position e - 1
            if (x1 != (h.x & 0xFFFFu) && b == 0) { status = QUERY_NONE; break; }
            const uint32_t last = e - 1u;
            const uint32_t ws = lds32(st.ranks + 4u * (kind + (s >> 4))), we = lds32(st.ranks + 4u * (kind + (last >> 4)));
            const uint32_t ones_s = (ws & 0xFFFFu) + static_cast<uint32_t>(__popc((ws >> 16) & ~(0xFFFFFFFFu << (s & 15u))));
            const uint32_t ones_e = (we & 0xFFFFu) + static_cast<uint32_t>(__popc((we >> 16) & ~(0xFFFFFFFEu << (last & 15u))));
            rs = b ? ones_s : s - ones_s;
            re = b ? ones_e : e - ones_e;
            if (rs >= re) { status = QUERY_NONE; break; }
        } else if (kind == KIND_SINGLE) {
            if (x1 != (h.x & 0xFFFFu)) { status = QUERY_NONE; break; }
        } else {
            status = kind == KIND_EMPTY ? QUERY_NONE : QUERY_DEFER;  // BWT::record() is None / a record the window does not decode
            break;
        }
        // two hops at once when the successor is a single-edge record leading to the pattern node after x1
        const uint32_t hop = b ? h.w : h.z;
        uint32_t offset;
        if (x2 == (hop & 0xFFFFu) && x2 != NODE_OUTSIDE) {
            offset = hop >> 16; idx = x2; pat += 2u * TILE_LINE;
        } else {
            offset = lds16(st.offs + 4u * idx + 2u * b); idx = x1; pat += TILE_LINE;
        }
        start = offset + rs; end = offset + re;
        if (pat >= pat_end) break;
        h = lds128(st.rec + 16u * idx);
    }
    q.idx = idx; q.start = start; q.end = end; q.i = seg + (pat - slot) / TILE_LINE;
    q.status = status == QUERY_ACTIVE && q.i >= k ? QUERY_FOUND : status;
}

// One record of the window, from its descriptor and shortcuts to its table entries.
__device__ __forceinline__ void stage_record(const Staged& st, uint32_t r, const Desc& d, const Quad& skip, uint32_t origin, uint32_t body_lo,
                                             uint32_t body_units) {
    const uint32_t fmt = d.fmt();
    uint32_t kind = KIND_DEFER, target0 = NODE_OUTSIDE, target1 = NODE_OUTSIDE, total = 0, offset0 = 0, offset1 = 0;
    if (fmt == FMT_EMPTY) kind = KIND_EMPTY;
    else if (d.total_len() < 0xFFFFu) {
        total = d.total_len();
        if (fmt == FMT_SINGLE && d.offset0() <= 0xFFFFu) {
            kind = KIND_SINGLE; target0 = window_index(d.node0(), origin, st.count); offset0 = d.offset0();
        } else if (fmt == FMT_DENSE2 && d.offset0() <= 0xFFFFu && d.offset1() <= 0xFFFFu) {
            const uint32_t unit0 = d.body() - body_lo;  // bodies are 32-byte aligned and lie in record order
            if (d.body() >= body_lo && unit0 + 2u * d.body_len() <= body_units && (unit0 / 2u) * 12u < KIND_DEFER) {
                kind = (unit0 / 2u) * 12u;
                target0 = window_index(d.node0(), origin, st.count); target1 = window_index(d.node1(), origin, st.count);
                offset0 = d.offset0(); offset1 = d.offset1();
            }
        }
    }
    // a shortcut whose landing node is not staged or whose offset does not fit is left out (the step then takes one hop)
    const uint32_t land0 = skip.y <= 0xFFFFu ? window_index(skip.x, origin, st.count) : NODE_OUTSIDE;
    const uint32_t land1 = skip.w <= 0xFFFFu ? window_index(skip.z, origin, st.count) : NODE_OUTSIDE;
    sts128(st.rec + 16u * r, target0 | (target1 << 16), total | (kind << 16), land0 | (skip.y << 16), land1 | (skip.w << 16));
    sts32(st.offs + 4u * r, offset0 | (offset1 << 16));
}

// One 192-bit dense block (layout.h) as twelve rank entries.
__device__ __forceinline__ void stage_block(const Staged& st, uint32_t blk, const Quad& lo, const Quad& hi) {
    const uint32_t bits[6] = {lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t ones = lo.x;
    const uint32_t at_block = st.ranks + 48u * blk;
#pragma unroll
    for (uint32_t j = 0; j < 6; j += 2) {
        const uint32_t a0 = bits[j] & 0xFFFFu, a1 = bits[j] >> 16, b0 = bits[j + 1] & 0xFFFFu, b1 = bits[j + 1] >> 16;
        const uint32_t o1 = ones + static_cast<uint32_t>(__popc(a0)), o2 = o1 + static_cast<uint32_t>(__popc(a1));
        const uint32_t o3 = o2 + static_cast<uint32_t>(__popc(b0));
        sts128(at_block + 8u * j, (ones & 0xFFFFu) | (a0 << 16), (o1 & 0xFFFFu) | (a1 << 16), (o2 & 0xFFFFu) | (b0 << 16), (o3 & 0xFFFFu) | (b1 << 16));
        ones = o3 + static_cast<uint32_t>(__popc(b1));
    }
}

constexpr uint32_t SMEM_HEADER = 128;  // control words, keeps the staged arrays 128-byte aligned

// Bytes of shared memory a plan needs.
__host__ __device__ inline uint32_t window_smem_bytes(uint32_t max_records, uint32_t body_cap, uint32_t threads) {
    return SMEM_HEADER + max_records * RECORD_BYTES + (body_cap / 2u) * 48u + (threads / 32u) * TILE_BYTES;
}

template <class T, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) k_find_window(IndexView ix, WindowPlan wp, const T* __restrict__ patterns,
                                                                const uint32_t* __restrict__ perm, const uint32_t* __restrict__ bucket_end,
                                                                uint32_t k, gbwt_b200_state* __restrict__ out,
                                                                uint32_t* __restrict__ deferred, uint32_t* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char smem[];
    volatile uint32_t* ctrl = reinterpret_cast<volatile uint32_t*>(smem);  // [0] window ticket, [1] next query slot
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t base = static_cast<uint32_t>(ix.offset), records = static_cast<uint32_t>(ix.records);
    Staged st;
    st.rec = smem_addr(smem) + SMEM_HEADER;
    st.offs = st.rec + 16u * wp.max_records;
    st.ranks = st.offs + 4u * wp.max_records;
    uint32_t tile = st.ranks + (wp.body_cap / 2u) * 48u + (tid >> 5) * TILE_BYTES;
    // (opaque to the compiler: it would otherwise recompute the shared-window addresses from special registers in every step)
    asm volatile("" : "+r"(st.rec), "+r"(st.offs), "+r"(st.ranks), "+r"(tile));
    const uint32_t slot = tile + 2u * lane;
    sts16(slot + SEGMENT * TILE_LINE, NODE_OUTSIDE);  // the line after a full segment
    // sector loads need every row segment on a 32-byte boundary
    const bool aligned = (reinterpret_cast<uintptr_t>(patterns) & 31u) == 0 && (k * sizeof(T)) % 32u == 0;
    for (;;) {
        __syncthreads();  // everybody has left the previous window: its shared memory and ctrl[] may be reused
        if (tid == 0) { ctrl[0] = atomicAdd(&counters[0], 1u); ctrl[1] = 0; }
        __syncthreads();
        const uint32_t w = ctrl[0];
        if (w >= wp.windows) break;
        // (bucket_end[] has one entry per sort bucket, 2^fine of them per window)
        const uint32_t fine_buckets = ((records - 1u) >> (wp.wshift - wp.fine)) + 1u, after = (w + 1u) << wp.fine;
        const uint32_t q_begin = w == 0 ? 0u : __ldg(bucket_end + (w << wp.fine) - 1u);
        const uint32_t q_end = __ldg(bucket_end + (after < fine_buckets ? after : fine_buckets) - 1u);
        if (q_begin >= q_end) continue;
        // Queries are handed out 32 at a time per warp; with wp.prefetch the pattern rows of a warp's NEXT 32 queries are
        // requested from L2 before it works on the current ones.
        uint32_t at = 0, q_cur = 0;
        if (lane == 0) at = q_begin + atomicAdd(const_cast<uint32_t*>(&ctrl[1]), 32u);
        at = __shfl_sync(0xFFFFFFFFu, at, 0);
        if (at + lane < q_end) {
            q_cur = __ldg(perm + at + lane);
            if (wp.prefetch) prefetch_row(patterns + static_cast<size_t>(q_cur) * k, k);
        }
        const uint32_t r0 = w << wp.wshift;
        st.lo = r0 > wp.margin ? r0 - wp.margin : 0u;
        const uint32_t want_hi = r0 + (1u << wp.wshift) + wp.margin;
        const uint32_t hi = want_hi < records ? want_hi : records;
        st.count = hi - st.lo;
        const uint32_t origin = base + st.lo;
        const uint32_t body_lo = __ldg(ix.stage_body + st.lo / STAGE_GRANULE);
        const uint32_t body_hi = __ldg(ix.stage_body + (hi + STAGE_GRANULE - 1u) / STAGE_GRANULE);
        const uint32_t body_units = body_hi - body_lo < wp.body_cap ? body_hi - body_lo : wp.body_cap;
        // decode the window into shared memory: descriptors + shortcuts, then the dense blocks as {ones before, 16 bits}
        // entries (requesting two records or blocks per thread before decoding either was measured: the extra live
        // registers spill under the 64-register budget of two 512-thread CTAs per SM, 6.2 -> 5.0 G queries/s)
        for (uint32_t r = tid; r < st.count; r += THREADS) {
            Desc d;
            load_sector(reinterpret_cast<const Unit16*>(ix.desc + st.lo + r), d.a, d.b);
            const Quad skip = load_quad(ix.skips + st.lo + r);
            stage_record(st, r, d, skip, origin, body_lo, body_units);
        }
        for (uint32_t blk = tid; blk < body_units / 2u; blk += THREADS) {
            Quad lo, hi;
            load_sector(ix.bodies + body_lo + 2u * blk, lo, hi);
            stage_block(st, blk, lo, hi);
        }
        __syncthreads();
        while (at < q_end) {
            uint32_t next_at = 0, q_next = 0;
            if (lane == 0) next_at = q_begin + atomicAdd(const_cast<uint32_t*>(&ctrl[1]), 32u);
            next_at = __shfl_sync(0xFFFFFFFFu, next_at, 0);
            if (next_at + lane < q_end) {
                q_next = __ldg(perm + next_at + lane);
                if (wp.prefetch) prefetch_row(patterns + static_cast<size_t>(q_next) * k, k);
            }
            const bool mine = at + lane < q_end;
            const T* row = mine ? patterns + static_cast<size_t>(q_cur) * k : nullptr;
            WindowQuery q;
            q.idx = 0; q.start = 0; q.end = 0; q.i = 0; q.status = mine && k != 0 ? QUERY_ACTIVE : QUERY_NONE;
            for (uint32_t seg = 0; seg < k; seg += SEGMENT) {
                if (!__any_sync(0xFFFFFFFFu, q.status == QUERY_ACTIVE)) break;
                const uint32_t n_seg = k - seg < SEGMENT ? k - seg : SEGMENT;
                load_tile<T>(q.status == QUERY_ACTIVE ? row : nullptr, seg, n_seg, aligned && n_seg % (32u / sizeof(T)) == 0, tile, origin, st.count, lane);
                if (n_seg < SEGMENT) sts16(slot + n_seg * TILE_LINE, NODE_OUTSIDE);  // the line after a short segment
                __syncwarp();
                if (q.status == QUERY_ACTIVE) {
                    if (seg == 0) window_find(st, lds16(slot), k, q);
                    if (q.status == QUERY_ACTIVE) window_extend(st, slot, seg, n_seg, k, q);
                }
                __syncwarp();  // the tile is rewritten in the next round
            }
            if (mine) {
                if (q.status == QUERY_DEFER || q.status == QUERY_ACTIVE) {
                    deferred[atomicAdd(&counters[1], 1u)] = q_cur;
                } else {
                    const bool found = q.status == QUERY_FOUND;
                    gbwt_b200_state result;
                    result.node = found ? q.idx + origin : 0u; result.start = found ? q.start : 0u; result.end = found ? q.end : 0u;
                    store_state(out + q_cur, result);
                }
            }
            at = next_at; q_cur = q_next;
        }
    }
}

// Pattern reader of the kernels that keep the chunk in registers (the general loop for deferred queries and the plain
// 32-bit kernels): same interface, no shared memory.
template <class T>
struct PlainRowReader {
    const T* p;
    __device__ __forceinline__ bool node(uint32_t i, uint32_t& out) {
        const uint64_t v = __ldg(p + i);
        out = static_cast<uint32_t>(v);
        return (v >> 32) == 0;
    }
};

// The deferred queries, by the general rounds loop (every record format, every edge case).
template <class T>
__global__ void __launch_bounds__(BLOCK_THREADS) k_find_deferred(IndexView ix, const T* __restrict__ patterns,
                                                                  const uint32_t* __restrict__ deferred, const uint32_t* __restrict__ counters,
                                                                  uint32_t k, gbwt_b200_state* __restrict__ out) {
    const uint32_t n = counters[1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t q = __ldg(deferred + i);
        PlainRowReader<T> rd{patterns + static_cast<size_t>(q) * k};
        gbwt_b200_state result;
        query_find_extend_rounds<true>(ix, rd, k, result);
        store_state(out + q, result);
    }
}

template <class T>
__global__ void __launch_bounds__(BLOCK_THREADS) k_window_keys(IndexView ix, uint32_t wshift, const T* __restrict__ patterns, size_t n,
                                                                size_t k, uint32_t* __restrict__ keys, uint32_t* __restrict__ counts) {
    for (size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < n; q += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint64_t node = first_node(patterns, q, k);
        uint64_t rec;
        const uint32_t b = record_of(ix, node, rec) ? static_cast<uint32_t>(rec >> wshift) : 0u;
        keys[q] = b;
        atomicAdd(counts + 1 + b, 1u);
    }
}

// Plain kernels for 32-bit pattern nodes: the general rounds loop, with or without the locality permutation.
template <bool RUNS>
__global__ void __launch_bounds__(BLOCK_THREADS) k_find_extend_u32(IndexView ix, const uint32_t* __restrict__ patterns,
                                                                    const uint32_t* __restrict__ perm, size_t n, size_t k,
                                                                    gbwt_b200_state* __restrict__ out) {
    GBWT_FOR_EACH_QUERY(q, n, perm) {
        gbwt_b200_state result;
        PlainRowReader<uint32_t> rd{patterns + q * k};
        query_find_extend_rounds<RUNS>(ix, rd, static_cast<uint32_t>(k), result);
        store_state(out + q, result);
    }
}

int env_or(const char* name, int fallback) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : fallback;
}

template <class T, int THREADS, int CTAS>
int launch_window_variant(const IndexView& ix, const WindowPlan& plan, const T* patterns, const uint32_t* perm,
                          const uint32_t* bucket_end, size_t k, gbwt_b200_state* out, uint32_t* deferred, uint32_t* counters,
                          int sm_count, cudaStream_t stream) {
    auto kernel = k_find_window<T, THREADS, CTAS>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.smem_bytes));
    if (e != cudaSuccess) return static_cast<int>(e);
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(plan.windows, static_cast<uint64_t>(sm_count) * CTAS));
    kernel<<<grid, THREADS, plan.smem_bytes, stream>>>(ix, plan, patterns, perm, bucket_end, static_cast<uint32_t>(k), out, deferred, counters);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace

bool plan_windows(const IndexView& ix, uint64_t body_units, WindowPlan& plan) {
    if (!ix.edges_valid || ix.records < 2 || ix.stage_body == nullptr) return false;
    // Shared memory per CTA (two CTAs of 512 threads per SM by default): 40 bytes per staged record, the pattern
    // columns, and what is left holds the bitvectors (48 bytes per 32-byte block of the global layout).
    const uint32_t threads = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW_THREADS", 512));
    const uint32_t smem_kb = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW_SMEM_KB", threads >= 1024 ? 224 : (threads >= 512 ? 112 : 55)));
    uint32_t window = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW", threads >= 1024 ? 1024 : (threads >= 512 ? 512 : 256)));
    uint32_t margin = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW_MARGIN", threads >= 1024 ? 128 : 96));
    if (threads != 256 && threads != 512 && threads != 1024) return false;
    if (window < STAGE_GRANULE || (window & (window - 1)) != 0 || smem_kb > 226 || smem_kb < 16) return false;
    margin = (margin + STAGE_GRANULE - 1) / STAGE_GRANULE * STAGE_GRANULE;
    plan.wshift = 0;
    while ((1u << plan.wshift) < window) plan.wshift++;
    plan.margin = margin;
    plan.max_records = window + 2 * margin;
    const uint64_t fixed = window_smem_bytes(plan.max_records, 0, threads);
    const uint64_t budget = static_cast<uint64_t>(smem_kb) * 1024;
    if (fixed + 4096 > budget) return false;
    plan.body_cap = static_cast<uint32_t>((budget - fixed) / 48) * 2;
    // no point in reserving more than the average window needs several times over
    const uint64_t avg_units = body_units * plan.max_records / std::max<uint64_t>(1, ix.records);
    plan.body_cap = static_cast<uint32_t>(std::min<uint64_t>(plan.body_cap, std::max<uint64_t>(256, 4 * avg_units + 64))) & ~1u;
    plan.windows = static_cast<uint32_t>(((ix.records - 1) >> plan.wshift) + 1);
    plan.threads = threads;
    plan.prefetch = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW_PREFETCH", 0));
    plan.fine = static_cast<uint32_t>(std::min<int>(std::max(0, env_or("GBWT_B200_WINDOW_FINE", 0)), static_cast<int>(plan.wshift)));
    plan.smem_bytes = window_smem_bytes(plan.max_records, plan.body_cap, threads);
    return true;
}

template <class T>
void launch_window_keys(const IndexView& ix, const WindowPlan& plan, const T* patterns, size_t n, size_t k, uint32_t* keys,
                        uint32_t* counts, unsigned grid, cudaStream_t stream) {
    k_window_keys<T><<<grid, BLOCK_THREADS, 0, stream>>>(ix, plan.wshift - plan.fine, patterns, n, k, keys, counts);
}

template <class T>
int launch_find_window(const IndexView& ix, const WindowPlan& plan, const T* patterns, const uint32_t* perm,
                       const uint32_t* bucket_end, size_t n, size_t k, gbwt_b200_state* out, uint32_t* deferred,
                       uint32_t* counters, int sm_count, cudaStream_t stream) {
    (void)n;
    if (plan.threads == 1024) return launch_window_variant<T, 1024, 1>(ix, plan, patterns, perm, bucket_end, k, out, deferred, counters, sm_count, stream);
    if (plan.threads == 256) return launch_window_variant<T, 256, 4>(ix, plan, patterns, perm, bucket_end, k, out, deferred, counters, sm_count, stream);
    return launch_window_variant<T, 512, 2>(ix, plan, patterns, perm, bucket_end, k, out, deferred, counters, sm_count, stream);
}

template <class T>
void launch_find_deferred(const IndexView& ix, const T* patterns, const uint32_t* deferred, const uint32_t* counters, size_t k,
                          gbwt_b200_state* out, unsigned grid, cudaStream_t stream) {
    k_find_deferred<T><<<grid, BLOCK_THREADS, 0, stream>>>(ix, patterns, deferred, counters, static_cast<uint32_t>(k), out);
}

void launch_find_extend_u32(const IndexView& ix, bool runs, const uint32_t* patterns, const uint32_t* perm, size_t n, size_t k,
                            gbwt_b200_state* out, unsigned grid, cudaStream_t stream) {
    if (runs) k_find_extend_u32<true><<<grid, BLOCK_THREADS, 0, stream>>>(ix, patterns, perm, n, k, out);
    else k_find_extend_u32<false><<<grid, BLOCK_THREADS, 0, stream>>>(ix, patterns, perm, n, k, out);
}

template void launch_window_keys<uint64_t>(const IndexView&, const WindowPlan&, const uint64_t*, size_t, size_t, uint32_t*, uint32_t*, unsigned, cudaStream_t);
template void launch_window_keys<uint32_t>(const IndexView&, const WindowPlan&, const uint32_t*, size_t, size_t, uint32_t*, uint32_t*, unsigned, cudaStream_t);
template int launch_find_window<uint64_t>(const IndexView&, const WindowPlan&, const uint64_t*, const uint32_t*, const uint32_t*, size_t, size_t,
                                          gbwt_b200_state*, uint32_t*, uint32_t*, int, cudaStream_t);
template int launch_find_window<uint32_t>(const IndexView&, const WindowPlan&, const uint32_t*, const uint32_t*, const uint32_t*, size_t, size_t,
                                          gbwt_b200_state*, uint32_t*, uint32_t*, int, cudaStream_t);
template void launch_find_deferred<uint64_t>(const IndexView&, const uint64_t*, const uint32_t*, const uint32_t*, size_t, gbwt_b200_state*, unsigned, cudaStream_t);
template void launch_find_deferred<uint32_t>(const IndexView&, const uint32_t*, const uint32_t*, const uint32_t*, size_t, gbwt_b200_state*, unsigned, cudaStream_t);

}  // namespace gbwt_b200
