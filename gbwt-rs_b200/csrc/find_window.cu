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
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }

// ---- pattern rows -----------------------------------------------------------------------------------------------------
// A thread reads its own pattern row one 32-byte sector at a time and always has the next sector in flight (the rows
// of a sorted batch are scattered over HBM: this is the one load of the loop that pays DRAM latency, and nothing
// depends on it for a whole chunk). The loads bypass L1 (every byte is used once) and ask L2 for the whole row on
// first touch. The nodes of the current chunk live in the thread's own column of a small shared array (word j of
// thread t at [j][t]: every lane always hits its own bank), so that node(i) is one conflict-free shared load however
// i moves -- with the chunk in registers the select by i & 3 compiled to branches that split the warp.

template <class T>
struct RowReader;

// 64-bit nodes: chunks of 4. A chunk with a node that does not fit 32 bits fails as a whole (the caller defers the query
// to the general kernel, which looks at the nodes one by one).
template <>
struct RowReader<uint64_t> {
    const uint64_t* p;
    uint32_t k, base, slot, stride;  // slot: shared address of this thread's column, stride: bytes between its words
    uint64_t n0, n1, n2, n3;
    bool bad, vec;
    __device__ __forceinline__ void fetch(uint32_t b) {
        n0 = n1 = n2 = n3 = 0;
        if (b >= k) return;
        if (vec && k - b >= 4) {
            asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(n0), "=l"(n1), "=l"(n2), "=l"(n3) : "l"(p + b));
        } else {
            n0 = __ldg(p + b);
            if (b + 1 < k) n1 = __ldg(p + b + 1);
            if (b + 2 < k) n2 = __ldg(p + b + 2);
            if (b + 3 < k) n3 = __ldg(p + b + 3);
        }
    }
    __device__ __forceinline__ RowReader(const uint64_t* row, uint32_t len, uint32_t slot_addr, uint32_t stride_bytes)
        : p(row), k(len), base(0xFFFFFFFFu), slot(slot_addr), stride(stride_bytes), bad(false), vec((reinterpret_cast<uintptr_t>(row) & 31) == 0) {
        fetch(0);
    }
    // nodes are asked for in non-decreasing chunk order; false = some node of the chunk does not fit 32 bits
    __device__ __forceinline__ bool node(uint32_t i, uint32_t& out) {
        const uint32_t b = i & ~3u;
        if (b != base) {
            base = b;
            sts32(slot, static_cast<uint32_t>(n0)); sts32(slot + stride, static_cast<uint32_t>(n1));
            sts32(slot + 2 * stride, static_cast<uint32_t>(n2)); sts32(slot + 3 * stride, static_cast<uint32_t>(n3));
            bad = ((n0 | n1 | n2 | n3) >> 32) != 0;
            fetch(b + 4);
        }
        out = lds32(slot + (i & 3u) * stride);
        return !bad;
    }
};

// 32-bit nodes: one 32-byte sector (8 nodes) per load, handed to the shared column four at a time like the 64-bit
// reader (a first version with 16-byte loads made L2 fetch every row several times: 27.6 GB of DRAM reads per 2^25
// queries instead of 7.7).
template <>
struct RowReader<uint32_t> {
    const uint32_t* p;
    uint32_t k, base, slot, stride;
    uint32_t n0, n1, n2, n3, n4, n5, n6, n7;  // the sector that holds chunk `base` (or, before the first call, node 0)
    bool vec;
    __device__ __forceinline__ void fetch(uint32_t b) {
        n0 = n1 = n2 = n3 = n4 = n5 = n6 = n7 = 0;
        if (b >= k) return;
        if (vec && k - b >= 8) {
            asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(n0), "=r"(n1), "=r"(n2), "=r"(n3), "=r"(n4), "=r"(n5), "=r"(n6), "=r"(n7) : "l"(p + b));
        } else {
            n0 = __ldg(p + b);
            if (b + 1 < k) n1 = __ldg(p + b + 1);
            if (b + 2 < k) n2 = __ldg(p + b + 2);
            if (b + 3 < k) n3 = __ldg(p + b + 3);
            if (b + 4 < k) n4 = __ldg(p + b + 4);
            if (b + 5 < k) n5 = __ldg(p + b + 5);
            if (b + 6 < k) n6 = __ldg(p + b + 6);
            if (b + 7 < k) n7 = __ldg(p + b + 7);
        }
    }
    __device__ __forceinline__ RowReader(const uint32_t* row, uint32_t len, uint32_t slot_addr, uint32_t stride_bytes)
        : p(row), k(len), base(0xFFFFFFFFu), slot(slot_addr), stride(stride_bytes), vec((reinterpret_cast<uintptr_t>(row) & 31) == 0) {
        fetch(0);
    }
    // chunks are visited in order, none skipped (the search loop advances by at most two nodes)
    __device__ __forceinline__ bool node(uint32_t i, uint32_t& out) {
        const uint32_t b = i & ~3u;
        if (b != base) {
            base = b;
            if ((b & 4u) == 0) {
                sts32(slot, n0); sts32(slot + stride, n1); sts32(slot + 2 * stride, n2); sts32(slot + 3 * stride, n3);
            } else {
                sts32(slot, n4); sts32(slot + stride, n5); sts32(slot + 2 * stride, n6); sts32(slot + 3 * stride, n7);
                fetch(b + 4);  // the next sector, four nodes early
            }
        }
        out = lds32(slot + (i & 3u) * stride);
        return true;
    }
};

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
// record times 16 steps. Arranged for the loop, not like the global layout (reading the 32-byte global records in
// place put every 16-byte access on one of four bank groups: 13 wavefronts per LDS.128 instead of 4, see
// profiles/r2_find_window_v1_tma_raw_layout_ncu.txt):
//   hot[r]   16 B  {node0, node1, total_len, kind}: kind = word index of the record's bitvector in `words` for a dense
//                  record, or KIND_SINGLE / KIND_EMPTY / KIND_DEFER (run-length body, outdegree > 2, body not staged)
//   pair[r]  16 B  the two-hop shortcut {landing node, offset} per edge (layout.h, IndexView::skips), as in HBM
//   offs[r]   8 B  {offset0, offset1}: only read when a step cannot take the shortcut
//   words[]   8 B  {ones before this word, 32 bits}: rank at position p of a record is ONE 8-byte load at kind + p / 32
//                  (the 192-bit blocks of the global layout hold 6 such words each, so p / 32 needs no block arithmetic)
constexpr uint32_t KIND_SINGLE = 0xFFFFFFFFu, KIND_EMPTY = 0xFFFFFFFEu, KIND_DEFER = 0xFFFFFFFDu;  // anything below: dense

struct Staged {
    uint32_t hot, pair, offs, words;  // shared-space addresses
    uint32_t lo, count;               // staged records [lo, lo + count)
};

enum : uint32_t { QUERY_FOUND = 0, QUERY_NONE = 1, QUERY_DEFER = 2 };

// rank1(p) and the bit at p of the dense record whose first word is `kind`
__device__ __forceinline__ uint32_t staged_rank1(const Staged& st, uint32_t kind, uint32_t p, uint32_t& bit) {
    const uint2 w = lds64(st.words + 8u * (kind + (p >> 5)));
    const uint32_t sh = p & 31u;
    bit = (w.y >> sh) & 1u;
    return w.x + static_cast<uint32_t>(__popc(w.y & ((1u << sh) - 1u)));
}

// One query against the staged window. QUERY_FOUND: (node, start, end) is the reference's SearchState; QUERY_NONE: the
// reference returns None; QUERY_DEFER: the window could not decide and the general kernel redoes the query from the
// start. The loop has ONE exit (every failure breaks out with its status), so the lanes of a warp reconverge after
// every step instead of carrying a stack of divergent returns.
template <class Reader>
__device__ __forceinline__ uint32_t window_query(const Staged& st, uint32_t base, Reader& rd, uint32_t k, uint32_t& node, uint32_t& start,
                                                 uint32_t& end) {
    if (k == 0) return QUERY_NONE;
    uint32_t x;
    if (!rd.node(0, x)) return QUERY_DEFER;
    uint32_t idx = x - base - st.lo;
    if (idx >= st.count || idx + st.lo == 0) return QUERY_DEFER;  // (record 0 is the endmarker: find() is None; let the general code say so)
    uint4 h = lds128(st.hot + 16u * idx);
    start = 0; end = h.z; node = x;
    if (h.w == KIND_DEFER) return QUERY_DEFER;
    if (h.w == KIND_EMPTY || end == 0) return QUERY_NONE;  // GBWT::find: no record
    uint32_t i = 1, status = QUERY_FOUND;
    while (i < k) {
        uint32_t x1;
        if (!rd.node(i, x1)) { status = QUERY_DEFER; break; }
        const uint32_t total = h.z, kind = h.w;
        const uint32_t s = start < total ? start : total, e = end < total ? end : total;
        // GBWT::extend: below first_node, or (Record::follow) an empty range
        if (x1 == 0 || s >= e) { status = QUERY_NONE; break; }
        uint32_t b = 0, rs = s, re = e;
        if (kind < KIND_DEFER) {
            // dense record: rank1(s), and rank1(e) = rank1(e - 1) + bit(e - 1); a range of one position needs one lookup
            b = x1 == h.y ? 1u : 0u;
            if (x1 != h.x && x1 != h.y) { status = QUERY_NONE; break; }
            uint32_t bit;
            const uint32_t ones_s = staged_rank1(st, kind, s, bit);
            uint32_t ones_e = ones_s + bit;
            if (e - 1u != s) { ones_e = staged_rank1(st, kind, e - 1u, bit); ones_e += bit; }
            rs = b ? ones_s : s - ones_s;
            re = b ? ones_e : e - ones_e;
            if (rs >= re) { status = QUERY_NONE; break; }
        } else if (kind == KIND_SINGLE) {
            if (x1 != h.x) { status = QUERY_NONE; break; }
        } else {
            status = kind == KIND_EMPTY ? QUERY_NONE : QUERY_DEFER;  // BWT::record() is None / a record the window does not decode
            break;
        }
        // two hops at once when the successor is a single-edge record leading to the pattern node after x1
        const uint2 hop = lds64(st.pair + 16u * idx + 8u * b);
        uint32_t x2 = 0;
        if (i + 1 < k && hop.x != 0 && !rd.node(i + 1, x2)) { status = QUERY_DEFER; break; }
        if (x2 == hop.x && x2 != 0) {
            start = hop.y + rs; end = hop.y + re;
            node = x2; i += 2;
        } else {
            const uint32_t edge_offset = lds32(st.offs + 8u * idx + 4u * b);
            start = edge_offset + rs; end = edge_offset + re;
            node = x1; i += 1;
        }
        if (i >= k) break;
        idx = node - base - st.lo;
        if (idx >= st.count) { status = QUERY_DEFER; break; }
        h = lds128(st.hot + 16u * idx);
    }
    return status;
}

constexpr uint32_t SMEM_HEADER = 128;  // control words, keeps the staged arrays 128-byte aligned
constexpr uint32_t PATTERN_WORDS = 4;  // words of the pattern column per thread (RowReader::CHUNK)

// Bytes of shared memory a plan needs.
__host__ __device__ inline uint32_t window_smem_bytes(uint32_t max_records, uint32_t body_cap, uint32_t threads) {
    return SMEM_HEADER + max_records * 40u + (body_cap / 2u) * 48u + threads * PATTERN_WORDS * 4u;
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
    st.hot = smem_addr(smem) + SMEM_HEADER;
    st.pair = st.hot + 16u * wp.max_records;
    st.offs = st.pair + 16u * wp.max_records;
    st.words = st.offs + 8u * wp.max_records;
    const uint32_t pattern_slot = st.words + (wp.body_cap / 2u) * 48u + 4u * tid;
    for (;;) {
        __syncthreads();  // everybody has left the previous window: its shared memory and ctrl[] may be reused
        if (tid == 0) { ctrl[0] = atomicAdd(&counters[0], 1u); ctrl[1] = 0; }
        __syncthreads();
        const uint32_t w = ctrl[0];
        if (w >= wp.windows) break;
        const uint32_t q_begin = w == 0 ? 0u : __ldg(bucket_end + w - 1), q_end = __ldg(bucket_end + w);
        if (q_begin >= q_end) continue;
        // Queries are handed out 32 at a time per warp; the pattern rows of a warp's NEXT 32 queries are requested from
        // L2 before it works on the current ones, so that the row reads of the loop find them there (the rows of a
        // sorted batch are scattered over HBM, and a thread has only one sector of its row in flight at a time).
        uint32_t slot = 0, q_cur = 0;
        if (lane == 0) slot = q_begin + atomicAdd(const_cast<uint32_t*>(&ctrl[1]), 32u);
        slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
        if (slot + lane < q_end) {
            q_cur = __ldg(perm + slot + lane);
            if (wp.prefetch) prefetch_row(patterns + static_cast<size_t>(q_cur) * k, k);
        }
        const uint32_t r0 = w << wp.wshift;
        st.lo = r0 > wp.margin ? r0 - wp.margin : 0u;
        const uint32_t want_hi = r0 + (1u << wp.wshift) + wp.margin;
        const uint32_t hi = want_hi < records ? want_hi : records;
        st.count = hi - st.lo;
        const uint32_t body_lo = __ldg(ix.stage_body + st.lo / STAGE_GRANULE);
        const uint32_t body_hi = __ldg(ix.stage_body + (hi + STAGE_GRANULE - 1u) / STAGE_GRANULE);
        const uint32_t body_units = body_hi - body_lo < wp.body_cap ? body_hi - body_lo : wp.body_cap;
        // decode the window into shared memory: descriptors + shortcuts ...
        for (uint32_t r = tid; r < st.count; r += THREADS) {
            Desc d;
            load_sector(reinterpret_cast<const Unit16*>(ix.desc + st.lo + r), d.a, d.b);
            const Quad skip = load_quad(ix.skips + st.lo + r);
            const uint32_t fmt = d.fmt();
            uint32_t kind = KIND_DEFER, node1 = 0;
            if (fmt == FMT_SINGLE) kind = KIND_SINGLE;
            else if (fmt == FMT_EMPTY) kind = KIND_EMPTY;
            else if (fmt == FMT_DENSE2) {
                const uint32_t unit0 = d.body() - body_lo;  // bodies are 32-byte aligned and lie in record order
                if (d.body() >= body_lo && unit0 + 2u * d.body_len() <= body_units) { kind = (unit0 / 2u) * 6u; node1 = d.node1(); }
            }
            sts128(st.hot + 16u * r, d.node0(), node1, fmt == FMT_EMPTY ? 0u : d.total_len(), kind);
            sts128(st.pair + 16u * r, skip.x, skip.y, skip.z, skip.w);
            sts64(st.offs + 8u * r, d.offset0(), d.offset1());
        }
        // ... and the dense blocks as {ones before, 32 bits} words
        for (uint32_t blk = tid; blk < body_units / 2u; blk += THREADS) {
            Quad lo, hi;
            load_sector(ix.bodies + body_lo + 2u * blk, lo, hi);
            const uint32_t bits[6] = {lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            uint32_t ones = lo.x;
            const uint32_t at = st.words + 48u * blk;
#pragma unroll
            for (uint32_t j = 0; j < 6; j += 2) {
                const uint32_t next = ones + static_cast<uint32_t>(__popc(bits[j]));
                sts128(at + 8u * j, ones, bits[j], next, bits[j + 1]);
                ones = next + static_cast<uint32_t>(__popc(bits[j + 1]));
            }
        }
        __syncthreads();
        while (slot < q_end) {
            uint32_t next_slot = 0, q_next = 0;
            if (lane == 0) next_slot = q_begin + atomicAdd(const_cast<uint32_t*>(&ctrl[1]), 32u);
            next_slot = __shfl_sync(0xFFFFFFFFu, next_slot, 0);
            if (next_slot + lane < q_end) {
                q_next = __ldg(perm + next_slot + lane);
                if (wp.prefetch) prefetch_row(patterns + static_cast<size_t>(q_next) * k, k);
            }
            if (slot + lane < q_end) {
                const uint32_t q = q_cur;
                RowReader<T> rd(patterns + static_cast<size_t>(q) * k, k, pattern_slot, 4u * THREADS);
                uint32_t node = 0, start = 0, end = 0;
                const uint32_t status = window_query(st, base, rd, k, node, start, end);
                if (status == QUERY_DEFER) {
                    deferred[atomicAdd(&counters[1], 1u)] = q;
                } else {
                    const bool found = status == QUERY_FOUND;
                    gbwt_b200_state result;
                    result.node = found ? node : 0u; result.start = found ? start : 0u; result.end = found ? end : 0u;
                    store_state(out + q, result);
                }
            }
            slot = next_slot; q_cur = q_next;
        }
    }
}

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
    plan.smem_bytes = window_smem_bytes(plan.max_records, plan.body_cap, threads);
    return true;
}

template <class T>
void launch_window_keys(const IndexView& ix, const WindowPlan& plan, const T* patterns, size_t n, size_t k, uint32_t* keys,
                        uint32_t* counts, unsigned grid, cudaStream_t stream) {
    k_window_keys<T><<<grid, BLOCK_THREADS, 0, stream>>>(ix, plan.wshift, patterns, n, k, keys, counts);
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
