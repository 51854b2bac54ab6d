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

// ---- mbarrier / bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP) -------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_global_to_shared(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done, tries = 0;
    do {
        // (a copy that never completes would otherwise hang the device: fail the launch instead)
        if (++tries > (1u << 24)) __trap();
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (done == 0);
}

// ---- pattern rows -----------------------------------------------------------------------------------------------------
// A thread reads its own pattern row a chunk of four nodes at a time and always has the next chunk in flight (the
// rows of a sorted batch are scattered over HBM; this is the one load that pays DRAM latency, and nothing depends on
// it for four nodes). The loads bypass L1 (every byte is used once) and ask L2 to fetch the whole row on first touch.

template <class T>
struct RowReader;

template <>
struct RowReader<uint64_t> {
    const uint64_t* p;
    uint32_t k, base;
    uint32_t c0, c1, c2, c3, bad;
    uint64_t n0, n1, n2, n3;
    bool vec;
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
    __device__ __forceinline__ RowReader(const uint64_t* row, uint32_t len)
        : p(row), k(len), base(0xFFFFFFFFu), c0(0), c1(0), c2(0), c3(0), bad(0), vec((reinterpret_cast<uintptr_t>(row) & 31) == 0) {
        fetch(0);
    }
    // nodes are asked for in increasing order of i; false = the node does not fit 32 bits (it cannot be in the index)
    __device__ __forceinline__ bool node(uint32_t i, uint32_t& out) {
        const uint32_t b = i & ~3u;
        if (b != base) {
            base = b;
            c0 = static_cast<uint32_t>(n0); c1 = static_cast<uint32_t>(n1);
            c2 = static_cast<uint32_t>(n2); c3 = static_cast<uint32_t>(n3);
            bad = ((n0 >> 32) != 0 ? 1u : 0u) | ((n1 >> 32) != 0 ? 2u : 0u) | ((n2 >> 32) != 0 ? 4u : 0u) | ((n3 >> 32) != 0 ? 8u : 0u);
            fetch(b + 4);
        }
        const uint32_t j = i & 3u;
        out = j == 0 ? c0 : (j == 1 ? c1 : (j == 2 ? c2 : c3));
        return ((bad >> j) & 1u) == 0;
    }
};

template <>
struct RowReader<uint32_t> {
    const uint32_t* p;
    uint32_t k, base;
    uint32_t c0, c1, c2, c3;
    uint32_t n0, n1, n2, n3;
    bool vec;
    __device__ __forceinline__ void fetch(uint32_t b) {
        n0 = n1 = n2 = n3 = 0;
        if (b >= k) return;
        if (vec && k - b >= 4) {
            asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(n0), "=r"(n1), "=r"(n2), "=r"(n3) : "l"(p + b));
        } else {
            n0 = __ldg(p + b);
            if (b + 1 < k) n1 = __ldg(p + b + 1);
            if (b + 2 < k) n2 = __ldg(p + b + 2);
            if (b + 3 < k) n3 = __ldg(p + b + 3);
        }
    }
    __device__ __forceinline__ RowReader(const uint32_t* row, uint32_t len)
        : p(row), k(len), base(0xFFFFFFFFu), c0(0), c1(0), c2(0), c3(0), vec((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        fetch(0);
    }
    __device__ __forceinline__ bool node(uint32_t i, uint32_t& out) {
        const uint32_t b = i & ~3u;
        if (b != base) {
            base = b;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            fetch(b + 4);
        }
        const uint32_t j = i & 3u;
        out = j == 0 ? c0 : (j == 1 ? c1 : (j == 2 ? c2 : c3));
        return true;
    }
};

// First node of pattern row q as a 64-bit value (the sort key is computed from it).
__device__ __forceinline__ uint64_t first_node(const uint64_t* patterns, size_t q, size_t k) { return __ldg(patterns + q * k); }
__device__ __forceinline__ uint64_t first_node(const uint32_t* patterns, size_t q, size_t k) { return __ldg(patterns + q * k); }

// ---- the staged window -----------------------------------------------------------------------------------------------

struct Staged {
    const uint4* desc;   // two per record: {total_len, meta, node0, offset0}, {body, body_len, node1, offset1}
    const uint2* skip;   // two per record: {landing node, offset} over edge 0 and over edge 1
    const uint4* body;   // 16-byte units, the first one is unit `body_lo` of IndexView::bodies
    uint32_t lo, count;  // staged records [lo, lo + count)
    uint32_t body_lo, body_units;
};

// rank1 of position p inside a dense body that starts at staged unit `unit0`, and the bit at p (layout.h: 32-byte
// blocks {ones_before, c0 | c1 << 8, 192 bits}): the block header and ONE 64-bit word, both 8-byte shared loads.
__device__ __forceinline__ uint32_t staged_rank1(const Staged& st, uint32_t unit0, uint32_t p, uint32_t& bit) {
    const uint32_t blk = __umulhi(p, 0xAAAAAAABu) >> 7;  // p / 192
    const uint32_t r = p - blk * DENSE_BITS;
    const uint32_t j = r >> 6, sh = r & 63u;
    const uint2* block = reinterpret_cast<const uint2*>(st.body + unit0 + 2u * blk);
    const uint2 hdr = block[0];
    const uint2 w = block[1 + j];
    const uint64_t word = (static_cast<uint64_t>(w.y) << 32) | w.x;
    const uint32_t sub = ((hdr.y << 8) >> (8u * j)) & 0xFFu;  // 0, c0, c1
    bit = static_cast<uint32_t>(word >> sh) & 1u;
    return hdr.x + sub + static_cast<uint32_t>(__popcll(word & ((1ull << sh) - 1ull)));
}

enum : int { QUERY_DONE = 0, QUERY_DEFER = 1 };

// One query against the staged window. QUERY_DONE: `out` holds the reference's answer. QUERY_DEFER: the window could
// not decide; the general kernel redoes the query from the start.
template <class Reader>
__device__ __forceinline__ int window_query(const Staged& st, uint32_t base, Reader& rd, uint32_t k, gbwt_b200_state& out) {
    set_none(out);
    if (k == 0) return QUERY_DONE;
    uint32_t x;
    if (!rd.node(0, x)) return QUERY_DONE;  // GBWT::find: not a node of the index
    uint32_t idx = x - base - st.lo;
    if (idx >= st.count || idx + st.lo == 0) return QUERY_DEFER;  // (record 0 is the endmarker: find() is None; let the general code say so)
    uint4 da = st.desc[2u * idx];
    uint32_t start = 0, end = da.x, node = x;
    uint32_t fmt = (da.y >> 16) & 0xFFu;
    if (fmt == FMT_EMPTY || end == 0) return QUERY_DONE;
    uint32_t i = 1;
    while (i < k) {
        uint32_t x1;
        if (!rd.node(i, x1) || x1 == 0) return QUERY_DONE;  // GBWT::extend: below first_node / not an edge target
        const uint32_t total = da.x;
        const uint32_t s = start < total ? start : total, e = end < total ? end : total;
        if (s >= e) return QUERY_DONE;
        uint32_t b, edge_offset, rs, re;
        if (fmt == FMT_SINGLE) {
            if (x1 != da.z) return QUERY_DONE;
            b = 0; edge_offset = da.w; rs = s; re = e;
        } else if (fmt == FMT_DENSE2) {
            const uint4 db = st.desc[2u * idx + 1u];
            if (x1 == da.z) { b = 0; edge_offset = da.w; }
            else if (x1 == db.z) { b = 1; edge_offset = db.w; }
            else return QUERY_DONE;
            const uint32_t unit0 = db.x - st.body_lo;
            const uint32_t last_blk = __umulhi(e - 1u, 0xAAAAAAABu) >> 7;
            if (unit0 + 2u * last_blk + 2u > st.body_units) return QUERY_DEFER;  // the body did not fit the window
            // rank1(s), and rank1(e) = rank1(e - 1) + bit(e - 1); a range of one position needs one lookup
            uint32_t bit;
            const uint32_t ones_s = staged_rank1(st, unit0, s, bit);
            uint32_t ones_e = ones_s + bit;
            if (e - 1u != s) { ones_e = staged_rank1(st, unit0, e - 1u, bit); ones_e += bit; }
            rs = b ? ones_s : s - ones_s;
            re = b ? ones_e : e - ones_e;
            if (rs >= re) return QUERY_DONE;
        } else if (fmt == FMT_EMPTY) {
            return QUERY_DONE;  // BWT::record() is None
        } else {
            return QUERY_DEFER;  // run-length body or outdegree > 2
        }
        // two hops at once when the successor is a single-edge record leading to the pattern node after x1
        bool hopped = false;
        if (i + 1 < k) {
            const uint2 sk = st.skip[2u * idx + b];
            uint32_t x2;
            if (sk.x != 0 && rd.node(i + 1, x2) && x2 == sk.x) {
                start = sk.y + rs; end = sk.y + re;
                node = x2; i += 2;
                hopped = true;
            }
        }
        if (!hopped) {
            start = edge_offset + rs; end = edge_offset + re;
            node = x1; i += 1;
        }
        if (i >= k) break;
        idx = node - base - st.lo;
        if (idx >= st.count) return QUERY_DEFER;
        da = st.desc[2u * idx];
        fmt = (da.y >> 16) & 0xFFu;
    }
    out.node = node; out.start = start; out.end = end;
    return QUERY_DONE;
}

constexpr uint32_t SMEM_HEADER = 128;  // mbarrier + control words, keeps the staged arrays 128-byte aligned

template <class T, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) k_find_window(IndexView ix, WindowPlan wp, const T* __restrict__ patterns,
                                                                const uint32_t* __restrict__ perm, const uint32_t* __restrict__ bucket_end,
                                                                uint32_t k, gbwt_b200_state* __restrict__ out,
                                                                uint32_t* __restrict__ deferred, uint32_t* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    volatile uint32_t* ctrl = reinterpret_cast<volatile uint32_t*>(smem + 16);  // [0] window ticket, [1] next query slot
    uint4* s_desc = reinterpret_cast<uint4*>(smem + SMEM_HEADER);
    uint4* s_skip = s_desc + 2u * wp.max_records;
    uint4* s_body = s_skip + wp.max_records;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t base = static_cast<uint32_t>(ix.offset), records = static_cast<uint32_t>(ix.records);
    if (tid == 0) mbar_init(bar, 1);
    uint32_t parity = 0;
    for (;;) {
        __syncthreads();  // everybody has left the previous window: its shared memory and ctrl[] may be reused
        if (tid == 0) ctrl[0] = atomicAdd(&counters[0], 1u);
        __syncthreads();
        const uint32_t w = ctrl[0];
        if (w >= wp.windows) break;
        const uint32_t q_begin = w == 0 ? 0u : __ldg(bucket_end + w - 1), q_end = __ldg(bucket_end + w);
        if (q_begin >= q_end) continue;
        const uint32_t r0 = w << wp.wshift;
        Staged st;
        st.lo = r0 > wp.margin ? r0 - wp.margin : 0u;
        const uint32_t want_hi = r0 + (1u << wp.wshift) + wp.margin;
        const uint32_t hi = want_hi < records ? want_hi : records;
        st.count = hi - st.lo;
        st.body_lo = __ldg(ix.stage_body + st.lo / STAGE_GRANULE);
        const uint32_t body_hi = __ldg(ix.stage_body + (hi + STAGE_GRANULE - 1u) / STAGE_GRANULE);
        st.body_units = body_hi - st.body_lo < wp.body_cap ? body_hi - st.body_lo : wp.body_cap;
        st.desc = s_desc; st.skip = reinterpret_cast<const uint2*>(s_skip); st.body = s_body;
        if (tid == 0) {
            ctrl[1] = q_begin;
            mbar_expect_tx(bar, st.count * 48u + st.body_units * 16u);
            bulk_global_to_shared(s_desc, ix.desc + st.lo, st.count * 32u, bar);
            bulk_global_to_shared(s_skip, ix.skips + st.lo, st.count * 16u, bar);
            if (st.body_units != 0) bulk_global_to_shared(s_body, ix.bodies + st.body_lo, st.body_units * 16u, bar);
        }
        __syncthreads();  // ctrl[1] is set
        mbar_wait(bar, parity);
        parity ^= 1u;
        // the window's queries, 32 at a time per warp
        for (;;) {
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(const_cast<uint32_t*>(&ctrl[1]), 32u);
            slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
            if (slot >= q_end) break;
            const uint32_t at = slot + lane;
            if (at < q_end) {
                const uint32_t q = __ldg(perm + at);
                RowReader<T> rd(patterns + static_cast<size_t>(q) * k, k);
                gbwt_b200_state result;
                if (window_query(st, base, rd, k, result) == QUERY_DONE) store_state(out + q, result);
                else deferred[atomicAdd(&counters[1], 1u)] = q;
            }
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
        RowReader<T> rd(patterns + static_cast<size_t>(q) * k, k);
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
        RowReader<uint32_t> rd(patterns + q * k, static_cast<uint32_t>(k));
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
    // Shared memory per CTA (two CTAs of 512 threads per SM by default): what is left after descriptors and
    // shortcuts (48 bytes per staged record) holds the bodies.
    const uint32_t threads = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW_THREADS", 512));
    const uint32_t smem_kb = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW_SMEM_KB", threads >= 1024 ? 200 : (threads >= 512 ? 100 : 50)));
    uint32_t window = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW", threads >= 1024 ? 1024 : (threads >= 512 ? 512 : 256)));
    uint32_t margin = static_cast<uint32_t>(env_or("GBWT_B200_WINDOW_MARGIN", 128));
    if (threads != 256 && threads != 512 && threads != 1024) return false;
    if (window < STAGE_GRANULE || (window & (window - 1)) != 0 || smem_kb > 227 || smem_kb < 16) return false;
    margin = (margin + STAGE_GRANULE - 1) / STAGE_GRANULE * STAGE_GRANULE;
    plan.wshift = 0;
    while ((1u << plan.wshift) < window) plan.wshift++;
    plan.margin = margin;
    plan.max_records = window + 2 * margin;
    const uint64_t fixed = SMEM_HEADER + static_cast<uint64_t>(plan.max_records) * 48;
    const uint64_t budget = static_cast<uint64_t>(smem_kb) * 1024;
    if (fixed + 4096 > budget) return false;
    plan.body_cap = static_cast<uint32_t>((budget - fixed) / 16);
    // no point in reserving more than the average window needs several times over
    const uint64_t avg_units = body_units * plan.max_records / std::max<uint64_t>(1, ix.records);
    plan.body_cap = static_cast<uint32_t>(std::min<uint64_t>(plan.body_cap, std::max<uint64_t>(256, 4 * avg_units + 64)));
    plan.windows = static_cast<uint32_t>(((ix.records - 1) >> plan.wshift) + 1);
    plan.threads = threads;
    plan.smem_bytes = static_cast<uint32_t>(fixed + static_cast<uint64_t>(plan.body_cap) * 16);
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
