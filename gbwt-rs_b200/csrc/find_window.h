// find_window.h -- host-side launchers of the record-window search kernels (find_window.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gbwt_b200.h"
#include "layout.h"

namespace gbwt_b200 {

// How a batch is cut into record windows. One bucket of the locality sort = one window of 2^wshift records; a CTA
// decodes the window plus `margin` records either side (descriptors, two-hop shortcuts and the contiguous bodies)
// into shared memory and resolves the window's queries from there.
struct WindowPlan {
    uint32_t wshift;       // log2(records per window)
    uint32_t margin;       // records staged before and after the window (multiple of STAGE_GRANULE)
    uint32_t windows;      // number of windows = buckets of the sort
    uint32_t max_records;  // (1 << wshift) + 2 * margin
    uint32_t body_cap;     // 16-byte units of shared memory set aside for bodies
    uint32_t threads;      // CTA size: 256, 512 or 1024
    uint32_t smem_bytes;   // dynamic shared memory per CTA
    uint32_t prefetch;     // ask L2 for the pattern rows of a warp's next 32 queries (GBWT_B200_WINDOW_PREFETCH)
    float edge_span;       // what the plan was made from (a batch of patterns of another length is planned again)
    uint64_t body_units;
    uint32_t direct_cap;   // 0: perm is sorted exactly; else window w owns perm[w cap .. (w + 1) cap) (one-pass placement, set per launch)
    uint32_t wide;         // 1: the instantiation that also decodes DENSE4 / byte-per-run records
    uint32_t aux_cap;      // wide records (up to four edges, DENSE4 or byte-per-run body) a window can hold
    uint32_t fine;         // the sort's buckets are 2^(wshift - fine) records: 2^fine buckets per window, so that the queries a
                           // warp takes together start within a few records of each other (GBWT_B200_WINDOW_FINE)
};

// Chooses the plan for an index (GBWT_B200_WINDOW / _MARGIN / _SMEM_KB / _THREADS override). False = do not use
// the window kernel for this index.
// `edge_span` = mean distance, in records, from a record to the targets of its edges (the margin is what a pattern of 32
// nodes travels at that rate).
// `wide`: the index has DENSE4 or byte-per-run records, which the kernel answers from shared memory on a slower path.
// `pattern_len`: the plan of an index is made for patterns of 32 nodes; a batch of longer or shorter patterns gets its own
// (same windows, other margins).
bool plan_windows(const IndexView& ix, uint64_t body_units, double edge_span, bool wide, uint32_t pattern_len, WindowPlan& plan);

// keys[q] = window of query q's first node (0 when it has no record), counts[1 + window] += 1.
// T = uint64_t or uint32_t pattern nodes.
template <class T>
void launch_window_keys(const IndexView& ix, const WindowPlan& plan, const T* patterns, size_t n, size_t k, uint32_t* keys,
                        uint32_t* counts, unsigned grid, cudaStream_t stream);

// One-pass placement for the window search (plan.direct_cap slots per window, filled[w] = queries that asked for window w,
// zero on entry): see k_window_place_direct. Queries that find their window full are appended to `deferred` (counters[1]).
template <class T>
void launch_window_place_direct(const IndexView& ix, const WindowPlan& plan, const T* patterns, size_t n, size_t k, uint32_t* filled,
                                uint32_t* slots, uint32_t* deferred, uint32_t* counters, unsigned grid, cudaStream_t stream);

// The search itself. bucket_end[w] = end of window w's slots in perm (what the scatter leaves in the cursor array);
// counters[0] = window ticket, counters[1] = number of deferred queries (both zero on entry); queries the window
// kernel cannot finish from shared memory are listed in `deferred` and finished by launch_find_deferred.
template <class T>
int launch_find_window(const IndexView& ix, const WindowPlan& plan, const T* patterns, const uint32_t* perm,
                       const uint32_t* bucket_end, size_t n, size_t k, gbwt_b200_state* out, uint32_t* deferred,
                       uint32_t* counters, int sm_count, cudaStream_t stream);

template <class T>
void launch_find_deferred(const IndexView& ix, const T* patterns, const uint32_t* deferred, const uint32_t* counters, size_t k,
                          gbwt_b200_state* out, unsigned grid, cudaStream_t stream);

// Checkpointed path extraction from record windows (find_window.cu: k_extract_window). plan_extract_windows derives its
// plan from the search plan; counters[0] = work ticket (zero on entry).
bool plan_extract_windows(const WindowPlan& search, size_t sequences, WindowPlan& plan);
int launch_extract_window(const IndexView& ix, const CheckpointView& cv, const WindowPlan& plan, const uint64_t* ids, size_t m,
                          const uint64_t* out_offsets, uint64_t base_offset, uint64_t* nodes, uint64_t* lengths, uint32_t* counters, int sm_count,
                          cudaStream_t stream);

// Bidirectional searches from record windows (find_window.cu: k_bd_window, then k_bd_deferred for what it could not
// decide). perm / bucket_end: the batch sorted by the window of path[first]; counters as for launch_find_window.
bool plan_bd_windows(const WindowPlan& search, WindowPlan& plan);
// launch_bd_place is the placement step of the sort for these batches: besides perm[slot] = q it writes packed[slot] =
// {element offset of path[start] in `nodes` (two words), end - start, first - start} (length 0: the search is None), so that
// the window kernel reads one coalesced 16-byte record per search instead of five scattered 8-byte values.
void launch_bd_place(const uint32_t* keys, size_t n, uint32_t* cursor, uint32_t* perm, const uint64_t* offsets, uint64_t base_offset,
                     const uint64_t* first, const uint64_t* start, const uint64_t* end, uint4* packed, unsigned grid, cudaStream_t stream);
// The same as a one-pass placement with plan.direct_cap slots per window (see launch_window_place_direct).
void launch_bd_place_direct(const IndexView& ix, const WindowPlan& plan, const uint64_t* nodes, size_t n, uint32_t* filled, uint32_t* perm,
                            const uint64_t* offsets, uint64_t base_offset, const uint64_t* first, const uint64_t* start, const uint64_t* end,
                            uint4* packed, uint32_t* deferred, uint32_t* counters, unsigned grid, cudaStream_t stream);
int launch_bd_window(const IndexView& ix, const WindowPlan& plan, const uint64_t* nodes, const uint64_t* offsets, uint64_t base_offset,
                     const uint64_t* first, const uint64_t* start, const uint64_t* end, const uint32_t* perm, const uint4* packed,
                     const uint32_t* bucket_end, gbwt_b200_bdstate* out, uint32_t* deferred, uint32_t* counters, int sm_count,
                     cudaStream_t stream);

// Plain (no window) kernels for 32-bit patterns, same dispatch as the 64-bit ones of kernels.cuh.
void launch_find_extend_u32(const IndexView& ix, bool runs, const uint32_t* patterns, const uint32_t* perm, size_t n, size_t k,
                            gbwt_b200_state* out, unsigned grid, cudaStream_t stream);

}  // namespace gbwt_b200
