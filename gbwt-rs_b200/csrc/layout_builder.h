// layout_builder.h -- K0: turns the reference's BWT byte stream into the device layout of layout.h.
// Runs once at load time on the host (parallel over records); the result is copied to HBM verbatim.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "layout.h"
#include "sds_loader.h"

namespace gbwt_b200 {

// Allocator whose resize() leaves trivially constructible elements uninitialised: the big arrays of the layout (GBs for a
// pangenome index) are cleared and filled by all host threads together instead of being zeroed -- and their pages first
// touched -- by the one thread that sizes them.
template <class T>
struct UninitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = UninitAllocator<U>; };
    UninitAllocator() = default;
    template <class U> UninitAllocator(const UninitAllocator<U>&) {}
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        if constexpr (sizeof...(Args) == 0) ::new (static_cast<void*>(p)) U;
        else ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
};
template <class T> using BigVector = std::vector<T, UninitAllocator<T>>;

struct HostLayout {
    BigVector<RecordDesc> desc;
    BigVector<uint64_t> bodies;    // raw 16-byte units, two words each
    std::vector<Edge> edges;       // edge lists of records with sigma > 2
    std::vector<Edge> endmarker;   // Record::decompress() of record 0 (src/gbwt.rs:413-414)
    BigVector<uint64_t> skips;     // two words per record (IndexView::skips)
    std::vector<uint32_t> stage_body;  // body offset of every STAGE_GRANULE-th record (IndexView::stage_body)
    uint64_t edges_total = 0, edges_local = 0;  // edges, and those whose target is within STAGE_LOCAL records
    uint64_t edges_local_span = 0;              // sum over the local edges of the distance to the target, in records
    bool edges_valid = true;
    uint64_t total_length = 0;    // sum of the record lengths (IndexView::walk_limit)
    uint64_t format_counts[FMT_COUNT] = {0, 0, 0, 0, 0, 0, 0};
    uint64_t checkpointed_records = 0;  // run bodies that carry a checkpoint table (layout.h)
};

// `policy` is GBWT_B200_LAYOUT_AUTO or GBWT_B200_LAYOUT_RUNS. Returns a GBWT_B200_* status.
int build_layout(const ParsedGBWT& in, int policy, HostLayout& out, std::string& err);

}  // namespace gbwt_b200
