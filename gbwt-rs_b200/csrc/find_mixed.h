// find_mixed.h -- launchers of the lean find/extend kernels for indexes that also hold other records than
// single-edge and dense ones (find_mixed.cu, its own translation unit; see find_lean.cuh).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/gbwt_b200.h"
#include "layout.h"

namespace gbwt_b200 {

void launch_find_extend_lean_mixed(const IndexView& ix, const uint64_t* patterns, const uint32_t* perm, size_t n, size_t k,
                                   gbwt_b200_state* out, unsigned grid, cudaStream_t stream);
void launch_find_extend_ragged_lean_mixed(const IndexView& ix, const uint64_t* nodes, const uint64_t* offsets, uint64_t base,
                                          const uint32_t* perm, size_t n, gbwt_b200_state* out, unsigned grid, cudaStream_t stream);

}  // namespace gbwt_b200
