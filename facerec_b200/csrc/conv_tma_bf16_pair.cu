// conv_tma_kernel variants, CTA pairs (cta_group::2): bf16 operands.
#include "conv_tma_launch.cuh"

namespace frb {

cudaError_t conv_tma_launch_bf16_pair(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                               const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR, const CUtensorMap& tmI) {
  return conv_tma_dispatch<true, false, true>(res, act, grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
}

}  // namespace frb
