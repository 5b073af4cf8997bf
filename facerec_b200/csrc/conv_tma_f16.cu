// conv_tma_kernel variants: fp16 operands.
#include "conv_tma_launch.cuh"

namespace frb {

cudaError_t conv_tma_launch_f16(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                               const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR, const CUtensorMap& tmI) {
  return conv_tma_dispatch<false, false, false>(res, act, grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
}

}  // namespace frb
