// Instantiation + dispatch of conv_tma_kernel over (residual kind, activation kind).  Included by one translation
// unit per (element type, stall accounting) pair so the variants compile in parallel.
#pragma once
#include "conv_tma.cuh"

namespace frb {

template <bool kBF16, bool kStall, int kRes, int kAct, bool kPair, bool kPool = false>
static cudaError_t conv_tma_launch_one(int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a, const CUtensorMap& tmA,
                                       const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR, const CUtensorMap& tmI) {
  auto* kernel = conv_tma_kernel<kBF16, kStall, kRes, kAct, kPair, kPool>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTmaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = a.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (a.cta2) {                          // CTA pairs: clusters of two, one per TPC
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kernel, a, tmA, tmB, tmO, tmR, tmI);
}

// res: ConvRes of the launch; act: 0 per-channel slope, 1 ReLU, 2 linear (an upsampled residual always takes 0)
template <bool kBF16, bool kStall, bool kPair>
static cudaError_t conv_tma_dispatch(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                                     const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR, const CUtensorMap& tmI) {
  if (res == RES_POST_ACT_UPSAMPLE) return conv_tma_launch_one<kBF16, kStall, 2, 0, kPair>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
  if (res == RES_PRE_ACT) {
    if (act == 1) return conv_tma_launch_one<kBF16, kStall, 1, 1, kPair>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
    if (act == 2) return conv_tma_launch_one<kBF16, kStall, 1, 2, kPair>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
    return conv_tma_launch_one<kBF16, kStall, 1, 0, kPair>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
  }
  if (act == 1 && a.pool_out != nullptr) return conv_tma_launch_one<kBF16, kStall, 0, 1, kPair, true>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
  if (act == 1) return conv_tma_launch_one<kBF16, kStall, 0, 1, kPair>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
  if (act == 2) return conv_tma_launch_one<kBF16, kStall, 0, 2, kPair>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
  return conv_tma_launch_one<kBF16, kStall, 0, 0, kPair>(grid, smem, st, a, tmA, tmB, tmO, tmR, tmI);
}

}  // namespace frb
