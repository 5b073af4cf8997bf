// Probe: can two warps of one CTA issue tcgen05.mma streams (different accumulators) at the same time, each with its
// own tcgen05.commit?  Prints cycles for one warp alone and for both together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_dual tools/umma_dual_probe.cu -lcuda
#include <cstdio>
#include "../facerec_b200/csrc/common.cuh"
using namespace frb;

__global__ void __launch_bounds__(96, 1) dual_kernel(int iters, int both, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = smem;                       // 2 x 16 KB
  uint8_t* sB = sA + 2 * 16384;             // 16 KB
  uint64_t* done = reinterpret_cast<uint64_t*>(sB + 16384);   // [2]
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (3 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) { mbar_init(&done[0], 1); mbar_init(&done[1], 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc<512>(slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp < 2 && (both || warp == 0)) {
    const uint32_t idesc = make_idesc_f16(128, 64, 0);
    const uint32_t a0 = desc_lo(smem_u32(sA + warp * 16384)), b0 = desc_lo(smem_u32(sB)), hi = desc_hi(1024);
    const uint32_t d = tmem + warp * 64;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        tc_mma_f16(d, pack_desc(a0, hi), pack_desc(b0, hi), idesc, 1u);
        tc_mma_f16(d, pack_desc(a0 + 2, hi), pack_desc(b0 + 2, hi), idesc, 1u);
        tc_mma_f16(d, pack_desc(a0 + 4, hi), pack_desc(b0 + 4, hi), idesc, 1u);
        tc_mma_f16(d, pack_desc(a0 + 6, hi), pack_desc(b0 + 6, hi), idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit_a(smem_u32(&done[warp]));
    __syncwarp();
    mbar_wait_a(smem_u32(&done[warp]), 0, nullptr, 0);
    if (lane == 0) out[warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_dyn(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 3 * 16384 + 64 + 1024;
  cudaFuncSetAttribute(dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  for (int both : {0, 1}) {
    cudaMemset(d, 0, 16);
    dual_kernel<<<1, 96, smem>>>(iters, both, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("both %d: CUDA error %s\n", both, cudaGetErrorString(e)); return 1; }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("both %d: warp 0 %.1f cycles per MMA, warp 1 %.1f cycles per MMA (own stream each)\n", both, (double)h[0] / (4.0 * iters),
           (double)h[1] / (4.0 * iters));
  }
  return 0;
}
