// Probe: issue rate of tcgen05.mma (cta_group::1, kind::f16, M = 128) from one thread, for N = 64 / 128 / 256,
// with and without a tcgen05.commit per 4 MMAs.  One CTA; operands are whatever is in shared memory.
#include <cstdio>
#include <vector>
#include "../facerec_b200/csrc/common.cuh"
using namespace frb;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int commit_each, int sbo, int vary, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = smem;                 // 64 KB region
  uint8_t* sB = smem + 65536;         // 3 blocks of up to 256 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 3 * 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (65536 + 3 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc<512>(slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, 0);
    const uint64_t a = make_sw128_kmajor_desc_sbo(smem_u32(sA), sbo);
    const uint64_t b = make_sw128_kmajor_desc(smem_u32(sB));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // vary == 1: every MMA reads a different A window and B block, like the 9 taps of a 3x3 conv
      const uint64_t ao = vary ? static_cast<uint64_t>((it % 9) * 8) : 0, bo = vary ? static_cast<uint64_t>((it % 3) * (N * 128 >> 4)) : 0;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(tmem, a + ao + k4 * 2, b + bo + k4 * 2, idesc, 1u);
      if (commit_each) tc_commit(&bars[1]);
    }
    const long long t1 = clock64();
    tc_commit(&bars[0]);
    mbar_wait(&bars[0], 0, nullptr, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_dyn(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 65536 + 3 * 32768 + 64 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int N : {64, 128, 256})
    for (int commit_each : {0, 1})
      for (int vary : {0, 1}) {
        const int sbo = 1280;
        const int iters = 500;
        rate_kernel<<<1, 128, smem>>>(N, iters, commit_each, sbo, vary, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("N %3d commit_each %d vary %d sbo %4d : issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (floor %d)\n", N, commit_each, vary, sbo,
               (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), N / 2);
      }
  return 0;
}
