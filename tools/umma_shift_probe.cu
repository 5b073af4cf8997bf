// Probe: can a tcgen05 SW128 K-major A descriptor start at an arbitrary 128-byte row of a TMA-written tile, with a
// stride between 8-row groups that is not a multiple of 1024 B?  (Needed to read shifted 3x3-conv windows straight
// out of one halo tile in shared memory.)  B = identity, so D[i][n] tells which A row / chunk row i really read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_probe tools/umma_shift_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../facerec_b200/csrc/common.cuh"

using namespace frb;

constexpr int kPix = 256;   // A rows available in smem
constexpr int kN = 64;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift_rows,
             int sbo_bytes, int base_off, float* __restrict__ out /*[128][64]*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = smem;                       // 256 rows x 128 B
  uint8_t* sB = smem + kPix * 128;          // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kN * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar, 1);
      mbar_init(done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<64>(slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, kPix * 128 + kN * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0, nullptr, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, kN, 0);
    const uint64_t a = desc_sw128(smem_u32(sA) + shift_rows * 128, sbo_bytes, base_off);
    const uint64_t b = desc_sw128(smem_u32(sB), 1024, 0);
    for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(tmem, a + k4 * 2, b + k4 * 2, idesc, k4 ? 1u : 0u);
    tc_commit(done);
  }
  mbar_wait(done, 0, nullptr, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < kN; n0 += 16) {
    uint32_t v[16];
    __syncwarp();
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + n0, v);
    tc_wait_ld();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * kN + n0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem, 64);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 2;
  EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
  std::vector<__half> hA(kPix * 64), hB(kN * 64);
  for (int p = 0; p < kPix; ++p)
    for (int c = 0; c < 64; ++c) hA[p * 64 + c] = __float2half(static_cast<float>(p * 8 + c / 8));   // row id * 8 + 16-byte chunk id (exact in fp16)
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, 128 * kN * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  cuuint64_t gdA[2] = {64, kPix}, gs[1] = {128}, gdB[2] = {64, kN};
  cuuint32_t boxA[2] = {64, kPix}, boxB[2] = {64, kN}, es[2] = {1, 1};
  if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, gdA, gs, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 3;
  if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, gdB, gs, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 3;
  const size_t smem = kPix * 128 + kN * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> hO(128 * kN);
  struct Case { int shift, sbo, bo; };
  std::vector<Case> cases;
  for (int shift : {0, 1, 2, 3, 5, 9, 10, 11})
    for (int sbo : {1024, 1280, 1408})
      for (int bo : {0, -1}) cases.push_back({shift, sbo, bo});
  for (const Case& c : cases) {
    const int bo = c.bo < 0 ? (c.shift & 7) : c.bo;
    if (c.bo < 0 && bo == 0) continue;
    probe_kernel<<<1, 128, smem>>>(tmA, tmB, c.shift, c.sbo, bo, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d sbo %d bo %d: CUDA error %s\n", c.shift, c.sbo, bo, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    int ok_rows = 0, ok_all = 0;
    int first_bad = -1;
    for (int i = 0; i < 128; ++i) {
      const int p = c.shift + (i / 8) * (c.sbo / 128) + (i % 8);
      bool row_ok = true, all_ok = true;
      for (int n = 0; n < kN; ++n) {
        const float want = static_cast<float>(p * 8 + n / 8);
        const float got = hO[i * kN + n];
        if (static_cast<int>(got) / 8 != p) row_ok = false;
        if (got != want) all_ok = false;
      }
      ok_rows += row_ok;
      ok_all += all_ok;
      if (!all_ok && first_bad < 0) first_bad = i;
    }
    printf("shift %2d sbo %4d base_off %d : rows from the right pixel %3d/128, fully exact %3d/128", c.shift, c.sbo, bo, ok_rows, ok_all);
    if (first_bad >= 0) {
      printf("  first bad row %d reads:", first_bad);
      for (int n = 0; n < 64; n += 8) printf(" %d.%d", static_cast<int>(hO[first_bad * kN + n]) / 8, static_cast<int>(hO[first_bad * kN + n]) % 8);
    }
    printf("\n");
  }
  return 0;
}
