// Probe: what in the MMA issuer's per-stage loop keeps the tensor pipe from running back to back?
// One CTA.  Warp 0 runs conv_tma's issuer loop shape — warp-wide, warp-uniform descriptor words, the tcgen05 ops under
// elect_one() — over `iters` stages of kPerStage MMAs (M = 128, N given); warp 1 plays the TMA producer (arrives on
// the stage's full barrier after the issuer's commit released it).  Template switches add, per stage:
//   kWait: mbarrier wait on the full barrier      kFence: tcgen05.fence::after_thread_sync
//   kCommit: tcgen05.commit to the empty barrier  kSync: __syncwarp() after the elected block
//   kSpin: both sides poll with mbarrier.test_wait instead of mbarrier.try_wait (which may suspend the thread)
// Output: cycles per MMA against the pipe floor (56 / 64 / 128 for N = 64 / 128 / 256).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_issue tools/umma_issue_probe.cu -lcuda
#include <cstdio>
#include "../facerec_b200/csrc/common.cuh"
using namespace frb;


template <int kPerStage, bool kWait, bool kFence, bool kCommit, bool kSync, bool kSpin, int kStages>
__global__ void __launch_bounds__(64 + 12 * 32, 1) issue_kernel(int N, int iters, int spinners, int one_lane, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = smem;                       // 4 x 32 KB (ring slots alias beyond 4 stages: timing only)
  uint8_t* sB = sA + 4 * 32768;             // 32 KB
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + 32768);   // [kStages]
  uint64_t* empty = full + 16;                                 // [kStages]
  uint64_t* done = empty + 16;
  uint64_t* idle = done + 2;                                   // completes when the issuer is finished (what spinners wait on)
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (5 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
      mbar_init(done, 1);
      mbar_init(idle, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp >= 2) {
    // stand-ins for epilogue warps parked on a barrier (tfull) while the issuer works: all lanes poll, or lane 0 only
    if (warp - 2 < spinners) {
      if (!one_lane || lane == 0) mbar_wait(idle, 0, nullptr, 0);
      __syncwarp();
    }
  } else if (warp == 1) {
    if (kWait && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        if (kCommit) {
          if (kSpin) { while (!mbar_test_wait_a(smem_u32(&empty[stage]), phase ^ 1)) {} }
          else mbar_wait(&empty[stage], phase ^ 1, nullptr, 0);
        }
        mbar_arrive(&full[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const uint32_t idesc = make_idesc_f16(128, N, 0);
    const uint32_t a0 = desc_lo(smem_u32(sA)), b0 = desc_lo(smem_u32(sB)), hi = desc_hi(1024);
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t a_ring = a0;
    int slot = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (kWait) {
        if (kSpin) { while (!mbar_test_wait_a(full0 + stage * 8, phase)) {} }
        else mbar_wait_a(full0 + stage * 8, phase, nullptr, 0);
      }
      if (kFence) tc_fence_after();
      if (elect_one()) {
        tc_mma_f16(tmem, pack_desc(a_ring, hi), pack_desc(b0, hi), idesc, 1u);
        tc_mma_f16(tmem, pack_desc(a_ring + 2, hi), pack_desc(b0 + 2, hi), idesc, 1u);
        tc_mma_f16(tmem, pack_desc(a_ring + 4, hi), pack_desc(b0 + 4, hi), idesc, 1u);
        tc_mma_f16(tmem, pack_desc(a_ring + 6, hi), pack_desc(b0 + 6, hi), idesc, 1u);
        if (kPerStage == 8) {
          tc_mma_f16(tmem, pack_desc(a_ring + 1024, hi), pack_desc(b0 + 1024, hi), idesc, 1u);
          tc_mma_f16(tmem, pack_desc(a_ring + 1026, hi), pack_desc(b0 + 1026, hi), idesc, 1u);
          tc_mma_f16(tmem, pack_desc(a_ring + 1028, hi), pack_desc(b0 + 1028, hi), idesc, 1u);
          tc_mma_f16(tmem, pack_desc(a_ring + 1030, hi), pack_desc(b0 + 1030, hi), idesc, 1u);
        }
        if (kCommit) tc_commit_a(empty0 + stage * 8);
      }
      if (kSync) __syncwarp();
      a_ring += 2048;
      if (++slot == 4) { slot = 0; a_ring = a0; }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0) {
      tc_commit(done);
      mbar_wait(done, 0, nullptr, 0);
      out[0] = clock64() - t0;
      mbar_arrive(idle);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_dyn(tmem, 512); }
}

template <int kPerStage, bool kWait, bool kFence, bool kCommit, bool kSync, bool kSpin, int kStages>
void run(long long* d, const char* what, int spinners = 0, int one_lane = 0) {
  const size_t smem = 5 * 32768 + 512 + 1024;
  auto* k = issue_kernel<kPerStage, kWait, kFence, kCommit, kSync, kSpin, kStages>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4000;
  for (int N : {64, 128}) {
    k<<<1, 64 + 12 * 32, smem>>>(N, iters, spinners, one_lane, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s (%s)\n", cudaGetErrorString(e), what); exit(1); }
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N %3d, %d MMAs per stage, ring of %d, %-42s: %6.1f cycles per MMA, %6.0f per stage\n", N, kPerStage, kStages, what,
           (double)h / ((double)iters * kPerStage), (double)h / iters);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<4, false, false, false, false, false, 4>(d, "bare loop");
  run<4, true, true, true, true, false, 4>(d, "+ ring hand-off");
  run<4, true, true, true, true, false, 4>(d, "+ ring hand-off, 12 warps parked (all)", 12, 0);
  run<4, true, true, true, true, false, 4>(d, "+ ring hand-off, 12 warps parked (1 lane)", 12, 1);
  run<8, true, true, true, true, false, 4>(d, "+ ring hand-off");
  run<8, true, true, true, true, false, 4>(d, "+ ring hand-off, 12 warps parked (all)", 12, 0);
  run<8, true, true, true, true, false, 4>(d, "+ ring hand-off, 12 warps parked (1 lane)", 12, 1);
  return 0;
}
