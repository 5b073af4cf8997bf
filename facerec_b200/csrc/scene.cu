// Scene-cut statistics on the GPU: the per-frame work of facerec/scene.py:81-133 for a whole
// batch of frames already resident in HBM.  Everything that decides a cut is reproduced in float64 with
// the reference's operation order where it matters for a discrete outcome:
//   luminance  Y = (0.299*R + 0.587*G) + 0.114*B   (no FMA; its truncation picks the histogram bin, :42-50, :57-62)
//   histogram  256 integer bins over the centre 2:1 crop (:102-106, :57-59)  -> exact
//   cdf        sequential float64 cumsum of count/N (:59-60)                  -> exact
// The three means (MAFD, MAFD of the equalised frame, frame variance; :111-127) are float64 block
// reductions: they differ from numpy's pairwise sums only in the last bits, far below the thresholds of :65-79.
#include "host.h"
#include "common.cuh"

namespace frb {

__device__ __forceinline__ double luminance(const uint8_t* px) {
  return __dadd_rn(__dadd_rn(__dmul_rn(0.299, static_cast<double>(px[0])), __dmul_rn(0.587, static_cast<double>(px[1]))),
                   __dmul_rn(0.114, static_cast<double>(px[2])));
}

// pass 1: integer histogram of floor(Y) over rows [row0, row1) of every frame
__global__ void scene_hist_kernel(const uint8_t* __restrict__ frames, int H, int W, int row0, int row1,
                                  unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint8_t* f = frames + static_cast<size_t>(b) * H * W * 3 + static_cast<size_t>(row0) * W * 3;
  const long long n = static_cast<long long>(row1 - row0) * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(luminance(f + i * 3));     // truncation, as astype(np.uint8) on values in [0, 255]
    atomicAdd(&sh[k], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[b * 256 + i], sh[i]);
}

// pass 2: cdf[j] = sum_{i<=j} hist[i]/N, sequential float64 (np.cumsum order); one thread per frame
__global__ void scene_cdf_kernel(const unsigned int* __restrict__ hist, int B, double n, double* __restrict__ cdf) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double acc = 0.0;
  for (int i = 0; i < 256; ++i) {
    acc = __dadd_rn(acc, __ddiv_rn(static_cast<double>(hist[b * 256 + i]), n));
    cdf[b * 256 + i] = acc;
  }
}

// pass 3: sum |Y_t - Y_{t-1}| and sum |eq_t - eq_{t-1}| (eq = cdf[floor(Y)] * 255) against the previous frame
// (frame b-1 of the batch, or the carried-over last frame of the previous batch for b == 0)
__global__ void scene_diff_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ prev_frame,
                                  const double* __restrict__ cdf, const double* __restrict__ prev_cdf, int H, int W,
                                  int row0, int row1, double* __restrict__ sums) {
  __shared__ double c_cur[256], c_prev[256];
  __shared__ double red[2][32];
  const int b = blockIdx.y;
  const uint8_t* cur = frames + static_cast<size_t>(b) * H * W * 3 + static_cast<size_t>(row0) * W * 3;
  const uint8_t* prv = (b == 0 ? prev_frame : frames + static_cast<size_t>(b - 1) * H * W * 3) + static_cast<size_t>(row0) * W * 3;
  const double* pc = b == 0 ? prev_cdf : cdf + (b - 1) * 256;
  if (b == 0 && prev_frame == nullptr) return;     // very first frame of the shard: nothing to compare with
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    c_cur[i] = __dmul_rn(cdf[b * 256 + i], 255.0);
    c_prev[i] = __dmul_rn(pc[i], 255.0);
  }
  __syncthreads();
  const long long n = static_cast<long long>(row1 - row0) * W;
  double s0 = 0.0, s1 = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double y = luminance(cur + i * 3), yp = luminance(prv + i * 3);
    s0 += fabs(y - yp);
    s1 += fabs(c_cur[static_cast<int>(y)] - c_prev[static_cast<int>(yp)]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    s0 = lane < nw ? red[0][lane] : 0.0;
    s1 = lane < nw ? red[1][lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      atomicAdd(&sums[b * 2 + 0], s0);
      atomicAdd(&sums[b * 2 + 1], s1);
    }
  }
}

int run_scene_stats(frb_ctx* ctx, const uint8_t* d_frames, int B, int H, int W, int crop, const uint8_t* d_prev_frame,
                    const double* d_prev_cdf, uint32_t* d_hist, double* d_cdf, double* d_sums, cudaStream_t st) {
  if (B <= 0 || B > 65535 || H <= 0 || W <= 0) return set_err(ctx, FRB_ERR_INVALID, "scene_stats: bad shape");
  int row0 = 0, row1 = H;
  if (crop && static_cast<double>(W) / H < 2.0) {           // centre 2:1 crop (facerec/scene.py:102-106)
    const int inset = static_cast<int>((H - 0.5 * W) / 2);
    row0 = inset;
    row1 = H - inset;                                      // numpy img[inset:-inset]; inset == 0 would be empty
    if (inset == 0) row1 = 0;
  }
  if (row1 <= row0) return set_err(ctx, FRB_ERR_INVALID, "scene_stats: empty crop (the reference fails here too)");
  FRB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * 256 * B, st));
  FRB_CUDA(ctx, cudaMemsetAsync(d_sums, 0, sizeof(double) * 2 * B, st));
  const long long n = static_cast<long long>(row1 - row0) * W;
  const int gx = static_cast<int>(std::max(1LL, std::min<long long>((n + 1023) / 1024, std::max(1, ctx->num_sms * 4 / B))));
  scene_hist_kernel<<<dim3(gx, B), 256, 0, st>>>(d_frames, H, W, row0, row1, d_hist);
  scene_cdf_kernel<<<(B + 63) / 64, 64, 0, st>>>(d_hist, B, static_cast<double>(n), d_cdf);
  scene_diff_kernel<<<dim3(gx, B), 256, 0, st>>>(d_frames, d_prev_frame, d_cdf, d_prev_cdf, H, W, row0, row1, d_sums);
  ctx->launches += 3;
  return check_last(ctx, "scene kernels");
}

}  // namespace frb
