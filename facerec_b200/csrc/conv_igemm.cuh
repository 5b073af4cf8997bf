// K2 — implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[m, n] = epilogue( sum_k A[m, k] * Wt[n, k] ),   m = (b, oh, ow),  k = (r, s, c),  n = cout
//
// * activations: NHWC 16-bit (fp16 by default, bf16 selectable), arbitrary pixel stride
//   (`in_ld`) so a conv can read a channel slice of a wider concat buffer;
// * weights: [Cout_pad, K_pad] K-major 16-bit, K ordered (r, s, c), fetched by TMA
//   (2D tiled, 128-byte swizzle) straight into the UMMA canonical layout;
// * A operand: either TMA (1x1 / stride 1 convs are plain GEMMs) or an im2col gather by four
//   producer warps using 16-byte cp.async with zero-fill for padding, written in the same
//   128-byte swizzled layout the tensor core expects;
// * accumulators: fp32 in TMEM, double buffered so the epilogue of tile i overlaps the
//   main loop of tile i+1 (persistent CTAs, one per SM);
// * epilogue: bias (BN folded on the host), residual add (pre-activation, or post-activation
//   from a nearest-upsampled coarser map for the FPN top-down path), per-channel negative
//   slope (1 = linear, 0 = ReLU, 0.1 = LeakyReLU; lets fused sibling convs differ),
//   16-bit or fp32 store with an output pixel stride + channel offset (concat for free),
//   or the RetinaFace head scatter into loc / conf / landm.
//
// Warp roles (320 threads): 0-3 epilogue (TMEM lane quarter = warp id), 4 MMA issuer + TMEM
// allocator, 5 TMA producer, 6-9 im2col gather producers.
#pragma once
#include "common.cuh"

namespace frb {

enum ConvAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };
enum ConvRes : int { RES_NONE = 0, RES_PRE_ACT = 1, RES_POST_ACT_UPSAMPLE = 2 };
enum ConvOut : int { OUT_16BIT = 0, OUT_FP32 = 1, OUT_HEAD_SCATTER = 2 };

constexpr int kBM = 128;                 // tile rows (output pixels) == UMMA M
constexpr int kBK = 64;                  // K elements per pipeline stage (one 128 B swizzle row)
constexpr int kAStageBytes = kBM * kBK * 2;
constexpr int kMaxStages = 8;
constexpr int kConvThreads = 320;
constexpr int kGatherLag = 2;            // cp.async groups kept in flight per gather thread

struct ConvArgs {
  const void* in;      // NHWC 16-bit, already offset to the first input channel
  int in_ld;           // elements between consecutive pixels
  int B, H, W, Cin;
  int Ho, Wo, Cout;
  int kh, kw, sh, sw, ph, pw;
  int K;               // kh*kw*Cin
  int num_kb;          // ceil(K / 64)
  int M;               // B*Ho*Wo
  int BN;              // tile columns == UMMA N (multiple of 16, <= 256)
  int n_tiles, total_tiles;
  int stages;          // smem pipeline depth
  int a_tma;           // A operand through TMA (1x1, stride 1, no padding)
  const float* bias;   // [n_tiles*BN]
  void* out;           // already offset to the first output channel
  int out_ld;
  int out_mode;        // ConvOut
  const float* slope;  // [n_tiles*BN] per-channel negative slope: 1 = linear, 0 = ReLU, 0.1 = LeakyReLU
  const void* res;     // residual source (16-bit), already channel-offset
  int res_ld;
  int res_mode;        // ConvRes
  int rH, rW;          // residual map size for RES_POST_ACT_UPSAMPLE
  float* loc;          // OUT_HEAD_SCATTER: [B, P_total, 4] / [.., 2] / [.., 10]
  float* conf;
  float* landm;
  int P_total, P_off;
  int* watchdog;       // mapped pinned host memory, may be null
};

template <bool kBF16>
struct Elem16;
template <>
struct Elem16<false> {
  using T = __half;
  using T2 = __half2;
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t relu2(uint32_t u) {
    __half2 h = __hmax2(*reinterpret_cast<const __half2*>(&u), __float2half2_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    __half2 h = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  // relu2(from_f2(a, b)) in one instruction (rounding is monotonic and keeps the sign: same bits for every non-NaN input)
  static __device__ __forceinline__ uint32_t from_f2_relu(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
};
template <>
struct Elem16<true> {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t relu2(uint32_t u) {
    __nv_bfloat162 h = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&u), __float2bfloat162_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    __nv_bfloat162 h = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t from_f2_relu(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
};

template <bool kBF16>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvArgs p, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA) {
  using E = Elem16<kBF16>;
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte aligned stage bases.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const int b_stage_bytes = p.BN * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.stages * kAStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + p.stages * b_stage_bytes);
  uint64_t* full_bar = bars;                       // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // TMEM columns: two accumulator buffers of BN fp32 columns, power of two >= 32.
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * p.BN)) tmem_cols <<= 1;

  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmB);
    if (p.a_tma) tma_prefetch_desc(&tmA);
  }
  if (warp == 4) {
    if (lane == 0) {
      const uint32_t full_count = p.a_tma ? 1u : 5u;  // TMA thread (+ one lane per gather warp)
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], full_count);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tfull_bar[b], 1);
        mbar_init(&tempty_bar[b], 4);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_dyn(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first_tile = blockIdx.x;
  const int tile_step = gridDim.x;

  if (warp < 4) {
    // ================================================================ epilogue
    const int row = warp * 32 + lane;
    int tile_iter = 0;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++tile_iter) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const int buf = tile_iter & 1;
      const uint32_t tphase = (tile_iter >> 1) & 1;
      mbar_wait(&tfull_bar[buf], tphase, p.watchdog, 100 + buf);
      tc_fence_after();

      const int m = m_tile * kBM + row;
      const bool row_ok = m < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(buf * p.BN);
      const int n_base = n_tile * p.BN;

      // Per-row addressing, computed once per tile.
      size_t out_row = static_cast<size_t>(m) * p.out_ld;
      size_t res_row = 0;
      int hb = 0, hpix = 0;
      if (p.res_mode == RES_PRE_ACT) {
        res_row = static_cast<size_t>(m) * p.res_ld;
      } else if (p.res_mode == RES_POST_ACT_UPSAMPLE && row_ok) {
        const int ow = m % p.Wo;
        const int t = m / p.Wo;
        const int oh = t % p.Ho;
        const int b = t / p.Ho;
        const int rh = (oh * p.rH) / p.Ho;   // nearest: floor(dst * in / out)
        const int rw = (ow * p.rW) / p.Wo;
        res_row = (static_cast<size_t>(b) * p.rH * p.rW + static_cast<size_t>(rh) * p.rW + rw) * p.res_ld;
      }
      if (p.out_mode == OUT_HEAD_SCATTER) {
        const int hw = p.Ho * p.Wo;
        hb = m / hw;
        hpix = m - hb * hw;
      }

      for (int n0 = 0; n0 < p.BN; n0 += 16) {
        uint32_t v[16];
        __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the per-row store branch
        tmem_ld16(taddr + n0, v);
        tc_wait_ld();
        const int n = n_base + n0;
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bz = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
          x[j + 0] = __uint_as_float(v[j + 0]) + bz.x;
          x[j + 1] = __uint_as_float(v[j + 1]) + bz.y;
          x[j + 2] = __uint_as_float(v[j + 2]) + bz.z;
          x[j + 3] = __uint_as_float(v[j + 3]) + bz.w;
        }
        // tcgen05.ld above is warp-collective; everything below is per row
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int ng = n + g * 8;
          if (!row_ok || ng >= p.Cout) break;   // Cout is a multiple of 8 on every 16-bit/fp32 path
          float* y = x + g * 8;
          if (p.res_mode == RES_PRE_ACT) {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(static_cast<const typename E::T*>(p.res) + res_row + ng));
            float2 f;
            f = E::to_f2(r.x); y[0] += f.x; y[1] += f.y;
            f = E::to_f2(r.y); y[2] += f.x; y[3] += f.y;
            f = E::to_f2(r.z); y[4] += f.x; y[5] += f.y;
            f = E::to_f2(r.w); y[6] += f.x; y[7] += f.y;
          }
          {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.slope + ng));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.slope + ng + 4));
            const float sl[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = y[j] > 0.f ? y[j] : y[j] * sl[j];
          }
          if (p.res_mode == RES_POST_ACT_UPSAMPLE) {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(static_cast<const typename E::T*>(p.res) + res_row + ng));
            float2 f;
            f = E::to_f2(r.x); y[0] += f.x; y[1] += f.y;
            f = E::to_f2(r.y); y[2] += f.x; y[3] += f.y;
            f = E::to_f2(r.z); y[4] += f.x; y[5] += f.y;
            f = E::to_f2(r.w); y[6] += f.x; y[7] += f.y;
          }
          if (p.out_mode == OUT_16BIT) {
            uint4 o;
            o.x = E::from_f2(y[0], y[1]);
            o.y = E::from_f2(y[2], y[3]);
            o.z = E::from_f2(y[4], y[5]);
            o.w = E::from_f2(y[6], y[7]);
            *reinterpret_cast<uint4*>(static_cast<typename E::T*>(p.out) + out_row + ng) = o;
          } else if (p.out_mode == OUT_FP32) {
            float* o = static_cast<float*>(p.out) + out_row + ng;
            *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
          } else {
            // Fused head columns per pixel: [0,4) class logits (2 anchors x 2), [4,12) box
            // deltas (2 x 4), [12,32) landmark deltas (2 x 10).  Prior index = P_off + 2*pix + anchor.
            const size_t prior0 = static_cast<size_t>(hb) * p.P_total + p.P_off + 2 * static_cast<size_t>(hpix);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = ng + j;
              if (c < 4) p.conf[prior0 * 2 + c] = y[j];
              else if (c < 12) p.loc[prior0 * 4 + (c - 4)] = y[j];
              else if (c < 32) p.landm[prior0 * 10 + (c - 12)] = y[j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  } else if (warp == 4) {
    // ============================================================== MMA issuer
    const uint32_t idesc = make_idesc_f16(kBM, p.BN, kBF16 ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int tile_iter = 0;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++tile_iter) {
      const int buf = tile_iter & 1;
      const uint32_t tphase = (tile_iter >> 1) & 1;
      mbar_wait(&tempty_bar[buf], tphase ^ 1, p.watchdog, 200 + buf);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * p.BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase, p.watchdog, 210 + stage);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t a_desc = make_sw128_kmajor_desc(smem_u32(smem_a + stage * kAStageBytes));
          const uint64_t b_desc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * b_stage_bytes));
#pragma unroll
          for (int k4 = 0; k4 < kBK / 16; ++k4) {
            // advancing 16 K-elements = 32 bytes inside the swizzle row = +2 in the (addr >> 4) field
            tc_mma_f16(d_tmem, a_desc + static_cast<uint64_t>(k4 * 2), b_desc + static_cast<uint64_t>(k4 * 2), idesc,
                       (kb | k4) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (kb == p.num_kb - 1) tc_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ============================================================ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = static_cast<uint32_t>(b_stage_bytes + (p.a_tma ? kAStageBytes : 0));
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, p.watchdog, 300 + stage);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_2d(smem_b + stage * b_stage_bytes, &tmB, &full_bar[stage], kb * kBK, n_tile * p.BN);
          if (p.a_tma) tma_load_2d(smem_a + stage * kAStageBytes, &tmA, &full_bar[stage], kb * kBK, m_tile * kBM);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (!p.a_tma) {
    // ========================================================= im2col gather
    // 128 threads: thread g owns 16-byte chunk q = g % 8 of rows r0 + 16*i (i = 0..7), so
    // 8 consecutive threads read one pixel's 128 contiguous bytes and fill one smem row.
    const int g = threadIdx.x - 192;
    const int q = g & 7;
    const int r0 = g >> 3;
    const uint32_t dst_off = static_cast<uint32_t>((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((q ^ (r0 & 7)) << 4));
    const typename E::T* in = static_cast<const typename E::T*>(p.in);

    int stage = 0;
    uint32_t phase = 0;
    int arr_stage = 0;       // next stage whose "full" arrival is still owed by this warp
    int in_flight = 0;       // committed cp.async groups not yet signalled
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
      const int m_tile = tile / p.n_tiles;
      int ih0[8], iw0[8], pix[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m_tile * kBM + r0 + 16 * i;
        if (m < p.M) {
          const int ow = m % p.Wo;
          const int t = m / p.Wo;
          const int oh = t % p.Ho;
          const int b = t / p.Ho;
          ih0[i] = oh * p.sh - p.ph;
          iw0[i] = ow * p.sw - p.pw;
          pix[i] = b * p.H * p.W;
        } else {
          ih0[i] = -(1 << 20);   // fails the bounds test for every tap -> zero fill
          iw0[i] = 0;
          pix[i] = 0;
        }
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, p.watchdog, 400 + stage);
        const int k = kb * kBK + q * 8;
        const bool k_ok = k < p.K;
        const int tap = k / p.Cin;
        const int c = k - tap * p.Cin;
        const int r = tap / p.kw;
        const int s = tap - r * p.kw;
        const uint32_t dst = smem_base + static_cast<uint32_t>(stage * kAStageBytes) + dst_off;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ih = ih0[i] + r;
          const int iw = iw0[i] + s;
          const bool ok = k_ok && static_cast<unsigned>(ih) < static_cast<unsigned>(p.H) &&
                          static_cast<unsigned>(iw) < static_cast<unsigned>(p.W);
          const typename E::T* src =
              ok ? in + (static_cast<size_t>(pix[i] + ih * p.W + iw) * p.in_ld + c) : in;
          cp_async16_zfill(dst + i * 2048, src, ok ? 16u : 0u);
        }
        cp_async_commit();
        ++in_flight;
        if (in_flight > kGatherLag) {
          cp_async_wait<kGatherLag>();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[arr_stage]);
          if (++arr_stage == p.stages) arr_stage = 0;
          --in_flight;
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    // drain: everything still in flight lands, then the owed arrivals are paid in order
    cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncwarp();
    while (in_flight > 0) {
      if (lane == 0) mbar_arrive(&full_bar[arr_stage]);
      if (++arr_stage == p.stages) arr_stage = 0;
      --in_flight;
    }
  }

  // -------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, tmem_cols);
  }
}

}  // namespace frb
