// Context lifecycle, error plumbing and the host side of K2 (conv layer preparation + launch).
#include <stdio.h>
#include <string.h>
#include <algorithm>

#include "host.h"
#include "conv_igemm.cuh"
#include "conv_tma.cuh"

namespace frb {

int set_err(frb_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx != nullptr) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    ctx->err = buf;
  }
  return code;
}

static void append_watchdog(frb_ctx* ctx) {
  if (ctx != nullptr && ctx->wd_host != nullptr && ctx->wd_host[0] != 0) {
    char buf[384];
    const int* pr = ctx->wd_host + 16 + 8 * (ctx->wd_host[1] >= 0 && ctx->wd_host[1] < 148 ? ctx->wd_host[1] : 0);
    snprintf(buf, sizeof(buf), " [pipeline watchdog: wait code %d in block %d; its progress (stall-counter build only): issuer A %d, "
             "issuer B %d, producer %d, epilogue groups %d %d %d]", ctx->wd_host[0], ctx->wd_host[1], pr[0], pr[1], pr[2], pr[3], pr[4], pr[5]);
    ctx->err += buf;
    ctx->wd_host[0] = 0;
  }
}

int check_cuda(frb_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return FRB_OK;
  set_err(ctx, FRB_ERR_CUDA, "CUDA error %s (%s) at: %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
  append_watchdog(ctx);
  return FRB_ERR_CUDA;
}

int check_last(frb_ctx* ctx, const char* what) { return check_cuda(ctx, cudaGetLastError(), what); }

int ctx_scratch(frb_ctx* ctx, int slot, size_t bytes, void** out) {
  if (slot < 0 || slot >= frb_ctx::kScratchSlots) return set_err(ctx, FRB_ERR_INVALID, "bad scratch slot %d", slot);
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) {
      FRB_CUDA(ctx, cudaDeviceSynchronize());   // the old buffer may still be in use by queued work
      FRB_CUDA(ctx, cudaFree(ctx->scratch[slot]));
      ctx->scratch[slot] = nullptr;
      ctx->scratch_bytes[slot] = 0;
    }
    const size_t want = bytes + bytes / 8 + 256;
    FRB_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], want));
    ctx->scratch_bytes[slot] = want;
  }
  *out = ctx->scratch[slot];
  return FRB_OK;
}

// 64 x 64 identity in the activation type: B operand of the MMAs that add a residual slab into the accumulator
int ctx_identity(frb_ctx* ctx, void** out) {
  void*& ident = ctx->d_ident[ctx->use_bf16 ? 1 : 0];
  if (ident == nullptr) {
    std::vector<uint16_t> h(64 * 64, 0);
    for (int i = 0; i < 64; ++i) h[i * 64 + i] = ctx->use_bf16 ? 0x3F80 : 0x3C00;   // 1.0
    FRB_CUDA(ctx, cudaMalloc(&ident, h.size() * 2));
    FRB_CUDA(ctx, cudaMemcpy(ident, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  *out = ident;
  return FRB_OK;
}

// ----------------------------------------------------------------------------- conv host side
static int pick_bn(int cout_pad16, int cap) {
  cap = std::max(16, std::min(256, cap)) & ~15;
  if (cout_pad16 <= cap) return cout_pad16;
  // several N tiles: keep tiles a multiple of the 64-column store slab
  for (int bn = cap & ~63; bn >= 64; bn -= 64)
    if (cout_pad16 % bn == 0) return bn;
  for (int bn = cap; bn >= 16; bn -= 16)
    if (cout_pad16 % bn == 0 && bn >= cap / 2) return bn;
  return cap;  // no good divisor: pad Cout up to a multiple of cap
}

static uint16_t f32_to_16(float f, int bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  }
  __half h = __float2half(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

int encode_2d(frb_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
              uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode_tiled(tm, ctx->use_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                                 const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(ctx, FRB_ERR_CUDA,
                   "cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r,
                   base, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes,
                   box_inner, box_outer);
  return FRB_OK;
}

int conv_prepare(frb_ctx* ctx, ConvLayer& L, const float* h_w, const float* h_bias, const float* h_slope) {
  if (L.Cin % 8 != 0) return set_err(ctx, FRB_ERR_INVALID, "conv: Cin=%d must be a multiple of 8", L.Cin);
  L.K = L.kh * L.kw * L.Cin;
  L.K_pad = (L.K + kBK - 1) / kBK * kBK;
  const int cout16 = (L.Cout + 15) / 16 * 16;
  L.BN = pick_bn(cout16, L.bn_cap > 0 ? std::min(L.bn_cap, ctx->bn_cap) : ctx->bn_cap);
  L.n_tiles = (cout16 + L.BN - 1) / L.BN;
  L.Cout_pad = L.n_tiles * L.BN;

  std::vector<uint16_t> w16(static_cast<size_t>(L.Cout_pad) * L.K_pad, 0);
  for (int n = 0; n < L.Cout; ++n)
    for (int k = 0; k < L.K; ++k)
      w16[static_cast<size_t>(n) * L.K_pad + k] = f32_to_16(h_w[static_cast<size_t>(n) * L.K + k], ctx->use_bf16);
  std::vector<float> bias(L.Cout_pad, 0.f);
  if (h_bias != nullptr)
    for (int n = 0; n < L.Cout; ++n) bias[n] = h_bias[n];

  std::vector<float> slope(L.Cout_pad, 1.f);
  if (h_slope != nullptr)
    for (int n = 0; n < L.Cout; ++n) slope[n] = h_slope[n];
  bool all0 = true, all1 = true;
  for (int n = 0; n < L.Cout; ++n) { all0 = all0 && slope[n] == 0.f; all1 = all1 && slope[n] == 1.f; }
  L.act_kind = all0 ? 1 : (all1 ? 2 : 0);

  FRB_CUDA(ctx, cudaMalloc(&L.d_w, w16.size() * 2));
  FRB_CUDA(ctx, cudaMalloc(&L.d_bias, bias.size() * 4));
  FRB_CUDA(ctx, cudaMalloc(&L.d_slope, slope.size() * 4));
  FRB_CUDA(ctx, cudaMemcpy(L.d_w, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
  FRB_CUDA(ctx, cudaMemcpy(L.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  FRB_CUDA(ctx, cudaMemcpy(L.d_slope, slope.data(), slope.size() * 4, cudaMemcpyHostToDevice));
  L.has_tmB2 = 0;
  if (L.BN % 16 == 0 && L.BN >= 32) {
    int rc2 = encode_2d(ctx, &L.tmB2, L.d_w, L.K_pad, L.Cout_pad, static_cast<uint64_t>(L.K_pad) * 2, kBK, L.BN / 2);
    if (rc2 != FRB_OK) return rc2;
    L.has_tmB2 = 1;
  }
  return encode_2d(ctx, &L.tmB, L.d_w, L.K_pad, L.Cout_pad, static_cast<uint64_t>(L.K_pad) * 2, kBK, L.BN);
}

void conv_release(ConvLayer& L) {
  if (L.d_w) cudaFree(L.d_w);
  if (L.d_bias) cudaFree(L.d_bias);
  if (L.d_slope) cudaFree(L.d_slope);
  if (L.d_bias_skip) cudaFree(L.d_bias_skip);
  L.d_bias_skip = nullptr;
  L.d_w = nullptr;
  L.d_bias = nullptr;
  L.d_slope = nullptr;
}

typedef CUresult (*EncodeTiledFnT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// rank-n tiled map over 16-bit elements, 128-byte swizzle; strides[i] is the byte stride of dim i+1
static int encode_nd(frb_ctx* ctx, CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, const uint32_t* estr) {
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides[i];
  CUresult r = ctx->encode_tiled(tm, ctx->use_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                 static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, b, e,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(ctx, FRB_ERR_CUDA, "cuTensorMapEncodeTiled(rank %d) failed (%d): dims %llu %llu %llu %llu box %u %u %u %u",
                   rank, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1],
                   rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return FRB_OK;
}

// NHWC activation slice [B,H,W,C] (pixel stride ld) as a 4-D map with a patch box
static int encode_patch(frb_ctx* ctx, CUtensorMap* tm, const void* base, int C, int W, int H, int B, int ld, int bw, int bh,
                        int bb, int sw, int sh) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
  const uint32_t box[4] = {(uint32_t)kSlab, (uint32_t)((bw - 1) * sw + 1), (uint32_t)((bh - 1) * sh + 1), (uint32_t)bb};
  const uint32_t est[4] = {1, (uint32_t)sw, (uint32_t)sh, 1};
  return encode_nd(ctx, tm, base, 4, dims, str, box, est);
}

// Spatial patch of <= 128 output pixels that wastes the fewest tile slots.
static void pick_patch(int B, int Ho, int Wo, int sw, int sh, int& bw, int& bh, int& bb) {
  double best = -1.0;
  bw = bh = bb = 1;
  for (int w = 1; w <= std::min(Wo, 128); ++w) {
    if ((w - 1) * sw + 1 > 256) break;
    for (int h = 1; h <= std::min(Ho, 128 / w); ++h) {
      if ((h - 1) * sh + 1 > 256) break;
      const int b = std::max(1, std::min(B, 128 / (w * h)));
      const double tiles = static_cast<double>((Wo + w - 1) / w) * ((Ho + h - 1) / h) * ((B + b - 1) / b);
      const double eff = static_cast<double>(B) * Ho * Wo / (tiles * 128.0) + 1e-6 * w;
      if (eff > best) { best = eff; bw = w; bh = h; bb = b; }
    }
  }
}

// (multiplier, shift) with n / d == (umulhi(n, mul) + n) >> shift for every 0 <= n < 2^31 (Granlund-Montgomery)
static void fastdiv_make(int d, uint32_t* fd) {
  uint32_t l = 0;
  while ((1ull << l) < static_cast<uint64_t>(d)) ++l;
  fd[0] = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - static_cast<uint64_t>(d))) / static_cast<uint64_t>(d) + 1);
  fd[1] = l;
}

static bool tma_path_ok(frb_ctx* ctx, const ConvLayer& L, const ConvIO& io) {
  const bool linear = L.kh == 1 && L.kw == 1 && L.sh == 1 && L.sw == 1 && L.ph == 0 && L.pw == 0;
  // the fused RetinaFace heads (32 linear columns) leave through the fp32 head epilogue of the TMA kernel
  const bool head = io.out_mode == OUT_HEAD_SCATTER && linear && L.Cout == 32 && L.BN == 32 && L.n_tiles == 1 &&
                    L.act_kind == 2 && io.res_mode == RES_NONE && (L.Cin % kBK) == 0 && (io.P_total % 2) == 0 && (io.P_off % 2) == 0;
  if (io.pool_out != nullptr) return true;                          // conv_launch_tma checks the shape it needs
  if (ctx->force_gather || (io.out_mode != OUT_16BIT && !head)) return false;
  if (!linear && (L.Cin % kBK) != 0) return false;                 // taps are fetched as whole 64-channel blocks
  if (!linear && io.res_mode == RES_POST_ACT_UPSAMPLE) return false;
  if (L.n_tiles > 1 && (L.BN % kSlab) != 0) return false;          // store slabs must not spill into the next N tile
  if (io.res_mode == RES_PRE_ACT && ((io.res_ld % 8) != 0 || (reinterpret_cast<uintptr_t>(io.res) & 15) != 0))
    return false;                                                    // the residual is a TMA operand too
  return true;
}

static int conv_launch_tma(frb_ctx* ctx, const ConvLayer& L, const ConvIO& io, cudaStream_t stream) {
  ConvTmaArgs a;
  memset(&a, 0, sizeof(a));
  // stem over the space-to-depth input: io.H x io.W is the padded s2d grid [H/2, W/2 + 3], every output
  // pixel reads a 4-row window of 4 consecutive s2d pixels (64 contiguous elements per row)
  const int Ho = L.special ? io.H : conv_out_dim(io.H, L.kh, L.sh, L.ph);
  const int Wo = L.special ? io.W - 3 : conv_out_dim(io.W, L.kw, L.sw, L.pw);
  const long long M = static_cast<long long>(io.B) * Ho * Wo;
  if (M <= 0 || M > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "conv: bad problem size");
  const bool linear = L.kh == 1 && L.kw == 1 && L.sh == 1 && L.sw == 1 && L.ph == 0 && L.pw == 0;
  a.patch = linear ? 0 : 1;
  a.kw = L.kw; a.sh = L.sh; a.sw = L.sw; a.ph = L.ph; a.pw = L.pw;
  a.cin_blocks = (L.Cin + kBK - 1) / kBK;
  a.num_kb = L.kh * L.kw * a.cin_blocks;
  a.BN = L.BN; a.n_tiles = L.n_tiles;
  fastdiv_make(a.n_tiles, a.fd_n);
  a.bias = L.d_bias; a.slope = L.d_slope; a.act_kind = L.act_kind;
  a.res_mode = io.res_mode; a.res = io.res; a.res_ld = io.res_ld; a.rH = io.rH; a.rW = io.rW;
  a.Ho = Ho; a.Wo = Wo; a.M = static_cast<int>(M); a.B = io.B;
  a.watchdog = ctx->wd_dev;
  a.stall = ctx->trace_mode != 0 ? ctx->trace_dev : nullptr;
  a.skip_loads = ctx->skip_loads;
  a.dual_issue = ctx->dual_issue;
  {
    // Programmatic dependent launch pays where launches are short: InceptionResnetV1's 105 launches of 15-35 us gain
    // 2.6 us each, the 0.3-1.5 ms launches of RetinaFace-R50 at 1080p lose ~5 us each, with either trigger placement
    // (profiles/r01_pdl_ab.txt).  Estimated duration: the larger of the tensor time and the HBM time of the launch.
    const double flop = 2.0 * static_cast<double>(M) * L.Cout * L.kh * L.kw * L.Cin;
    const double bytes = 2.0 * static_cast<double>(M) * (L.Cout + static_cast<double>(L.Cin) / (L.sh * L.sw));
    const double est_us = std::max(flop / 6.0e8, bytes / 5.0e6);
    a.pdl = !ctx->layer_times && (ctx->pdl == 2 || (ctx->pdl == 1 && est_us < 100.0));
  }
  if (ctx->trace_dev) cudaMemsetAsync(ctx->trace_dev, 0, 8, stream);

  CUtensorMap tmA, tmO, tmR, tmI;
  memset(&tmI, 0, sizeof(tmI));
  if (io.res_mode == RES_PRE_ACT) {
    // 64 x 64 identity in the activation type: B operand of the MMAs that add the residual into the accumulator
    void* ident = nullptr;
    int irc = ctx_identity(ctx, &ident);
    if (irc == FRB_OK) irc = encode_2d(ctx, &tmI, ident, 64, 64, 128, 64, 64);
    if (irc != FRB_OK) return irc;
  }
  int rc;
  long long m_tiles;
  if (linear) {
    m_tiles = (M + kBM - 1) / kBM;
    a.a_bytes = kAStageBytes;
    a.bw = kBM; a.bh = 1; a.bb = 1; a.tiles_w = 1; a.tiles_h = 1;
    if ((rc = encode_2d(ctx, &tmA, io.in, L.Cin, M, static_cast<uint64_t>(io.in_ld) * 2, kBK, kBM)) != FRB_OK) return rc;
    if (io.out_mode == OUT_HEAD_SCATTER) {
      a.head_loc = io.loc; a.head_conf = io.conf; a.head_landm = io.landm;
      a.head_P_total = io.P_total; a.head_P_off = io.P_off; a.head_hw = Ho * Wo;
      tmO = tmA;                                   // never used: the head epilogue stores from registers
    } else if ((rc = encode_2d(ctx, &tmO, io.out, L.Cout, M, static_cast<uint64_t>(io.out_ld) * 2, kSlab, kBM)) != FRB_OK) return rc;
    tmR = tmO;
    if (io.res_mode == RES_PRE_ACT &&
        (rc = encode_2d(ctx, &tmR, io.res, L.Cout, M, static_cast<uint64_t>(io.res_ld) * 2, kSlab, kBM)) != FRB_OK)
      return rc;
  } else {
    pick_patch(io.B, Ho, Wo, L.sw, L.sh, a.bw, a.bh, a.bb);
    // halo mode: stride-1 k x k conv read from one halo tile per 64-channel block (patch fixed at 8 wide x 16 tall),
    // including the s2d stem's 4 x 1 window conv.  What it buys is fewer barrier waits in the MMA issuer (each costs the
    // tensor pipe ~140 cycles whatever the ring depth, tools/umma_issue_probe.cu) and k*k-fold less TMA traffic:
    //  * resident weights: the halo tile is all that has to arrive, every tap of a channel block goes out under ONE
    //    wait (3x3 64->64: 4 924 -> 2 637 cycles per tile, stem 2.59 -> 1.91 ms inside the forward) - default;
    //  * streamed weights, a whole filter row of taps per B stage (fits for BN <= 128 with 3-wide filters): one wait
    //    per row (3x3 128->128: 12 949 -> 7 974 cycles per tile) - default when the 8 x 16 patch wastes little;
    //  * streamed weights, one tap per stage (BN > 128): a wash, on request only.
    // FRB_OPT_HALO = 1 forces it wherever it is possible, -1 turns it off.
    if (ctx->halo_mode >= 0 && io.res_mode != RES_PRE_ACT && L.kh * L.kw > 1 && L.sh == 1 && L.sw == 1 && (L.Cin % kBK) == 0) {
      // The per-tap path adds taps-major, the halo paths channel-block-major: which one a layer takes must not depend
      // on the batch size, or a frame's bits would depend on its batch.  The patch efficiency the halo tiles are
      // compared with is therefore the one of a reference batch of 64 images.
      constexpr int kRefBatch = 64;
      int rw, rh, rb;
      pick_patch(kRefBatch, Ho, Wo, L.sw, L.sh, rw, rh, rb);
      const double eff_patch = static_cast<double>(kRefBatch) * Ho * Wo /
                               (128.0 * ((Wo + rw - 1) / rw) * ((Ho + rh - 1) / rh) * ((kRefBatch + rb - 1) / rb));
      const double eff_halo = static_cast<double>(Ho) * Wo / (128.0 * ((Wo + 7) / 8) * ((Ho + 15) / 16));
      const int halo_bytes = ((8 + L.kw - 1) * (16 + L.kh - 1) * 128 + 1023) & ~1023;
      const bool resident_w = L.n_tiles * L.kh * L.kw * ((L.Cin + kBK - 1) / kBK) * L.BN * 128 <= kResidentWeightBytes;
      const bool row_fits = L.kw > 1 && 2 * L.kw * L.BN * 128 + 2 * halo_bytes <=
                                            226 * 1024 - (kOutBufs * kSlabBytes + (4 * kMaxStages + 20) * 8 + 2048 + 1024);
      const bool want = ctx->halo_mode > 0 || resident_w || (row_fits && eff_halo >= 0.9 * eff_patch);
      // streamed weights, TWO 8 x 16 patches side by side per tile (one (16 + kw - 1)-wide halo tile, two accumulators):
      // every weight stage (one tap) feeds two M = 128 MMAs, i.e. half the weight bytes through L2 per output pixel and
      // twice the MMAs per barrier wait.  Pays for N tiles of 192 / 256 (FPN merge1 1.98 -> 1.85 ms, SSH1 3x3 1.73 ->
      // 1.59 ms); with N = 128 it replaces the row-per-stage path (12 MMAs per wait) by 8 per wait and loses 8 %
      // (profiles/r01_pdl_ab.txt).
      const double eff_pair = static_cast<double>(Ho) * Wo / (256.0 * ((Wo + 15) / 16) * ((Ho + 15) / 16));
      const bool pair = !resident_w && (ctx->mpair == 2 || (ctx->mpair == 1 && L.BN >= 192)) && eff_pair >= 0.85 * eff_patch;
      if (pair) {
        a.halo = 1;
        a.mpair = 1;
        a.bw = 16; a.bh = 16; a.bb = 1;
        a.kh = L.kh;
        a.halo_w = 16 + L.kw - 1;
        a.halo_h = 16 + L.kh - 1;
        a.halo_bytes = (a.halo_w * a.halo_h * 128 + 1023) & ~1023;
      } else if (want && eff_halo >= 0.85 * eff_patch) {
        a.halo = 1;
        a.bw = 8; a.bh = 16; a.bb = 1;
        a.kh = L.kh;
        a.halo_w = 8 + L.kw - 1;
        a.halo_h = 16 + L.kh - 1;
        a.halo_bytes = halo_bytes;
      }
    }
    a.tiles_w = (Wo + a.bw - 1) / a.bw;
    a.tiles_h = (Ho + a.bh - 1) / a.bh;
    m_tiles = static_cast<long long>(a.tiles_w) * a.tiles_h * ((io.B + a.bb - 1) / a.bb);
    a.a_bytes = a.bw * a.bh * a.bb * 128;
    if (L.special) {
      // overlapping windows: dim0 = 64 elements starting at s2d pixel w, dim1 steps one s2d pixel (16 elements)
      const uint64_t dims[4] = {64, (uint64_t)Wo, (uint64_t)io.H, (uint64_t)io.B};
      const uint64_t str[3] = {32, (uint64_t)io.W * 32, (uint64_t)io.H * io.W * 32};
      const uint32_t box[4] = {64, (uint32_t)(a.halo ? a.halo_w : a.bw), (uint32_t)(a.halo ? a.halo_h : a.bh), (uint32_t)a.bb};
      const uint32_t est[4] = {1, 1, 1, 1};
      if ((rc = encode_nd(ctx, &tmA, io.in, 4, dims, str, box, est)) != FRB_OK) return rc;
    } else if (a.halo) {
      if ((rc = encode_patch(ctx, &tmA, io.in, L.Cin, io.W, io.H, io.B, io.in_ld, a.halo_w, a.halo_h, 1, 1, 1)) != FRB_OK) return rc;
    } else if ((rc = encode_patch(ctx, &tmA, io.in, L.Cin, io.W, io.H, io.B, io.in_ld, a.bw, a.bh, a.bb, L.sw, L.sh)) != FRB_OK) return rc;
    if (io.pool_out != nullptr) {
      // fused max pool (3x3 / 2 / pad 1): needs the 8 x 16 halo patches, one N tile, ReLU, no residual
      if (!a.halo || a.mpair || a.bw != 8 || a.bh != 16 || a.bb != 1 || L.n_tiles != 1 || io.res_mode != RES_NONE || L.act_kind != 1 ||
          (io.pool_ld % 8) != 0 || (reinterpret_cast<uintptr_t>(io.pool_out) & 15) != 0)
        return set_err(ctx, FRB_ERR_INVALID, "conv: this layer cannot take the fused max pool");
      a.pool_out = io.pool_out; a.pool_ld = io.pool_ld;
      a.pool_Ho = conv_out_dim(Ho, 3, 2, 1); a.pool_Wo = conv_out_dim(Wo, 3, 2, 1);
      tmO = tmA;                                   // never used: nothing but the pooled tensor is written
    } else if ((rc = encode_patch(ctx, &tmO, io.out, L.Cout, Wo, Ho, io.B, io.out_ld, a.mpair ? 8 : a.bw, a.bh, a.bb, 1, 1)) != FRB_OK) return rc;
    tmR = tmO;
    if (io.res_mode == RES_PRE_ACT &&
        (rc = encode_patch(ctx, &tmR, io.res, L.Cout, Wo, Ho, io.B, io.res_ld, a.bw, a.bh, a.bb, 1, 1)) != FRB_OK)
      return rc;
  }
  if (m_tiles * L.n_tiles > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "conv: too many tiles");
  a.total_tiles = static_cast<int>(m_tiles * L.n_tiles);
  fastdiv_make(a.tiles_w, a.fd_w);
  fastdiv_make(a.tiles_h, a.fd_h);

  // small weight matrices (the bandwidth-bound early layers) stay resident in shared memory for the whole kernel
  const int w_bytes = L.n_tiles * a.num_kb * L.BN * 128;
  a.b_resident = w_bytes <= kResidentWeightBytes ? 1 : 0;
  // ... or just the N tile a CTA works on: with a grid that is a multiple of n_tiles, tile = blockIdx + k * grid keeps
  // tile % n_tiles fixed per CTA, and the weight slice is fetched once instead of once per tile (the 1x1 expansions of
  // layer2/3 move 1/3 fewer bytes through L2)
  const int slice_bytes = a.num_kb * L.BN * 128;
  if (!a.b_resident && !a.halo && ctx->weight_slice && L.n_tiles > 1 && slice_bytes <= kResidentSliceBytes &&
      a.total_tiles >= 4 * ctx->num_sms && L.n_tiles <= ctx->num_sms / 4)
    a.b_resident = 2;
  // CTA pairs: every path but the HBM-bound 1x1 layers with resident weights or weight slices.  Two M tiles per pair and
  // N tile, half of each weight block (streamed or resident) per CTA; results are the same bits
  // (a tile's k-blocks accumulate in the same order), so the choice may depend on the amount of work.
  a.cta2 = (ctx->cta2 && a.stall == nullptr && a.b_resident != 2 && (a.halo || !a.b_resident) && L.has_tmB2 && m_tiles >= 2 &&
            (io.res_mode != RES_PRE_ACT || (L.BN % kSlab) == 0) && io.out_mode != OUT_HEAD_SCATTER &&
            ((m_tiles + 1) / 2) * L.n_tiles >= ctx->num_sms / 2) ? 1 : 0;
  if (a.cta2) a.total_tiles = static_cast<int>(((m_tiles + 1) / 2) * L.n_tiles);
  const int b_rows = a.cta2 ? L.BN / 2 : L.BN;
  const int res_bytes = (a.b_resident == 2 ? slice_bytes : w_bytes) / (a.cta2 ? 2 : 1);
  const int base_fixed = kOutBufs * kSlabBytes + (a.b_resident ? res_bytes : 0) + (io.res_mode == RES_PRE_ACT ? kIdentBytes : 0) +
                         (4 * kMaxStages + 20) * 8 + 2048 + 1024;
  const int total_budget = 226 * 1024 - base_fixed;
  // residual slabs ride the ring as single extra A blocks: one k-block per stage then
  a.kbps = (!a.halo && a.num_kb >= 2 && ctx->kbps2 && io.res_mode != RES_PRE_ACT) ? 2 : 1;
  // halo mode with streamed weights: a whole filter row of taps per B stage when two such stages and two halo tiles fit
  a.brow = 1;
  if (a.halo && !a.b_resident && L.kw > 1 && 2 * L.kw * L.BN * 128 + 2 * a.halo_bytes <= total_budget) a.brow = L.kw;
  int stage_bytes = (a.halo ? 0 : a.kbps * kAStageBytes) + (a.b_resident ? 0 : (a.halo ? a.brow : a.kbps) * b_rows * 128);
  if (a.kbps == 2 && total_budget / stage_bytes < ctx->kbps2_min_stages) {   // too few double stages to cover the TMA latency
    a.kbps = 1;
    stage_bytes /= 2;
  }
  if (a.halo) {
    // TMA latency is hidden by bytes in flight: with resident weights the whole budget goes to halo buffers,
    // otherwise the halo ring gets about a third and the weight ring the rest
    const int for_a = a.b_resident ? total_budget
                                   : std::max(2 * a.halo_bytes, std::min(total_budget / 3, total_budget - 2 * a.brow * L.BN * 128));
    a.a_stages = std::max(2, std::min(kMaxStages, for_a / a.halo_bytes));
  }
  const int fixed = base_fixed + a.a_stages * a.halo_bytes;
  const int budget = 226 * 1024 - fixed;
  a.stages = stage_bytes > 0 ? std::max(2, std::min(std::min(kMaxStages, ctx->max_stages), budget / stage_bytes)) : 2;
  const size_t smem = static_cast<size_t>(a.stages) * stage_bytes + fixed;
  if (smem > 227 * 1024) return set_err(ctx, FRB_ERR_INVALID, "conv: shared-memory plan does not fit (%zu bytes)", smem);
  int grid = std::min(a.total_tiles, ctx->num_sms);
  if (a.b_resident == 2) grid = grid / L.n_tiles * L.n_tiles;
  if (a.cta2) grid = 2 * std::min(a.total_tiles, ctx->num_sms / 2);
  if (a.cta2 && io.res_mode == RES_PRE_ACT) {
    // each CTA of a pair holds half of the identity's rows
    void* ident = ctx->d_ident[ctx->use_bf16 ? 1 : 0];
    if ((rc = encode_2d(ctx, &tmI, ident, 64, 64, 128, 64, 32)) != FRB_OK) return rc;
  }
  const CUtensorMap& tmBsel = a.cta2 ? L.tmB2 : L.tmB;
  const int act = io.res_mode == RES_POST_ACT_UPSAMPLE ? 0 : L.act_kind;
  cudaError_t le;
#ifdef FRB_DEBUG_BUILD
  if (a.stall != nullptr && !ctx->use_bf16) le = conv_tma_launch_stall(io.res_mode, act, grid, smem, stream, a, tmA, tmBsel, tmO, tmR, tmI);
  else
#endif
  if (a.cta2) le = ctx->use_bf16 ? conv_tma_launch_bf16_pair(io.res_mode, act, grid, smem, stream, a, tmA, tmBsel, tmO, tmR, tmI)
                                 : conv_tma_launch_f16_pair(io.res_mode, act, grid, smem, stream, a, tmA, tmBsel, tmO, tmR, tmI);
  else if (ctx->use_bf16) le = conv_tma_launch_bf16(io.res_mode, act, grid, smem, stream, a, tmA, tmBsel, tmO, tmR, tmI);
  else le = conv_tma_launch_f16(io.res_mode, act, grid, smem, stream, a, tmA, tmBsel, tmO, tmR, tmI);
  FRB_CUDA(ctx, le);
  ctx->launches++;
  return check_last(ctx, "conv_tma_kernel launch");
}

// Whether a conv of this shape takes the 8 x 16 halo patches the fused max pool needs: asked by the planner, answered by the
// launch logic itself on a dry run of its patch choice (no launch).
bool conv_pool_fusable(frb_ctx* ctx, const ConvLayer& L, int H, int W) {
  if (!(L.special || ((L.Cin % kBK) == 0 && L.kh * L.kw > 1 && L.sh == 1 && L.sw == 1)) || L.n_tiles != 1 || L.act_kind != 1 ||
      ctx->halo_mode < 0 || ctx->force_gather)
    return false;
  const int Ho = L.special ? H : conv_out_dim(H, L.kh, L.sh, L.ph), Wo = L.special ? W - 3 : conv_out_dim(W, L.kw, L.sw, L.pw);
  constexpr int kRefBatch = 64;
  int rw, rh, rb;
  pick_patch(kRefBatch, Ho, Wo, L.sw, L.sh, rw, rh, rb);
  const double eff_patch = static_cast<double>(kRefBatch) * Ho * Wo /
                           (128.0 * ((Wo + rw - 1) / rw) * ((Ho + rh - 1) / rh) * ((kRefBatch + rb - 1) / rb));
  const double eff_halo = static_cast<double>(Ho) * Wo / (128.0 * ((Wo + 7) / 8) * ((Ho + 15) / 16));
  const bool resident_w = L.n_tiles * L.kh * L.kw * ((L.Cin + kBK - 1) / kBK) * L.BN * 128 <= kResidentWeightBytes;
  return resident_w && eff_halo >= 0.85 * eff_patch;      // conv_launch_tma: `want` (resident weights) and the efficiency gate
}

int conv_launch(frb_ctx* ctx, const ConvLayer& L, const ConvIO& io, cudaStream_t stream) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.in = io.in;
  a.in_ld = io.in_ld;
  a.B = io.B; a.H = io.H; a.W = io.W; a.Cin = L.Cin;
  a.Ho = conv_out_dim(io.H, L.kh, L.sh, L.ph);
  a.Wo = conv_out_dim(io.W, L.kw, L.sw, L.pw);
  a.Cout = L.Cout;
  a.kh = L.kh; a.kw = L.kw; a.sh = L.sh; a.sw = L.sw; a.ph = L.ph; a.pw = L.pw;
  a.K = L.K;
  a.num_kb = L.K_pad / kBK;
  const long long M = static_cast<long long>(io.B) * a.Ho * a.Wo;
  if (M <= 0 || M > 0x7fffffffLL || static_cast<long long>(io.B) * io.H * io.W > 0x7fffffffLL)
    return set_err(ctx, FRB_ERR_INVALID, "conv: bad problem size B=%d H=%d W=%d", io.B, io.H, io.W);
  a.M = static_cast<int>(M);
  a.BN = L.BN;
  a.n_tiles = L.n_tiles;
  const long long m_tiles = (M + kBM - 1) / kBM;
  if (m_tiles * L.n_tiles > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "conv: too many tiles");
  a.total_tiles = static_cast<int>(m_tiles * L.n_tiles);
  a.bias = L.d_bias;
  a.out = io.out; a.out_ld = io.out_ld; a.out_mode = io.out_mode;
  a.slope = L.d_slope;
  a.res = io.res; a.res_ld = io.res_ld; a.res_mode = io.res_mode; a.rH = io.rH; a.rW = io.rW;
  a.loc = io.loc; a.conf = io.conf; a.landm = io.landm; a.P_total = io.P_total; a.P_off = io.P_off;
  a.watchdog = ctx->wd_dev;

  if ((reinterpret_cast<uintptr_t>(io.in) & 15) || (io.in_ld % 8) != 0)
    return set_err(ctx, FRB_ERR_INVALID, "conv: input must be 16-byte aligned with in_ld %% 8 == 0");
  if (io.out_mode != OUT_HEAD_SCATTER && io.pool_out == nullptr &&
      ((reinterpret_cast<uintptr_t>(io.out) & 15) || (io.out_ld % 8) != 0 || (L.Cout % 8) != 0))
    return set_err(ctx, FRB_ERR_INVALID, "conv: output must be 16-byte aligned with out_ld %% 8 == 0 and Cout %% 8 == 0");
  if (io.res_mode != RES_NONE && ((reinterpret_cast<uintptr_t>(io.res) & 15) || (io.res_ld % 8) != 0))
    return set_err(ctx, FRB_ERR_INVALID, "conv: residual must be 16-byte aligned with res_ld %% 8 == 0");

  if (L.special || tma_path_ok(ctx, L, io)) return conv_launch_tma(ctx, L, io, stream);

  a.a_tma = (!ctx->force_gather && L.kh == 1 && L.kw == 1 && L.sh == 1 && L.sw == 1 && L.ph == 0 && L.pw == 0) ? 1 : 0;
  CUtensorMap tmA;
  memset(&tmA, 0, sizeof(tmA));
  if (a.a_tma) {
    int rc = encode_2d(ctx, &tmA, io.in, static_cast<uint64_t>(L.Cin), static_cast<uint64_t>(a.M),
                       static_cast<uint64_t>(io.in_ld) * 2, kBK, kBM);
    if (rc != FRB_OK) return rc;
  }

  const int stage_bytes = kAStageBytes + L.BN * 128;
  const int budget = 200 * 1024;
  a.stages = std::max(3, std::min(kMaxStages, budget / stage_bytes));
  const size_t smem = static_cast<size_t>(a.stages) * stage_bytes + (2 * kMaxStages + 4) * 8 + 16 + 1024;

  const int grid = std::min(a.total_tiles, ctx->num_sms);
  if (ctx->use_bf16) {
    FRB_CUDA(ctx, cudaFuncSetAttribute(conv_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_igemm_kernel<true><<<grid, kConvThreads, smem, stream>>>(a, L.tmB, tmA);
  } else {
    FRB_CUDA(ctx, cudaFuncSetAttribute(conv_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_igemm_kernel<false><<<grid, kConvThreads, smem, stream>>>(a, L.tmB, tmA);
  }
  ctx->launches++;
  return check_last(ctx, "conv_igemm_kernel launch");
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_version(void) { return 100; }

int frb_init(int device, frb_ctx** out) {
  if (out == nullptr) return FRB_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return FRB_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FRB_ERR_NO_DEVICE;
  if (prop.major != 10) return FRB_ERR_NO_DEVICE;  // tcgen05 / TMEM kernels: sm_100 family only
  if (cudaSetDevice(device) != cudaSuccess) return FRB_ERR_CUDA;
  frb_ctx* ctx = new frb_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->wd_host), 64 + 148 * 32, cudaHostAllocMapped) == cudaSuccess) {
    memset(ctx->wd_host, 0, 64 + 148 * 32);
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->wd_dev), ctx->wd_host, 0) != cudaSuccess)
      ctx->wd_dev = nullptr;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete ctx;
    return FRB_ERR_CUDA;
  }
  ctx->encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  *out = ctx;
  return FRB_OK;
}

int frb_destroy(frb_ctx* ctx) {
  if (ctx == nullptr) return FRB_ERR_INVALID;
  for (auto& kv : ctx->nets) destroy_net(kv.second);
  ctx->nets.clear();
  if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
  if (ctx->wd_host) cudaFreeHost(ctx->wd_host);
  for (cudaEvent_t e : ctx->lt_events) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
    if (ctx->d_ident[i]) cudaFree(ctx->d_ident[i]);
  if (ctx->trace_dev) cudaFree(ctx->trace_dev);
  for (int i = 0; i < frb_ctx::kScratchSlots; ++i)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  delete ctx;
  return FRB_OK;
}

const char* frb_last_error(frb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int frb_set_option(frb_ctx* ctx, int option, int value) {
  if (ctx == nullptr) return FRB_ERR_INVALID;
  switch (option) {
    case FRB_OPT_BF16: ctx->use_bf16 = value ? 1 : 0; return FRB_OK;
    case FRB_OPT_BN_CAP:
      if (value < 16 || value > 256 || value % 16) return set_err(ctx, FRB_ERR_INVALID, "bn_cap must be 16..256 step 16");
      ctx->bn_cap = value;
      return FRB_OK;
    case FRB_OPT_HALO: ctx->halo_mode = value > 0 ? 1 : (value < 0 ? -1 : 0); return FRB_OK;
    case FRB_OPT_KBPS2:
      ctx->kbps2 = value ? 1 : 0;
      if (value >= 2) ctx->kbps2_min_stages = value;
      return FRB_OK;
    case FRB_OPT_LAYER_TIMES: ctx->layer_times = value ? 1 : 0; return FRB_OK;
    case FRB_OPT_DUAL_ISSUE: ctx->dual_issue = value; return FRB_OK;
    case FRB_OPT_PDL: ctx->pdl = value; return FRB_OK;
    case FRB_OPT_MPAIR: ctx->mpair = value; return FRB_OK;
    case FRB_OPT_WEIGHT_SLICE: ctx->weight_slice = value; return FRB_OK;
    case FRB_OPT_GRAPHS: ctx->graphs = value ? 1 : 0; return FRB_OK;
    case FRB_OPT_CTA2: ctx->cta2 = value ? 1 : 0; return FRB_OK;
    case FRB_OPT_FUSE_TAIL2:
      if (ctx->fuse_tail2 != (value ? 1 : 0)) {
        ctx->fuse_tail2 = value ? 1 : 0;
        for (auto& kv : ctx->nets) net_drop_plans(kv.second);
      }
      return FRB_OK;
    case FRB_OPT_FUSE_POOL:
      if (ctx->fuse_pool != (value ? 1 : 0)) {
        ctx->fuse_pool = value ? 1 : 0;
        for (auto& kv : ctx->nets) net_drop_plans(kv.second);
      }
      return FRB_OK;
    case FRB_OPT_FUSE_DWPW:
      if (value < 0 || value > 2) return set_err(ctx, FRB_ERR_INVALID, "fuse_dwpw must be 0, 1 or 2");
      if (ctx->fuse_dwpw != value) {
        ctx->fuse_dwpw = value;
        for (auto& kv : ctx->nets) net_drop_plans(kv.second);
      }
      return FRB_OK;
    case FRB_OPT_FUSE_TAIL:
      if (value < 0 || value > 3) return set_err(ctx, FRB_ERR_INVALID, "fuse_tail must be 0..3");
      if (ctx->fuse_tail != value) {
        ctx->fuse_tail = value;
        for (auto& kv : ctx->nets) net_drop_plans(kv.second);     // the op list of a forward changes
      }
      return FRB_OK;
#ifndef FRB_DEBUG_BUILD
    case FRB_OPT_FORCE_GATHER:
    case FRB_OPT_SKIP_LOADS:
    case FRB_OPT_MAX_STAGES:
    case FRB_OPT_TRACE:
      return set_err(ctx, FRB_ERR_STATE, "option %d is a tuning experiment of the debug build (include/facerec_b200_debug.h)", option);
#else
    case FRB_OPT_FORCE_GATHER: ctx->force_gather = value ? 1 : 0; return FRB_OK;
    case FRB_OPT_SKIP_LOADS: ctx->skip_loads = value ? 1 : 0; return FRB_OK;
    case FRB_OPT_MAX_STAGES: ctx->max_stages = value < 2 ? 2 : value; return FRB_OK;
    case FRB_OPT_TRACE:
      ctx->trace_mode = value;
      if (value && ctx->trace_dev == nullptr) {
        if (cudaMalloc(&ctx->trace_dev, (1 + 3 * 4000) * sizeof(long long)) != cudaSuccess) return set_err(ctx, FRB_ERR_CUDA, "trace alloc failed");
      } else if (!value && ctx->trace_dev != nullptr) {
        cudaFree(ctx->trace_dev);
        ctx->trace_dev = nullptr;
      }
      return FRB_OK;
#endif
  }
  return set_err(ctx, FRB_ERR_INVALID, "unknown option %d", option);
}

long long frb_launch_count(frb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int frb_debug_trace(frb_ctx* ctx, long long* h_out, int max_records) {
  if (ctx == nullptr || h_out == nullptr) return FRB_ERR_INVALID;
  if (ctx->trace_dev == nullptr) return set_err(ctx, FRB_ERR_STATE, "frb_debug_trace: FRB_OPT_TRACE is off (debug build only)");
  if (cudaDeviceSynchronize() != cudaSuccess) return FRB_ERR_CUDA;
  const int n = std::min(max_records, 4000);
  if (cudaMemcpy(h_out, ctx->trace_dev, (1 + 3 * static_cast<size_t>(n)) * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return FRB_ERR_CUDA;
  return FRB_OK;
}

int frb_debug_layer_times(frb_ctx* ctx, double* h_rows, int max_rows) {
  if (ctx == nullptr || h_rows == nullptr) return -FRB_ERR_INVALID;
  const int n = std::min<int>(max_rows, static_cast<int>(ctx->lt_info.size() / 6));
  if (n <= 0 || ctx->lt_events.size() < static_cast<size_t>(n) + 1) return 0;
  if (cudaEventSynchronize(ctx->lt_events[n]) != cudaSuccess) return -FRB_ERR_CUDA;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->lt_events[i], ctx->lt_events[i + 1]);
    for (int j = 0; j < 5; ++j) h_rows[i * 6 + j] = ctx->lt_info[i * 6 + j];
    h_rows[i * 6 + 5] = ms;
  }
  return n;
}

int frb_conv2d(frb_ctx* ctx, const void* d_in16, int B, int H, int W, int Cin, const float* h_w, const float* h_bias,
               int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int act, float slope, const void* d_res16,
               int res_mode, int rH, int rW, void* d_out, int out_fp32, int iters, float* h_ms, void* stream) {
  if (ctx == nullptr) return FRB_ERR_INVALID;
  FRB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvLayer L;
  L.Cin = Cin; L.Cout = Cout; L.kh = kh; L.kw = kw; L.sh = sh; L.sw = sw; L.ph = ph; L.pw = pw;
  if (res_mode == RES_PRE_ACT) L.bn_cap = 128;
  std::vector<float> sl(Cout, act == 0 ? 1.f : (act == 1 ? 0.f : slope));
  int rc = conv_prepare(ctx, L, h_w, h_bias, sl.data());
  if (rc != FRB_OK) { conv_release(L); return rc; }
  ConvIO io;
  io.in = d_in16; io.in_ld = Cin; io.B = B; io.H = H; io.W = W;
  io.out = d_out; io.out_ld = Cout; io.out_mode = out_fp32 ? OUT_FP32 : OUT_16BIT;
  io.res = d_res16; io.res_ld = Cout; io.res_mode = res_mode; io.rH = rH; io.rW = rW;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h_ms != nullptr) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  rc = conv_launch(ctx, L, io, st);  // warm-up / the functional result
  if (rc == FRB_OK && h_ms != nullptr && iters > 0) {
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters && rc == FRB_OK; ++i) rc = conv_launch(ctx, L, io, st);
    cudaEventRecord(e1, st);
  }
  if (rc == FRB_OK) rc = check_cuda(ctx, cudaStreamSynchronize(st), "frb_conv2d sync");
  if (rc == FRB_OK && h_ms != nullptr) {
    float ms = 0.f;
    if (iters > 0) cudaEventElapsedTime(&ms, e0, e1);
    *h_ms = iters > 0 ? ms / iters : 0.f;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  conv_release(L);
  return rc;
}

}  // extern "C"
