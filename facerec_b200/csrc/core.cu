// Context lifecycle, error plumbing and the host side of K2 (conv layer preparation + launch).
#include <stdio.h>
#include <string.h>
#include <algorithm>

#include "host.h"
#include "conv_igemm.cuh"

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
    char buf[128];
    snprintf(buf, sizeof(buf), " [pipeline watchdog: wait code %d in block %d]", ctx->wd_host[0], ctx->wd_host[1]);
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
  if (slot < 0 || slot >= 8) return set_err(ctx, FRB_ERR_INVALID, "bad scratch slot %d", slot);
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

// ----------------------------------------------------------------------------- conv host side
static int pick_bn(int cout_pad16, int cap) {
  cap = std::max(16, std::min(256, cap)) & ~15;
  if (cout_pad16 <= cap) return cout_pad16;
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
  L.BN = pick_bn(cout16, ctx->bn_cap);
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

  FRB_CUDA(ctx, cudaMalloc(&L.d_w, w16.size() * 2));
  FRB_CUDA(ctx, cudaMalloc(&L.d_bias, bias.size() * 4));
  FRB_CUDA(ctx, cudaMalloc(&L.d_slope, slope.size() * 4));
  FRB_CUDA(ctx, cudaMemcpy(L.d_w, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
  FRB_CUDA(ctx, cudaMemcpy(L.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  FRB_CUDA(ctx, cudaMemcpy(L.d_slope, slope.data(), slope.size() * 4, cudaMemcpyHostToDevice));
  return encode_2d(ctx, &L.tmB, L.d_w, L.K_pad, L.Cout_pad, static_cast<uint64_t>(L.K_pad) * 2, kBK, L.BN);
}

void conv_release(ConvLayer& L) {
  if (L.d_w) cudaFree(L.d_w);
  if (L.d_bias) cudaFree(L.d_bias);
  if (L.d_slope) cudaFree(L.d_slope);
  L.d_w = nullptr;
  L.d_bias = nullptr;
  L.d_slope = nullptr;
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
  if (io.out_mode != OUT_HEAD_SCATTER &&
      ((reinterpret_cast<uintptr_t>(io.out) & 15) || (io.out_ld % 8) != 0 || (L.Cout % 8) != 0))
    return set_err(ctx, FRB_ERR_INVALID, "conv: output must be 16-byte aligned with out_ld %% 8 == 0 and Cout %% 8 == 0");
  if (io.res_mode != RES_NONE && ((reinterpret_cast<uintptr_t>(io.res) & 15) || (io.res_ld % 8) != 0))
    return set_err(ctx, FRB_ERR_INVALID, "conv: residual must be 16-byte aligned with res_ld %% 8 == 0");

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
  if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->wd_host), 64, cudaHostAllocMapped) == cudaSuccess) {
    memset(ctx->wd_host, 0, 64);
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
  if (ctx->wd_host) cudaFreeHost(ctx->wd_host);
  for (int i = 0; i < 8; ++i)
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
    case FRB_OPT_FORCE_GATHER: ctx->force_gather = value ? 1 : 0; return FRB_OK;
  }
  return set_err(ctx, FRB_ERR_INVALID, "unknown option %d", option);
}

long long frb_launch_count(frb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int frb_conv2d(frb_ctx* ctx, const void* d_in16, int B, int H, int W, int Cin, const float* h_w, const float* h_bias,
               int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int act, float slope, const void* d_res16,
               int res_mode, int rH, int rW, void* d_out, int out_fp32, int iters, float* h_ms, void* stream) {
  if (ctx == nullptr) return FRB_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvLayer L;
  L.Cin = Cin; L.Cout = Cout; L.kh = kh; L.kw = kw; L.sh = sh; L.sw = sw; L.ph = ph; L.pw = pw;
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
