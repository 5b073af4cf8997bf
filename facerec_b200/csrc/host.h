// Internal host-side declarations shared by the translation units of libfacerec_b200.so.
// Nothing here crosses the C ABI; see include/facerec_b200.h for the exported surface.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdarg.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/facerec_b200.h"
#include "../../include/facerec_b200_debug.h"

namespace frb {

struct Net;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace frb

struct frb_ctx {
  int device = 0;
  int num_sms = 148;
  int use_bf16 = 0;            // 0: fp16 activations/weights (default), 1: bf16
  int bn_cap = 256;            // largest conv tile N the planner may pick
  int force_gather = 0;        // debugging: never use the TMA A path
  void* d_ident[2] = {nullptr, nullptr};   // 64 x 64 identity (fp16, bf16): B operand of the residual MMAs
  int dual_issue = 1;          // FRB_OPT_DUAL_ISSUE
  int mpair = 1;               // FRB_OPT_MPAIR: 0 never, 1 for N tiles of 192 and more, 2 wherever possible
  int weight_slice = 1;        // FRB_OPT_WEIGHT_SLICE
  int pdl = 1;                 // FRB_OPT_PDL: 0 never, 1 short launches, 2 always
  int skip_loads = 0;          // FRB_OPT_SKIP_LOADS: timing experiment, results are garbage
  int max_stages = 64;         // FRB_OPT_MAX_STAGES: cap on the conv_tma stage ring (latency-sensitivity experiments)
  int kbps2_min_stages = 3;    // FRB_OPT_KBPS2 value >= 2 sets it: fewest double stages that must fit for two k-blocks per stage
  int kbps2 = 1;               // two 64-wide k-blocks per pipeline stage (0: one)
  int layer_times = 0;         // FRB_OPT_LAYER_TIMES: run_plan brackets every op with events (debug only)
  std::vector<cudaEvent_t> lt_events;      // lt_events[0] = before the first op, [i+1] = after op i
  std::vector<double> lt_info;             // 6 doubles per op: kind, M, N, K, flop, (ms filled by frb_debug_layer_times)
  int trace_mode = 0;          // FRB_OPT_TRACE: per-CTA stall counters of the conv_tma roles
  long long* trace_dev = nullptr;  // FRB_OPT_TRACE: device buffer for conv_tma timeline records
  cudaStream_t capture_stream = nullptr;   // graph capture of short forwards happens here, replay on the caller's stream
  int cta2 = 1;                // FRB_OPT_CTA2: CTA pairs (cta_group::2) for per-tap / linear convs with streamed weights
  int graphs = 0;              // FRB_OPT_GRAPHS: short forwards (InceptionResnetV1, MobileNet RetinaFace at small sizes) replay a CUDA graph
  int fuse_tail = 3;           // FRB_OPT_FUSE_TAIL: layer1 of R50 runs conv3 + skip and the next block's conv1 as one launch (conv_tail.cu)
  int fuse_tail2 = 1;          // FRB_OPT_FUSE_TAIL2: the same fusion in layer2 (streamed weights, conv_tail2_kernel)
  int fuse_pool = 1;           // FRB_OPT_FUSE_POOL: the ResNet stem's max pool runs in the stem conv's epilogue
  int fuse_dwpw = 1;           // FRB_OPT_FUSE_DWPW: MobileNet blocks (depthwise 3x3 + pointwise 1x1) as one launch where the weights fit
  int halo_mode = 0;           // halo tiles for stride-1 k x k convs: 0 where it pays (see conv_launch_tma), 1 always, -1 never
  std::string err;
  int* wd_host = nullptr;      // watchdog words (mapped pinned memory)
  int* wd_dev = nullptr;
  frb::EncodeTiledFn encode_tiled = nullptr;
  std::map<int, frb::Net*> nets;
  long long launches = 0;      // kernels launched through this context (bench bookkeeping)
  bool attr_nms = false;       // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: tracked per context
  bool attr_dist = false;
  // frb_knn_fit: the database whose split planes / norms are kept in scratch slots 0 / 1 (and its max |x|^2, scale exponent)
  const float* knn_fit_db = nullptr;
  int knn_fit_m = 0, knn_fit_d = 0;
  unsigned int knn_fit_max_norm_bits = 0;
  int knn_fit_exp = 0;
  int knn_redo_last = 0;       // queries of the last frb_knn_topk call that needed the exact three-segment pass
  static constexpr int kScratchSlots = 12;   // 0-7 per-call scratch of the distance stage, 8-10 the fitted kNN database
  void* scratch[kScratchSlots] = {};
  size_t scratch_bytes[kScratchSlots] = {};
};

namespace frb {

int set_err(frb_ctx* ctx, int code, const char* fmt, ...);
int check_cuda(frb_ctx* ctx, cudaError_t e, const char* what);
// Maps a failed launch / sync to an error string, appending the pipeline watchdog record if set.
int check_last(frb_ctx* ctx, const char* what);

#define FRB_CUDA(ctx, expr)                                   \
  do {                                                        \
    int _rc = frb::check_cuda((ctx), (expr), #expr);          \
    if (_rc != 0) return _rc;                                 \
  } while (0)

// ------------------------------------------------------------------ conv
struct ConvLayer {
  int Cin = 0, Cout = 0, kh = 1, kw = 1, sh = 1, sw = 1, ph = 0, pw = 0;
  int K = 0, K_pad = 0, BN = 0, Cout_pad = 0, n_tiles = 0;
  int act_kind = 0;          // 0 per-channel slope, 1 ReLU on every channel, 2 linear on every channel
  int special = 0;           // 1: ResNet stem over the space-to-depth input (window conv, see nets.cu)
  int direct3 = 0;           // 1: first conv of a network (3 real input channels of NHWC8, 3x3): CUDA-core kernel K3
  int bn_cap = 0;            // per-layer cap on the N tile (0 = context default); residual layers use 128
  void* d_w = nullptr;       // [Cout_pad, K_pad] 16-bit, K-major, K ordered (r, s, c)
  float* d_bias = nullptr;   // [Cout_pad]
  float* d_slope = nullptr;  // [Cout_pad] negative slope per output channel (1 linear, 0 relu)
  float* d_bias_skip = nullptr;   // [Cout_pad] this layer's bias + the skip conv's (conv3 of a block whose skip conv K2t fuses)
  CUtensorMap tmB;
  CUtensorMap tmB2;          // the same matrix with a box of BN / 2 rows: one CTA's half of a weight block in pair mode
  int has_tmB2 = 0;
};

struct ConvIO {
  const void* in = nullptr;  // pre-offset to first input channel
  int in_ld = 0;
  int B = 0, H = 0, W = 0;
  void* out = nullptr;       // pre-offset to first output channel
  int out_ld = 0;
  int out_mode = 0;          // ConvOut
  const void* res = nullptr;
  int res_ld = 0;
  int res_mode = 0;          // ConvRes
  int rH = 0, rW = 0;
  float* loc = nullptr;
  float* conf = nullptr;
  float* landm = nullptr;
  int P_total = 0, P_off = 0;
  // 3x3 / stride 2 / pad 1 max pool fused into the conv's epilogue (the ResNet stem): `out` is not written; pool_out is
  // [B, pool_Ho, pool_Wo, pool_ld] and must have been prepared by launch_pool_border_zero
  void* pool_out = nullptr;
  int pool_ld = 0;
};

inline int conv_out_dim(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }

// 2D tiled tensor map over a row-major 16-bit matrix [outer, inner], 128-byte swizzle.
int encode_2d(frb_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
              uint32_t box_inner, uint32_t box_outer);
// 64 x 64 identity matrix in the context's activation type (device memory owned by the context)
int ctx_identity(frb_ctx* ctx, void** out);
// Grow-only device scratch owned by the context (slot = independent buffer).
int ctx_scratch(frb_ctx* ctx, int slot, size_t bytes, void** out);

// h_w: [Cout][kh][kw][Cin] fp32 (BN already folded); h_bias: [Cout] or null.
// h_slope: [Cout] per-channel negative slope or null (= linear everywhere).
int conv_prepare(frb_ctx* ctx, ConvLayer& L, const float* h_w, const float* h_bias, const float* h_slope);
void conv_release(ConvLayer& L);
int conv_launch(frb_ctx* ctx, const ConvLayer& L, const ConvIO& io, cudaStream_t stream);
// true when a conv over an H x W input runs on 8 x 16 halo patches, i.e. can take ConvIO::pool_out (the fused 3x3 / 2 max pool)
bool conv_pool_fusable(frb_ctx* ctx, const ConvLayer& L, int H, int W);
// conv_tail.cu: y = relu(conv3(in) + res) [M, 256] and y1 = relu(conv1_next(y)) [M, L1.Cout] in one launch (K2t)
// With Ld != null the skip is Ld's 1x1 conv of `res` (= its input x) computed inside, bias1 = conv3's bias + Ld's.
bool conv_tail_applicable(const ConvLayer& L3, const ConvLayer& L1);
bool conv_tail_skip_applicable(const ConvLayer& Ld);
// K2t2: the same for layer2 (conv3 128 -> 512 + skip, next conv1 512 -> 128; both weight matrices streamed)
bool conv_tail2_applicable(const ConvLayer& L3, const ConvLayer& L1);
int conv_tail2_launch(frb_ctx* ctx, const ConvLayer& L3, const ConvLayer& L1, const void* in, int in_ld, long long M,
                      const void* res, int res_ld, void* out1, int out1_ld, void* out2, int out2_ld, cudaStream_t stream);
int conv_tail_launch(frb_ctx* ctx, const ConvLayer& L3, const ConvLayer& L1, const void* in, int in_ld, long long M,
                     const void* res, int res_ld, void* out1, int out1_ld, void* out2, int out2_ld, const ConvLayer* Ld,
                     const float* bias1, cudaStream_t stream);

// ------------------------------------------------------------------ glue kernels (glue.cu)
int launch_preprocess_s2d(frb_ctx*, const uint8_t* d_frames, int B, int H, int W, int src_is_rgb, void* d_out, cudaStream_t);
int launch_preprocess(frb_ctx*, const uint8_t* d_frames, int B, int H, int W, int src_is_rgb, int out_c, void* d_out,
                      cudaStream_t);
int launch_conv3x3_c3_direct(frb_ctx*, const void* in, int B, int H, int W, int stride, int pad, const void* d_w16, int K_pad,
                             int Cout, const float* d_bias, const float* d_slope, void* out, int out_ld, cudaStream_t);
int launch_dwconv3x3(frb_ctx*, const void* in, int B, int H, int W, int C, int stride, const void* d_w9c,
                     const float* d_bias, int act, float slope, void* out, cudaStream_t);
// K4f: depthwise 3x3 + pointwise 1x1 of a MobileNet block in one launch (the depthwise output stays in shared memory)
bool dwpw_fusable(int C, int Cout);
int launch_dwpw_fused(frb_ctx*, const void* in, int B, int H, int W, int C, int stride, const void* d_dw_w9c,
                      const float* d_dw_bias, float dw_slope, const void* d_pw_w16, int K_pad, int Cout, const float* d_pw_bias,
                      const float* d_pw_slope, void* out, int out_ld, cudaStream_t);
// zeroes the positions of a pooled tensor that the fused stem + pool epilogue max-reduces into (row % 8 == 0 or col % 4 == 0)
int launch_pool_border_zero(frb_ctx*, void* out, int B, int Ho, int Wo, int C, int out_ld, cudaStream_t);
int launch_maxpool3x3s2(frb_ctx*, const void* in, int in_ld, int B, int H, int W, int C, int pad, void* out, int out_ld,
                        cudaStream_t);
int launch_avgpool(frb_ctx*, const void* in, int B, int HW, int C, void* out, cudaStream_t);
int launch_l2norm(frb_ctx*, float* x, int rows, int cols, cudaStream_t);
int launch_prior_decode_filter(frb_ctx*, const float* d_loc, const float* d_conf, const float* d_landm, int B, int H,
                               int W, float thr, int max_cand, float* d_cand, int32_t* d_count, int32_t* d_prior_idx,
                               cudaStream_t);
int launch_nms(frb_ctx*, const float* d_cand, const int32_t* d_count, const int32_t* d_prior_idx, int B, int max_cand,
               float iou_thr, int legacy_plus_one, int32_t* d_keep, int32_t* d_nkeep, cudaStream_t);
int launch_gather_kept(frb_ctx*, const float* d_cand, const int32_t* d_keep, const int32_t* d_nkeep, int B, int max_cand,
                       int max_out, float* d_out, int32_t* d_nout, cudaStream_t);
int launch_crop_resize(frb_ctx*, const uint8_t* d_frames, int B, int H, int W, const int32_t* d_boxes, int F,
                       int out_c, void* d_out, cudaStream_t);

// ------------------------------------------------------------------ distance stage (dist.cu)
int run_pdist(frb_ctx*, const float* d_X, int N, int D, int row0, int rows, int layout, void* d_out, cudaStream_t);
int run_knn_topk(frb_ctx*, const float* d_Q, int Q, const float* d_DB, int M, int D, int k, int32_t* d_idx, float* d_dist,
                 cudaStream_t);
int run_knn_fit(frb_ctx*, const float* d_DB, int M, int D, cudaStream_t);
int run_knn_vote(frb_ctx*, const int32_t* d_idx, const int32_t* d_label_idx, int Q, int k, int n_classes, float* d_proba,
                 cudaStream_t);
int run_segment_mean(frb_ctx*, const float* d_X, const int32_t* d_offsets, int S, int D, float* d_out, cudaStream_t);

// ------------------------------------------------------------------ clustering (linkage.cu)
int run_linkage_complete(frb_ctx*, float* d_D, int n, int32_t* d_zxy, float* d_zd, cudaStream_t);

// ------------------------------------------------------------------ scene cuts (scene.cu)
int run_scene_stats(frb_ctx*, const uint8_t* d_frames, int B, int H, int W, int crop, const uint8_t* d_prev_frame,
                    const double* d_prev_cdf, uint32_t* d_hist, double* d_cdf, double* d_sums, cudaStream_t);

// ------------------------------------------------------------------ nets (nets.cu)
struct WeightBlob;
Net* build_net(frb_ctx* ctx, int model_id, const void* blob, size_t bytes);
int net_forward_retina(frb_ctx* ctx, Net* net, const void* d_in, int B, int H, int W, float* d_loc, float* d_conf,
                       float* d_landm, cudaStream_t stream);
int net_forward_irv1(frb_ctx* ctx, Net* net, const void* d_in, int F, float* d_emb, cudaStream_t stream);
void destroy_net(Net* net);
void net_drop_plans(Net* net);
long long net_workspace_bytes(Net* net, int B, int H, int W);

}  // namespace frb
