// extern "C" surface of libfacerec_b200.so (see include/facerec_b200.h): argument checks and
// dispatch only; kernels live in glue.cu / conv_igemm.cuh / dist.cu, graphs in nets.cu.
#include <stdio.h>
#include <string.h>

#include "host.h"

using namespace frb;

#define FRB_NEED(ctx, cond, msg)                                  \
  do {                                                            \
    if (!(cond)) return set_err((ctx), FRB_ERR_INVALID, "%s", (msg)); \
  } while (0)

// Every entry point runs on the context's device whatever the caller's current device is (two contexts on two GPUs in
// one process are legal); the work itself goes to the caller's stream, which must belong to that device.  A context is
// meant to be driven from ONE stream at a time: the per-network activation arena and the scratch buffers are shared by
// all calls on the context.
#define FRB_ENTER(ctx)                                                        \
  do {                                                                        \
    if ((ctx) == nullptr) return FRB_ERR_INVALID;                             \
    int _cur = -1;                                                            \
    if (cudaGetDevice(&_cur) != cudaSuccess || _cur != (ctx)->device)         \
      FRB_CUDA((ctx), cudaSetDevice((ctx)->device));                          \
  } while (0)

static Net* find_net(frb_ctx* ctx, int model_id) {
  auto it = ctx->nets.find(model_id);
  return it == ctx->nets.end() ? nullptr : it->second;
}

extern "C" {

int frb_load_model(frb_ctx* ctx, int model_id, const void* h_blob, size_t bytes) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, h_blob != nullptr && bytes > 0, "frb_load_model: empty blob");
  FRB_NEED(ctx, model_id >= FRB_RETINA_MNET025 && model_id <= FRB_IRV1, "frb_load_model: unknown model id");
  Net* net = build_net(ctx, model_id, h_blob, bytes);
  if (net == nullptr) return FRB_ERR_INVALID;   // message set by build_net
  Net* old = find_net(ctx, model_id);
  if (old != nullptr) {
    cudaDeviceSynchronize();
    destroy_net(old);
  }
  ctx->nets[model_id] = net;
  return FRB_OK;
}

int frb_num_priors(int H, int W) {
  int p = 0;
  for (int s = 8; s <= 32; s *= 2) p += 2 * ((H + s - 1) / s) * ((W + s - 1) / s);
  return p;
}

int frb_input_layout(int model_id) { return model_id == FRB_RETINA_R50 ? FRB_LAYOUT_S2D16 : FRB_LAYOUT_NHWC8; }

size_t frb_input_elems(int layout, int B, int H, int W) {
  if (layout == FRB_LAYOUT_S2D16) return static_cast<size_t>(B) * ((H + 1) / 2) * ((W + 1) / 2 + 3) * 16;
  return static_cast<size_t>(B) * H * W * 8;
}

int frb_preprocess_u8(frb_ctx* ctx, const uint8_t* d_frames, int B, int H, int W, int src_is_rgb, int layout,
                      void* d_out16, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_frames && d_out16 && B > 0 && H > 0 && W > 0, "frb_preprocess_u8: bad arguments");
  if (layout == FRB_LAYOUT_S2D16)
    return launch_preprocess_s2d(ctx, d_frames, B, H, W, src_is_rgb, d_out16, static_cast<cudaStream_t>(stream));
  FRB_NEED(ctx, layout == FRB_LAYOUT_NHWC8, "frb_preprocess_u8: unknown layout");
  return launch_preprocess(ctx, d_frames, B, H, W, src_is_rgb, 8, d_out16, static_cast<cudaStream_t>(stream));
}

int frb_retina_forward(frb_ctx* ctx, int model_id, const void* d_in16, int B, int H, int W, float* d_loc, float* d_conf,
                       float* d_landm, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, model_id == FRB_RETINA_MNET025 || model_id == FRB_RETINA_R50, "frb_retina_forward: not a RetinaFace model id");
  FRB_NEED(ctx, d_in16 && d_loc && d_conf && d_landm, "frb_retina_forward: null pointer");
  Net* net = find_net(ctx, model_id);
  if (net == nullptr) return set_err(ctx, FRB_ERR_STATE, "frb_retina_forward: model %d not loaded", model_id);
  return net_forward_retina(ctx, net, d_in16, B, H, W, d_loc, d_conf, d_landm, static_cast<cudaStream_t>(stream));
}

int frb_prior_decode_filter(frb_ctx* ctx, const float* d_loc, const float* d_conf, const float* d_landm, int B, int H,
                            int W, float score_thr, int max_cand, float* d_cand, int32_t* d_count, int32_t* d_prior_idx,
                            void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_loc && d_conf && d_landm && d_cand && d_count && d_prior_idx, "frb_prior_decode_filter: null pointer");
  FRB_NEED(ctx, H > 0 && W > 0, "frb_prior_decode_filter: bad frame size");
  return launch_prior_decode_filter(ctx, d_loc, d_conf, d_landm, B, H, W, score_thr, max_cand, d_cand, d_count,
                                    d_prior_idx, static_cast<cudaStream_t>(stream));
}

int frb_nms(frb_ctx* ctx, const float* d_cand, const int32_t* d_count, const int32_t* d_prior_idx, int B, int max_cand,
            float iou_thr, int legacy_plus_one, int32_t* d_keep, int32_t* d_nkeep, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_cand && d_count && d_prior_idx && d_keep && d_nkeep, "frb_nms: null pointer");
  return launch_nms(ctx, d_cand, d_count, d_prior_idx, B, max_cand, iou_thr, legacy_plus_one, d_keep, d_nkeep,
                    static_cast<cudaStream_t>(stream));
}

int frb_gather_kept(frb_ctx* ctx, const float* d_cand, const int32_t* d_keep, const int32_t* d_nkeep, int B, int max_cand,
                    int max_out, float* d_out, int32_t* d_nout, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_cand && d_keep && d_nkeep && d_out && d_nout, "frb_gather_kept: null pointer");
  return launch_gather_kept(ctx, d_cand, d_keep, d_nkeep, B, max_cand, max_out, d_out, d_nout,
                            static_cast<cudaStream_t>(stream));
}

int frb_crop_resize_std(frb_ctx* ctx, const uint8_t* d_frames_rgb, int B, int H, int W, const int32_t* d_boxes, int F,
                        void* d_out16, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_frames_rgb && (F == 0 || (d_boxes && d_out16)), "frb_crop_resize_std: null pointer");
  FRB_NEED(ctx, B > 0 && H > 0 && W > 0 && F >= 0, "frb_crop_resize_std: bad shape");
  return launch_crop_resize(ctx, d_frames_rgb, B, H, W, d_boxes, F, 8, d_out16, static_cast<cudaStream_t>(stream));
}

int frb_irv1_forward(frb_ctx* ctx, const void* d_in16, int F, float* d_emb, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_in16 && d_emb, "frb_irv1_forward: null pointer");
  Net* net = find_net(ctx, FRB_IRV1);
  if (net == nullptr) return set_err(ctx, FRB_ERR_STATE, "frb_irv1_forward: InceptionResnetV1 not loaded");
  return net_forward_irv1(ctx, net, d_in16, F, d_emb, static_cast<cudaStream_t>(stream));
}

int frb_scene_stats(frb_ctx* ctx, const uint8_t* d_frames_rgb, int B, int H, int W, int crop, const uint8_t* d_prev_frame,
                    const double* d_prev_cdf, uint32_t* d_hist, double* d_cdf, double* d_sums, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_frames_rgb && d_hist && d_cdf && d_sums, "frb_scene_stats: null pointer");
  FRB_NEED(ctx, (d_prev_frame == nullptr) == (d_prev_cdf == nullptr), "frb_scene_stats: prev frame and prev cdf go together");
  return run_scene_stats(ctx, d_frames_rgb, B, H, W, crop, d_prev_frame, d_prev_cdf, d_hist, d_cdf, d_sums,
                         static_cast<cudaStream_t>(stream));
}

int frb_segment_mean(frb_ctx* ctx, const float* d_X, const int32_t* d_offsets, int S, int D, float* d_out, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, S == 0 || (d_X && d_offsets && d_out), "frb_segment_mean: null pointer");
  return run_segment_mean(ctx, d_X, d_offsets, S, D, d_out, static_cast<cudaStream_t>(stream));
}

int frb_pdist(frb_ctx* ctx, const float* d_X, int N, int D, int row0, int rows, int layout, void* d_out, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_X && d_out, "frb_pdist: null pointer");
  return run_pdist(ctx, d_X, N, D, row0, rows, layout, d_out, static_cast<cudaStream_t>(stream));
}

int frb_linkage_complete(frb_ctx* ctx, float* d_D, int N, int32_t* d_merge_pairs, float* d_merge_dist, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_D && d_merge_pairs && d_merge_dist, "frb_linkage_complete: null pointer");
  return run_linkage_complete(ctx, d_D, N, d_merge_pairs, d_merge_dist, static_cast<cudaStream_t>(stream));
}

int frb_knn_topk(frb_ctx* ctx, const float* d_Q, int Q, const float* d_DB, int M, int D, int k, int32_t* d_idx,
                 float* d_dist, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, Q == 0 || (d_Q && d_DB && d_idx && d_dist), "frb_knn_topk: null pointer");
  return run_knn_topk(ctx, d_Q, Q, d_DB, M, D, k, d_idx, d_dist, static_cast<cudaStream_t>(stream));
}

long long frb_workspace_bytes(frb_ctx* ctx, int model_id, int B, int H, int W) {
  if (ctx == nullptr) return -1;
  Net* net = find_net(ctx, model_id);
  if (net == nullptr || B <= 0) return -1;
  if (model_id == FRB_IRV1) { H = 160; W = 160; }
  return net_workspace_bytes(net, B, H, W);
}

int frb_knn_fit(frb_ctx* ctx, const float* d_DB, int M, int D, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, d_DB != nullptr, "frb_knn_fit: null pointer");
  return run_knn_fit(ctx, d_DB, M, D, static_cast<cudaStream_t>(stream));
}

int frb_knn_redo_count(frb_ctx* ctx) { return ctx ? ctx->knn_redo_last : -1; }

int frb_knn_vote(frb_ctx* ctx, const int32_t* d_idx, const int32_t* d_label_idx, int Q, int k, int n_classes,
                 float* d_proba, void* stream) {
  FRB_ENTER(ctx);
  FRB_NEED(ctx, Q == 0 || (d_idx && d_label_idx && d_proba), "frb_knn_vote: null pointer");
  return run_knn_vote(ctx, d_idx, d_label_idx, Q, k, n_classes, d_proba, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
