/* facerec_b200 — C ABI of libfacerec_b200.so (CUDA, sm_100a only).
 *
 * Drop-in boundary for the per-frame face path of MoMaF/facerec.  The reference has no FFI of
 * its own: its operator API for this path is three duck-typed Python call sites plus two
 * library calls.  Each entry point below names the reference site whose arithmetic it replaces
 * (paths relative to the reference checkout).  The Python adapters in facerec_b200/ present the
 * reference-side objects (detect(), embeddings(), linkage input, kNN classifier) on top of this
 * ABI; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 (FRB_OK) or a negative frb_status; frb_last_error(ctx) gives text;
 *  - `d_` pointers are device pointers owned by the caller; `h_` pointers are host pointers;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *    asynchronous on that stream unless stated otherwise;
 *  - one frb_ctx per (process, device); a ctx is not thread-safe;
 *  - no torch / pybind types cross this boundary;
 *  - there is no CPU fallback: without an sm_100 device frb_init fails.
 */
#ifndef FACEREC_B200_H_
#define FACEREC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct frb_ctx frb_ctx;

enum frb_status {
  FRB_OK = 0,
  FRB_ERR_INVALID = -1,   /* bad argument / unmet precondition */
  FRB_ERR_CUDA = -2,      /* CUDA runtime or driver error (text in frb_last_error) */
  FRB_ERR_STATE = -3,     /* e.g. model not loaded */
  FRB_ERR_NO_DEVICE = -4  /* no sm_100 GPU */
};

enum frb_model {
  FRB_RETINA_MNET025 = 0, /* RetinaFace, MobileNetV1-0.25 backbone, FPN/SSH 64 ch  */
  FRB_RETINA_R50 = 1,     /* RetinaFace, ResNet-50 backbone, FPN/SSH 256 ch        */
  FRB_IRV1 = 2            /* FaceNet InceptionResnetV1, 512-d                      */
};

enum frb_pdist_layout {
  FRB_CONDENSED_F32 = 0,  /* scipy pdist order, float  */
  FRB_CONDENSED_F64 = 1,  /* scipy pdist order, double */
  FRB_SQUARE_F32 = 2,     /* rows x N row-major, float */
  FRB_SQUARE_SYM_F32 = 3  /* N x N, only j >= i computed and mirrored: exactly symmetric (input of frb_linkage_complete) */
};

enum frb_option {
  FRB_OPT_BF16 = 0,        /* 1: bf16 activations/weights instead of fp16 (set before frb_load_model) */
  FRB_OPT_BN_CAP = 1,      /* largest conv tile N (multiple of 16, <= 256)                           */
  FRB_OPT_HALO = 3,        /* stride-1 k x k convs reading one halo tile per channel block instead of one TMA box
                              per filter tap: 0 (default) where it pays (resident weights, or a filter row of
                              streamed weights per stage), 1 wherever possible, -1 never                          */
  FRB_OPT_KBPS2 = 5,       /* 0: one 64-wide k-block per pipeline stage instead of two                     */
  FRB_OPT_LAYER_TIMES = 6, /* 1: network forwards bracket every op with events, read by frb_debug_layer_times */
  FRB_OPT_DUAL_ISSUE = 9,  /* 0: a single MMA issuer warp per CTA (default 1: two, alternating tiles, where safe)   */
  FRB_OPT_PDL = 10,        /* programmatic dependent launch of the conv kernels (set-up under the previous kernel's   */
                           /* tail): 0 never, 1 (default) launches estimated under 100 us, 2 always                  */
  FRB_OPT_WEIGHT_SLICE = 11,/* 0: stream the weights of every tile (default 1: a CTA keeps its N tile's weights in    */
                           /* shared memory where the slice fits and the grid can be a multiple of the N tiles)      */
  FRB_OPT_MPAIR = 12,      /* halo-mode convs with streamed weights work on two 8 x 16 patches per weight stage:     */
                           /* 0 never, 1 (default) for N tiles of 192 and more, 2 wherever possible                  */
  FRB_OPT_CTA2 = 14,       /* 0: never run conv CTAs as pairs (default 1: convs on the per-tap / linear path with streamed weights   */
                           /* use cta_group::2 - M = 256 MMAs over the two SMs of a TPC, each CTA loading half of the weight block)  */
  FRB_OPT_FUSE_TAIL = 15,  /* 0: every conv of ResNet-50 is a launch of its own; 1: in layer1 a bottleneck's conv3 + skip and the   */
                           /* next bottleneck's conv1 run as ONE launch that never re-reads the 256-channel tensor (same bits as 0);  */
                           /* 2: as 1, but not across the layer1 -> layer2 edge; 3 (default): as 1, and the 1x1 skip conv of          */
                           /* layer1's first block is computed inside that launch too (its output is never rounded to 16 bits, so     */
                           /* results differ from 0 by that one rounding)                                                           */
  FRB_OPT_FUSE_TAIL2 = 18, /* 0: FRB_OPT_FUSE_TAIL stops at layer1 (default 1: layer2's conv3 + skip -> next conv1 pairs are fused as well,     */
                           /* with both weight matrices streamed; same bits)                                                            */
  FRB_OPT_FUSE_POOL = 17,  /* 0: the ResNet-50 stem and its 3x3 / stride 2 max pool are two launches (default 1: the pool runs in the stem's  */
                           /* epilogue, the stem output is never stored; the pooled tensor has the same bits)                            */
  FRB_OPT_FUSE_DWPW = 16,  /* 0: depthwise and pointwise convs of RetinaFace-MobileNet0.25 are separate launches; 1 (default): one launch */
                           /* per block for the blocks with up to 32 input channels (the depthwise output stays in shared memory; same   */
                           /* rounding points, other summation order); 2: every block whose weights fit                                  */
  FRB_OPT_GRAPHS = 13      /* 1: short forwards (InceptionResnetV1, RetinaFace-MobileNet0.25 up to 64 x 360p) are captured once per  */
                           /* shape into a CUDA graph and replayed.  Default 0: measured SLOWER than stream launches with         */
                           /* programmatic dependent launch (the kernels are GPU-bound at 10-35 us, and a replayed graph loses    */
                           /* the set-up overlap: 1.49 vs 1.23 ms at 64 faces, profiles/r02_stages_v1.json)                        */
  /* values 2, 4, 7, 8 are timing / tracing experiments that only a debug build of the library honours:
     include/facerec_b200_debug.h */
};

/* ---- lifecycle ---------------------------------------------------------------------------- */
int frb_version(void);
int frb_init(int device, frb_ctx** out);
int frb_destroy(frb_ctx* ctx);
const char* frb_last_error(frb_ctx* ctx); /* valid until the next call on ctx */
int frb_set_option(frb_ctx* ctx, int option, int value);
long long frb_launch_count(frb_ctx* ctx); /* kernels launched through ctx so far */
/* FRB_OPT_LAYER_TIMES: per-op rows (kind 0 conv / 1 dwconv / 2 maxpool / 3 avgpool / 4 l2norm, M, N, K, flop, ms) of the
 * LAST network forward; returns the row count (<= max_rows) or a negative frb_status. */
int frb_debug_layer_times(frb_ctx* ctx, double* h_rows, int max_rows);

/* Weights.  `h_blob` is the self-describing fp32 blob written by facerec_b200/weights.py
 * (magic "FRBW", entry table of name/shape/offset, then data): one "<layer>.w" [Cout,kh,kw,Cin]
 * and "<layer>.b" [Cout] per convolution with BatchNorm already folded on the host.
 * Replaces: model construction at facerec/extract.py:392 and facerec/detector.py:64. */
int frb_load_model(frb_ctx* ctx, int model_id, const void* h_blob, size_t bytes);

/* Bytes of the activation arena a forward of `model_id` needs for B frames of H x W (B faces for FRB_IRV1; H, W ignored
 * there), or -1 if the model is not loaded.  The library allocates the arena itself on the first forward of a shape and only
 * ever grows it: one warm-up forward per shape leaves the hot path free of allocations. */
long long frb_workspace_bytes(frb_ctx* ctx, int model_id, int B, int H, int W);

/* ---- per-frame path ----------------------------------------------------------------------- */

/* Layout of the 16-bit network input produced by K1. */
enum frb_layout {
  FRB_LAYOUT_NHWC8 = 0, /* [B,H,W,8]: 3 real channels + 5 zero (MobileNet0.25 stem)                        */
  FRB_LAYOUT_S2D16 = 1  /* [B,H/2,W/2+3,16]: 2x2 space-to-depth blocks (dy,dx,c4), 2 zero pixels left and 1
                           right = the stem's horizontal padding (ResNet-50 stem; odd H/W round up)           */
};
int frb_input_layout(int model_id);                              /* the layout frb_retina_forward expects */
size_t frb_input_elems(int layout, int B, int H, int W);         /* 16-bit elements of that buffer        */

/* K1  u8 HWC frames -> 16-bit network input: RetinaFace input = BGR minus (104,117,123), no scaling.
 * Replaces the input transform inside RetinaFace.predict (facerec/detector.py:68); frames arrive RGB
 * from facerec/extract.py:300, so src_is_rgb=1 swaps channels back. */
int frb_preprocess_u8(frb_ctx* ctx, const uint8_t* d_frames, int B, int H, int W, int src_is_rgb, int layout,
                      void* d_out16, void* stream);

/* K2-K5  RetinaFace forward -> loc [B,P,4], conf [B,P,2] (logits), landm [B,P,10], fp32.
 * d_in16 is the K1 output in frb_input_layout(model_id); H, W are the FRAME dimensions.
 * Replaces the network inside RetinaFace.predict (facerec/detector.py:64,68). */
int frb_retina_forward(frb_ctx* ctx, int model_id, const void* d_in16, int B, int H, int W,
                       float* d_loc, float* d_conf, float* d_landm, void* stream);
/* number of priors for an H x W input: 2 * sum over steps {8,16,32} of ceil(H/s)*ceil(W/s) */
int frb_num_priors(int H, int W);

/* K6  PriorBox decode + softmax + score>thr + compaction; priors are computed in-kernel.
 * cand row = x1,y1,x2,y2,score, 10 landmark coords (pixels).  Candidates beyond max_cand are
 * dropped but still counted in d_count (caller checks for overflow).
 * Replaces anchor decode / confidence filter inside RetinaFace.predict(img, threshold=0.95)
 * (facerec/detector.py:68). */
int frb_prior_decode_filter(frb_ctx* ctx, const float* d_loc, const float* d_conf, const float* d_landm, int B, int H,
                            int W, float score_thr, int max_cand, float* d_cand /*[B,max_cand,15]*/,
                            int32_t* d_count /*[B]*/, int32_t* d_prior_idx /*[B,max_cand]*/, void* stream);

/* K7  per-frame sort by score (descending, ties by lower prior index) + greedy IoU NMS.
 * d_keep holds indices into the frame's cand rows, in score order. */
int frb_nms(frb_ctx* ctx, const float* d_cand, const int32_t* d_count, const int32_t* d_prior_idx, int B, int max_cand,
            float iou_thr, int legacy_plus_one, int32_t* d_keep /*[B,max_cand]*/, int32_t* d_nkeep /*[B]*/,
            void* stream);

/* K7b  kept detections of every frame -> dense rows out[b, i, :] = cand[b, keep[b, i], :] for
 * i < min(nkeep[b], max_out); d_nout[b] = nkeep[b] (caller checks nout <= max_out).  Replaces the
 * list comprehension input of facerec/detector.py:70 (one small D2H per batch). */
int frb_gather_kept(frb_ctx* ctx, const float* d_cand, const int32_t* d_keep, const int32_t* d_nkeep, int B, int max_cand,
                    int max_out, float* d_out /*[B,max_out,15]*/, int32_t* d_nout /*[B]*/, void* stream);

/* K8  crop + bilinear resize (cv2.INTER_LINEAR u8 fixed-point semantics) to 160x160 +
 * fixed_image_standardization (x-127.5)/128 -> 16-bit NHWC with 8 channels (3 real + 5 zero).
 * d_boxes rows = (frame, x, y, w, h): ints AFTER the host-side cropBox clamping.
 * Replaces keras_facenet cropBox + the resize/standardise part of FaceNet.embeddings
 * (facerec/extract.py:161-166). */
int frb_crop_resize_std(frb_ctx* ctx, const uint8_t* d_frames_rgb, int B, int H, int W, const int32_t* d_boxes /*[F,5]*/,
                        int F, void* d_out16 /*[F,160,160,8]*/, void* stream);

/* K2,K9-K11  InceptionResnetV1 forward incl. last_linear + BatchNorm1d + L2 normalisation.
 * Replaces FaceNet.embeddings model.predict (facerec/extract.py:166) and the explicit
 * normalisation of facerec/utils/utils.py:20-22. */
int frb_irv1_forward(frb_ctx* ctx, const void* d_in16 /*[F,160,160,8]*/, int F, float* d_emb /*[F,512]*/, void* stream);

/* Generic K2 entry (unit tests, micro-benchmarks): one convolution with the fused epilogue.
 * h_w [Cout,kh,kw,Cin] fp32 and h_bias [Cout] are host pointers; the call prepares the layer,
 * launches, synchronises the stream and releases the layer.  act: 0 none, 1 relu, 2 leaky(slope).
 * res_mode: 0 none, 1 add d_res before activation ([B,Ho,Wo,Cout]), 2 add nearest-upsampled
 * d_res [B,rH,rW,Cout] after activation.  out_fp32: 0 -> 16-bit out, 1 -> fp32 out. */
int frb_conv2d(frb_ctx* ctx, const void* d_in16, int B, int H, int W, int Cin, const float* h_w, const float* h_bias,
               int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int act, float slope, const void* d_res16,
               int res_mode, int rH, int rW, void* d_out, int out_fp32, int iters, float* h_ms /*nullable*/,
               void* stream);

/* Scene-cut statistics of a batch of RGB frames (the per-frame work of SceneChangeDetector.update,
 * facerec/scene.py:81-133, called at facerec/extract.py:301).  Per frame b: d_hist[b] = 256-bin histogram of
 * floor(luminance) over the centre 2:1 crop (crop != 0), d_cdf[b] = its float64 cumulative distribution,
 * d_sums[b] = { sum |Y_b - Y_(b-1)|, sum |eq_b - eq_(b-1)| } over the cropped pixels, eq = cdf[floor(Y)] * 255.
 * Frame 0 is compared with d_prev_frame / d_prev_cdf (last frame of the previous batch; both NULL at the
 * start of a shard, then d_sums[0] stays 0).  The threshold cascade (scene.py:65-79) stays on the host. */
int frb_scene_stats(frb_ctx* ctx, const uint8_t* d_frames_rgb, int B, int H, int W, int crop, const uint8_t* d_prev_frame,
                    const double* d_prev_cdf, uint32_t* d_hist /*[B,256]*/, double* d_cdf /*[B,256]*/,
                    double* d_sums /*[B,2]*/, void* stream);

/* ---- embedding-distance stage ------------------------------------------------------------- */

/* K12  per-trajectory mean of per-frame embeddings (facerec/cluster.py:26-29). */
int frb_segment_mean(frb_ctx* ctx, const float* d_X /*[n,D]*/, const int32_t* d_offsets /*[S+1]*/, int S, int D,
                     float* d_out /*[S,D]*/, void* stream);

/* K13  Euclidean distances of rows [row0,row0+rows) of X against all of X (split-bf16 tensor-core
 * GEMM, fp32 accumulate, d = sqrt(max(|a|^2+|b|^2-2ab,0))).  Condensed layouts receive only the
 * j>i entries of those rows, at their scipy pdist offsets in a full-size condensed vector.
 * Replaces the implicit pdist inside scipy linkage (facerec/cluster.py:41,199). */
int frb_pdist(frb_ctx* ctx, const float* d_X, int N, int D, int row0, int rows, int layout, void* d_out, void* stream);

/* Complete-linkage clustering of the N x N fp32 distance matrix d_D (FRB_SQUARE_F32 output of frb_pdist; it is
 * overwritten), nearest-neighbour-chain algorithm with scipy's tie rules.  Outputs the N-1 merges in the order
 * they are found: d_merge_pairs[k] = (dying cluster, surviving cluster) as ORIGINAL point indices with
 * pairs[k][0] < pairs[k][1], d_merge_dist[k] = merge height.  The host sorts them by height (stable) and
 * relabels to get scipy's Z (facerec_b200/distance.py::gpu_linkage_complete).
 * Replaces the host side of scipy.cluster.hierarchy.linkage(y, "complete") (facerec/cluster.py:41,199). */
int frb_linkage_complete(frb_ctx* ctx, float* d_D /*[N,N]*/, int N, int32_t* d_merge_pairs /*[N-1,2]*/,
                         float* d_merge_dist /*[N-1]*/, void* stream);

/* K14  k smallest Euclidean distances (k <= 16), never materialising Q x M.  Ties -> lower DB index.
 * One bf16 tensor-core pass keeps 24 candidates per query, which are re-scored in fp32; queries whose k best cannot be
 * proven from that (worst-case rounding bound) are re-run through the exact split-bf16 pass.  The call synchronises
 * `stream` once (it reads the count of such queries).
 * Replaces KNeighborsClassifier.fit/kneighbors (facerec/classify_knn.py:212,227). */
int frb_knn_topk(frb_ctx* ctx, const float* d_Q, int Q, const float* d_DB, int M, int D, int k,
                 int32_t* d_idx /*[Q,k] ascending distance*/, float* d_dist /*[Q,k]*/, void* stream);

/* The `fit` half of KNeighborsClassifier.fit / predict_proba (facerec/classify_knn.py:212): prepares the database once
 * (power-of-two scaling, fp16 hi / lo planes, |x|^2) and keeps it inside ctx; later frb_knn_topk calls that pass the SAME
 * d_DB, M, D reuse it instead of preparing the database on every call (0.65 ms per 10^6 rows).  The caller must not change
 * the database in place without calling frb_knn_fit again.  Optional: frb_knn_topk works without it. */
int frb_knn_fit(frb_ctx* ctx, const float* d_DB, int M, int D, void* stream);

/* Diagnostics: how many queries of the LAST frb_knn_topk call on ctx could not be proven from the one-pass bf16 scan
 * and went through the exact split-bf16 pass (results are exact either way; this only costs time). */
int frb_knn_redo_count(frb_ctx* ctx);

/* K15  neighbour labels -> class histogram / k (predict_proba with uniform weights,
 * facerec/classify_knn.py:227). */
int frb_knn_vote(frb_ctx* ctx, const int32_t* d_idx, const int32_t* d_label_idx /*[M] in [0,n_classes)*/, int Q, int k,
                 int n_classes, float* d_proba /*[Q,n_classes]*/, void* stream);

/* ---- host-side SORT tracker ---------------------------------------------------
 * Replaces: class Sort / KalmanBoxTracker, facerec/sort.py:82-378, as driven by facerec/extract.py:312-327.
 * No GPU work and no frb_ctx: plain host state.  Detections are rows x1,y1,x2,y2,score (float64, as the reference
 * builds them at extract.py:319); ids are the running global detection ids of Sort.update (sort.py:251-253). */
typedef struct frb_sort frb_sort;
frb_sort* frb_sort_create(int max_age, int min_hits, double iou_threshold);           /* sort.py:204-222 */
void frb_sort_destroy(frb_sort* s);
int frb_sort_update(frb_sort* s, const double* h_dets /*[n,5]*/, int n, long long frame, long long* h_ids /*[n]*/);
int frb_sort_has_valid_tracker(frb_sort* s, long long detection_id);                  /* 1 / 0, -1: unknown id; sort.py:297-308 */
int frb_sort_get_detection_bbox(frb_sort* s, long long detection_id, double* h_box4); /* sort.py:321-331 */
void frb_sort_kill_trackers(frb_sort* s);                                             /* sort.py:333-336 */
int frb_sort_nan_removed(frb_sort* s);  /* trackers dropped for NaN predictions so far (the reference prints a warning each) */
/* sort.py:338-378: forget trackers idle for >= expiry_age frames (current_frame = LLONG_MAX: all); returns how many
 * VALID trajectories were popped; read them with the two calls below before the next pop. */
int frb_sort_pop_expired(frb_sort* s, long long expiry_age, long long current_frame);
int frb_sort_expired_info(frb_sort* s, int i, long long* first_frame, int* len);
int frb_sort_expired_boxes(frb_sort* s, int i, double* h_boxes /*[len,4]*/, unsigned char* h_detected /*[len]*/);

/* ---- features.jsonl writer helper ---------------------------------------------
 * JSON text of a float32 vector, byte-identical to Python's json.dumps([float(x) for x in v], separators=(",",":"))
 * (the "embeddings" value of a features line, facerec/extract.py:186-195).  Needs cap >= 2 + 33 * n; returns the
 * text length (no terminator) or -1. */
long long frb_format_f32_json(const float* h_v, int n, char* h_out, long long cap);

#ifdef __cplusplus
}
#endif
#endif /* FACEREC_B200_H_ */
