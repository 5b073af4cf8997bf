"""Thin torch-tensor wrappers over the C ABI (include/facerec_b200.h).

PyTorch is only the owner of device memory and streams here: every function passes raw
`data_ptr()`s and explicit shapes to libfacerec_b200.so and launches on the current stream.
Nothing in this module computes on the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import Context, default_context, ptr, current_stream

FACE = 160


def act_dtype(ctx: Context) -> torch.dtype:
    return torch.bfloat16 if getattr(ctx, "bf16", False) else torch.float16


def _dev(ctx: Context) -> torch.device:
    return torch.device("cuda", ctx.device)


def _need_cuda(t: torch.Tensor, name: str, dtype=None):
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must have dtype {dtype}, got {t.dtype}")


def load_model(ctx: Context, model_id: int, blob: bytes) -> None:
    buf = (C.c_char * len(blob)).from_buffer_copy(blob)
    ctx.check(ctx.lib.frb_load_model(ctx.handle, model_id, C.addressof(buf), len(blob)), "frb_load_model")


def num_priors(H: int, W: int) -> int:
    return int(_lib.load_library().frb_num_priors(H, W))


def input_layout(model_id: int) -> int:
    return int(_lib.load_library().frb_input_layout(model_id))


def preprocess_u8(ctx: Context, frames: torch.Tensor, src_is_rgb: bool = True, layout: int = _lib.FRB_LAYOUT_NHWC8) -> torch.Tensor:
    """[B,H,W,3] uint8 -> 16-bit network input (BGR minus mean): [B,H,W,8] or, for FRB_LAYOUT_S2D16,
    [B,H/2,W/2+3,16] (2x2 space-to-depth blocks with the stem's horizontal zero padding)."""
    _need_cuda(frames, "frames", torch.uint8)
    B, H, W, ch = frames.shape
    if ch != 3:
        raise ValueError("frames must be [B,H,W,3]")
    shape = (B, (H + 1) // 2, (W + 1) // 2 + 3, 16) if layout == _lib.FRB_LAYOUT_S2D16 else (B, H, W, 8)
    out = torch.empty(shape, dtype=act_dtype(ctx), device=frames.device)
    ctx.check(ctx.lib.frb_preprocess_u8(ctx.handle, ptr(frames), B, H, W, int(src_is_rgb), layout, ptr(out),
                                        current_stream()), "frb_preprocess_u8")
    return out


def retina_forward(ctx: Context, model_id: int, x16: torch.Tensor, H: Optional[int] = None, W: Optional[int] = None):
    """x16 = preprocess_u8(..., layout=input_layout(model_id)); H, W = frame size (needed for the S2D layout)."""
    _need_cuda(x16, "x16", act_dtype(ctx))
    B = x16.shape[0]
    if input_layout(model_id) == _lib.FRB_LAYOUT_S2D16:
        if H is None:
            H, W = 2 * x16.shape[1], 2 * (x16.shape[2] - 3)
        if tuple(x16.shape) != (B, (H + 1) // 2, (W + 1) // 2 + 3, 16):
            raise ValueError("x16 does not have the S2D16 layout of an HxW frame")
    else:
        if x16.shape[3] != 8:
            raise ValueError("x16 must be [B,H,W,8]")
        H, W = x16.shape[1], x16.shape[2]
    P = num_priors(H, W)
    loc = torch.empty((B, P, 4), dtype=torch.float32, device=x16.device)
    conf = torch.empty((B, P, 2), dtype=torch.float32, device=x16.device)
    landm = torch.empty((B, P, 10), dtype=torch.float32, device=x16.device)
    ctx.check(ctx.lib.frb_retina_forward(ctx.handle, model_id, ptr(x16), B, H, W, ptr(loc), ptr(conf), ptr(landm),
                                         current_stream()), "frb_retina_forward")
    return loc, conf, landm


def prior_decode_filter(ctx: Context, loc, conf, landm, H: int, W: int, score_thr: float = 0.95, max_cand: int = 1024):
    for t, n in ((loc, "loc"), (conf, "conf"), (landm, "landm")):
        _need_cuda(t, n, torch.float32)
    B = loc.shape[0]
    cand = torch.empty((B, max_cand, 15), dtype=torch.float32, device=loc.device)
    count = torch.empty((B,), dtype=torch.int32, device=loc.device)
    pidx = torch.empty((B, max_cand), dtype=torch.int32, device=loc.device)
    ctx.check(ctx.lib.frb_prior_decode_filter(ctx.handle, ptr(loc), ptr(conf), ptr(landm), B, H, W, float(score_thr),
                                              max_cand, ptr(cand), ptr(count), ptr(pidx), current_stream()),
              "frb_prior_decode_filter")
    return cand, count, pidx


def nms(ctx: Context, cand, count, pidx, iou_thr: float = 0.4, legacy_plus_one: bool = True):
    _need_cuda(cand, "cand", torch.float32)
    _need_cuda(count, "count", torch.int32)
    _need_cuda(pidx, "pidx", torch.int32)
    B, max_cand, _ = cand.shape
    keep = torch.empty((B, max_cand), dtype=torch.int32, device=cand.device)
    nkeep = torch.empty((B,), dtype=torch.int32, device=cand.device)
    ctx.check(ctx.lib.frb_nms(ctx.handle, ptr(cand), ptr(count), ptr(pidx), B, max_cand, float(iou_thr),
                              int(legacy_plus_one), ptr(keep), ptr(nkeep), current_stream()), "frb_nms")
    return keep, nkeep


def gather_kept(ctx: Context, cand, keep, nkeep, max_out: int = 128):
    B, max_cand, _ = cand.shape
    out = torch.empty((B, max_out, 15), dtype=torch.float32, device=cand.device)
    nout = torch.empty((B,), dtype=torch.int32, device=cand.device)
    ctx.check(ctx.lib.frb_gather_kept(ctx.handle, ptr(cand), ptr(keep), ptr(nkeep), B, max_cand, max_out, ptr(out), ptr(nout),
                                      current_stream()), "frb_gather_kept")
    return out, nout


def crop_resize_std(ctx: Context, frames_rgb: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """frames [B,H,W,3] uint8 RGB, boxes [F,5] int32 (frame,x,y,w,h inside the frame) -> [F,160,160,8] 16-bit."""
    _need_cuda(frames_rgb, "frames_rgb", torch.uint8)
    _need_cuda(boxes, "boxes", torch.int32)
    B, H, W, _ = frames_rgb.shape
    F = boxes.shape[0]
    out = torch.empty((F, FACE, FACE, 8), dtype=act_dtype(ctx), device=frames_rgb.device)
    ctx.check(ctx.lib.frb_crop_resize_std(ctx.handle, ptr(frames_rgb), B, H, W, ptr(boxes), F, ptr(out), current_stream()),
              "frb_crop_resize_std")
    return out


def irv1_forward(ctx: Context, faces16: torch.Tensor) -> torch.Tensor:
    _need_cuda(faces16, "faces16", act_dtype(ctx))
    F = faces16.shape[0]
    if tuple(faces16.shape[1:]) != (FACE, FACE, 8):
        raise ValueError("faces16 must be [F,160,160,8]")
    emb = torch.empty((F, 512), dtype=torch.float32, device=faces16.device)
    if F:
        ctx.check(ctx.lib.frb_irv1_forward(ctx.handle, ptr(faces16), F, ptr(emb), current_stream()), "frb_irv1_forward")
    return emb


def conv2d(ctx: Context, x16: torch.Tensor, w: np.ndarray, bias: Optional[np.ndarray], stride=(1, 1), pad=(0, 0),
           act: int = 0, slope: float = 0.0, res16: Optional[torch.Tensor] = None, res_mode: int = 0,
           out_fp32: bool = False, iters: int = 0):
    """Generic K2 entry: x16 [B,H,W,Cin] 16-bit, w [Cout,kh,kw,Cin] fp32 (host).  Returns (out, ms_per_launch)."""
    _need_cuda(x16, "x16", act_dtype(ctx))
    B, H, W, Cin = x16.shape
    w = np.ascontiguousarray(w, dtype=np.float32)
    Cout, kh, kw, cin2 = w.shape
    assert cin2 == Cin
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    Ho = (H + 2 * pad[0] - kh) // stride[0] + 1
    Wo = (W + 2 * pad[1] - kw) // stride[1] + 1
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32 if out_fp32 else act_dtype(ctx), device=x16.device)
    rH = rW = 0
    if res16 is not None:
        _need_cuda(res16, "res16", act_dtype(ctx))
        rH, rW = res16.shape[1], res16.shape[2]
    ms = C.c_float(0)
    ctx.check(ctx.lib.frb_conv2d(ctx.handle, ptr(x16), B, H, W, Cin, ptr(w), ptr(b), Cout, kh, kw, stride[0], stride[1],
                                 pad[0], pad[1], act, float(slope), ptr(res16), res_mode, rH, rW, ptr(out), int(out_fp32),
                                 iters, C.byref(ms) if iters > 0 else None, current_stream()), "frb_conv2d")
    return out, float(ms.value)


def segment_mean(ctx: Context, X: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    _need_cuda(X, "X", torch.float32)
    _need_cuda(offsets, "offsets", torch.int32)
    S = offsets.numel() - 1
    out = torch.empty((S, X.shape[1]), dtype=torch.float32, device=X.device)
    ctx.check(ctx.lib.frb_segment_mean(ctx.handle, ptr(X), ptr(offsets), S, X.shape[1], ptr(out), current_stream()),
              "frb_segment_mean")
    return out


def pdist(ctx: Context, X: torch.Tensor, row0: int = 0, rows: Optional[int] = None, layout: int = _lib.FRB_CONDENSED_F32,
          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Euclidean distances of rows [row0,row0+rows) against all rows of X (fp32 [N,D])."""
    _need_cuda(X, "X", torch.float32)
    N, D = X.shape
    rows = N - row0 if rows is None else rows
    if out is None:
        if layout in (_lib.FRB_SQUARE_F32, _lib.FRB_SQUARE_SYM_F32):
            out = torch.empty((rows, N), dtype=torch.float32, device=X.device)
        else:
            dt = torch.float64 if layout == _lib.FRB_CONDENSED_F64 else torch.float32
            out = torch.zeros((N * (N - 1) // 2,), dtype=dt, device=X.device)
    ctx.check(ctx.lib.frb_pdist(ctx.handle, ptr(X), N, D, row0, rows, layout, ptr(out), current_stream()), "frb_pdist")
    return out


def knn_topk(ctx: Context, Q: torch.Tensor, DB: torch.Tensor, k: int):
    _need_cuda(Q, "Q", torch.float32)
    _need_cuda(DB, "DB", torch.float32)
    nq, D = Q.shape
    idx = torch.empty((nq, k), dtype=torch.int32, device=Q.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
    ctx.check(ctx.lib.frb_knn_topk(ctx.handle, ptr(Q), nq, ptr(DB), DB.shape[0], D, k, ptr(idx), ptr(dist),
                                   current_stream()), "frb_knn_topk")
    return idx, dist


def knn_fit(ctx: Context, DB: torch.Tensor) -> None:
    """Prepare the database once (KNeighborsClassifier.fit); later knn_topk calls with the SAME tensor reuse it."""
    _need_cuda(DB, "DB", torch.float32)
    ctx.check(ctx.lib.frb_knn_fit(ctx.handle, ptr(DB), DB.shape[0], DB.shape[1], current_stream()), "frb_knn_fit")


def workspace_bytes(ctx: Context, model_id: int, B: int, H: int = 160, W: int = 160) -> int:
    return int(ctx.lib.frb_workspace_bytes(ctx.handle, model_id, B, H, W))


def knn_redo_count(ctx: Context) -> int:
    """Queries of the last knn_topk call that went through the exact split-bf16 pass (diagnostics)."""
    return int(ctx.lib.frb_knn_redo_count(ctx.handle))


def knn_vote(ctx: Context, idx: torch.Tensor, label_idx: torch.Tensor, n_classes: int) -> torch.Tensor:
    _need_cuda(idx, "idx", torch.int32)
    _need_cuda(label_idx, "label_idx", torch.int32)
    nq, k = idx.shape
    proba = torch.empty((nq, n_classes), dtype=torch.float32, device=idx.device)
    ctx.check(ctx.lib.frb_knn_vote(ctx.handle, ptr(idx), ptr(label_idx), nq, k, n_classes, ptr(proba), current_stream()),
              "frb_knn_vote")
    return proba
