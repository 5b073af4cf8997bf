"""Batched detect -> crop -> embed over frames resident in HBM: the hot path of facerec/extract.py's
per-frame loop (extract.py:279-342) without the tracker in between.

    pipe = FacePipeline(backbone="r50")
    out = pipe.detect_embed(frames_u8)          # host or CUDA [B,H,W,3] uint8 RGB

Per batch: K1 preprocess -> RetinaFace forward -> K6 decode/filter -> K7 NMS -> K7b gather (one small D2H)
-> host box arithmetic of extract.py:160-164 (bbox_float_to_int, xywh, cropBox margin 16) -> K8 crop/
resize/standardise -> InceptionResnetV1 forward + L2 norm -> embeddings.  Faces are embedded on frames
whose index satisfies `index % save_every == 0`, the reference's gate (extract.py:336,349); the
extractor (facerec_b200/extract.py) additionally applies the SORT validity gate and Kalman boxes.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import ops
from ._lib import Context, default_context
from .detector import B200RetinaFaceDetector
from .embedder import B200FaceNet

MARGIN = 16   # int(0.1 * 160), facerec/extract.py:163


def bbox_float_to_int(bbox_float, max_w, max_h, padding: float = 0.0) -> List[int]:
    """facerec/extract.py:38-53 in float32: optional min(w,h)*padding, clip to the frame, round half to even."""
    b = np.array(bbox_float, dtype=np.float32)
    pad = padding * np.minimum(b[2] - b[0], b[3] - b[1])
    b += np.array([-pad, -pad, pad, pad])
    b = np.minimum(np.maximum(b, [0, 0, 0, 0]), [max_w, max_h, max_w, max_h])
    return [int(c) for c in np.round(b)]


def boxes_to_rects(boxes: np.ndarray, frame_idx: np.ndarray, W: int, H: int, margin: int = MARGIN) -> np.ndarray:
    """Vectorised extract.py:160-164 for float boxes [n,4]: tight int box -> xywh -> cropBox rectangle.
    Returns int32 [n,5] rows (frame, x, y, w, h) inside the frame; w or h may be <= 0 for degenerate boxes."""
    if len(boxes) == 0:
        return np.zeros((0, 5), np.int32)
    b = np.asarray(boxes, dtype=np.float32).copy()
    b = np.minimum(np.maximum(b, np.float32(0)), np.array([W, H, W, H], dtype=np.float32))
    t = np.round(b).astype(np.int64)                       # tight box x1,y1,x2,y2 (half-to-even)
    x = t[:, 0] - margin
    y = t[:, 1] - margin
    w = (t[:, 2] - t[:, 0]) + 2 * margin
    h = (t[:, 3] - t[:, 1]) + 2 * margin
    w = np.where(x < 0, w + x, w)
    x = np.maximum(x, 0)
    h = np.where(y < 0, h + y, h)
    y = np.maximum(y, 0)
    w = np.minimum(x + w, W) - x
    h = np.minimum(y + h, H) - y
    return np.stack([np.asarray(frame_idx, np.int64), x, y, w, h], 1).astype(np.int32)


@dataclass
class BatchResult:
    dets: List[np.ndarray]          # per frame float32 [K,15], NMS order, after the min-size filter
    face_frame: np.ndarray          # [F] frame index (within the batch) of every embedded face
    face_det: np.ndarray            # [F] row of dets[face_frame] the face came from
    rects: np.ndarray               # [F,5] int32 crop rectangles (frame, x, y, w, h)
    embeddings: np.ndarray          # [F,512] float32 unit rows


class FacePipeline:
    def __init__(self, backbone: str = "r50", min_face_size: int = 20, save_every: int = 5, seed: int = 0,
                 ctx: Optional[Context] = None, device: int = 0, max_faces_per_frame: int = 750,
                 det_weights=None, emb_weights=None):
        self.ctx = ctx if ctx is not None else default_context(device)
        self.device = torch.device("cuda", self.ctx.device)
        self.detector = B200RetinaFaceDetector(min_face_size, backbone, det_weights, seed, ctx=self.ctx)
        self.embedder = B200FaceNet(weights=emb_weights, seed=seed, ctx=self.ctx)
        self.save_every = int(save_every)
        self.max_out = int(max_faces_per_frame)
        self.detector.keep_top_k = self.max_out          # the launch / collect halves of the detector use its own cap
        self.min_size = min_face_size
        self._pin_in = None
        self._pin_det = None

    # ---- stages ---------------------------------------------------------------------------
    def detect_device(self, frames_u8: torch.Tensor):
        """-> (dets [B,max_out,15] fp32, counts [B] int32) on the device."""
        cand, keep, nkeep, count = self.detector.detect_device(frames_u8)
        dets, nout = ops.gather_kept(self.ctx, cand, keep, nkeep, self.max_out)
        return dets, nout, count

    def _dets_to_host(self, dets: torch.Tensor, nout: torch.Tensor, count: torch.Tensor):
        B = dets.shape[0]
        if self._pin_det is None or self._pin_det[0].shape[0] != B:
            self._pin_det = (torch.empty(tuple(dets.shape), dtype=torch.float32).pin_memory(),
                             torch.empty((2, B), dtype=torch.int32).pin_memory())
        pd, pn = self._pin_det
        pd.copy_(dets, non_blocking=True)
        pn[0].copy_(nout, non_blocking=True)
        pn[1].copy_(count, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n, c = pn[0].numpy(), pn[1].numpy()
        if (c > self.detector.max_candidates).any() or (n > self.max_out).any():
            warnings.warn(f"detection overflow: {int(c.max())} candidates / {int(n.max())} faces in one frame; kept the "
                          f"{self.detector.max_candidates} / {self.max_out} highest-scoring", RuntimeWarning, stacklevel=2)
        out = []
        for b in range(B):
            d = pd[b, :min(int(n[b]), self.max_out)].numpy()
            if self.min_size > 0 and len(d):
                d = d[np.minimum(d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]) >= self.min_size]
            out.append(d.copy())
        return out

    def select_faces(self, dets: List[np.ndarray], first_index: int, W: int, H: int):
        ff, fd, boxes = [], [], []
        for b, d in enumerate(dets):
            if (first_index + b) % self.save_every != 0:
                continue
            for i in range(len(d)):
                ff.append(b); fd.append(i); boxes.append(d[i, :4])
        ff = np.asarray(ff, np.int64)
        fd = np.asarray(fd, np.int64)
        rects = boxes_to_rects(np.asarray(boxes, np.float32).reshape(-1, 4), ff, W, H)
        ok = (rects[:, 3] > 0) & (rects[:, 4] > 0) if len(rects) else np.zeros(0, bool)
        return ff[ok], fd[ok], rects[ok]

    # ---- whole step ------------------------------------------------------------------------
    def step_device(self, frames_u8: torch.Tensor, first_index: int = 0, want_host_embeddings: bool = True) -> BatchResult:
        """One pass of the hot path over a batch already resident in HBM."""
        B, H, W, _ = frames_u8.shape
        dets_d, nout, count = self.detect_device(frames_u8)
        dets = self._dets_to_host(dets_d, nout, count)
        ff, fd, rects = self.select_faces(dets, first_index, W, H)
        if len(rects):
            emb = self.embedder.embed_device(frames_u8, torch.from_numpy(rects).to(self.device, non_blocking=True))
            emb_h = emb.cpu().numpy() if want_host_embeddings else emb
        else:
            emb_h = np.zeros((0, 512), np.float32)
        return BatchResult(dets, ff, fd, rects, emb_h)

    # ---- the same step in two halves, so that two batches can be in flight (what facerec_b200.extract does) -----------
    def step_launch(self, frames_u8: torch.Tensor, first_index: int = 0):
        """First half of a step over a GPU-resident batch: enqueue detection and the copy of the kept detections to pinned
        host memory.  Nothing here waits for the GPU."""
        return self.detector.detect_batch_launch(frames_u8), frames_u8, first_index

    def step_embed(self, handle):
        """Second half: wait for THIS batch's detections only, pick the faces on the host, enqueue crop + embed and the copy
        of the embeddings.  Work enqueued after step_launch(next batch) runs behind it, so the GPU stays busy while the host
        is in here."""
        dh, frames_u8, first_index = handle
        _, H, W, _ = frames_u8.shape
        dets = []
        for d in self.detector.detect_batch_collect(dh):
            if self.min_size > 0 and len(d):
                d = d[np.minimum(d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]) >= self.min_size]
            dets.append(d)
        ff, fd, rects = self.select_faces(dets, first_index, W, H)
        eh = self.embedder.embed_boxes_launch(frames_u8, rects, checked=True) if len(rects) else None
        return dets, ff, fd, rects, eh

    def step_collect(self, handle) -> BatchResult:
        dets, ff, fd, rects, eh = handle
        emb = self.embedder.embed_collect(eh).copy() if eh is not None else np.zeros((0, 512), np.float32)
        return BatchResult(dets, ff, fd, rects, emb)

    def run_pipelined(self, batches):
        """batches: iterable of (frames_u8 on the GPU, first_index).  Yields one BatchResult per batch, in order, equal to
        step_device's, with two batches in flight: detection of batch k+1 is enqueued before the host collects batch k."""
        prev = pend = None
        for frames_u8, first_index in batches:
            h = self.step_launch(frames_u8, first_index)
            if prev is not None:
                h2 = self.step_embed(prev)
                if pend is not None:
                    yield self.step_collect(pend)
                pend = h2
            prev = h
        if prev is not None:
            h2 = self.step_embed(prev)
            if pend is not None:
                yield self.step_collect(pend)
            yield self.step_collect(h2)

    def to_device(self, frames: np.ndarray) -> torch.Tensor:
        if self._pin_in is None or tuple(self._pin_in.shape) != tuple(frames.shape):
            self._pin_in = torch.empty(tuple(frames.shape), dtype=torch.uint8).pin_memory()
        self._pin_in.numpy()[...] = frames
        return self._pin_in.to(self.device, non_blocking=True)

    def detect_embed(self, frames, first_index: int = 0) -> BatchResult:
        """Host-facing entry: frames [B,H,W,3] uint8 RGB (numpy or CUDA tensor)."""
        if isinstance(frames, torch.Tensor):
            dev = frames if frames.is_cuda else frames.to(self.device)
        else:
            dev = self.to_device(np.ascontiguousarray(frames))
        return self.step_device(dev, first_index)
