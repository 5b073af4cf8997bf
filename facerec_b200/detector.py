"""RetinaFace detector with the reference's `detect()` contract on the B200 path.

Drop-in for `RetinaFaceDetector` of facerec/detector.py:54-79 (and, through the shared result
schema of facerec/detector.py:3-13, for the `FaceNetDetector` selected at facerec/extract.py:418):

    detector = B200RetinaFaceDetector(min_face_size=20)
    faces = detector.detect(frame_rgb)      # [{"box": [x1,y1,x2,y2], "keypoints": {...}}, ...]

Same semantics as the reference wrapper: score threshold 0.95 (`predict(img, threshold=0.95)`,
detector.py:68), faces kept when min(w, h) >= min_face_size (detector.py:79), keypoints truncated with
int() (detector.py:72-77), plain Python floats/ints so the caller can `json.dump` them
(facerec/extract.py:191).  Addition: `detect_batch(frames)` runs a whole batch of equally sized
frames through one pass of the CUDA path (the reference calls detect() once per frame at batch 1).

All compute is in libfacerec_b200.so: u8->16-bit preprocess, RetinaFace forward on tcgen05 tensor
cores, PriorBox decode + score filter + compaction, per-frame NMS.  The host only formats results.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops, weights as W
from ._lib import Context, default_context

KEYPOINT_NAMES = ("left_eye", "right_eye", "nose", "mouth_left", "mouth_right")


class B200RetinaFaceDetector:
    def __init__(self, min_face_size: int = 0, backbone: str = "mnet025", weights: Optional[Dict[str, np.ndarray]] = None,
                 seed: int = 0, threshold: float = 0.95, nms_threshold: float = 0.4, max_candidates: int = 5000,
                 keep_top_k: int = 750, ctx: Optional[Context] = None, device: int = 0):
        """max_candidates / keep_top_k are the published RetinaFace pipeline's pre-NMS top_k (5000) and post-NMS
        keep_top_k (750): a frame with more candidates degrades to its highest-scoring ones (with a warning) instead of
        failing.  Up to 8192 candidates per frame are supported; the 1024 best enter the NMS suppression step."""
        if backbone not in ("mnet025", "r50"):
            raise ValueError("backbone must be 'mnet025' or 'r50'")
        self.min_size = min_face_size
        self.threshold = float(threshold)
        self.nms_threshold = float(nms_threshold)
        self.max_candidates = int(max_candidates)
        self.keep_top_k = int(keep_top_k)
        self._host_pool: Dict[tuple, list] = {}     # pinned result buffers, reused across batches (no allocation per batch)
        self.ctx = ctx if ctx is not None else default_context(device)
        self.device = torch.device("cuda", self.ctx.device)
        self.model_name = "retina_" + backbone
        self.model_id = W.MODEL_IDS[self.model_name]
        if weights is None:
            weights = W.synth_weights(self.model_name, seed)   # no checkpoints offline: seeded random init
        ops.load_model(self.ctx, self.model_id, W.write_blob(weights))
        self._pinned: Dict[tuple, torch.Tensor] = {}

    # ---- device-side pipeline -------------------------------------------------------------
    def detect_device(self, frames_u8: torch.Tensor):
        """frames_u8: [B,H,W,3] uint8 RGB on the GPU.  Returns device tensors (cand [B,C,15], keep [B,C], nkeep [B],
        count [B]); rows cand[b, keep[b,:nkeep[b]]] are the detections of frame b in score order."""
        B, H, W_, _ = frames_u8.shape
        x16 = ops.preprocess_u8(self.ctx, frames_u8, True, ops.input_layout(self.model_id))
        loc, conf, landm = ops.retina_forward(self.ctx, self.model_id, x16, H, W_)
        cand, count, pidx = ops.prior_decode_filter(self.ctx, loc, conf, landm, H, W_, self.threshold, self.max_candidates)
        keep, nkeep = ops.nms(self.ctx, cand, count, pidx, self.nms_threshold, True)
        return cand, keep, nkeep, count

    def _to_device(self, frames: np.ndarray) -> torch.Tensor:
        key = tuple(frames.shape)
        buf = self._pinned.get(key)
        if buf is None:
            buf = torch.empty(key, dtype=torch.uint8).pin_memory()
            self._pinned = {key: buf}
        buf.numpy()[...] = frames
        return buf.to(self.device, non_blocking=True)

    # ---- reference-facing API -------------------------------------------------------------
    def detect_batch_raw(self, frames) -> List[np.ndarray]:
        """Per frame: float32 [K,15] rows (x1,y1,x2,y2,score, 5 landmark x/y pairs), NMS order, before the min-size filter."""
        if isinstance(frames, torch.Tensor) and frames.is_cuda:
            dev = frames
        else:
            arr = np.ascontiguousarray(np.stack(list(frames)) if not isinstance(frames, np.ndarray) else frames)
            if arr.ndim != 4 or arr.shape[3] != 3 or arr.dtype != np.uint8:
                raise ValueError("frames must be [B,H,W,3] uint8")
            dev = self._to_device(arr)
        return self.detect_batch_collect(self.detect_batch_launch(dev))

    def _host_buffers(self, B: int):
        key = (B, self.keep_top_k)
        pool = self._host_pool.setdefault(key, [])
        if pool:
            return pool.pop()
        return (torch.empty((B, self.keep_top_k, 15), dtype=torch.float32, pin_memory=True),
                torch.empty((2, B), dtype=torch.int32, pin_memory=True))

    def detect_batch_launch(self, frames_dev: torch.Tensor):
        """Enqueue the whole detection pass for a GPU-resident batch plus the device-to-host copy of the kept detections
        (dense [B, keep_top_k, 15] rows via frb_gather_kept, into pooled pinned buffers) and return a handle; nothing
        here waits for the GPU."""
        cand, keep, nkeep, count = self.detect_device(frames_dev)
        dets, nout = ops.gather_kept(self.ctx, cand, keep, nkeep, self.keep_top_k)
        hd, hn = self._host_buffers(frames_dev.shape[0])
        hd.copy_(dets, non_blocking=True)
        hn[0].copy_(nout, non_blocking=True)
        hn[1].copy_(count, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return (hd, hn), ev, (dets, nout, count)   # the device tensors stay alive until the copies have run

    def detect_batch_collect(self, handle) -> List[np.ndarray]:
        (hd, hn), ev, _ = handle
        ev.synchronize()
        n, count = hn[0].numpy(), hn[1].numpy()
        if (count > self.max_candidates).any():
            warnings.warn(f"{int((count > self.max_candidates).sum())} frame(s) had more than max_candidates="
                          f"{self.max_candidates} priors above the score threshold (max {int(count.max())}): the overflow "
                          "was dropped before NMS", RuntimeWarning, stacklevel=2)
        if (n > self.keep_top_k).any():
            warnings.warn(f"a frame kept {int(n.max())} faces after NMS; only the keep_top_k={self.keep_top_k} highest-scoring "
                          "are returned", RuntimeWarning, stacklevel=2)
        out = [hd[b, :min(int(n[b]), self.keep_top_k)].numpy().copy() for b in range(len(n))]
        self._host_pool.setdefault((len(n), self.keep_top_k), []).append((hd, hn))
        return out

    def format(self, dets: np.ndarray) -> List[dict]:
        out = []
        for d in dets:
            x1, y1, x2, y2 = float(d[0]), float(d[1]), float(d[2]), float(d[3])
            if not np.isfinite(d[:15]).all():     # never produced by frb_nms; guards hand-made rows (NaN < min_size is False)
                continue
            if min(x2 - x1, y2 - y1) < self.min_size:
                continue
            kp = {n: (int(d[5 + 2 * t]), int(d[6 + 2 * t])) for t, n in enumerate(KEYPOINT_NAMES)}
            out.append({"box": [x1, y1, x2, y2], "keypoints": kp})
        return out

    def detect_batch(self, frames) -> List[List[dict]]:
        return [self.format(d) for d in self.detect_batch_raw(frames)]

    def detect(self, img: np.ndarray) -> List[dict]:
        assert len(img.shape) == 3 and img.shape[2] == 3
        return self.detect_batch(np.ascontiguousarray(img)[None])[0]
