"""FaceNet InceptionResnetV1 embedder with the reference's `embeddings()` / `cropBox` contract.

Drop-in for `keras_facenet.FaceNet(key)` as used at facerec/extract.py:392,164-167:

    embedders = {k: B200FaceNet(k) for k in facenet_models}
    cropped = cropBox(img_np, detection={"box": [x, y, w, h]}, margin=16)
    vec = embedders[k].embeddings([cropped])[0].tolist()

`embeddings(crops)` = bilinear resize to 160x160 (cv2 INTER_LINEAR uint8 semantics) ->
fixed_image_standardization -> InceptionResnetV1 -> L2 normalisation, all in libfacerec_b200.so.
`embed_boxes(frames_on_gpu, rects)` is the batched form used by the extractor: the crop is fused into
the resize kernel, which reads the frames already resident in HBM (no per-face host round trip).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops, weights as W
from ._lib import Context, default_context

FACE_IMAGE_SIZE = 160


def cropBox(image: np.ndarray, detection: dict, margin: int = 0) -> np.ndarray:
    """keras_facenet.utils.cropBox as called at facerec/extract.py:164: absolute margin, clamp at 0,
    far edges clamped by slicing."""
    x, y, w, h = crop_rect(detection["box"], image.shape[1], image.shape[0], margin)
    return image[y:y + h, x:x + w]


def crop_rect(box_xywh, W_: int, H: int, margin: int = 0):
    """The rectangle cropBox selects, as (x, y, w, h) fully inside the W_ x H frame."""
    x, y, w, h = (int(v) for v in box_xywh)
    x -= margin
    y -= margin
    w += 2 * margin
    h += 2 * margin
    if x < 0:
        w += x
        x = 0
    if y < 0:
        h += y
        y = 0
    return x, y, max(min(x + w, W_) - x, 0), max(min(y + h, H) - y, 0)


class B200FaceNet:
    _loaded: Dict[int, tuple] = {}

    def __init__(self, key: str = "20180402-114759", weights: Optional[Dict[str, np.ndarray]] = None, seed: int = 0,
                 ctx: Optional[Context] = None, device: int = 0, max_batch: int = 1024):
        self.key = key
        self.ctx = ctx if ctx is not None else default_context(device)
        self.device = torch.device("cuda", self.ctx.device)
        self.max_batch = int(max_batch)
        if weights is None:
            weights = W.synth_weights("irv1", seed)
        ops.load_model(self.ctx, W.MODEL_IDS["irv1"], W.write_blob(weights))

    def embed_device(self, frames_u8: torch.Tensor, rects: torch.Tensor) -> torch.Tensor:
        """frames [B,H,W,3] uint8 RGB (GPU), rects [F,5] int32 (frame, x, y, w, h) (GPU) -> [F,512] fp32 unit rows (GPU)."""
        outs = []
        for i in range(0, rects.shape[0], self.max_batch):
            r = rects[i:i + self.max_batch].contiguous()
            # the network plan (arena, tile grids) is cached per batch size: round the face count up to a
            # power of two (>= 8) by repeating the first rectangle, and drop the padding rows afterwards
            n = r.shape[0]
            bucket = 8
            while bucket < n:
                bucket *= 2
            if bucket != n:
                r = torch.cat([r, r[:1].expand(bucket - n, 5)], 0).contiguous()
            faces = ops.crop_resize_std(self.ctx, frames_u8, r)
            outs.append(ops.irv1_forward(self.ctx, faces)[:n])
        if not outs:
            return torch.empty((0, 512), dtype=torch.float32, device=self.device)
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def embed_boxes(self, frames_u8: torch.Tensor, rects: np.ndarray) -> np.ndarray:
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 5)
        B, H, W_, _ = frames_u8.shape
        if len(rects):
            ok = ((rects[:, 0] >= 0) & (rects[:, 0] < B) & (rects[:, 1] >= 0) & (rects[:, 2] >= 0) & (rects[:, 3] > 0) &
                  (rects[:, 4] > 0) & (rects[:, 1] + rects[:, 3] <= W_) & (rects[:, 2] + rects[:, 4] <= H))
            if not ok.all():
                raise ValueError(f"crop rectangle outside its frame or empty: {rects[~ok][0].tolist()}")
        return self.embed_collect(self.embed_boxes_launch(frames_u8, rects, checked=True))

    def embed_boxes_launch(self, frames_u8: torch.Tensor, rects: np.ndarray, checked: bool = False):
        """Enqueue crop + resize + InceptionResnetV1 for `rects` and the copy of the embeddings to pinned host memory;
        returns a handle for embed_collect.  Nothing here waits for the GPU."""
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 5)
        if not checked and len(rects):
            B, H, W_, _ = frames_u8.shape
            ok = ((rects[:, 0] >= 0) & (rects[:, 0] < B) & (rects[:, 1] >= 0) & (rects[:, 2] >= 0) & (rects[:, 3] > 0) &
                  (rects[:, 4] > 0) & (rects[:, 1] + rects[:, 3] <= W_) & (rects[:, 2] + rects[:, 4] <= H))
            if not ok.all():
                raise ValueError(f"crop rectangle outside its frame or empty: {rects[~ok][0].tolist()}")
        rp = torch.empty(rects.shape, dtype=torch.int32, pin_memory=True)
        rp.numpy()[...] = rects
        emb = self.embed_device(frames_u8, rp.to(self.device, non_blocking=True))
        host = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True)
        host.copy_(emb, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev, (emb, rp)

    def embed_collect(self, handle) -> np.ndarray:
        host, ev, _ = handle
        ev.synchronize()
        return host.numpy()

    def embeddings(self, images: Sequence[np.ndarray]) -> np.ndarray:
        """images: list of HxWx3 uint8 crops of any size -> float32 [len, 512]."""
        images = list(images)
        if not images:
            return np.zeros((0, 512), np.float32)
        hmax = max(im.shape[0] for im in images)
        wmax = max(im.shape[1] for im in images)
        pack = np.zeros((len(images), hmax, wmax, 3), np.uint8)
        rects = np.zeros((len(images), 5), np.int32)
        for i, im in enumerate(images):
            if im.ndim != 3 or im.shape[2] != 3 or im.dtype != np.uint8 or im.shape[0] == 0 or im.shape[1] == 0:
                raise ValueError("each image must be a non-empty HxWx3 uint8 array")
            pack[i, :im.shape[0], :im.shape[1]] = im
            rects[i] = (i, 0, 0, im.shape[1], im.shape[0])
        return self.embed_boxes(torch.from_numpy(pack).to(self.device), rects)
