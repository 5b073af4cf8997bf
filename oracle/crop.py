"""Crop + resize + standardise restatement (facerec/extract.py:38-53,157-166).

`bbox_float_to_int` follows facerec/extract.py:38-53; `crop_box` follows keras_facenet.utils.cropBox
as called at facerec/extract.py:161-164; `resize_linear_u8` restates cv2.resize(INTER_LINEAR) for
uint8 (11-bit fixed-point coefficients) and is pinned against cv2 itself in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np


def bbox_float_to_int(bbox, max_w, max_h, padding=0.0):
    """facerec/extract.py:38-53: optional min(w,h)*padding, clip to the frame, round half to even."""
    x1, y1, x2, y2 = bbox
    if padding > 0:
        p = min(x2 - x1, y2 - y1) * padding
        x1, y1, x2, y2 = x1 - p, y1 - p, x2 + p, y2 + p
    clipped = np.clip(np.array([x1, y1, x2, y2]), 0, [max_w, max_h, max_w, max_h])
    return tuple(int(v) for v in np.round(clipped).astype(int))


def crop_rect(box_xywh, W, H, margin=16):
    """Rectangle that `image[y:y+h, x:x+w]` of cropBox selects, after numpy's far-edge clamping."""
    x, y, w, h = box_xywh
    x -= margin; y -= margin; w += 2 * margin; h += 2 * margin
    if x < 0:
        w += x; x = 0
    if y < 0:
        h += y; y = 0
    x2, y2 = min(x + w, W), min(y + h, H)
    return x, y, max(x2 - x, 0), max(y2 - y, 0)


def crop_box(image, box_xywh, margin=16):
    x, y, w, h = crop_rect(box_xywh, image.shape[1], image.shape[0], margin)
    return image[y:y + h, x:x + w]


def _coeffs(dsize, ssize, horizontal):
    scale = 1.0 / (float(dsize) / float(ssize))
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = f - s.astype(np.float32)
    if horizontal:
        lo = s < 0
        f[lo] = 0; s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0; s[hi] = ssize - 1
    a0 = np.clip(np.rint((np.float32(1) - f) * np.float32(2048)), -32768, 32767).astype(np.int32)
    a1 = np.clip(np.rint(f * np.float32(2048)), -32768, 32767).astype(np.int32)
    return s, a0, a1


def resize_linear_u8(img: np.ndarray, dsize=160) -> np.ndarray:
    """cv2.resize(img, (dsize, dsize), interpolation=cv2.INTER_LINEAR) for HxWxC uint8, bit-exact."""
    h, w = img.shape[:2]
    sx, ax0, ax1 = _coeffs(dsize, w, True)
    sy, by0, by1 = _coeffs(dsize, h, False)
    sx1 = np.minimum(sx + 1, w - 1)
    r0 = np.clip(sy, 0, h - 1)
    r1 = np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int32)
    rows0 = src[r0][:, sx] * ax0[None, :, None] + src[r0][:, sx1] * ax1[None, :, None]
    rows1 = src[r1][:, sx] * ax0[None, :, None] + src[r1][:, sx1] * ax1[None, :, None]
    out = (((by0[:, None, None] * (rows0 >> 4)) >> 16) + ((by1[:, None, None] * (rows1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def standardize(face_u8: np.ndarray) -> np.ndarray:
    """fixed_image_standardization: (x - 127.5) / 128, float32."""
    return ((face_u8.astype(np.float32) - np.float32(127.5)) / np.float32(128.0)).astype(np.float32)
