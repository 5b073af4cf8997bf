"""numpy restatement of RetinaFace post-processing and of the reference's detect() post-filter.

PriorBox / decode / softmax / threshold / NMS follow the published RetinaFace formulation
(the arithmetic sits in the un-vendored `retinaface` package called at
facerec/detector.py:68 with threshold=0.95).  `detect_from_raw` follows facerec/detector.py:66-79
(min-size filter, box passthrough, int() keypoints, key names).  All arithmetic is float32 in a fixed
operation order so the CUDA kernels (frb_prior_decode_filter, frb_nms) can match it to the ulp.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
MIN_SIZES = [[16, 32], [64, 128], [256, 512]]
STEPS = [8, 16, 32]
VAR0, VAR1 = F32(0.1), F32(0.2)
KEYPOINT_NAMES = ["left_eye", "right_eye", "nose", "mouth_left", "mouth_right"]


def num_priors(H: int, W: int) -> int:
    return sum(2 * math.ceil(H / s) * math.ceil(W / s) for s in STEPS)


def priorbox(H: int, W: int) -> np.ndarray:
    """[P,4] float32 (cx, cy, w, h), normalised; level-major, then row, col, anchor."""
    out = []
    for k, step in enumerate(STEPS):
        fh, fw = math.ceil(H / step), math.ceil(W / step)
        i, j = np.meshgrid(np.arange(fh, dtype=F32), np.arange(fw, dtype=F32), indexing="ij")
        cx = ((j + F32(0.5)) * F32(step)) / F32(W)
        cy = ((i + F32(0.5)) * F32(step)) / F32(H)
        lvl = np.empty((fh, fw, 2, 4), dtype=F32)
        for a, ms in enumerate(MIN_SIZES[k]):
            lvl[:, :, a, 0] = cx
            lvl[:, :, a, 1] = cy
            lvl[:, :, a, 2] = F32(ms) / F32(W)
            lvl[:, :, a, 3] = F32(ms) / F32(H)
        out.append(lvl.reshape(-1, 4))
    return np.concatenate(out, 0)


def face_scores(conf: np.ndarray) -> np.ndarray:
    """softmax(conf, -1)[..., 1] in float32 with the max-subtraction form."""
    conf = conf.astype(F32)
    m = conf.max(-1, keepdims=True)
    e = np.exp(conf - m)
    return (e[..., 1] / (e[..., 0] + e[..., 1])).astype(F32)


def decode(loc, landm, priors, H, W):
    """boxes [P,4] x1,y1,x2,y2 and landmarks [P,10] in pixels."""
    loc = loc.astype(F32)
    landm = landm.astype(F32)
    pc, ps = priors[:, :2], priors[:, 2:]
    cxcy = pc + (loc[:, :2] * VAR0) * ps
    wh = ps * np.exp(loc[:, 2:] * VAR1)
    x1y1 = cxcy - wh * F32(0.5)
    x2y2 = x1y1 + wh
    scale = np.array([W, H], dtype=F32)
    boxes = np.concatenate([x1y1 * scale, x2y2 * scale], 1).astype(F32)
    lm = np.empty_like(landm)
    for t in range(5):
        lm[:, 2 * t:2 * t + 2] = (pc + (landm[:, 2 * t:2 * t + 2] * VAR0) * ps) * scale
    return boxes, lm.astype(F32)


def nms(boxes, scores, prior_idx, iou_thr=0.4, plus_one=True):
    """Greedy IoU NMS; order = score descending, ties by lower prior index.  Returns kept positions."""
    n = len(scores)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.lexsort((prior_idx, -scores.astype(F32)))
    po = F32(1.0 if plus_one else 0.0)
    b = boxes.astype(F32)
    area = ((b[:, 2] - b[:, 0]) + po) * ((b[:, 3] - b[:, 1]) + po)
    keep = []
    removed = np.zeros(n, dtype=bool)
    for oi, i in enumerate(order):
        if removed[i]:
            continue
        keep.append(i)
        rest = order[oi + 1:]
        iw = np.maximum(F32(0), (np.minimum(b[i, 2], b[rest, 2]) - np.maximum(b[i, 0], b[rest, 0])) + po)
        ih = np.maximum(F32(0), (np.minimum(b[i, 3], b[rest, 3]) - np.maximum(b[i, 1], b[rest, 1])) + po)
        inter = iw * ih
        ovr = inter / ((area[i] + area[rest]) - inter)
        removed[rest[ovr > F32(iou_thr)]] = True
    return np.asarray(keep, dtype=np.int64)


def postprocess(loc, conf, landm, H, W, score_thr=0.95, iou_thr=0.4, plus_one=True):
    """One frame of raw head outputs -> dets [K,15] (x1,y1,x2,y2,score, 10 landmark coords), prior ids [K]."""
    priors = priorbox(H, W)
    scores = face_scores(conf)
    idx = np.nonzero(scores > F32(score_thr))[0]
    boxes, lm = decode(loc[idx], landm[idx], priors[idx], H, W)
    keep = nms(boxes, scores[idx], idx, iou_thr, plus_one)
    dets = np.concatenate([boxes[keep], scores[idx][keep, None], lm[keep]], 1).astype(F32)
    return dets, idx[keep]


def detect_from_dets(dets: np.ndarray, min_size: int = 0):
    """facerec/detector.py:70-79: min-size filter, box passthrough, int()-truncated keypoints."""
    out = []
    for d in dets:
        x1, y1, x2, y2 = (float(v) for v in d[:4])
        if min(x2 - x1, y2 - y1) < min_size:
            continue
        kp = {n: (int(d[5 + 2 * t]), int(d[6 + 2 * t])) for t, n in enumerate(KEYPOINT_NAMES)}
        out.append({"box": [x1, y1, x2, y2], "keypoints": kp})
    return out
