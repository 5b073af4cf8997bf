"""Deterministic synthetic video for benchmarks and parity tests (there is no dataset offline).

`frame(seed, idx, H, W)` is counter-based: any frame of any shard is reproducible without generating
its predecessors, which is what frame sharding (facerec/extract.py:230-234) needs.  A clip is a
sequence of scenes separated by hard cuts; a scene is a smooth low-contrast background plus a few
high-contrast textured blobs ("faces") that stay put, with a faint global flicker so consecutive frames
are not byte-identical.  Random-init detectors respond to local contrast, so the blobs are what the
calibrated RetinaFace heads fire on, and static blobs let SORT form tracks (min_hits = 3).
"""
from __future__ import annotations

import functools

import numpy as np

SCENE_LEN = 192          # frames between hard cuts (8 s at 24 fps)


def _rng(*key) -> np.random.RandomState:
    h = 1469598103
    for k in key:
        h = (h * 1099511 + int(k) + 12345) % (2 ** 31 - 1)
    return np.random.RandomState(h)


@functools.lru_cache(maxsize=4)
def _scene(seed: int, scene: int, H: int, W: int, vary_level: bool = False) -> np.ndarray:
    r = _rng(seed, scene, H, W)
    level = 118.0
    if vary_level:   # hard cuts that also change the frame mean: food for the scene-cut detector tests
        level = float(_rng(seed, scene, 99).uniform(60, 190))
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    s = H / 1080.0
    img = np.empty((H, W, 3), np.float32)
    for c in range(3):
        img[..., c] = (level + 10 * np.sin(xx / (r.uniform(150, 400) * s) + r.uniform(0, 6))
                       * np.cos(yy / (r.uniform(150, 400) * s) + r.uniform(0, 6)))
    n_obj = r.randint(3, 9)
    for _ in range(n_obj):
        rad = r.uniform(25, 110) * s
        cx, cy = r.uniform(rad, W - rad), r.uniform(rad, H - rad)
        d2 = ((xx - cx) / rad) ** 2 + ((yy - cy) / (rad * r.uniform(1.0, 1.3))) ** 2
        mask = np.clip(1.5 - 1.5 * d2, 0, 1)
        fx, fy = r.uniform(0.15, 0.5) / max(s, 0.25), r.uniform(0.15, 0.5) / max(s, 0.25)
        for c in range(3):
            tex = 95 * np.sin(xx * fx + r.uniform(0, 6)) * np.sin(yy * fy + r.uniform(0, 6)) + r.uniform(-25, 25)
            img[..., c] += mask * tex
    return img


def frame(seed: int, idx: int, H: int, W: int, vary_level: bool = False, scene_len: int = SCENE_LEN) -> np.ndarray:
    """RGB uint8 [H,W,3]."""
    base = _scene(seed, idx // scene_len, H, W, vary_level)
    flicker = 2.0 * np.sin(idx * 0.7)
    return np.clip(base + flicker, 0, 255).astype(np.uint8)


def clip(seed: int, start: int, n: int, H: int, W: int) -> np.ndarray:
    return np.stack([frame(seed, start + i, H, W) for i in range(n)])


def embeddings(n: int, seed: int, centres: int = 2800, noise: float = 0.05, d: int = 512, mean_of: int = 0, device="cuda"):
    """Synthetic face embeddings for the distance stage: a mixture of `centres` identity directions on the
    unit sphere, rows = normalize(c + noise * g).  `mean_of` > 1 averages that many such rows (trajectory means are
    not unit norm, facerec/cluster.py:29).  Generated with torch on `device`, same seed -> same matrix on every rank."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    c = torch.nn.functional.normalize(torch.randn(centres, d, generator=g, device=device), dim=1)
    idx = torch.randint(0, centres, (n,), generator=g, device=device)
    x = torch.nn.functional.normalize(c[idx] + noise * torch.randn(n, d, generator=g, device=device), dim=1)
    for _ in range(max(mean_of, 1) - 1):
        x = x + torch.nn.functional.normalize(c[idx] + noise * torch.randn(n, d, generator=g, device=device), dim=1)
    if mean_of and mean_of > 1:
        x = x / mean_of
    return x.contiguous()
