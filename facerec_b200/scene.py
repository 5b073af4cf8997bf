"""Scene-cut detector of the per-frame loop (facerec/extract.py:301 -> facerec/scene.py:81-140), host side.

Same statistics and threshold cascade as the reference (luminance, centre 2:1 crop, 256-bin histogram
equalisation, MAFD / SDMAFD / ADFV, facerec/scene.py:42-133), restated with one `bincount` over the
truncated luminance instead of `np.histogram` over a flattened float copy (identical bin counts: every
luminance value v lands in bin floor(v)).  It is not part of the detect+embed metric; SURVEY 8(f1) lists
a GPU version as the next step because at 1080p this host code costs tens of ms per frame.
"""
from __future__ import annotations

import numpy as np


class SceneChangeDetector:
    def __init__(self, grayscale: bool = False, crop: bool = True, movie_id: int = 0):
        self.grayscale = grayscale
        self.crop = crop
        self.movie_id = movie_id
        self.frame_counter = 0
        self.prev_img = None
        self.prev_img_eq = None
        self.prev_mafd_eq = None
        self.prev_fv_eq = None

    @staticmethod
    def check(mafd, mafd_eq, sdmafd_eq, adfv_eq) -> bool:
        """Threshold cascade of facerec/scene.py:65-79."""
        if mafd < 14 or mafd_eq < 40:
            return False
        if 58 < mafd_eq < 100 and mafd < 100 and adfv_eq > 23:
            return True
        if mafd_eq < 85 and mafd > 170:
            return True
        if adfv_eq < 2 or sdmafd_eq < 5:
            return False
        return bool(mafd_eq > 50 and mafd > 35 and sdmafd_eq > 50 and adfv_eq > 50)

    def update(self, img: np.ndarray) -> bool:
        assert img.ndim == 3 and img.shape[2] == 3
        self.frame_counter += 1
        if self.grayscale:
            y = img[..., 0].astype(np.float32)
        else:
            y = 0.299 * img[:, :, 0] + 0.587 * img[:, :, 1] + 0.114 * img[:, :, 2]
        h, w = y.shape
        if self.crop and w / h < 2:
            inset = int((h - 0.5 * w) / 2)
            y = y[inset:-inset, :]
        k = y.astype(np.uint8)
        cdf = np.cumsum(np.bincount(k.ravel(), minlength=256) / k.size)
        y_eq = cdf[k] * 255
        cut = False
        if self.prev_img is not None:
            mafd = np.abs(y - self.prev_img).mean()
            mafd_eq = np.abs(y_eq - self.prev_img_eq).mean()
            sdmafd_eq = None if self.prev_mafd_eq is None else mafd_eq - self.prev_mafd_eq
            fv_eq = np.abs(y_eq - mafd_eq).mean()
            adfv_eq = None if self.prev_fv_eq is None else np.abs(fv_eq - self.prev_fv_eq)
            self.prev_fv_eq = fv_eq
            self.prev_mafd_eq = mafd_eq
            if sdmafd_eq is not None and adfv_eq is not None:
                cut = self.check(mafd, mafd_eq, sdmafd_eq, adfv_eq)
        self.prev_img = y
        self.prev_img_eq = y_eq
        return cut


class B200SceneChangeDetector:
    """Batched GPU version: the pixel work of `update` runs in libfacerec_b200.so (frb_scene_stats) on frames already
    resident in HBM; the scalar recurrences (SDMAFD, ADFV) and the threshold cascade stay on the host.
    `update_batch(frames_u8_cuda)` returns one bool per frame, equal to calling `update` frame by frame."""

    def __init__(self, grayscale: bool = False, crop: bool = True, movie_id: int = 0, ctx=None):
        import torch
        from ._lib import default_context
        if grayscale:
            raise ValueError("grayscale=True is not implemented on the GPU path (the extractor uses grayscale=False)")
        self.crop = crop
        self.movie_id = movie_id
        self.ctx = ctx if ctx is not None else default_context(0)
        self.frame_counter = 0
        self.prev_frame = None      # last frame of the previous batch (device)
        self.prev_cdf = None
        self.prev_mafd_eq = None
        self.prev_fv_eq = None
        self._torch = torch

    def update_batch(self, frames):
        return self.update_collect(self.stats_launch(frames))

    def stats_launch(self, frames):
        """Enqueue the pixel statistics of a batch (and their copies to pinned host memory); batches must be launched
        in clip order, the kernel reads the previous batch's last frame / CDF."""
        torch = self._torch
        from ._lib import ptr, current_stream
        B, H, W, _ = frames.shape
        dev = frames.device
        hist = torch.empty((B, 256), dtype=torch.int32, device=dev)
        cdf = torch.empty((B, 256), dtype=torch.float64, device=dev)
        sums = torch.empty((B, 2), dtype=torch.float64, device=dev)
        self.ctx.check(self.ctx.lib.frb_scene_stats(self.ctx.handle, ptr(frames), B, H, W, int(self.crop), ptr(self.prev_frame),
                                                    ptr(self.prev_cdf), ptr(hist), ptr(cdf), ptr(sums), current_stream()),
                       "frb_scene_stats")
        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (hist, cdf, sums)]
        for hbuf, t in zip(host, (hist, cdf, sums)):
            hbuf.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.prev_frame = frames[B - 1].clone()
        self.prev_cdf = cdf[B - 1].clone()
        return host, ev, (hist, cdf, sums)

    def update_collect(self, handle):
        """The scalar recurrences and the threshold cascade over a launched batch: one bool per frame."""
        host, ev, _ = handle
        ev.synchronize()
        h = host[0].numpy().astype(np.float64)
        c = host[1].numpy()
        s = host[2].numpy()
        B = h.shape[0]
        n = h[0].sum()
        cuts = []
        for b in range(B):
            self.frame_counter += 1
            cut = False
            if self.frame_counter > 1:
                mafd = s[b, 0] / n
                mafd_eq = s[b, 1] / n
                sdmafd_eq = None if self.prev_mafd_eq is None else mafd_eq - self.prev_mafd_eq
                fv_eq = float((h[b] * np.abs(c[b] * 255 - mafd_eq)).sum() / n)     # mean |eq - mafd_eq| via the histogram
                adfv_eq = None if self.prev_fv_eq is None else abs(fv_eq - self.prev_fv_eq)
                self.prev_fv_eq = fv_eq
                self.prev_mafd_eq = mafd_eq
                if sdmafd_eq is not None and adfv_eq is not None:
                    cut = SceneChangeDetector.check(mafd, mafd_eq, sdmafd_eq, adfv_eq)
            cuts.append(bool(cut))
        return cuts
