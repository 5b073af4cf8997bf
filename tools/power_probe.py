"""Burst vs sustained: one conv layer timed over few and many back-to-back launches, with SM clock and board power
sampled meanwhile (nvidia-smi), in per-tap, two-k-block and halo modes.  Shows how much of the gap between the
per-layer table (short bursts) and the whole forward (power-capped) is clocks and how much is the kernel.
Usage: python tools/power_probe.py [--only merge1] > gpurun_out/power_probe.txt"""
import argparse
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facerec_b200 import ops, _lib  # noqa: E402
from tools.bench_conv import SHAPES  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows = []
        self.stop = False

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((float(o[0]), float(o[1])))
            except Exception:
                pass
            time.sleep(0.05)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="merge1")
    a = ap.parse_args()
    for name, H, W, Cin, Cout, k, s, p in SHAPES:
        if a.only not in name:
            continue
        for mode in ("kbps2", "kbps1", "halo"):
            ctx = _lib.Context(0)
            ctx.set_option(_lib.FRB_OPT_BN_CAP, 128)
            if mode == "halo":
                ctx.set_option(_lib.FRB_OPT_HALO, 1)
            if mode == "kbps1":
                ctx.set_option(_lib.FRB_OPT_KBPS2, 0)
            for batch, iters in ((8, 10), (8, 4000), (64, 3), (64, 500)):
                x = torch.randn(batch, H, W, Cin, device="cuda").half()
                w = (np.random.RandomState(0).standard_normal((Cout, k, k, Cin)) / np.sqrt(k * k * Cin)).astype(np.float32)
                time.sleep(1.0)                     # let the board cool back to burst clocks
                smp = Sampler()
                smp.start()
                out, ms = ops.conv2d(ctx, x, w, None, (s, s), (p, p), 1, 0.0, iters=iters)
                smp.stop = True
                smp.join()
                fl = 2.0 * batch * out.shape[1] * out.shape[2] * Cout * k * k * Cin
                clk = np.median([r[0] for r in smp.rows]) if smp.rows else 0
                pw = np.max([r[1] for r in smp.rows]) if smp.rows else 0
                print(f"{name:18s} {mode:6s} batch {batch:3d} iters {iters:5d}  {ms:8.3f} ms/launch  {fl/ms/1e9:7.1f} TFLOP/s  "
                      f"sm {clk:6.0f} MHz  max power {pw:5.0f} W  ({len(smp.rows)} samples)", flush=True)
                del x, out
            ctx.close()


if __name__ == "__main__":
    main()
