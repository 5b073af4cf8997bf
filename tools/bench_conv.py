"""Micro-benchmark of the K2 conv kernel on the layer shapes of RetinaFace-R50 @1080p and
InceptionResnetV1 (per-launch CUDA-event time from frb_conv2d, TFLOP/s, fraction of measured peak).
Usage: python tools/bench_conv.py [--batch 8] [--bn-cap 128] > gpurun_out/conv_bench.txt"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facerec_b200 import ops, _lib  # noqa: E402

SHAPES = [
    # name, H, W, Cin, Cout, k, stride, pad
    ("r50.conv1 7x7s2", 1080, 1920, 8, 64, 7, 2, 3),
    ("r50.l1.conv1 1x1", 270, 480, 256, 64, 1, 1, 0),
    ("r50.l1.conv2 3x3", 270, 480, 64, 64, 3, 1, 1),
    ("r50.l1.conv3 1x1", 270, 480, 64, 256, 1, 1, 0),
    ("r50.l2.conv2 3x3", 135, 240, 128, 128, 3, 1, 1),
    ("r50.l2.conv3 1x1", 135, 240, 128, 512, 1, 1, 0),
    ("r50.l2.conv1 1x1", 135, 240, 512, 128, 1, 1, 0),
    ("r50.l3.conv2 3x3", 68, 120, 256, 256, 3, 1, 1),
    ("r50.l3.conv3 1x1", 68, 120, 256, 1024, 1, 1, 0),
    ("r50.l3.conv1 1x1", 68, 120, 1024, 256, 1, 1, 0),
    ("r50.l4.conv2 3x3", 34, 60, 512, 512, 3, 1, 1),
    ("r50.l4.conv3 1x1", 34, 60, 512, 2048, 1, 1, 0),
    ("fpn.merge1 3x3", 135, 240, 256, 256, 3, 1, 1),
    ("ssh1.A 3x3", 135, 240, 256, 192, 3, 1, 1),
    ("ssh1.B 3x3", 135, 240, 64, 128, 3, 1, 1),
    ("head1 1x1", 135, 240, 256, 32, 1, 1, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--bn-cap", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="", help="substring filter on the layer name")
    ap.add_argument("--no-halo", action="store_true")
    ap.add_argument("--halo", action="store_true")
    ap.add_argument("--kbps1", action="store_true")
    a = ap.parse_args()
    ctx = _lib.Context(0)
    ctx.set_option(_lib.FRB_OPT_BN_CAP, a.bn_cap)
    if a.halo:
        ctx.set_option(_lib.FRB_OPT_HALO, 1)
    if a.kbps1:
        ctx.set_option(_lib.FRB_OPT_KBPS2, 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops", 1590.0)
    print(f"# batch {a.batch} bn_cap {a.bn_cap} peak {peak} TFLOP/s")
    tot_f = tot_t = 0.0
    for name, H, W, Cin, Cout, k, s, p in SHAPES:
        if a.only and a.only not in name:
            continue
        x = torch.randn(a.batch, H, W, Cin, device="cuda").half()
        w = (np.random.RandomState(0).standard_normal((Cout, k, k, Cin)) / np.sqrt(k * k * Cin)).astype(np.float32)
        out, ms = ops.conv2d(ctx, x, w, None, (s, s), (p, p), 1, 0.0, iters=a.iters)
        Ho, Wo = out.shape[1], out.shape[2]
        fl = 2.0 * a.batch * Ho * Wo * Cout * k * k * Cin
        by = 2.0 * (x.numel() + out.numel())
        tf = fl / ms / 1e9
        print(f"{name:22s} M={a.batch*Ho*Wo:8d} N={Cout:5d} K={k*k*Cin:5d}  {ms:8.3f} ms  {tf:8.1f} TFLOP/s  {tf/peak*100:5.1f}% peak  {by/ms/1e6:7.0f} GB/s io")
        tot_f += fl
        tot_t += ms
        del x, out
    print(f"# sum: {tot_t:.2f} ms, {tot_f/tot_t/1e9:.1f} TFLOP/s aggregate")


if __name__ == "__main__":
    main()
