"""One conv shape timed with and without CTA pairs on the same box (frb_conv2d, 20 back-to-back launches each).
Usage: python tools/conv_ab.py B,H,W,Cin,Cout,kh,kw,stride,ph,pw [...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib

for shape in sys.argv[1:]:
    B, H, W, Cin, Cout, kh, kw, s, ph, pw = (int(v) for v in shape.split(","))
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    w = (np.random.RandomState(0).standard_normal((Cout, kh, kw, Cin)) / np.sqrt(kh * kw * Cin)).astype(np.float32)
    Ho, Wo = (H + 2 * ph - kh) // s + 1, (W + 2 * pw - kw) // s + 1
    fl = 2.0 * B * Ho * Wo * Cout * kh * kw * Cin
    out = {}
    for pairs in (0, 1, 0, 1):
        ctx = _lib.Context(0)
        ctx.set_option(_lib.FRB_OPT_CTA2, pairs)
        ops.conv2d(ctx, x, w, None, (s, s), (ph, pw), 1, 0.0, iters=10)
        _, ms = ops.conv2d(ctx, x, w, None, (s, s), (ph, pw), 1, 0.0, iters=20)
        out.setdefault(pairs, []).append(ms)
        ctx.close()
    print(f"{shape}: single {min(out[0]):.3f} ms ({fl / min(out[0]) / 1e9:.0f} TFLOP/s)  pairs {min(out[1]):.3f} ms ({fl / min(out[1]) / 1e9:.0f} TFLOP/s)", flush=True)
