"""Where one conv_tma launch waits (FRB_OPT_TRACE = 2): stall cycles of the epilogue (thread 0), MMA issuer and TMA
producer, averaged over CTAs, as fractions of each role's lifetime.
Usage: python tools/stall_conv.py --shape B,H,W,Cin,Cout,k,stride,pad [--res] [--bn-cap 128]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--shape", action="append", default=[])
ap.add_argument("--res", action="store_true")
ap.add_argument("--bn-cap", type=int, default=256)
ap.add_argument("--halo", type=int, default=0, help="FRB_OPT_HALO value")
ap.add_argument("--max-stages", type=int, default=64)
ap.add_argument("--kbps1", action="store_true")
ap.add_argument("--skip-loads", action="store_true")
ap.add_argument("--no-plain", action="store_true", help="skip the uninstrumented timing run")
ap.add_argument("--slice", type=int, default=1, help="FRB_OPT_WEIGHT_SLICE")
ap.add_argument("--issue", type=int, default=1, help="FRB_OPT_DUAL_ISSUE: 0 warp A only, 1 both, 2 warp B only")
a = ap.parse_args()
for shape in a.shape:
    f = [int(v) for v in shape.split(",")]
    if len(f) == 10:                      # B,H,W,Cin,Cout,kh,kw,stride,ph,pw
        B, H, W, Cin, Cout, kh, kw, s, ph, pw = f
    else:
        B, H, W, Cin, Cout, k, s, p = f
        kh = kw = k
        ph = pw = p
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    w = (np.random.RandomState(0).standard_normal((Cout, kh, kw, Cin)) / np.sqrt(kh * kw * Cin)).astype(np.float32)
    Ho, Wo = (H + 2 * ph - kh) // s + 1, (W + 2 * pw - kw) // s + 1
    res = torch.randn(B, Ho, Wo, Cout, device="cuda").half() if a.res else None
    plain = _lib.Context(0)
    plain.set_option(_lib.FRB_OPT_BN_CAP, a.bn_cap)
    plain.set_option(_lib.FRB_OPT_HALO, a.halo)
    plain.set_option(_lib.FRB_OPT_MAX_STAGES, a.max_stages)
    plain.set_option(_lib.FRB_OPT_KBPS2, 0 if a.kbps1 else 1)
    plain.set_option(_lib.FRB_OPT_SKIP_LOADS, int(a.skip_loads))
    plain.set_option(_lib.FRB_OPT_DUAL_ISSUE, a.issue)
    plain.set_option(_lib.FRB_OPT_WEIGHT_SLICE, a.slice)
    ms0 = 1.0
    if not a.no_plain:
        ops.conv2d(plain, x, w, None, (s, s), (ph, pw), 1, 0.0, res, 1 if a.res else 0, iters=20)       # clocks up
        _, ms0 = ops.conv2d(plain, x, w, None, (s, s), (ph, pw), 1, 0.0, res, 1 if a.res else 0, iters=20)
    plain.close()
    ctx = _lib.Context(0)
    ctx.set_option(_lib.FRB_OPT_BN_CAP, a.bn_cap)
    ctx.set_option(_lib.FRB_OPT_TRACE, 2)
    ctx.set_option(_lib.FRB_OPT_HALO, a.halo)
    ctx.set_option(_lib.FRB_OPT_MAX_STAGES, a.max_stages)
    ctx.set_option(_lib.FRB_OPT_KBPS2, 0 if a.kbps1 else 1)
    ctx.set_option(_lib.FRB_OPT_SKIP_LOADS, int(a.skip_loads))
    ctx.set_option(_lib.FRB_OPT_DUAL_ISSUE, a.issue)
    ctx.set_option(_lib.FRB_OPT_WEIGHT_SLICE, a.slice)
    _, ms = ops.conv2d(ctx, x, w, None, (s, s), (ph, pw), 1, 0.0, res, 1 if a.res else 0, iters=1)
    buf = np.zeros(1 + 3 * 4000, np.int64)
    ctx.check(ctx.lib.frb_debug_trace(ctx.handle, buf.ctypes.data, 4000), "trace")
    c = buf[1:1 + 148 * 16].reshape(148, 16).astype(np.float64)
    c = c[c[:, 0] > 0]
    m = c.mean(0)
    fl = 2.0 * B * Ho * Wo * Cout * kh * kw * Cin
    print(f"== {shape}{' +res' if a.res else ''}: {ms0:.3f} ms plain ({fl / ms0 / 1e9:.0f} TFLOP/s), {ms:.3f} ms counted; "
          f"{m[5]:.1f} tiles per CTA, {m[0] / max(m[5], 1):.0f} cycles per tile")
    print(f"   epilogue t0 : accumulator wait {m[1] / m[0] * 100:5.1f}%  residual wait {m[2] / m[0] * 100:5.1f}%  "
          f"staging-slab (store drain) wait {m[3] / m[0] * 100:5.1f}%  barriers {m[4] / m[0] * 100:5.1f}%")
    print(f"   MMA issuer  : TMEM-buffer wait {m[7] / m[6] * 100:5.1f}%  operand (full) wait {m[8] / m[6] * 100:5.1f}%")
    print(f"   TMA producer: ring-slot (empty) wait {m[10] / m[9] * 100:5.1f}%  residual-set wait {m[11] / m[9] * 100:5.1f}%", flush=True)
    ctx.close()
