"""Timeline of CTA 0 of one conv_tma launch (FRB_OPT_TRACE): per-tile gaps between the pipeline roles."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="8,270,480,64,64,3,1,1")   # B,H,W,Cin,Cout,k,stride,pad
ap.add_argument("--no-halo", action="store_true")
ap.add_argument("--res", action="store_true", help="add a residual tile (ResNet conv3 + skip)")
a = ap.parse_args()
B, H, W, Cin, Cout, k, s, p = (int(v) for v in a.shape.split(","))
ctx = _lib.Context(0)
ctx.set_option(_lib.FRB_OPT_TRACE, 1)
if a.no_halo:
    ctx.set_option(_lib.FRB_OPT_HALO, 0)
x = torch.randn(B, H, W, Cin, device="cuda").half()
w = (np.random.RandomState(0).standard_normal((Cout, k, k, Cin)) / np.sqrt(k * k * Cin)).astype(np.float32)
Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
res = torch.randn(B, Ho, Wo, Cout, device="cuda").half() if a.res else None
out, ms = ops.conv2d(ctx, x, w, None, (s, s), (p, p), 1, 0.0, res, 1 if a.res else 0, iters=1)
buf = np.zeros(1 + 3 * 4000, np.int64)
ctx.check(ctx.lib.frb_debug_trace(ctx.handle, buf.ctypes.data, 4000), "trace")
n = int(min(buf[0], 4000))
rec = buf[1:1 + 3 * n].reshape(n, 3)
t0 = rec[:, 2].min()
names = {10: "epi.ready", 11: "epi.acc_avail", 13: "epi.tmem_read", 14: "epi.slab_free", 12: "epi.store_issued",
         20: "mma.want_tmem", 21: "mma.tmem_free", 22: "mma.halo_landed", 23: "mma.issued_all", 30: "tma.halo_issued"}
print(f"{ms:.3f} ms per launch; {n} records")
# per-tile waits of the epilogue: for the accumulator (10 -> 11) and, summed over slabs, for a free staging slab (13 -> 14)
acc_wait, slab_wait, last13 = {}, {}, {}
ready = {}
for e, tl, c in rec:
    e, tl, c = int(e), int(tl), int(c)
    if e == 10: ready[tl] = c
    elif e == 11 and tl in ready: acc_wait[tl] = c - ready[tl]
    elif e == 13: last13[tl] = c
    elif e == 14 and tl in last13: slab_wait[tl] = slab_wait.get(tl, 0) + c - last13[tl]
mid = sorted(acc_wait)[5:-2]
if mid:
    print("epilogue, mean cycles per tile: waiting for the accumulator %.0f, waiting for a free staging slab %.0f" %
          (np.mean([acc_wait[t] for t in mid]), np.mean([slab_wait.get(t, 0) for t in mid])))
by_tile = {}
durs = [int(c) for e, t, c in rec if e == 24]
rec = rec[rec[:, 0] != 24]
t0 = rec[:, 2].min()
for e, t, c in rec:
    by_tile.setdefault(int(t), {})[int(e)] = int(c - t0)
if durs:
    print('cycles inside the tap loop per channel block (lane 0, clock64 delta):', durs[:3], '... mean', float(np.mean(durs[3:])))
for t in sorted(by_tile)[:3] + sorted(by_tile)[20:30]:
    d = by_tile[t]
    print(t, "  ".join(f"{names.get(e, e)}={d[e]}" for e in sorted(d, key=lambda e: d[e])))
tiles = sorted(by_tile)
if len(tiles) > 12:
    st = [by_tile[t].get(12) for t in tiles if 12 in by_tile[t]]
    print("mean cycles between consecutive tile completions:", float(np.diff(st[5:]).mean()))
