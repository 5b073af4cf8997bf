"""Per-op times of one network forward under SUSTAINED load (FRB_OPT_LAYER_TIMES): the forward is run back to back
for a second first so the board sits at its power-capped clocks, then the last forward's event brackets are read.
Usage: python tools/layer_times.py [retina_r50|retina_mnet025|irv1] [batch] > gpurun_out/layer_times.txt"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib, synth, weights as W  # noqa: E402

KINDS = ["conv", "dwconv", "maxpool", "avgpool", "l2norm"]


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "retina_r50"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    H, Wd = int(os.environ.get("FRB_H", "1080")), int(os.environ.get("FRB_W", "1920"))
    ctx = _lib.Context(0)
    if os.environ.get("FRB_DUAL"):
        ctx.set_option(_lib.FRB_OPT_DUAL_ISSUE, int(os.environ["FRB_DUAL"]))
    if os.environ.get("FRB_KBPS2"):
        ctx.set_option(_lib.FRB_OPT_KBPS2, int(os.environ["FRB_KBPS2"]))
    if os.environ.get("FRB_SLICE"):
        ctx.set_option(_lib.FRB_OPT_WEIGHT_SLICE, int(os.environ["FRB_SLICE"]))
    if os.environ.get("FRB_BN_CAP"):
        ctx.set_option(_lib.FRB_OPT_BN_CAP, int(os.environ["FRB_BN_CAP"]))
    if os.environ.get("FRB_MPAIR"):
        ctx.set_option(_lib.FRB_OPT_MPAIR, int(os.environ["FRB_MPAIR"]))
    if os.environ.get("FRB_PDL"):
        ctx.set_option(_lib.FRB_OPT_PDL, int(os.environ["FRB_PDL"]))
    if os.environ.get("FRB_CTA2"):
        ctx.set_option(_lib.FRB_OPT_CTA2, int(os.environ["FRB_CTA2"]))
    if os.environ.get("FRB_FUSE"):
        ctx.set_option(_lib.FRB_OPT_FUSE_TAIL, int(os.environ["FRB_FUSE"]))
    if os.environ.get("FRB_TAIL2"):
        ctx.set_option(_lib.FRB_OPT_FUSE_TAIL2, int(os.environ["FRB_TAIL2"]))
    if os.environ.get("FRB_POOL"):
        ctx.set_option(_lib.FRB_OPT_FUSE_POOL, int(os.environ["FRB_POOL"]))
    if os.environ.get("FRB_DWPW"):
        ctx.set_option(_lib.FRB_OPT_FUSE_DWPW, int(os.environ["FRB_DWPW"]))
    if os.environ.get("FRB_HALO"):
        ctx.set_option(_lib.FRB_OPT_HALO, int(os.environ["FRB_HALO"]))
    mid = W.MODEL_IDS[model]
    ops.load_model(ctx, mid, W.write_blob(W.synth_weights(model, 0)))
    if model == "irv1":
        x16 = (torch.randn(B, 160, 160, 8, device="cuda") * 0.5).half()
        x16[..., 3:] = 0
        run = lambda: ops.irv1_forward(ctx, x16)
    else:
        fr = torch.from_numpy(synth.clip(1234, 0, B, H, Wd)).cuda()
        x16 = ops.preprocess_u8(ctx, fr, True, ops.input_layout(mid))
        run = lambda: ops.retina_forward(ctx, mid, x16, H, Wd)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / reps
    ctx.set_option(_lib.FRB_OPT_LAYER_TIMES, 1)
    for _ in range(3):
        run()
    rows = np.zeros((512, 6), np.float64)
    n = ctx.lib.frb_debug_layer_times(ctx.handle, rows.ctypes.data_as(C.c_void_p), 512)
    assert n > 0, n
    rows = rows[:n]
    tot = rows[:, 5].sum()
    print(f"# {model} batch {B}: {plain:.3f} ms per forward over {reps} back-to-back forwards; {n} ops, bracketed sum {tot:.3f} ms")
    print("# idx kind        M      N      K        ms   TFLOP/s   share   io GB/s (in+out 16-bit, K/taps unknown -> M*(K'+N)*2)")
    order = np.argsort(-rows[:, 5])
    rank = {int(i): r for r, i in enumerate(order)}
    for i, (kind, M, N, K, fl, ms) in enumerate(rows):
        tf = fl / ms / 1e9 if ms > 0 else 0.0
        print(f"{i:4d} {KINDS[int(kind)]:8s} {int(M):9d} {int(N):5d} {int(K):6d} {ms:9.3f} {tf:9.1f} {ms / tot * 100:6.1f}%  rank {rank[i]:3d}")
    conv = rows[rows[:, 0] == 0]
    print(f"# conv total {conv[:, 5].sum():.3f} ms for {conv[:, 4].sum() / 1e12:.2f} TFLOP -> {conv[:, 4].sum() / conv[:, 5].sum() / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
