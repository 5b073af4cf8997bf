"""One warm-up + one RetinaFace-R50 forward over a 64-frame 1080p batch (the roofline unit of bench.py), for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib, synth, weights as W
B, H, Wd = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 1080, 1920
ctx = _lib.Context(0)
mid = W.MODEL_IDS["retina_r50"]
ops.load_model(ctx, mid, W.write_blob(W.synth_weights("retina_r50", 0)))
fr = torch.from_numpy(synth.clip(1234, 0, B, H, Wd)).cuda()
x16 = ops.preprocess_u8(ctx, fr, True, ops.input_layout(mid))
if os.environ.get("FRB_FUSE"):
    ctx.set_option(_lib.FRB_OPT_FUSE_TAIL, int(os.environ["FRB_FUSE"]))
n0 = 0
for _ in range(2):
    n0 = ctx.launch_count()
    ops.retina_forward(ctx, mid, x16, H, Wd)
torch.cuda.synchronize()
print("launches per forward:", ctx.launch_count() - n0)
