"""CUDA-event time of every stage of one detect+embed step (64 x 1080p, RetinaFace-R50 + InceptionResnetV1), back to back
like bench.py runs them: where the ~2 ms between the RetinaFace forward and the whole step go."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib, synth, weights as W
from facerec_b200.pipeline import FacePipeline

B, H, Wd = 64, 1080, 1920
pipe = FacePipeline("r50")
ctx, det = pipe.ctx, pipe.detector
mid = det.model_id
frames = torch.from_numpy(synth.clip(1234, 0, B, H, Wd)).cuda()
for _ in range(3):
    r = pipe.step_device(frames, 0)
torch.cuda.synchronize()
rects = torch.from_numpy(r.rects).cuda()
names, evs = [], []
def mark(n):
    e = torch.cuda.Event(enable_timing=True); e.record(); names.append(n); evs.append(e)
reps = 10
acc = {}
for _ in range(reps):
    names.clear(); evs.clear()
    mark("start")
    x16 = ops.preprocess_u8(ctx, frames, True, ops.input_layout(mid)); mark("preprocess (s2d)")
    loc, conf, landm = ops.retina_forward(ctx, mid, x16, H, Wd); mark("retina forward")
    cand, count, pidx = ops.prior_decode_filter(ctx, loc, conf, landm, H, Wd, det.threshold, det.max_candidates); mark("decode + filter")
    keep, nkeep = ops.nms(ctx, cand, count, pidx, det.nms_threshold, True); mark("nms")
    dets, nout = ops.gather_kept(ctx, cand, keep, nkeep, pipe.max_out); mark("gather kept")
    emb = pipe.embedder.embed_device(frames, rects); mark(f"crop + irv1 ({len(rects)} faces)")
    torch.cuda.synchronize()
    for i in range(1, len(evs)):
        acc[names[i]] = acc.get(names[i], 0.0) + evs[i - 1].elapsed_time(evs[i]) / reps
tot = sum(acc.values())
for k, v in acc.items():
    print(f"{k:32s} {v:8.3f} ms  {v / tot * 100:5.1f} %")
print(f"{'sum':32s} {tot:8.3f} ms")
