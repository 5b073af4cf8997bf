"""GPU time of the extractor's per-batch stages at configs[1] (64 x 1080p): scene statistics, detection (launch + collect),
crop + embed of 188 faces, and the host -> device copy of the batch.  Measured: 1.1 / 35.5 / 2.9 / 7.2 ms."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from facerec_b200 import synth, _lib
from facerec_b200.scene import B200SceneChangeDetector
from facerec_b200.pipeline import FacePipeline
ctx = _lib.Context(0)
fr = torch.from_numpy(synth.clip(1234, 0, 64, 1080, 1920)).cuda()
sc = B200SceneChangeDetector(False, True, "m", ctx=ctx)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("scene stats launch+collect ms", t(lambda: sc.update_collect(sc.stats_launch(fr))))
pipe = FacePipeline("r50", min_face_size=20, save_every=5, ctx=ctx)
print("detect launch+collect ms", t(lambda: pipe.detector.detect_batch_collect(pipe.detector.detect_batch_launch(fr))))
rects = np.array([[i % 64, 100 + (i % 7) * 50, 100, 160, 160] for i in range(188)], np.int32)
print("embed 188 faces ms", t(lambda: pipe.embedder.embed_collect(pipe.embedder.embed_boxes_launch(fr, rects))))
pinned = torch.empty(fr.shape, dtype=torch.uint8).pin_memory()
print("h2d 398MB ms", t(lambda: fr.copy_(pinned, non_blocking=True)))
