"""How busy is the GPU while the extractor runs on preloaded 1080p frames?  torch.profiler (CUPTI) over one process_video
call: total kernel + memcpy time on the device against the wall time of the loop, and the largest gaps.
Measured (1 280 frames, 20 batches of 64): wall 1 033 ms, device busy 942 ms (91 %): 40 ms of kernels + a 7.2 ms upload per batch.
Uploading on a copy stream into a device ring of the extractor's own was tried: wall 1 003-1 027 ms, not kept."""
import contextlib, io, os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import extract, _lib
from facerec_b200.detector import B200RetinaFaceDetector
from facerec_b200.embedder import B200FaceNet
from facerec_b200.sort import NativeSort as Sort
from torch.profiler import profile, ProfilerActivity

ctx = _lib.Context(0)
det = B200RetinaFaceDetector(20, "r50", ctx=ctx)
emb = B200FaceNet(ctx=ctx)
with tempfile.TemporaryDirectory() as tmp:
    ex = extract.Extractor(det, emb, Sort(5, 3, 0.5), tmp, 1, 0, save_images=False, batch=64)
    src = extract.PreloadedSource(100001, int(sys.argv[1]) if len(sys.argv) > 1 else 512, 1920, 1080)
    with contextlib.redirect_stdout(io.StringIO()):
        ex.process_video(src)
    src.seek(0)
    ex.tracker = Sort(5, 3, 0.5)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            ex.process_video(src)
        torch.cuda.synchronize()
        wall = time.time() - t0
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    busy = sum(e.time_range.end - e.time_range.start for e in evs) / 1e3
    kern = {}
    for e in evs:
        kern[e.name[:40]] = kern.get(e.name[:40], 0.0) + (e.time_range.end - e.time_range.start) / 1e3
    gaps = []
    for a, b in zip(evs, evs[1:]):
        g = (b.time_range.start - a.time_range.end) / 1e3
        if g > 0.5:
            gaps.append((round(g, 2), a.name[:30], b.name[:30]))
    print(f"wall {wall * 1e3:.1f} ms, device busy {busy:.1f} ms ({busy / (wall * 1e3):.0%}), timing {ex.timing}")
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1])[:12]:
        print(f"  {v:9.2f} ms  {k}")
    print("gaps > 0.5 ms:", len(gaps), "sum", round(sum(g[0] for g in gaps), 1), "ms; largest:", sorted(gaps, reverse=True)[:12])
