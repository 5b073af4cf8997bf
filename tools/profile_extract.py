"""cProfile of the whole extractor on a synthetic clip: where does the host time go?"""
import contextlib, cProfile, io, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import extract, _lib
from facerec_b200.detector import B200RetinaFaceDetector
from facerec_b200.embedder import B200FaceNet
from facerec_b200.sort import NativeSort as Sort
spec = sys.argv[1] if len(sys.argv) > 1 else "synth:100001:480:640x360"
backbone = sys.argv[2] if len(sys.argv) > 2 else "mnet025"
ctx = _lib.Context(0)
det = B200RetinaFaceDetector(20, backbone, ctx=ctx)
emb = B200FaceNet(ctx=ctx)
with tempfile.TemporaryDirectory() as tmp:
    ex = extract.Extractor(det, emb, Sort(5, 3, 0.5), tmp, 1, 0, save_images=False, batch=64)
    f = spec.split(":")
    w, h = (int(v) for v in f[3].split("x"))
    src = extract.PreloadedSource(int(f[1]), int(f[2]), w, h)   # frame synthesis (25 ms per 1080p frame) stays outside
    with contextlib.redirect_stdout(io.StringIO()):
        ex.process_video(src)                                    # warm-up (arena, first-call costs)
    src.seek(0)
    ex.tracker = Sort(5, 3, 0.5)
    pr = cProfile.Profile()
    t0 = time.time()
    pr.enable()
    with contextlib.redirect_stdout(io.StringIO()):
        counts = ex.process_video(src)
    pr.disable()
    dt = time.time() - t0
n = int(spec.split(":")[2])
print(f"{spec} {backbone}: {n} frames in {dt:.2f} s = {n/dt:.1f} frames/s (under cProfile), {counts}")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:60]))
