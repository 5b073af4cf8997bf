"""Stage-level throughput: InceptionResnetV1 faces/s vs batch, RetinaFace frames/s vs batch, and the whole extractor
(decode-free synthetic source -> detect -> host scene/SORT -> embed -> JSONL) in frames/s with a host-time breakdown."""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib, synth, weights as W  # noqa: E402


def ev_time(fn, iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--extract-frames", type=int, default=960)
    ap.add_argument("--extract-frames-1080p", type=int, default=768)
    a = ap.parse_args()
    ctx = _lib.Context(0)
    out = {}
    ops.load_model(ctx, W.MODEL_IDS["irv1"], W.write_blob(W.synth_weights("irv1", 0)))
    for F in (64, 256, 1024):
        faces = torch.randn(F, 160, 160, 8, device="cuda").half()
        l0 = ctx.launch_count()
        ms = ev_time(lambda: ops.irv1_forward(ctx, faces), 5)
        out[f"irv1_F{F}"] = {"ms": ms, "faces_per_s": F / ms * 1e3, "tflops": 2.835 * F / ms, "launches": (ctx.launch_count() - l0) // 6}
    for model, H, Wd in (("retina_mnet025", 360, 640), ("retina_r50", 1080, 1920)):
        mid = W.MODEL_IDS[model]
        ops.load_model(ctx, mid, W.write_blob(W.synth_weights(model, 0)))
        for B in (8, 64):
            fr = torch.from_numpy(synth.clip(1234, 0, B, H, Wd)).cuda()
            x16 = ops.preprocess_u8(ctx, fr, True, ops.input_layout(mid))
            ms = ev_time(lambda: ops.retina_forward(ctx, mid, x16, H, Wd), 3)
            out[f"{model}_B{B}"] = {"ms": ms, "frames_per_s": B / ms * 1e3}
    # whole extractor on the 360p clip (config 0 shape) and a short 1080p run
    from facerec_b200 import extract
    from facerec_b200.detector import B200RetinaFaceDetector
    from facerec_b200.embedder import B200FaceNet
    from facerec_b200.sort import NativeSort as Sort
    emb = B200FaceNet(ctx=ctx)
    for backbone, (W_, H_), n in (("mnet025", (640, 360), a.extract_frames), ("r50", (1920, 1080), a.extract_frames_1080p)):
        det = B200RetinaFaceDetector(20, backbone, ctx=ctx)
        src = extract.PreloadedSource(100001, n, W_, H_)
        with tempfile.TemporaryDirectory() as tmp:
            for rep in range(2):            # first pass warms plans / arena; the second is the measurement
                src.seek(0)
                ex = extract.Extractor(det, emb, Sort(5, 3, 0.5), tmp, 1, 0, save_images=False, batch=64)
                t0 = time.time()
                with contextlib.redirect_stdout(io.StringIO()):
                    counts = ex.process_video(src)
                dt = time.time() - t0
        out[f"extract_{backbone}_{W_}x{H_}"] = {"n_frames": n, "seconds": dt, "frames_per_s": n / dt, **counts,
                                                "main_thread_seconds": {k: round(v, 4) for k, v in ex.timing.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
