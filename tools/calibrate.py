"""Calibrate the class-head bias of the seeded random-init RetinaFace weights.

With plain random init the face score sits near 0.5 everywhere, so nothing passes the reference's 0.95
threshold (facerec/detector.py:68) and "identical kept-face sets" would be vacuous.  This tool runs the
fp32 oracle on a few frames of the synthetic clip and picks the bias shift that lets a target number
of priors per frame through the threshold.  Output: facerec_b200/calib.json (committed), read by
facerec_b200.weights.synth_weights.  Test infrastructure: it imports oracle/, the product does not.

Usage: python tools/calibrate.py [--model retina_r50 --height 1080 --width 1920 --target 60]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import weights as W, synth  # noqa: E402
from oracle import nets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="retina_r50")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--target", type=float, default=60.0, help="priors per frame above the 0.95 threshold")
    ap.add_argument("--frames", type=int, nargs="*", default=[0, 200, 400])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--clip-seed", type=int, default=1234)
    a = ap.parse_args()
    backbone = a.model.split("_", 1)[1]
    w = W.synth_weights(a.model, a.seed, calib={})
    P = nets.params_from_weights(w)
    ds = []
    for idx in a.frames:
        f = synth.frame(a.clip_seed, idx, a.height, a.width)
        x = torch.from_numpy(f[..., ::-1].astype(np.float32) - np.array([104, 117, 123], np.float32)).permute(2, 0, 1)[None]
        with torch.no_grad():
            _, conf, _ = nets.retina_forward(P, x.contiguous(), backbone)
        ds.append((conf[0, :, 1] - conf[0, :, 0]).numpy())
        print(f"frame {idx}: d std {ds[-1].std():.3f} max {ds[-1].max():.3f}", flush=True)
    d = np.concatenate(ds)
    thr = float(np.log(0.95 / 0.05))
    q = 1.0 - a.target / len(ds[0])
    bias = thr - float(np.quantile(d, q))
    per_frame = [int((x + bias > thr).sum()) for x in ds]
    print(f"bias {bias:.4f}: candidates per frame {per_frame}")
    path = W.CALIB_PATH
    calib = W.load_calib()
    calib[a.model] = {"cls_gain": 1.0, "cls_bias": round(bias, 4), "height": a.height, "width": a.width,
                      "target_per_frame": a.target, "weight_seed": a.seed, "clip_seed": a.clip_seed}
    with open(path, "w") as f:
        json.dump(calib, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
