"""One warm-up + one launch of every memory-bound glue kernel (K1, K6, K7, K7b, K8) at the configs[1] batch
(64 x 1080p, RetinaFace-R50 outputs) and of the distance-stage kernels (pdist, kNN) at a reduced size, for the ncu
metrics pass (gpu__time_duration, dram__bytes_read/write, lts__t_bytes, tensor-pipe active):

  python tools/kernels_once.py > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:'preprocess|prior_decode|nms_kernel|gather_kept|crop_resize|split_rows|dist|knn|maxpool' \
      --csv --log-file gpurun_out/glue_dist_launches.csv python tools/kernels_once.py

Prints the algorithmic bytes / FLOPs of every launch so that the csv can be turned into GB/s and TFLOP/s
(tools/summarize_launches.py).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib, synth, weights as W  # noqa: E402
from facerec_b200.pipeline import FacePipeline  # noqa: E402
from tools.bench_dist import embeddings  # noqa: E402

B, H, Wd = int(os.environ.get("FRB_B", "64")), 1080, 1920
PN = int(os.environ.get("FRB_PDIST_N", "20000"))
KQ, KM = int(os.environ.get("FRB_KNN_Q", "8192")), int(os.environ.get("FRB_KNN_M", "262144"))


def main():
    ctx = _lib.Context(0)
    info = {}
    if os.environ.get("FRB_SKIP_GLUE", "0") == "0":
        info = glue(ctx)
    if os.environ.get("FRB_SKIP_DIST", "0") == "0":
        info.update(dist(ctx))
    print(json.dumps(info))


def glue(ctx):
    pipe = FacePipeline("r50", min_face_size=20, save_every=1, ctx=ctx)
    mid = pipe.detector.model_id
    fr = torch.from_numpy(synth.clip(1234, 0, B, H, Wd)).cuda()
    x16 = ops.preprocess_u8(ctx, fr, True, ops.input_layout(mid))
    loc, conf, landm = ops.retina_forward(ctx, mid, x16, H, Wd)
    torch.cuda.synchronize()
    rects = None
    for it in range(2):      # the second pass is the one to read (index 1 of each kernel in the csv)
        ops.preprocess_u8(ctx, fr, True, _lib.FRB_LAYOUT_S2D16)
        ops.preprocess_u8(ctx, fr, True, _lib.FRB_LAYOUT_NHWC8)
        cand, count, pidx = ops.prior_decode_filter(ctx, loc, conf, landm, H, Wd, 0.95, pipe.detector.max_candidates)
        keep, nkeep = ops.nms(ctx, cand, count, pidx, 0.4, True)
        dets, nout = ops.gather_kept(ctx, cand, keep, nkeep, pipe.max_out)
        if rects is None:
            d = pipe._dets_to_host(dets, nout, count)
            _, _, rects = pipe.select_faces(d, 0, Wd, H)
            rects_d = torch.from_numpy(rects).cuda()
        ops.crop_resize_std(ctx, fr, rects_d)
        torch.cuda.synchronize()
    P = ops.num_priors(H, Wd)
    crop_px = int((rects[:, 3].astype(np.int64) * rects[:, 4]).sum())
    info = {
        "batch": B, "H": H, "W": Wd, "faces": int(len(rects)), "candidates": int(count.sum().item()),
        "algorithmic_bytes": {
            "preprocess_s2d_kernel": B * H * Wd * 3 + B * ((H + 1) // 2) * ((Wd + 1) // 2 + 3) * 32,
            "preprocess_kernel": B * H * Wd * 3 + B * H * Wd * 16,
            "prior_decode_filter_kernel": B * P * 8 + int(count.sum().item()) * (16 + 40 + 64),
            "crop_resize_kernel": crop_px * 3 + len(rects) * 160 * 160 * 16,
        },
    }
    return info


def dist(ctx):
    # ---- distance stage, reduced sizes (same kernels, same tile shapes as configs[3] / configs[4])
    info = {}
    X = embeddings(PN, 7, mean_of=2)
    cond = torch.empty(PN * (PN - 1) // 2, dtype=torch.float32, device="cuda")
    DB = embeddings(KM, 11, centres=10000)
    Q = embeddings(KQ, 13, centres=10000)
    for it in range(2):
        ops.pdist(ctx, X, 0, PN, _lib.FRB_CONDENSED_F32, cond)
        ops.knn_topk(ctx, Q, DB, 10)
        torch.cuda.synchronize()
    info["pdist"] = {"N": PN, "D": 512, "algorithmic_flop": PN * (PN - 1) / 2 * 512 * 2, "out_bytes": PN * (PN - 1) // 2 * 4}
    info["knn"] = {"Q": KQ, "M": KM, "D": 512, "algorithmic_flop": 2.0 * KQ * KM * 512,
                   "min_bytes": (KQ + KM) * 512 * 2 + KQ * 10 * 8, "queries_through_exact_pass": ops.knn_redo_count(ctx)}
    return info


if __name__ == "__main__":
    main()
