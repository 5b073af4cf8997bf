"""Where the distance GEMM's time goes (debug build: FRB_DEBUG_BUILD=1 python -m facerec_b200.build --force):
the same launches with the epilogue reduced to the barrier hand-over (FRB_DIST_DBG=1) and with a shallower stage ring
(FRB_DIST_STAGES)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib, synth  # noqa: E402


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ctx = _lib.Context(0)
    PN, KQ, KM = 30000, 16384, 524288
    X = synth.embeddings(PN, 7, mean_of=2)
    cond = torch.empty(PN * (PN - 1) // 2, dtype=torch.float32, device="cuda")
    DB = synth.embeddings(KM, 11, centres=10000)
    Q = synth.embeddings(KQ, 13, centres=10000)
    out = {}
    for dbg in (0, 1):
        for stages in (8, 3):
            os.environ["FRB_DIST_DBG"] = str(dbg)
            os.environ["FRB_DIST_STAGES"] = str(stages)
            pd = timed(lambda: ops.pdist(ctx, X, 0, PN, _lib.FRB_CONDENSED_F32, cond))
            kn = timed(lambda: ops.knn_topk(ctx, Q, DB, 10))
            out[f"dbg{dbg}_stages{stages}"] = {"pdist_ms": pd, "pdist_issued_tflops": 3 * PN * (PN - 1) / 2 * 1024 / pd / 1e9,
                                              "knn_ms": kn, "knn_tflops_one_pass": 2.0 * KQ * KM * 512 / kn / 1e9,
                                              "knn_redo": ops.knn_redo_count(ctx)}
            print(f"dbg{dbg}_stages{stages}", out[f"dbg{dbg}_stages{stages}"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
