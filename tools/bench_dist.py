"""Embedding-distance stage benchmarks (BASELINE.json configs[3] and configs[4]) on one GPU.

  pdist : N x 512 trajectory-mean embeddings -> condensed Euclidean distances (facerec/cluster.py:41)
  knn   : Q cluster/frame vectors vs M actor embeddings, k = 10, fused top-k (facerec/classify_knn.py:212,227)

Reports CUDA-event times, algorithmic TFLOP/s (2*N*N*D/2 for the triangle, 2*Q*M*D for kNN; the split-bf16 scheme
issues 3x that many tensor-core FLOPs) and the CPU library time on a bounded sub-sample.
Usage: python tools/bench_dist.py [--n 50000] [--q 50000] [--m 1000000] [--cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib  # noqa: E402


from facerec_b200.synth import embeddings  # noqa: E402


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--q", type=int, default=50000)
    ap.add_argument("--m", type=int, default=1000000)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--link-n", type=int, default=50000)
    a = ap.parse_args()
    ctx = _lib.Context(0)
    out = {}
    X = embeddings(a.n, 7, mean_of=2)
    cond = torch.empty(a.n * (a.n - 1) // 2, dtype=torch.float32, device="cuda")
    ms = timed(lambda: ops.pdist(ctx, X, 0, a.n, _lib.FRB_CONDENSED_F32, cond))
    fl = a.n * (a.n - 1) / 2 * 512 * 2
    out["pdist"] = {"N": a.n, "D": 512, "ms": ms, "algorithmic_tflops": fl / ms / 1e9, "tensor_tflops_issued": 3 * fl / ms / 1e9,
                    "out_GB": cond.numel() * 4 / 1e9, "write_GBps": cond.numel() * 4 / ms / 1e6}
    # checksum-of-checksums property at full size: row-sharded == whole
    whole = float(cond.double().sum())
    cond2 = torch.zeros_like(cond)
    for r0 in range(0, a.n, 12500):
        ops.pdist(ctx, X, r0, min(12500, a.n - r0), _lib.FRB_CONDENSED_F32, cond2)
    out["pdist"]["sharded_equals_whole"] = bool(torch.equal(cond, cond2))
    out["pdist"]["checksum"] = whole
    del cond, cond2
    DB = embeddings(a.m, 11, centres=10000)
    Q = embeddings(a.q, 13, centres=10000)
    ms = timed(lambda: ops.knn_topk(ctx, Q, DB, 10), iters=2)
    fl = 2.0 * a.q * a.m * 512
    out["knn"] = {"Q": a.q, "M": a.m, "k": 10, "ms": ms, "algorithmic_tflops": fl / ms / 1e9, "tensor_tflops_issued": 3 * fl / ms / 1e9}
    # complete linkage on the GPU at a size whose square fp32 matrix we also want to time end to end
    from facerec_b200 import distance
    n_l = min(a.n, a.link_n)
    Xl = X[:n_l].contiguous()
    t0 = time.time()
    D = ops.pdist(ctx, Xl, layout=_lib.FRB_SQUARE_SYM_F32)
    pairs = torch.empty((n_l - 1, 2), dtype=torch.int32, device="cuda")
    dist_ = torch.empty((n_l - 1,), dtype=torch.float32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.check(ctx.lib.frb_linkage_complete(ctx.handle, D.data_ptr(), n_l, pairs.data_ptr(), dist_.data_ptr(), _lib.current_stream()), "linkage")
    e1.record(); torch.cuda.synchronize()
    Z = distance._label_merges(pairs.cpu().numpy(), dist_.cpu().numpy().astype(np.float64), n_l)
    out["linkage"] = {"N": n_l, "kernel_ms": e0.elapsed_time(e1), "wall_s_incl_pdist_and_relabel": time.time() - t0,
                      "max_height": float(Z[-1, 2])}
    del D
    if a.cpu:
        from scipy.cluster.hierarchy import linkage as sp_linkage
        n_c = 8000
        cpd = ops.pdist(ctx, X[:n_c].contiguous(), layout=_lib.FRB_CONDENSED_F64).cpu().numpy()
        t0 = time.time(); sp_linkage(cpd, method="complete"); dt = time.time() - t0
        out["linkage"]["cpu_scipy_linkage_s_at_n8000"] = dt
        out["linkage"]["cpu_scipy_linkage_s_extrapolated"] = dt * (n_l / n_c) ** 2
        from scipy.spatial.distance import pdist
        from sklearn.neighbors import KNeighborsClassifier
        n_s = 4000
        t0 = time.time(); pdist(X[:n_s].cpu().numpy().astype(np.float64)); dt = time.time() - t0
        out["pdist"]["cpu_scipy_s_at_n4000"] = dt
        out["pdist"]["cpu_scipy_s_extrapolated"] = dt * (a.n / n_s) ** 2
        m_s, q_s = min(a.m, 100000), 500
        y = np.arange(m_s) % 1000
        knn = KNeighborsClassifier(n_neighbors=10).fit(DB[:m_s].cpu().numpy(), y)
        t0 = time.time(); knn.predict_proba(Q[:q_s].cpu().numpy()); dt = time.time() - t0
        out["knn"]["cpu_sklearn_s_sample"] = dt
        out["knn"]["cpu_sklearn_s_extrapolated"] = dt * (a.q / q_s) * (a.m / m_s)
        out["cpu_cores"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
