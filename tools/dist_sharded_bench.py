"""Row-sharded distance stage over the GPUs of one node (BASELINE.json configs[3], configs[4]); run under torchrun.

  torchrun --nproc-per-node N tools/dist_sharded_bench.py [--n 50000 --q 50000 --m 1000000]

pdist: each rank owns N/world embedding rows (as frame sharding leaves them), one NCCL all_gather makes every rank
self-sufficient, each rank fills its equal-area slice of the condensed vector.  kNN: queries sharded, database
replicated.  Times are CUDA events, max over ranks; correctness: rank slices are compared with a single-GPU run."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import ops, _lib, distance  # noqa: E402
from tools.bench_dist import embeddings  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--q", type=int, default=50000)
    ap.add_argument("--m", type=int, default=1000000)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    X = embeddings(a.n, 7, mean_of=2)                      # same seed on every rank -> same matrix
    lo, hi = rank * a.n // world, (rank + 1) * a.n // world
    mine = X[lo:hi].contiguous()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    distance.pdist_sharded(mine, ctx)                       # warm-up
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    Xall = distance.all_gather_rows(mine)
    ev[1].record()
    sl, off, n = distance.pdist_sharded(mine, ctx)
    ev[2].record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    r0, rows = distance.balanced_row_blocks(n, world)[rank]
    ref = distance._gpu_row_block(ctx, X, r0, rows)        # same rows from the un-gathered matrix
    ok = torch.tensor([int(torch.equal(ref, sl) and torch.equal(Xall, X))], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    DB = embeddings(a.m, 11, centres=10000)
    Q = embeddings(a.q, 13, centres=10000)
    qlo, qhi = rank * a.q // world, (rank + 1) * a.q // world
    distance.knn_sharded(Q[qlo:qhi].contiguous(), DB, 10, ctx)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx, d = distance.knn_sharded(Q[qlo:qhi].contiguous(), DB, 10, ctx)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    tk = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "pdist": {"N": a.n, "all_gather_ms": float(t[0]), "pdist_ms_max_rank": float(t[1]),
                                                    "slices_equal_single_gpu": bool(ok.item())},
                          "knn": {"Q": a.q, "M": a.m, "ms_max_rank": float(tk[0]),
                                  "algorithmic_tflops_total": 2.0 * a.q * a.m * 512 / float(tk[0]) / 1e9}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
