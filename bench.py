#!/usr/bin/env python3
"""Headline benchmark: video frames/s of detect+embed (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[1]): 1920x1080 synthetic clip, RetinaFace-ResNet50 + InceptionResnetV1,
64-frame batches, seeded random-init weights.  A step = one pass of the hot path over one 64-frame
batch: K1 preprocess -> RetinaFace forward (tcgen05 implicit-GEMM convs) -> K6 decode/filter -> K7 NMS ->
host box arithmetic -> K8 crop/resize -> InceptionResnetV1 + L2 norm, faces embedded on every
`save_every`-th frame (the reference's gate, facerec/extract.py:336).  Frames shard over ranks exactly
like the reference's --shard-i/--n-shards split (no data-path collective): weak scaling.

  python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA path
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # the CPU oracle on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU
  python bench.py --workload pdist|knn [--impl reference]        # configs[3] / configs[4]: the embedding-distance stage

The default line also carries a `distance` object: configs[3] (50k x 512 pdist, rows sharded over the ranks, one NCCL
all_gather of the embeddings) and configs[4] (50k queries x 1M actor embeddings, fused top-10, queries sharded), each with its own
device-timed rate, end-to-end rate and roofline, so that the driver's 1/2/4/8-GPU runs exercise the path's only collective.

`value`  : frames/s with the batch already resident in HBM when the timed region starts, K steps through
           FacePipeline.run_pipelined (two batches in flight, the schedule of facerec_b200.extract: detection of step k+1 is
           enqueued before the host collects step k; every step's results are complete inside the timed region).
           `ms_per_step_sequential` is the same K steps one at a time (FacePipeline.step_device).
`e2e`    : the same from pinned HOST frames: every step copies its own 64 frames host -> device (a copy stream runs ahead,
           as the extractor's prefetch does) and reads its detections + embeddings back (D2H) inside the timed region.
`roofline`: the RetinaFace-R50 forward (64 launches: conv_tma_kernel, 3 + 3 fused conv_tail / conv_tail2 kernels, one zero fill), timed with
           CUDA events on the launching stream inside the timed region; algorithmic work
           449.71 GFLOP/frame x 64 frames per launch of the forward.  `traffic` is not measured by this
           script: it is the dram__bytes sum of the committed ncu launch list of the same forward (profiles/r02_fwd_traffic.json).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R50_GFLOP_PER_FRAME_1080P = 449.71    # SURVEY App. C (82 convs, 1920x1080)
IRV1_GFLOP_PER_FACE = 2.835
METRIC = "video_frames_per_sec_detect_embed"


def conv_gflop_retina(backbone: str, H: int, W: int) -> float:
    """Analytic conv FLOPs per frame of the RetinaFace graph (2 FLOP per MAC), any input size."""
    from facerec_b200 import weights as Wt
    o = lambda n, k, s, p: (n + 2 * p - k) // s + 1
    macs = 0
    if backbone == "r50":
        h, w = o(H, 7, 2, 3), o(W, 7, 2, 3)
        macs += h * w * 64 * 49 * 3
        h, w = o(h, 3, 2, 1), o(w, 3, 2, 1)
        cin = 64
        feats = []
        for l, (nb, planes) in enumerate(zip(Wt.R50_BLOCKS, Wt.R50_PLANES)):
            for i in range(nb):
                s = 2 if (i == 0 and l > 0) else 1
                macs += h * w * planes * cin
                ho, wo = o(h, 3, s, 1), o(w, 3, s, 1)
                macs += ho * wo * planes * planes * 9 + ho * wo * planes * 4 * planes
                if i == 0:
                    macs += ho * wo * planes * 4 * cin
                cin, h, w = planes * 4, ho, wo
            if l >= 1:
                feats.append((h, w, cin))
        C = 256
    else:
        h, w = o(H, 3, 2, 1), o(W, 3, 2, 1)
        macs += h * w * 8 * 27
        cin = 8
        feats = []
        for i, (s, cout) in enumerate(zip(Wt.MNET_STRIDES, Wt.MNET_COUT), start=1):
            h, w = o(h, 3, s, 1), o(w, 3, s, 1)
            macs += h * w * cin * 9 + h * w * cin * cout
            cin = cout
            if i in (5, 11, 13):
                feats.append((h, w, cin))
        C = 64
    for (h, w, c) in feats:
        macs += h * w * c * C
    for (h, w, _) in feats[:2]:
        macs += h * w * C * C * 9
    q = C // 4
    for (h, w, _) in feats:
        macs += h * w * 9 * (C * 2 * q + C * q + 3 * q * q) + h * w * C * 32
    return 2.0 * macs / 1e9


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained figure)"


def cited_traffic(backbone, B, H, W):
    """DRAM bytes of one forward from the committed ncu launch list (tools/summarize_launches.py): a CITED constant,
    bench.py does not run under a profiler."""
    p = os.path.join(ROOT, "profiles", "r02_fwd_traffic.json")
    if (backbone, B, H, W) != ("r50", 64, 1080, 1920) or not os.path.exists(p):
        return {"traffic": None}
    d = json.load(open(p))
    return {"traffic": d["dram_bytes"],
            "traffic_unit": f"bytes per forward, CITED from the ncu launch list {d['source']} ({d['dram_read_bytes'] / 1e9:.1f} GB read + "
                            f"{d['dram_write_bytes'] / 1e9:.1f} GB written over {d['launches']} launches; layer-by-layer minimum (every layer output written "
                            f"and read once, SURVEY App. C) 2.60 GB/frame x 64 = 166 GB: the fused tails / stem pool remove 45 GB of it; "
                            f"L2 -> SM {d['l2_bytes'] / 1e9:.0f} GB; time-weighted tensor-pipe active {d['tensor_pipe_active_pct_time_weighted']:.1f} %)"}


def make_batch(seed, first, B, H, W):
    from facerec_b200 import synth
    return synth.clip(seed, first, B, H, W)


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_PARAMS = {}


def cpu_reference_fps(frames: np.ndarray, first_index: int, backbone: str, save_every: int, threads: int):
    """The fp32 oracle driven like the reference loop (one frame per detect call, one face per embed call):
    returns (frames/s, faces embedded)."""
    import torch
    from facerec_b200 import weights as Wt
    from facerec_b200.pipeline import boxes_to_rects
    from oracle import nets, postproc, crop
    torch.set_num_threads(threads)
    H, W = frames.shape[1:3]
    if backbone not in _CPU_PARAMS:      # model construction is outside every timed region (as on the GPU arm)
        _CPU_PARAMS[backbone] = (nets.params_from_weights(Wt.synth_weights("retina_" + backbone, 0)),
                                 nets.params_from_weights(Wt.synth_weights("irv1", 0)))
    Pd, Pe = _CPU_PARAMS[backbone]
    t0 = time.time()
    faces = 0
    with torch.no_grad():
        for b in range(len(frames)):
            f = frames[b]
            x = torch.from_numpy(f[..., ::-1].astype(np.float32) - np.array([104, 117, 123], np.float32))
            loc, conf, landm = nets.retina_forward(Pd, x.permute(2, 0, 1)[None].contiguous(), backbone)
            dets, _ = postproc.postprocess(loc[0].numpy(), conf[0].numpy(), landm[0].numpy(), H, W)
            if (first_index + b) % save_every != 0:
                continue
            dets = dets[np.minimum(dets[:, 2] - dets[:, 0], dets[:, 3] - dets[:, 1]) >= 20] if len(dets) else dets
            rects = boxes_to_rects(dets[:, :4], np.zeros(len(dets), np.int64), W, H)
            for (_, x0, y0, w, h) in rects:
                if w <= 0 or h <= 0:
                    continue
                face = crop.standardize(crop.resize_linear_u8(f[y0:y0 + h, x0:x0 + w]))
                nets.irv1_forward(Pe, torch.from_numpy(face).permute(2, 0, 1)[None].contiguous())
                faces += 1
    dt = time.time() - t0
    return len(frames) / dt, faces, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(1, a.ref_frames)
    # sample index 0 lands on the save_every gate so every step also embeds that frame's faces
    frames = make_batch(a.clip_seed, 0, sample, a.height, a.width)
    for _ in range(a.warmup):
        cpu_reference_fps(frames[:1], 0, a.backbone, a.save_every, threads)
    faces = 0
    dt = 0.0
    for _ in range(a.steps):
        _, f, d = cpu_reference_fps(frames, 0, a.backbone, a.save_every, threads)
        faces += f
        dt += d
    fps = a.steps * sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(a, sample_frames=sample),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} frame(s) of the {a.width}x{a.height} clip per step, fp32 PyTorch oracle, "
                                   f"batch 1 per frame/face like facerec/extract.py:303,166; {faces} faces embedded"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(a, sample_frames=None):
    c = {"workload": f"configs[1]: {a.width}x{a.height} synthetic clip, RetinaFace-{a.backbone} + InceptionResnetV1, "
                     f"{a.batch}-frame batches, save_every={a.save_every}, random-init seeded weights",
         "global_batch_frames": a.batch * a.gpus, "frames_per_gpu_per_step": a.batch, "parallelism": f"frame-shard x{a.gpus}",
         "l2_policy": "inputs larger than L2 (one 64-frame 1080p batch = 398 MB u8, 1.06 GB after K1)"}
    if sample_frames is not None:
        c["sample_frames_per_step"] = sample_frames
    return c



# ------------------------------------------------------------------------------------------ distance stage
PDIST_N, KNN_Q, KNN_M, KNN_K, EMB_D = 50000, 50000, 1000000, 10, 512


def _timed_region(fn, steps, warmup, world, dist, device):
    """W warm-up calls, then K timed calls between barrier + synchronize; CUDA events on the launching stream; max over ranks."""
    import torch
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / steps


def load_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_burst_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)"
    return 1590.0, "fallback (B200_PROFILING.md burst figure)"


def bench_pdist(a, ctx, rank, world, steps, warmup):
    """configs[3]: N x 512 trajectory means -> condensed Euclidean distances (facerec/cluster.py:41), rows sharded over
    the ranks by equal triangle area, one all_gather of the embeddings (the path's only exchange)."""
    import torch
    import torch.distributed as dist
    from facerec_b200 import synth, distance, _lib
    dev = torch.device("cuda", ctx.device)
    N = a.pdist_n
    X = synth.embeddings(N, 7, mean_of=2, device=dev)                     # same seed on every rank -> same matrix
    lo, hi = rank * N // world, (rank + 1) * N // world
    mine = X[lo:hi].contiguous()                                           # the rows this rank's frame shard produced
    r0, rows = distance.balanced_row_blocks(N, world)[rank]
    beg, end = distance.condensed_offset(N, r0), distance.condensed_offset(N, r0 + rows)
    out = torch.empty((max(end - beg, 1),), dtype=torch.float32, device=dev)
    pairs = N * (N - 1) // 2

    def step():
        Xall = distance.all_gather_rows(mine) if world > 1 else mine
        ctx.check(ctx.lib.frb_pdist(ctx.handle, Xall.data_ptr(), N, EMB_D, r0, rows, _lib.FRB_CONDENSED_F32,
                                    out.data_ptr() - beg * 4, _lib.current_stream()), "frb_pdist")

    l0 = ctx.launch_count()
    ms = _timed_region(step, steps, warmup, world, dist, dev)
    launches = (ctx.launch_count() - l0) // (steps + warmup)
    # the kernel alone (rank 0's block), for the roofline
    Xall = distance.all_gather_rows(mine) if world > 1 else mine
    kms = _timed_region(lambda: ctx.check(ctx.lib.frb_pdist(ctx.handle, Xall.data_ptr(), N, EMB_D, r0, rows, _lib.FRB_CONDENSED_F32,
                                                            out.data_ptr() - beg * 4, _lib.current_stream()), "frb_pdist"),
                        steps, 1, world, dist, dev)
    # end to end through the host-facing call: pinned host embeddings in, this rank's slice of the condensed vector out
    host_in = mine.cpu().pin_memory()
    host_out = torch.empty((end - beg,), dtype=torch.float32, pin_memory=True)

    def step_e2e():
        m = host_in.to(dev, non_blocking=True)
        Xa = distance.all_gather_rows(m) if world > 1 else m
        ctx.check(ctx.lib.frb_pdist(ctx.handle, Xa.data_ptr(), N, EMB_D, r0, rows, _lib.FRB_CONDENSED_F32,
                                    out.data_ptr() - beg * 4, _lib.current_stream()), "frb_pdist")
        host_out.copy_(out[:end - beg], non_blocking=True)

    ems = _timed_region(step_e2e, max(1, min(steps, 3)), 1, world, dist, dev)
    flop = pairs * 2.0 * EMB_D
    peak, peak_src = load_burst_peak()
    hbm, _ = load_hbm_peak()
    my_flop = (end - beg) * 2.0 * EMB_D
    rec = {"metric": "pdist_pairs_per_sec", "value": pairs / (ms / 1e3), "unit": "pairs/s", "ms_per_step": ms,
           "config": {"workload": f"configs[3]: {N} x {EMB_D} trajectory-mean embeddings -> condensed fp32 Euclidean distances, "
                                  f"rows sharded over {world} GPU(s) by equal triangle area" +
                                  (", one NCCL all_gather of the embeddings per step" if world > 1 else ""),
                      "N": N, "D": EMB_D, "parallelism": f"row-shard x{world}",
                      "l2_policy": f"output slice {4e-9 * (end - beg):.2f} GB per rank, larger than L2"},
           "e2e": {"value": pairs / (ems / 1e3), "unit": "pairs/s", "ms_per_step": ems, "h2d_bytes_per_step": int(host_in.numel() * 4),
                   "d2h_bytes_per_step": int((end - beg) * 4),
                   "api": "pinned host embeddings -> H2D -> (all_gather) -> frb_pdist -> D2H of this rank's condensed slice"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "tensor", "achieved": my_flop / kms / 1e9, "peak": peak, "unit": "TFLOP/s",
                        "frac": my_flop / kms / 1e9 / peak, "traffic": None,
                        "kernel": "dist_gemm_kernel (cta_group::2, 256x256 tiles, split-bf16: 3 tensor passes per algorithmic FLOP)",
                        "algorithmic_flop_per_launch": my_flop, "launch_ms": kms, "peak_source": peak_src,
                        "issued_tflops": 3 * my_flop / kms / 1e9,
                        "write_gbs": (end - beg) * 4 / kms / 1e6, "write_frac_of_hbm_copy_peak": (end - beg) * 4 / kms / 1e6 / hbm}}
    del out, X
    return rec


def bench_knn(a, ctx, rank, world, steps, warmup):
    """configs[4]: Q cluster means vs M actor-image embeddings, fused cosine/Euclidean top-10 + vote
    (facerec/classify_knn.py:212,227); queries sharded over the ranks, database replicated (loaded once, like fit())."""
    import torch
    import torch.distributed as dist
    from facerec_b200 import synth, ops
    dev = torch.device("cuda", ctx.device)
    Qn, M = a.knn_q, a.knn_m
    DB = synth.embeddings(M, 11, centres=10000, device=dev)
    Q = synth.embeddings(Qn, 13, centres=10000, device=dev)
    labels = (torch.arange(M, device=dev, dtype=torch.int32) % 1000).contiguous()   # 1000 actors
    lo, hi = rank * Qn // world, (rank + 1) * Qn // world
    mine = Q[lo:hi].contiguous()
    ops.knn_fit(ctx, DB)                    # KNeighborsClassifier.fit (facerec/classify_knn.py:212): outside the timed predict calls

    def step():
        idx, _ = ops.knn_topk(ctx, mine, DB, KNN_K)
        return ops.knn_vote(ctx, idx, labels, 1000)

    l0 = ctx.launch_count()
    ms = _timed_region(step, steps, warmup, world, dist, dev)
    launches = (ctx.launch_count() - l0) // (steps + warmup)
    redo = ops.knn_redo_count(ctx)
    host_q = mine.cpu().pin_memory()
    host_idx = torch.empty((hi - lo, KNN_K), dtype=torch.int32, pin_memory=True)
    host_d = torch.empty((hi - lo, KNN_K), dtype=torch.float32, pin_memory=True)

    def step_e2e():
        q = host_q.to(dev, non_blocking=True)
        idx, d = ops.knn_topk(ctx, q, DB, KNN_K)
        host_idx.copy_(idx, non_blocking=True)
        host_d.copy_(d, non_blocking=True)

    ems = _timed_region(step_e2e, max(1, min(steps, 3)), 1, world, dist, dev)
    flop = 2.0 * (hi - lo) * M * EMB_D
    peak, peak_src = load_burst_peak()
    rec = {"metric": "knn_queries_per_sec", "value": Qn / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms,
           "config": {"workload": f"configs[4]: {Qn} cluster-mean queries vs {M} actor-image embeddings x {EMB_D}, k = {KNN_K}, "
                                  f"fused top-k + uniform vote, queries sharded over {world} GPU(s), database replicated",
                      "Q": Qn, "M": M, "D": EMB_D, "k": KNN_K, "parallelism": f"query-shard x{world}",
                      "l2_policy": f"database {M * EMB_D * 2 / 1e9:.2f} GB (bf16 hi plane) streamed per query block, larger than L2"},
           "e2e": {"value": Qn / (ems / 1e3), "unit": "queries/s", "ms_per_step": ems, "h2d_bytes_per_step": int(host_q.numel() * 4),
                   "d2h_bytes_per_step": int((hi - lo) * KNN_K * 8),
                   "api": "pinned host queries -> H2D -> frb_knn_topk (database resident, as after fit()) -> D2H of indices + distances"},
           "gpu_launches": int(launches), "queries_through_exact_pass": int(redo),
           "roofline": {"bound": "tensor", "achieved": flop / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": flop / ms / 1e9 / peak,
                        "traffic": None,
                        "kernel": "dist_gemm_kernel (cta_group::2, 256x256 tiles, resident A, one bf16 pass) + knn_merge_kernel (exact re-rank); "
                                  "timed as the whole frb_knn_topk + frb_knn_vote step",
                        "algorithmic_flop_per_launch": flop, "launch_ms": ms, "peak_source": peak_src}}
    del DB, Q
    return rec


def cpu_pdist_rate(n_s, threads):
    """scipy pdist (what linkage(vectors, 'complete') runs first, facerec/cluster.py:41) on a sample: pairs/s."""
    from scipy.spatial.distance import pdist
    rng = np.random.RandomState(7)
    x = rng.standard_normal((n_s, EMB_D)).astype(np.float32)
    t0 = time.time()
    pdist(x)
    dt = time.time() - t0
    return n_s * (n_s - 1) / 2 / dt, dt


def cpu_knn_rate(q_s, m_s, threads):
    """sklearn KNeighborsClassifier(10).fit().predict_proba() (facerec/classify_knn.py:212,227) on a sample: queries/s
    against the FULL-size database equivalent (time scales linearly in the database rows)."""
    from sklearn.neighbors import KNeighborsClassifier
    rng = np.random.RandomState(11)
    db = rng.standard_normal((m_s, EMB_D)).astype(np.float32)
    q = rng.standard_normal((q_s, EMB_D)).astype(np.float32)
    knn = KNeighborsClassifier(n_neighbors=KNN_K, weights="uniform").fit(db, np.arange(m_s) % 1000)
    t0 = time.time()
    knn.predict_proba(q)
    dt = time.time() - t0
    return q_s / dt, dt


def run_distance_reference(a):
    """--impl reference for the distance workloads: the library calls the reference makes, on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    if a.workload == "pdist":
        n_s = 6000
        for i in range(a.warmup + a.steps):
            r, dt = cpu_pdist_rate(n_s if i >= a.warmup else 1000, threads)
            if i >= a.warmup:
                vals.append((r, dt))
        value = float(np.mean([v for v, _ in vals]))
        metric, unit = "pdist_pairs_per_sec", "pairs/s"
        sample = f"scipy.spatial.distance.pdist on {n_s} x {EMB_D} fp32 rows per step ({n_s * (n_s - 1) // 2} pairs; scipy pdist is single-threaded)"
        cores = 1
        cfg = {"workload": f"configs[3]: {a.pdist_n} x {EMB_D} pdist, sampled at N = {n_s}", "N": a.pdist_n, "sample_N": n_s}
    else:
        q_s, m_s = 256, 200000
        for i in range(a.warmup + a.steps):
            r, dt = cpu_knn_rate(q_s if i >= a.warmup else 16, m_s, threads)
            if i >= a.warmup:
                vals.append((r * m_s / a.knn_m, dt))
        value = float(np.mean([v for v, _ in vals]))
        metric, unit = "knn_queries_per_sec", "queries/s"
        sample = (f"sklearn KNeighborsClassifier(10).predict_proba on {q_s} queries x {m_s} database rows per step, rate scaled "
                  f"linearly to the {a.knn_m}-row database")
        cores = threads
        cfg = {"workload": f"configs[4]: {a.knn_q} queries vs {a.knn_m} x {EMB_D}, k = {KNN_K}, sampled", "Q": a.knn_q, "M": a.knn_m}
    dt = float(np.mean([d for _, d in vals]))
    print(json.dumps({"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
                      "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "fp64 (scipy) / fp32 (sklearn)", "data": "synthetic", "config": cfg,
                      "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "reference", "sample": sample},
                      "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_distance(a):
    """--workload pdist|knn: one JSON line with the same contract as the default line."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from facerec_b200._lib import Context
    ctx = Context(local)
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.time()
    rec = (bench_pdist if a.workload == "pdist" else bench_knn)(a, ctx, rank, world, a.steps, max(a.warmup, 3))
    t1 = time.time()
    if rank == 0:
        clocks = sampler.window(t0, t1) if sampler else {}
        if sampler:
            sampler.stop()
        line = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16 split (hi + lo), fp32 accumulate" if a.workload == "pdist" else
                "bf16 scan + fp32 exact re-rank (split-bf16 pass for undecidable queries)", "data": "synthetic"}
        line.update({k: rec[k] for k in rec if k not in line})
        line["clocks"] = clocks
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            if a.workload == "pdist":
                v, dt = cpu_pdist_rate(6000, threads)
                line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": 1, "kind": "reference",
                                        "sample": f"scipy pdist (the first half of linkage(vectors, 'complete'), facerec/cluster.py:41) on 6000 x 512 rows, {dt:.1f} s, single-threaded by construction"}
            else:
                v, dt = cpu_knn_rate(256, 200000, threads)
                line["cpu_baseline"] = {"value": v * 200000 / a.knn_m, "unit": "queries/s", "cores": threads, "kind": "reference",
                                        "sample": f"sklearn KNeighborsClassifier(10).predict_proba, 256 queries x 200000 rows ({dt:.1f} s), rate scaled to {a.knn_m} rows"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from facerec_b200._lib import Context
    from facerec_b200.pipeline import FacePipeline
    ctx = Context(local)   # raises without an sm_100 GPU / the built library: no fallback
    pipe = FacePipeline(a.backbone, min_face_size=20, save_every=a.save_every, ctx=ctx)
    H, W, B = a.height, a.width, a.batch
    # this rank's shard of the clip: the reference's split, shard_len frames per rank (extract.py:230-232)
    first = rank * 100000
    frames_h = make_batch(a.clip_seed, first, B, H, W)
    pinned = torch.from_numpy(frames_h).pin_memory()
    frames_d = pinned.to(pipe.device)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also plans the activation arena)
    res = None
    for _ in range(max(a.warmup, 3)):
        res = pipe.step_device(frames_d, first)
    faces_per_step = len(res.rects)
    dets_per_step = int(sum(len(d) for d in res.dets))

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- timed region 1: inputs resident in HBM ------------------------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fwd_ev = []
    barrier()
    l0 = ctx.launch_count()
    t0 = time.time()
    ev[0].record(stream)
    # two batches in flight (FacePipeline.run_pipelined, the schedule of facerec_b200.extract): detection of step k+1 is
    # enqueued before the host collects step k, so the GPU does not idle while the host picks the faces to embed.  Every
    # step's detection, crop, embedding and both result copies complete inside the timed region (the generator drains).
    for _ in pipe.run_pipelined((frames_d, first) for _ in range(a.steps)):
        pass
    ev[1].record(stream)
    barrier()
    t1 = time.time()
    launches = ctx.launch_count() - l0
    ms_total = ev[0].elapsed_time(ev[1])
    # the same K steps one after the other (step_device: each step waits for its own results before the next starts)
    es = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    es[0].record(stream)
    for _ in range(a.steps):
        pipe.step_device(frames_d, first)
    es[1].record(stream)
    torch.cuda.synchronize()
    ms_sequential = es[0].elapsed_time(es[1]) / a.steps

    # ---- RetinaFace forward alone, CUDA events on the launching stream (roofline) ---------
    from facerec_b200 import ops
    x16 = ops.preprocess_u8(ctx, frames_d, True, ops.input_layout(pipe.detector.model_id))
    ops.retina_forward(ctx, pipe.detector.model_id, x16, H, W)
    torch.cuda.synchronize()
    l1 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        ops.retina_forward(ctx, pipe.detector.model_id, x16, H, W)
    e1.record(stream)
    torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / a.steps
    fwd_launches = (ctx.launch_count() - l1) // a.steps
    del x16

    # ---- timed region 2: end to end from pinned host frames --------------------------------
    # Every step copies ITS OWN 64 frames host->device (pinned memory, a second stream so the copy of step k+1
    # overlaps the compute of step k, as the extractor's prefetch does) and reads detections + embeddings back.
    copy_stream = torch.cuda.Stream()
    NB = 4                                   # device buffers: two batches in flight on the GPU + the copy running ahead
    dbuf = [torch.empty_like(frames_d) for _ in range(NB)]
    ready = [torch.cuda.Event() for _ in range(NB)]
    released = [None] * NB                   # event after the last GPU work that reads buffer i

    def h2d(k):
        with torch.cuda.stream(copy_stream):
            if released[k % NB] is not None:
                copy_stream.wait_event(released[k % NB])
            dbuf[k % NB].copy_(pinned, non_blocking=True)
            ready[k % NB].record(copy_stream)

    def host_batches(n):
        h2d(0)
        for k in range(n):
            if k + 1 < n:
                h2d(k + 1)
            stream.wait_event(ready[k % NB])
            yield dbuf[k % NB], first

    def run_e2e(n):
        k = 0
        for r in pipe.run_pipelined(host_batches(n)):     # r: detections + embeddings of batch k on the host
            released[k % NB] = torch.cuda.Event()
            released[k % NB].record(stream)
            k += 1

    run_e2e(2)
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev2[0].record(stream)
    run_e2e(a.steps)
    ev2[1].record(stream)
    barrier()
    t2 = time.time()
    ms_e2e = ev2[0].elapsed_time(ev2[1])

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=pipe.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    # ---- the embedding-distance stage at its configured sizes (configs[3], configs[4]); every rank takes part
    dist_recs = {}
    if not a.no_distance:
        del dbuf, frames_d
        torch.cuda.empty_cache()
        for name, fn in (("pdist", bench_pdist), ("knn", bench_knn)):
            try:
                dist_recs[name] = fn(a, ctx, rank, world, 3, 2)
            except Exception as e:      # the headline number must survive a failure here; the failure is reported, not hidden
                dist_recs[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    if rank == 0:
        clocks = sampler.window(t0, t1) if sampler else {}
        if sampler and not clocks.get("samples"):      # nvidia-smi delivered nothing inside the first region: use both
            clocks = sampler.window(t0, t2)
        if sampler:
            sampler.stop()
        peak, peak_src = load_peaks()
        gflop = conv_gflop_retina(a.backbone, H, W)
        fwd_tflops = gflop * B / fwd_ms          # GFLOP / ms = TFLOP/s
        frames_total = world * B * a.steps
        h2d = int(frames_h.nbytes + faces_per_step * 20)
        d2h = int(B * pipe.max_out * 15 * 4 + 2 * B * 4 + faces_per_step * 512 * 4)
        line = {
            "metric": METRIC, "value": frames_total / (ms_total / 1e3), "unit": "frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (fp32 accumulate; heads/decode/NMS fp32)",
            "data": "synthetic", "config": workload_config(a),
            "faces_per_sec": world * faces_per_step * a.steps / (ms_total / 1e3),
            "faces_per_step_per_gpu": faces_per_step, "detections_per_step_per_gpu": dets_per_step,
            "e2e": {"value": frames_total / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / a.steps,
                    "api": "FacePipeline.run_pipelined (two batches in flight, as facerec_b200.extract runs) on frames copied from pinned host "
                           "memory every step (copy stream running ahead) + D2H of every step's detections and embeddings"},
            "ms_per_step_sequential": ms_sequential,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": fwd_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": fwd_tflops / peak,
                         **cited_traffic(a.backbone, B, H, W),
                         "kernel": f"one frb_retina_forward over {B} frames = {fwd_launches} launches: conv_tma_kernel (single-CTA and CTA-pair "
                                   "instantiations; stem with the max pool in its epilogue; heads included), 3 x conv_tail_kernel + 3 x conv_tail2_kernel "
                                   "(layer1 / layer2: conv3 + skip + next conv1 fused) and the zero fill of the pooled tensor's seam positions",
                         "algorithmic_gflop_per_frame": gflop, "launch_ms": fwd_ms, "peak_source": peak_src},
        }
        if dist_recs:
            line["distance"] = dist_recs
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            fps, faces, dt = cpu_reference_fps(frames_h[:a.cpu_baseline_frames], first, a.backbone, a.save_every, threads)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"first {a.cpu_baseline_frames} frame(s) of the same batch ({dt:.1f} s of CPU work), fp32 "
                                              f"PyTorch oracle, batch 1 per frame/face; {faces} faces embedded"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--backbone", default="r50", choices=["r50", "mnet025"])
    ap.add_argument("--save-every", type=int, default=5)
    ap.add_argument("--clip-seed", type=int, default=1234)
    ap.add_argument("--ref-frames", type=int, default=5,
                    help="frames per step of the CPU arm / cpu_baseline sample (5 = one save_every period: 5 detects, 1 embed pass)")
    ap.add_argument("--cpu-baseline-frames", type=int, default=25,
                    help="frames of the batch the cpu_baseline leg runs through the oracle (25 = five save_every periods, ~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="frames", choices=["frames", "pdist", "knn"],
                    help="frames = configs[1] (the headline); pdist = configs[3]; knn = configs[4]")
    ap.add_argument("--no-distance", action="store_true", help="default line without the `distance` object")
    ap.add_argument("--pdist-n", type=int, default=PDIST_N)
    ap.add_argument("--knn-q", type=int, default=KNN_Q)
    ap.add_argument("--knn-m", type=int, default=KNN_M)
    a = ap.parse_args()
    if a.workload != "frames":
        if a.impl == "reference":
            run_distance_reference(a)
        else:
            run_distance(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
