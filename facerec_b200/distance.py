"""Embedding-distance stage on the B200 path, behind the two library calls the reference makes.

* cluster.py:41 / :199  `linkage(vectors, method="complete")` -> `linkage(gpu_pdist(vectors), "complete")`
  (scipy accepts a condensed distance vector; everything after it — fcluster, split/merge, relabel —
  is unchanged host code).
* classify_knn.py:212,227  `KNeighborsClassifier(n_neighbors=k, weights="uniform").fit(X, y)` /
  `.classes_` / `.predict_proba(V)` -> `B200KNeighborsClassifier` with the same three members.

Multi-GPU: rows are sharded over ranks; the only exchange is one all_gather of the
embedding matrix (NCCL over NVLink via torch.distributed, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops, _lib
from ._lib import Context, default_context


def _ctx(ctx):
    return ctx if ctx is not None else default_context(0)


def gpu_pdist(vectors: np.ndarray, ctx: Optional[Context] = None, dtype=np.float64) -> np.ndarray:
    """Euclidean condensed distance vector (scipy pdist order) of float32 rows, computed on the GPU."""
    ctx = _ctx(ctx)
    X = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).to(torch.device("cuda", ctx.device))
    layout = _lib.FRB_CONDENSED_F64 if np.dtype(dtype) == np.float64 else _lib.FRB_CONDENSED_F32
    return ops.pdist(ctx, X, layout=layout).cpu().numpy()


def balanced_row_blocks(n: int, parts: int):
    """Split rows [0,n) into `parts` contiguous blocks with equal upper-triangle area (row i owns n-1-i entries)."""
    total = n * (n - 1) // 2
    bounds = [0]
    for p in range(1, parts):
        target = total * p / parts
        # entries in rows [0, r): r*n - r(r+1)/2  >= target
        r = int(np.ceil(((2 * n - 1) - np.sqrt(max((2 * n - 1) ** 2 - 8 * target, 0.0))) / 2))
        bounds.append(min(max(r, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1] - bounds[i]) for i in range(parts)]


def condensed_offset(n: int, i: int) -> int:
    """Index of entry (i, i+1) in the condensed vector."""
    return i * n - i * (i + 1) // 2


def all_gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather ragged row shards ([n_r, D] per rank) into the full [N, D] matrix on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def _gpu_row_block(ctx, X: torch.Tensor, r0: int, rows: int) -> torch.Tensor:
    """Condensed entries of rows [r0, r0+rows) (all j > i), fp32, computed by frb_pdist."""
    n = X.shape[0]
    beg, end = condensed_offset(n, r0), condensed_offset(n, r0 + rows)
    # the kernel addresses a full-size condensed vector: hand it a base shifted so this block lands at 0
    out = torch.empty((max(end - beg, 1),), dtype=torch.float32, device=X.device)
    ctx.check(ctx.lib.frb_pdist(ctx.handle, X.data_ptr(), n, X.shape[1], r0, rows, _lib.FRB_CONDENSED_F32,
                                out.data_ptr() - beg * 4, _lib.current_stream()), "frb_pdist")
    return out[:end - beg]


def pdist_sharded(local_rows: torch.Tensor, ctx: Optional[Context] = None, group=None, block_fn=None):
    """Row-sharded pdist: all_gather the embeddings (the path's one exchange step), then this rank
    fills its area-balanced row block.  Returns (this rank's slice of the condensed vector, its offset, N).
    `block_fn(X, r0, rows)` computes the slice; default = the CUDA kernel (tests inject a CPU stand-in)."""
    import torch.distributed as dist
    X = all_gather_rows(local_rows, group).contiguous()
    n = X.shape[0]
    r0, rows = balanced_row_blocks(n, dist.get_world_size(group))[dist.get_rank(group)]
    if block_fn is None:
        c = _ctx(ctx)
        block_fn = lambda X_, a, b: _gpu_row_block(c, X_, a, b)
    return block_fn(X, r0, rows), condensed_offset(n, r0), n


def knn_sharded(local_queries: torch.Tensor, db: torch.Tensor, k: int, ctx: Optional[Context] = None):
    """Query-sharded kNN: every rank holds the whole database and answers its own queries; no exchange
    is needed on the data path (the database is replicated once at load time)."""
    return ops.knn_topk(_ctx(ctx), local_queries, db, k)


class B200KNeighborsClassifier:
    """Brute-force Euclidean kNN with uniform votes: the subset of sklearn's KNeighborsClassifier that
    facerec/classify_knn.py:212-227 uses (fit, classes_, predict_proba) plus kneighbors()."""

    def __init__(self, n_neighbors: int = 5, weights: str = "uniform", ctx: Optional[Context] = None):
        if weights != "uniform":
            raise ValueError("only weights='uniform' is implemented (the reference's setting)")
        if not 1 <= n_neighbors <= 16:
            raise ValueError("n_neighbors must be in 1..16")
        self.n_neighbors = n_neighbors
        self.ctx = _ctx(ctx)
        self.device = torch.device("cuda", self.ctx.device)

    def fit(self, X, y):
        X = np.ascontiguousarray(X, dtype=np.float32)
        y = np.asarray(y)
        if X.ndim != 2 or len(X) != len(y) or len(X) < self.n_neighbors:
            raise ValueError("fit: X must be [M,D] with M == len(y) >= n_neighbors")
        self.classes_, inv = np.unique(y, return_inverse=True)
        self._db = torch.from_numpy(X).to(self.device)
        self._labels = torch.from_numpy(inv.astype(np.int32)).to(self.device)
        ops.knn_fit(self.ctx, self._db)        # the database's fp16 planes and norms are prepared once, not per query batch
        return self

    def kneighbors(self, V, n_neighbors: Optional[int] = None):
        k = n_neighbors or self.n_neighbors
        Q = torch.from_numpy(np.ascontiguousarray(V, dtype=np.float32)).to(self.device)
        idx, dist = ops.knn_topk(self.ctx, Q, self._db, k)
        return dist.cpu().numpy().astype(np.float64), idx.cpu().numpy().astype(np.int64)

    def predict_proba(self, V) -> np.ndarray:
        Q = torch.from_numpy(np.ascontiguousarray(V, dtype=np.float32)).to(self.device)
        idx, _ = ops.knn_topk(self.ctx, Q, self._db, self.n_neighbors)
        p32 = ops.knn_vote(self.ctx, idx, self._labels, len(self.classes_)).cpu().numpy()
        # the kernel returns count/k in fp32; sklearn's value is the float64 quotient of the same integer count
        return np.rint(p32.astype(np.float64) * self.n_neighbors) / self.n_neighbors

    def predict(self, V) -> np.ndarray:
        return self.classes_[np.argmax(self.predict_proba(V), axis=1)]


def _label_merges(pairs: np.ndarray, dist: np.ndarray, n: int) -> np.ndarray:
    """Raw NN-chain merges -> scipy's linkage matrix Z: stable sort by height, then union-find relabelling of the
    two children to cluster ids (new clusters are numbered n, n+1, ... in sorted order; lower id first)."""
    order = np.argsort(dist, kind="mergesort")
    pairs, dist = pairs[order], dist[order]
    parent = np.arange(2 * n - 1)
    size = np.ones(2 * n - 1, dtype=np.int64)
    Z = np.empty((n - 1, 4), dtype=np.float64)

    def find(x):
        root = x
        while parent[root] != root:
            root = parent[root]
        while parent[x] != root:
            parent[x], x = root, parent[x]
        return root

    nxt = n
    for k in range(n - 1):
        a, b = find(int(pairs[k, 0])), find(int(pairs[k, 1]))
        if a > b:
            a, b = b, a
        parent[a] = parent[b] = nxt
        size[nxt] = size[a] + size[b]
        Z[k] = (a, b, dist[k], size[nxt])
        nxt += 1
    return Z


def gpu_linkage_complete(vectors: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """scipy.cluster.hierarchy.linkage(vectors, method="complete") with both the distance matrix and the clustering on
    the GPU (facerec/cluster.py:41): returns the float64 linkage matrix Z that `fcluster` consumes.  Identical to
    scipy's Z for the same (fp32) distances, ties included."""
    ctx = _ctx(ctx)
    X = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).to(torch.device("cuda", ctx.device))
    n = X.shape[0]
    D = ops.pdist(ctx, X, layout=_lib.FRB_SQUARE_SYM_F32)
    pairs = torch.empty((n - 1, 2), dtype=torch.int32, device=X.device)
    dist = torch.empty((n - 1,), dtype=torch.float32, device=X.device)
    ctx.check(ctx.lib.frb_linkage_complete(ctx.handle, D.data_ptr(), n, pairs.data_ptr(), dist.data_ptr(),
                                           _lib.current_stream()), "frb_linkage_complete")
    return _label_merges(pairs.cpu().numpy(), dist.cpu().numpy().astype(np.float64), n)
