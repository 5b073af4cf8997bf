#! /usr/bin/env python3
"""Trajectory clustering with the distance matrix on the GPU: the reference's facerec/cluster.py with
one call swapped.

`linkage(vectors, method="complete")` (facerec/cluster.py:41, and :199 for the re-clustering of oversize
clusters) computes an implicit Euclidean pdist in fp64 on one CPU thread; here the condensed distance
vector comes from `frb_pdist` (split-bf16 tensor-core GEMM) and scipy's linkage /
fcluster consume it.  Everything after that — size-capped re-clustering, split_and_merge, relabel,
size-ordered labels, clusters.json — follows facerec/cluster.py:61-109,163-225.  Same CLI.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
from typing import Callable, Optional

import numpy as np
from scipy.cluster import hierarchy

from .formats import read_features, get_vectors, EMB_NAME


def read_trajectories(data_dir: str, vector_map, emb_name: str = EMB_NAME):
    """trajectories.jsonl -> (trajectory dicts, per-trajectory mean embedding) (facerec/cluster.py:19-31)."""
    trajectories, means = [], []
    with open(os.path.join(data_dir, "trajectories.jsonl")) as fh:
        for line in fh:
            t = json.loads(line)
            trajectories.append(t)
            means.append(get_vectors(t, vector_map, emb_name).mean(axis=0))
    return trajectories, np.array(means)


def cluster_once(vectors: np.ndarray, n_clusters: int, pdist_fn: Optional[Callable[[np.ndarray], np.ndarray]],
                 linkage_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> np.ndarray:
    """Complete-linkage clustering cut at n_clusters (facerec/cluster.py:33-45).  `linkage_fn(vectors) -> Z` does
    distances + linkage in one go (GPU: distance.gpu_linkage_complete); otherwise scipy links `pdist_fn(vectors)`."""
    if vectors.shape[0] == 0:
        return np.array([], dtype=np.int32)
    if vectors.shape[0] == 1:
        return np.array([1], dtype=np.int32)
    link = linkage_fn(vectors) if linkage_fn is not None else hierarchy.linkage(pdist_fn(vectors), method="complete")
    return hierarchy.fcluster(link, t=n_clusters, criterion="maxclust")


def split_and_merge(clusters: np.ndarray, min_size: int = 20, max_size: int = 40) -> np.ndarray:
    """Cap cluster sizes at max_size by splitting in index order, then bundle undersized neighbours
    (facerec/cluster.py:61-101; min_size is a preference, not a guarantee)."""
    out = np.zeros(clusters.size, dtype=np.int32)
    nxt = 0
    for ci, n in zip(*np.unique(clusters, return_counts=True)):
        idx = np.where(ci == clusters)[0]
        if n > max_size:
            parts = (n + max_size - 1) // max_size
            step = (n + parts - 1) // parts
            for i in range(parts):
                out[idx[i * step:(i + 1) * step]] = nxt
                nxt += 1
        else:
            out[idx] = nxt
            nxt += 1
    ids, counts = np.unique(out, return_counts=True)
    small = np.where(counts < min_size)[0]
    ids, counts = ids[small], counts[small]
    prev = 0
    for upper in range(1, small.size):
        if counts[prev:upper].sum() >= min_size or upper == small.size - 1:
            out[np.isin(out, ids[prev:upper])] = ids[prev]
            prev = upper
    return out


def relabel(clusters: np.ndarray) -> np.ndarray:
    """Dense labels 0..n-1 in sorted order of the old labels (facerec/cluster.py:103-109)."""
    return np.unique(clusters, return_inverse=True)[1].astype(np.int32)


def cluster_trajectories(trajectories, embeddings: np.ndarray, size: int, min_size: int, max_size: int,
                         pdist_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                         linkage_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> np.ndarray:
    """facerec/cluster.py:163-214: N//size clusters, re-cluster the oversize ones, label by descending size.
    Default: distance matrix AND complete linkage on the GPU; pass pdist_fn to keep scipy's linkage on the host."""
    if pdist_fn is None and linkage_fn is None:
        from .distance import gpu_linkage_complete
        linkage_fn = gpu_linkage_complete
    n = len(trajectories)
    n_clusters = n // size
    print(f"N={n} size={size} n_clusters={n_clusters}")
    clusters = cluster_once(embeddings, n_clusters, pdist_fn, linkage_fn)
    for ci, cnt in zip(*np.unique(clusters, return_counts=True)):
        if cnt > max_size:
            idx = np.where(ci == clusters)[0]
            sub = cluster_once(embeddings[idx], (cnt + max_size - 1) // max_size, pdist_fn, linkage_fn)
            sub = split_and_merge(sub, min_size, max_size)
            clusters[idx] = clusters.max() + 1 + sub
    clusters = relabel(clusters)
    ids, counts = np.unique(clusters, return_counts=True)
    print(f"Number of clusters: {len(ids)}")
    order = np.argsort(np.argsort(-counts))[ids]      # rank by size; ties keep the lower label first
    return order[clusters]


def write_clusters(clusters: np.ndarray, data_dir: str, movie_id: int):
    """clusters.json (facerec/cluster.py:216-225)."""
    path = os.path.join(data_dir, "clusters.json")
    with open(path, "w") as fh:
        json.dump({"clusters": [int(c) for c in clusters], "movie_id": movie_id}, fh, indent=None, separators=(",", ":"))
        fh.write("\n")
    print(f"Wrote trajectory clusters to: {path}")


def main(argv=None):
    p = argparse.ArgumentParser(allow_abbrev=True, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--size", type=int, default=18, help="preferred size of output clusters (in trajectory count)")
    p.add_argument("--min-size", type=int, default=12, help="preferred minimum size of output clusters")
    p.add_argument("--max-size", type=int, default=24, help="maximum size of output clusters")
    p.add_argument("--path", type=str, default=".", help="directory in which to find JSON files")
    args = p.parse_args(argv)
    for data_dir in glob.glob(args.path):
        data_dir = data_dir.rstrip("/")
        movie_id = int(os.path.basename(data_dir).split("-")[0])
        print(f"Clustering: {data_dir}")
        vector_map = read_features(data_dir)
        trajectories, means = read_trajectories(data_dir, vector_map)
        write_clusters(cluster_trajectories(trajectories, means, args.size, args.min_size, args.max_size), data_dir, movie_id)


if __name__ == "__main__":
    main()
