#! /usr/bin/env python3
"""Actor classification of trajectory clusters with the neighbour search on the GPU: the reference's
facerec/classify_knn.py with the classifier swapped.

`KNeighborsClassifier(n_neighbors=k, weights="uniform").fit(X, y)` / `.predict_proba(vectors)`
(facerec/classify_knn.py:212,227; brute-force Euclidean, one call per trajectory) becomes ONE fused
distance + top-k pass over every per-frame vector of the film (`frb_knn_topk` + `frb_knn_vote`), after
which the per-trajectory and per-cluster means, the p > 0.05 filter, the 5-decimal rounding and
predictions.json follow facerec/classify_knn.py:218-256.  Same CLI, same actor-images.zip input
(:40-52) and class-balanced up-sampling (:101-143).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import zipfile
from typing import Optional

import numpy as np

from .formats import read_features, get_vectors, EMB_NAME

ACTOR_ID_PREFIX = "momaf:elonet_henkilo_"


def read_actor_embeddings(zip_path: str, emb_name: str = EMB_NAME):
    """[(actor_id, vector)] from the <image>.json members that carry a face box (facerec/classify_knn.py:40-52)."""
    out = []
    with zipfile.ZipFile(zip_path) as z:
        for name in z.namelist():
            if name.endswith(".json"):
                d = json.loads(z.read(name))
                if "box" in d:
                    out.append((int(d["actorID"]), d["embeddings"][emb_name]))
    return out


def get_class_data(embeddings, min_samples: int = 20):
    """Class-balanced training set: every actor contributes max(min over actors, min_samples) vectors, short lists
    are repeated cyclically (facerec/classify_knn.py:101-143)."""
    if len(embeddings) == 0:
        return np.empty((0, 0), dtype=np.float32), np.empty((0,), dtype=np.int32)
    actors = {}
    for a, v in embeddings:
        actors.setdefault(a, []).append(v)
    dim = len(embeddings[-1][1])
    n_samples = max(min(len(v) for v in actors.values()), min_samples)
    X = np.zeros((len(actors) * n_samples, dim), dtype=np.float32)
    y = np.zeros(len(actors) * n_samples, dtype=np.int32)
    for i, (actor_id, vecs) in enumerate(actors.items()):
        reps = (n_samples + len(vecs) - 1) // len(vecs)
        X[i * n_samples:(i + 1) * n_samples] = np.array((vecs * reps)[:n_samples], dtype=np.float32)
        y[i * n_samples:(i + 1) * n_samples] = actor_id
    return X, y


def classify(data_dir: str, X: np.ndarray, y: np.ndarray, k: int, save_p_higher: float = 0.05, knn=None):
    """facerec/classify_knn.py:186-258 with all trajectories' vectors classified in one batched call."""
    movie_id = int(os.path.basename(data_dir).split("-")[0])
    predictions_file = os.path.join(data_dir, "predictions.json")
    vector_map = read_features(data_dir)
    with open(os.path.join(data_dir, "clusters.json")) as fh:
        clusters = np.array(json.load(fh)["clusters"])
    uniq = sorted(set(clusters))
    if len(np.unique(y)) < 3 or len(X) == 0:
        with open(predictions_file, "w") as fh:
            json.dump({int(ci): {} for ci in uniq}, fh)      # bare dict, as the reference writes it (:204-208)
        print(f"Not enough actor data. Wrote empty predictions: {predictions_file}")
        return
    if knn is None:
        from .distance import B200KNeighborsClassifier
        knn = B200KNeighborsClassifier(n_neighbors=k, weights="uniform")
    knn.fit(X, y)
    classes = knn.classes_
    per_traj = []
    with open(os.path.join(data_dir, "trajectories.jsonl")) as fh:
        for line in fh:
            per_traj.append(get_vectors(json.loads(line), vector_map, EMB_NAME))
    lens = [len(v) for v in per_traj]
    allv = np.concatenate([v for v in per_traj if len(v)], 0) if any(lens) else np.zeros((0, X.shape[1]), np.float32)
    proba = knn.predict_proba(allv) if len(allv) else np.zeros((0, len(classes)))
    trajectory_preds, pos = [], 0
    for n in lens:
        if n > 0:
            trajectory_preds.append(proba[pos:pos + n].mean(axis=0))
            pos += n
        else:
            trajectory_preds.append(np.zeros(X.shape[1], dtype=np.float32))   # the reference's quirk (:230)
    trajectory_preds = np.array(trajectory_preds)
    assert len(trajectory_preds) == len(clusters)
    cluster_preds = {}
    for ci in uniq:
        pred = trajectory_preds[np.where(clusters == ci)[0]].mean(axis=0)
        cluster_preds[int(ci)] = {f"{ACTOR_ID_PREFIX}{classes[i]}": round(float(pred[i]), 5)
                                  for i in np.where(pred > save_p_higher)[0]}
    with open(predictions_file, "w") as fh:
        json.dump({"movie_id": movie_id, "predictions": cluster_preds}, fh)
    print(f"Wrote predictions: {predictions_file}")


def main(argv=None):
    p = argparse.ArgumentParser(allow_abbrev=True, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--path", type=str, default=".", help="path to JSON data directory for a film")
    p.add_argument("--actors-dir", type=str, default=".", help="directory to find actor-images.zip")
    args = p.parse_args(argv)
    embeddings = read_actor_embeddings(args.actors_dir + "/actor-images.zip")
    for data_dir in glob.glob(args.path):
        data_dir = data_dir.rstrip("/")
        k, min_samples = 10, 20
        print(f"Predicting for {data_dir} using k-NN with k={k} and min_samples={min_samples}")
        X, y = get_class_data(embeddings, min_samples)
        classify(data_dir, X, y, k)
        break                              # the reference handles the first directory only (:295)


if __name__ == "__main__":
    main()
