"""Readers of the on-disk formats the pipeline stages exchange, as the reference
reads them (facerec/utils/utils.py:90-120)."""
from __future__ import annotations

import json
import os

import numpy as np

EMB_NAME = "20170512-110547"     # the key cluster.py:17 and classify_knn.py:13 read


def read_features(data_dir: str):
    """features.jsonl -> {frame: {tuple(box): {model_key: vector}}} (facerec/utils/utils.py:90-105)."""
    print("Reading features.")
    vector_map = {}
    nb = 0
    with open(os.path.join(data_dir, "features.jsonl")) as fh:
        for line in fh:
            obj = json.loads(line)
            vector_map.setdefault(obj["frame"], {})[tuple(obj["box"])] = obj["embeddings"]
            nb += 1
    print(f"  features read for {len(vector_map)} frames and {nb} boxes")
    return vector_map


def get_vectors(trajectory: dict, vector_map, emb_name: str = EMB_NAME) -> np.ndarray:
    """All stored embeddings along a trajectory, float32 [n, D] (facerec/utils/utils.py:107-120)."""
    vectors = []
    for frame, bbs in enumerate(trajectory["bbs"], start=trajectory["start"]):
        hit = vector_map.get(frame, {}).get(tuple(bbs))
        if hit is not None:
            vectors.append(np.array(hit[emb_name], dtype=np.float32))
    return np.array(vectors, dtype=np.float32)
