"""Generate tests/golden/*.json by running the REFERENCE's own runnable code in this container.

Run here only (needs the read-only reference checkout, default /root/reference); the fixtures it writes are
committed so that tests can run anywhere (the GPU box has no reference checkout).  What is runnable:

  * `bbox_float_to_int`            facerec/extract.py:38-53   (function extracted by AST: the module itself
                                                                imports tensorflow / keras_facenet)
  * SORT                           facerec/sort.py            (imported verbatim; `filterpy.kalman` is absent
                                                                offline, so a shim restating filterpy's
                                                                published KalmanFilter predict/update is
                                                                injected before import)
  * `read_features`, `get_vectors` facerec/utils/utils.py:90-120
  * `cluster_trajectories`         facerec/cluster.py:163-214 (matplotlib stubbed)
  * `get_class_data`, `classify`   facerec/classify_knn.py:101-143,186-258
  * `SceneChangeDetector`          facerec/scene.py           (matplotlib stubbed)

Usage: python tools/make_golden.py [--reference /root/reference]
"""
import argparse
import ast
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
import types
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EMB_NAME = "20170512-110547"


class ShimKalmanFilter:
    """filterpy.kalman.KalmanFilter subset used by facerec/sort.py:96-121,134,157 (published equations)."""

    def __init__(self, dim_x, dim_z):
        self.x = np.zeros((dim_x, 1))
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.R = np.eye(dim_z)
        self._I = np.eye(dim_x)

    def predict(self):
        self.x = np.dot(self.F, self.x)
        self.P = np.dot(np.dot(self.F, self.P), self.F.T) + self.Q

    def update(self, z):
        y = z - np.dot(self.H, self.x)
        PHT = np.dot(self.P, self.H.T)
        S = np.dot(self.H, PHT) + self.R
        K = np.dot(PHT, np.linalg.inv(S))
        self.x = self.x + np.dot(K, y)
        I_KH = self._I - np.dot(K, self.H)
        self.P = np.dot(np.dot(I_KH, self.P), I_KH.T) + np.dot(np.dot(K, self.R), K.T)


def install_stubs():
    fp = types.ModuleType("filterpy")
    fk = types.ModuleType("filterpy.kalman")
    fk.KalmanFilter = ShimKalmanFilter
    fp.kalman = fk
    sys.modules["filterpy"] = fp
    sys.modules["filterpy.kalman"] = fk
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)


def extract_function(path, name):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {"np": np}
            exec(compile(ast.Module([node], []), path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def detection_script(seed, n_frames=60):
    """Synthetic detections: a few boxes drifting, with dropouts, a newcomer and a scene cut."""
    rng = np.random.RandomState(seed)
    tracks = [dict(x=rng.uniform(50, 500), y=rng.uniform(50, 250), w=rng.uniform(40, 120), vx=rng.uniform(-3, 3),
                   vy=rng.uniform(-2, 2), t0=int(rng.randint(0, 20)), t1=int(rng.randint(25, n_frames)))
              for _ in range(5)]
    frames = []
    cut = n_frames // 2 + int(rng.randint(0, 5))
    for f in range(n_frames):
        dets = []
        for t in tracks:
            if t["t0"] <= f < t["t1"] and rng.rand() > 0.12:
                x = t["x"] + t["vx"] * (f - t["t0"]) + rng.normal(0, 1.0)
                y = t["y"] + t["vy"] * (f - t["t0"]) + rng.normal(0, 1.0)
                w = t["w"] + rng.normal(0, 1.0)
                dets.append([x, y, x + w, y + 1.2 * w, 0.95])
        if rng.rand() < 0.08:
            x, y, w = rng.uniform(0, 550), rng.uniform(0, 280), rng.uniform(20, 60)
            dets.append([x, y, x + w, y + w, 0.95])
        frames.append({"dets": dets, "cut": f == cut})
    return frames


def run_sort(sort_mod, script, min_trajectory=3):
    """The tracker calls of facerec/extract.py:312-358 (update, kill on cut, delayed validity/box queries, pop_expired)."""
    sort_mod.KalmanBoxTracker.count = 0
    mt = sort_mod.Sort(max_age=5, min_hits=3, iou_threshold=0.5)
    out = {"ids": [], "queries": [], "expired": []}
    buf = []

    def dump(trks):
        for trk in trks:
            out["expired"].append({"first_frame": trk.first_frame, "initial_hits": trk.initial_hits,
                                   "boxes": [[float(v) for v in b] for b, _ in trk.history],
                                   "detected": [bool(d) for _, d in trk.history]})

    def query(item):
        f, ids = item
        for d in ids:
            valid = bool(mt.has_valid_tracker(d))
            box = [float(v) for v in mt.get_detection_bbox(d)]
            out["queries"].append({"frame": f, "det": int(d), "valid": valid, "box": box})

    for f, fr in enumerate(script):
        if fr["cut"]:
            mt.kill_trackers()
        dets = np.array(fr["dets"], dtype=float).reshape(-1, 5)
        ids = mt.update(dets, frame=f)
        out["ids"].append([int(i) for i in ids])
        buf.append((f, list(ids)))
        dump(mt.pop_expired(2 * min_trajectory, f))
        if len(buf) == min_trajectory:
            query(buf.pop(0))
    for item in buf:
        query(item)
    dump(mt.pop_expired(expiry_age=0))
    return out


def embedding_world(seed, n_traj=240, n_ident=14, dim=64):
    """Small film: trajectories with per-frame unit embeddings around identity centres + an actor DB."""
    rng = np.random.RandomState(seed)
    cent = rng.standard_normal((n_ident, dim))
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    feats, trajs = [], []
    frame = 0
    for t in range(n_traj):
        ident = int(rng.randint(0, n_ident))
        length = int(rng.randint(3, 14))
        x, y, w = int(rng.randint(0, 400)), int(rng.randint(0, 200)), int(rng.randint(50, 120))
        bbs = []
        for i in range(length):
            box = [x + i, y, x + i + w, y + w]
            bbs.append(box)
            if (frame + i) % 5 == 0 or i == 0:
                v = cent[ident] + 0.25 * rng.standard_normal(dim)
                v /= np.linalg.norm(v)
                feats.append({"frame": frame + i, "tag": f"1:{frame + i}:{box[0]}_{box[1]}_{box[2]}_{box[3]}",
                              "embeddings": {EMB_NAME: [float(np.float32(c)) for c in v]}, "box": box,
                              "keypoints": {}, "w": 640, "h": 360})
        trajs.append({"start": frame, "len": length, "bbs": bbs, "detected": [True] * length, "w": 640, "h": 360,
                      "object_type": "face", "index": t, "movie_id": 100001})
        frame += int(rng.randint(1, 6))
    actors = []
    for a in range(n_ident - 4):            # some identities have no actor images
        for j in range(int(rng.randint(6, 30))):
            v = cent[a] + 0.25 * rng.standard_normal(dim)
            v /= np.linalg.norm(v)
            actors.append({"actorID": 5000 + a * 7, "actorname": f"Actor {a}", "box": [0, 0, 10, 10],
                           "embeddings": {EMB_NAME: [float(np.float32(c)) for c in v]}})
    return feats, trajs, actors


def write_world(d, feats, trajs, actors):
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "features.jsonl"), "w") as f:
        for o in feats:
            f.write(json.dumps(o, separators=(",", ":")) + "\n")
    with open(os.path.join(d, "trajectories.jsonl"), "w") as f:
        for o in trajs:
            f.write(json.dumps(o, separators=(",", ":")) + "\n")
    with zipfile.ZipFile(os.path.join(d, "actor-images.zip"), "w") as z:
        for i, a in enumerate(actors):
            z.writestr(f"img{i:04d}.json", json.dumps(a))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    ref = os.path.join(a.reference, "facerec")
    os.makedirs(GOLD, exist_ok=True)
    install_stubs()
    sys.path.insert(0, ref)

    # 1. bbox_float_to_int
    fn = extract_function(os.path.join(ref, "extract.py"), "bbox_float_to_int")
    rng = np.random.RandomState(0)
    cases = []
    for _ in range(400):
        x1, y1 = rng.uniform(-30, 640), rng.uniform(-30, 360)
        w, h = rng.uniform(1, 300), rng.uniform(1, 300)
        if rng.rand() < 0.3:   # exact .5 boundaries exercise round-half-to-even
            x1, y1, w, h = np.floor(x1) + 0.5, np.floor(y1) + 0.5, np.floor(w), np.floor(h) + 0.5
        box = [float(x1), float(y1), float(x1 + w), float(y1 + h)]
        pad = float(rng.choice([0.0, 0.0, 0.10]))
        cases.append({"box": box, "w": 640, "h": 360, "padding": pad, "out": fn(box, 640, 360, pad)})
    json.dump(cases, open(os.path.join(GOLD, "bbox_float_to_int.json"), "w"))

    # 2. SORT
    import sort as ref_sort
    gold = []
    for seed in (1, 2, 3, 4):
        script = detection_script(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            res = run_sort(ref_sort, script)
        gold.append({"seed": seed, "script": script, "result": res})
    json.dump(gold, open(os.path.join(GOLD, "sort_runs.json"), "w"))

    # 3-5. readers, clustering, kNN classification on a small synthetic film
    import cluster as ref_cluster
    import classify_knn as ref_knn
    from utils.utils import read_features, get_vectors
    feats, trajs, actors = embedding_world(7)
    tmp = tempfile.mkdtemp()
    try:
        d = os.path.join(tmp, "100001-data")
        write_world(d, feats, trajs, actors)
        with contextlib.redirect_stdout(io.StringIO()):
            vm = read_features(d)
            trajectories, means = ref_cluster.read_trajectories(d, vm)
            clusters = ref_cluster.cluster_trajectories(trajectories, means, 18, 12, 24)
            ref_cluster.write_clusters(clusters, d, 100001)
            emb = ref_knn.read_actor_embeddings(os.path.join(d, "actor-images.zip"))
            X, y = ref_knn.get_class_data(emb, 20)
            ref_knn.classify(d, X, y, 10)
        gold = {"features": feats, "trajectories": trajs, "actors": actors,
                "mean_embeddings": [[float(c) for c in m] for m in means],
                "clusters_json": open(os.path.join(d, "clusters.json")).read(),
                "class_X_shape": list(X.shape), "class_y": [int(v) for v in y],
                "class_X_checksum": float(np.float64(X).sum()),
                "predictions_json": open(os.path.join(d, "predictions.json")).read()}
        json.dump(gold, open(os.path.join(GOLD, "film_small.json"), "w"))
    finally:
        shutil.rmtree(tmp)

    # 6. scene-cut detector
    import scene as ref_scene
    sys.path.insert(0, ROOT)
    from facerec_b200 import synth
    det = ref_scene.SceneChangeDetector(grayscale=False, crop=True, movie_id=100001)
    rows = []
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(60):
            cut = bool(det.update(synth.frame(77, i, 180, 320, vary_level=True, scene_len=7)))
            rows.append({"index": i, "cut": cut, "prev_mafd_eq": float(det.prev_mafd_eq or 0), "prev_fv_eq": float(det.prev_fv_eq or 0)})
    json.dump({"H": 180, "W": 320, "clip_seed": 77, "scene_len": 7, "rows": rows}, open(os.path.join(GOLD, "scene_cuts.json"), "w"))
    print("wrote", sorted(os.listdir(GOLD)))


# ---------------------------------------------------------------------------------------------------------
# 7. the reference's own process_video loop (facerec/extract.py:199-369) driven with scripted fakes
def scripted_faces(index, n_frames=48):
    """Deterministic per-frame detections (the fake detector's answer), keyed by frame index."""
    script = detection_script(11, n_frames)
    out = []
    for d in script[index % n_frames]["dets"]:
        x1, y1, x2, y2 = (float(np.float32(v)) for v in d[:4])
        kp = {n: (int(x1 + (i + 1) * (x2 - x1) / 6), int(y1 + (y2 - y1) / 2)) for i, n in
              enumerate(["left_eye", "right_eye", "nose", "mouth_left", "mouth_right"])}
        out.append({"box": [x1, y1, x2, y2], "keypoints": kp})
    return out


def frame_index_of(img):
    """Frames of the fixture clip carry their index in the blue channel of the top-left 4x4 block."""
    return int(img[0, 0, 2]) + 256 * int(img[0, 4, 2])


def fake_embedding(crop):
    c = crop.astype(np.float64)
    return np.array([c.mean(), c[..., 0].mean(), c[..., 2].std(), crop.shape[0], crop.shape[1], float(crop[0, 0, 1]),
                     float(crop[-1, -1, 0]), 1.0], dtype=np.float32)


def fixture_clip_frames(n=48, H=180, W=320):
    sys.path.insert(0, ROOT)
    from facerec_b200 import synth
    out = []
    for i in range(n):
        f = synth.frame(77, i, H, W, vary_level=True, scene_len=17).copy()
        f[0:4, 0:4, 2] = i % 256
        f[0:4, 4:8, 2] = i // 256
        out.append(f)
    return out


def write_fixture_clip(path, frames):
    import cv2
    H, W = frames[0].shape[:2]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 24.0, (W, H))
    assert vw.isOpened()
    for f in frames:
        vw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    vw.release()


def run_reference_extract(reference):
    ref = os.path.join(reference, "facerec")
    for name in ("tensorflow", "mtcnn", "retinaface", "keras_facenet", "keras_facenet.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mtcnn"].MTCNN = object
    sys.modules["retinaface"].RetinaFace = object

    def cropBox(image, detection, margin):      # keras_facenet.utils.cropBox semantics
        x, y, w, h = detection["box"]
        x -= margin; y -= margin; w += 2 * margin; h += 2 * margin
        if x < 0:
            w += x; x = 0
        if y < 0:
            h += y; y = 0
        return image[y:y + h, x:x + w]

    class FaceNet:
        def __init__(self, key=None):
            self.key = key

        def embeddings(self, crops):
            return np.stack([fake_embedding(c) for c in crops])

    sys.modules["keras_facenet"].FaceNet = FaceNet
    sys.modules["keras_facenet.utils"].cropBox = cropBox
    sys.modules["keras_facenet"].utils = sys.modules["keras_facenet.utils"]
    import importlib
    extract = importlib.import_module("extract")

    class FakeDetector:
        def detect(self, img):
            return scripted_faces(frame_index_of(img))

    gold = {}
    tmp = tempfile.mkdtemp()
    try:
        frames = fixture_clip_frames()
        clip = os.path.join(tmp, "100002-fixture.avi")
        write_fixture_clip(clip, frames)
        for n_shards, shard_i in ((1, 0), (2, 0), (2, 1)):
            import sort as ref_sort
            ref_sort.KalmanBoxTracker.count = 0
            extract.detector = FakeDetector()
            extract.embedders = {k: FaceNet(k) for k in extract.facenet_models}
            extract.multi_tracker = ref_sort.Sort(max_age=5, min_hits=3, iou_threshold=0.5)
            out = os.path.join(tmp, f"out_{n_shards}_{shard_i}")
            opt = extract.Options(out_path=out, n_shards=n_shards, shard_i=shard_i, save_every=5, min_trajectory=3,
                                  display_width=None, display_height=None, max_trajectory_age=5, save_images=False)
            with contextlib.redirect_stdout(io.StringIO()):
                extract.process_video(clip, opt)
            files = {}
            for root, _, names in os.walk(out):
                for nme in names:
                    files[os.path.relpath(os.path.join(root, nme), out)] = open(os.path.join(root, nme)).read()
            gold[f"{n_shards}_{shard_i}"] = files
    finally:
        shutil.rmtree(tmp)
    json.dump(gold, open(os.path.join(GOLD, "extract_runs.json"), "w"))
    print("extract golden:", {k: {n: len(v) for n, v in f.items()} for k, f in gold.items()})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ns = ap.parse_args()
    main()
    run_reference_extract(ns.reference)
