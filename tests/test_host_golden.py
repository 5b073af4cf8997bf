"""Host logic against golden fixtures produced by the REFERENCE's own runnable code (tools/make_golden.py):
bbox arithmetic, SORT, scene cuts, format readers, clustering / kNN aggregation, and the whole per-shard
driver loop.  CPU only; the distance calls are injected with scipy / scikit-learn (the libraries the
reference itself calls) so that only the host logic is under test here — the GPU kernels behind the
same interfaces are covered by tests/test_gpu_*.py."""
import io
import json
import os
import contextlib

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


def test_bbox_float_to_int_matches_reference():
    from facerec_b200.pipeline import bbox_float_to_int, boxes_to_rects
    from facerec_b200.embedder import crop_rect
    cases = _gold("bbox_float_to_int.json")
    for c in cases:
        assert bbox_float_to_int(c["box"], c["w"], c["h"], c["padding"]) == c["out"], c
    # the vectorised rectangle path used by the batched pipeline = tight box -> xywh -> cropBox rectangle
    plain = [c for c in cases if c["padding"] == 0.0]
    rects = boxes_to_rects(np.array([c["box"] for c in plain], np.float32), np.zeros(len(plain), np.int64), 640, 360)
    for c, r in zip(plain, rects):
        t = c["out"]
        assert tuple(r[1:]) == crop_rect([t[0], t[1], t[2] - t[0], t[3] - t[1]], 640, 360, 16)


def _run_sort(script, native=False):
    from facerec_b200 import sort as S
    S.KalmanBoxTracker.count = 0
    mt = (S.NativeSort if native else S.Sort)(max_age=5, min_hits=3, iou_threshold=0.5)
    out = {"ids": [], "queries": [], "expired": []}
    buf = []

    def dump(trks):
        for t in trks:
            out["expired"].append({"first_frame": t.first_frame, "initial_hits": getattr(t, "initial_hits", None),
                                   "boxes": [[float(v) for v in b] for b, _ in t.history],
                                   "detected": [bool(d) for _, d in t.history]})

    def query(item):
        f, ids = item
        for d in ids:
            out["queries"].append({"frame": f, "det": int(d), "valid": bool(mt.has_valid_tracker(d)),
                                   "box": [float(v) for v in mt.get_detection_bbox(d)]})

    for f, fr in enumerate(script):
        if fr["cut"]:
            mt.kill_trackers()
        ids = mt.update(np.array(fr["dets"], dtype=float).reshape(-1, 5), frame=f)
        out["ids"].append([int(i) for i in ids])
        buf.append((f, list(ids)))
        dump(mt.pop_expired(6, f))
        if len(buf) == 3:
            query(buf.pop(0))
    for item in buf:
        query(item)
    dump(mt.pop_expired(expiry_age=0))
    return out


def test_sort_matches_reference_runs():
    for run in _gold("sort_runs.json"):
        got = _run_sort(run["script"])
        ref = run["result"]
        assert got["ids"] == ref["ids"]
        assert len(got["queries"]) == len(ref["queries"]) and len(got["expired"]) == len(ref["expired"])
        for g, r in zip(got["queries"], ref["queries"]):
            assert (g["frame"], g["det"], g["valid"]) == (r["frame"], r["det"], r["valid"])
            assert g["box"] == r["box"]                      # bit-identical Kalman posterior
        for g, r in zip(got["expired"], ref["expired"]):
            assert g == r


def test_native_sort_matches_reference_runs():
    """csrc/sort.cu behind the C ABI: ids, validity and every integer box the extractor derives are identical to the
    reference runs; float boxes agree to 1e-9 (plain loops vs. LAPACK / BLAS in the 4x4 inverse and 8x8 products)."""
    from facerec_b200.pipeline import bbox_float_to_int
    for run in _gold("sort_runs.json"):
        got = _run_sort(run["script"], native=True)
        ref = run["result"]
        assert got["ids"] == ref["ids"]
        assert len(got["queries"]) == len(ref["queries"]) and len(got["expired"]) == len(ref["expired"])
        for g, r in zip(got["queries"], ref["queries"]):
            assert (g["frame"], g["det"], g["valid"]) == (r["frame"], r["det"], r["valid"])
            np.testing.assert_allclose(g["box"], r["box"], rtol=1e-9, atol=1e-9)
            assert bbox_float_to_int(g["box"], 640, 360) == bbox_float_to_int(r["box"], 640, 360)
        for g, r in zip(got["expired"], ref["expired"]):
            assert g["first_frame"] == r["first_frame"] and g["detected"] == r["detected"]
            np.testing.assert_allclose(np.array(g["boxes"]), np.array(r["boxes"]), rtol=1e-9, atol=1e-9)
            assert [bbox_float_to_int(b, 640, 360) for b in g["boxes"]] == [bbox_float_to_int(b, 640, 360) for b in r["boxes"]]


def test_scene_detector_matches_reference():
    from facerec_b200 import synth, scene
    g = _gold("scene_cuts.json")
    det = scene.SceneChangeDetector(False, True, 1)
    cuts = 0
    for r in g["rows"]:
        c = det.update(synth.frame(g["clip_seed"], r["index"], g["H"], g["W"], vary_level=True, scene_len=g["scene_len"]))
        assert c == r["cut"]
        assert abs((det.prev_mafd_eq or 0) - r["prev_mafd_eq"]) < 1e-9
        cuts += c
    assert cuts >= 3


def _film_dir(tmp_path, film):
    d = tmp_path / "100001-data"
    d.mkdir()
    with open(d / "features.jsonl", "w") as fh:
        for o in film["features"]:
            fh.write(json.dumps(o, separators=(",", ":")) + "\n")
    with open(d / "trajectories.jsonl", "w") as fh:
        for o in film["trajectories"]:
            fh.write(json.dumps(o, separators=(",", ":")) + "\n")
    import zipfile
    with zipfile.ZipFile(d / "actor-images.zip", "w") as z:
        for i, a in enumerate(film["actors"]):
            z.writestr(f"img{i:04d}.json", json.dumps(a))
    return str(d)


def test_cluster_and_classify_host_logic_match_reference(tmp_path):
    from scipy.spatial.distance import pdist
    from sklearn.neighbors import KNeighborsClassifier
    from facerec_b200 import cluster, classify_knn, formats
    film = _gold("film_small.json")
    d = _film_dir(tmp_path, film)
    with contextlib.redirect_stdout(io.StringIO()):
        vm = formats.read_features(d)
        traj, means = cluster.read_trajectories(d, vm)
        assert np.array_equal(means.astype(np.float32), np.array(film["mean_embeddings"], np.float32))
        labels = cluster.cluster_trajectories(traj, means, 18, 12, 24, pdist_fn=lambda v: pdist(v.astype(np.float64)))
        cluster.write_clusters(labels, d, 100001)
        assert open(os.path.join(d, "clusters.json")).read() == film["clusters_json"]
        emb = classify_knn.read_actor_embeddings(os.path.join(d, "actor-images.zip"))
        X, y = classify_knn.get_class_data(emb, 20)
        assert list(X.shape) == film["class_X_shape"] and [int(v) for v in y] == film["class_y"]
        assert float(np.float64(X).sum()) == film["class_X_checksum"]
        classify_knn.classify(d, X, y, 10, knn=KNeighborsClassifier(n_neighbors=10, weights="uniform"))
    assert open(os.path.join(d, "predictions.json")).read() == film["predictions_json"]


def test_split_and_merge_and_relabel_edge_cases():
    from facerec_b200 import cluster
    c = np.array([1] * 50 + [2] * 3 + [3] * 4 + [4] * 30, dtype=np.int32)
    out = cluster.split_and_merge(c, 12, 24)
    ids, counts = np.unique(out, return_counts=True)
    assert counts.max() <= 24 and counts.sum() == len(c)
    assert np.array_equal(cluster.relabel(np.array([7, 3, 7, 9])), [1, 0, 1, 2])
    assert len(cluster.cluster_once(np.zeros((0, 4)), 3, None)) == 0
    assert list(cluster.cluster_once(np.zeros((1, 4)), 3, None)) == [1]


# ------------------------------------------------------------------------------------------------ driver loop
def _scripted_faces(index, n_frames=48):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLD), "..", "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


class _FakeDetector:
    def __init__(self, mg):
        self.mg = mg

    def detect_batch_raw(self, frames):
        return [self.mg.frame_index_of(f.numpy()) for f in frames]

    def format(self, index):
        return self.mg.scripted_faces(index)


class _FakeEmbedder:
    def __init__(self, mg):
        import torch
        self.mg = mg
        self.device = torch.device("cpu")

    def embed_boxes(self, frames, rects):
        return np.stack([self.mg.fake_embedding(frames[f, y:y + h, x:x + w].numpy()) for f, x, y, w, h in rects])


@pytest.mark.parametrize("n_shards,shard_i,batch,native,threads", [(1, 0, 64, False, 1), (1, 0, 7, True, 1), (2, 0, 5, False, 1),
                                                                    (2, 1, 16, True, 1), (1, 0, 7, True, 3), (2, 1, 5, True, 4)])
def test_extractor_reproduces_reference_process_video(tmp_path, n_shards, shard_i, batch, native, threads):
    """facerec/extract.py:199-369 was run verbatim (fake detector / embedder) by tools/make_golden.py on a lossless
    clip; the batched extractor must write byte-identical features / trajectories / scene_changes files, with one
    decoder thread or several (each seeks to its own batches of the file)."""
    from facerec_b200 import extract, sort as S
    mg = _scripted_faces(0)
    clip = str(tmp_path / "100002-fixture.avi")
    mg.write_fixture_clip(clip, mg.fixture_clip_frames())
    S.KalmanBoxTracker.count = 0
    ex = extract.Extractor(_FakeDetector(mg), _FakeEmbedder(mg), (S.NativeSort if native else S.Sort)(5, 3, 0.5), str(tmp_path / "out"), n_shards, shard_i,
                           save_every=5, min_trajectory=3, max_trajectory_age=5, save_images=False, batch=batch,
                           decode_threads=threads)
    with contextlib.redirect_stdout(io.StringIO()):
        ex.process_video(clip)
    gold = _gold("extract_runs.json")[f"{n_shards}_{shard_i}"]
    for rel, text in gold.items():
        got = open(os.path.join(str(tmp_path / "out"), rel)).read()
        assert got == text, rel


def test_shard_arithmetic_and_movie_id():
    from facerec_b200.extract import shard_range, movie_id_from_path
    assert shard_range(720, 1, 0, 5) == (0, 720, 720)
    assert shard_range(721, 8, 7, 5) == (637, 721, 721)
    assert shard_range(721, 8, 0, 5) == (0, 91, 96)
    assert movie_id_from_path("/x/100001-synthA.avi") == 100001
    assert movie_id_from_path("film-12-part-7.mp4") == 127


def test_feature_line_is_byte_identical_to_json_dump():
    import io as _io
    from facerec_b200.extract import feature_line, FACENET_KEYS
    rng = np.random.RandomState(0)
    vec = rng.standard_normal(512).astype(np.float32).tolist()
    rec = {"frame": 35, "tag": "100001:35:10_20_110_140", "box": [10, 20, 110, 140],
           "keypoints": {"left_eye": (30, 50), "right_eye": (70, 51), "nose": (50, 80), "mouth_left": (35, 110), "mouth_right": (66, 111)},
           "w": 640, "h": 360}
    ref = _io.StringIO()
    json.dump({"frame": rec["frame"], "tag": rec["tag"], "embeddings": {k: vec for k in FACENET_KEYS}, "box": rec["box"],
               "keypoints": rec["keypoints"], "w": 640, "h": 360}, ref, indent=None, separators=(",", ":"))
    assert feature_line(rec, vec, FACENET_KEYS) == ref.getvalue()


def test_native_float_json_is_byte_identical_to_json_dumps():
    """csrc/jsonfmt.cu: the text of a float32 vector equals json.dumps(v.tolist()) over 30 decades of magnitude and the
    repr layout boundaries (1e-4, 1e16, zeros, subnormals, float32 max)."""
    from facerec_b200.extract import format_f32_json, feature_line, FACENET_KEYS
    rng = np.random.RandomState(0)
    v = np.concatenate([rng.standard_normal(2000).astype(np.float32) * np.float32(s)
                        for s in (1, 1e-3, 1e-6, 1e-12, 1e5, 1e10, 1e18, 1e-30, 1e30)] +
                       [np.array([0.0, -0.0, 1.0, -1.0, 0.5, 1e16, 9.999999e15, 1e-4, 9.9999e-5, 123456.0, 1e22, 3.4028235e38,
                                  1.4e-45, 1e-5, 2.0 ** 24, 16777217.0, 0.1, 100.0, 1e15], np.float32)])
    assert format_f32_json(v) == json.dumps(v.tolist(), separators=(",", ":"))
    e = rng.standard_normal(512).astype(np.float32)
    e /= np.linalg.norm(e)
    rec = {"frame": 35, "tag": "100001:35:10_20_110_140", "box": [10, 20, 110, 140],
           "keypoints": {"left_eye": (30, 50), "right_eye": (70, 51), "nose": (50, 80), "mouth_left": (35, 110), "mouth_right": (66, 111)},
           "w": 640, "h": 360}
    assert feature_line(rec, e, FACENET_KEYS) == feature_line(rec, e.tolist(), FACENET_KEYS)


def test_native_sort_rejects_non_finite_detections_like_scipy():
    """A NaN box used to spin forever inside the C++ assignment (round-1 advisor finding); the reference raises
    ValueError from scipy's linear_sum_assignment on the same input (facerec/sort.py:31)."""
    from facerec_b200 import sort as S
    for cls in (S.NativeSort, S.Sort):
        S.KalmanBoxTracker.count = 0
        mt = cls(5, 3, 0.5)
        mt.update(np.array([[10, 10, 50, 50, 0.95]], float), frame=0)
        with pytest.raises(ValueError):
            mt.update(np.array([[np.nan, 10, 50, 50, 0.95]], float), frame=1)
