#! /usr/bin/env python3
"""Per-shard face extractor: the reference's facerec/extract.py on the B200 path.

Same CLI flags and defaults (facerec/extract.py:376-387), same shard arithmetic with the
`max_trajectory_age` overlap (:230-232,279), same per-frame order of operations (:301-342) and the
same three per-shard outputs (:248-260,361-363), so `merge_shards.py`, `cluster.py` and
`classify_knn.py` of the reference consume the result unchanged:

    python -m facerec_b200.extract --n-shards 8 --shard-i $RANK --out-path ./data 100001-film.avi

What differs is *how* the work is scheduled: frames are decoded into batches, RetinaFace runs once
per batch on the GPU (the reference calls detect() per frame at batch 1), the sequential part
(scene cut -> SORT -> 3-frame delayed validity gate -> Kalman box -> integer crop rectangle) runs
on the host exactly in the reference's order, and the face crops of a whole batch are resized and
embedded in one InceptionResnetV1 pass from the frames still resident in HBM.  One process per GPU;
shards never talk to each other (the reference's SLURM-array model, scripts/extract.sh:12).

The reference embeds every face with four keras_facenet models and stores one vector per model key
(facerec/extract.py:24-25,166-167); downstream only reads `'20170512-110547'` (facerec/cluster.py:17,
facerec/classify_knn.py:13).  Here one 512-d model runs and its vector is written under every key in
`--emb-keys` (default: the reference's four), so readers keyed on any of them keep working.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import sys
from time import time
from typing import List, Optional

import numpy as np

FACENET_KEYS = ["20180402-114759", "20180408-102900", "20170511-185253", "20170512-110547"]
FACE_IMAGE_SIZE = 160
SAVE_FACE_PADDING = 0.10


def movie_id_from_path(path: str) -> int:
    """Leading digits of the basename, or all digit groups joined (facerec/extract.py:237-246)."""
    parts = os.path.splitext(os.path.basename(path))[0].split("-")
    if parts[0].isdigit():
        return int(parts[0])
    return int("".join(p for p in parts if p.isdigit()))


def shard_range(n_total: int, n_shards: int, shard_i: int, max_trajectory_age: int):
    """[beg, end) of the shard and the overlapped end actually processed (facerec/extract.py:230-232,279)."""
    assert 0 <= shard_i < n_shards, "Bad shard index."
    shard_len = (n_total + n_shards - 1) // n_shards
    beg = shard_len * shard_i
    end = min(beg + shard_len, n_total)
    return beg, end, min(end + max_trajectory_age, n_total)


# --------------------------------------------------------------------------------------- frame sources
class VideoSource:
    """cv2.VideoCapture reader with the reference's DAR handling (facerec/extract.py:204-228,284-300)."""

    def __init__(self, path: str, display_width: Optional[int] = None, display_height: Optional[int] = None):
        import cv2
        self.cv2 = cv2
        self.path = path
        self.cap = cv2.VideoCapture(path)
        self.n_frames = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))
        assert self.n_frames > 0, "Invalid video file <" + path + "> cwd=" + os.getcwd()
        self.fps = self.cap.get(cv2.CAP_PROP_FPS)
        self.video_w = int(self.cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.video_h = int(self.cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        if display_width is None or display_height is None:
            par = (self.cap.get(cv2.CAP_PROP_SAR_NUM) or 1.0) / (self.cap.get(cv2.CAP_PROP_SAR_DEN) or 1.0)
            self.height = self.video_h
            self.width = round(self.video_h * (self.video_w / self.video_h) * par)
        else:
            self.width, self.height = display_width, display_height

    def reopen(self) -> "VideoSource":
        """A second, independent handle on the same file (one per decoder thread)."""
        return VideoSource(self.path, self.width, self.height)

    def seek(self, index: int):
        assert self.cap.set(self.cv2.CAP_PROP_POS_FRAMES, index), f"Couldn't set start frame to: {index}"

    def read(self) -> Optional[np.ndarray]:
        ok, frame = self.cap.read()
        if not ok:
            return None
        if self.width != self.video_w or self.height != self.video_h:
            frame = self.cv2.resize(frame, (self.width, self.height))
        return self.cv2.cvtColor(frame, self.cv2.COLOR_BGR2RGB)

    def read_into(self, dst: np.ndarray) -> bool:
        """Same as read(), but the colour conversion writes straight into `dst` (a slot of a pinned batch buffer)."""
        ok, frame = self.cap.read()
        if not ok:
            return False
        if self.width != self.video_w or self.height != self.video_h:
            frame = self.cv2.resize(frame, (self.width, self.height))
        self.cv2.cvtColor(frame, self.cv2.COLOR_BGR2RGB, dst=dst)
        return True

    def close(self):
        self.cap.release()


class SyntheticSource:
    """facerec_b200.synth clip behind the same interface (path 'synth:<movie_id>:<n_frames>:<W>x<H>[:seed]')."""

    def __init__(self, spec: str):
        from . import synth
        f = spec.split(":")
        self.movie_id = int(f[1])
        self.n_frames = int(f[2])
        self.width, self.height = (int(v) for v in f[3].split("x"))
        self.seed = int(f[4]) if len(f) > 4 else 1234
        self.fps = 24.0
        self.video_w, self.video_h = self.width, self.height
        self._synth = synth
        self._pos = 0

    def seek(self, index: int):
        self._pos = index

    def read(self) -> Optional[np.ndarray]:
        if self._pos >= self.n_frames:
            return None
        fr = self._synth.frame(self.seed, self._pos, self.height, self.width)
        self._pos += 1
        return fr

    def close(self):
        pass


class PreloadedSource:
    """Frames of a synthetic clip generated up front (stands in for a decoder that keeps up): benchmarks only."""

    def __init__(self, movie_id: int, n_frames: int, width: int, height: int, seed: int = 1234):
        from . import synth
        self.movie_id, self.n_frames, self.width, self.height = movie_id, n_frames, width, height
        self.frames = [synth.frame(seed, i, height, width) for i in range(n_frames)]
        self._pos = 0

    def seek(self, index: int):
        self._pos = index

    def read(self) -> Optional[np.ndarray]:
        if self._pos >= self.n_frames:
            return None
        self._pos += 1
        return self.frames[self._pos - 1]

    def close(self):
        pass


_RING_CACHE = {}


def _pinned_ring(n_slots: int, shape, pin: bool):
    """Batch buffers the decoder threads fill.  Page-locking host memory costs ~0.25 s per GB (3.6 GB for nine 64 x 1080p
    slots), so the ring of a given shape is allocated once per process and reused by later shards / files."""
    import torch
    key = (tuple(shape), bool(pin))
    ring = _RING_CACHE.setdefault(key, [])
    while len(ring) < n_slots:
        ring.append(torch.empty(tuple(shape), dtype=torch.uint8, pin_memory=pin))
    return ring[:n_slots]


class BatchReader:
    """T decoder threads filling a ring of batch buffers in order.  Thread t decodes batches t, t + T, ... (each with its
    own handle on the source when T > 1, seeking to the batch's first frame); `next()` hands the batches out in order as
    (numpy view, torch view) of a ring slot, None at the end of the stream, and re-raises a decoder's exception.
    A thread may run at most T batches ahead of the consumer, which is what keeps a slot from being overwritten while
    the extractor still reads it (ring = T + 5 slots, see Extractor.process_video)."""

    def __init__(self, src, ring, batch: int, first_frame: int, n_frames: int, n_threads: int = 1):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.src, self.ring, self.batch = src, ring, batch
        self.first, self.n_frames = first_frame, n_frames
        self.n_batches = (n_frames + batch - 1) // batch
        self.T = max(1, min(n_threads, max(self.n_batches, 1)))
        self.cv = threading.Condition()
        self.ready = {}               # batch index -> frames decoded (or an exception)
        self.taken = 0                # batches handed to the consumer so far
        self.stop = False
        self.copiers = ThreadPoolExecutor(max_workers=4)     # numpy releases the GIL while it copies a frame into its slot
        self.threads = [threading.Thread(target=self._run, args=(t,), daemon=True) for t in range(self.T)]
        for th in self.threads:
            th.start()

    def _run(self, tid: int):
        src = None
        try:
            src = self.src if tid == 0 else self.src.reopen()
            into = getattr(src, "read_into", None)
            pos = None                                  # frame index the handle is positioned at
            for k in range(tid, self.n_batches, self.T):
                with self.cv:
                    while not self.stop and k >= self.taken + self.T + 1:
                        self.cv.wait()
                    if self.stop:
                        return
                start = self.first + k * self.batch
                want = min(self.batch, self.n_frames - k * self.batch)
                if pos != start and (self.T > 1 or k > 0):
                    src.seek(start)                     # thread 0 of a single-thread reader was positioned by the caller
                arr = self.ring[k % len(self.ring)].numpy()
                n = 0
                copies = []
                while n < want:
                    if into is not None:
                        if not into(arr[n]):
                            break
                    else:
                        fr = src.read()
                        if fr is None:
                            break
                        copies.append(self.copiers.submit(np.copyto, arr[n], fr))
                    n += 1
                for c in copies:
                    c.result()
                pos = start + n
                with self.cv:
                    self.ready[k] = n
                    self.cv.notify_all()
                if n < want:                            # the stream ended early: later batches do not exist
                    return
        except BaseException as e:                      # noqa: BLE001 - surface decode errors on the consumer side
            with self.cv:
                self.ready[-1] = e
                self.cv.notify_all()
        finally:
            if src is not None and tid != 0:
                src.close()

    def next(self):
        k = self.taken
        if k >= self.n_batches:
            return None
        with self.cv:
            while k not in self.ready and -1 not in self.ready:
                self.cv.wait()
            if k not in self.ready:
                raise self.ready[-1]
            n = self.ready.pop(k)
            self.taken += 1
            if n < min(self.batch, self.n_frames - k * self.batch):
                self.n_batches = k + 1              # short read: this is the last batch
            self.cv.notify_all()
        if n == 0:
            return None
        t = self.ring[k % len(self.ring)]
        return t.numpy()[:n], t[:n]

    def close(self):
        with self.cv:
            self.stop = True
            self.cv.notify_all()
        self.copiers.shutdown(wait=False)


# --------------------------------------------------------------------------------------- writers
def trajectory_record(trk, max_w: int, max_h: int) -> dict:
    """One trajectories.jsonl line (facerec/extract.py:55-80); `index`/`movie_id` are added by merge_shards.py."""
    from .pipeline import bbox_float_to_int
    return {"start": trk.first_frame, "len": len(trk.history),
            "bbs": [bbox_float_to_int(b, max_w, max_h) for b, _ in trk.history],
            "detected": [bool(d) for _, d in trk.history], "w": max_w, "h": max_h, "object_type": "face"}


def dump_line(obj, file):
    # json.dumps takes the C encoder (json.dump walks the object in Python); the text is identical
    file.write(json.dumps(obj, indent=None, separators=(",", ":")))
    file.write("\n")


_FMT_BUF = None


def format_f32_json(vec: np.ndarray) -> str:
    """json.dumps(vec.tolist(), separators=(",", ":")) for a float32 vector, formatted in libfacerec_b200.so
    (csrc/jsonfmt.cu): ~10 us instead of ~150 us per 512-float embedding."""
    global _FMT_BUF
    import ctypes as C
    from . import _lib
    v = np.ascontiguousarray(vec, dtype=np.float32).reshape(-1)
    cap = 2 + 33 * len(v)
    if _FMT_BUF is None or len(_FMT_BUF) < cap:
        _FMT_BUF = C.create_string_buffer(max(cap, 2 + 33 * 512))
    n = _lib.load_library().frb_format_f32_json(v.ctypes.data, len(v), _FMT_BUF, len(_FMT_BUF))
    if n < 0:
        raise RuntimeError("frb_format_f32_json failed")
    return C.string_at(_FMT_BUF, n).decode("ascii")


def feature_line(rec: dict, vec, keys) -> str:
    """One features.jsonl line (facerec/extract.py:186-195), byte-identical to
    json.dump({...,"embeddings": {k: vec for k in keys},...}, separators=(",", ":")): the 512-float vector is
    serialised once and spliced in under every key instead of being encoded len(keys) times.  `vec` is a list of
    Python floats or a float32 numpy vector (then the text comes from format_f32_json)."""
    d = lambda o: json.dumps(o, indent=None, separators=(",", ":"))
    v = format_f32_json(vec) if isinstance(vec, np.ndarray) else d(vec)
    emb = ",".join(d(k) + ":" + v for k in keys)
    return ('{"frame":' + d(rec["frame"]) + ',"tag":' + d(rec["tag"]) + ',"embeddings":{' + emb + '},"box":' + d(rec["box"]) +
            ',"keypoints":' + d(rec["keypoints"]) + ',"w":' + d(rec["w"]) + ',"h":' + d(rec["h"]) + "}")


def save_thumbnail(img_np, filtered_box, d_width, d_height, path):
    """Padded face thumbnail (facerec/extract.py:170-185): 10 % padded box, <=160 px, grayscale if R == G, JPEG q65."""
    from PIL import Image, ImageOps
    from .pipeline import bbox_float_to_int
    padded = Image.fromarray(img_np).crop(tuple(bbox_float_to_int(filtered_box, d_width, d_height, SAVE_FACE_PADDING)))
    padded.thumbnail((FACE_IMAGE_SIZE, FACE_IMAGE_SIZE), resample=Image.BILINEAR)
    a = np.array(padded).reshape((-1, 3))
    if np.all(a[:, 0] == a[:, 1]):
        padded = ImageOps.grayscale(padded)
    padded.save(path, quality=65)


# --------------------------------------------------------------------------------------- the extractor
class Extractor:
    def __init__(self, detector, embedder, tracker, out_path: str = "./data", n_shards: int = 256, shard_i: int = 0,
                 save_every: int = 5, min_trajectory: int = 3, max_trajectory_age: int = 5, save_images: bool = True,
                 emb_keys: Optional[List[str]] = None, batch: int = 64, display_width=None, display_height=None,
                 decode_threads: int = 1):
        self.detector, self.embedder, self.tracker = detector, embedder, tracker
        self.out_path = out_path.rstrip("/")
        self.n_shards, self.shard_i = n_shards, shard_i
        self.save_every, self.min_trajectory, self.max_trajectory_age = save_every, min_trajectory, max_trajectory_age
        self.save_images = save_images
        self.emb_keys = list(emb_keys) if emb_keys else list(FACENET_KEYS)
        self.batch = batch
        self.display_width, self.display_height = display_width, display_height
        self.decode_threads = decode_threads

    def process_video(self, file: str):
        import torch
        from .pipeline import bbox_float_to_int
        from .embedder import crop_rect
        from .scene import SceneChangeDetector, B200SceneChangeDetector

        if not isinstance(file, str):      # any object with the VideoSource interface (+ movie_id), e.g. preloaded frames
            src, synthetic = file, True
        else:
            synthetic = file.startswith("synth:")
            src = SyntheticSource(file) if synthetic else VideoSource(file, self.display_width, self.display_height)
        movie_id = src.movie_id if synthetic else movie_id_from_path(file)
        d_width, d_height = src.width, src.height
        beg, end, end_overlap = shard_range(src.n_frames, self.n_shards, self.shard_i, self.max_trajectory_age)
        src.seek(beg)

        base = f"{self.out_path}/{movie_id}-data"
        dirs = {k: f"{base}/{k}" for k in ("features", "trajectories", "scene_changes", "images")}
        for d in dirs.values():
            os.makedirs(d, exist_ok=True)
        suffix = f"{movie_id}_{beg}-{end}"
        features_file = open(f"{dirs['features']}/features_{suffix}.jsonl", "w")
        trajectories_file = open(f"{dirs['trajectories']}/trajectories_{suffix}.jsonl", "w")
        print(f"Shard {self.shard_i + 1} / {self.n_shards}: frames {beg} - {end} (+{end_overlap - end} overlap) of "
              f"{src.n_frames}, {d_width}x{d_height}, batches of {self.batch}, saving every 1/{self.save_every} frames")

        device = self.embedder.device
        # scene cuts: batched on the GPU when the frames live there, else the host restatement
        gpu_scene = device.type == "cuda"
        scene = (B200SceneChangeDetector(False, True, movie_id, ctx=getattr(self.embedder, "ctx", None)) if gpu_scene
                 else SceneChangeDetector(grayscale=False, crop=True, movie_id=movie_id))
        tracker = self.tracker
        scene_changes: List[int] = []
        buf = []                      # frames waiting for their tracker verdict (min_trajectory delay)
        counts = {"boxes": 0, "frames": 0, "traj": 0}
        device = self.embedder.device
        recent = {}                   # batch id -> frames on the GPU; delayed frames may live in an earlier batch

        def schedule(fd, pending):
            """process_frame_new (facerec/extract.py:140-197) up to the embedding call, which is deferred to the batch."""
            n = 0
            for face in fd["faces"]:
                if not tracker.has_valid_tracker(face["detection_id"]):
                    continue
                box = tracker.get_detection_bbox(face["detection_id"])
                tight = bbox_float_to_int(box, d_width, d_height)
                x, y, w, h = crop_rect([tight[0], tight[1], tight[2] - tight[0], tight[3] - tight[1]], d_width, d_height, 16)
                tag = fd["tag"] + ":{}_{}_{}_{}".format(*tight)
                rec = {"frame": fd["index"], "tag": tag, "embeddings": None, "box": tight,
                       "keypoints": face["keypoints"], "w": d_width, "h": d_height}
                pending.append((fd["slot"], (x, y, w, h), rec))
                if self.save_images:
                    save_thumbnail(fd["img_np"], box, d_width, d_height, f"{dirs['images']}/{tag}.jpeg")
                n += 1
            counts["boxes"] += n
            counts["frames"] += int(n > 0)

        def flush_sync(pending, batches):
            """Embedders without the launch / collect split (test doubles): embed and write in one go."""
            for bid, frames_d in batches.items():
                idx = [i for i, p in enumerate(pending) if p[0][0] == bid]
                if not idx:
                    continue
                rects = np.array([[pending[i][0][1], *pending[i][1]] for i in idx], np.int32)
                if ((rects[:, 3] <= 0) | (rects[:, 4] <= 0)).any():
                    raise ValueError("empty face crop (the reference would fail in cv2.resize here)")
                emb = self.embedder.embed_boxes(frames_d, rects)
                for i, e in zip(idx, emb):
                    pending[i][2]["embeddings"] = e.tolist()
            for _, _, rec in pending:
                features_file.write(feature_line(rec, rec["embeddings"], self.emb_keys))
                features_file.write("\n")
            pending.clear()

        def embed_launch(pending, batches):
            """Enqueue the embeddings of every pending face from the GPU-resident batch it came from (no waiting)."""
            handles = []
            for bid, frames_d in batches.items():
                idx = [i for i, p in enumerate(pending) if p[0][0] == bid]
                if not idx:
                    continue
                rects = np.array([[pending[i][0][1], *pending[i][1]] for i in idx], np.int32)
                if ((rects[:, 3] <= 0) | (rects[:, 4] <= 0)).any():
                    raise ValueError("empty face crop (the reference would fail in cv2.resize here)")
                handles.append((idx, self.embedder.embed_boxes_launch(frames_d, rects)))
            return handles

        def write_features(pending, handles):
            """Wait for the launched embeddings, then write the feature lines in order."""
            for idx, h in handles:
                emb = self.embedder.embed_collect(h)
                for i, e in zip(idx, emb):
                    pending[i][2]["embeddings"] = e if e.dtype == np.float32 else e.tolist()
            for _, _, rec in pending:
                features_file.write(feature_line(rec, rec["embeddings"], self.emb_keys))
                features_file.write("\n")
            pending.clear()

        # Frame ingest: decoder threads fill (pinned) batch buffers while the GPU works on earlier ones;
        # cv2 releases the GIL while decoding.  Thread t owns batches t, t + T, ...: it opens the file itself, seeks to
        # the first frame of its batch (the reference relies on the same CAP_PROP_POS_FRAMES seek for its shards,
        # facerec/extract.py:233) and decodes straight INTO the batch's ring slot (no np.stack, no pin_memory copy: both
        # were a full extra pass over 400 MB per 64 x 1080p batch).  One thread decodes 1080p mp4v at ~180 frames/s, far
        # below what the GPU path consumes, hence --decode-threads.
        from time import perf_counter
        t_setup0 = perf_counter()
        pin = device.type == "cuda"
        n_to_read = end_overlap - beg
        reopen = getattr(src, "reopen", None)
        n_threads = max(1, int(getattr(self, "decode_threads", 1))) if reopen is not None else 1
        # slots: T being filled + 1 ready + two in flight on the GPU + one walked by the host + one still referenced by
        # the min_trajectory-delayed frames of the batch before it
        ring = _pinned_ring(n_threads + 5, (self.batch, d_height, d_width, 3), pin)
        reader = BatchReader(src, ring, self.batch, beg, n_to_read, n_threads)
        from collections import deque
        from time import perf_counter
        tm = self.timing = {"wait_reader": 0.0, "launch": 0.0, "wait_detect": 0.0, "host_loop": 0.0, "wait_embed_write": 0.0,
                            "setup": perf_counter() - t_setup0, "loop": 0.0}
        t_loop0 = perf_counter()
        staged_embedder = hasattr(self.embedder, "embed_boxes_launch")
        staged_detector = hasattr(self.detector, "detect_batch_launch")
        inflight = deque()            # batches whose detection (and scene statistics) are running on the GPU
        state = {"next_frame": beg, "next_batch": 0, "done": False}

        def launch_next():
            """Take the next decoded batch, start its upload, detection and scene statistics; never waits for the GPU."""
            if state["done"]:
                return
            t0 = perf_counter()
            item = reader.next()
            t1 = perf_counter()
            tm["wait_reader"] += t1 - t0
            if item is None:
                state["done"] = True
                return
            if isinstance(item, BaseException):
                raise item
            frames_h, frames_t = item
            frames_d = frames_t.to(device, non_blocking=True)
            det_h = self.detector.detect_batch_launch(frames_d) if staged_detector else None
            scene_h = scene.stats_launch(frames_d) if gpu_scene else None
            inflight.append((state["next_batch"], state["next_frame"], frames_h, frames_d, det_h, scene_h))
            state["next_batch"] += 1
            state["next_frame"] += len(frames_h)
            tm["launch"] += perf_counter() - t1

        # Two batches in flight: while the host walks batch k (SORT, scheduling, JSON) the GPU runs detection of
        # batch k+1; embed(k) is enqueued before detect(k+2), so waiting for it never waits for a whole detection pass.
        launch_next()
        launch_next()
        while inflight:
            batch_id, f, frames_h, frames_d, det_h, scene_h = inflight.popleft()
            t0 = perf_counter()
            dets = self.detector.detect_batch_collect(det_h) if staged_detector else self.detector.detect_batch_raw(frames_d)
            cuts = scene.update_collect(scene_h) if gpu_scene else None
            t1 = perf_counter()
            pending = []
            for b in range(len(frames_h)):
                fi = f + b
                cut = cuts[b] if gpu_scene else scene.update(frames_h[b])
                faces = self.detector.format(dets[b])
                buf.append({"index": fi, "img_np": frames_h[b], "faces": faces, "tag": f"{movie_id}:{fi}", "slot": (batch_id, b)})
                if cut:
                    scene_changes.append(fi)
                    tracker.kill_trackers()
                detections = np.array([[*face["box"], 0.95] for face in faces]).reshape((-1, 5))
                ids = tracker.update(detections, frame=fi)
                for i, face in enumerate(faces):
                    face["detection_id"] = ids[i]
                for trk in tracker.pop_expired(2 * self.min_trajectory, fi):
                    dump_line(trajectory_record(trk, d_width, d_height), trajectories_file)
                    counts["traj"] += 1
                if len(buf) == self.min_trajectory:
                    fd = buf.pop(0)
                    if fd["index"] % self.save_every == 0:
                        schedule(fd, pending)
            recent[batch_id] = frames_d
            t2 = perf_counter()
            if staged_embedder:
                handles = embed_launch(pending, recent)
                launch_next()
                write_features(pending, handles)
            else:
                flush_sync(pending, recent)
                launch_next()
            t3 = perf_counter()
            tm["wait_detect"] += t1 - t0
            tm["host_loop"] += t2 - t1
            tm["wait_embed_write"] += t3 - t2
            oldest = min([fd["slot"][0] for fd in buf], default=batch_id)
            for k in [k for k in recent if k < oldest]:
                del recent[k]

        pending = []
        for fd in buf:
            if fd["index"] % self.save_every == 0:
                schedule(fd, pending)
        if staged_embedder:
            write_features(pending, embed_launch(pending, recent))
        else:
            flush_sync(pending, recent)
        for trk in tracker.pop_expired(expiry_age=0):
            dump_line(trajectory_record(trk, d_width, d_height), trajectories_file)
            counts["traj"] += 1
        tm["loop"] = perf_counter() - t_loop0
        with open(f"{dirs['scene_changes']}/scene_changes_{suffix}.json", "w") as fh:
            json.dump({"frame_indices": [c for c in scene_changes if beg <= c < end]}, fh, indent=None, separators=(",", ":"))
        features_file.close()
        trajectories_file.close()
        reader.close()
        src.close()
        print(f"Saved {counts['boxes']} boxes from {counts['frames']} different frames")
        print(f"and {counts['traj']} trajectories.")
        return counts


def read_display_size(file_basename: str, aspects_path: str = "aspect_ratios.csv"):
    """Optional ./aspect_ratios.csv override (facerec/extract.py:398-413)."""
    if not os.path.exists(aspects_path):
        return None, None
    with open(aspects_path, "r") as fh:
        reader = csv.reader(fh, delimiter=",", quotechar='"')
        headers = next(reader)
        name_i, w_i, h_i = (headers.index(c) for c in ("filename", "display_width", "display_height"))
        for row in reader:
            if row[name_i] == file_basename:
                return int(row[w_i]), int(row[h_i])
    return None, None


def main(argv=None):
    p = argparse.ArgumentParser(allow_abbrev=True, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--n-shards", type=int, default=256, help="sets the number of shards")
    p.add_argument("--shard-i", type=int, required=True, help="indicates specific shard")
    p.add_argument("--save-every", type=int, default=5, help="interval between saved frame images")
    p.add_argument("--iou-threshold", type=float, default=0.5, help="required area overlap between matches")
    p.add_argument("--min-trajectory", type=int, default=3, help="minimum allowed trajectory length")
    p.add_argument("--max-trajectory-age", type=int, default=5, help="maximum allowed break in trajectory")
    p.add_argument("--min-face-size", type=int, default=20, help="minimum allowed face size in unknown units")
    p.add_argument("--face-threshold", type=float, default=0.95, help="minimum allowed face detection score")
    p.add_argument("--out-path", type=str, default="./data", help="storage directory")
    p.add_argument("--no-images", action="store_true", help="if set, no images are stored")
    # additions of the B200 path
    p.add_argument("--backbone", default="r50", choices=["r50", "mnet025"], help="RetinaFace backbone")
    p.add_argument("--batch", type=int, default=64, help="frames per GPU batch")
    p.add_argument("--device", type=int, default=int(os.environ.get("LOCAL_RANK", "0")), help="CUDA device of this shard")
    p.add_argument("--emb-keys", nargs="*", default=FACENET_KEYS, help="keys the embedding is stored under")
    p.add_argument("--detector-weights", default=None,
                   help="RetinaFace weights: .frbw blob, .npz under this repo's layer names, or a PyTorch state dict under the "
                        "public RetinaFace names (facerec_b200/convert.py)")
    p.add_argument("--embedder-weights", default=None,
                   help="InceptionResnetV1 weights (.frbw / .npz / PyTorch state dict with the public FaceNet names)")
    p.add_argument("--weight-seed", type=int, default=0,
                   help="seed of the RANDOM-INIT weights used when no weight file is given (benchmarks and parity tests only)")
    p.add_argument("--decode-threads", type=int, default=1, help="decoder threads feeding the pinned frame ring (each opens the file and decodes its own frame ranges)")
    p.add_argument("file")
    args = p.parse_args(argv)

    start = time()
    from ._lib import Context
    from .detector import B200RetinaFaceDetector
    from .embedder import B200FaceNet
    from .sort import NativeSort as Sort   # csrc/sort.cu; the numpy restatement in sort.py stays as the cross-check
    from . import convert
    ctx = Context(args.device)
    det_w = convert.load_weights(args.detector_weights, "retina_" + args.backbone) if args.detector_weights else None
    emb_w = convert.load_weights(args.embedder_weights, "irv1") if args.embedder_weights else None
    if det_w is None or emb_w is None:
        import sys
        print("WARNING: running on seeded RANDOM-INIT weights (" +
              ", ".join(n for n, w in (("detector", det_w), ("embedder", emb_w)) if w is None) +
              "): detections and embeddings are only meaningful for benchmarking and parity tests. "
              "Pass --detector-weights / --embedder-weights for real use; stored actor embeddings must come from the same "
              "embedder weights.", file=sys.stderr)
    detector = B200RetinaFaceDetector(args.min_face_size, args.backbone, weights=det_w, seed=args.weight_seed,
                                      threshold=args.face_threshold, ctx=ctx)
    embedder = B200FaceNet(weights=emb_w, seed=args.weight_seed, ctx=ctx)
    tracker = Sort(max_age=args.max_trajectory_age, min_hits=args.min_trajectory, iou_threshold=args.iou_threshold)
    dw, dh = read_display_size(os.path.basename(args.file))
    ex = Extractor(detector, embedder, tracker, args.out_path, args.n_shards, args.shard_i, args.save_every,
                   args.min_trajectory, args.max_trajectory_age, not args.no_images, args.emb_keys, args.batch, dw, dh,
                   decode_threads=args.decode_threads)
    ex.process_video(args.file)
    minutes, seconds = divmod(time() - start, 60)
    print(f"Completed in {int(minutes)} minutes, {int(seconds)} seconds.")


if __name__ == "__main__":
    main()
