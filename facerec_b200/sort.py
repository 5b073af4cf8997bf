"""Host-side SORT tracker for the extractor: same interface and arithmetic as the
reference's modified SORT (facerec/sort.py), without the filterpy dependency.

It sits between detect and embed on the per-frame path: it decides which detections get embedded
(`has_valid_tracker`, facerec/sort.py:297-308) and supplies the Kalman-posterior box that is cropped
(`get_detection_bbox`, facerec/sort.py:321-331).  It is sequential, tiny (tens of boxes per frame) and
CPU-natural, so it is host code, not a kernel.

Kalman model (facerec/sort.py:96-121): state [cx, cy, area, ratio, d/dt of those], constant
velocity, measurement = first four; R = diag(1,1,10,10), P0 = diag(10,10,10,10,1e4,1e4,1e4,1e4),
Q = diag(1,1,1,1,.01,.01,.01,1e-4).  predict / update are the textbook equations filterpy's
KalmanFilter implements (P = FPF' + Q;  K = PH'(HPH'+R)^-1, x += K(z-Hx), Joseph-form P update),
evaluated with the same dense float64 products so results agree to the last bit with a filterpy-backed
run of the reference file (tests/golden/sort_*.json were produced that way by tools/make_golden.py).
"""
from __future__ import annotations

import sys
from typing import Dict, List, Optional

import numpy as np
from scipy.optimize import linear_sum_assignment

_F = np.eye(8)
_F[:4, 4:] = np.eye(4)
_H = np.eye(4, 8)
_R = np.diag([1.0, 1.0, 10.0, 10.0])
_P0 = np.diag([10.0, 10.0, 10.0, 10.0, 10000.0, 10000.0, 10000.0, 10000.0])
_Q = np.diag([1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.01, 0.0001])
_I8 = np.eye(8)


def box_to_z(b):
    """[x1,y1,x2,y2] -> column [cx, cy, area, w/h] (facerec/sort.py:55-67)."""
    w = b[2] - b[0]
    h = b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w * h, w / float(h)]).reshape(4, 1)


def state_to_box(x):
    """state column -> [x1,y1,x2,y2] (facerec/sort.py:70-79)."""
    v = x[:, 0]
    w = np.sqrt(v[2] * v[3])
    h = v[2] / w
    return np.array([v[0] - w / 2.0, v[1] - h / 2.0, v[0] + w / 2.0, v[1] + h / 2.0])


def iou_matrix(a, b):
    """IoU of every box in a [n,>=4] with every box in b [m,>=4] -> [n,m] (facerec/sort.py:35-52)."""
    a = a[:, None, :]
    b = b[None, :, :]
    iw = np.maximum(0.0, np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]))
    ih = np.maximum(0.0, np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]))
    inter = iw * ih
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    return inter / (area_a + area_b - inter)


class KalmanBoxTracker:
    """One tracked face.  `history` holds (box, is_posterior) per frame since `first_frame`."""

    count = 0   # process-global tracker id, never reset (facerec/sort.py:87,122-123)

    def __init__(self, bbox, first_frame: int):
        self.first_frame = first_frame
        self.x = np.zeros((8, 1))
        self.x[:4] = box_to_z(bbox)
        self.P = _P0.copy()
        self.id = KalmanBoxTracker.count
        KalmanBoxTracker.count += 1
        self.history = [(self.get_state(), True)]
        self.time_since_update = 0
        self.hits = 1
        self.initial_hits = 1
        self.had_nan_preds = False

    def predict(self):
        # area and ratio must not go negative (facerec/sort.py:148-152)
        if (self.x[6] + self.x[2]) < 1e-3:
            self.x[6] *= 0.0
        if (self.x[7] + self.x[3]) < 1e-3:
            self.x[7] *= 0.0
        self.time_since_update += 1
        self.x = np.dot(_F, self.x)
        self.P = np.dot(np.dot(_F, self.P), _F.T) + _Q
        box = self.get_state()
        self.history.append((box, False))
        return box

    def update(self, bbox):
        z = box_to_z(bbox)
        y = z - np.dot(_H, self.x)
        PHT = np.dot(self.P, _H.T)
        S = np.dot(_H, PHT) + _R
        K = np.dot(PHT, np.linalg.inv(S))
        self.x = self.x + np.dot(K, y)
        I_KH = _I8 - np.dot(K, _H)
        self.P = np.dot(np.dot(I_KH, self.P), I_KH.T) + np.dot(np.dot(K, _R), K.T)
        self.history[-1] = (self.get_state(), True)   # the posterior replaces this frame's prediction
        self.time_since_update = 0
        self.hits += 1
        if len(self.history) == self.hits:
            self.initial_hits += 1

    def get_state(self):
        return state_to_box(self.x)

    def __len__(self):
        return len(self.history)


def associate(detections, trackers, iou_threshold):
    """Hungarian assignment on IoU (facerec/sort.py:172-201).  Returns (matches [k,2], unmatched detection ids)."""
    if len(trackers) == 0 or len(detections) == 0:
        return np.empty((0, 2), dtype=int), np.arange(len(detections))
    iou = iou_matrix(detections, trackers)
    iou[iou < iou_threshold] = -1.0
    rows, cols = linear_sum_assignment(iou, maximize=True)
    matches = [(r, c) for r, c in zip(rows, cols) if iou[r, c] >= iou_threshold]
    matches = np.array(matches, dtype=np.int32).reshape(-1, 2)
    # set-difference order, as the reference builds it
    unmatched = np.array(list(set(range(len(detections))) - set(matches[:, 0])))
    return matches, unmatched


class Sort:
    def __init__(self, max_age: int = 5, min_hits: int = 3, iou_threshold: float = 0.5):
        self.max_age = max_age
        self.min_hits = min_hits
        self.iou_threshold = iou_threshold
        self.trackers: List[KalmanBoxTracker] = []
        self.frame_count = 0
        self.detection_count = 0
        self.tracker_id_map: Dict[int, list] = {}        # tracker id -> detection ids
        self.detection_id_map: Dict[int, KalmanBoxTracker] = {}
        self.frame_map: Dict[int, int] = {}              # detection id -> frame

    def update(self, detections: np.ndarray, frame: int):
        """detections: [n,5] rows x1,y1,x2,y2,score.  Call for every frame.  Returns one global id per detection."""
        self.frame_count += 1
        predicted = np.zeros((len(self.trackers), 5))
        bad = []
        for t, trk in enumerate(self.trackers):
            pos = trk.predict()
            predicted[t, :4] = pos
            if np.any(np.isnan(pos)):
                bad.append(t)
        if bad:
            predicted = predicted[~np.isnan(predicted).any(axis=1)]
            for t in reversed(bad):
                print("WARNING: Removed tracker with NaN predictions.")
                self.trackers[t].had_nan_preds = True
                self.trackers.pop(t)
        assert detections.shape[1] == 5
        matched, unmatched = associate(detections, predicted, self.iou_threshold)

        ids = self.detection_count + np.arange(len(detections))
        self.detection_count += len(detections)
        for d in ids:
            self.frame_map[d] = frame
        for di, ti in matched:
            trk = self.trackers[ti]
            trk.update(detections[di])
            self.tracker_id_map[trk.id].append(ids[di])
            self.detection_id_map[ids[di]] = trk
        # stop following expired trackers and ones that missed a hit while starting (facerec/sort.py:277-282)
        for i in reversed(range(len(self.trackers))):
            trk = self.trackers[i]
            expired = trk.time_since_update > self.max_age and len(trk) >= self.min_hits
            not_started = len(trk) <= self.min_hits and trk.initial_hits < len(trk)
            if expired or not_started:
                self.trackers.pop(i)
        for di in unmatched:
            trk = KalmanBoxTracker(detections[di], frame)
            self.trackers.append(trk)
            self.tracker_id_map[trk.id] = [ids[di]]
            self.detection_id_map[ids[di]] = trk
        return ids

    def has_valid_tracker(self, detection_id) -> bool:
        trk = self.detection_id_map.get(detection_id)
        assert trk is not None, "Tried to access non-existent tracker <" + str(detection_id) + ">"
        return trk.initial_hits >= self.min_hits and not trk.had_nan_preds

    def get_detection_bbox(self, detection_id):
        trk = self.detection_id_map.get(detection_id)
        assert trk is not None, "Tried to access non-existent tracker!"
        i = self.frame_map.get(detection_id) - trk.first_frame
        assert 0 <= i < len(trk), "Faulty frame index!"
        return trk.history[i][0]

    def kill_trackers(self):
        self.trackers = []

    def pop_expired(self, expiry_age: int, current_frame: Optional[int] = None):
        """Forget trackers without detections for `expiry_age` frames; returns the valid ones with their
        trailing predictions trimmed (facerec/sort.py:338-378)."""
        if current_frame is None:
            current_frame = sys.maxsize
        out = []
        for trk_id in list(self.tracker_id_map.keys()):
            det_ids = self.tracker_id_map[trk_id]
            trk = self.detection_id_map[det_ids[0]]
            age = current_frame - (trk.first_frame + len(trk) - trk.time_since_update - 1)
            assert age >= 0, "Age less than zero?"
            if age >= expiry_age:
                del self.tracker_id_map[trk_id]
                for d in det_ids:
                    del self.detection_id_map[d]
                    del self.frame_map[d]
                if trk.initial_hits >= self.min_hits:
                    trk.history = trk.history[:len(trk) - trk.time_since_update]
                    out.append(trk)
        return out


class _ExpiredTrack:
    """What pop_expired hands to the trajectory writer: first_frame + history of (box, is_posterior)."""

    def __init__(self, first_frame, boxes, detected):
        self.first_frame = first_frame
        self.history = [(boxes[i], bool(detected[i])) for i in range(len(boxes))]

    def __len__(self):
        return len(self.history)


class NativeSort:
    """Same interface as `Sort`, state and arithmetic in libfacerec_b200.so (csrc/sort.cu): the per-frame
    host cost drops from ~1 ms of numpy calls to a few microseconds.  Ids, validity and integer boxes are identical to
    `Sort`; float boxes agree to ~1e-12 (LAPACK / BLAS vs. plain loops in the 4x4 inverse and 8x8 products).
    Caveat: on EXACT IoU ties (duplicate detection boxes in one frame) the shortest-augmenting-path assignment here and
    scipy's linear_sum_assignment may pick different, equally optimal matchings (35 of 300 fuzz runs with duplicated
    boxes, 0 of 300 without); NMS never emits duplicate boxes, so the extractor does not meet that case.
    Non-finite detections (or a 0/0 IoU of two degenerate boxes) raise ValueError, as scipy does inside the reference."""

    def __init__(self, max_age: int = 5, min_hits: int = 3, iou_threshold: float = 0.5):
        import ctypes as C
        from . import _lib
        self._C = C
        self.lib = _lib.load_library()
        self.max_age, self.min_hits, self.iou_threshold = max_age, min_hits, iou_threshold
        self.handle = C.c_void_p(self.lib.frb_sort_create(max_age, min_hits, float(iou_threshold)))
        if not self.handle.value:
            raise RuntimeError("frb_sort_create failed")
        self._nan_seen = 0

    def __del__(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.frb_sort_destroy(self.handle)
            self.handle = self._C.c_void_p()

    def update(self, detections: np.ndarray, frame: int):
        det = np.ascontiguousarray(detections, dtype=np.float64).reshape(-1, 5)
        ids = np.empty(len(det), np.int64)
        rc = self.lib.frb_sort_update(self.handle, det.ctypes.data, len(det), int(frame), ids.ctypes.data)
        if rc != 0:
            if not np.isfinite(det).all():
                raise ValueError("matrix contains invalid numeric entries")   # what scipy raises inside the reference Sort
            raise ValueError(f"frb_sort_update failed ({rc}): invalid IoU matrix")
        nan = self.lib.frb_sort_nan_removed(self.handle)
        for _ in range(nan - self._nan_seen):
            print("WARNING: Removed tracker with NaN predictions.")
        self._nan_seen = nan
        return ids

    def has_valid_tracker(self, detection_id) -> bool:
        r = self.lib.frb_sort_has_valid_tracker(self.handle, int(detection_id))
        assert r >= 0, "Tried to access non-existent tracker <" + str(detection_id) + ">"
        return bool(r)

    def get_detection_bbox(self, detection_id):
        box = np.empty(4, np.float64)
        rc = self.lib.frb_sort_get_detection_bbox(self.handle, int(detection_id), box.ctypes.data)
        assert rc == 0, "Tried to access non-existent tracker!"
        return box

    def kill_trackers(self):
        self.lib.frb_sort_kill_trackers(self.handle)

    def pop_expired(self, expiry_age: int, current_frame: Optional[int] = None):
        C = self._C
        n = self.lib.frb_sort_pop_expired(self.handle, int(expiry_age),
                                          int(current_frame) if current_frame is not None else 2 ** 63 - 1)
        out = []
        for i in range(n):
            first, ln = C.c_longlong(), C.c_int()
            self.lib.frb_sort_expired_info(self.handle, i, C.byref(first), C.byref(ln))
            boxes = np.empty((ln.value, 4), np.float64)
            det = np.empty(ln.value, np.uint8)
            self.lib.frb_sort_expired_boxes(self.handle, i, boxes.ctypes.data, det.ctypes.data)
            out.append(_ExpiredTrack(first.value, boxes, det))
        return out
