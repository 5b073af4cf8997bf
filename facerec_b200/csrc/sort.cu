// Host-side SORT tracker: the reference's modified SORT (facerec/sort.py:82-378) behind the C ABI.
// Sequential and tiny (an 8-state Kalman filter per face, Hungarian assignment over tens of boxes) — host code, no
// kernels; it lives in this library so the extractor's per-frame loop does not pay ~1 ms of Python + numpy per frame.
// Same arithmetic as facerec_b200/sort.py (which is bit-identical to the reference): the products with the 0/1
// matrices F and H are written out as the exact sums they are; the 4x4 inverse and the dense 8x8 products may differ
// from LAPACK / BLAS in the last bit, integer boxes, ids and validity do not (tests/test_host_golden.py).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>

#include "../../include/facerec_b200.h"

namespace {

struct Tracker {
  long long id = 0, first_frame = 0;
  double x[8];
  double P[8][8];
  std::vector<std::array<double, 4>> boxes;   // one per frame since first_frame
  std::vector<uint8_t> posterior;             // 1: Kalman posterior of a matched detection, 0: prediction only
  int time_since_update = 0, hits = 1, initial_hits = 1;
  bool had_nan = false;

  void state_box(double* b) const {           // facerec/sort.py:70-79
    const double w = std::sqrt(x[2] * x[3]);
    const double h = x[2] / w;
    b[0] = x[0] - w / 2.0; b[1] = x[1] - h / 2.0; b[2] = x[0] + w / 2.0; b[3] = x[1] + h / 2.0;
  }
  void push(bool post) {
    std::array<double, 4> b;
    state_box(b.data());
    boxes.push_back(b);
    posterior.push_back(post ? 1 : 0);
  }
};

void box_to_z(const double* b, double* z) {   // facerec/sort.py:55-67
  const double w = b[2] - b[0], h = b[3] - b[1];
  z[0] = b[0] + w / 2.0; z[1] = b[1] + h / 2.0; z[2] = w * h; z[3] = w / h;
}

const double kQ[8] = {1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.01, 0.0001};
const double kR[4] = {1.0, 1.0, 10.0, 10.0};
const double kP0[8] = {10.0, 10.0, 10.0, 10.0, 10000.0, 10000.0, 10000.0, 10000.0};

void predict(Tracker& t) {                    // facerec/sort.py:143-161
  if (t.x[6] + t.x[2] < 1e-3) t.x[6] *= 0.0;
  if (t.x[7] + t.x[3] < 1e-3) t.x[7] *= 0.0;
  t.time_since_update += 1;
  for (int i = 0; i < 4; ++i) t.x[i] = t.x[i] + t.x[i + 4];                     // F x
  double FP[8][8];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) FP[i][j] = i < 4 ? t.P[i][j] + t.P[i + 4][j] : t.P[i][j];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) t.P[i][j] = (j < 4 ? FP[i][j] + FP[i][j + 4] : FP[i][j]) + (i == j ? kQ[i] : 0.0);
  t.push(false);
}

// inverse of a 4x4 by LU with partial pivoting (what numpy.linalg.inv does through gesv)
void inv4(const double S[4][4], double out[4][4]) {
  double a[4][4], b[4][4];
  memcpy(a, S, sizeof(a));
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) b[i][j] = i == j ? 1.0 : 0.0;
  for (int k = 0; k < 4; ++k) {
    int p = k;
    for (int i = k + 1; i < 4; ++i)
      if (std::fabs(a[i][k]) > std::fabs(a[p][k])) p = i;
    if (p != k) {
      for (int j = 0; j < 4; ++j) { std::swap(a[k][j], a[p][j]); std::swap(b[k][j], b[p][j]); }
    }
    const double r = 1.0 / a[k][k];
    for (int i = k + 1; i < 4; ++i) {
      a[i][k] *= r;
      for (int j = k + 1; j < 4; ++j) a[i][j] -= a[i][k] * a[k][j];
    }
  }
  for (int c = 0; c < 4; ++c) {                // L y = P e_c, then U x = y
    for (int i = 1; i < 4; ++i)
      for (int k = 0; k < i; ++k) b[i][c] -= a[i][k] * b[k][c];
    for (int i = 3; i >= 0; --i) {
      for (int k = i + 1; k < 4; ++k) b[i][c] -= a[i][k] * b[k][c];
      b[i][c] /= a[i][i];
    }
  }
  memcpy(out, b, sizeof(b));
}

void update(Tracker& t, const double* bbox) { // facerec/sort.py:125-141 + filterpy KalmanFilter.update
  double z[4], y[4];
  box_to_z(bbox, z);
  for (int i = 0; i < 4; ++i) y[i] = z[i] - t.x[i];                              // z - H x
  double S[4][4], Si[4][4], K[8][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) S[i][j] = t.P[i][j] + (i == j ? kR[i] : 0.0);   // H P H' + R
  inv4(S, Si);
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += t.P[i][k] * Si[k][j];                     // (P H') S^-1
      K[i][j] = s;
    }
  for (int i = 0; i < 8; ++i) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s += K[i][k] * y[k];
    t.x[i] = t.x[i] + s;
  }
  double A[8][8], AP[8][8], Pn[8][8];           // A = I - K H; P = A P A' + K R K'
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) A[i][j] = (i == j ? 1.0 : 0.0) - (j < 4 ? K[i][j] : 0.0);
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double s = 0.0;
      for (int k = 0; k < 8; ++k) s += A[i][k] * t.P[k][j];
      AP[i][j] = s;
    }
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double s = 0.0;
      for (int k = 0; k < 8; ++k) s += AP[i][k] * A[j][k];
      double kr = 0.0;
      for (int k = 0; k < 4; ++k) kr += K[i][k] * kR[k] * K[j][k];
      Pn[i][j] = s + kr;
    }
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) t.P[i][j] = Pn[i][j];
  std::array<double, 4> b;
  t.state_box(b.data());
  t.boxes.back() = b;                          // the posterior replaces this frame's prediction
  t.posterior.back() = 1;
  t.time_since_update = 0;
  t.hits += 1;
  if (static_cast<int>(t.boxes.size()) == t.hits) t.initial_hits += 1;
}

// Rectangular assignment maximising the total of `w` [n x m] (shortest augmenting paths with potentials, rows <= cols
// after an optional transpose): returns for every row the matched column or -1.
std::vector<int> assign_max(const std::vector<double>& w, int n, int m) {
  const bool tr = n > m;
  const int R = tr ? m : n, C = tr ? n : m;
  auto cost = [&](int r, int c) { return tr ? -w[static_cast<size_t>(c) * m + r] : -w[static_cast<size_t>(r) * m + c]; };
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(R + 1, 0.0), v(C + 1, 0.0);
  std::vector<int> p(C + 1, 0), way(C + 1, 0);
  for (int i = 1; i <= R; ++i) {
    p[0] = i;
    int j0 = 0;
    std::vector<double> minv(C + 1, INF);
    std::vector<char> used(C + 1, 0);
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = INF;
      int j1 = 0;
      for (int j = 1; j <= C; ++j)
        if (!used[j]) {
          const double cur = cost(i0 - 1, j - 1) - u[i0] - v[j];
          if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
          if (minv[j] < delta) { delta = minv[j]; j1 = j; }
        }
      if (!(delta < INF)) return std::vector<int>();   // non-finite weights: no augmenting path can be ranked (caller reports it)
      for (int j = 0; j <= C; ++j)
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0);
  }
  std::vector<int> row_to_col(n, -1);
  for (int j = 1; j <= C; ++j)
    if (p[j] != 0) {
      if (tr) row_to_col[j - 1] = p[j] - 1;
      else row_to_col[p[j] - 1] = j - 1;
    }
  return row_to_col;
}

}  // namespace

struct frb_sort {
  int max_age, min_hits;
  double iou_threshold;
  long long next_tracker_id = 0, detection_count = 0;
  std::vector<std::shared_ptr<Tracker>> trackers;
  std::map<long long, std::vector<long long>> tracker_dets;               // tracker id -> detection ids (creation order)
  std::unordered_map<long long, std::shared_ptr<Tracker>> det_tracker;    // detection id -> tracker
  std::unordered_map<long long, long long> det_frame;                     // detection id -> frame
  std::vector<std::shared_ptr<Tracker>> expired;                          // result of the last pop_expired
  int nan_removed = 0;
};

extern "C" {

frb_sort* frb_sort_create(int max_age, int min_hits, double iou_threshold) {
  frb_sort* s = new frb_sort();
  s->max_age = max_age; s->min_hits = min_hits; s->iou_threshold = iou_threshold;
  return s;
}

void frb_sort_destroy(frb_sort* s) { delete s; }

int frb_sort_update(frb_sort* s, const double* dets, int n, long long frame, long long* out_ids) {
  if (s == nullptr || n < 0 || (n > 0 && (dets == nullptr || out_ids == nullptr))) return FRB_ERR_INVALID;
  // A non-finite detection would put NaN into the IoU matrix; the reference's linear_sum_assignment raises ValueError on
  // that input (facerec/sort.py:31 via scipy).  Refuse it before any tracker state is touched.
  for (int i = 0; i < 5 * n; ++i)
    if (!std::isfinite(dets[i])) return FRB_ERR_INVALID;
  // predictions of the live trackers; NaN ones are dropped for good (facerec/sort.py:228-242)
  std::vector<std::shared_ptr<Tracker>> live;
  std::vector<std::array<double, 4>> pred;
  for (auto& t : s->trackers) {
    predict(*t);
    const std::array<double, 4>& b = t->boxes.back();
    if (std::isnan(b[0]) || std::isnan(b[1]) || std::isnan(b[2]) || std::isnan(b[3])) {
      t->had_nan = true;
      s->nan_removed++;
      continue;
    }
    live.push_back(t);
    pred.push_back(b);
  }
  s->trackers.swap(live);
  const int m = static_cast<int>(pred.size());
  std::vector<int> match(n, -1);
  if (n > 0 && m > 0) {
    std::vector<double> iou(static_cast<size_t>(n) * m);
    for (int i = 0; i < n; ++i) {
      const double* a = dets + 5 * i;
      for (int j = 0; j < m; ++j) {
        const std::array<double, 4>& b = pred[j];
        const double iw = std::max(0.0, std::min(a[2], b[2]) - std::max(a[0], b[0]));
        const double ih = std::max(0.0, std::min(a[3], b[3]) - std::max(a[1], b[1]));
        const double inter = iw * ih;
        double v = inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter);
        if (v < s->iou_threshold) v = -1.0;
        if (std::isnan(v)) return FRB_ERR_INVALID;     // 0 / 0 of two degenerate boxes: scipy raises on the same matrix
        iou[static_cast<size_t>(i) * m + j] = v;
      }
    }
    std::vector<int> r2c = assign_max(iou, n, m);
    if (r2c.empty()) return FRB_ERR_INVALID;
    for (int i = 0; i < n; ++i)
      if (r2c[i] >= 0 && iou[static_cast<size_t>(i) * m + r2c[i]] >= s->iou_threshold) match[i] = r2c[i];
  }
  for (int i = 0; i < n; ++i) {
    out_ids[i] = s->detection_count + i;
    s->det_frame[out_ids[i]] = frame;
  }
  s->detection_count += n;
  for (int i = 0; i < n; ++i)
    if (match[i] >= 0) {
      std::shared_ptr<Tracker>& t = s->trackers[match[i]];
      update(*t, dets + 5 * i);
      s->tracker_dets[t->id].push_back(out_ids[i]);
      s->det_tracker[out_ids[i]] = t;
    }
  // stop following expired trackers and ones that missed a hit while starting (facerec/sort.py:277-282)
  for (int i = static_cast<int>(s->trackers.size()) - 1; i >= 0; --i) {
    const Tracker& t = *s->trackers[i];
    const int len = static_cast<int>(t.boxes.size());
    const bool expired = t.time_since_update > s->max_age && len >= s->min_hits;
    const bool not_started = len <= s->min_hits && t.initial_hits < len;
    if (expired || not_started) s->trackers.erase(s->trackers.begin() + i);
  }
  for (int i = 0; i < n; ++i)
    if (match[i] < 0) {
      auto t = std::make_shared<Tracker>();
      t->id = s->next_tracker_id++;
      t->first_frame = frame;
      memset(t->x, 0, sizeof(t->x));
      box_to_z(dets + 5 * i, t->x);
      for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b) t->P[a][b] = a == b ? kP0[a] : 0.0;
      t->push(true);
      s->trackers.push_back(t);
      s->tracker_dets[t->id] = {out_ids[i]};
      s->det_tracker[out_ids[i]] = t;
    }
  return FRB_OK;
}

int frb_sort_has_valid_tracker(frb_sort* s, long long det_id) {
  if (s == nullptr) return -1;
  auto it = s->det_tracker.find(det_id);
  if (it == s->det_tracker.end()) return -1;
  return it->second->initial_hits >= s->min_hits && !it->second->had_nan ? 1 : 0;
}

int frb_sort_get_detection_bbox(frb_sort* s, long long det_id, double* out4) {
  if (s == nullptr || out4 == nullptr) return FRB_ERR_INVALID;
  auto it = s->det_tracker.find(det_id);
  auto fr = s->det_frame.find(det_id);
  if (it == s->det_tracker.end() || fr == s->det_frame.end()) return FRB_ERR_INVALID;
  const long long i = fr->second - it->second->first_frame;
  if (i < 0 || i >= static_cast<long long>(it->second->boxes.size())) return FRB_ERR_INVALID;
  memcpy(out4, it->second->boxes[i].data(), 4 * sizeof(double));
  return FRB_OK;
}

void frb_sort_kill_trackers(frb_sort* s) {
  if (s) s->trackers.clear();
}

int frb_sort_nan_removed(frb_sort* s) { return s ? s->nan_removed : 0; }

int frb_sort_pop_expired(frb_sort* s, long long expiry_age, long long current_frame) {
  if (s == nullptr) return -1;
  s->expired.clear();
  for (auto it = s->tracker_dets.begin(); it != s->tracker_dets.end();) {
    std::shared_ptr<Tracker> t = s->det_tracker[it->second[0]];
    const long long last = t->first_frame + static_cast<long long>(t->boxes.size()) - t->time_since_update - 1;
    if (current_frame - last >= expiry_age) {
      for (long long d : it->second) {
        s->det_tracker.erase(d);
        s->det_frame.erase(d);
      }
      it = s->tracker_dets.erase(it);
      if (t->initial_hits >= s->min_hits) {
        const size_t keep = t->boxes.size() - t->time_since_update;   // trim the trailing predictions
        t->boxes.resize(keep);
        t->posterior.resize(keep);
        s->expired.push_back(t);
      }
    } else {
      ++it;
    }
  }
  return static_cast<int>(s->expired.size());
}

int frb_sort_expired_info(frb_sort* s, int i, long long* first_frame, int* len) {
  if (s == nullptr || i < 0 || i >= static_cast<int>(s->expired.size())) return FRB_ERR_INVALID;
  if (first_frame) *first_frame = s->expired[i]->first_frame;
  if (len) *len = static_cast<int>(s->expired[i]->boxes.size());
  return FRB_OK;
}

int frb_sort_expired_boxes(frb_sort* s, int i, double* boxes, unsigned char* detected) {
  if (s == nullptr || i < 0 || i >= static_cast<int>(s->expired.size()) || boxes == nullptr || detected == nullptr)
    return FRB_ERR_INVALID;
  const Tracker& t = *s->expired[i];
  for (size_t k = 0; k < t.boxes.size(); ++k) {
    memcpy(boxes + 4 * k, t.boxes[k].data(), 4 * sizeof(double));
    detected[k] = t.posterior[k];
  }
  return FRB_OK;
}

}  // extern "C"
