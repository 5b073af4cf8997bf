"""End-to-end parity of the CUDA path against the fp32 oracle on the synthetic clip (north_star tolerances:
box IoU >= 0.99 with identical kept-face sets, embedding cosine >= 0.999).  Writes one JSON record.

  python tools/parity_report.py --backbone mnet025 --height 360 --width 640 --frames 240
  python tools/parity_report.py --backbone r50 --height 1080 --width 1920 --frames 64 --stride 1

What is compared, per frame:
  * the PRODUCT output (FacePipeline.detect_embed: kept faces after NMS and the min-size filter, crop rectangles,
    embeddings) with the oracle's kept faces (oracle/nets.py -> oracle/postproc.py -> oracle/crop.py);
  * a face of the oracle is "reproduced" when a product face has IoU >= 0.99 with it (north_star).

Every face that is NOT reproduced is then explained or not.  The detector makes three discrete decisions per prior
(score > 0.95, min(w, h) >= min_size, greedy NMS at IoU 0.4); a 16-bit network cannot agree with an fp32 one on a
decision whose margin is smaller than its own rounding noise, so a face may legitimately differ when a decision on
ITS path sits inside a declared band:
    score band   |score - 0.95|          < 1e-3   (of the oracle's or the GPU's score of that prior)
    size band    |min(w, h) - min_size|   < 0.25 px
    nms band     some other candidate q decides its fate: nms_iou = |IoU(p, q) - 0.4| < 0.01 on either side;
                 nms_order = the scores of p and q are within 1e-3 of each other (the greedy order of the two flips);
                 nms_cascade = q itself is kept on one side only (its own difference is explained where q is judged) or
                 q's score sits in the score band
    iou band     the SAME prior is kept on both sides but the two boxes have IoU in [0.97, 0.99) (sub-pixel wobble
                 of a small box; reported separately, never counted as identical)
`frames_identical`: kept sets equal, every face IoU >= 0.99.  `frames_identical_outside_band`: every difference of the
frame is explained by a band.  `unexplained` lists the rest with everything needed to reproduce them.
The decisions are replayed from the GPU's RAW head outputs (frb_retina_forward) with the oracle's post-processing,
which frb_prior_decode_filter / frb_nms reproduce bit for bit (tests/test_gpu_glue.py), so prior ids are available.

Embedding cosine: `cos_min_same_rect` compares faces whose integer crop rectangle is the same on both sides (the
embedder alone); `cos_min_all` is unconditional: GPU embedding of the GPU's crop vs oracle embedding of the oracle's
crop, for every reproduced face.  The crop rectangle is np.round() of the float box (facerec/extract.py:38-53): a
coordinate within the box noise of a rounding boundary (k + 0.5) lands on the other integer and moves the crop by one
pixel; `rect_differs_max_dist_to_half` is the largest distance to such a boundary among the faces whose rectangles
differ (it must stay within the box noise: `rect_differs_unexplained` counts those further than 0.06 px).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerec_b200 import synth, weights as W, _lib, ops  # noqa: E402
from facerec_b200.pipeline import FacePipeline, boxes_to_rects  # noqa: E402
from oracle import nets, postproc, crop  # noqa: E402

SCORE_THR = 0.95
SCORE_BAND = 1e-3
SIZE_BAND = 0.25
NMS_IOU = 0.4
NMS_BAND = 0.01
IOU_OK = 0.99
IOU_WOBBLE = 0.97


def iou(a, b, po=0.0):
    iw = max(0.0, min(a[2], b[2]) - max(a[0], b[0]) + po)
    ih = max(0.0, min(a[3], b[3]) - max(a[1], b[1]) + po)
    inter = iw * ih
    return inter / (((a[2] - a[0]) + po) * ((a[3] - a[1]) + po) + ((b[2] - b[0]) + po) * ((b[3] - b[1]) + po) - inter)


def _side(loc, conf, landm, H, Wd, min_size):
    """Replay of the post-processing on one side's raw outputs -> dict prior id -> record."""
    pri = postproc.priorbox(H, Wd)
    scores = postproc.face_scores(conf)
    wide = np.nonzero(scores > np.float32(SCORE_THR - 4 * SCORE_BAND))[0]      # candidates incl. the band below the threshold
    boxes, _ = postproc.decode(loc[wide], landm[wide], pri[wide], H, Wd)
    dets, kept_ids = postproc.postprocess(loc, conf, landm, H, Wd)
    size_ok = np.minimum(dets[:, 2] - dets[:, 0], dets[:, 3] - dets[:, 1]) >= min_size if len(dets) else np.zeros(0, bool)
    return {"scores": scores, "cand": {int(p): (boxes[i], float(scores[p])) for i, p in enumerate(wide)},
            "kept": {int(p): dets[i] for i, p in enumerate(kept_ids)},
            "final": {int(p): dets[i] for i, p in enumerate(kept_ids) if size_ok[i]}}


def _explain(p, me, other, min_size):
    """Why may prior p be a final face on side `me` and not on side `other`?  Returns a list of band names."""
    why = []
    s_me, s_ot = float(me["scores"][p]), float(other["scores"][p])
    if abs(s_me - SCORE_THR) < SCORE_BAND or abs(s_ot - SCORE_THR) < SCORE_BAND:
        why.append("score")
    for side in (me, other):
        if p in side["cand"]:
            b = side["cand"][p][0]
            if abs(float(min(b[2] - b[0], b[3] - b[1])) - min_size) < SIZE_BAND:
                why.append("size")
                break
    # NMS: p is kept on `me`; on `other` it passed the threshold but was suppressed (or its suppressor differs)
    if p in other["cand"] and p not in other["kept"]:
        bp, sp = other["cand"][p]
        for q, (bq, sq) in other["cand"].items():
            if q == p:
                continue
            v = iou(bp, bq, 1.0)
            if v <= NMS_IOU - NMS_BAND:
                continue
            near_thr = abs(v - NMS_IOU) < NMS_BAND
            flip = abs(sq - sp) < SCORE_BAND or (q in me["cand"] and abs(me["cand"][q][1] - me["cand"][p][1]) < SCORE_BAND)
            # q decides p's fate on one side and not on the other: its own presence may be borderline too
            q_border = abs(float(me["scores"][q]) - SCORE_THR) < SCORE_BAND or abs(float(other["scores"][q]) - SCORE_THR) < SCORE_BAND
            vm = iou(me["cand"][p][0], me["cand"][q][0], 1.0) if (q in me["cand"] and p in me["cand"]) else v
            if near_thr or abs(vm - NMS_IOU) < NMS_BAND:
                why.append("nms_iou")
                break
            if flip:
                why.append("nms_order")
                break
            if q_border or ((q in other["kept"]) != (q in me["kept"])):
                why.append("nms_cascade")
                break
    return why


def run(backbone, H, Wd, n_frames, first=0, stride=7, seed=1234, min_size=20, batch=16, verbose=False):
    ctx = _lib.default_context(0)
    pipe = FacePipeline(backbone, min_face_size=min_size, save_every=1, ctx=ctx)
    idx = [first + i * stride for i in range(n_frames)]
    Pd = nets.params_from_weights(W.synth_weights("retina_" + backbone, 0))
    Pe = nets.params_from_weights(W.synth_weights("irv1", 0))
    st = {"frames": n_frames, "frames_identical": 0, "frames_identical_outside_band": 0, "oracle_faces": 0, "gpu_faces": 0,
          "matched_iou99": 0, "min_iou_matched": 1.0, "max_landmark_err_px": 0.0,
          "cos_min_same_rect": 1.0, "cos_n_same_rect": 0, "cos_min_all": 1.0, "cos_n_all": 0, "cos_below_0999_all": 0,
          "rect_differs": 0, "rect_differs_max_dist_to_half": 0.0, "rect_differs_unexplained": 0,
          "explained": {"score": 0, "size": 0, "nms_iou": 0, "nms_order": 0, "nms_cascade": 0, "iou_wobble": 0}, "unexplained": [],
          "max_abs_logit_err": 0.0, "max_abs_score_err": 0.0}
    mid = pipe.detector.model_id
    with torch.no_grad():
        for c0 in range(0, n_frames, batch):
            ids = idx[c0:c0 + batch]
            frames = np.stack([synth.frame(seed, i, H, Wd) for i in ids])
            res = pipe.detect_embed(frames)
            dev = torch.from_numpy(frames).to(pipe.device)
            x16 = ops.preprocess_u8(ctx, dev, True, ops.input_layout(mid))
            g_loc, g_conf, g_landm = (t.cpu().numpy() for t in ops.retina_forward(ctx, mid, x16, H, Wd))
            del x16, dev
            for b in range(len(ids)):
                x = torch.from_numpy(frames[b][..., ::-1].astype(np.float32) - np.array([104, 117, 123], np.float32))
                loc, conf, landm = nets.retina_forward(Pd, x.permute(2, 0, 1)[None].contiguous(), backbone)
                o = _side(loc[0].numpy(), conf[0].numpy(), landm[0].numpy(), H, Wd, min_size)
                g = _side(g_loc[b], g_conf[b], g_landm[b], H, Wd, min_size)
                st["max_abs_logit_err"] = max(st["max_abs_logit_err"], float(np.abs(g_conf[b] - conf[0].numpy()).max()))
                st["max_abs_score_err"] = max(st["max_abs_score_err"], float(np.abs(g["scores"] - o["scores"]).max()))
                dets = np.stack(list(o["final"].values())) if o["final"] else np.zeros((0, 15), np.float32)
                got = res.dets[b]
                st["oracle_faces"] += len(dets)
                st["gpu_faces"] += len(got)
                # ---- product output vs oracle: IoU matching (north_star's criterion)
                used = set()
                identical = len(dets) == len(got)
                for d in dets:
                    best, bj = 0.0, -1
                    for j, gg in enumerate(got):
                        if j not in used:
                            v = iou(d, gg)
                            if v > best:
                                best, bj = v, j
                    if best < IOU_OK:
                        identical = False
                        continue
                    used.add(bj)
                    st["matched_iou99"] += 1
                    st["min_iou_matched"] = min(st["min_iou_matched"], best)
                    st["max_landmark_err_px"] = max(st["max_landmark_err_px"], float(np.abs(got[bj][5:] - d[5:]).max()))
                    rect = boxes_to_rects(d[None, :4], np.zeros(1, np.int64), Wd, H)[0]
                    k = np.nonzero((res.face_frame == b) & (res.face_det == bj))[0]
                    if len(k) and rect[3] > 0 and rect[4] > 0:
                        _, x0, y0, w, h = rect
                        face = crop.standardize(crop.resize_linear_u8(frames[b][y0:y0 + h, x0:x0 + w]))
                        ref = nets.irv1_forward(Pe, torch.from_numpy(face).permute(2, 0, 1)[None].contiguous())[0].numpy()
                        cs = float((ref * res.embeddings[k[0]]).sum())
                        st["cos_min_all"] = min(st["cos_min_all"], cs)
                        st["cos_n_all"] += 1
                        st["cos_below_0999_all"] += int(cs < 0.999)
                        if tuple(res.rects[k[0]][1:]) == tuple(rect[1:]):
                            st["cos_min_same_rect"] = min(st["cos_min_same_rect"], cs)
                            st["cos_n_same_rect"] += 1
                        else:
                            st["rect_differs"] += 1
                            # which coordinate rounded the other way, and how close to the boundary was it?
                            tb = np.minimum(np.maximum(d[:4], 0), [Wd, H, Wd, H])
                            tg = np.minimum(np.maximum(got[bj][:4], 0), [Wd, H, Wd, H])
                            moved = np.round(tb) != np.round(tg)
                            dist_half = np.abs((tb - np.floor(tb)) - 0.5)
                            dd = float(dist_half[moved].max()) if moved.any() else 1.0
                            st["rect_differs_max_dist_to_half"] = max(st["rect_differs_max_dist_to_half"], dd)
                            st["rect_differs_unexplained"] += int(dd > 0.06)
                if len(got) != len(used):
                    identical = False
                st["frames_identical"] += int(identical)
                # ---- explain every difference of the frame from the replayed decisions
                ok_band = True
                if not identical:
                    of, gf = set(o["final"]), set(g["final"])
                    for p in sorted(of ^ gf):
                        me, other = (o, g) if p in of else (g, o)
                        why = _explain(p, me, other, min_size)
                        if why:
                            st["explained"][why[0]] += 1
                        else:
                            ok_band = False
                            bo = o["cand"].get(p, (None, None))[0]
                            bg = g["cand"].get(p, (None, None))[0]
                            st["unexplained"].append({"frame": ids[b], "prior": p, "side": "oracle_only" if p in of else "gpu_only",
                                                      "score_oracle": float(o["scores"][p]), "score_gpu": float(g["scores"][p]),
                                                      "box_oracle": None if bo is None else [float(v) for v in bo],
                                                      "box_gpu": None if bg is None else [float(v) for v in bg],
                                                      "kept_oracle": p in o["kept"], "kept_gpu": p in g["kept"]})
                    for p in sorted(of & gf):
                        v = iou(o["final"][p], g["final"][p])
                        if v < IOU_OK:
                            if v >= IOU_WOBBLE:
                                st["explained"]["iou_wobble"] += 1
                            else:
                                ok_band = False
                                st["unexplained"].append({"frame": ids[b], "prior": p, "side": "both", "iou": v,
                                                          "box_oracle": [float(x_) for x_ in o["final"][p][:4]],
                                                          "box_gpu": [float(x_) for x_ in g["final"][p][:4]]})
                st["frames_identical_outside_band"] += int(identical or ok_band)
                if verbose and not identical:
                    print(f"frame {ids[b]}: oracle {len(dets)} gpu {len(got)} band_ok {ok_band}", file=sys.stderr)
    st["face_match_rate"] = st["matched_iou99"] / max(st["oracle_faces"], 1)
    st["frac_identical"] = st["frames_identical"] / n_frames
    st["frac_identical_outside_band"] = st["frames_identical_outside_band"] / n_frames
    st["cos_frac_ge_0999_all"] = 1.0 - st["cos_below_0999_all"] / max(st["cos_n_all"], 1)
    st["cos_min"] = st["cos_min_all"]
    st["cos_n"] = st["cos_n_all"]
    st.update({"backbone": backbone, "H": H, "W": Wd, "first": first, "stride": stride,
               "bands": {"score": SCORE_BAND, "size_px": SIZE_BAND, "nms_iou": NMS_BAND, "iou_wobble": [IOU_WOBBLE, IOU_OK]},
               "dtype": "fp16 activations, fp32 accumulate"})
    st["unexplained"] = st["unexplained"][:40]
    return st


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="mnet025")
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--stride", type=int, default=7)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    out = run(a.backbone, a.height, a.width, a.frames, first=a.first, stride=a.stride, batch=a.batch, verbose=a.v)
    print(json.dumps({k: (v.item() if hasattr(v, "item") else v) for k, v in out.items()}))
