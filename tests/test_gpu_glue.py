"""Parity of the memory-bound glue kernels (K1, K6, K7, K8) against the numpy oracle: bit-exact for the
integer / index work, last-ulp (expf) tolerance for decoded boxes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_preprocess_matches_numpy(ctx):
    from facerec_b200 import ops
    rng = np.random.RandomState(0)
    for shape in [(2, 36, 50, 3), (1, 7, 9, 3), (3, 360, 640, 3)]:
        fr = rng.randint(0, 256, shape).astype(np.uint8)
        out = ops.preprocess_u8(ctx, torch.from_numpy(fr).cuda(), True).float().cpu().numpy()
        ref = np.zeros(shape[:3] + (8,), np.float32)
        ref[..., 0] = fr[..., 2].astype(np.float32) - 104
        ref[..., 1] = fr[..., 1].astype(np.float32) - 117
        ref[..., 2] = fr[..., 0].astype(np.float32) - 123
        assert np.array_equal(out, ref)
        out_bgr = ops.preprocess_u8(ctx, torch.from_numpy(fr).cuda(), False).float().cpu().numpy()
        assert np.array_equal(out_bgr[..., 0], fr[..., 0].astype(np.float32) - 104)


def _raw(rng, B, H, W, frac=0.01):
    from oracle import postproc
    P = postproc.num_priors(H, W)
    loc = rng.standard_normal((B, P, 4)).astype(np.float32)
    landm = rng.standard_normal((B, P, 10)).astype(np.float32) * 2
    conf = (0.5 * rng.standard_normal((B, P, 2))).astype(np.float32)
    hot = rng.rand(B, P) < frac
    conf[..., 1] += hot * rng.uniform(3, 12, (B, P)).astype(np.float32)
    return loc, conf, landm


@pytest.mark.parametrize("H,W,frac", [(360, 640, 0.01), (135, 97, 0.05), (1080, 1920, 0.001)])
def test_decode_filter_nms_matches_oracle(ctx, H, W, frac):
    from facerec_b200 import ops
    from oracle import postproc
    rng = np.random.RandomState(H + W)
    B = 3
    loc, conf, landm = _raw(rng, B, H, W, frac)
    assert ops.num_priors(H, W) == postproc.num_priors(H, W)
    tl, tc, tm = (torch.from_numpy(a).cuda() for a in (loc, conf, landm))
    cand, count, pidx = ops.prior_decode_filter(ctx, tl, tc, tm, H, W, 0.95, 1024)
    keep, nkeep = ops.nms(ctx, cand, count, pidx, 0.4, True)
    cand, count, pidx, keep, nkeep = (t.cpu().numpy() for t in (cand, count, pidx, keep, nkeep))
    for b in range(B):
        scores = postproc.face_scores(conf[b])
        sel = np.nonzero(scores > np.float32(0.95))[0]
        assert count[b] == len(sel) <= 1024
        got_sel = pidx[b, :count[b]]
        assert np.array_equal(np.sort(got_sel), sel)            # identical candidate set (order is arbitrary)
        dets, kept_priors = postproc.postprocess(loc[b], conf[b], landm[b], H, W)
        k = keep[b, :nkeep[b]]
        assert np.array_equal(pidx[b, k], kept_priors)           # identical kept set, identical order
        got = cand[b, k]
        np.testing.assert_allclose(got[:, 4], dets[:, 4], rtol=0, atol=2e-7)   # expf vs np.exp: <= 2 ulp
        np.testing.assert_allclose(got[:, :4], dets[:, :4], rtol=0, atol=2e-3)    # expf ulp * 1920 px
        np.testing.assert_allclose(got[:, 5:], dets[:, 5:], rtol=0, atol=1e-3)


def test_nms_empty_and_overflow(ctx):
    from facerec_b200 import ops
    rng = np.random.RandomState(3)
    H, W = 64, 64
    loc, conf, landm = _raw(rng, 2, H, W, 0.0)
    conf[0, :, 1] = -20          # frame 0: nothing passes
    conf[1, :, 1] = 20           # frame 1: everything passes -> overflow is reported through count
    tl, tc, tm = (torch.from_numpy(a).cuda() for a in (loc, conf, landm))
    cand, count, pidx = ops.prior_decode_filter(ctx, tl, tc, tm, H, W, 0.95, 32)
    keep, nkeep = ops.nms(ctx, cand, count, pidx)
    count, nkeep = count.cpu().numpy(), nkeep.cpu().numpy()
    assert count[0] == 0 and nkeep[0] == 0
    assert count[1] == loc.shape[1] and 1 <= nkeep[1] <= 32


def test_nms_crowded_frame_keeps_top_candidates_and_drops_nonfinite(ctx):
    """More than 1024 priors above the threshold (a crowd, or a miscalibrated class head): the 1024 highest-scoring
    candidates go through NMS (the published pipeline trims to its top 5000 the same way) and nothing raises; a
    non-finite box (exp overflow of a 16-bit regression) is never kept and suppresses nothing."""
    from facerec_b200 import ops
    from oracle import postproc
    rng = np.random.RandomState(11)
    H, W = 360, 640
    loc, conf, landm = _raw(rng, 2, H, W, 0.30)            # ~2 800 hot priors per frame
    loc[1, :, 2] = np.where(rng.rand(loc.shape[1]) < 0.05, 600.0, loc[1, :, 2])   # exp(120) = inf -> x2 = -inf + inf
    tl, tc, tm = (torch.from_numpy(a).cuda() for a in (loc, conf, landm))
    cand, count, pidx = ops.prior_decode_filter(ctx, tl, tc, tm, H, W, 0.95, 5000)
    keep, nkeep = ops.nms(ctx, cand, count, pidx, 0.4, True)
    cand, count, pidx, keep, nkeep = (t.cpu().numpy() for t in (cand, count, pidx, keep, nkeep))
    for b in range(2):
        scores = postproc.face_scores(conf[b])
        sel = np.nonzero(scores > np.float32(0.95))[0]
        assert 1024 < len(sel) == count[b] <= 5000
        assert np.array_equal(np.sort(pidx[b, :count[b]]), sel)
        # NMS in isolation: the oracle is fed the GPU's own decoded candidates (thousands of scores packed below 1.0
        # differ from numpy's by an ulp of expf here and there, which is decode parity, tested above, not NMS)
        c = cand[b, :count[b]]
        pr = pidx[b, :count[b]]
        order = np.lexsort((pr, -c[:, 4]))[:1024]            # the 1024 best: score descending, lower prior first
        finite = np.isfinite(c[order, :4]).all(1)
        assert finite.all() == (b == 0)
        sub = order[finite]
        k = postproc.nms(c[sub, :4], c[sub, 4], pr[sub], 0.4, True)
        got = keep[b, :nkeep[b]]
        assert np.array_equal(pr[got], pr[sub][k])
        assert np.isfinite(cand[b, got]).all()


def test_crop_resize_standardize_matches_cv2_semantics(ctx):
    from facerec_b200 import ops
    from oracle import crop
    rng = np.random.RandomState(7)
    B, H, W = 2, 90, 120
    frames = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    boxes = []
    for f in range(B):
        for _ in range(6):
            x, y = rng.randint(0, W - 4), rng.randint(0, H - 4)
            w, h = rng.randint(2, W - x + 1), rng.randint(2, H - y + 1)
            boxes.append((f, x, y, w, h))
    boxes.append((0, 0, 0, W, H))          # whole frame
    boxes.append((1, 5, 5, 1, 1))          # single pixel
    boxes.append((1, 10, 10, 160, 80)[:1] + (0, 0, 120, 80))
    bx = np.asarray(boxes, np.int32)
    out = ops.crop_resize_std(ctx, torch.from_numpy(frames).cuda(), torch.from_numpy(bx).cuda()).float().cpu().numpy()
    for i, (f, x, y, w, h) in enumerate(bx):
        ref = crop.standardize(crop.resize_linear_u8(frames[f, y:y + h, x:x + w]))
        assert np.array_equal(out[i, :, :, :3], ref), i       # (u8 - 127.5)/128 is exact in fp16
        assert not out[i, :, :, 3:].any()


@pytest.mark.parametrize("B,H,W", [(2, 18, 26), (1, 16, 24), (2, 17, 25), (1, 20, 40), (1, 7, 4)])
def test_preprocess_space_to_depth_layout(ctx, B, H, W):
    """W * 3 a multiple of 4 takes the 32-bit-load path, the others the byte path; odd sizes leave the missing row /
    column of the last 2x2 block zero."""
    from facerec_b200 import ops, _lib
    rng = np.random.RandomState(4)
    fr = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    out = ops.preprocess_u8(ctx, torch.from_numpy(fr).cuda(), True, _lib.FRB_LAYOUT_S2D16).float().cpu().numpy()
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    assert out.shape == (B, H2, W2 + 3, 16)
    bgr = np.zeros((B, 2 * H2, 2 * W2, 4), np.float32)
    bgr[:, :H, :W, :3] = fr[..., ::-1].astype(np.float32) - np.array([104, 117, 123], np.float32)
    ref = np.zeros_like(out)
    for dy in range(2):
        for dx in range(2):
            ref[:, :, 2:-1, (dy * 2 + dx) * 4:(dy * 2 + dx) * 4 + 4] = bgr[:, dy::2, dx::2, :]
    assert np.array_equal(out, ref)


def test_scene_stats_match_host_detector(ctx):
    """GPU scene-cut statistics vs the host restatement (itself pinned to facerec/scene.py by tests/golden/scene_cuts.json)."""
    from facerec_b200 import synth, scene
    H, W = 180, 320
    frames = np.stack([synth.frame(77, i, H, W, vary_level=True, scene_len=7) for i in range(60)])
    host = scene.SceneChangeDetector(False, True, 1)
    ref, ref_stats = [], []
    for f in frames:
        ref.append(bool(host.update(f)))
        ref_stats.append((host.prev_mafd_eq or 0.0, host.prev_fv_eq or 0.0))
    gpu = scene.B200SceneChangeDetector(False, True, 1, ctx=ctx)
    got, got_stats = [], []
    for lo, hi in ((0, 1), (1, 17), (17, 18), (18, 60)):          # ragged batches, state carried across calls
        got += gpu.update_batch(torch.from_numpy(frames[lo:hi]).cuda())
        got_stats.append((gpu.prev_mafd_eq or 0.0, gpu.prev_fv_eq or 0.0))
    assert got == ref and sum(ref) >= 3
    for (lo, hi), g in zip(((0, 1), (1, 17), (17, 18), (18, 60)), got_stats):
        r = ref_stats[hi - 1]
        assert abs(g[0] - r[0]) <= 1e-9 * max(1.0, abs(r[0])) and abs(g[1] - r[1]) <= 1e-9 * max(1.0, abs(r[1]))
    # no crop needed for a 2:1 frame, and the wide-frame path takes all rows
    wide = np.stack([synth.frame(5, i, 100, 200) for i in range(3)])
    h2 = scene.SceneChangeDetector(False, True, 1)
    g2 = scene.B200SceneChangeDetector(False, True, 1, ctx=ctx)
    assert g2.update_batch(torch.from_numpy(wide).cuda()) == [bool(h2.update(f)) for f in wide]
    assert abs(g2.prev_mafd_eq - h2.prev_mafd_eq) < 1e-9
