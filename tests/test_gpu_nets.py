"""Whole-network parity: the CUDA graphs (fused layers, concat-by-offset buffers, fp16 activations,
fp32 accumulation) against the fp32 PyTorch oracle on the same seeded weights and inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _frames(rng, B, H, W):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    out = np.empty((B, H, W, 3), np.uint8)
    for b in range(B):
        img = np.zeros((H, W, 3), np.float32)
        for c in range(3):
            img[..., c] = 127 + 60 * np.sin(xx / rng.uniform(7, 40) + rng.uniform(0, 6)) * np.cos(yy / rng.uniform(7, 40))
        img += rng.standard_normal((H, W, 3)) * 12
        out[b] = np.clip(img, 0, 255).astype(np.uint8)
    return out


@pytest.mark.parametrize("model,backbone,B,H,W", [("retina_mnet025", "mnet025", 2, 360, 640),
                                                  ("retina_r50", "r50", 2, 135, 240),
                                                  ("retina_r50", "r50", 1, 270, 481)])
def test_retina_forward_matches_oracle(ctx, model, backbone, B, H, W):
    from facerec_b200 import ops, weights
    from oracle import nets, postproc
    w = weights.synth_weights(model, 0)
    ops.load_model(ctx, weights.MODEL_IDS[model], weights.write_blob(w))
    rng = np.random.RandomState(H)
    fr = _frames(rng, B, H, W)
    mid = weights.MODEL_IDS[model]
    x16 = ops.preprocess_u8(ctx, torch.from_numpy(fr).cuda(), True, ops.input_layout(mid))
    loc, conf, landm = (t.cpu() for t in ops.retina_forward(ctx, mid, x16, H, W))
    P = nets.params_from_weights(w)
    x = torch.from_numpy(fr[..., ::-1].astype(np.float32) - np.array([104, 117, 123], np.float32)).permute(0, 3, 1, 2)
    with torch.no_grad():
        rloc, rconf, rlandm = nets.retina_forward(P, x.contiguous(), backbone)
    assert loc.shape == rloc.shape == (B, postproc.num_priors(H, W), 4)
    for got, ref, name in ((loc, rloc, "loc"), (conf, rconf, "conf"), (landm, rlandm, "landm")):
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= 0.02 * scale + 0.02, f"{name}: max err {err:.4g} vs scale {scale:.4g}"
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 5e-3, f"{name}: relative L2 error {rel:.3g}"


def test_irv1_forward_matches_oracle(ctx):
    from facerec_b200 import ops, weights
    from oracle import nets, crop
    w = weights.synth_weights("irv1", 0)
    ops.load_model(ctx, weights.MODEL_IDS["irv1"], weights.write_blob(w))
    rng = np.random.RandomState(1)
    fr = _frames(rng, 2, 200, 260)
    boxes = np.array([[0, 10, 20, 120, 150], [0, 100, 30, 160, 160], [1, 0, 0, 260, 200], [1, 50, 60, 33, 47],
                      [1, 200, 150, 60, 50]], np.int32)
    faces = ops.crop_resize_std(ctx, torch.from_numpy(fr).cuda(), torch.from_numpy(boxes).cuda())
    emb = ops.irv1_forward(ctx, faces).cpu()
    P = nets.params_from_weights(w)
    x = np.stack([crop.standardize(crop.resize_linear_u8(fr[f, y:y + h, x0:x0 + ww])) for f, x0, y, ww, h in boxes])
    with torch.no_grad():
        ref = nets.irv1_forward(P, torch.from_numpy(x).permute(0, 3, 1, 2).contiguous())
    cos = (emb * ref).sum(1)
    assert torch.allclose(emb.norm(dim=1), torch.ones(len(boxes)), atol=1e-5)
    assert cos.min().item() >= 0.999, cos
    # distinct faces stay distinct (the test is not vacuous)
    assert (ref @ ref.T - torch.eye(len(boxes))).max().item() < 0.9999


def test_forwards_are_bit_reproducible(ctx):
    """The same input gives the same bits every time, with one and with two MMA issuer warps: no output depends on the
    timing of the pipeline roles (a stage-split between the issuers was dropped for breaking exactly this)."""
    from facerec_b200 import ops, weights, _lib
    rng = np.random.RandomState(3)
    fr = torch.from_numpy(_frames(rng, 3, 270, 480)).cuda()
    mid = weights.MODEL_IDS["retina_r50"]
    ops.load_model(ctx, mid, weights.write_blob(weights.synth_weights("retina_r50", 0)))
    x16 = ops.preprocess_u8(ctx, fr, True, ops.input_layout(mid))
    runs = [[t.clone() for t in ops.retina_forward(ctx, mid, x16, 270, 480)] for _ in range(3)]
    for r in runs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(runs[0], r))
    solo = _lib.Context(0)
    try:
        solo.set_option(_lib.FRB_OPT_DUAL_ISSUE, 0)
        ops.load_model(solo, mid, weights.write_blob(weights.synth_weights("retina_r50", 0)))
        one = ops.retina_forward(solo, mid, x16, 270, 480)
        assert all(torch.equal(a, b) for a, b in zip(runs[0], one))
    finally:
        solo.close()
    ops.load_model(ctx, weights.MODEL_IDS["irv1"], weights.write_blob(weights.synth_weights("irv1", 0)))
    faces = torch.zeros(24, 160, 160, 8)
    faces[..., :3] = torch.randn(24, 160, 160, 3, generator=torch.Generator().manual_seed(1)) * 0.5
    f16 = faces.half().cuda()
    e = [ops.irv1_forward(ctx, f16).clone() for _ in range(3)]
    assert torch.equal(e[0], e[1]) and torch.equal(e[0], e[2])


@pytest.mark.parametrize("pdl", [0, 2])
def test_dependent_launch_policy_does_not_change_results(ctx, pdl):
    """FRB_OPT_PDL only moves a kernel's set-up under the previous kernel's tail: plain stream order (0), the default
    (short launches only) and 'always' (2) give the same bits, back to back, for both networks."""
    from facerec_b200 import ops, weights, _lib
    rng = np.random.RandomState(5)
    fr = torch.from_numpy(_frames(rng, 2, 270, 480)).cuda()
    mid = weights.MODEL_IDS["retina_r50"]
    irv1 = weights.MODEL_IDS["irv1"]
    faces = torch.zeros(40, 160, 160, 8)
    faces[..., :3] = torch.randn(40, 160, 160, 3, generator=torch.Generator().manual_seed(2)) * 0.5
    f16 = faces.half().cuda()
    outs = []
    for c, opt in ((ctx, None), (_lib.Context(0), pdl)):
        try:
            if opt is not None:
                c.set_option(_lib.FRB_OPT_PDL, opt)
            ops.load_model(c, mid, weights.write_blob(weights.synth_weights("retina_r50", 0)))
            ops.load_model(c, irv1, weights.write_blob(weights.synth_weights("irv1", 0)))
            x16 = ops.preprocess_u8(c, fr, True, ops.input_layout(mid))
            res = []
            for _ in range(3):                      # back to back: set-up of launch k+1 overlaps the tail of launch k
                res = [t.clone() for t in ops.retina_forward(c, mid, x16, 270, 480)] + [ops.irv1_forward(c, f16).clone()]
            outs.append(res)
        finally:
            if opt is not None:
                c.close()
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))


def test_graph_replay_equals_eager_launches(ctx):
    """FRB_OPT_GRAPHS = 1 (opt-in): short forwards (InceptionResnetV1, RetinaFace-MobileNet0.25 at 360p) are captured into a CUDA
    graph on their second call and replayed through net-owned staging buffers; results are the bits of the eager launches, for
    changing input / output pointers and interleaved batch sizes (one graph per shape)."""
    from facerec_b200 import ops, weights, _lib
    rng = np.random.RandomState(8)
    irv1 = weights.MODEL_IDS["irv1"]
    mnet = weights.MODEL_IDS["retina_mnet025"]
    gen = torch.Generator().manual_seed(4)
    faces = {n: torch.zeros(n, 160, 160, 8) for n in (8, 32)}
    for n in faces:
        faces[n][..., :3] = torch.randn(n, 160, 160, 3, generator=gen) * 0.5
    fr = [torch.from_numpy(_frames(rng, 4, 360, 640)).cuda() for _ in range(2)]
    eager = _lib.Context(0)
    ctx.set_option(_lib.FRB_OPT_GRAPHS, 1)
    try:
        eager.set_option(_lib.FRB_OPT_GRAPHS, 0)
        want = {}
        for c in (eager, ctx):
            ops.load_model(c, irv1, weights.write_blob(weights.synth_weights("irv1", 0)))
            ops.load_model(c, mnet, weights.write_blob(weights.synth_weights("retina_mnet025", 0)))
        for n in faces:
            want[n] = ops.irv1_forward(eager, faces[n].half().cuda()).clone()
        want_det = [[t.clone() for t in ops.retina_forward(eager, mnet, ops.preprocess_u8(eager, f, True), 360, 640)] for f in fr]
        l0 = ctx.launch_count()
        for rep in range(4):                          # call 1 eager, call 2 captures, calls 3+ replay; shapes interleaved
            for n in (8, 32):
                got = ops.irv1_forward(ctx, faces[n].half().cuda().clone())      # a fresh input pointer every time
                assert torch.equal(got, want[n]), (rep, n)
            for f, w in zip(fr, want_det):
                got = ops.retina_forward(ctx, mnet, ops.preprocess_u8(ctx, f, True), 360, 640)
                assert all(torch.equal(a, b) for a, b in zip(got, w)), rep
        assert ctx.launch_count() - l0 > 4 * 2 * 100   # replays are counted with the kernels they contain
    finally:
        ctx.set_option(_lib.FRB_OPT_GRAPHS, 0)
        eager.close()


@pytest.mark.parametrize("B,H,W", [(2, 270, 480), (3, 135, 97), (1, 64, 64), (9, 270, 480), (64, 135, 240)])
def test_fused_bottleneck_tail_does_not_change_results(ctx, B, H, W):
    """FRB_OPT_FUSE_TAIL = 1: conv3 + skip and the next block's conv1 as one launch (conv_tail.cu: layer1 with resident weights,
    layer2 with streamed weights in two halves of 256 channels) accumulate in the
    same order as the two separate launches: heads identical bit for bit, also on ragged last tiles (M % 128 != 0) and
    with several tiles per CTA (the last two shapes: 574 and 1 040 tiles over 148 CTAs),
    toggled back and forth on ONE context (the cached plans are rebuilt)."""
    from facerec_b200 import ops, weights, _lib
    rng = np.random.RandomState(11)
    fr = torch.from_numpy(_frames(rng, B, H, W)).cuda()
    mid = weights.MODEL_IDS["retina_r50"]
    c = _lib.Context(0)
    try:
        ops.load_model(c, mid, weights.write_blob(weights.synth_weights("retina_r50", 0)))
        x16 = ops.preprocess_u8(c, fr, True, ops.input_layout(mid))
        outs = []
        for fuse in (1, 0, 1, 3):
            c.set_option(_lib.FRB_OPT_FUSE_TAIL, fuse)
            n0 = c.launch_count()
            outs.append([t.clone() for t in ops.retina_forward(c, mid, x16, H, W)])
            outs[-1].append(c.launch_count() - n0)
    finally:
        c.close()
    assert outs[0][3] == outs[2][3] == outs[1][3] - 6          # three pairs in layer1 (one of them across the layer edge), three in layer2
    assert outs[3][3] == outs[1][3] - 7                        # ... and layer1.0's skip conv computed inside the first of them
    for k in range(3):
        assert torch.isfinite(outs[0][k]).all()
        assert torch.equal(outs[0][k], outs[1][k]) and torch.equal(outs[0][k], outs[2][k])
        # level 3 skips ONE rounding to 16 bits (the skip conv's output): not the same bits, the same network
        ref = outs[1][k].float()
        assert (outs[3][k].float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("B,H,W", [(2, 360, 640), (3, 135, 97)])
def test_fused_mobilenet_blocks_match_separate_launches(ctx, B, H, W):
    """FRB_OPT_FUSE_DWPW: depthwise 3x3 + pointwise 1x1 in one launch (glue.cu K4f, mma.sync GEMM) keeps the rounding points
    of the two separate launches (16-bit depthwise output, fp32 accumulation); only the summation order of the pointwise
    GEMM differs.  0 = separate, 1 = default (blocks up to 32 channels), 2 = every block that fits."""
    from facerec_b200 import ops, weights, _lib
    rng = np.random.RandomState(12)
    fr = torch.from_numpy(_frames(rng, B, H, W)).cuda()
    mid = weights.MODEL_IDS["retina_mnet025"]
    c = _lib.Context(0)
    try:
        ops.load_model(c, mid, weights.write_blob(weights.synth_weights("retina_mnet025", 0)))
        x16 = ops.preprocess_u8(c, fr, True, ops.input_layout(mid))
        outs = []
        for fuse in (0, 1, 2):
            c.set_option(_lib.FRB_OPT_FUSE_DWPW, fuse)
            n0 = c.launch_count()
            outs.append([t.clone() for t in ops.retina_forward(c, mid, x16, H, W)])
            outs[-1].append(c.launch_count() - n0)
    finally:
        c.close()
    assert outs[1][3] == outs[0][3] - 4 and outs[2][3] == outs[0][3] - 12
    for k in range(3):
        ref = outs[0][k].float()
        for o in (outs[1], outs[2]):
            assert torch.isfinite(o[k]).all()
            assert (o[k].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("B,H,W", [(2, 270, 480), (3, 135, 97), (1, 64, 64), (2, 1080, 1920), (5, 203, 311)])
def test_fused_stem_max_pool_does_not_change_results(ctx, B, H, W):
    """FRB_OPT_FUSE_POOL: the ResNet stem's 3x3 / stride 2 max pool computed in the stem conv's epilogue (inner positions of
    an 8 x 16 patch stored, seam positions max-reduced into the pre-zeroed tensor).  max is exact and order independent:
    heads identical bit for bit with the separate max-pool launch, on ragged patches and odd sizes too, and repeatable."""
    from facerec_b200 import ops, weights, _lib
    rng = np.random.RandomState(13)
    fr = torch.from_numpy(_frames(rng, B, H, W)).cuda()
    mid = weights.MODEL_IDS["retina_r50"]
    c = _lib.Context(0)
    try:
        ops.load_model(c, mid, weights.write_blob(weights.synth_weights("retina_r50", 0)))
        x16 = ops.preprocess_u8(c, fr, True, ops.input_layout(mid))
        outs = []
        for fuse in (1, 0, 1):
            c.set_option(_lib.FRB_OPT_FUSE_POOL, fuse)
            n0 = c.launch_count()
            outs.append([t.clone() for t in ops.retina_forward(c, mid, x16, H, W)])
            outs[-1].append(c.launch_count() - n0)
    finally:
        c.close()
    assert outs[0][3] == outs[1][3] == outs[2][3]              # stem + zero fill instead of stem + max pool
    for k in range(3):
        assert torch.isfinite(outs[0][k]).all()
        assert torch.equal(outs[0][k], outs[1][k]) and torch.equal(outs[0][k], outs[2][k])


def test_fusions_in_bf16_give_the_bits_of_the_separate_launches(ctx):
    """The fused kernels are templated on the element type: with FRB_OPT_BF16 the bit-preserving fusions (tails at level 1 in
    layer1 and layer2, stem pool) still reproduce the separate launches exactly, also with several tiles per CTA."""
    from facerec_b200 import ops, weights, _lib
    rng = np.random.RandomState(14)
    B, H, W = 9, 270, 480
    fr = torch.from_numpy(_frames(rng, B, H, W)).cuda()
    mid = weights.MODEL_IDS["retina_r50"]
    c = _lib.Context(0)
    try:
        c.set_option(_lib.FRB_OPT_BF16, 1)
        ops.load_model(c, mid, weights.write_blob(weights.synth_weights("retina_r50", 0)))
        x16 = ops.preprocess_u8(c, fr, True, ops.input_layout(mid))
        outs = []
        for tail, tail2, pool in ((1, 1, 1), (0, 0, 0), (1, 0, 1), (2, 1, 0)):
            c.set_option(_lib.FRB_OPT_FUSE_TAIL, tail)
            c.set_option(_lib.FRB_OPT_FUSE_TAIL2, tail2)
            c.set_option(_lib.FRB_OPT_FUSE_POOL, pool)
            outs.append([t.clone() for t in ops.retina_forward(c, mid, x16, H, W)])
    finally:
        c.close()
    for o in outs[1:]:
        for k in range(3):
            assert torch.isfinite(o[k]).all()
            assert torch.equal(outs[0][k], o[k])
