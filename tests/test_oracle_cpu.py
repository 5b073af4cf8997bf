"""CPU suite: pins of the oracle against the libraries the reference calls, the weight blob, the C-ABI export list
and the multi-rank host logic (gloo, world size 2).  No GPU compute here."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_resize_restatement_is_bit_exact_with_cv2():
    import cv2
    from oracle import crop
    rng = np.random.RandomState(0)
    for _ in range(120):
        h, w = rng.randint(1, 420), rng.randint(1, 420)
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        assert np.array_equal(cv2.resize(img, (160, 160), interpolation=cv2.INTER_LINEAR), crop.resize_linear_u8(img))
    assert np.array_equal(crop.standardize(np.array([[0, 255]], np.uint8)), np.array([[-127.5 / 128, 127.5 / 128]], np.float32))


def test_crop_box_matches_product_crop_rect():
    from oracle import crop
    from facerec_b200.embedder import crop_rect, cropBox
    rng = np.random.RandomState(1)
    img = rng.randint(0, 256, (90, 120, 3)).astype(np.uint8)
    for _ in range(200):
        box = [int(rng.randint(-5, 110)), int(rng.randint(-5, 85)), int(rng.randint(1, 80)), int(rng.randint(1, 80))]
        assert crop.crop_rect(box, 120, 90, 16) == crop_rect(box, 120, 90, 16)
        a, b = crop.crop_box(img, box, 16), cropBox(img, {"box": box}, 16)
        assert a.shape == b.shape and np.array_equal(a, b)


def test_nms_oracle_matches_torchvision():
    import torchvision
    from oracle import postproc
    rng = np.random.RandomState(2)
    for n in (1, 7, 60, 300):
        xy = rng.uniform(0, 300, (n, 2)).astype(np.float32)
        wh = rng.uniform(10, 90, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = rng.uniform(0.95, 1, n).astype(np.float32)
        keep = postproc.nms(boxes, scores, np.arange(n), 0.4, plus_one=False)
        ref = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.4).numpy()
        assert np.array_equal(keep, ref)


def test_priorbox_layout_and_decode_identity():
    from oracle import postproc
    for H, W in ((360, 640), (1080, 1920), (135, 97)):
        pri = postproc.priorbox(H, W)
        assert len(pri) == postproc.num_priors(H, W)
    assert postproc.num_priors(360, 640) == 9520 and postproc.num_priors(1080, 1920) == 85200
    pri = postproc.priorbox(360, 640)
    # level 0, cell (0,0): centre (4,4) px, anchors 16 and 32 px
    assert np.allclose(pri[0] * [640, 360, 640, 360], [4, 4, 16, 16]) and np.allclose(pri[1] * [640, 360, 640, 360], [4, 4, 32, 32])
    boxes, lm = postproc.decode(np.zeros((len(pri), 4), np.float32), np.zeros((len(pri), 10), np.float32), pri, 360, 640)
    assert np.allclose(boxes[0], [-4, -4, 12, 12], atol=1e-4) and np.allclose(lm[0], [4, 4] * 5, atol=1e-4)
    s = postproc.face_scores(np.array([[0.0, np.log(19.0)]], np.float32))
    assert abs(float(s[0]) - 0.95) < 1e-6


def test_detect_post_filter_follows_reference_wrapper():
    from oracle import postproc
    d = np.zeros((2, 15), np.float32)
    d[0, :5] = [10.2, 20.7, 50.9, 39.9, 0.99]           # min side 19.2 -> dropped at min_size 20
    d[1, :5] = [10.2, 20.7, 50.9, 41.0, 0.99]
    d[1, 5:] = np.arange(10) + 0.9
    out = postproc.detect_from_dets(d, 20)
    assert len(out) == 1 and out[0]["keypoints"]["left_eye"] == (0, 1) and out[0]["keypoints"]["mouth_right"] == (8, 9)
    assert all(isinstance(v, float) for v in out[0]["box"])


def test_weight_factory_and_blob_roundtrip():
    from facerec_b200 import weights as W
    w = W.synth_weights("retina_mnet025", 3)
    w2 = W.synth_weights("retina_mnet025", 3)
    assert all(np.array_equal(w[k], w2[k]) for k in w)
    assert sum(v.size for v in w.values()) == 423264
    r = W.read_blob(W.write_blob(w))
    assert list(r) == list(w) and all(np.array_equal(r[k], w[k]) for k in w)
    specs = {s.name: s for s in W.model_specs("irv1")}
    assert len(specs) == 133 and specs["repeat_2.3.branch1.1"].kw == 7 and specs["last_linear"].cout == 512
    assert sum(int(np.prod([s.cout, s.kh, s.kw, s.cin])) + s.cout for s in W.model_specs("retina_r50")) == 27264608


def test_fold_bn_equals_torch_batchnorm_eval():
    from facerec_b200 import weights as W
    rng = np.random.RandomState(0)
    w = rng.standard_normal((6, 3, 3, 4)).astype(np.float32)
    g, b, m, v = (rng.uniform(0.5, 1.5, 6).astype(np.float32) for _ in range(4))
    wf, bf = W.fold_bn(w, g, b, m, v, 1e-3)
    x = torch.randn(2, 4, 9, 9)
    bn = torch.nn.BatchNorm2d(6, eps=1e-3).eval()
    bn.weight.data, bn.bias.data = torch.from_numpy(g), torch.from_numpy(b)
    bn.running_mean.data, bn.running_var.data = torch.from_numpy(m), torch.from_numpy(v)
    ref = bn(torch.nn.functional.conv2d(x, torch.from_numpy(w).permute(0, 3, 1, 2)))
    got = torch.nn.functional.conv2d(x, torch.from_numpy(wf).permute(0, 3, 1, 2), torch.from_numpy(bf))
    assert torch.allclose(ref, got, atol=1e-5)


def test_oracle_networks_shapes():
    from facerec_b200 import weights as W
    from oracle import nets, postproc
    P = nets.params_from_weights(W.synth_weights("retina_mnet025", 0))
    with torch.no_grad():
        loc, conf, landm = nets.retina_forward(P, torch.zeros(1, 3, 97, 135), "mnet025")
    n = postproc.num_priors(97, 135)
    assert loc.shape == (1, n, 4) and conf.shape == (1, n, 2) and landm.shape == (1, n, 10)


def test_library_exports_every_header_symbol():
    """The C-ABI library loads (no GPU needed) and exports exactly what include/facerec_b200.h (+ the debug-only additions
    of include/facerec_b200_debug.h) declares."""
    import ctypes
    from facerec_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "facerec_b200.h")).read()
    header += open(os.path.join(ROOT, "include", "facerec_b200_debug.h")).read()
    # the timing experiments are not part of the public header
    assert "SKIP_LOADS" not in open(os.path.join(ROOT, "include", "facerec_b200.h")).read()
    declared = set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.frb_version() == 100
    assert lib.frb_num_priors(1080, 1920) == 85200
    # every enumerator of the header has the same value in the ctypes mirror, and the other way round
    enums = {k: int(v) for k, v in re.findall(r"\b(FRB_[A-Z0-9_]+)\s*=\s*(-?\d+)", header)}
    mirror = {k: v for k, v in vars(_lib).items() if k.startswith("FRB_") and isinstance(v, int)}
    assert enums and {k: enums.get(k) for k in mirror} == mirror, set(enums.items()) ^ set(mirror.items())
    assert {k for k in enums if k.startswith("FRB_OPT_")} == {k for k in mirror if k.startswith("FRB_OPT_")}


def test_product_fails_loudly_without_gpu():
    from facerec_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.FrbError):
        _lib.Context(0)


def test_product_never_imports_oracle():
    for root, _, names in os.walk(os.path.join(ROOT, "facerec_b200")):
        for n in names:
            if n.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, n)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), n


def test_balanced_row_blocks_cover_the_triangle():
    from facerec_b200.distance import balanced_row_blocks, condensed_offset
    for n, parts in ((50000, 8), (1000, 3), (7, 8), (2, 2)):
        blocks = balanced_row_blocks(n, parts)
        assert blocks[0][0] == 0 and sum(r for _, r in blocks) == n
        assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(parts - 1))
        if n >= 1000:
            areas = [condensed_offset(n, a + r) - condensed_offset(n, a) for a, r in blocks]
            assert max(areas) / (sum(areas) / parts) < 1.01


_GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from scipy.spatial.distance import pdist
from facerec_b200 import distance
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=rank, world_size=world)
rng = np.random.RandomState(0)
X = rng.standard_normal((301, 32)).astype(np.float32)
lo, hi = (0, 130) if rank == 0 else (130, 301)          # ragged shards, as produced by frame sharding
def cpu_block(Xa, r0, rows):
    n = Xa.shape[0]
    full = pdist(Xa.numpy().astype(np.float64)).astype(np.float32)
    return torch.from_numpy(full[distance.condensed_offset(n, r0):distance.condensed_offset(n, r0 + rows)])
sl, off, n = distance.pdist_sharded(torch.from_numpy(X[lo:hi]), group=None, block_fn=cpu_block)
assert n == 301
ref = pdist(X.astype(np.float64)).astype(np.float32)
assert np.array_equal(sl.numpy(), ref[off:off + len(sl)])
parts = [None, None]
dist.all_gather_object(parts, (off, sl.numpy()))
if rank == 0:
    whole = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
    assert np.array_equal(whole, ref)
    print("OK")
dist.destroy_process_group()
"""


def test_row_sharded_pdist_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0][0]


def test_oracle_r50_backbone_matches_torchvision():
    """Pins the FLOP-heavy part of the detector oracle: the functional ResNet-50 body of oracle/nets.py, fed the
    BN-folded weights of a torchvision.models.resnet50 with randomised BatchNorm statistics, reproduces torchvision's
    own layer2 / layer3 / layer4 activations (the RetinaFace-R50 backbone named by north_star)."""
    import torch
    import torchvision
    from facerec_b200.weights import fold_bn
    from oracle import nets
    torch.manual_seed(0)
    m = torchvision.models.resnet50(weights=None).eval()
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)

    W = {}

    def put(name, conv, bn):
        w, b = fold_bn(conv.weight.detach().numpy().transpose(0, 2, 3, 1), bn.weight.detach().numpy(), bn.bias.detach().numpy(),
                       bn.running_mean.numpy(), bn.running_var.numpy(), bn.eps)
        W[name + ".w"], W[name + ".b"] = w, b

    put("conv1", m.conv1, m.bn1)
    for l in range(1, 5):
        for i, blk in enumerate(getattr(m, f"layer{l}")):
            p = f"layer{l}.{i}"
            put(p + ".conv1", blk.conv1, blk.bn1)
            put(p + ".conv2", blk.conv2, blk.bn2)
            put(p + ".conv3", blk.conv3, blk.bn3)
            if blk.downsample is not None:
                put(p + ".down", blk.downsample[0], blk.downsample[1])
    x = torch.randn(2, 3, 96, 128, generator=g)
    with torch.no_grad():
        y = m.maxpool(m.relu(m.bn1(m.conv1(x))))
        y = m.layer1(y)
        ref = []
        for l in (m.layer2, m.layer3, m.layer4):
            y = l(y)
            ref.append(y)
        got = nets.r50_features(nets.params_from_weights(W), x)
    assert [tuple(t.shape) for t in got] == [(2, 512, 12, 16), (2, 1024, 6, 8), (2, 2048, 3, 4)]
    for a, b in zip(got, ref):
        assert float((a - b).abs().max()) <= 2e-4 * float(b.abs().max())


def _fake_state_dict(model, rng):
    """A PyTorch-style state dict under the PUBLIC layer names, random values, plus what folding it must give."""
    from facerec_b200 import weights as W
    sd, want = {}, {}

    def conv_bn(conv, bn, ours, sp, eps):
        w = rng.standard_normal((sp.cout, sp.cin, sp.kh, sp.kw)).astype(np.float32)
        g, b = rng.uniform(0.5, 1.5, sp.cout).astype(np.float32), rng.standard_normal(sp.cout).astype(np.float32)
        m, v = rng.standard_normal(sp.cout).astype(np.float32), rng.uniform(0.5, 1.5, sp.cout).astype(np.float32)
        sd[conv + ".weight"] = torch.from_numpy(w)
        for k, a in (("weight", g), ("bias", b), ("running_mean", m), ("running_var", v)):
            sd[f"{bn}.{k}"] = torch.from_numpy(a)
        sd[bn + ".num_batches_tracked"] = torch.tensor(1)
        want[ours + ".w"], want[ours + ".b"] = W.fold_bn(w.transpose(0, 2, 3, 1), g, b, m, v, eps)

    def conv_bias(name, ours, sp):
        w = rng.standard_normal((sp.cout, sp.cin, sp.kh, sp.kw)).astype(np.float32)
        b = rng.standard_normal(sp.cout).astype(np.float32)
        sd[name + ".weight"], sd[name + ".bias"] = torch.from_numpy(w), torch.from_numpy(b)
        want[ours + ".w"], want[ours + ".b"] = w.transpose(0, 2, 3, 1), b

    specs = {sp.name: sp for sp in W.model_specs(model)}
    if model == "irv1":
        for n, sp in specs.items():
            if n == "last_linear":
                w = rng.standard_normal((512, 1792)).astype(np.float32)
                g, b = rng.uniform(0.5, 1.5, 512).astype(np.float32), rng.standard_normal(512).astype(np.float32)
                m, v = rng.standard_normal(512).astype(np.float32), rng.uniform(0.5, 1.5, 512).astype(np.float32)
                sd["last_linear.weight"] = torch.from_numpy(w)
                for k, a in (("weight", g), ("bias", b), ("running_mean", m), ("running_var", v)):
                    sd[f"last_bn.{k}"] = torch.from_numpy(a)
                want[n + ".w"], want[n + ".b"] = W.fold_bn(w[:, None, None, :], g, b, m, v, 1e-3)
            elif sp.bn:
                conv_bn(n + ".conv", n + ".bn", n, sp, 1e-3)
            else:
                conv_bias(n, n, sp)
        return sd, want
    if model == "retina_mnet025":
        conv_bn("body.stage1.0.0", "body.stage1.0.1", "body.0", specs["body.0"], 1e-5)
        blocks = [("stage1", i) for i in range(1, 6)] + [("stage2", i) for i in range(6)] + [("stage3", i) for i in range(2)]
        for n, (st, i) in enumerate(blocks, start=1):
            conv_bn(f"body.{st}.{i}.0", f"body.{st}.{i}.1", f"body.{n}.dw", specs[f"body.{n}.dw"], 1e-5)
            conv_bn(f"body.{st}.{i}.3", f"body.{st}.{i}.4", f"body.{n}.pw", specs[f"body.{n}.pw"], 1e-5)
    else:
        conv_bn("body.conv1", "body.bn1", "conv1", specs["conv1"], 1e-5)
        for l, nb in enumerate(W.R50_BLOCKS, start=1):
            for i in range(nb):
                for c in (1, 2, 3):
                    conv_bn(f"body.layer{l}.{i}.conv{c}", f"body.layer{l}.{i}.bn{c}", f"layer{l}.{i}.conv{c}", specs[f"layer{l}.{i}.conv{c}"], 1e-5)
                if i == 0:
                    conv_bn(f"body.layer{l}.{i}.downsample.0", f"body.layer{l}.{i}.downsample.1", f"layer{l}.{i}.down", specs[f"layer{l}.{i}.down"], 1e-5)
    for k in (1, 2, 3):
        conv_bn(f"fpn.output{k}.0", f"fpn.output{k}.1", f"fpn.out{k}", specs[f"fpn.out{k}"], 1e-5)
    for k in (1, 2):
        conv_bn(f"fpn.merge{k}.0", f"fpn.merge{k}.1", f"fpn.merge{k}", specs[f"fpn.merge{k}"], 1e-5)
    for k in (1, 2, 3):
        for ours, theirs in (("c3", "conv3X3"), ("c5_1", "conv5X5_1"), ("c5_2", "conv5X5_2"), ("c7_2", "conv7X7_2"), ("c7_3", "conv7x7_3")):
            conv_bn(f"ssh{k}.{theirs}.0", f"ssh{k}.{theirs}.1", f"ssh{k}.{ours}", specs[f"ssh{k}.{ours}"], 1e-5)
        for ours, theirs in (("cls", "ClassHead"), ("box", "BboxHead"), ("ldm", "LandmarkHead")):
            conv_bias(f"{theirs}.{k - 1}.conv1x1", f"head{k}.{ours}", specs[f"head{k}.{ours}"])
    return sd, want


@pytest.mark.parametrize("model", ["retina_mnet025", "retina_r50", "irv1"])
def test_checkpoint_converter_folds_public_state_dicts(model, tmp_path):
    """--detector-weights / --embedder-weights: a state dict under the public RetinaFace / InceptionResnetV1 names (with a
    DataParallel 'module.' prefix) converts to exactly the BN-folded [Cout,kh,kw,Cin] tensors frb_load_model expects, and
    survives the .pt -> .frbw -> .npz round trip."""
    from facerec_b200 import convert, weights as W
    rng = np.random.RandomState(3)
    sd, want = _fake_state_dict(model, rng)
    path = str(tmp_path / "ckpt.pth")
    torch.save({"module." + k: v for k, v in sd.items()}, path)
    got = convert.load_weights(path, model)
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape and np.array_equal(got[k], np.ascontiguousarray(want[k], dtype=np.float32)), k
    blob = str(tmp_path / "w.frbw")
    open(blob, "wb").write(W.write_blob(got))
    again = convert.load_weights(blob, model)
    assert all(np.array_equal(again[k], got[k]) for k in got)
    np.savez(str(tmp_path / "w.npz"), **got)
    assert all(np.array_equal(convert.load_weights(str(tmp_path / "w.npz"), model)[k], got[k]) for k in got)
    bad = dict(got)
    bad.pop(next(iter(bad)))
    np.savez(str(tmp_path / "bad.npz"), **bad)
    with pytest.raises(ValueError):
        convert.load_weights(str(tmp_path / "bad.npz"), model)


def test_conv_tail_barrier_protocol_has_no_deadlock_or_parity_aliasing():
    """The barrier protocol of conv_tail_kernel (fused bottleneck tail) replayed under random interleavings: the shipped
    protocol passes every configuration; without in-order slot taking the model finds the parity aliasing that hung the
    first GPU version (so the model is able to see that class of bug)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import conv_tail_protocol_sim as sim
    assert sim.sweep(in_order=True, seeds=15) == []
    assert sim.sweep2(seeds=12) == []                      # layer2 variant (two halves per tile, two load rings, two issuers)
    assert any("aliased" in str(d) or "deadlock" in str(d) for _, d in sim.sweep(in_order=False, seeds=15))
