"""Checkpoint conversion: public PyTorch state dicts -> the FRBW weight blob of frb_load_model.

The reference gets its weights from third-party packages at run time (`RetinaFace(quality="normal")`,
facerec/detector.py:64; `FaceNet(key)`, facerec/extract.py:392); none are available offline, so everything in this
repo that is measured runs on seeded random-init weights (weights.synth_weights).  For real use the extractor takes
`--detector-weights` / `--embedder-weights` (facerec_b200/extract.py), which accept

  * `*.frbw`   a blob written by `weights.write_blob` (what frb_load_model reads),
  * `*.npz`    arrays under this repo's layer names: "<layer>.w" [Cout,kh,kw,Cin] and "<layer>.b" [Cout], BN folded,
  * `*.pt|*.pth|*.bin`   a PyTorch state dict under the PUBLIC names of the architectures BASELINE.json names:
        RetinaFace as published with the paper's PyTorch port (MobileNetV1-0.25 `body.stage{1,2,3}.*`, torchvision
        ResNet-50 `body.conv1/bn1/layer{1..4}.*`, `fpn.output{k}/merge{k}`, `ssh{k}.conv3X3/conv5X5_1/conv5X5_2/
        conv7X7_2/conv7x7_3`, `ClassHead/BboxHead/LandmarkHead.{k}.conv1x1`), and
        FaceNet InceptionResnetV1 (`conv2d_1a.conv/bn`, `repeat_{1,2,3}.{i}.branch*.conv/bn`, `mixed_6a/7a.*`, `block8.*`,
        `last_linear`, `last_bn`).
    Conv weights are [Cout,Cin,kh,kw] there; BatchNorm (eval) is folded into the preceding conv here.

Embeddings written by the extractor are only comparable with embeddings computed by the SAME weights: an existing
actor database (facerec/prepare-actors.py) must be recomputed when the embedder changes.
"""
from __future__ import annotations

import os
from typing import Dict, Mapping

import numpy as np

from . import weights as W


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float32)


def _conv_w(sd: Mapping, key: str) -> np.ndarray:
    """[Cout,Cin,kh,kw] -> [Cout,kh,kw,Cin]"""
    return np.ascontiguousarray(_np(sd[key]).transpose(0, 2, 3, 1))


def _conv_bn(sd: Mapping, conv: str, bn: str, eps: float):
    w = _conv_w(sd, conv + ".weight")
    return W.fold_bn(w, _np(sd[bn + ".weight"]), _np(sd[bn + ".bias"]), _np(sd[bn + ".running_mean"]),
                     _np(sd[bn + ".running_var"]), eps)


def _strip(sd: Mapping) -> Dict[str, object]:
    """DataParallel checkpoints prefix every key with 'module.'."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def retina_from_state_dict(sd: Mapping, backbone: str) -> Dict[str, np.ndarray]:
    sd = _strip(sd)
    out: Dict[str, np.ndarray] = {}

    def put(name, wb):
        out[name + ".w"], out[name + ".b"] = wb

    if backbone == "mnet025":
        put("body.0", _conv_bn(sd, "body.stage1.0.0", "body.stage1.0.1", 1e-5))
        blocks = [("stage1", i) for i in range(1, 6)] + [("stage2", i) for i in range(6)] + [("stage3", i) for i in range(2)]
        for n, (stage, i) in enumerate(blocks, start=1):
            p = f"body.{stage}.{i}"
            put(f"body.{n}.dw", _conv_bn(sd, p + ".0", p + ".1", 1e-5))      # [C,1,3,3] -> [C,3,3,1]
            put(f"body.{n}.pw", _conv_bn(sd, p + ".3", p + ".4", 1e-5))
    elif backbone == "r50":
        put("conv1", _conv_bn(sd, "body.conv1", "body.bn1", 1e-5))
        for l, nb in enumerate(W.R50_BLOCKS, start=1):
            for i in range(nb):
                p, q = f"body.layer{l}.{i}", f"layer{l}.{i}"
                for c in (1, 2, 3):
                    put(f"{q}.conv{c}", _conv_bn(sd, f"{p}.conv{c}", f"{p}.bn{c}", 1e-5))
                if i == 0:
                    put(q + ".down", _conv_bn(sd, p + ".downsample.0", p + ".downsample.1", 1e-5))
    else:
        raise ValueError(backbone)
    for k in (1, 2, 3):
        put(f"fpn.out{k}", _conv_bn(sd, f"fpn.output{k}.0", f"fpn.output{k}.1", 1e-5))
    for k in (1, 2):
        put(f"fpn.merge{k}", _conv_bn(sd, f"fpn.merge{k}.0", f"fpn.merge{k}.1", 1e-5))
    ssh = {"c3": "conv3X3", "c5_1": "conv5X5_1", "c5_2": "conv5X5_2", "c7_2": "conv7X7_2", "c7_3": "conv7x7_3"}
    for k in (1, 2, 3):
        for ours, theirs in ssh.items():
            put(f"ssh{k}.{ours}", _conv_bn(sd, f"ssh{k}.{theirs}.0", f"ssh{k}.{theirs}.1", 1e-5))
        for ours, theirs in (("cls", "ClassHead"), ("box", "BboxHead"), ("ldm", "LandmarkHead")):
            p = f"{theirs}.{k - 1}.conv1x1"
            out[f"head{k}.{ours}.w"] = _conv_w(sd, p + ".weight")
            out[f"head{k}.{ours}.b"] = _np(sd[p + ".bias"])
    return out


def irv1_from_state_dict(sd: Mapping) -> Dict[str, np.ndarray]:
    sd = _strip(sd)
    out: Dict[str, np.ndarray] = {}
    for sp in W.irv1_specs():
        n = sp.name
        if n == "last_linear":
            w = _np(sd["last_linear.weight"])[:, None, None, :]               # Linear [512,1792] -> 1x1 conv
            out[n + ".w"], out[n + ".b"] = W.fold_bn(w, _np(sd["last_bn.weight"]), _np(sd["last_bn.bias"]),
                                                     _np(sd["last_bn.running_mean"]), _np(sd["last_bn.running_var"]), 1e-3)
        elif sp.bn:
            out[n + ".w"], out[n + ".b"] = _conv_bn(sd, n + ".conv", n + ".bn", 1e-3)
        else:                                                                   # the block's plain 1x1 `conv2d` with bias
            out[n + ".w"] = _conv_w(sd, n + ".weight")
            out[n + ".b"] = _np(sd[n + ".bias"])
    return out


def check_complete(model: str, weights: Mapping[str, np.ndarray]) -> None:
    for sp in W.model_specs(model):
        w, b = weights.get(sp.name + ".w"), weights.get(sp.name + ".b")
        want = (sp.cout, sp.kh, sp.kw, sp.cin)
        if w is None or b is None:
            raise ValueError(f"{model}: weights for layer '{sp.name}' are missing")
        if tuple(w.shape) != want or tuple(b.shape) != (sp.cout,):
            raise ValueError(f"{model}: layer '{sp.name}' has shape {tuple(w.shape)} / {tuple(b.shape)}, expected {want} / ({sp.cout},)")


def load_weights(path: str, model: str) -> Dict[str, np.ndarray]:
    """Read `path` (.frbw / .npz / PyTorch state dict) as the weight dict of `model`
    ("retina_mnet025", "retina_r50" or "irv1"), validated against the architecture table."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".frbw":
        with open(path, "rb") as f:
            w = W.read_blob(f.read())
    elif ext == ".npz":
        with np.load(path) as z:
            w = {k: np.asarray(z[k], dtype=np.float32) for k in z.files}
    elif ext in (".pt", ".pth", ".bin"):
        import torch
        sd = torch.load(path, map_location="cpu", weights_only=True)
        if isinstance(sd, dict) and "state_dict" in sd:
            sd = sd["state_dict"]
        w = irv1_from_state_dict(sd) if model == "irv1" else retina_from_state_dict(sd, model.split("_", 1)[1])
    else:
        raise ValueError(f"{path}: unknown weight file type (expected .frbw, .npz, .pt/.pth/.bin)")
    check_complete(model, w)
    return w
