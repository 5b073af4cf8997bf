"""fp32 PyTorch restatement of the networks named by BASELINE.json north_star.

Functional style: weights come from the same FRBW blob the CUDA library loads (BatchNorm already
folded, see facerec_b200/weights.py), so both sides compute conv(x, w) + b with identical numbers.
Reference call sites these stand in for: RetinaFace.predict network part (facerec/detector.py:64,68)
and FaceNet(key).embeddings model.predict (facerec/extract.py:166,392).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

MNET_STRIDES = [1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1]
R50_BLOCKS = [3, 4, 6, 3]


def params_from_weights(weights: Dict[str, np.ndarray]) -> Dict[str, tuple]:
    """{"x.w": [Cout,kh,kw,Cin], "x.b": [Cout]} -> {"x": (OIHW tensor, bias tensor)}"""
    P = {}
    for k, v in weights.items():
        if k.endswith(".w"):
            n = k[:-2]
            w = torch.from_numpy(np.ascontiguousarray(v.transpose(0, 3, 1, 2))).float()
            P[n] = (w, torch.from_numpy(weights[n + ".b"]).float())
    return P


def _act(y, act, slope=0.1):
    if act == "relu":
        return F.relu(y)
    if act == "leaky":
        return F.leaky_relu(y, slope)
    return y


def _conv(x, P, name, stride=1, padding=0, act="none", groups=1):
    w, b = P[name]
    return _act(F.conv2d(x, w, b, stride=stride, padding=padding, groups=groups), act)


def r50_features(P, x):
    """torchvision-style ResNet-50 body: outputs of layer2, layer3, layer4 (C = 512 / 1024 / 2048 at
    stride 8 / 16 / 32).  Pinned against torchvision.models.resnet50 itself in tests/test_oracle_cpu.py."""
    y = _conv(x, P, "conv1", 2, 3, "relu")
    y = F.max_pool2d(y, 3, 2, 1)
    feats = []
    for l, nb in enumerate(R50_BLOCKS, start=1):
        for i in range(nb):
            s = 2 if (i == 0 and l > 1) else 1
            p = f"layer{l}.{i}"
            o = _conv(y, P, p + ".conv1", 1, 0, "relu")
            o = _conv(o, P, p + ".conv2", s, 1, "relu")   # torchvision v1.5: stride on the 3x3
            o = _conv(o, P, p + ".conv3", 1, 0, "none")
            idn = _conv(y, P, p + ".down", s, 0, "none") if i == 0 else y
            y = F.relu(o + idn)
        if l >= 2:
            feats.append(y)
    return feats


def retina_forward(P, x, backbone: str):
    """x: [B,3,H,W] float32, BGR minus (104,117,123).  Returns loc [B,P,4], conf logits [B,P,2], landm [B,P,10]."""
    if backbone == "mnet025":
        neck = "leaky"
        y = _conv(x, P, "body.0", 2, 1, "leaky")
        feats = []
        for i, s in enumerate(MNET_STRIDES, start=1):
            y = _conv(y, P, f"body.{i}.dw", s, 1, "leaky", groups=y.shape[1])
            y = _conv(y, P, f"body.{i}.pw", 1, 0, "leaky")
            if i in (5, 11, 13):
                feats.append(y)
    elif backbone == "r50":
        neck = "relu"
        feats = r50_features(P, x)
    else:
        raise ValueError(backbone)
    o1 = _conv(feats[0], P, "fpn.out1", 1, 0, neck)
    o2 = _conv(feats[1], P, "fpn.out2", 1, 0, neck)
    o3 = _conv(feats[2], P, "fpn.out3", 1, 0, neck)
    up3 = F.interpolate(o3, size=o2.shape[2:], mode="nearest")
    o2 = _conv(o2 + up3, P, "fpn.merge2", 1, 1, neck)
    up2 = F.interpolate(o2, size=o1.shape[2:], mode="nearest")
    o1 = _conv(o1 + up2, P, "fpn.merge1", 1, 1, neck)
    locs, confs, lands = [], [], []
    for k, f in enumerate((o1, o2, o3), start=1):
        c3 = _conv(f, P, f"ssh{k}.c3", 1, 1)
        c51 = _conv(f, P, f"ssh{k}.c5_1", 1, 1, neck)
        c5 = _conv(c51, P, f"ssh{k}.c5_2", 1, 1)
        c72 = _conv(c51, P, f"ssh{k}.c7_2", 1, 1, neck)
        c7 = _conv(c72, P, f"ssh{k}.c7_3", 1, 1)
        s = F.relu(torch.cat([c3, c5, c7], 1))
        B = s.shape[0]
        confs.append(_conv(s, P, f"head{k}.cls").permute(0, 2, 3, 1).reshape(B, -1, 2))
        locs.append(_conv(s, P, f"head{k}.box").permute(0, 2, 3, 1).reshape(B, -1, 4))
        lands.append(_conv(s, P, f"head{k}.ldm").permute(0, 2, 3, 1).reshape(B, -1, 10))
    return torch.cat(locs, 1), torch.cat(confs, 1), torch.cat(lands, 1)


def _block35(x, P, p):
    b0 = _conv(x, P, p + ".branch0", act="relu")
    b1 = _conv(_conv(x, P, p + ".branch1.0", act="relu"), P, p + ".branch1.1", 1, 1, "relu")
    b2 = _conv(x, P, p + ".branch2.0", act="relu")
    b2 = _conv(_conv(b2, P, p + ".branch2.1", 1, 1, "relu"), P, p + ".branch2.2", 1, 1, "relu")
    out = _conv(torch.cat([b0, b1, b2], 1), P, p + ".conv2d")
    return F.relu(out * 0.17 + x)


def _block17(x, P, p):
    b0 = _conv(x, P, p + ".branch0", act="relu")
    b1 = _conv(x, P, p + ".branch1.0", act="relu")
    b1 = _conv(b1, P, p + ".branch1.1", 1, (0, 3), "relu")
    b1 = _conv(b1, P, p + ".branch1.2", 1, (3, 0), "relu")
    out = _conv(torch.cat([b0, b1], 1), P, p + ".conv2d")
    return F.relu(out * 0.10 + x)


def _block8(x, P, p, scale=0.20, relu=True):
    b0 = _conv(x, P, p + ".branch0", act="relu")
    b1 = _conv(x, P, p + ".branch1.0", act="relu")
    b1 = _conv(b1, P, p + ".branch1.1", 1, (0, 1), "relu")
    b1 = _conv(b1, P, p + ".branch1.2", 1, (1, 0), "relu")
    out = _conv(torch.cat([b0, b1], 1), P, p + ".conv2d") * scale + x
    return F.relu(out) if relu else out


def irv1_forward(P, x, normalize: bool = True):
    """x: [F,3,160,160] float32 RGB, already (x-127.5)/128.  Returns [F,512] (L2-normalised)."""
    y = _conv(x, P, "conv2d_1a", 2, 0, "relu")
    y = _conv(y, P, "conv2d_2a", 1, 0, "relu")
    y = _conv(y, P, "conv2d_2b", 1, 1, "relu")
    y = F.max_pool2d(y, 3, 2)
    y = _conv(y, P, "conv2d_3b", 1, 0, "relu")
    y = _conv(y, P, "conv2d_4a", 1, 0, "relu")
    y = _conv(y, P, "conv2d_4b", 2, 0, "relu")
    for i in range(5):
        y = _block35(y, P, f"repeat_1.{i}")
    b0 = _conv(y, P, "mixed_6a.branch0", 2, 0, "relu")
    b1 = _conv(y, P, "mixed_6a.branch1.0", act="relu")
    b1 = _conv(b1, P, "mixed_6a.branch1.1", 1, 1, "relu")
    b1 = _conv(b1, P, "mixed_6a.branch1.2", 2, 0, "relu")
    y = torch.cat([b0, b1, F.max_pool2d(y, 3, 2)], 1)
    for i in range(10):
        y = _block17(y, P, f"repeat_2.{i}")
    b0 = _conv(_conv(y, P, "mixed_7a.branch0.0", act="relu"), P, "mixed_7a.branch0.1", 2, 0, "relu")
    b1 = _conv(_conv(y, P, "mixed_7a.branch1.0", act="relu"), P, "mixed_7a.branch1.1", 2, 0, "relu")
    b2 = _conv(_conv(y, P, "mixed_7a.branch2.0", act="relu"), P, "mixed_7a.branch2.1", 1, 1, "relu")
    b2 = _conv(b2, P, "mixed_7a.branch2.2", 2, 0, "relu")
    y = torch.cat([b0, b1, b2, F.max_pool2d(y, 3, 2)], 1)
    for i in range(5):
        y = _block8(y, P, f"repeat_3.{i}")
    y = _block8(y, P, "block8", 1.0, False)
    y = F.adaptive_avg_pool2d(y, 1)
    y = _conv(y, P, "last_linear").flatten(1)          # Linear + BatchNorm1d folded
    return F.normalize(y, p=2, dim=1) if normalize else y
