"""Architecture tables, seeded synthetic weights and the FRBW weight-blob format.

The reference never ships weights (they are downloaded by third-party packages at run time:
`FaceNet(key)` at facerec/extract.py:392, `RetinaFace(quality="normal")` at facerec/detector.py:64)
and there is no network here, so benchmarks and parity tests use *seeded random-init weights of the
named architectures*.  This module is the single source of truth for

  * the convolution list of each model (name, Cout, kh, kw, Cin, following activation),
  * a deterministic numpy weight factory (He-scaled convs, randomised BatchNorm statistics folded
    into the conv, damped residual branches so activations stay O(1) through ~100 layers,
    class-head gain/bias constants taken from `calib.json`),
  * the blob the C ABI loads (`frb_load_model`): little-endian
        "FRBW" u32 version=1  u32 n_entries  u64 data_offset
        n x { u16 name_len, name, u32 ndim, u32 dims[ndim], u64 first_float }
        float32 data
    with one "<layer>.w" [Cout,kh,kw,Cin] and one "<layer>.b" [Cout] per convolution, BatchNorm
    already folded.

Real checkpoints (e.g. converted PyTorch state dicts) go through `fold_bn` + `write_blob` too.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, NamedTuple, Tuple

import numpy as np

MODEL_IDS = {"retina_mnet025": 0, "retina_r50": 1, "irv1": 2}

_HERE = os.path.dirname(os.path.abspath(__file__))
CALIB_PATH = os.path.join(_HERE, "calib.json")


class ConvSpec(NamedTuple):
    name: str
    cout: int
    kh: int
    kw: int
    cin: int
    act: str          # "relu" | "leaky" | "none"
    bn: bool = True   # conv(no bias)+BN, else conv with bias
    eps: float = 1e-5
    gain: float = 1.0  # extra multiplier on the He scale (residual damping, input normalisation)
    depthwise: bool = False


MNET_STRIDES = [1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1]
MNET_COUT = [16, 32, 32, 64, 64, 128, 128, 128, 128, 128, 128, 256, 256]
R50_BLOCKS = [3, 4, 6, 3]
R50_PLANES = [64, 128, 256, 512]


def retina_specs(backbone: str) -> List[ConvSpec]:
    """RetinaFace conv list (public PyTorch formulation)."""
    s: List[ConvSpec] = []
    if backbone == "mnet025":
        # input is pixel - mean (|x| up to ~150): the first conv carries the 1/64 normalisation
        s.append(ConvSpec("body.0", 8, 3, 3, 3, "leaky", gain=1 / 64))
        cin = 8
        for i, cout in enumerate(MNET_COUT, start=1):
            s.append(ConvSpec(f"body.{i}.dw", cin, 3, 3, 1, "leaky", depthwise=True))
            s.append(ConvSpec(f"body.{i}.pw", cout, 1, 1, cin, "leaky"))
            cin = cout
        in_ch, out_ch, neck = [64, 128, 256], 64, "leaky"
    elif backbone == "r50":
        s.append(ConvSpec("conv1", 64, 7, 7, 3, "relu", gain=1 / 64))
        cin = 64
        for l, (nb, planes) in enumerate(zip(R50_BLOCKS, R50_PLANES), start=1):
            for i in range(nb):
                p = f"layer{l}.{i}"
                s.append(ConvSpec(p + ".conv1", planes, 1, 1, cin, "relu"))
                s.append(ConvSpec(p + ".conv2", planes, 3, 3, planes, "relu"))
                # damped residual branch keeps the trunk variance O(1) over 16 blocks
                s.append(ConvSpec(p + ".conv3", planes * 4, 1, 1, planes, "none", gain=0.35))
                if i == 0:
                    s.append(ConvSpec(p + ".down", planes * 4, 1, 1, cin, "none", gain=0.8))
                cin = planes * 4
        in_ch, out_ch, neck = [512, 1024, 2048], 256, "relu"
    else:
        raise ValueError(backbone)
    for k in (1, 2, 3):
        s.append(ConvSpec(f"fpn.out{k}", out_ch, 1, 1, in_ch[k - 1], neck))
    for k in (1, 2):
        s.append(ConvSpec(f"fpn.merge{k}", out_ch, 3, 3, out_ch, neck))
    q = out_ch // 4
    for k in (1, 2, 3):
        s.append(ConvSpec(f"ssh{k}.c3", 2 * q, 3, 3, out_ch, "none"))
        s.append(ConvSpec(f"ssh{k}.c5_1", q, 3, 3, out_ch, neck))
        s.append(ConvSpec(f"ssh{k}.c5_2", q, 3, 3, q, "none"))
        s.append(ConvSpec(f"ssh{k}.c7_2", q, 3, 3, q, neck))
        s.append(ConvSpec(f"ssh{k}.c7_3", q, 3, 3, q, "none"))
    for k in (1, 2, 3):
        s.append(ConvSpec(f"head{k}.cls", 4, 1, 1, out_ch, "none", bn=False))
        s.append(ConvSpec(f"head{k}.box", 8, 1, 1, out_ch, "none", bn=False))
        s.append(ConvSpec(f"head{k}.ldm", 20, 1, 1, out_ch, "none", bn=False))
    return s


def irv1_specs() -> List[ConvSpec]:
    """FaceNet InceptionResnetV1 conv list; BasicConv2d = conv + BN(eps 1e-3) + ReLU."""
    e = 1e-3
    s: List[ConvSpec] = []

    def bc(name, cout, kh, kw, cin):
        s.append(ConvSpec(name, cout, kh, kw, cin, "relu", eps=e))

    bc("conv2d_1a", 32, 3, 3, 3)
    bc("conv2d_2a", 32, 3, 3, 32)
    bc("conv2d_2b", 64, 3, 3, 32)
    bc("conv2d_3b", 80, 1, 1, 64)
    bc("conv2d_4a", 192, 3, 3, 80)
    bc("conv2d_4b", 256, 3, 3, 192)
    for i in range(5):
        p = f"repeat_1.{i}"
        bc(p + ".branch0", 32, 1, 1, 256)
        bc(p + ".branch1.0", 32, 1, 1, 256)
        bc(p + ".branch1.1", 32, 3, 3, 32)
        bc(p + ".branch2.0", 32, 1, 1, 256)
        bc(p + ".branch2.1", 32, 3, 3, 32)
        bc(p + ".branch2.2", 32, 3, 3, 32)
        s.append(ConvSpec(p + ".conv2d", 256, 1, 1, 96, "none", bn=False))
    bc("mixed_6a.branch0", 384, 3, 3, 256)
    bc("mixed_6a.branch1.0", 192, 1, 1, 256)
    bc("mixed_6a.branch1.1", 192, 3, 3, 192)
    bc("mixed_6a.branch1.2", 256, 3, 3, 192)
    for i in range(10):
        p = f"repeat_2.{i}"
        bc(p + ".branch0", 128, 1, 1, 896)
        bc(p + ".branch1.0", 128, 1, 1, 896)
        bc(p + ".branch1.1", 128, 1, 7, 128)
        bc(p + ".branch1.2", 128, 7, 1, 128)
        s.append(ConvSpec(p + ".conv2d", 896, 1, 1, 256, "none", bn=False))
    bc("mixed_7a.branch0.0", 256, 1, 1, 896)
    bc("mixed_7a.branch0.1", 384, 3, 3, 256)
    bc("mixed_7a.branch1.0", 256, 1, 1, 896)
    bc("mixed_7a.branch1.1", 256, 3, 3, 256)
    bc("mixed_7a.branch2.0", 256, 1, 1, 896)
    bc("mixed_7a.branch2.1", 256, 3, 3, 256)
    bc("mixed_7a.branch2.2", 256, 3, 3, 256)
    for p in [f"repeat_3.{i}" for i in range(5)] + ["block8"]:
        bc(p + ".branch0", 192, 1, 1, 1792)
        bc(p + ".branch1.0", 192, 1, 1, 1792)
        bc(p + ".branch1.1", 192, 1, 3, 192)
        bc(p + ".branch1.2", 192, 3, 1, 192)
        s.append(ConvSpec(p + ".conv2d", 1792, 1, 1, 384, "none", bn=False))
    # Linear(1792 -> 512, no bias) + BatchNorm1d(512, eps 1e-3), stored as a 1x1 conv
    s.append(ConvSpec("last_linear", 512, 1, 1, 1792, "none", eps=e))
    return s


def model_specs(model: str) -> List[ConvSpec]:
    if model == "retina_mnet025":
        return retina_specs("mnet025")
    if model == "retina_r50":
        return retina_specs("r50")
    if model == "irv1":
        return irv1_specs()
    raise ValueError(f"unknown model {model!r}")


def fold_bn(w: np.ndarray, gamma, beta, mean, var, eps: float) -> Tuple[np.ndarray, np.ndarray]:
    """conv(no bias) followed by BatchNorm (eval) -> conv with bias.  w is [Cout, kh, kw, Cin]."""
    scale = (gamma / np.sqrt(var + eps)).astype(np.float32)
    return (w * scale[:, None, None, None]).astype(np.float32), (beta - mean * scale).astype(np.float32)


def load_calib() -> dict:
    if os.path.exists(CALIB_PATH):
        with open(CALIB_PATH) as f:
            return json.load(f)
    return {}


def synth_weights(model: str, seed: int = 0, calib: dict | None = None) -> Dict[str, np.ndarray]:
    """Deterministic random-init weights for `model`: {"<layer>.w": [Cout,kh,kw,Cin], "<layer>.b": [Cout]}."""
    calib = load_calib() if calib is None else calib
    c = calib.get(model, {})
    cls_gain = float(c.get("cls_gain", 1.0))
    cls_bias = float(c.get("cls_bias", 0.0))
    out: Dict[str, np.ndarray] = {}
    for li, sp in enumerate(model_specs(model)):
        rng = np.random.RandomState((seed * 1000003 + li * 7919 + 12345) % (2 ** 31 - 1))
        fan_in = sp.kh * sp.kw * sp.cin
        if sp.act == "relu":
            g = np.sqrt(2.0)
        elif sp.act == "leaky":
            g = np.sqrt(2.0 / (1.0 + 0.1 ** 2))
        else:
            g = 1.0
        std = g * sp.gain / np.sqrt(fan_in)
        w = (rng.standard_normal((sp.cout, sp.kh, sp.kw, sp.cin)) * std).astype(np.float32)
        if sp.bn:
            gamma = rng.uniform(0.8, 1.2, sp.cout).astype(np.float32)
            beta = (rng.standard_normal(sp.cout) * 0.05).astype(np.float32)
            mean = (rng.standard_normal(sp.cout) * 0.05).astype(np.float32)
            var = rng.uniform(0.8, 1.2, sp.cout).astype(np.float32)
            w, b = fold_bn(w, gamma, beta, mean, var, sp.eps)
        else:
            b = (rng.standard_normal(sp.cout) * 0.05).astype(np.float32)
        if sp.name.endswith(".cls"):
            # channel order per anchor: (background, face); a wide logit gap with a calibrated shift keeps
            # the 0.95 score threshold (facerec/detector.py:68) away from the bulk of the priors
            w = (w * cls_gain).astype(np.float32)
            b = (b * 0).astype(np.float32)
            b[1::2] += cls_bias
        out[sp.name + ".w"] = w
        out[sp.name + ".b"] = b
    return out


def write_blob(weights: Dict[str, np.ndarray]) -> bytes:
    names = list(weights.keys())
    table = bytearray()
    off = 0
    for n in names:
        a = np.ascontiguousarray(weights[n], dtype=np.float32)
        nb = n.encode()
        table += struct.pack("<H", len(nb)) + nb + struct.pack("<I", a.ndim)
        table += struct.pack(f"<{a.ndim}I", *a.shape) + struct.pack("<Q", off)
        off += a.size
    data_off = (20 + len(table) + 15) // 16 * 16
    head = b"FRBW" + struct.pack("<IIQ", 1, len(names), data_off)
    buf = bytearray(head) + table
    buf += b"\0" * (data_off - len(buf))
    for n in names:
        buf += np.ascontiguousarray(weights[n], dtype=np.float32).tobytes()
    return bytes(buf)


def read_blob(blob: bytes) -> Dict[str, np.ndarray]:
    if blob[:4] != b"FRBW":
        raise ValueError("bad magic")
    version, n, data_off = struct.unpack_from("<IIQ", blob, 4)
    if version != 1:
        raise ValueError("unsupported blob version")
    pos = 20
    out: Dict[str, np.ndarray] = {}
    for _ in range(n):
        (nl,) = struct.unpack_from("<H", blob, pos)
        pos += 2
        name = blob[pos:pos + nl].decode()
        pos += nl
        (nd,) = struct.unpack_from("<I", blob, pos)
        pos += 4
        dims = struct.unpack_from(f"<{nd}I", blob, pos)
        pos += 4 * nd
        (eo,) = struct.unpack_from("<Q", blob, pos)
        pos += 8
        cnt = int(np.prod(dims)) if nd else 1
        out[name] = np.frombuffer(blob, dtype="<f4", count=cnt, offset=data_off + eo * 4).reshape(dims).copy()
    return out
