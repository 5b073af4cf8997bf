"""K2 parity: the tcgen05 implicit-GEMM convolution against a plain PyTorch fp32 conv of the same
16-bit-rounded operands (tolerance: one fp16 output rounding + fp32 accumulation-order noise)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x16, w, b, stride, pad, act, slope, res16=None, res_mode=0):
    x = x16.float().cpu().permute(0, 3, 1, 2)
    wt = torch.from_numpy(w).half().float().permute(0, 3, 1, 2)
    y = F.conv2d(x, wt, None if b is None else torch.from_numpy(b), stride=stride, padding=pad)
    if res_mode == 1:
        y = y + res16.float().cpu().permute(0, 3, 1, 2)
    if act == 1:
        y = F.relu(y)
    elif act == 2:
        y = F.leaky_relu(y, slope)
    if res_mode == 2:
        r = res16.float().cpu().permute(0, 3, 1, 2)
        y = y + F.interpolate(r, size=y.shape[2:], mode="nearest")
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # B, H, W, Cin, Cout, kh, kw, stride, pad, act, res_mode, out_fp32
    (2, 20, 24, 64, 128, 1, 1, (1, 1), (0, 0), 1, 0, False),     # plain GEMM, TMA A path
    (1, 17, 17, 96, 32, 1, 1, (1, 1), (0, 0), 0, 0, False),      # K tail (96 -> 128), small N
    (2, 19, 23, 32, 64, 3, 3, (1, 1), (1, 1), 2, 0, False),      # 3x3 pad 1, leaky
    (2, 36, 40, 8, 8, 3, 3, (2, 2), (1, 1), 2, 0, False),        # MobileNet first conv (Cin padded to 8)
    (1, 64, 80, 8, 64, 7, 7, (2, 2), (3, 3), 1, 0, False),       # ResNet stem
    (2, 8, 8, 128, 128, 1, 7, (1, 1), (0, 3), 1, 0, False),      # Block17 1x7
    (2, 8, 8, 128, 128, 7, 1, (1, 1), (3, 0), 1, 0, False),      # Block17 7x1
    (1, 36, 36, 192, 256, 3, 3, (2, 2), (0, 0), 1, 0, False),    # stride 2, no pad, two N tiles
    (2, 17, 17, 96, 256, 1, 1, (1, 1), (0, 0), 1, 1, False),     # residual before activation
    (1, 23, 40, 128, 64, 1, 1, (1, 1), (0, 0), 2, 2, False),     # FPN lateral + nearest-upsampled top-down
    (3, 1, 1, 1792, 512, 1, 1, (1, 1), (0, 0), 0, 0, True),      # last_linear, fp32 out
    (4, 68, 120, 64, 64, 3, 3, (1, 1), (1, 1), 1, 0, False),     # > 148 tiles: persistent loop, TMEM double buffer
    (1, 30, 30, 256, 512, 1, 1, (2, 2), (0, 0), 0, 0, False),    # strided 1x1 (ResNet downsample): strided patch boxes
    (2, 21, 37, 64, 128, 3, 3, (1, 1), (1, 1), 1, 1, False),     # patch tiling + TMA-staged residual, ragged borders
    (5, 8, 8, 128, 384, 3, 3, (1, 1), (1, 1), 1, 0, False),      # patches spanning images (8x8 maps), two 192-wide N tiles
    (30, 3, 3, 192, 192, 1, 3, (1, 1), (0, 1), 1, 0, False),     # Block8 1x3 on 3x3 maps: 14 images per patch
    (3, 17, 17, 192, 256, 3, 3, (2, 2), (0, 0), 1, 0, False),    # Mixed_6a stride-2 reduction 17 -> 8
    (2, 40, 56, 256, 80, 1, 1, (1, 1), (0, 0), 2, 0, False),     # Cout = 80: partial last store slab
    (1, 135, 240, 64, 256, 1, 1, (1, 1), (0, 0), 1, 1, False),   # ResNet conv3 + skip, 254 M-tiles x 2 N-tiles
    (2, 48, 64, 128, 128, 3, 3, (1, 1), (1, 1), 1, 1, False),    # halo mode: 2 channel blocks, streamed weights, residual
    (1, 47, 61, 64, 64, 3, 3, (1, 1), (1, 1), 2, 0, False),      # halo mode, ragged in both dims, resident weights
    (1, 32, 64, 128, 128, 1, 7, (1, 1), (0, 3), 1, 0, False),    # halo mode with a 1x7 filter (14 x 16 halo)
    (1, 32, 48, 256, 256, 3, 3, (1, 1), (1, 1), 1, 0, False),    # halo mode: 4 channel blocks x 9 taps, BN = 256
    (3, 270, 480, 64, 64, 3, 3, (1, 1), (1, 1), 1, 0, False),    # ResNet layer1 conv2 at full 1080p/4 size
    (2, 40, 44, 128, 96, 7, 1, (1, 1), (3, 0), 1, 0, False),     # column filter: aligned halo tiles by default, 2 channel blocks
    (1, 50, 40, 64, 64, 4, 1, (1, 1), (2, 0), 2, 0, False),      # 4 x 1 window (the shape of the s2d stem), resident weights
    (4, 68, 120, 128, 512, 1, 1, (1, 1), (0, 0), 1, 1, False),   # ResNet conv3 + skip with a resident weight slice per CTA
    (4, 68, 120, 128, 576, 1, 1, (1, 1), (0, 0), 2, 0, False),   # no skip: two k-blocks per stage, three 192-wide N tiles, 147 CTAs
    (4, 68, 120, 128, 576, 1, 1, (1, 1), (0, 0), 1, 1, False),   # nine 64-wide N tiles: grid of 144 CTAs
    # CTA pairs (cta_group::2): per-tap / linear path, streamed weights, >= 74 (M pair, N tile) units
    (3, 68, 120, 1024, 256, 1, 1, (1, 1), (0, 0), 1, 0, False),  # layer3 conv1: 16 k-blocks, two per stage, BN = 256 split 128 + 128
    (3, 68, 120, 256, 1024, 1, 1, (1, 1), (0, 0), 1, 1, False),  # layer3 conv3 + skip: residual slabs x identity halves, 8 N tiles
    (4, 135, 240, 128, 128, 3, 3, (2, 2), (1, 1), 1, 0, False),  # layer2.0 conv2: strided 3x3, patch boxes, 9 taps x 2 blocks
    (2, 135, 240, 256, 512, 1, 1, (2, 2), (0, 0), 0, 0, False),  # layer2.0 downsample: strided 1x1, two N tiles
    (3, 67, 119, 512, 256, 1, 1, (1, 1), (0, 0), 2, 2, False),   # FPN lateral + upsampled top-down, ODD number of M tiles
    (5, 33, 59, 192, 320, 3, 3, (2, 2), (1, 1), 1, 0, False),    # ragged strided patches, 160-wide N tiles (80 + 80), odd M tiles
    (3, 135, 240, 128, 128, 3, 3, (1, 1), (1, 1), 1, 0, False),  # pairs on halo tiles, a filter row of streamed weights per stage
    (2, 135, 240, 256, 256, 3, 3, (1, 1), (1, 1), 1, 0, False),  # pairs on PAIRED halo tiles (FPN merge1): 4 patches per weight stage
    (2, 133, 239, 64, 64, 3, 3, (1, 1), (1, 1), 1, 0, False),    # pairs on halo tiles with resident weights (32 rows per CTA), ragged
]
SLICE_CASES = CASES[-12:-9]
PAIR_CASES = CASES[-9:]


@pytest.mark.parametrize("case", CASES, ids=[f"c{i}" for i in range(len(CASES))])
def test_conv2d_matches_torch(ctx, case):
    from facerec_b200 import ops
    B, H, W, Cin, Cout, kh, kw, stride, pad, act, res_mode, out_fp32 = case
    g = torch.Generator().manual_seed(1234 + Cin * 7 + Cout)
    x16 = (torch.randn(B, H, W, Cin, generator=g)).half().cuda()
    w = (torch.randn(Cout, kh, kw, Cin, generator=g) / np.sqrt(kh * kw * Cin)).numpy().astype(np.float32)
    b = (torch.randn(Cout, generator=g) * 0.1).numpy().astype(np.float32)
    Ho = (H + 2 * pad[0] - kh) // stride[0] + 1
    Wo = (W + 2 * pad[1] - kw) // stride[1] + 1
    res16 = None
    if res_mode == 1:
        res16 = torch.randn(B, Ho, Wo, Cout, generator=g).half().cuda()
    elif res_mode == 2:
        res16 = torch.randn(B, (Ho + 1) // 2, (Wo + 1) // 2, Cout, generator=g).half().cuda()
    out, _ = ops.conv2d(ctx, x16, w, b, stride, pad, act, 0.1, res16, res_mode, out_fp32)
    torch.cuda.synchronize()
    ref = _ref(x16, w, b, stride, pad, act, 0.1, res16, res_mode)
    got = out.float().cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs()
    tol = 2e-3 * ref.abs() + 2e-3
    assert bool((err <= tol).all()), f"max err {err.max().item():.4g} at ref {ref.flatten()[err.argmax()].item():.4g}"


@pytest.mark.parametrize("pair", [0, 2], ids=["single", "paired"])
@pytest.mark.parametrize("case", [c for c in CASES if c[5] * c[6] > 1 and c[7] == (1, 1) and c[3] % 64 == 0],
                         ids=lambda c: f"{c[1]}x{c[2]}x{c[3]}")
def test_conv2d_halo_mode_matches_torch(case, pair):
    """FRB_OPT_HALO: stride-1 k x k convs fed from one halo tile per channel block (shifted UMMA descriptors);
    FRB_OPT_MPAIR: two patches per halo tile and weight stage where the weights are streamed."""
    from facerec_b200 import ops, _lib
    c = _lib.Context(0)
    try:
        c.set_option(_lib.FRB_OPT_HALO, 1)
        c.set_option(_lib.FRB_OPT_MPAIR, pair)
        B, H, W, Cin, Cout, kh, kw, stride, pad, act, res_mode, out_fp32 = case
        g = torch.Generator().manual_seed(99 + Cin + Cout)
        x16 = torch.randn(B, H, W, Cin, generator=g).half().cuda()
        w = (torch.randn(Cout, kh, kw, Cin, generator=g) / np.sqrt(kh * kw * Cin)).numpy().astype(np.float32)
        b = (torch.randn(Cout, generator=g) * 0.1).numpy().astype(np.float32)
        res16 = torch.randn(B, H, W, Cout, generator=g).half().cuda() if res_mode == 1 else None
        out, _ = ops.conv2d(c, x16, w, b, stride, pad, act, 0.1, res16, res_mode, out_fp32)
        torch.cuda.synchronize()
        ref = _ref(x16, w, b, stride, pad, act, 0.1, res16, res_mode)
        err = (out.float().cpu() - ref).abs()
        assert bool((err <= 2e-3 * ref.abs() + 2e-3).all()), err.max().item()
    finally:
        c.close()


@pytest.mark.parametrize("case", SLICE_CASES, ids=lambda c: f"{c[4]}r{c[10]}")
def test_conv2d_weight_slice_is_bit_identical_to_streaming(ctx, case):
    """FRB_OPT_WEIGHT_SLICE changes where a tile's weights come from, not the order its k-blocks are accumulated in."""
    from facerec_b200 import ops, _lib
    B, H, W, Cin, Cout, kh, kw, stride, pad, act, res_mode, out_fp32 = case
    g = torch.Generator().manual_seed(7 + Cout)
    x16 = torch.randn(B, H, W, Cin, generator=g).half().cuda()
    w = (torch.randn(Cout, kh, kw, Cin, generator=g) / np.sqrt(Cin)).numpy().astype(np.float32)
    res16 = torch.randn(B, H, W, Cout, generator=g).half().cuda() if res_mode == 1 else None
    sliced, _ = ops.conv2d(ctx, x16, w, None, stride, pad, act, 0.1, res16, res_mode)
    c = _lib.Context(0)
    try:
        c.set_option(_lib.FRB_OPT_WEIGHT_SLICE, 0)
        streamed, _ = ops.conv2d(c, x16, w, None, stride, pad, act, 0.1, res16, res_mode)
        torch.cuda.synchronize()
        assert torch.equal(sliced, streamed)
    finally:
        c.close()


def test_conv2d_bf16_option():
    from facerec_b200 import ops, _lib
    c = _lib.Context(0)
    try:
        c.set_option(_lib.FRB_OPT_BF16, 1)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(2, 12, 12, 64, generator=g).bfloat16().cuda()
        w = (torch.randn(96, 3, 3, 64, generator=g) / 24).numpy().astype(np.float32)
        out, _ = ops.conv2d(c, x, w, None, (1, 1), (1, 1), 1, 0.0)
        torch.cuda.synchronize()
        ref = F.relu(F.conv2d(x.float().cpu().permute(0, 3, 1, 2), torch.from_numpy(w).bfloat16().float().permute(0, 3, 1, 2),
                              padding=1)).permute(0, 2, 3, 1)
        err = (out.float().cpu() - ref).abs()
        assert bool((err <= 1e-2 * ref.abs() + 1e-2).all()), err.max().item()
    finally:
        c.close()


def test_conv2d_rejects_bad_shapes(ctx):
    from facerec_b200 import ops, _lib
    x = torch.zeros(1, 8, 8, 12, dtype=torch.float16, device="cuda")   # Cin not a multiple of 8
    with pytest.raises(_lib.FrbError):
        ops.conv2d(ctx, x, np.zeros((16, 1, 1, 12), np.float32), None)


@pytest.mark.parametrize("case", PAIR_CASES, ids=lambda c: f"{c[1]}x{c[2]}x{c[3]}x{c[4]}")
def test_conv2d_cta_pairs_give_the_same_bits(ctx, case):
    """FRB_OPT_CTA2: two CTAs of a TPC share one M = 256 tcgen05.mma and split the weight block; the k-blocks of a tile
    still accumulate in the same order, so the output is bit-identical to the single-CTA launch."""
    from facerec_b200 import ops, _lib
    B, H, W, Cin, Cout, kh, kw, stride, pad, act, res_mode, out_fp32 = case
    g = torch.Generator().manual_seed(77 + Cin + Cout)
    x16 = torch.randn(B, H, W, Cin, generator=g).half().cuda()
    w = (torch.randn(Cout, kh, kw, Cin, generator=g) / np.sqrt(kh * kw * Cin)).numpy().astype(np.float32)
    b = (torch.randn(Cout, generator=g) * 0.1).numpy().astype(np.float32)
    Ho = (H + 2 * pad[0] - kh) // stride[0] + 1
    Wo = (W + 2 * pad[1] - kw) // stride[1] + 1
    res16 = None
    if res_mode == 1:
        res16 = torch.randn(B, Ho, Wo, Cout, generator=g).half().cuda()
    elif res_mode == 2:
        res16 = torch.randn(B, (Ho + 1) // 2, (Wo + 1) // 2, Cout, generator=g).half().cuda()
    single = _lib.Context(0)
    try:
        single.set_option(_lib.FRB_OPT_CTA2, 0)
        a, _ = ops.conv2d(single, x16, w, b, stride, pad, act, 0.1, res16, res_mode, out_fp32)
        c, _ = ops.conv2d(ctx, x16, w, b, stride, pad, act, 0.1, res16, res_mode, out_fp32)
        torch.cuda.synchronize()
        assert torch.equal(a, c)
    finally:
        single.close()
