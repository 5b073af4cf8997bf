"""Size-independent properties at BASELINE.json's full sizes (configs[1], [3], [4]), where the CPU oracle would take
minutes to hours: batch-composition invariance and frame-permutation equivariance of detect at 1080p / ResNet50,
unit-norm and position-independent embeddings at 512 faces, pdist spot checks + metric properties at N = 50 000,
sorted / in-range / spot-checked kNN at 50 000 x 1 000 000, monotone valid linkage from the GPU NN-chain."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _detect_rows(det, frames_dev):
    return det.detect_batch_raw(frames_dev)


def test_r50_1080p_batch_invariance_and_permutation(ctx):
    """The detections of a frame do not depend on what else is in the batch or where it sits in it: same kept rows,
    bit-identical floats (every output pixel is the same sequence of MMAs whatever tile it lands in)."""
    from facerec_b200 import synth
    from facerec_b200.detector import B200RetinaFaceDetector
    det = B200RetinaFaceDetector(20, "r50", ctx=ctx)
    frames = torch.from_numpy(synth.clip(1234, 40, 8, 1080, 1920)).cuda()
    whole = _detect_rows(det, frames)
    assert sum(len(r) for r in whole) >= 8                      # not vacuous
    rev = _detect_rows(det, frames.flip(0).contiguous())
    for i in range(8):
        assert np.array_equal(whole[i], rev[7 - i]), i
    for i in (0, 5):
        alone = _detect_rows(det, frames[i:i + 1].contiguous())[0]
        assert np.array_equal(whole[i], alone), i
    three = _detect_rows(det, frames[2:5].contiguous())
    for j in range(3):
        assert np.array_equal(whole[2 + j], three[j]), j


def test_irv1_512_faces_unit_norm_and_position_independent(ctx):
    from facerec_b200 import ops, weights as W
    ops.load_model(ctx, W.MODEL_IDS["irv1"], W.write_blob(W.synth_weights("irv1", 0)))
    g = torch.Generator().manual_seed(7)
    faces = torch.zeros(512, 160, 160, 8)
    faces[..., :3] = torch.randn(512, 160, 160, 3, generator=g) * 0.5
    faces[300] = faces[3]                                       # duplicates far apart in the batch
    faces[511] = faces[3]
    e = ops.irv1_forward(ctx, faces.half().cuda()).cpu().numpy()
    assert e.shape == (512, 512) and np.isfinite(e).all()
    assert np.abs(np.linalg.norm(e.astype(np.float64), axis=1) - 1).max() < 1e-5
    assert np.array_equal(e[3], e[300]) and np.array_equal(e[3], e[511])
    small = ops.irv1_forward(ctx, faces[:8].half().cuda()).cpu().numpy()
    assert np.array_equal(small[3], e[3])                       # batch of 8 vs batch of 512: same bits


def _unit_rows(n, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


def test_pdist_50k_properties(ctx):
    """configs[3]: 50 000 x 512.  Condensed float32 (5 GB): spot-check 20 000 random pairs against float64, duplicates
    at distance 0, triangle inequality on random triples, and the square layout of a 4 096-row slice is symmetric."""
    from facerec_b200 import ops, _lib
    n, d = 50_000, 512
    x = _unit_rows(n, d, 11)
    x[12345] = x[77]
    x[49999] = x[77]
    cond = ops.pdist(ctx, x, layout=_lib.FRB_CONDENSED_F32)
    assert cond.numel() == n * (n - 1) // 2

    def at(i, j):
        i, j = torch.minimum(i, j), torch.maximum(i, j)
        return cond[n * i - i * (i + 1) // 2 + (j - i - 1)]

    g = torch.Generator(device="cuda").manual_seed(5)
    i = torch.randint(0, n, (20_000,), generator=g, device="cuda")
    j = torch.randint(0, n, (20_000,), generator=g, device="cuda")
    keep = i != j
    i, j = i[keep], j[keep]
    ref = (x[i].double() - x[j].double()).norm(dim=1)
    assert (at(i, j).double() - ref).abs().max().item() < 2e-5
    dup = torch.tensor([77, 77, 12345], device="cuda"), torch.tensor([12345, 49999, 49999], device="cuda")
    # exact duplicates: d^2 = ||a||^2 + ||b||^2 - 2ab carries the dropped lo*lo term of the split-bf16 product
    # (<= 2^-17 relative to ||a|| ||b||), and the square root turns that 8e-6 into 3e-3
    assert at(*dup).abs().max().item() < 5e-3
    k = torch.randint(0, n, (len(i),), generator=g, device="cuda")
    ok = (k != i) & (k != j)
    assert bool((at(i[ok], j[ok]) <= at(i[ok], k[ok]) + at(k[ok], j[ok]) + 1e-5).all())
    del cond
    sq = ops.pdist(ctx, x[:4096].contiguous(), layout=_lib.FRB_SQUARE_SYM_F32)
    assert torch.equal(sq, sq.t()) and float(sq.diagonal().abs().max()) == 0.0


def test_knn_50k_by_1m_properties(ctx):
    """configs[4]: 50 000 queries against 1 000 000 database rows, k = 10.  Distances ascending, indices in range and
    unique per row, planted exact matches found first, spot-checked rows equal a float64 brute force."""
    from facerec_b200 import ops
    q, m, d, k = 50_000, 1_000_000, 512, 10
    db = _unit_rows(m, d, 21)
    qs = _unit_rows(q, d, 22)
    planted = torch.arange(0, q, 997, device="cuda")
    qs[planted] = db[(planted * 13) % m]
    idx, dist = ops.knn_topk(ctx, qs, db, k)
    assert idx.shape == (q, k) and dist.shape == (q, k)
    assert bool((idx >= 0).all()) and bool((idx < m).all())
    assert bool((dist[:, 1:] >= dist[:, :-1]).all())
    assert bool((idx.sort(dim=1).values.diff(dim=1) != 0).all())
    assert torch.equal(idx[planted, 0].long(), (planted * 13) % m) and float(dist[planted, 0].max()) < 5e-3
    rows = torch.tensor([1, 2, 4242, 31337, 49_999], device="cuda")     # not planted (those sit at the near-zero floor above)
    full = torch.cdist(qs[rows].double(), db.double())
    ref_d, ref_i = full.topk(k, dim=1, largest=False)
    assert torch.equal(idx[rows].long(), ref_i)
    assert (dist[rows].double() - ref_d).abs().max().item() < 2e-5


def test_gpu_linkage_20k_valid_and_monotone(ctx):
    """NN-chain complete linkage straight from the GPU distance matrix: a valid scipy linkage matrix with
    non-decreasing merge heights and cluster sizes that add up."""
    from scipy.cluster import hierarchy
    from facerec_b200.distance import gpu_linkage_complete
    n = 20_000
    x = _unit_rows(n, 512, 31)
    Z = gpu_linkage_complete(x.cpu().numpy(), ctx=ctx)
    assert Z.shape == (n - 1, 4) and hierarchy.is_valid_linkage(Z)
    assert bool((np.diff(Z[:, 2]) >= 0).all()) and Z[-1, 3] == n
    labels = hierarchy.fcluster(Z, 0.9 * Z[-1, 2], criterion="distance")
    assert 1 < labels.max() < n
