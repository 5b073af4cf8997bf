"""K12-K15 parity against the libraries the reference calls: scipy pdist/linkage (facerec/cluster.py:41)
and scikit-learn brute kNN (facerec/classify_knn.py:212,227)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mixture(rng, n, d=512, centres=40, noise=0.05, unit=True):
    c = rng.standard_normal((centres, d))
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    x = c[rng.randint(0, centres, n)] + noise * rng.standard_normal((n, d))
    if unit:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


@pytest.mark.parametrize("n,d", [(300, 512), (1000, 512), (130, 100), (2, 512)])
def test_pdist_condensed_matches_scipy(ctx, n, d):
    from scipy.spatial.distance import pdist
    from facerec_b200 import ops, _lib
    rng = np.random.RandomState(n)
    x = _mixture(rng, n, d, unit=False)
    ref = pdist(x.astype(np.float64))
    X = torch.from_numpy(x).cuda()
    got32 = ops.pdist(ctx, X, layout=_lib.FRB_CONDENSED_F32).cpu().numpy()
    got64 = ops.pdist(ctx, X, layout=_lib.FRB_CONDENSED_F64).cpu().numpy()
    assert got32.shape == ref.shape
    np.testing.assert_allclose(got32, ref, rtol=0, atol=3e-5)
    assert np.array_equal(got64.astype(np.float32), got32)
    sq = ops.pdist(ctx, X, layout=_lib.FRB_SQUARE_F32).cpu().numpy()
    from scipy.spatial.distance import squareform
    full = squareform(ref)
    off = ~np.eye(n, dtype=bool)
    np.testing.assert_allclose(sq[off], full[off], rtol=0, atol=3e-5)
    assert not np.diag(sq).any()                      # the kernel pins d(i,i) to exactly 0


def test_pdist_row_sharded_equals_whole(ctx):
    from facerec_b200 import ops, _lib
    rng = np.random.RandomState(5)
    x = _mixture(rng, 700)
    X = torch.from_numpy(x).cuda()
    whole = ops.pdist(ctx, X).cpu().numpy()
    out = torch.zeros_like(torch.from_numpy(whole)).cuda()
    for r0, r in ((0, 100), (100, 333), (433, 267)):
        ops.pdist(ctx, X, r0, r, _lib.FRB_CONDENSED_F32, out)
    assert np.array_equal(out.cpu().numpy(), whole)


def test_linkage_labels_match_reference_call(ctx):
    """cluster.py:41-42 — linkage(vectors, 'complete') + fcluster(maxclust) on GPU distances vs on raw vectors."""
    from scipy.cluster.hierarchy import linkage, fcluster
    from facerec_b200 import ops, _lib
    rng = np.random.RandomState(11)
    n = 1500
    # trajectory means are NOT unit norm (cluster.py:29)
    x = np.stack([_mixture(rng, 1, centres=1)[0] for _ in range(1)])  # warm rng
    cent = rng.standard_normal((80, 512)); cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    xs = []
    for i in range(n):
        L = rng.randint(3, 30)
        v = cent[rng.randint(0, 80)] + 0.05 * rng.standard_normal((L, 512))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        xs.append(v.astype(np.float32).mean(axis=0))
    x = np.stack(xs).astype(np.float32)
    k = n // 18
    ref = fcluster(linkage(x, method="complete"), t=k, criterion="maxclust")
    d = ops.pdist(ctx, torch.from_numpy(x).cuda(), layout=_lib.FRB_CONDENSED_F64).cpu().numpy()
    got = fcluster(linkage(d, method="complete"), t=k, criterion="maxclust")
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("q,m,k", [(257, 5000, 10), (5, 300, 10), (1000, 20000, 10), (64, 129, 16)])
def test_knn_topk_matches_bruteforce(ctx, q, m, k):
    from facerec_b200 import ops
    rng = np.random.RandomState(q + m)
    db = _mixture(rng, m, centres=50)
    qs = _mixture(rng, q, centres=50)
    idx, dist = ops.knn_topk(ctx, torch.from_numpy(qs).cuda(), torch.from_numpy(db).cuda(), k)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    d = np.sqrt(np.maximum(((qs.astype(np.float64)[:, None, :] - db.astype(np.float64)[None, :, :]) ** 2).sum(-1), 0)) \
        if q * m <= 2_000_000 else None
    if d is None:
        qq = (qs.astype(np.float64) ** 2).sum(1)[:, None]
        dd = (db.astype(np.float64) ** 2).sum(1)[None, :]
        d = np.sqrt(np.maximum(qq + dd - 2 * qs.astype(np.float64) @ db.astype(np.float64).T, 0))
    ref_idx = np.argsort(d, axis=1, kind="stable")[:, :k]
    ref_d = np.take_along_axis(d, ref_idx, 1)
    np.testing.assert_allclose(dist, ref_d, rtol=0, atol=3e-5)
    # index sets agree except where two candidates are closer than the fp32 resolution
    same = (np.sort(idx, 1) == np.sort(ref_idx, 1)).all(1)
    for r in np.nonzero(~same)[0]:
        gap = np.abs(np.take_along_axis(d[r][None], idx[r][None].astype(np.int64), 1) - ref_d[r]).max()
        assert gap < 3e-5
    assert same.mean() > 0.99
    assert (idx[:, 0] == ref_idx[:, 0]).mean() > 0.995          # kNN top-1


def test_knn_tight_clusters_take_the_exact_pass(ctx):
    """Database rows closer to each other than the bf16 rounding bound: the one-pass candidates cannot be proven, the
    queries are re-run through the exact split-bf16 pass and still match the fp64 brute force."""
    from facerec_b200 import ops
    rng = np.random.RandomState(77)
    db = _mixture(rng, 3000, centres=20, noise=0.0005)
    qs = _mixture(rng, 300, centres=20, noise=0.0005)
    idx, dist = ops.knn_topk(ctx, torch.from_numpy(qs).cuda(), torch.from_numpy(db).cuda(), 10)
    assert ops.knn_redo_count(ctx) > 200                       # (almost) every query was undecidable in bf16
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    d = np.sqrt(np.maximum(((qs.astype(np.float64)[:, None, :] - db.astype(np.float64)[None, :, :]) ** 2).sum(-1), 0))
    ref_idx = np.argsort(d, axis=1, kind="stable")[:, :10]
    ref_d = np.take_along_axis(d, ref_idx, 1)
    np.testing.assert_allclose(dist, ref_d, rtol=0, atol=3e-5)
    got_d = np.take_along_axis(d, idx.astype(np.int64), 1)
    assert np.abs(got_d - ref_d).max() < 3e-5                  # the same neighbours up to fp32-resolution ties
    # well separated data: nothing needs the exact pass
    db2 = rng.standard_normal((3000, 512)).astype(np.float32)
    db2 /= np.linalg.norm(db2, axis=1, keepdims=True)
    ops.knn_topk(ctx, torch.from_numpy(_mixture(rng, 300, centres=50)).cuda(), torch.from_numpy(db2).cuda(), 10)
    assert ops.knn_redo_count(ctx) < 60


def test_knn_fit_reuses_the_prepared_database(ctx):
    """frb_knn_fit (KNeighborsClassifier.fit): later frb_knn_topk calls with the same database tensor skip its preparation
    and give the same bits; another database, or the same one after a new fit, is never confused with it."""
    from facerec_b200 import ops
    rng = np.random.RandomState(21)
    db = torch.from_numpy(_mixture(rng, 7000, centres=60)).cuda()
    db2 = torch.from_numpy(_mixture(rng, 7000, centres=60)).cuda()
    q = torch.from_numpy(_mixture(rng, 400, centres=60)).cuda()
    plain = [t.clone() for t in ops.knn_topk(ctx, q, db, 10)]
    plain2 = [t.clone() for t in ops.knn_topk(ctx, q, db2, 10)]
    l0 = ctx.launch_count()
    ops.knn_topk(ctx, q, db, 10)
    unfitted = ctx.launch_count() - l0
    ops.knn_fit(ctx, db)
    ops.pdist(ctx, db[:300].contiguous())                       # other distance calls must not disturb the fitted planes
    l0 = ctx.launch_count()
    fitted = ops.knn_topk(ctx, q, db, 10)
    assert ctx.launch_count() - l0 == unfitted - 2              # no max-abs / split pass over the database
    assert all(torch.equal(a, b) for a, b in zip(plain, fitted))
    other = ops.knn_topk(ctx, q, db2, 10)                       # a different tensor: prepared per call as before
    assert all(torch.equal(a, b) for a, b in zip(plain2, other))
    from facerec_b200 import weights as W
    assert ops.workspace_bytes(ctx, W.MODEL_IDS["irv1"], 8) in (-1,) or ops.workspace_bytes(ctx, W.MODEL_IDS["irv1"], 8) > 0


def test_knn_vote_matches_sklearn_predict_proba(ctx):
    from sklearn.neighbors import KNeighborsClassifier
    from facerec_b200 import ops
    rng = np.random.RandomState(2)
    n_cls, per = 30, 40
    cent = rng.standard_normal((n_cls, 512)); cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    X = np.concatenate([cent[c] + 0.08 * rng.standard_normal((per, 512)) for c in range(n_cls)]).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    y = np.repeat(np.arange(100, 100 + n_cls), per).astype(np.int32)
    V = (cent[rng.randint(0, n_cls, 200)] + 0.3 * rng.standard_normal((200, 512))).astype(np.float32)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    knn = KNeighborsClassifier(n_neighbors=10, weights="uniform").fit(X, y)
    ref = knn.predict_proba(V)
    idx, _ = ops.knn_topk(ctx, torch.from_numpy(V).cuda(), torch.from_numpy(X).cuda(), 10)
    lab = torch.from_numpy(np.searchsorted(knn.classes_, y).astype(np.int32)).cuda()
    proba = ops.knn_vote(ctx, idx, lab, n_cls).cpu().numpy()
    assert np.abs(proba - ref).max() < 1e-6


def test_segment_mean_matches_numpy(ctx):
    from facerec_b200 import ops
    rng = np.random.RandomState(9)
    lens = rng.randint(1, 40, 57)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = rng.standard_normal((off[-1], 512)).astype(np.float32)
    got = ops.segment_mean(ctx, torch.from_numpy(x).cuda(), torch.from_numpy(off).cuda()).cpu().numpy()
    ref = np.stack([x[off[i]:off[i + 1]].mean(axis=0) for i in range(len(lens))])
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n,d,dup", [(2, 8, False), (37, 16, False), (700, 512, False), (1500, 512, False), (400, 4, True),
                                     (4000, 512, False), (3000, 3, True)])
def test_gpu_linkage_equals_scipy_on_same_distances(ctx, n, d, dup):
    """frb_linkage_complete + host relabelling == scipy linkage(method='complete') fed the SAME fp32 distances
    (bit-identical Z, ties included: `dup` plants duplicate points and a coarse grid so many distances tie)."""
    from scipy.cluster.hierarchy import linkage, fcluster
    from facerec_b200 import ops, _lib, distance
    rng = np.random.RandomState(n)
    if dup:
        x = rng.randint(0, 4, (n, d)).astype(np.float32)
    else:
        x = _mixture(rng, n, d, centres=max(2, n // 18), unit=False)
    Z = distance.gpu_linkage_complete(x, ctx)
    cond = ops.pdist(ctx, torch.from_numpy(x).cuda(), layout=_lib.FRB_CONDENSED_F32).cpu().numpy().astype(np.float64)
    ref = linkage(cond, method="complete")
    assert np.array_equal(Z, ref)
    if n >= 36:
        k = n // 18
        assert np.array_equal(fcluster(Z, t=k, criterion="maxclust"), fcluster(ref, t=k, criterion="maxclust"))


def test_label_merges_matches_scipy_convention():   # (host-only logic, runs with the gpu-marked module)
    from facerec_b200.distance import _label_merges
    Z = _label_merges(np.array([[0, 3], [1, 2], [2, 3]], np.int32), np.array([0.5, 0.2, 0.9]), 4)
    assert Z.tolist() == [[1.0, 2.0, 0.2, 2.0], [0.0, 3.0, 0.5, 2.0], [4.0, 5.0, 0.9, 4.0]]
