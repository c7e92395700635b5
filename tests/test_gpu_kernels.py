"""Per-kernel parity of the CUDA path (through the C-ABI) against the oracle / numpy.
Runs on the B200 box only."""
import numpy as np
import pytest
import scipy.sparse as sps
import torch

from oracle import polara_oracle as po
from tests.helpers import check_topk_against_scores, random_seen_csr, subspace_gap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from polara_b200.engine import get_engine
    return get_engine(0)


def _rand_csr(rng, m, n, density, heavy_col=False):
    a = sps.random(m, n, density=density, random_state=np.random.RandomState(rng.integers(1 << 30)), format="csr",
                   dtype=np.float32)
    a.data = np.rint(1 + 4 * a.data).astype(np.float32)
    if heavy_col:   # one very popular column and one very long row: exercises the block-cooperative path
        col = np.zeros((m, 1), dtype=np.float32); col[rng.random(m) < 0.9] = 2.0
        a = sps.hstack([a[:, :-1], sps.csr_matrix(col)]).tocsr()
        row = np.zeros((1, n), dtype=np.float32); row[0, rng.random(n) < 0.8] = 3.0
        a = sps.vstack([a[:-1], sps.csr_matrix(row)]).tocsr()
    a.sort_indices()
    return a


@pytest.mark.parametrize("m,n,density,ell,heavy", [(1000, 700, 0.02, 32, False), (6000, 9000, 0.004, 64, True),
                                                   (300, 50, 0.3, 96, False), (5000, 6000, 0.01, 160, True),
                                                   (17, 5, 0.5, 32, False)])
def test_spmm_matches_scipy(eng, m, n, density, ell, heavy):
    rng = np.random.default_rng(0)
    a = _rand_csr(rng, m, n, density, heavy)
    x = rng.standard_normal((n, ell)).astype(np.float32)
    a_dev = eng.upload_csr(a.indptr, a.indices, a.data, a.shape)
    y = eng.spmm(a_dev, eng.upload(x)).cpu().numpy()
    ref = a.astype(np.float64) @ x.astype(np.float64)
    scale = np.abs(a).astype(np.float64) @ np.abs(x).astype(np.float64) + 1e-6
    assert np.max(np.abs(y - ref) / scale) < 5e-6
    # determinism: bit-identical on a second run
    y2 = eng.spmm(a_dev, eng.upload(x)).cpu().numpy()
    assert np.array_equal(y, y2)


@pytest.mark.parametrize("ell,ldx", [(50, 64), (1, 32), (33, 40), (70, 96), (130, 160)])
def test_spmm_live_columns_only(eng, ell, ldx):
    """ell need not be a multiple of 32: X is read up to column ell (the rest may hold anything), Y comes back in whole
    groups of 32 columns with zeros beyond ell, and the live columns are bit-identical to the padded call."""
    rng = np.random.default_rng(5)
    a = _rand_csr(rng, 700, 900, 0.02, True)
    x = rng.standard_normal((900, ldx)).astype(np.float32)
    x_poison = x.copy(); x_poison[:, ell:] = np.nan
    a_dev = eng.upload_csr(a.indptr, a.indices, a.data, a.shape)
    y = eng.spmm(a_dev, eng.upload(x_poison), ell=ell).cpu().numpy()
    assert y.shape == (700, (ell + 31) // 32 * 32)
    assert not y[:, ell:].any() and np.isfinite(y).all()
    x_zero = x.copy(); x_zero[:, ell:] = 0
    full = eng.spmm(a_dev, eng.upload(np.ascontiguousarray(np.pad(x_zero, ((0, 0), (0, y.shape[1] + 32 - ldx))))), ell=y.shape[1]).cpu().numpy()
    assert np.array_equal(y[:, :ell], full[:, :ell])


def test_spmm_empty_rows_and_empty_matrix(eng):
    a = sps.csr_matrix((np.array([1.0, 2.0], dtype=np.float32), (np.array([3, 3]), np.array([0, 4]))), shape=(9, 5))
    x = np.arange(5 * 32, dtype=np.float32).reshape(5, 32)
    a_dev = eng.upload_csr(a.indptr, a.indices, a.data, a.shape)
    y = eng.spmm(a_dev, eng.upload(x)).cpu().numpy()
    np.testing.assert_array_equal(y, (a @ x))
    z = sps.csr_matrix((4, 5), dtype=np.float32)
    z_dev = eng.upload_csr(z.indptr, np.zeros(0, np.int32), np.zeros(0, np.float32), z.shape)
    y = eng.spmm(z_dev, eng.upload(x)).cpu().numpy()
    assert y.shape == (4, 32) and not y.any()


@pytest.mark.parametrize("m,n,density,heavy", [(2000, 1500, 0.01, True), (64, 9000, 0.02, False), (5, 3, 0.6, False)])
def test_transpose_matches_scipy(eng, m, n, density, heavy):
    rng = np.random.default_rng(1)
    a = _rand_csr(rng, m, n, density, heavy)
    a_dev = eng.upload_csr(a.indptr, a.indices, a.data, a.shape)
    t = eng.transpose(a_dev)
    ref = a.T.tocsr(); ref.sort_indices()
    np.testing.assert_array_equal(t.indptr.cpu().numpy(), ref.indptr)
    np.testing.assert_array_equal(t.indices.cpu().numpy(), ref.indices)
    np.testing.assert_array_equal(t.values.cpu().numpy(), ref.data)


def test_rescale_matches_oracle(eng, golden):
    g = golden("kernels_small")
    a = sps.csr_matrix((g["a_data"], g["a_indices"], g["a_indptr"]), shape=tuple(g["a_shape"]))
    for rs, cs in ((0.7, 1.0), (1.0, 0.4), (0.8, 0.4)):
        a_dev = eng.upload_csr(a.indptr, a.indices, a.data, a.shape)
        eng.rescale(a_dev, rs, cs)
        got = sps.csr_matrix((a_dev.values.cpu().numpy(), a.indices, a.indptr), shape=a.shape).toarray()
        ref = po.scaled_training_matrix(a, rs, cs).toarray()
        np.testing.assert_allclose(got, ref, rtol=2e-6)
    # and against the recorded reference outputs
    a_dev = eng.upload_csr(a.indptr, a.indices, a.data, a.shape)
    eng.rescale(a_dev, 0.7, 1.0)
    got = sps.csr_matrix((a_dev.values.cpu().numpy(), a.indices, a.indptr), shape=a.shape).toarray()
    np.testing.assert_allclose(got, g["sc_rows"], rtol=2e-6)


@pytest.mark.parametrize("n,c,rank", [(5000, 64, 10), (777, 130, 40), (40, 5, 5), (3600, 5, 3)])
def test_tall_svd_matches_numpy(eng, n, c, rank):
    rng = np.random.default_rng(2)
    base = rng.standard_normal((n, c)) * (0.8 ** np.arange(c))
    m = base.astype(np.float32)
    u, s, vt = eng.tall_svd(eng.upload(m), rank, want_vt=True)
    uu, ss, vvt = np.linalg.svd(m.astype(np.float64), full_matrices=False)
    np.testing.assert_allclose(s.cpu().numpy(), ss[:rank], rtol=2e-5)
    u = u[:, :rank].cpu().numpy(); vt = vt.cpu().numpy()
    # compare the rank-`rank` reconstruction (sign/rotation free)
    rec = (u * s.cpu().numpy()) @ vt
    ref = (uu[:, :rank] * ss[:rank]) @ vvt[:rank]
    assert np.abs(rec - ref).max() < 2e-4 * ss[0]
    np.testing.assert_allclose(u.T @ u, np.eye(rank), atol=5e-5)


def _planted(n_users, n_items, per_user, rank, seed):
    from polara_b200.synth import planted_ratings
    u, i, r = planted_ratings(n_users, n_items, per_user, rank=rank, decay=0.75, seed=seed)
    return sps.csr_matrix((r, (u, i)), shape=(n_users, n_items)), (u, i, r)


@pytest.mark.parametrize("rank,ell", [(10, 32), (16, 64)])
def test_rsvd_matches_arpack(eng, rank, ell):
    a, _ = _planted(3000, 1200, 50, 24, seed=3)
    a_dev = eng.upload_csr(a.indptr, a.indices, a.data, a.shape)
    at_dev = eng.transpose(a_dev)
    v, sigma, u, iters = eng.rsvd(a_dev, at_dev, rank, ell, max_iters=16, tol=1e-8, seed=1, want_u=True)
    v_ref, s_ref, u_ref = po.svd_build(a, rank, return_u=True)
    np.testing.assert_allclose(sigma.cpu().numpy(), s_ref, rtol=1e-4)
    assert subspace_gap(v[:, :rank].cpu().numpy(), v_ref) < 1e-2
    assert subspace_gap(u[:, :rank].cpu().numpy(), u_ref) < 1e-2
    vv = v[:, :rank].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(vv.T @ vv, np.eye(rank), atol=1e-4)
    # padding columns of the device buffer stay zero (they are read by the SpMM)
    assert not v[:, rank:].any()


@pytest.mark.parametrize("kernel", ["simt", "tcgen05"])
@pytest.mark.parametrize("m,n,r,k,filt", [(200, 1000, 10, 10, True), (333, 4097, 50, 10, True), (64, 300, 7, 25, True),
                                          (130, 2500, 128, 10, False), (50, 40, 5, 10, True), (1, 513, 16, 3, True)])
def test_score_topk_matches_oracle(eng, kernel, m, n, r, k, filt):
    rng = np.random.default_rng(5)
    eng.set_score_kernel(kernel)
    e = (rng.standard_normal((m, r)) * (0.9 ** np.arange(r))).astype(np.float32)
    v = rng.standard_normal((n, r)).astype(np.float32)
    per_row = rng.integers(0, min(n, 40), size=m)
    if n <= 64:
        per_row[: m // 2] = n - 4           # fewer unseen items than k: seen items must re-enter in score order
    rows, cols, indptr = random_seen_csr(rng, m, n, per_row)
    e_dev, v_dev = eng.upload(e), eng.upload(v)
    seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32))) if filt else None
    ids, sc = eng.score_topk(e_dev, v_dev, r, k, seen=seen, want_scores=True)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    s64 = e.astype(np.float64) @ v.astype(np.float64).T
    tol = 4e-6 * np.abs(e).astype(np.float64).sum(1).max() * np.abs(v).max()
    frac = check_topk_against_scores(ids, s64, rows if filt else [], cols if filt else [], k, tol)
    assert frac > 0.99
    # reported scores are the canonical fp32 scores of the reported items
    got = np.take_along_axis(s64, ids, axis=1)
    np.testing.assert_allclose(sc, got, atol=tol)


def test_score_kernels_agree_bitwise(eng):
    """The tcgen05 kernel (bf16 filter + exact rescoring) must return exactly what the
    exact fp32 SIMT kernel returns: same ids, same scores."""
    rng = np.random.default_rng(6)
    m, n, r, k = 700, 20000, 50, 10
    e = rng.standard_normal((m, r)).astype(np.float32)
    v = rng.standard_normal((n, r)).astype(np.float32)
    rows, cols, indptr = random_seen_csr(rng, m, n, rng.integers(0, 200, size=m))
    e_dev, v_dev = eng.upload(e), eng.upload(v)
    seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32)))
    out = {}
    for kernel in ("simt", "tcgen05"):
        eng.set_score_kernel(kernel)
        ids, sc = eng.score_topk(e_dev, v_dev, r, k, seen=seen, want_scores=True)
        out[kernel] = (ids.cpu().numpy(), sc.cpu().numpy())
    np.testing.assert_array_equal(out["simt"][0], out["tcgen05"][0])
    np.testing.assert_array_equal(out["simt"][1], out["tcgen05"][1])


def _score_case(eng, rng, m, n, r, k, scale=1.0, kernel="tcgen05"):
    e = (rng.standard_normal((m, r)) * (0.9 ** np.arange(r)) * scale).astype(np.float32)
    v = rng.standard_normal((n, r)).astype(np.float32)
    rows, cols, indptr = random_seen_csr(rng, m, n, rng.integers(0, min(n, 40), size=m))
    e_dev, v_dev = eng.upload(e), eng.upload(v)
    seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32)))
    out = {}
    for kern in ("simt", kernel):
        eng.set_score_kernel(kern)
        ids, sc = eng.score_topk(e_dev, v_dev, r, k, seen=seen, want_scores=True)
        out[kern] = (ids.cpu().numpy(), sc.cpu().numpy())
    return out


def test_score_tc_ignores_stale_shared_memory(eng):
    """A launch must not pick up the k-th-score exchange entries a previous launch left in shared memory: run a problem
    with 100x larger scores first (its thresholds would wipe out every candidate of the second one), then a different
    problem with single-tile work items, and compare with the exact kernel."""
    rng = np.random.default_rng(21)
    _score_case(eng, rng, 333, 4097, 50, 10, scale=100.0)
    out = _score_case(eng, rng, 333, 4097, 50, 10)
    np.testing.assert_array_equal(out["simt"][0], out["tcgen05"][0])
    np.testing.assert_array_equal(out["simt"][1], out["tcgen05"][1])


@pytest.mark.parametrize("m,n,r,k", [(333, 4097, 50, 10), (200, 1000, 10, 10), (1000, 3000, 16, 25), (2000, 20000, 50, 10),
                                     (129, 700, 33, 5)])
def test_score_pair_mode_agrees_bitwise(eng, m, n, r, k):
    """PB200_TC_PAIR=1: CTA pairs issue tcgen05.mma.cta_group::2 (each CTA stages half of every item tile); results
    must stay bit-identical to the exact kernel."""
    import os
    rng = np.random.default_rng(22)
    os.environ["PB200_TC_PAIR"] = "1"
    try:
        _score_case(eng, rng, m, n, r, k, scale=50.0)          # leaves different thresholds behind in shared memory
        out = _score_case(eng, rng, m, n, r, k)
    finally:
        os.environ.pop("PB200_TC_PAIR", None)
    np.testing.assert_array_equal(out["simt"][0], out["tcgen05"][0])
    np.testing.assert_array_equal(out["simt"][1], out["tcgen05"][1])


def test_score_topk_sharded_merge_equals_unsharded(eng):
    rng = np.random.default_rng(7)
    m, n, r, k = 300, 6000, 32, 10
    e = rng.standard_normal((m, r)).astype(np.float32)
    v = rng.standard_normal((n, r)).astype(np.float32)
    rows, cols, indptr = random_seen_csr(rng, m, n, rng.integers(0, 60, size=m))
    e_dev, v_dev = eng.upload(e), eng.upload(v)
    seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32)))
    full = eng.score_topk(e_dev, v_dev, r, k, seen=seen).cpu().numpy()
    parts = []
    bounds = [0, 1500, 3100, 6000]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        # seen ids stay global: the kernel compares local id + item_offset
        parts.append(eng.score_topk_cands(e_dev, eng.upload(v[lo:hi]), r, k, seen=seen, item_offset=lo))
    stacked = torch.stack(parts).contiguous()
    merged = eng.merge_cands(stacked, len(parts), m, k).cpu().numpy()
    np.testing.assert_array_equal(merged, full)


def test_score_dense_matches_numpy(eng):
    rng = np.random.default_rng(8)
    e = rng.standard_normal((3, 20)).astype(np.float32)
    v = rng.standard_normal((999, 20)).astype(np.float32)
    s = eng.score_dense(eng.upload(e), eng.upload(v), 20).cpu().numpy()
    np.testing.assert_allclose(s, e.astype(np.float64) @ v.astype(np.float64).T, atol=1e-4)


def test_ttm_matches_reference_fixture(eng, golden):
    g = golden("kernels_small")
    idx, val, shp = g["ttm_idx"], g["ttm_val"], tuple(int(s) for s in g["ttm_shape"])
    u, v = g["ttm_u"], g["ttm_v"]      # u: [n1 x 3], v: [n2 x 2] ; fixture = ttm3d_seq(idx,val,shp, v, u, ((2,0),(1,0)))
    i0, i1, i2 = (eng.upload(idx[:, c].astype(np.int32)) for c in range(3))
    vals = eng.upload(val.astype(np.float32))
    seg, a1, a2, vv = eng.coo_group(i0, shp[0], i1, i2, vals)
    out = eng.ttm(shp[0], seg, a2, a1, vv, eng.upload(v.astype(np.float32)), v.shape[1],
                  eng.upload(u.astype(np.float32)), u.shape[1])
    got = out[:, : v.shape[1] * u.shape[1]].cpu().numpy().reshape(shp[0], v.shape[1], u.shape[1])
    np.testing.assert_allclose(got, g["ttm0"], rtol=2e-5, atol=2e-5)
    # few-segment variant: group by mode 2, contract modes 1 and 0
    seg2, b0, b1, vv2 = eng.coo_group(i2, shp[2], i0, i1, vals)
    w0 = np.random.default_rng(0).standard_normal((shp[0], 4)).astype(np.float32)
    red = eng.ttm_reduce(shp[2], seg2, b1, b0, vv2, eng.upload(u.astype(np.float32)), u.shape[1], eng.upload(w0), 4)
    ref = po.ttm3d(idx, val, shp, u, w0.astype(np.float64), 2, 1, 0).reshape(shp[2], -1)
    np.testing.assert_allclose(red.cpu().numpy(), ref, rtol=2e-5, atol=2e-5)
