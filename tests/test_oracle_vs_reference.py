"""Oracle vs the LIVE reference on fresh seeds (build container only; skipped
where /root/reference does not exist, e.g. on the GPU box)."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import polara_oracle as po
from oracle.ref_shim import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference checkout absent")


@pytest.mark.parametrize("seed", [1, 2])
def test_downvote_topk_rescale_live(seed):
    import_reference()
    from polara.recommender.models import RecommenderModel
    from polara.preprocessing.matrices import rescale_matrix
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((20, 50))
    rows = np.repeat(np.arange(20), 4)
    cols = np.concatenate([rng.choice(50, 4, replace=False) for _ in range(20)])
    ref = s.copy()
    RecommenderModel.downvote_seen_items(ref, (rows, cols))
    mine = po.downvote_seen_items(s.copy(), rows, cols)
    np.testing.assert_array_equal(mine, ref)
    for row in range(20):
        np.testing.assert_array_equal(po.topsort(ref[row], 6), RecommenderModel.topsort(ref[row], 6))
    a = sps.random(40, 30, density=0.2, random_state=seed, format="csr")
    for scaling, axis in ((0.4, 0), (0.8, 1), (1, 0)):
        np.testing.assert_allclose(po.rescale_matrix(a, scaling, axis).toarray(),
                                   rescale_matrix(a, scaling, axis).toarray(), rtol=1e-14)


def test_hooi_live():
    import_reference()
    from polara.lib.tensor import hooi
    rng = np.random.default_rng(3)
    shp = (40, 30, 5)
    nnz = 900
    idx = np.unique(np.stack([rng.integers(0, s, nnz) for s in shp], axis=1), axis=0).astype(np.intp)
    val = np.ones(len(idx))
    ref = hooi(idx, val, shp, (4, 3, 2), num_iters=6, growth_tol=1e-4, seed=5)
    mine = po.hooi(idx, val, shp, (4, 3, 2), num_iters=6, growth_tol=1e-4, seed=5)
    for a, b in zip(mine[:3], ref[:3]):
        sv = np.linalg.svd(a.T @ b, compute_uv=False)
        assert sv.min() > 1 - 1e-9
    np.testing.assert_allclose(np.linalg.norm(mine[3]), np.linalg.norm(ref[3]), rtol=1e-10)
