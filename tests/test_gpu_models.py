"""End-to-end parity of the device models against fixtures recorded from the REAL
reference (tests/golden, made by oracle/make_golden.py).  B200 box only."""
import numpy as np
import pytest

from oracle import polara_oracle as po
from tests.helpers import subspace_gap

pytestmark = pytest.mark.gpu


def _svd_model(g, scaled=False):
    from polara_b200.host import ArrayData
    from polara_b200.models import B200ScaledSVD, B200SVDModel
    data = ArrayData.from_golden(g)
    model = (B200ScaledSVD if scaled else B200SVDModel)(data)
    model.verbose = False
    model.rank = int(g["rank"])
    if scaled:
        model.col_scaling = float(g["col_scaling"])
        model.row_scaling = float(g["row_scaling"])
    sp = float(g["switch_positive"])
    model.switch_positive = None if np.isnan(sp) else sp
    return model


@pytest.mark.parametrize("name", ["svd_warm_r10", "svd_known_r8", "svd_scaled_r10"])
def test_svd_model_reproduces_reference(golden, name):
    g = golden(name)
    model = _svd_model(g, scaled=bool(g["scaled"]))
    model.build()
    np.testing.assert_allclose(model.factors["singular_values"], g["singular_values"], rtol=2e-4)
    assert subspace_gap(model.factors["itemid"], g["item_factors"]) < 2e-2
    recs = model.get_recommendations()
    assert recs.dtype == np.int64 and recs.shape == g["recs"].shape
    assert (recs == g["recs"]).mean() > 0.97          # fp32 + subspace tolerance: a few near-tie swaps
    # evaluate(): hit counts within a couple of hits of the recorded reference numbers
    hits = model.evaluate("hits")
    ref = g["hits"]
    assert abs(hits.true_positive - ref[0]) <= 3
    assert abs(hits.false_negative - ref[3]) <= 3


@pytest.mark.parametrize("name", ["svd_warm_r10", "svd_known_r8"])
def test_scoring_with_reference_factors_is_exact(golden, name):
    """Feeding the reference's own factors isolates the scoring path: lists must match the
    reference's lists except where its f64 scores are tied to within fp32 resolution."""
    g = golden(name)
    model = _svd_model(g)
    f = model.data.fields
    model.factors = {f.userid: None, f.itemid: g["item_factors"].copy(), "singular_values": g["singular_values"]}
    model._is_ready = True
    for kernel in ("simt", "tcgen05"):
        model.score_kernel = kernel
        model._recommendations = None
        recs = model.get_recommendations()
        assert (recs == g["recs"]).mean() > 0.995, kernel
        model.topk = 25
        model.rank = int(g["rank_reduced"])          # rank truncation without rebuild (models.py:819-832)
        recs25 = model.get_recommendations()
        assert (recs25 == g["recs_top25"]).mean() > 0.995, kernel
        model.topk = 10
        assert (model.get_recommendations() == g["recs_reduced"]).mean() > 0.995
        model.filter_seen = False
        assert (model.get_recommendations() == g["recs_unfiltered"]).mean() > 0.995
        model.filter_seen = True
        model.factors = {f.userid: None, f.itemid: g["item_factors"].copy(), "singular_values": g["singular_values"]}
        model._rank = int(g["rank"])


@pytest.mark.parametrize("name,flat", [("coffee_small", None), ("coffee_flat34", [2, 3])])
def test_coffee_model_reproduces_reference(golden, name, flat):
    from polara_b200.host import ArrayData
    from polara_b200.models import B200CoffeeModel
    g = golden(name)
    model = B200CoffeeModel(ArrayData.from_golden(g))
    model.verbose = False
    model.mlrank = tuple(int(x) for x in g["mlrank"])
    model.seed = int(g["seed"])
    model.num_iters = int(g["num_iters"])
    model.growth_tol = float(g["growth_tol"])
    if flat is not None:
        model.flattener = flat
    model.build()
    for key, ref in (("userid", "u0"), ("itemid", "u1"), ("rating", "u2")):
        assert subspace_gap(model.factors[key], g[ref]) < 2e-2, key
    np.testing.assert_allclose(np.linalg.norm(model.factors["core"]), np.linalg.norm(g["core"]), rtol=1e-3)
    assert model.factors["core"].shape == g["core"].shape
    recs = model.get_recommendations()
    assert (recs == g["recs"]).mean() > 0.95
    # scoring alone, from the reference's factors: near-exact
    f = model.data.fields
    model.factors = {f.userid: g["u0"], f.itemid: g["u1"], f.feedback: g["u2"], "core": g["core"]}
    recs = model.get_recommendations()
    assert (recs == g["recs"]).mean() > 0.995


def test_threshold_zero_feedback_stays_seen():
    """models.py:191-211: zeroed (sub-threshold) feedback is dropped from P but still masked."""
    from polara_b200.host import ArrayData
    from polara_b200.models import B200SVDModel
    rng = np.random.default_rng(0)
    v = np.linalg.qr(rng.standard_normal((30, 4)))[0]
    user = np.array([0, 0, 0, 1, 1]); item = np.array([3, 7, 9, 1, 2])
    fdbk = np.array([5.0, 0.0, 4.0, 0.0, 3.0])
    data = ArrayData(np.zeros((1, 2), dtype=np.int64), np.ones(1), (2, 30), user, item, fdbk, (2, 30), warm_start=True)
    model = B200SVDModel(data)
    model.verbose = False
    model.rank = 4
    model.topk = 5
    model.factors = {"userid": None, "itemid": v, "singular_values": np.ones(4)}
    model._is_ready = True
    recs = model.get_recommendations()
    ref = po.recommend_svd(user, item, fdbk, (2, 30), v, topk=5)
    np.testing.assert_array_equal(recs, ref)
    assert 7 not in recs[0] and 1 not in recs[1]


def test_missing_inputs_raise():
    from polara_b200.host import ArrayData
    from polara_b200.models import B200SVDModel
    data = ArrayData(np.zeros((1, 2), dtype=np.int64), np.ones(1), (2, 30))
    model = B200SVDModel(data)
    with pytest.raises(NotImplementedError):
        model.build(operator=object())
    model.factors = {"userid": None, "itemid": np.zeros((30, 4)), "singular_values": np.ones(4)}
    with pytest.raises(ValueError):
        model.get_recommendations()        # no test data (data.py:840-841)


def test_streamed_fast_path_matches_plain():
    """The pinned-CSR fast path (user chunks, H2D overlapped with scoring) returns exactly what the plain path does."""
    import scipy.sparse as sps
    import torch
    from polara_b200.host import ArrayData
    from polara_b200.models import B200SVDModel
    from polara_b200.synth import popularity_csr
    m, n = 4 * 65536 + 777, 3000
    indptr, indices, values = popularity_csr(m, n, 12 * m, seed=9)
    v = np.linalg.qr(np.random.default_rng(1).standard_normal((n, 12)))[0] * (0.9 ** np.arange(12))
    data = ArrayData(np.zeros((1, 2), dtype=np.int64), np.ones(1), (m, n))
    data.test_csr = ((torch.from_numpy(indptr).pin_memory(), torch.from_numpy(indices).pin_memory(),
                      torch.from_numpy(values).pin_memory()), (m, n))
    model = B200SVDModel(data)
    model.verbose = False
    model.rank = 12
    model.factors = {"userid": None, "itemid": v, "singular_values": np.ones(12)}
    model._is_ready = True
    streamed = model.get_recommendations()
    model.stream_chunks = 1
    single = model.get_recommendations()
    np.testing.assert_array_equal(streamed, single)
    # and the COO route (what a polara data model feeds) gives the same lists
    rows = np.repeat(np.arange(m), np.diff(indptr))
    data2 = ArrayData(np.zeros((1, 2), dtype=np.int64), np.ones(1), (m, n), rows, indices.astype(np.int64),
                      values.astype(np.float64), (m, n), warm_start=True)
    model2 = B200SVDModel(data2)
    model2.verbose = False
    model2.rank = 12
    model2.factors = {"userid": None, "itemid": v, "singular_values": np.ones(12)}
    model2._is_ready = True
    np.testing.assert_array_equal(model2.get_recommendations(), single)
