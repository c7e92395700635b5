import sys, numpy as np
sys.path.insert(0, '/root/repo')
from polara_b200.engine import get_engine
from tests.helpers import random_seen_csr
eng = get_engine(0)
for (m, n, r, k) in ((64, 300, 7, 25), (50, 40, 5, 10), (1, 513, 16, 3)):
    rng = np.random.default_rng(5)
    e = (rng.standard_normal((m, r)) * (0.9 ** np.arange(r))).astype(np.float32)
    v = rng.standard_normal((n, r)).astype(np.float32)
    per_row = rng.integers(0, min(n, 40), size=m)
    rows, cols, indptr = random_seen_csr(rng, m, n, per_row)
    e_dev, v_dev = eng.upload(e), eng.upload(v)
    seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32)))
    for kern in ("simt", "tcgen05"):
        eng.set_score_kernel(kern)
        try:
            ids, sc = eng.score_topk(e_dev, v_dev, r, k, seen=seen, want_scores=True)
            eng.sync()
            print((m, n, r, k), kern, "ok", ids[0, :5].tolist())
        except Exception as ex:
            print((m, n, r, k), kern, "FAILED:", ex)
            raise
