"""Shared test helpers (CPU side)."""
import numpy as np
import scipy.sparse as sps

from oracle import polara_oracle as po


def subspace_gap(v_a, v_b):
    s = np.linalg.svd(np.asarray(v_a, dtype=np.float64).T @ np.asarray(v_b, dtype=np.float64), compute_uv=False)
    return float(np.sqrt(max(0.0, 1.0 - s.min() ** 2)))


def check_topk_against_scores(ids, scores64, seen_rows, seen_cols, k, tol):
    """``ids`` [m x k] must be, for every row, a valid top-k of the f64 oracle scores under
    the reference's order (unseen by score desc, then seen by score desc), allowing swaps
    only between items whose oracle scores differ by less than ``tol`` (fp32 near-ties).
    Returns the fraction of entries that agree exactly with the oracle order."""
    m, n = scores64.shape
    seen = sps.csr_matrix((np.ones(len(seen_rows), dtype=bool), (seen_rows, seen_cols)), shape=(m, n)).toarray() \
        if len(seen_rows) else np.zeros((m, n), dtype=bool)
    exact = 0
    for u in range(m):
        ref = po.rank_key_order(scores64[u], np.flatnonzero(seen[u]), k)
        mine = ids[u]
        assert len(set(mine.tolist())) == k, "duplicate items in row %d" % u
        n_unseen = n - seen[u].sum()
        # seen items may only appear after all unseen ones are exhausted
        assert not seen[u][mine[:min(k, n_unseen)]].any(), "seen item recommended in row %d" % u
        s_ref = scores64[u][ref]
        s_mine = scores64[u][mine]
        key_ref = np.where(seen[u][ref], -1e30, 0) + s_ref
        key_mine = np.where(seen[u][mine], -1e30, 0) + s_mine
        np.testing.assert_allclose(key_mine, key_ref, rtol=0, atol=tol,
                                   err_msg="row %d is not a top-%d within tolerance" % (u, k))
        exact += (mine == ref).sum()
    return exact / (m * k)


def random_seen_csr(rng, m, n, per_row):
    rows, cols = [], []
    for u in range(m):
        c = np.sort(rng.choice(n, size=min(n, per_row[u]), replace=False))
        rows.append(np.full(len(c), u))
        cols.append(c)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=m), out=indptr[1:])
    return rows, cols, indptr
