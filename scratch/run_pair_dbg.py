"""Diagnose the CTA-pair mode: which (row, item) pairs does it miss against the exact SIMT kernel?"""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from polara_b200.engine import get_engine
from tests.helpers import random_seen_csr
eng = get_engine(0)
for (m, n, r, k, seed) in ((333, 4097, 50, 10, 5), (2000, 20000, 50, 10, 6), (1000, 3000, 16, 10, 7)):
    rng = np.random.default_rng(seed)
    e = (rng.standard_normal((m, r)) * (0.9 ** np.arange(r))).astype(np.float32)
    v = rng.standard_normal((n, r)).astype(np.float32)
    rows, cols, indptr = random_seen_csr(rng, m, n, rng.integers(0, min(n, 40), size=m))
    e_dev, v_dev = eng.upload(e), eng.upload(v)
    seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32)))
    out = {}
    for name, kernel, pair in (("simt", "simt", "0"), ("tc", "tcgen05", "0"), ("pair", "tcgen05", "1")):
        os.environ["PB200_TC_PAIR"] = pair
        eng.set_score_kernel(kernel)
        ids, sc = eng.score_topk(e_dev, v_dev, r, k, seen=seen, want_scores=True)
        out[name] = (ids.cpu().numpy(), sc.cpu().numpy())
    norms = np.sqrt((v.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    order = np.argsort(-norms, kind="stable")
    pos_of = np.empty(n, dtype=np.int64); pos_of[order] = np.arange(n)
    print("case", (m, n, r), "tc==simt", np.array_equal(out["tc"][0], out["simt"][0]), "pair==simt", np.array_equal(out["pair"][0], out["simt"][0]))
    bad = np.flatnonzero((out["pair"][0] != out["simt"][0]).any(1))
    print("  rows differing:", len(bad), "of", m, "first:", bad[:20].tolist())
    miss = []
    for u in bad:
        for it in set(out["simt"][0][u].tolist()) - set(out["pair"][0][u].tolist()):
            p = pos_of[it]; miss.append((int(u), int(u) // 128, int(it), int(p), int(p) // 128, int(p) % 128))
    print("  missed (row, user_tile, item, pos, item_tile, col):")
    for t in miss[:40]: print("   ", t)
    if miss:
        cols_ = np.array([t[5] for t in miss]); ut = np.array([t[1] for t in miss]); it_ = np.array([t[4] for t in miss])
        print("  col<64:", int((cols_ < 64).sum()), "col>=64:", int((cols_ >= 64).sum()), " user tiles odd:", int((ut % 2).sum()), "even:", int((ut % 2 == 0).sum()))
        print("  item tiles hist:", np.bincount(it_).tolist())
