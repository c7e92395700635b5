"""How much of the norm-ordered sweep could be cut by the Cauchy-Schwarz bound ||e_u|| * ||v_pos|| < t_u ?
(CPU experiment on a reduced C2-like matrix; development aid.)"""
import sys, time
import numpy as np, scipy.sparse as sps
from scipy.sparse.linalg import svds
sys.path.insert(0, '.')
from polara_b200.synth import popularity_csr
n_users, n_items, nnz, r, k = 100_000, 100_000, 10_000_000, 50, 10
ip, ix, vl = popularity_csr(n_users, n_items, nnz, seed=20260924)
a = sps.csr_matrix((vl.astype(np.float64), ix, ip), shape=(n_users, n_items))
t0 = time.time()
_, s, vt = svds(a, k=r)
print('svds', time.time() - t0, 's')
v = np.ascontiguousarray(vt[::-1].T)
vn = np.linalg.norm(v, axis=1)
order = np.argsort(-vn)
vs = vn[order]
print('item norm quantiles (sorted desc): ', [float('%.4g' % vs[int(q * (n_items - 1))]) for q in (0, .001, .01, .05, .1, .25, .5, .9)])
users = np.random.default_rng(0).choice(n_users, 128 * 64, replace=False)
users.sort()
e = a[users] @ v
en = np.linalg.norm(e, axis=1)
sc = e @ v.T
sc[a[users].nonzero()] = -np.inf
kth = -np.partition(-sc, k - 1, axis=1)[:, k - 1]
ratio = kth / en                      # user u needs items with norm >= ratio_u only
pos = np.searchsorted(-vs, -ratio)    # first sweep position whose norm is below the user's cut
print('per-user cut position quantiles:', [int(np.quantile(pos, q)) for q in (.1, .5, .9, .99, 1.0)])
tile_cut = pos.reshape(-1, 128).max(axis=1)
print('per-128-user-tile cut (max over users): mean %.0f  median %.0f  max %d of %d items' % (tile_cut.mean(), np.median(tile_cut), tile_cut.max(), n_items))
print('fraction of tiles left to sweep: %.3f' % (np.ceil(tile_cut / 128).sum() / (len(tile_cut) * np.ceil(n_items / 128))))
