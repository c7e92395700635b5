import numpy as np, torch, sys, time
sys.path.insert(0,'/root/repo')
from polara_b200.engine import get_engine
from tests.helpers import random_seen_csr
eng = get_engine(0)
rng = np.random.default_rng(6)
for (m,n,r,k) in [(128,256,16,10),(128,1024,50,10),(700,20000,50,10),(5000,100000,50,10)]:
    e = rng.standard_normal((m, r)).astype(np.float32)
    v = rng.standard_normal((n, r)).astype(np.float32)
    rows, cols, indptr = random_seen_csr(rng, m, n, rng.integers(0, min(n//2,200), size=m))
    e_dev, v_dev = eng.upload(e), eng.upload(v)
    seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32)))
    out = {}
    for kernel in ("simt", "tcgen05"):
        eng.set_score_kernel(kernel)
        s0 = eng.stats()
        ids, sc = eng.score_topk(e_dev, v_dev, r, k, seen=seen, want_scores=True)
        torch.cuda.synchronize()
        st = eng.stats()
        out[kernel] = (ids.cpu().numpy(), sc.cpu().numpy())
        print(kernel, (m,n,r,k), 'kernel_us', st[4], 'rescored', st[1]-s0[1], 'err', hex(st[7]))
    same = np.array_equal(out["simt"][0], out["tcgen05"][0]); same_s = np.array_equal(out["simt"][1], out["tcgen05"][1])
    print('  ids equal', same, 'scores equal', same_s, 'mismatch rows', int((out["simt"][0]!=out["tcgen05"][0]).any(1).sum()))
    if not same:
        bad = np.flatnonzero((out["simt"][0]!=out["tcgen05"][0]).any(1))[:3]
        for b in bad:
            print('   row', b, out["simt"][0][b], out["tcgen05"][0][b]); print('      ', out["simt"][1][b], out["tcgen05"][1][b])
