import numpy as np, torch, sys, os
sys.path.insert(0,'/root/repo')
from polara_b200.engine import get_engine
from tests.helpers import random_seen_csr
eng = get_engine(0)
rng = np.random.default_rng(6)
m,n,r,k = 700,20000,50,10
e = rng.standard_normal((m, r)).astype(np.float32); v = rng.standard_normal((n, r)).astype(np.float32)
rows, cols, indptr = random_seen_csr(rng, m, n, rng.integers(0, 200, size=m))
e_dev, v_dev = eng.upload(e), eng.upload(v); seen = (eng.upload(indptr), eng.upload(cols.astype(np.int32)))
eng.set_score_kernel('simt'); ref = eng.score_topk(e_dev, v_dev, r, k, seen=seen).cpu().numpy()
eng.set_score_kernel('tcgen05')
try:
    ids = eng.score_topk(e_dev, v_dev, r, k, seen=seen); torch.cuda.synchronize()
    print(os.environ.get('PB200_TC_CLUSTER'), os.environ.get('PB200_TC_ABUFS'), 'equal', np.array_equal(ids.cpu().numpy(), ref))
except Exception as ex:
    print(os.environ.get('PB200_TC_CLUSTER'), os.environ.get('PB200_TC_ABUFS'), 'FAILED', str(ex)[:100])
    import ctypes
