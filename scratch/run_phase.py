import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo'); sys.argv=['x']
import bench
from polara_b200.engine import get_engine, DeviceCSR
from polara_b200.host import ArrayData
from polara_b200.models import B200SVDModel
eng = get_engine(0); dev = torch.device('cuda',0)
U, I, NNZ, R = 1000000, 100000, 100000000, 50
indptr, indices, values = bench.synth_csr_torch(U, I, int(NNZ*1.3), 20260924, dev)
shape=(U,I)
data = ArrayData(np.zeros((1,2),dtype=np.int64), np.ones(1), shape)
data.train_csr = (indptr.cpu(), indices.cpu(), values.cpu(), shape)
model = B200SVDModel(data); model.verbose=False; model.rank=R
model.build()
print('build', model.last_timings, 'nnz', indices.shape[0])
v_dev = model._device_factor('itemid'); p_dev = DeviceCSR(indptr, indices, values, shape)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r=fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n, r
ms, e = t(lambda: eng.spmm(p_dev, v_dev, ell=v_dev.shape[1])); print('spmm E=P.V ms', ms)
seen=(indptr, indices)
import os
for filt in (True,):
    s0=eng.stats(); ms, ids = t(lambda: eng.score_topk(e, v_dev, R, 10, seen=seen if filt else None)); st=eng.stats()
    print('dbg', os.environ.get('PB200_TC_DEBUG'), 'cluster', os.environ.get('PB200_TC_CLUSTER'), 'score_topk total ms', ms, 'filter', filt, 'main kernel ms', st[4]/1000, 'rescored/user/call', (st[1]-s0[1])/U/4, 'dbgstats', [st[i]-s0[i] for i in (2,3,5,6)])
