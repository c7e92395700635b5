import numpy as np, torch, sys, time
sys.path.insert(0,'/root/repo')
sys.argv=['x']
import bench
from polara_b200.engine import get_engine, DeviceCSR
from polara_b200.host import ArrayData
from polara_b200.models import B200SVDModel
eng = get_engine(0)
dev = torch.device('cuda',0)
U, I, NNZ, R = 100000, 100000, 10000000, 50
indptr, indices, values = bench.synth_csr_torch(U, I, NNZ, 1, dev)
shape=(U,I)
data = ArrayData(np.zeros((1,2),dtype=np.int64), np.ones(1), shape)
data.train_csr = (indptr.cpu(), indices.cpu(), values.cpu(), shape)
model = B200SVDModel(data); model.verbose=False; model.rank=R
model.build()
v_dev = model._device_factor('itemid')
p_dev = DeviceCSR(indptr, indices, values, shape)
e = eng.spmm(p_dev, v_dev, ell=v_dev.shape[1])
vn = torch.linalg.norm(v_dev[:, :R], dim=1); en = torch.linalg.norm(e[:, :R], dim=1)
print('vmax', float(vn.max()), 'v median', float(vn.median()), 'enorm median', float(en.median()))
seen=(indptr, indices)
for kernel in ('simt','tcgen05'):
    eng.set_score_kernel(kernel)
    for filt in (True, False):
        s0 = eng.stats()
        ids, sc = eng.score_topk(e, v_dev, R, 10, seen=seen if filt else None, want_scores=True)
        torch.cuda.synchronize(); st = eng.stats()
        print(kernel, 'filter', filt, 'kernel_us', st[4], 'rescored/user', (st[1]-s0[1])/U, 'err', hex(st[7]))
        if kernel=='simt': ref=(ids.clone(), sc.clone(), filt)
    
print('kth score median', float(sc[:,9].median()), 'top score median', float(sc[:,0].median()))
