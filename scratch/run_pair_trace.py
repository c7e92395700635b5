import sys, os, numpy as np, torch
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
v_dev = model._device_factor('itemid'); p_dev = DeviceCSR(indptr, indices, values, shape)
e = eng.spmm(p_dev, v_dev, ell=R)
seen=(indptr, indices)
ref = None
for pair in ("0", "1"):
    os.environ["PB200_TC_PAIR"] = pair
    os.environ.pop("PB200_TC_TRACE", None)
    for _ in range(2):
        ids = eng.score_topk(e, v_dev, R, 10, seen=seen)
    ms = []
    for _ in range(3):
        ids = eng.score_topk(e, v_dev, R, 10, seen=seen); ms.append(eng.last_score_kernel_ms())
    if ref is None: ref = ids.clone()
    print("pair", pair, "kernel ms", ms, "equal to pair=0:", bool(torch.equal(ids, ref)), "differing rows", int((ids != ref).any(1).sum()))
    os.environ["PB200_TC_TRACE"] = "/root/repo/gpurun_out/trace_pair%s.txt" % pair
    eng.score_topk(e, v_dev, R, 10, seen=seen)
