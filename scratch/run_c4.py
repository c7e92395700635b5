"""C4-shaped CoFFee build + scoring on one B200 (timing + sanity; the reference cannot run at this size)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.argv = ['x']
import bench
from polara_b200.host import ArrayData
from polara_b200.models import B200CoffeeModel
dev = torch.device('cuda', 0)
U, I, F, NNZ = 1_000_000, 50_000, 5, 50_000_000
indptr, indices, values = bench.synth_csr_torch(U, I, int(NNZ * 1.25), 7, dev)
nnz = indices.shape[0]
rows = torch.repeat_interleave(torch.arange(U, device=dev), indptr[1:] - indptr[:-1])
fdbk = (values - 1).to(torch.int64)                       # ratings 1..5 -> levels 0..4
idx = torch.stack([rows, indices.to(torch.int64), fdbk], 1).cpu().numpy()
val = np.ones(nnz)
print('tensor nnz', nnz)
# test users = first 200K users, all their triplets (known-user scenario)
m_test = 200_000
hi = int(indptr[m_test])
data = ArrayData(idx, val, (U, I, F), idx[:hi, 0], idx[:hi, 1], idx[:hi, 2], (m_test, I, F), n_feedback=F)
model = B200CoffeeModel(data); model.verbose = False
model.mlrank = (60, 60, 4); model.seed = 0; model.num_iters = 5
t0 = time.perf_counter(); model.build(); torch.cuda.synchronize(); t1 = time.perf_counter()
print('hooi build s', round(t1 - t0, 3), 'core norm trace', [round(x, 3) for x in model.core_norm_trace])
u1 = model.factors['itemid']; print('item factor orthonormality err', float(np.abs(u1.T @ u1 - np.eye(60)).max()))
t0 = time.perf_counter(); recs = model.get_recommendations(); t1 = time.perf_counter()
print('get_recommendations s', round(t1 - t0, 3), recs.shape, 'pairs/s', m_test * I / (t1 - t0))
