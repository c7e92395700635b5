import sys, os, numpy as np, torch
sys.path.insert(0,'/root/repo')
from polara_b200.engine import get_engine
eng = get_engine(0)
rng = np.random.default_rng(0)
m, n, r, k = 148*128*4, 100000, 50, 10
e = torch.randn(m, 64, device='cuda'); e[:, r:] = 0
v = torch.randn(n, 64, device='cuda') * torch.rand(n,1,device='cuda')**8; v[:, r:] = 0
eng.set_score_kernel('tcgen05')
for it in range(2):
    ids = eng.score_topk(e, v, r, k)
    torch.cuda.synchronize(); st = eng.stats()
print('dbg', os.environ.get('PB200_TC_DEBUG'), 'main kernel ms', st[4]/1000, 'tiles/SM', 4*391, 'cycles/tile @1.9GHz', st[4]*1e-6*1.9e9/(4*391))
