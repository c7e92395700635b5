import sys, numpy as np, torch, os
sys.path.insert(0,'/root/repo')
from polara_b200.engine import get_engine
eng = get_engine(0)
m, n, r, k = 148*128*4, 100000, 50, 10
g = torch.Generator(device='cuda'); g.manual_seed(1)
e = torch.randn(m, 64, device='cuda', generator=g); e[:, r:] = 0
v = torch.randn(n, 64, device='cuda', generator=g); v[:, r:] = 0          # flat norms: worst case for the filter
eng.set_score_kernel('tcgen05')
for it in range(2):
    s0 = eng.stats(); ids = eng.score_topk(e, v, r, k); torch.cuda.synchronize(); st = eng.stats()
print('flat-norm worst case: main kernel ms', st[4]/1000, 'rescored/user', (st[1]-s0[1])/m, 'pairs/s', m*n/(st[4]*1e-6))
