"""SpMM timings at C2 shapes (development aid): A Q and A^T W for every kernel variant, plain and panel-major."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from bench import synth_csr_torch
from polara_b200.engine import get_engine, DeviceCSR
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
eng = get_engine(0)
users, items, nnz_t = 1_000_000, 100_000, 100_000_000
ip, ix, vl = synth_csr_torch(users, items, int(nnz_t * 1.2), 20260924, dev)
a = DeviceCSR(ip, ix, vl, (users, items))
nnz = a.nnz
t0 = time.perf_counter(); at = eng.transpose(a); torch.cuda.synchronize(); print('transpose %.1f ms nnz %d' % ((time.perf_counter() - t0) * 1e3, nnz))
peak = 6566.7
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = []
for ell in (64, 96):
    q = torch.randn(items, ell, device=dev); w = torch.randn(users, ell, device=dev)
    yq = torch.empty(users, ell, device=dev); yw = torch.empty(items, ell, device=dev)
    t0 = time.perf_counter(); atb = eng.block_columns(at, eng.panel_cols_for(users, ell)); torch.cuda.synchronize()
    print('block_columns(A^T) ell %d: %d panels, %.1f ms' % (ell, atb.n_panels, (time.perf_counter() - t0) * 1e3))
    ref_q = ref_w = None
    for kern in ('window32', 'window'):
        eng.set_spmm_kernel(kern)
        for name, mat, x, y, rows, cols in (('A.Q', a, q, yq, users, items), ('At.W', at, w, yw, items, users), ('At.W panels', atb, w, yw, items, users)):
            ms = timeit(lambda: eng.spmm(mat, x, out=y))
            alg = 8 * nnz + 8 * (rows + 1) + 4 * ell * (rows + cols)
            gath = nnz * ell * 4
            print('ell %3d %-8s %-12s %7.3f ms  algorithmic %.0f GB/s (%.3f of hbm)  gather %.1f TB/s' % (ell, kern, name, ms, alg / ms / 1e6, alg / ms / 1e6 / peak, gath / ms / 1e9))
            res.append(dict(ell=ell, kernel=kern, product=name, ms=ms, alg_gbs=alg / ms / 1e6, gather_tbs=gath / ms / 1e9))
            yc = y.clone()
            if name == 'A.Q':
                if ref_q is None: ref_q = yc
                else: print('    max rel diff vs ldg', float(((yc - ref_q).abs().max() / ref_q.abs().max())))
            else:
                if ref_w is None: ref_w = yc
                else: print('    max rel diff vs ldg', float(((yc - ref_w).abs().max() / ref_w.abs().max())))
json.dump(res, open('gpurun_out/r2_spmm_bench.json', 'w'))
