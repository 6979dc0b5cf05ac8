"""Quick device timings of the individual kernels (CUDA events) on the config-2 shapes."""
import json
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from optconpy_b200 import problems as pb, device as dv


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    out = {}
    p = pb.drivcav_problem(N, 5e-3)
    M, A, J = p['M'], p['A'], p['J']
    NV, NP = p['NV'], p['NP']
    Nc = pb.convection_matrix(p, pb.analytic_vortex)
    Ft = -(0.5*M.T + 2e-3*(A.T + Nc.T))
    S = dv.sadpnt_matrix(Ft - 1.0*M.T, J)
    t = time.perf_counter()
    lu = dv.LU(S)
    out['lu_setup_s'] = time.perf_counter() - t
    out['lu_info'] = lu.info
    for opts, name in (({}, 'colamd_default'),):
        lu2 = dv.LU(S, lu_options=opts)
        out['lu_info_' + name] = lu2.info
        B = torch.randn((NV, 66), dtype=torch.float64, device='cuda')
        out['solve66_ms_' + name] = timeit(lambda: lu2.solve(B, nrows_out=NV))
    for k in (8, 24, 66, 116, 256):
        B = torch.randn((NV, k), dtype=torch.float64, device='cuda')
        ms = timeit(lambda: lu.solve(B, nrows_out=NV))
        out['solve_k%d_ms' % k] = ms
        out['solve_k%d_GBs' % k] = lu.algorithmic_bytes(k)/ms/1e6
    Md = dv.DeviceCSR(M.T)
    V = torch.randn((NV, 66), dtype=torch.float64, device='cuda')
    out['spmm66_ms'] = timeit(lambda: Md.matmul(V))
    Z = torch.randn((NV, 2000), dtype=torch.float64, device='cuda')
    out['gram_2000x66_ms'] = timeit(lambda: dv.gram(Z, V))
    Zs = torch.randn((NV, 256), dtype=torch.float64, device='cuda')
    out['gram_256x256_ms'] = timeit(lambda: dv.gram(Zs, Zs))
    G = dv.gram(Zs, Zs)
    out['eig256_ms'] = timeit(lambda: dv.sym_eig(G), reps=3, warm=1)
    G64 = G[:64, :64].contiguous()
    out['eig64_ms'] = timeit(lambda: dv.sym_eig(G64), reps=3, warm=1)
    # low-rank Z for compress
    R = torch.randn((NV, 150), dtype=torch.float64, device='cuda') * \
        torch.logspace(0, -9, 150, dtype=torch.float64, device='cuda')[None, :]
    Zl = R @ torch.randn((150, 3000), dtype=torch.float64, device='cuda')
    out['compress_3000_ms'] = timeit(lambda: dv.compress(Zl, thresh=5e-5, k=50), reps=3, warm=1)
    _, info = dv.compress(Zl, thresh=5e-5, k=50)
    out['compress_info'] = dict(kept=info['kept'], chol_rank=info['chol_rank'], sweeps=info['sweeps'])
    print(json.dumps(out, indent=1, default=str))


if __name__ == '__main__':
    main()
