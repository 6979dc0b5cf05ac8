"""Times the all-columns-at-once (panel) executor on one cavity saddle-point factor:
   python tools/wide_bench.py MESH_N k1,k2,...   (env: OCB_PANEL_T, OCB_PANEL_U, OCB_WIDE_LEGACY)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == '__main__':
    import torch
    from optconpy_b200 import problems as pb, device as dv
    N = int(sys.argv[1])
    ks = [int(v) for v in sys.argv[2].split(',')]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    prob = pb.drivcav_problem(N, 5e-3)
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    K = dv.sadpnt_matrix(-(0.5*M.T + 2e-3*(A.T + Nc.T)) - 1.0*M.T, J)
    t0 = time.perf_counter()
    lu = dv.LU(K, wide=True)
    setup = time.perf_counter() - t0
    n = K.shape[0]
    i = lu.info
    peak = dv.fp64_peak('dfma')
    if os.environ.get('OCB_DUMP_LEVELS'):
        import ctypes as C
        import numpy as np
        from optconpy_b200 import _cabi
        lib = _cabi.load()
        ns = lib.ocb_lu_panel_levels(lu.handle, None, 0)
        arr = np.zeros((max(ns, 1), 3), dtype=np.int64)
        lib.ocb_lu_panel_levels(lu.handle, arr.ctypes.data, ns)
        np.save(os.environ['OCB_DUMP_LEVELS'], arr)
    for k in ks:
        B = torch.randn((n, k), dtype=torch.float64, device='cuda')
        X = lu.solve(B)
        res = float(torch.linalg.norm(dv.DeviceCSR(K).matmul(X) - B)/torch.linalg.norm(B))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lu.solve(B, out=X)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/reps
        fl = 2.0*(i['nnzL'] + i['nnzU'])*k
        ab = 12*(i['nnzL'] + i['nnzU']) + 16*(n + 1) + 32*n*k
        print(json.dumps(dict(env={e: os.environ.get(e) for e in ('OCB_PANEL_T', 'OCB_PANEL_U', 'OCB_WIDE_LEGACY')
                                   if os.environ.get(e)},
                              n=n, k=k, nnz_LU=int(i['nnzL'] + i['nnzU']), sublevels=int(i['levelsL'] + i['levelsU']),
                              ms=round(ms, 3), fp64_TFs=round(fl/ms/1e9, 3), frac_fp64=round(fl/ms/1e9/peak, 4),
                              alg_GBs=round(ab/ms/1e6, 1), residual=res, setup_s=round(setup, 2))), flush=True)
