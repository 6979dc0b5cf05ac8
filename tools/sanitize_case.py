"""Small invocation of every solve executor for compute-sanitizer (memcheck / racecheck):
cluster column-panel kernel with clusters of 4 and 2 (one and two columns per panel), the
persistent panel executor, the fused ADI update and the Gram / compression kernels.
   compute-sanitizer --tool memcheck  python tools/sanitize_case.py
   compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('OCB_LU_WORKERS', '0')

if __name__ == '__main__':
    import numpy as np
    import scipy.sparse.linalg as spsla
    import torch
    from optconpy_b200 import problems as pb, device as dv
    import optconpy_b200.proj_ric_utils as gpru
    prob = pb.drivcav_problem(6, 1e-2)
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    Ft = -(0.5*M.T + 0.05*(A.T + Nc.T))
    K = dv.sadpnt_matrix(Ft - 1.0*M.T, J)
    n = K.shape[0]
    ref = spsla.splu(K)
    rng = np.random.default_rng(0)
    worst = 0.0
    for k_hint, ks in ((8, (1, 3)), (40, (40, 90))):          # clusters of 4, then of 2 (KP = 1 and 2)
        lu = dv.FactorJob([K], k_hint=k_hint).result()[0]
        for k in ks:
            B = rng.standard_normal((n, k))
            X = dv.to_host(lu.solve(dv.to_dev(B)))
            worst = max(worst, np.linalg.norm(X - ref.solve(B))/np.linalg.norm(X))
    luw = dv.LU(K, wide=True)                                  # persistent panel executor
    for k in (644, 700):
        B = rng.standard_normal((n, k))
        X = dv.to_host(luw.solve(dv.to_dev(B)))
        worst = max(worst, np.linalg.norm(X - ref.solve(B))/np.linalg.norm(X))
    W = rng.standard_normal((prob['NV'], 3))
    d = dict(adi_max_steps=12, adi_newZ_reltol=1e-6, ms=[-5.0, -2.0, -1.0])
    res = gpru.solve_proj_lyap_stein(amat=Ft, mmat=M.T, jmat=J, wmat=W, transposed=True, adi_dict=d)
    zc = gpru.compress_Zsvd(res['zfac'], thresh=1e-6)
    torch.cuda.synchronize()
    print('sanitize_case ok: worst solve error %.2e, adi steps %d, compressed %s -> %s'
          % (worst, len(res['adi_rel_newZ_norms']), res['zfac'].shape, zc.shape))
    assert worst < 1e-10
