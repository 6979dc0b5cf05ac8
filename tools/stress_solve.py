"""Determinism / race stress of the solve kernel: the same solve repeated must be bit-identical
(fixed reduction order) and equal to SuperLU to rounding.  python tools/stress_solve.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse.linalg as spsla
import torch
from optconpy_b200 import problems as pb, device as dv

bad = 0
for name, prob in (('cav10', pb.drivcav_problem(10, 1e-2)), ('chan22x8', pb.channel_problem(22, 8, 2e-2)),
                   ('cav25', pb.drivcav_problem(25, 5e-3))):
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex) if 'cav' in name else 0*M
    K = dv.sadpnt_matrix(-(0.5*M.T + 2e-2*(A.T + Nc.T)) - 1.0*M.T, J)
    lu = dv.LU(K)
    ref_lu = spsla.splu(K)
    n = K.shape[0]
    for k in (1, 4, 8, 16, 33, 41, 70):
        B = torch.randn((n, k), dtype=torch.float64, device='cuda')
        X0 = lu.solve(B).clone()
        ref = ref_lu.solve(B.cpu().numpy())
        err = np.linalg.norm(X0.cpu().numpy() - ref)/np.linalg.norm(ref)
        nbad = 0
        for it in range(300):
            X = lu.solve(B)
            if not torch.equal(X, X0):
                nbad += 1
        bad += nbad + (err > 1e-10)
        print('%-9s n=%5d k=%3d relerr %.1e  non-identical repeats %d/300' % (name, n, k, err, nbad))
print('STRESS', 'FAILED' if bad else 'ok')
