"""Minimal driver for profiling the solve kernel: N=25 cavity saddle-point LU, a few solves."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optconpy_b200 import problems as pb, device as dv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
k = int(sys.argv[2]) if len(sys.argv) > 2 else 66
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
p = pb.drivcav_problem(N, 5e-3)
M, A, J = p['M'], p['A'], p['J']
Nc = pb.convection_matrix(p, pb.analytic_vortex)
Ft = -(0.5*M.T + 2e-3*(A.T + Nc.T))
K = dv.sadpnt_matrix(Ft - 1.0*M.T, J)
dv.LU(K)             # first factorisation of the pattern: minimum-degree ordering, cached
lu = dv.LU(K)        # what every later shift / time step sees: reused ordering
B = torch.randn((p['NV'], k), dtype=torch.float64, device='cuda')
for _ in range(reps):
    X = lu.solve(B, nrows_out=p['NV'])
torch.cuda.synchronize()
print('ok', lu.info)
