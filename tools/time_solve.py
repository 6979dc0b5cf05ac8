"""Time the multi-RHS saddle-point solve kernel alone (CUDA events, inputs resident):
   python tools/time_solve.py [N] [k,k,...] [reps]"""
import os
import sys
import json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse.linalg as spsla
import torch
from optconpy_b200 import problems as pb, device as dv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
ks = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [8, 35, 66, 148, 256]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
p = pb.drivcav_problem(N, 5e-3)
M, A, J = p['M'], p['A'], p['J']
Nc = pb.convection_matrix(p, pb.analytic_vortex)
Ft = -(0.5*M.T + 2e-3*(A.T + Nc.T))
K = dv.sadpnt_matrix(Ft - 1.0*M.T, J)
dv.LU(K)   # caches the ordering of this pattern; the timed factor below reuses it
lu = dv.LU(K, wide=bool(os.environ.get('OCB_TIME_WIDE')))
print(json.dumps(lu.info))
n = K.shape[0]
ref_lu = spsla.splu(K)
for k in ks:
    B = torch.randn((n, k), dtype=torch.float64, device='cuda')
    X = lu.solve(B)
    torch.cuda.synchronize()
    ref = ref_lu.solve(B.cpu().numpy())
    err = np.linalg.norm(X.cpu().numpy() - ref)/np.linalg.norm(ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lu.solve(B, out=X)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3*e0.elapsed_time(e1)/reps
    ab = lu.algorithmic_bytes(k)
    print('k=%4d  %9.1f us/solve  %8.1f GB/s algorithmic  relerr %.1e  KP env %s'
          % (k, us, ab/us/1e3, err, os.environ.get('OCB_SPTRSM_KP')))
