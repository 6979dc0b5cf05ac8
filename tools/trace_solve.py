"""Per-sub-level timing of one solve (OCB_TRSM_TRACE=1): cycles of CTA 0 between barriers,
next to the size of the sub-level.  python tools/trace_solve.py [N] [k]"""
import os
import sys
os.environ['OCB_TRSM_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import ctypes as C
import numpy as np
import torch
from optconpy_b200 import problems as pb, device as dv, _cabi, _lu_worker
from test_lu_program import _program

N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
k = int(sys.argv[2]) if len(sys.argv) > 2 else 24
p = pb.drivcav_problem(N, 5e-3)
M, A, J = p['M'], p['A'], p['J']
Nc = pb.convection_matrix(p, pb.analytic_vortex)
Ft = -(0.5*M.T + 2e-3*(A.T + Nc.T))
K = dv.sadpnt_matrix(Ft - 1.0*M.T, J)
arrs = _lu_worker.factor_arrays(dv._csc_args(K, dict(dv.LU_OPTIONS)))
prog = _program(arrs, K.shape[0])
info, sub_ptr, slices = prog[0], prog[1], prog[2]
lu = dv.LU(K)
B = torch.randn((K.shape[0], k), dtype=torch.float64, device='cuda')
for _ in range(3):
    lu.solve(B)
torch.cuda.synchronize()
nsub = len(sub_ptr) - 1
buf = np.zeros(nsub + 1, dtype=np.int64)
_cabi.check(_cabi.load().ocb_debug_trace(buf.ctypes.data, nsub + 1), 'trace')
d = np.diff(buf)
print('sub-levels', nsub, 'total cycles', buf[-1] - buf[0], '= %.1f us at 1.965 GHz' % ((buf[-1]-buf[0])/1965.0))
print('%4s %7s %6s %7s %8s %8s' % ('sub', 'slices', 'rows', 'trips', 'maxtrip', 'cycles'))
for sb in range(nsub):
    sl = slices[sub_ptr[sb]:sub_ptr[sb+1]]
    print('%4d %7d %6d %7d %8d %8d' % (sb, len(sl), (sl[:, 2] >> 8).sum(), sl[:, 1].sum(), sl[:, 1].max(), d[sb]))
