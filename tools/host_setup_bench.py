"""Host cost of one shifted saddle-point factorisation, step by step (CPU only, no GPU):
SuperLU path vs. numeric-only refactorisation, and the analyse+pack step.  Minimum of N runs
(the boxes are shared: the median is noisy).

    python tools/host_setup_bench.py [--mesh 25] [--reps 15]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mesh', type=int, default=25)
    ap.add_argument('--reps', type=int, default=15)
    args = ap.parse_args()
    from optconpy_b200 import _lu_worker as w, device as dv, problems as pb
    prob = pb.drivcav_problem(args.mesh, 5e-3)
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)

    def mat(tau, p):
        return dv.sadpnt_matrix(-(0.5*M.T + tau*(A.T + Nc.T)) + p*M.T, J)
    opts = dict(dv.LU_OPTIONS)
    flags = dv._pack_flags(False, 40)
    a0 = dv._csc_args(mat(2e-3, -1.0), opts) + (232448, flags)
    q = w.order_only(a0)
    piv = w.static_pivots(a0 + (q,))
    shifts = [(3e-4, -5.0), (1e-3, -1.3), (2e-4, -3.0), (2.5e-4, -1.1)]
    jobs = [dv._csc_args(mat(t, p), opts) + (232448, flags, q, piv) for t, p in shifts]

    def best(fn):
        ts = []
        for i in range(args.reps):
            t0 = time.perf_counter()
            out = fn(jobs[i % len(jobs)])
            ts.append(time.perf_counter() - t0)
        return 1e3*min(ts), 1e3*float(np.median(ts)), out
    os.environ['OCB_REFACTOR'] = '0'
    w._build(jobs[0])
    slu = best(lambda a: w._build(a))
    r = slu[2]
    print('SuperLU path      : min %6.1f ms  median %6.1f   (factor %.1f + pack %.1f ms in the last run; %d bytes)'
          % (slu[0], slu[1], 1e3*r[3], 1e3*r[4], r[2]))
    os.environ['OCB_REFACTOR'] = '1'
    w._build(jobs[0])
    w._build(jobs[1])
    st = best(lambda a: w._build(a))
    r = st[2]
    assert r[6][2] == 'static', r[6]
    print('static-pivot path : min %6.1f ms  median %6.1f   (numeric %.1f + pack %.1f ms in the last run; %d bytes; backward error %.1e)'
          % (st[0], st[1], 1e3*r[3], 1e3*r[4], r[2], r[6][0]))
    rf = next(iter(w._REFAC.values()))
    nm = best(lambda a: rf.numeric(a[0]))
    print('  numeric refactor: min %6.1f ms  median %6.1f   (%d supernodes, largest front %d, %.0f MFLOP -> %.1f GFLOP/s)'
          % (nm[0], nm[1], rf.info['supernodes'], rf.info['max_front'], rf.info['flops']/1e6,
             rf.info['flops']/nm[0]/1e6))
    arrs = rf.numeric(jobs[0][0])
    am = w._amat_arrays(jobs[0])
    from optconpy_b200 import _cabi
    lib = _cabi.load()
    pk = best(lambda a: w._pack(lib, arrs, a[3][0], 232448, flags & ~2, None, None))
    pg = best(lambda a: w._pack(lib, arrs, a[3][0], 232448, flags & ~2, None, am))
    print('  analyse + pack  : min %6.1f ms  median %6.1f   (with the residual guard: min %.1f ms)'
          % (pk[0], pk[1], pg[0]))
    import ctypes as C
    buf = np.zeros(pk[2][2] + 4096, dtype=np.uint8)
    nb, be = C.c_int64(0), C.c_double(0.0)

    def into(a):
        return lib.ocb_lu_pack_host_checked(a[3][0], *[x.ctypes.data for x in arrs], 232448, flags & ~2,
                                            buf.ctypes.data, buf.nbytes, None, C.byref(nb),
                                            *[x.ctypes.data for x in am], C.byref(be))
    pi = best(into)
    print('  ... guarded, into a pinned-pool segment that holds the structure already: min %.1f ms  median %.1f'
          % (pi[0], pi[1]))


if __name__ == '__main__':
    main()
