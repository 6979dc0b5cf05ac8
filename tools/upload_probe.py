"""Where does the upload time of the factor images go while the GPU is busy?  Diagnostic for the
open issue of profiles/r01c_onestep_supernodes.md (with OCB_MERGE=1 the images of the e2e run
upload at 3.5-5 ms each instead of 0.5 ms and short device phases of the main thread stall).

    python tools/upload_probe.py [steps]            # run once with OCB_MERGE=0 and once with 1

Runs the setup pipeline of the DRE driver (FactorJob -> worker processes -> pinned segments ->
uploader thread) for `steps` time steps of the N=25 cavity while the main thread keeps the GPU
busy with solves of an earlier factor, and prints per image: segment path (pinned slot or
one-off segment), seconds in the allocator, seconds in ocb_lu_create_from_image, plus the main
thread's solve times and torch's cudaMalloc count."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import scipy.sparse as sps
    import torch
    from optconpy_b200 import device as dv, scenarios as sc
    import optconpy_b200.lin_alg_utils as glau

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    prob, cs, kw = sc.config2(glau, N=25)
    tm = kw['tmesh']
    M, A, J = prob['M'], prob['A'], prob['J']
    MT, AT = sps.csr_matrix(M.T), sps.csr_matrix(A.T)
    shifts = kw['nwtn_adi_dict']['ms']

    def mats_of(i):
        tk = len(tm) - 2 - i
        t, cts = tm[tk], tm[tk+1] - tm[tk]
        nmattd, _ = kw['get_tdpart'](time=t, **kw['gttdprtargs'])
        NT = sps.csr_matrix(nmattd.T)
        ft = -(0.5*MT + cts*(AT + NT))
        return [dv.sadpnt_matrix(ft + mu*MT, J) for mu in shifts] + [dv.sadpnt_matrix(MT + cts*(AT + NT), J)]

    pre = [mats_of(i) for i in range(steps)]
    # per-image timers around the two halves of LU.__init__
    log = []
    orig_arena = dv._new_arena

    def timed_arena(nbytes):
        t0 = time.perf_counter()
        out = orig_arena(nbytes)
        log.append(['arena', time.perf_counter() - t0, nbytes])
        return out
    dv._new_arena = timed_arena
    busy = dv.LU(pre[0][0])
    n = busy.n
    B = torch.randn((n, 58), dtype=torch.float64, device='cuda')
    X = torch.empty_like(B)
    la = 4
    jobs = {}

    def submit(i):
        jobs[i] = dv.FactorJob(pre[i], k_hint=74).start_upload()
    for i in range(min(la + 1, steps)):
        submit(i)
    m0 = torch.cuda.memory_stats().get('num_device_alloc', 0)
    solve_ms, waits = [], []
    for i in range(steps):
        t0 = time.perf_counter()
        lus = jobs.pop(i).result()
        waits.append(time.perf_counter() - t0)
        if i + la + 1 < steps:
            submit(i + la + 1)
        for _ in range(100):                      # ~ the device work of one time step
            t1 = time.perf_counter()
            lus[0].solve(B, out=X)
            torch.cuda.current_stream().synchronize()
            solve_ms.append(1e3*(time.perf_counter() - t1))
        del lus
    st = dv.STATS
    sm = np.array(solve_ms)
    print('OCB_MERGE=%s  images %d  upload+handle %.2f ms/image (allocator %.2f)  collect wait %.1f ms/step  '
          'main wait %.1f ms/step  solve median %.3f ms p99 %.3f max %.2f  cudaMallocs %d  pinned pool %s'
          % (os.environ.get('OCB_MERGE', dv.MERGE_DEFAULT), st['n_factor'],
             1e3*st['lu_analyse_upload_s']/max(st['n_factor'], 1), 1e3*st['lu_arena_s']/max(st['n_factor'], 1),
             1e3*st['lu_collect_wait_s']/steps, 1e3*np.mean(waits), np.median(sm), np.percentile(sm, 99), sm.max(),
             torch.cuda.memory_stats().get('num_device_alloc', 0) - m0,
             'none' if dv._SHM['pool'] is None else '%d x %.1f MB' % (len(dv._SHM['pool'].segs), dv._SHM['pool'].seg_bytes/1e6)))
    slow = sorted((x for x in log if x[0] == 'arena'), key=lambda x: -x[1])[:5]
    print('   slowest allocator calls (ms):', [round(1e3*x[1], 2) for x in slow])


if __name__ == '__main__':
    if os.environ.get('OCB_PROBE_CHILD') or 'OCB_MERGE' in os.environ:
        main()
    else:
        import subprocess
        for m in ('0', '1'):
            subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:],
                            env=dict(os.environ, OCB_MERGE=m, OCB_PROBE_CHILD='1'))
