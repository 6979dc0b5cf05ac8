"""Host-side timeline of the reference-facing stepper (bench.py's ``e2e`` leg).

    OCB_TIMELINE=/tmp/tl python tools/e2e_timeline.py [--steps 12] [--lookahead 4]

Runs ``solve_flow_daeric`` on the bench workload and prints, per time step, what the main
thread, the look-ahead thread, the upload thread and the LU worker processes were doing
(wall-clock; the GPU is synchronised at the end of every step as in bench.py).  Diagnostics
only - nothing here is a bench number."""
import argparse
import glob
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=12)
    ap.add_argument('--lookahead', type=int, default=4)
    ap.add_argument('--mesh', type=int, default=25)
    ap.add_argument('--show', type=int, default=3, help='steps printed event by event')
    ap.add_argument('--dump', default=None, help='write every event (and the step stamps) to this file')
    args = ap.parse_args()
    tdir = os.environ.get('OCB_TIMELINE')
    assert tdir, 'set OCB_TIMELINE=<dir>'
    shutil.rmtree(tdir, ignore_errors=True)
    os.makedirs(tdir)
    import torch
    import optconpy_b200.lin_alg_utils as glau
    import optconpy_b200.proj_ric_utils as gpru
    from optconpy_b200 import scenarios as sc, dre_stepper as ds, device as dv
    prob, cs, kw = sc.config2(glau, N=args.mesh)
    S, la = args.steps, args.lookahead
    kw['tmesh'] = kw['tmesh'][-(S+la+1):]
    tmp = tempfile.mkdtemp(prefix='ocb_tl_')
    kw['gtdtstrargs'] = dict(kw['gtdtstrargs'], data_prfx=os.path.join(tmp, 'tdst_'))
    stamps, mem = [], []

    def cb(tk):
        torch.cuda.synchronize()
        stamps.append(time.time())
        ms = torch.cuda.memory_stats()
        mem.append((ms.get('num_device_alloc', 0), ms.get('num_device_free', 0)))
    dv.reset_stats()
    stimes = {}
    try:
        ds.solve_flow_daeric(lau=glau, pru=gpru, store=ds.NpyStore(), step_callback=cb,
                             timing=stimes, lookahead=la, **kw)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ev = list(dv.TIMELINE)
    for f in glob.glob(os.path.join(tdir, 'worker_*.log')):
        pid = os.path.basename(f)[7:-4]
        for line in open(f):
            a, b, tf, tp = [float(x) for x in line.split()]
            ev.append(('worker ' + pid, 'factor %.1f + pack %.1f ms' % (1e3*tf, 1e3*tp), a, b))
    ev.sort(key=lambda e: e[2])
    if args.dump:
        with open(args.dump, 'w') as f:
            for th, label, a, b in ev:
                f.write('%.3f %.3f | %s | %s\n' % (1e3*(a - stamps[0]), 1e3*(b - stamps[0]), th, label))
            for i, s_ in enumerate(stamps):
                f.write('%.3f %.3f | STEP | end of step %d\n' % (1e3*(s_ - stamps[0]), 1e3*(s_ - stamps[0]), i))
    steps_ms = [1e3*(b - a) for a, b in zip(stamps[:-1], stamps[1:])]
    print('step ms:', ' '.join('%.1f' % s for s in steps_ms))
    print('workers %d, lookahead %d' % (dv._POOL['workers'], la))
    print('cudaMalloc/cudaFree counts at the step ends:', mem)
    print('CUDA_MODULE_LOADING =', os.environ.get('CUDA_MODULE_LOADING'))
    # steady-state window: steps la+2 .. S-1
    lo, hi = stamps[min(la + 2, len(stamps) - 2)], stamps[-1 - la] if len(stamps) > 2*la + 3 else stamps[-1]
    nst = sum(1 for s in stamps if lo < s <= hi)
    print('steady window: %d steps, %.1f ms/step' % (nst, 1e3*(hi - lo)/max(nst, 1)))
    busy = {}
    for th, label, a, b in ev:
        a2, b2 = max(a, lo), min(b, hi)
        if b2 > a2:
            key = (th if th.startswith('worker') else th + ' | ' + label.split(' ')[0])
            busy[key] = busy.get(key, 0.0) + b2 - a2
    wk = [v for k, v in busy.items() if k.startswith('worker')]
    print('worker busy fraction in the window: mean %.2f min %.2f max %.2f (n=%d)'
          % (sum(wk)/max(len(wk), 1)/(hi - lo), min(wk)/(hi - lo), max(wk)/(hi - lo), len(wk)))
    for k in sorted(k for k in busy if not k.startswith('worker')):
        print('  %-55s %7.2f ms/step' % (k, 1e3*busy[k]/max(nst, 1)))
    # event list of the slowest step after the start-up and of a few steady steps
    if len(steps_ms) > 3:
        worst = max(range(2, len(steps_ms)), key=lambda i: steps_ms[i])
        a0, b0 = stamps[worst], stamps[worst + 1]
        print('--- slowest step %d (%.1f ms): events overlapping it (ms relative to its start) ---'
              % (worst + 1, steps_ms[worst]))
        for th, label, a, b in ev:
            if b >= a0 and a <= b0:
                print('%9.2f %9.2f  %-28s %s' % (1e3*(a - a0), 1e3*(b - a0), th[:28], label))
    i0 = min(la + 3, len(stamps) - 2)
    a0, b0 = stamps[i0], stamps[min(i0 + args.show, len(stamps) - 1)]
    print('--- events between step stamps %d and %d (ms relative) ---' % (i0, i0 + args.show))
    for th, label, a, b in ev:
        if b >= a0 and a <= b0:
            print('%9.2f %9.2f  %-28s %s' % (1e3*(a - a0), 1e3*(b - a0), th[:28], label))
    for s in stamps:
        if a0 <= s <= b0:
            print('%9.2f            STEP END' % (1e3*(s - a0)))
    print('STATS', {k: round(v/(S+la), 4) for k, v in dv.STATS.items() if k.startswith('lu_')})


if __name__ == '__main__':
    main()
