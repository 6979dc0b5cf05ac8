"""Backward implicit-Euler stepper for the differential-algebraic Riccati equation.

Python-3 restatement of the *caller* of the hot path, ``solve_flow_daeric``
(reference ``solve_dae_ric.py:7-213``; it cannot be imported: Python 2 and it
needs ``dolfin_navier_scipy``).  The numerical work is delegated to two
modules with the reference's ``lin_alg_utils`` / ``proj_ric_utils`` interface,
passed in as ``lau`` / ``pru``: by default the CUDA-backed ones of this package;
the tests hand in the CPU oracle to obtain the reference trajectory with the
very same driver.

Semantics kept from the reference, in its order of operations:
  * consistency check of C^T (``solve_dae_ric.py:75-83``),
  * terminal values ``Zc = sqrt(gamma) M^-1 C~^T``, ``mtxtb = -M^T Zc Zc^T B~``
    (``:92-101``), ``w(T) = M^-T gamma C~^T y*(T)`` (``:107-108``),
  * per step: ``F_k^T = -(M^T/2 + tau (A^T + N^T))``, ``W_k = [M^T Zc, sqrt(tau) C~^T]``,
    Newton-ADI, compression (``:147-165``); then the feed-forward solve
    (``:173-194``) — including the quirk that the accumulated outer-Newton gain
    ``cnsmtxtb`` is updated with the PREVIOUS step's ``mtxtb`` (``:181``) before
    ``mtxtb`` is recomputed (``:189``),
  * memoisation: a step whose ``__Z`` entry already exists is not recomputed
    (``:143-145``).
Storage goes through a ``store`` object (``NpyStore`` = the reference's
``dou.save_npa`` / ``dou.load_npa`` on ``.npy`` files; ``MemStore`` keeps arrays
in memory).
"""
import os
import numpy as np

__all__ = ['NpyStore', 'MemStore', 'solve_flow_daeric', 'default_datastr']


class NpyStore(object):
    """``dou.save_npa`` / ``dou.load_npa`` shim: ``np.save`` appends ``.npy``
    (reference ``optcont_main.py:231`` loads ``veldict[t]+'.npy'``)."""

    def save(self, arr, fstring):
        d = os.path.dirname(fstring)
        if d and not os.path.isdir(d):
            os.makedirs(d)
        np.save(fstring, arr)

    def load(self, fstring):
        try:
            return np.load(fstring + '.npy')
        except (IOError, OSError):
            raise IOError(fstring)


class MemStore(dict):
    def save(self, arr, fstring):
        self[fstring] = None if arr is None else np.array(arr, copy=True)

    def load(self, fstring):
        try:
            return self[fstring]
        except KeyError:
            raise IOError(fstring)


def default_datastr(time=None, meshp=None, nu=None, Nts=None, data_prfx='', **kw):
    """``get_datastr`` of ``optcont_main.py:153-157``."""
    return (data_prfx + 'time{0}_nu{1}_mesh{2}_Nts{3}').format(time, nu, meshp, Nts)


def _terminal(lau, pru, mmat, bmat, cmat, mcmat, v_is_my, rmat, vmat, gamma):
    if v_is_my and mcmat is not None:
        tct = lau.apply_invsqrt_fromright(vmat, mcmat.T, output='dense')
    else:
        tct = lau.apply_sqrt_fromright(vmat, cmat.T, output='dense')
    tb = lau.apply_invsqrt_fromright(rmat, bmat, output='sparse')
    zc = np.sqrt(gamma)*lau.apply_massinv(mmat, tct)
    return tct, tb, zc


def solve_flow_daeric(mmat=None, amat=None, jmat=None, bmat=None,
                      cmat=None, rhsv=None, rhsp=None,
                      mcmat=None, v_is_my=False,
                      rmat=None, vmat=None,
                      gamma=1.0,
                      tmesh=None, ystarvec=None,
                      nwtn_adi_dict=None,
                      curnwtnsdict=None,
                      comprz_thresh=None, comprz_maxc=None, save_full_z=False,
                      get_tdpart=None, gttdprtargs=None,
                      get_datastr=None, gtdtstrargs=None,
                      check_c_consist=True,
                      lau=None, pru=None, store=None, verbose=False,
                      stepinfo=None, step_callback=None, lookahead=4, timing=None,
                      private_extensions=True, overlap_tail=True):
    """Same keyword signature as the reference's ``solve_flow_daeric`` plus
    ``lau``/``pru`` (backend modules), ``store`` and ``stepinfo`` (optional list
    that receives per-step diagnostics).  Returns the ``feedbackthroughdict``
    ``{t: dict(w=..., mtxtb=...)}`` of store keys.

    ``lookahead`` (number of steps, 0 = off): the coefficient matrices of a time step depend
    on ``t`` only, not on the Riccati solution, so when the backend offers
    ``pru.factors_async`` / ``lau.sadlu_async`` the sparse LU setup of the next
    ``lookahead`` steps is started (host worker processes) before the device work of step
    ``k``; the numbers are the same with or without it.

    ``overlap_tail`` (with look-ahead and a backend that offers ``pru.tail_thread_init``): the
    feed-forward half of step ``k`` runs on a helper thread while this thread drives the Riccati
    half of step ``k-1`` (independent chains); same numbers, ``step_callback`` is then called from
    that thread.

    ``private_extensions=False`` (with ``lookahead=0``): call the backend through the
    reference's signatures ONLY - no ``_factors`` / ``_lazy_zfac`` / ``sadlu`` keywords - i.e.
    exactly what ``solve_dae_ric.py:152-163,192-194`` executes (bench.py ``e2e_plain``)."""
    if lau is None or pru is None:
        from . import lin_alg_utils as _lau, proj_ric_utils as _pru
        lau, pru = lau or _lau, pru or _pru
    store = NpyStore() if store is None else store
    if timing is not None:      # wall seconds of this thread: waiting for look-ahead, storage
        import time as _time

        class _TimedStore(object):
            def __init__(self, inner):
                self.inner = inner

            def save(self, arr, fstring):
                t0 = _time.perf_counter()
                self.inner.save(arr, fstring)
                timing['store_s'] = timing.get('store_s', 0.0) + _time.perf_counter() - t0

            def load(self, fstring):
                return self.inner.load(fstring)
        store = _TimedStore(store)
    get_datastr = default_datastr if get_datastr is None else get_datastr
    gtdtstrargs = {} if gtdtstrargs is None else gtdtstrargs
    gttdprtargs = {} if gttdprtargs is None else gttdprtargs

    if check_c_consist:
        chk = mcmat if (v_is_my and mcmat is not None) else cmat
        if chk is not None:
            mic = lau.apply_massinv(mmat.T, chk.T)
            if np.linalg.norm(jmat @ mic) > 1e-12:
                raise Warning(('mcmat' if chk is mcmat else 'cmat') +
                              '.T needs to be in the kernel of J*M.-1')

    MT, AT, NV = mmat.T.tocsr(), amat.T.tocsr(), amat.shape[0]
    tE = tmesh[-1]
    gtdtstrargs.update(time=tE)
    key = get_datastr(**gtdtstrargs)

    tct_mat, tb_mat, Zc = _terminal(lau, pru, mmat, bmat, cmat, mcmat, v_is_my,
                                    rmat, vmat, gamma)
    mtxtb = -pru.get_mTzzTtb(MT, Zc, tb_mat)
    store.save(Zc, key + '__Z')
    store.save(mtxtb, key + '__mtxtb')
    if ystarvec is not None:
        wc = lau.apply_massinv(MT, gamma*np.dot(mcmat.T, ystarvec(tE)))
        store.save(wc, key + '__w')
    else:
        wc = None
    fbdict = {tE: dict(w=key + '__w', mtxtb=key + '__mtxtb')}
    if curnwtnsdict is not None:
        store.save(wc, curnwtnsdict[tE]['w'])
        store.save(mtxtb, curnwtnsdict[tE]['mtxtb'])

    can_prefetch = (bool(lookahead) and private_extensions and hasattr(pru, 'factors_async')
                    and hasattr(lau, 'sadlu_async'))

    def prepare(tk):
        """Everything of step tk that depends on t only (incl. background factorisations)."""
        t = tmesh[tk]
        cts = tmesh[tk+1] - t
        nmattd, rhsvtd = get_tdpart(time=t, **gttdprtargs)
        NT = nmattd.T
        ft_mat = -(0.5*MT + cts*(AT + NT))
        at_mat = MT + cts*(AT + NT)
        pre = dict(nmattd=nmattd, rhsvtd=rhsvtd, NT=NT, ft_mat=ft_mat, at_mat=at_mat,
                   fac=None, sadlu=None)
        if can_prefetch:
            # block width of the ADI right-hand sides of that step: [M^T Z (Z^T B), M^T Zc, C~^T]
            khint = zc_width[0] + tct_mat.shape[1] + tb_mat.shape[1]
            pre['fac'] = pru.factors_async(mmat=MT, amat=ft_mat, jmat=jmat, transposed=True,
                                           nwtn_adi_dict=nwtn_adi_dict, k_hint=khint)
            pre['sadlu'] = lau.sadlu_async(amat=at_mat, jmat=jmat)
        return pre

    # the preparation of step k-1 runs in a helper thread while this thread drives the device
    # work of step k (the C library and numpy/scipy release the GIL)
    pool = None
    old_switch = None
    if can_prefetch:
        import sys
        from concurrent.futures import ThreadPoolExecutor
        # CUDA's current device is per thread: the helper thread is bound to the device of THIS
        # thread (a fresh thread starts on device 0 whatever the caller selected)
        tinit = getattr(pru, 'lookahead_thread_init', None)
        pool = ThreadPoolExecutor(max_workers=1, initializer=tinit() if tinit is not None else None)
        # this thread drives the GPU with many short blocking calls; each one has to win the
        # GIL back from the helper threads (assembly, pickling), which by default may keep it
        # for 5 ms at a time
        old_switch = sys.getswitchinterval()
        sys.setswitchinterval(1e-4)
    depth = int(lookahead) if can_prefetch else 0
    zc_width = [Zc.shape[1]]        # latest factor width, read by the look-ahead thread
    ahead = {}

    # The two halves of a time step are independent chains: the Riccati part (Newton-ADI,
    # compression) of step k-1 needs Zc of step k only, the feed-forward part ("tail": gain,
    # feed-forward solve, stores) of step k needs Zc of step k and the tail of step k+1.  With a
    # backend that allows a second device-driving thread (``pru.tail_thread_init``) the tail of
    # step k runs on that thread, on its own stream, while this thread already drives the Riccati
    # part of step k-1; the numbers are the same (same operations on the same data, in the same
    # order within each chain).  At most one tail is outstanding.
    tail_pool = None
    if overlap_tail and can_prefetch and hasattr(pru, 'tail_thread_init'):
        from concurrent.futures import ThreadPoolExecutor as _TPE
        tail_pool = _TPE(max_workers=1, initializer=pru.tail_thread_init())
    tail_state = dict(wc=wc, mtxtb=mtxtb)
    pending_tail = [None]
    # A full (generation-2) pass of Python's cycle collector walks every object of the process -
    # millions after importing torch / scipy - with the GIL held: 15-50 ms during which NO Python
    # thread moves (measured: one time step of every run, always the same one, lost that much in
    # whichever call happened to be active).  The objects alive now are long-lived; park them in
    # the permanent generation for the duration of the loop, so that the passes that do happen
    # only look at what the loop itself created.
    frozen = False
    if can_prefetch:
        import gc
        gc.collect()
        gc.freeze()
        frozen = True

    def tail(tk, t, cts, key, pre, Zc, cnsw, cnsmtxtb, info):
        """Feed-forward half of step tk (``solve_dae_ric.py:173-207``)."""
        wc, mtxtb = tail_state['wc'], tail_state['mtxtb']
        rhsvtd = pre['rhsvtd']
        at_mat = pre['at_mat']
        ftilde = rhsvtd + rhsv
        if cnsw is not None:
            ftilde = rhsvtd + rhsv + cnsw
        # NB (reference quirk, solve_dae_ric.py:181): uses the previous step's mtxtb
        cnsmtxtb = cnsmtxtb + mtxtb if cnsmtxtb is not None else mtxtb

        mtxft = pru.get_mTzzTtb(MT, Zc, ftilde)
        fl1 = np.dot(mcmat.T, ystarvec(t))
        rhswc = MT @ wc + cts*(fl1 - mtxft)
        mtxtb = -pru.get_mTzzTtb(MT, Zc, tb_mat)
        xkw = dict(sadlu=pre['sadlu']) if pre['sadlu'] is not None else {}
        wc = lau.solve_sadpnt_smw(amat=at_mat, jmat=jmat,
                                  umat=cts*cnsmtxtb, vmat=tb_mat.T,
                                  rhsv=rhswc, **xkw)[:NV]

        if curnwtnsdict is not None:
            cnsw = cnsw + wc if cnsw is not None else wc
            cnsmtxtb = cnsmtxtb + mtxtb if cnsmtxtb is not None else mtxtb
            store.save(cnsw, curnwtnsdict[t]['w'])
            store.save(cnsmtxtb, curnwtnsdict[t]['mtxtb'])

        store.save(wc, key + '__w')
        store.save(mtxtb, key + '__mtxtb')
        tail_state['wc'], tail_state['mtxtb'] = wc, mtxtb
        fbdict.update({t: dict(w=key + '__w', mtxtb=key + '__mtxtb')})
        if stepinfo is not None:
            stepinfo.append(info)
        if step_callback is not None:
            step_callback(tk)

    def join_tail():
        if pending_tail[0] is not None:
            f, pending_tail[0] = pending_tail[0], None
            f.result()          # re-raises what the tail raised

    try:
        for tk in range(len(tmesh)-2, -1, -1):
            t = tmesh[tk]
            cts = tmesh[tk+1] - t
            if verbose:
                print('Time is {0}, timestep is {1}'.format(t, cts))
            key = get_datastr(**dict(gtdtstrargs, time=t))
            for tj in range(tk-1, max(tk-1-depth, -1), -1):
                if tj not in ahead:
                    ahead[tj] = pool.submit(prepare, tj)
            if timing is not None:
                _t0 = _time.perf_counter()
            pre = ahead.pop(tk).result() if tk in ahead else prepare(tk)
            if timing is not None:
                timing['prepare_wait_s'] = timing.get('prepare_wait_s', 0.0) + _time.perf_counter() - _t0

            cnsw, cnsmtxtb = None, None
            if curnwtnsdict is not None:
                # written by the tail of this very time step in an earlier sweep only: no
                # dependence on the tail that may still be running
                try:
                    cnsw = store.load(curnwtnsdict[t]['w'])
                    cnsmtxtb = store.load(curnwtnsdict[t]['mtxtb'])
                except IOError:
                    cnsw, cnsmtxtb = None, None

            info = dict(t=t, tau=cts)
            try:
                Zc = store.load(key + '__Z')
            except IOError:
                ft_mat = pre['ft_mat']
                w_mat = np.hstack([MT @ Zc, np.sqrt(cts)*tct_mat])
                oldfb = np.sqrt(cts)*cnsmtxtb if cnsmtxtb is not None else None
                xkw = dict(_factors=pre['fac']) if pre['fac'] is not None else {}
                if private_extensions and hasattr(pru, 'DeviceFactor'):
                    # the uncompressed factor is only compressed below: leave it on the device
                    xkw['_lazy_zfac'] = True
                nres = pru.proj_alg_ric_newtonadi(mmat=MT, amat=ft_mat, transposed=True,
                                                  mtxoldb=oldfb, jmat=jmat,
                                                  bmat=np.sqrt(cts)*tb_mat,
                                                  wmat=w_mat, z0=Zc,
                                                  nwtn_adi_dict=nwtn_adi_dict, **xkw)
                Zp = nres['zfac']
                info.update(nwtn_upd_fnorms=nres.get('nwtn_upd_fnorms'),
                            adi_steps=nres.get('adi_steps'), zp_cols=Zp.shape[1])
                if comprz_maxc is not None or comprz_thresh is not None:
                    Zc = pru.compress_Zsvd(Zp, thresh=comprz_thresh, k=comprz_maxc)
                else:
                    Zc = Zp = np.asarray(Zp)
                store.save(np.asarray(Zp) if save_full_z else Zc, key + '__Z')
            info.update(zc_cols=Zc.shape[1])
            zc_width[0] = Zc.shape[1]

            join_tail()                       # the tail of the previous step (it ran beside this Riccati part)
            if tail_pool is not None:
                pending_tail[0] = tail_pool.submit(tail, tk, t, cts, key, pre, Zc, cnsw, cnsmtxtb, info)
            else:
                tail(tk, t, cts, key, pre, Zc, cnsw, cnsmtxtb, info)
        join_tail()
    finally:
        if tail_pool is not None:
            tail_pool.shutdown(wait=True)
        if frozen:
            gc.unfreeze()
    if pool is not None:
        pool.shutdown(wait=True)
        sys.setswitchinterval(old_switch)
    return fbdict
