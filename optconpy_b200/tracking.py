"""Tracking-cost evaluation and a linear closed-loop simulator.

Host-side (numpy/scipy) post-processing around the hot path, needed to state the
north-star parity claim "the DRE trajectory of the tracking cost matches within 1e-8"
without FEniCS:

* :func:`eval_costfunc` restates ``optcont_main.py:213-264`` (trapezoidal
  ``int dy'W dy + u'u`` plus the terminal penalty ``dy(T)'V dy(T)``) on a ``store``
  (``dre_stepper.NpyStore`` / ``MemStore``) instead of ``np.load`` of path strings;
* :func:`simulate_closed_loop` stands in for ``snu.solve_nse(closed_loop=True,
  feedbackthroughdict=...)`` (``optcont_main.py:610-626``, dolfin_navier_scipy is absent):
  implicit Euler for the linearised (Oseen) flow with the time-varying gain and
  feed-forward the backward DRE sweep produced,

      M v' + (A + N(t)) v - tB (mtxtb(t)^T v + tB^T w(t)) - J^T p = f,   J v = 0.

Nothing here runs on the GPU; both the CUDA modules' and the oracle's feedback go through
this same code in the tests.
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

__all__ = ['eval_costfunc', 'simulate_closed_loop', 'extract_output', 'save_output_json',
           'load_json_dicts']


def eval_costfunc(V=None, W=None, R=None, cmat=None, ystar=None, bmat=None, tbmat=None,
                  tmesh=None, veldict=None, fbftdict=None, penau=True, store=None):
    """``dy(T)'V dy(T) + int_tmesh dy'W dy + u'u`` with ``dy = ystar - C v`` and
    ``u = mtxtb^T v + tB^T w`` (``optcont_main.py:213-264``)."""
    def _dywdy(t, V=None):
        cvel = store.load(veldict[t])
        delty = ystar(t) - np.asarray(cmat @ cvel)
        wm = W if V is None else V
        return float(np.dot(delty.T, np.asarray(wm @ delty))[0, 0])

    def _uru(t):
        if not penau:
            return 0.0
        if R is None and tbmat is not None:
            cvel = store.load(veldict[t])
            key = t if t in fbftdict else None
            curfb = np.dot(store.load(fbftdict[key]['mtxtb']).T, cvel)
            curft = np.asarray(tbmat.T @ store.load(fbftdict[key]['w']))
            return float(np.dot((curfb+curft).T, curfb+curft)[0, 0])
        raise NotImplementedError()

    cfv = 0.0
    old = _dywdy(tmesh[0]) + _uru(tmesh[0])
    for k, t in enumerate(tmesh[1:]):
        cts = t - tmesh[k]
        new = _dywdy(t) + _uru(t)
        cfv += 0.5*cts*(new + old)
        old = new
    return cfv + _dywdy(tmesh[-1], V=V)


def simulate_closed_loop(mmat=None, amat=None, jmat=None, rhsv=None, tb_mat=None, tmesh=None,
                         get_tdpart=None, gttdprtargs=None, fbftdict=None, store=None,
                         v0=None, key_prfx='vel_'):
    """Implicit-Euler closed-loop simulation on ``tmesh``; saves the velocities into ``store``
    and returns ``{t: key}`` (the ``dictofvels`` the cost functional reads).  ``fbftdict=None``
    simulates the uncontrolled flow."""
    gttdprtargs = {} if gttdprtargs is None else gttdprtargs
    NV, NP = mmat.shape[0], jmat.shape[0]
    v = np.zeros((NV, 1)) if v0 is None else np.asarray(v0, dtype=np.float64).reshape(NV, 1)
    tbd = None if tb_mat is None else np.asarray(tb_mat.todense() if sps.issparse(tb_mat) else tb_mat)
    veldict = {}
    store.save(v, key_prfx + repr(float(tmesh[0])))
    veldict[tmesh[0]] = key_prfx + repr(float(tmesh[0]))
    M, A, J = sps.csr_matrix(mmat), sps.csr_matrix(amat), sps.csr_matrix(jmat)
    for k in range(len(tmesh)-1):
        t1 = tmesh[k+1]
        tau = t1 - tmesh[k]
        nmat, rhstd = get_tdpart(time=t1, **gttdprtargs)
        sys_a = (M + tau*(A + sps.csr_matrix(nmat))).tolil()
        rhs = np.asarray(M @ v) + tau*(np.asarray(rhsv) + np.asarray(rhstd))
        if fbftdict is not None:
            key = t1 if t1 in fbftdict else None
            mtxtb = store.load(fbftdict[key]['mtxtb'])
            w = store.load(fbftdict[key]['w'])
            sys_a = sps.csr_matrix(sys_a) - tau*sps.csr_matrix(tbd @ mtxtb.T)   # NV x NV, rank 2 NU
            rhs = rhs + tau*(tbd @ (tbd.T @ w))
        K = sps.bmat([[sps.csr_matrix(sys_a), -J.T], [J, None]], format='csc')
        sol = spsla.splu(K).solve(np.vstack([rhs, np.zeros((NP, 1))]))
        v = sol[:NV].reshape(NV, 1)
        store.save(v, key_prfx + repr(float(t1)))
        veldict[t1] = key_prfx + repr(float(t1))
    return veldict


def extract_output(dictofpaths=None, tmesh=None, c_mat=None, ystarvec=None, store=None):
    """Output signals ``y(t) = C v(t)`` and the targets along ``tmesh`` as lists per component
    (``dou.extract_output`` as called at ``optcont_main.py:642-645``)."""
    ys, ystars = [], []
    for t in tmesh:
        y = np.asarray(c_mat @ store.load(dictofpaths[t])).ravel()
        ys.append(y.tolist())
        if ystarvec is not None:
            ystars.append(np.asarray(ystarvec(t)).ravel().tolist())
    yscomplist = [list(c) for c in zip(*ys)]
    ystarlist = [list(c) for c in zip(*ystars)] if ystars else None
    return yscomplist, ystarlist


def save_output_json(ycomp, tmesh, ystar=None, fstring=None):
    """The ``__sigout`` file of ``optcont_main.py:160-175``: one JSON object with the keys
    ``ycomp``, ``tmesh``, ``ystar`` (what ``plot_output.plot_optcont_json`` reads)."""
    import json
    if fstring is None:
        fstring = 'nonspecified_output'
    with open(fstring, mode='w') as jsfile:
        jsfile.write(json.dumps(dict(ycomp=ycomp, tmesh=list(tmesh), ystar=ystar)))
    return fstring


def load_json_dicts(StrToJs):
    """``optcont_main.py:178-182``."""
    import json
    with open(StrToJs) as fjs:
        return json.load(fjs)
