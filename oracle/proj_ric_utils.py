"""Oracle restatement of ``sadptprj_riclyap_adi.proj_ric_utils`` (TEST INFRASTRUCTURE).

Absent third-party module (see ``oracle/__init__.py``: parity unpinned).  The
signatures are the reference's call sites; the mathematics is what the
reference's test and driver encode:

* ``solve_proj_lyap_stein`` solves ``P^T (F^T X M + M^T X F + W W^T) P = 0``,
  ``X = Z Z^T`` (``tests/test_units_compfacres_compress.py:62-79``) by the
  Li-White low-rank ADI recurrence on saddle-point systems
  ``[[F^T + mu M^T, J^T], [J, 0]]``.
* ``proj_alg_ric_newtonadi`` is Newton-Kleinman on
  ``F^T X M + M^T X F - M^T X B B^T X M + W W^T = 0``
  (``solve_dae_ric.py:147-159``, ``optcont_main.py:488-492``).
"""
import numpy as np
import scipy.sparse as sps

from . import lin_alg_utils as lau

DEFAULT_SHIFTS = [-30.0, -20.0, -10.0, -5.0, -3.0, -1.0]

# Baseline variant "LU cached across Newton steps" (bench.py cpu_baseline):
# None = factorise the shifted matrices in every call, as a straightforward implementation
# does; a dict = reuse a factorisation whenever the same shifted matrix comes again (the
# Newton steps of one Riccati solve share them).  Numbers are identical either way.
LU_CACHE = None


def _shifted_lu(At, Mt, jmat, mu):
    if LU_CACHE is None:
        return lau.SadLU(lau.sadpnt_matrix(At + mu*Mt, jmat))
    key = (float(mu), At.shape, At.nnz, float(At.data.sum()), float(np.abs(At.data).sum()),
           float(Mt.data.sum()), id(jmat))
    if key not in LU_CACHE:
        LU_CACHE[key] = lau.SadLU(lau.sadpnt_matrix(At + mu*Mt, jmat))
    return LU_CACHE[key]


def _dense(a):
    if sps.issparse(a):
        return np.asarray(a.todense(), dtype=np.float64)
    return np.asarray(a, dtype=np.float64)


def solve_proj_lyap_stein(amat=None, jmat=None, wmat=None, mmat=None,
                          umat=None, vmat=None, transposed=False,
                          adi_dict=dict(adi_max_steps=150,
                                        adi_newZ_reltol=1e-8),
                          nwtn_adi_dict=None, **kw):
    """Low-rank ADI for the projected Lyapunov equation

        [F-UV]^T X M + M^T X [F-UV] + W W^T = 0   on  {J X M = 0, M^T X J^T = 0}

    Call site: ``tests/test_units_compfacres_compress.py:62-64``.  ``transposed``
    means ``amat, mmat`` are handed in already transposed
    (``solve_dae_ric.py:152-153``).  Returns ``dict(zfac=Z,
    adi_rel_newZ_norms=[...])``.
    """
    if nwtn_adi_dict is not None:
        adi_dict = nwtn_adi_dict
    if transposed:
        At, Mt = sps.csr_matrix(amat), sps.csr_matrix(mmat)
    else:
        At, Mt = sps.csr_matrix(amat.T), sps.csr_matrix(mmat.T)
    ms = list(adi_dict.get('ms', DEFAULT_SHIFTS))
    NV, NP = At.shape[0], jmat.shape[0]
    W = _dense(wmat)

    lus, aius, sinvs = [], [], []
    if umat is not None and vmat is not None:
        # (F - U V)^T = F^T - V^T U^T : SMW with "U" = V^T (dense), "V" = U^T
        ut = np.vstack([_dense(vmat).T, np.zeros((NP, vmat.shape[0]))])
        vt = sps.hstack([sps.csr_matrix(umat).T,
                         sps.csr_matrix((umat.shape[1], NP))], format='csr')
    else:
        ut, vt = None, None
    for mu in ms:
        alu = _shifted_lu(At, Mt, jmat, mu)
        lus.append(alu)
        if ut is not None:
            sinvs.append(lau.get_Sinv_smw(alu, umat=ut, vmat=vt))

    def shifted_solve(i, R):
        rhs = np.vstack([R, np.zeros((NP, R.shape[1]))])
        if ut is None:
            return lus[i](rhs)[:NV, :]
        return lau.app_smw_inv(lus[i], umat=ut, vmat=vt, rhsa=rhs,
                               Sinv=sinvs[i])[:NV, :]

    maxsteps = int(adi_dict['adi_max_steps'])
    reltol = adi_dict['adi_newZ_reltol']
    V = np.sqrt(-2.0*ms[0])*shifted_solve(0, W)
    blocks = [V]
    z_nsq = np.linalg.norm(V)**2
    rel_norms = [1.0]
    step = 1
    while step < maxsteps and rel_norms[-1] > reltol:
        i, ip = step % len(ms), (step-1) % len(ms)
        X = shifted_solve(i, np.asarray(Mt @ V))
        V = np.sqrt(ms[i]/ms[ip])*(V - (ms[i] + ms[ip])*X)
        blocks.append(V)
        v_nsq = np.linalg.norm(V)**2
        z_nsq += v_nsq
        rel_norms.append(np.sqrt(v_nsq/z_nsq))
        step += 1
    return dict(zfac=np.hstack(blocks), adi_rel_newZ_norms=rel_norms)


def get_mTzzTtb(MT, Z, tB, output=None):
    """``M^T (Z (Z^T tB))`` -> dense (NV, m)
    (``solve_dae_ric.py:101,183,189``; ``optcont_main.py:505-506``)."""
    ztb = Z.T @ tB if sps.issparse(tB) else np.dot(Z.T, tB)
    return np.asarray(MT @ np.dot(Z, np.asarray(ztb)))


def _probe_vec(n, nwtn_adi_dict):
    rng = np.random.default_rng(nwtn_adi_dict.get('probe_seed', 0))
    vec = rng.standard_normal((n, 1))
    return vec/np.linalg.norm(vec)


def proj_alg_ric_newtonadi(mmat=None, amat=None, jmat=None,
                           bmat=None, wmat=None, z0=None, mtxoldb=None,
                           transposed=False,
                           nwtn_adi_dict=dict(adi_max_steps=150,
                                              adi_newZ_reltol=1e-5,
                                              nwtn_max_steps=14,
                                              nwtn_upd_reltol=1e-8),
                           **kw):
    """Newton-Kleinman / LR-ADI for the projected algebraic Riccati equation

        F^T X M + M^T X F - M^T X B B^T X M + W W^T = 0,   X = Z Z^T.

    Call sites: ``optcont_main.py:488-492`` (non-transposed),
    ``solve_dae_ric.py:152-159`` (``transposed=True``: ``mmat=M^T``,
    ``amat=F^T``).  ``mtxoldb`` (NV, m): feedback of a previous outer Newton
    step, folded into the drift as ``F + B mtxoldb^T`` (only reachable with
    ``outernwtnstps > 1``; semantics unpinned).  Stops when the update norm
    falls below ``nwtn_upd_reltol`` (relative) or ``nwtn_upd_abstol``.
    Returns ``dict(zfac=Z, nwtn_upd_fnorms=[...], adi_steps=[...])``.
    """
    MT = sps.csr_matrix(mmat) if transposed else sps.csr_matrix(mmat.T)
    B = bmat
    W = _dense(wmat)
    znc = z0
    fnorms, adi_steps = [], []
    stp = 0
    maxstp = int(nwtn_adi_dict['nwtn_max_steps'])
    reltol = nwtn_adi_dict.get('nwtn_upd_reltol', 0.0)
    abstol = nwtn_adi_dict.get('nwtn_upd_abstol', 0.0)
    full = nwtn_adi_dict.get('full_upd_norm_check', False)
    while stp < maxstp:
        if znc is None:
            rhsadi, kfb = W, None
        else:
            kfb = get_mTzzTtb(MT, znc, B)                 # M^T Z Z^T B
            rhsadi = np.hstack([kfb, W])
        if mtxoldb is not None:
            kfb = -_dense(mtxoldb) if kfb is None else kfb - _dense(mtxoldb)
        res = solve_proj_lyap_stein(amat=amat, jmat=jmat, wmat=rhsadi, mmat=mmat,
                                    umat=B if kfb is not None else None,
                                    vmat=kfb.T if kfb is not None else None,
                                    transposed=transposed,
                                    adi_dict=nwtn_adi_dict)
        znn = res['zfac']
        adi_steps.append(len(res['adi_rel_newZ_norms']))
        if full:
            if znc is None:
                upd = np.linalg.norm(np.dot(znn.T, znn))
            else:
                upd = np.sqrt(np.abs(lau.comp_sqfnrm_factrd_diff(znn, znc)))
            ref = np.linalg.norm(np.dot(znn.T, znn))
        else:
            vec = _probe_vec(znn.shape[0], nwtn_adi_dict)
            nv = np.dot(znn, np.dot(znn.T, vec))
            cv = 0*nv if znc is None else np.dot(znc, np.dot(znc.T, vec))
            upd, ref = np.linalg.norm(nv - cv), np.linalg.norm(nv)
        fnorms.append(upd)
        znc = znn
        stp += 1
        if upd < abstol or upd < reltol*ref:
            break
    return dict(zfac=znc, nwtn_upd_fnorms=fnorms, adi_steps=adi_steps)


def compress_Zsvd(Z, k=None, thresh=None, shplot=False):
    """Column compression ``Zc = Z V_k`` by the SVD ``Z = U S V^T``: keep the
    singular values ``> thresh``, at most ``k`` of them
    (``solve_dae_ric.py:162``, ``optcont_main.py:498``,
    ``tests/test_units_compfacres_compress.py:92``).  ``Zc Zc^T = U_k S_k^2 U_k^T``."""
    U, s, Vt = np.linalg.svd(Z, full_matrices=False)
    keep = s.size
    if thresh is not None:
        keep = int(np.sum(s > thresh))
    if k is not None:
        keep = min(keep, int(k))
    return np.dot(Z, Vt[:keep, :].T)


def comp_proj_lyap_res_norm(Z, amat=None, mmat=None, wmat=None, jmat=None,
                            umat=None, vmat=None):
    """SQUARED Frobenius norm of the projected Lyapunov residual
    ``P^T (F^T Z Z^T M + M^T Z Z^T F + W W^T) P`` in factored form
    (positional call ``(Z, F, M, W, J)``,
    ``tests/test_units_compfacres_compress.py:82,104``)."""
    Ft, Mt = sps.csr_matrix(amat.T), sps.csr_matrix(mmat.T)
    ftz = np.asarray(Ft @ Z)
    if umat is not None and vmat is not None:
        ftz = ftz - np.dot(_dense(vmat).T, np.asarray(sps.csr_matrix(umat).T @ Z))
    mtz = np.asarray(Mt @ Z)
    stacked = np.hstack([ftz, mtz, _dense(wmat)])
    prj = lau.app_prj_via_sadpnt(amat=mmat, jmat=jmat, rhsv=stacked,
                                 transposedprj=True)
    ka, kb = ftz.shape[1], mtz.shape[1]
    return lau.comp_sqfnrm_factrd_lyap_res(prj[:, :ka], prj[:, ka:ka+kb],
                                           prj[:, ka+kb:])
