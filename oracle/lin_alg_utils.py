"""Oracle restatement of ``sadptprj_riclyap_adi.lin_alg_utils`` (TEST INFRASTRUCTURE).

Absent third-party module (see ``oracle/__init__.py``); signatures follow the
reference's call sites, cited per function.  scipy/SuperLU on the host.
``COLUMNWISE=True`` mirrors the upstream structure (one right-hand-side column
per SuperLU call in a Python loop — recollection, unverified); ``False`` hands
the whole block to ``SuperLU.solve`` (the best scipy can do).
"""
import numpy as np
import scipy.linalg as spla
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

COLUMNWISE = False


def _dense(a):
    if sps.issparse(a):
        return np.asarray(a.todense(), dtype=np.float64)
    return np.asarray(a, dtype=np.float64)


class SadLU(object):
    """LU handle of a sparse matrix; callable like ``spsla.factorized`` results."""

    def __init__(self, mat):
        self.lu = spsla.splu(sps.csc_matrix(mat))
        self.shape = mat.shape

    def __call__(self, rhs):
        return self.solve(rhs)

    def solve(self, rhs):
        rhs = np.asarray(rhs, dtype=np.float64)
        if rhs.ndim == 1 or not COLUMNWISE:
            return self.lu.solve(rhs)
        out = np.empty_like(rhs)
        for c in range(rhs.shape[1]):
            out[:, c] = self.lu.solve(np.ascontiguousarray(rhs[:, c]))
        return out


def mm_dnssps(A, v):
    """sparse-or-dense agnostic product (``optcont_main.py:232-236``)."""
    if sps.issparse(A) or sps.issparse(v):
        out = A @ v
        return out
    return np.dot(A, v)


def app_luinv_to_spmat(alu_solve, Z):
    """LU-inverse applied to a sparse matrix -> dense
    (``tests/test_units_compfacres_compress.py:71``)."""
    Zd = _dense(Z)
    out = np.zeros(Zd.shape)
    for c in range(Zd.shape[1]):
        out[:, c] = alu_solve(Zd[:, c])
    return out


def apply_massinv(M, rhsa, output=None):
    """``M^-1 rhsa`` for sparse or dense rhsa (``solve_dae_ric.py:77,81,100,108``;
    ``optcont_main.py:398`` with ``output='sparse'``)."""
    mlu = SadLU(M)
    res = mlu.solve(_dense(rhsa))
    if output == 'sparse':
        return sps.csr_matrix(res)
    return res


def _chol_lower(M):
    return spla.cholesky(_dense(M), lower=True)


def apply_sqrt_fromright(M, rhsa, output=None):
    """``rhsa M^{1/2}`` with the Cholesky factor as square root
    (``solve_dae_ric.py:94``); ``(rhsa L)(rhsa L)^T = rhsa M rhsa^T``."""
    L = _chol_lower(M)
    res = _dense(rhsa).dot(L)
    return sps.csr_matrix(res) if output == 'sparse' else res


def apply_invsqrt_fromright(M, rhsa, output=None):
    """``rhsa M^{-1/2}`` (``optcont_main.py:421,424``; ``solve_dae_ric.py:92,97``);
    ``(rhsa L^-T)(rhsa L^-T)^T = rhsa M^-1 rhsa^T``."""
    L = _chol_lower(M)
    res = spla.solve_triangular(L, _dense(rhsa).T, lower=True).T
    return sps.csr_matrix(res) if output == 'sparse' else res


def get_Sinv_smw(amat_lu, umat=None, vmat=None):
    """``(I - V A^-1 U)^-1`` for Sherman-Morrison-Woodbury."""
    aiu = amat_lu(_dense(umat)) if not COLUMNWISE else \
        app_luinv_to_spmat(amat_lu, umat)
    vaiu = vmat @ aiu if sps.issparse(vmat) else np.dot(vmat, aiu)
    return np.linalg.inv(np.eye(aiu.shape[1]) - vaiu)


def app_smw_inv(amat, umat=None, vmat=None, rhsa=None, Sinv=None, alu=None):
    """``(A - U V)^-1 rhsa`` by Sherman-Morrison-Woodbury:
    ``x = y + A^-1 U (I - V A^-1 U)^-1 V y``, ``y = A^-1 rhsa``  (SURVEY a11)."""
    if alu is None:
        alu = amat if callable(amat) else SadLU(amat)
    rhs = _dense(rhsa)
    y = alu(rhs)
    if umat is None:
        return y
    ud = _dense(umat)
    if Sinv is None:
        Sinv = get_Sinv_smw(alu, umat=ud, vmat=vmat)
    vy = vmat @ y if sps.issparse(vmat) else np.dot(vmat, y)
    # upstream applies A^-1 a second time to the corrected right-hand side
    crhs = rhs + np.dot(ud, np.dot(Sinv, vy))
    return alu(crhs)


def sadpnt_matrix(amat, jmat, jmatT=None):
    """``[[A, J^T], [J, 0]]`` in CSC."""
    nnpp = jmat.shape[0]
    if jmatT is None:
        jmatT = jmat.T
    return sps.bmat([[sps.csr_matrix(amat), sps.csr_matrix(jmatT)],
                     [sps.csr_matrix(jmat), sps.csr_matrix((nnpp, nnpp))]],
                    format='csc')


def solve_sadpnt_smw(amat=None, jmat=None, rhsv=None, jmatT=None,
                     umat=None, vmat=None, rhsp=None, sadlu=None,
                     return_alu=False):
    """Solve ``[[A - U V, J^T], [J, 0]] [v; p] = [rhsv; rhsp]`` by LU + SMW and
    return the full ``(NV+NP, r)`` array; callers slice ``[:NV]``
    (``solve_dae_ric.py:192-194``, ``optcont_main.py:510-514``)."""
    nnpp = jmat.shape[0]
    rv = _dense(rhsv)
    if rhsp is None:
        rhsp = np.zeros((nnpp, rv.shape[1]))
    alu = sadlu if sadlu is not None else SadLU(sadpnt_matrix(amat, jmat, jmatT))
    rhs = np.vstack([rv, _dense(rhsp)])
    if umat is not None:
        umate = np.vstack([_dense(umat), np.zeros((nnpp, umat.shape[1]))])
        vmate = sps.hstack([sps.csr_matrix(vmat),
                            sps.csr_matrix((vmat.shape[0], nnpp))], format='csr')
        sol = app_smw_inv(alu, umat=umate, vmat=vmate, rhsa=rhs)
    else:
        sol = alu(rhs)
    if return_alu:
        return sol, alu
    return sol


def app_prj_via_sadpnt(amat=None, jmat=None, rhsv=None, jmatT=None,
                       umat=None, vmat=None, transposedprj=False):
    """Discrete Leray projection through one saddle-point solve
    (``optcont_main.py:405-408``).  With ``P = I - A^-1 J^T S^-1 J``,
    ``S = J A^-1 J^T``: ``P rhsv = sadpnt^-1([A rhsv; 0])[:NV]`` and
    ``P^T rhsv = A sadpnt^-1([rhsv; 0])[:NV]`` (A symmetric, e.g. the mass matrix)."""
    NV = amat.shape[0]
    if transposedprj:
        sol = solve_sadpnt_smw(amat=amat, jmat=jmat, rhsv=rhsv, jmatT=jmatT,
                               umat=umat, vmat=vmat)[:NV, :]
        return np.asarray(amat @ sol)
    arhs = amat @ rhsv
    return solve_sadpnt_smw(amat=amat, jmat=jmat, rhsv=arhs, jmatT=jmatT,
                            umat=umat, vmat=vmat)[:NV, :]


def comp_sqfnrm_factrd_diff(zone, ztwo):
    """``||Z1 Z1^T - Z2 Z2^T||_F^2`` from small Gram matrices."""
    return (np.linalg.norm(np.dot(zone.T, zone))**2
            - 2*np.linalg.norm(np.dot(zone.T, ztwo))**2
            + np.linalg.norm(np.dot(ztwo.T, ztwo))**2)


def comp_sqfnrm_factrd_sum(zone, ztwo):
    """``||Z1 Z1^T + Z2 Z2^T||_F^2``."""
    return (np.linalg.norm(np.dot(zone.T, zone))**2
            + 2*np.linalg.norm(np.dot(zone.T, ztwo))**2
            + np.linalg.norm(np.dot(ztwo.T, ztwo))**2)


def comp_sqfnrm_factrd_lyap_res(A, B, C):
    """``||A B^T + B A^T + C C^T||_F^2`` without forming the big matrices:
    with the stacked Gram ``G = [A B C]^T [A B C]`` it is
    ``tr(D G D G)``, ``D = [[0,I,0],[I,0,0],[0,0,I]]``."""
    U = np.hstack([A, B, C])
    ka, kb = A.shape[1], B.shape[1]
    G = np.dot(U.T, U)
    DG = np.vstack([G[ka:ka+kb, :], G[:ka, :], G[ka+kb:, :]])
    return float(np.sum(DG * DG.T))
