"""CUDA-backed ``sadptprj_riclyap_adi.lin_alg_utils`` for the optconpy hot path.

Same module-level functions, keyword names, return shapes and error behaviour as
the reference's call sites expect (cited per function; SURVEY.md 8a rows
a6-a11).  Inputs and outputs are host objects (scipy sparse / numpy), exactly as
the reference driver passes them; every solve and product runs on the GPU
through the C ABI (``include/optconpy_b200.h``).  The sparse LU factorisation
is the separately timed host setup step (``device.LU``).  No CPU fallback.
"""
import numpy as np
import scipy.linalg as spla
import scipy.sparse as sps
import torch

from . import device as dv

__all__ = ['mm_dnssps', 'app_luinv_to_spmat', 'apply_massinv', 'apply_sqrt_fromright',
           'apply_invsqrt_fromright', 'get_Sinv_smw', 'app_smw_inv', 'solve_sadpnt_smw',
           'app_prj_via_sadpnt', 'comp_sqfnrm_factrd_diff', 'comp_sqfnrm_factrd_sum',
           'comp_sqfnrm_factrd_lyap_res', 'SadLU', 'sadlu_async']


def _dense(a):
    if sps.issparse(a):
        return np.asarray(a.todense(), dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    return a[:, None] if a.ndim == 1 else a


class SadLU(object):
    """Callable LU handle (what ``spsla.factorized`` returns in the reference):
    ``alu(rhs)`` solves on the device."""

    def __init__(self, mat, background=False):
        self.shape = mat.shape
        self._lu = None
        # background=True: the factorisation runs in a worker process, uploaded on first use
        self._job = dv.FactorJob([mat]).start_upload() if background else None
        if not background:
            self._lu = dv.LU(mat)

    @property
    def lu(self):
        if self._lu is None:
            self._lu = self._job.result()[0]
        return self._lu

    def __call__(self, rhs):
        rhs = np.asarray(rhs, dtype=np.float64)
        one_d = rhs.ndim == 1
        x = dv.to_host(self.lu.solve(dv.to_dev(rhs)))
        return x[:, 0] if one_d else x


def sadlu_async(amat=None, jmat=None, jmatT=None):
    """Start the factorisation of ``[[A, J^T], [J, 0]]`` in the background; pass the result
    to ``solve_sadpnt_smw(..., sadlu=...)``.  Extension of the reference interface."""
    dv.require_cuda()
    return SadLU(dv.sadpnt_matrix(amat, jmat, jmatT), background=True)


def mm_dnssps(A, v):
    """Sparse/dense agnostic product (``optcont_main.py:232-236``): sparse times a
    dense block goes through the SpMM kernel, small dense products stay on the host."""
    if sps.issparse(A) and not sps.issparse(v):
        vd = _dense(v)
        return dv.to_host(dv.DeviceCSR(A).matmul(dv.to_dev(vd)))
    if sps.issparse(A) or sps.issparse(v):
        return A @ v
    return np.dot(A, v)


def app_luinv_to_spmat(alu_solve, Z):
    """``A^-1 Z`` for a sparse Z -> dense
    (``tests/test_units_compfacres_compress.py:71``)."""
    return alu_solve(_dense(Z))


def apply_massinv(M, rhsa, output=None):
    """``M^-1 rhsa`` (``solve_dae_ric.py:77,81,100,108``; ``optcont_main.py:398``)."""
    res = SadLU(sps.csc_matrix(M))(_dense(rhsa))
    if output == 'sparse':
        return sps.csr_matrix(res)
    return res


def _chol_lower(M):
    return spla.cholesky(_dense(M), lower=True)


def apply_sqrt_fromright(M, rhsa, output=None):
    """``rhsa M^{1/2}`` for the small (2NY x 2NY) weight matrices
    (``solve_dae_ric.py:94``); host-side by design (8 x 8)."""
    res = _dense(rhsa).dot(_chol_lower(M))
    return sps.csr_matrix(res) if output == 'sparse' else res


def apply_invsqrt_fromright(M, rhsa, output=None):
    """``rhsa M^{-1/2}`` (``optcont_main.py:421,424``; ``solve_dae_ric.py:92,97``)."""
    res = spla.solve_triangular(_chol_lower(M), _dense(rhsa).T, lower=True).T
    return sps.csr_matrix(res) if output == 'sparse' else res


def get_Sinv_smw(amat_lu, umat=None, vmat=None):
    """``(I - V A^-1 U)^-1``; the m x m core is inverted on the host
    here only for API parity — the solvers below build it on the device."""
    aiu = amat_lu(_dense(umat))
    vaiu = vmat @ aiu if sps.issparse(vmat) else np.dot(vmat, aiu)
    return np.linalg.inv(np.eye(aiu.shape[1]) - vaiu)


def app_smw_inv(amat, umat=None, vmat=None, rhsa=None, Sinv=None, alu=None):
    """``(A - U V)^-1 rhsa`` by Sherman-Morrison-Woodbury on the device."""
    if alu is None:
        alu = amat if isinstance(amat, SadLU) else SadLU(amat)
    rhs = dv.to_dev(_dense(rhsa))
    n = alu.shape[0]
    if umat is None:
        return dv.to_host(alu.lu.solve(rhs))
    Ufb = dv.to_dev(_dense(umat))
    Vt = dv.DeviceCSR(sps.csr_matrix(vmat))
    return dv.to_host(alu.lu.smw_solve(rhs, n, Ufb=Ufb, Vt=Vt))


def solve_sadpnt_smw(amat=None, jmat=None, rhsv=None, jmatT=None,
                     umat=None, vmat=None, rhsp=None, sadlu=None, return_alu=False):
    """``[[A - U V, J^T], [J, 0]] [v; p] = [rhsv; rhsp]`` -> dense (NV+NP, r)
    (``solve_dae_ric.py:192-194``, ``optcont_main.py:510-514``)."""
    dv.require_cuda()
    NV, nnpp = amat.shape[0], jmat.shape[0]
    rv = _dense(rhsv)
    with dv.phase('sadpnt_factor_wait_upload'):
        alu = sadlu if sadlu is not None else SadLU(dv.sadpnt_matrix(amat, jmat, jmatT))
        alu.lu
    with dv.phase('sadpnt_solve'):
        rhs = rv if rhsp is None else np.vstack([rv, _dense(rhsp)])
        B = dv.to_dev(rhs)
        if umat is not None:
            Ufb = dv.to_dev(_dense(umat))
            Vt = dv.DeviceCSR(sps.csr_matrix(vmat))
            sol = alu.lu.smw_solve(B, NV, Ufb=Ufb, Vt=Vt)
        else:
            sol = alu.lu.solve(B)
        sol = dv.to_host(sol)
    if return_alu:
        return sol, alu
    return sol


def app_prj_via_sadpnt(amat=None, jmat=None, rhsv=None, jmatT=None,
                       umat=None, vmat=None, transposedprj=False):
    """Discrete Leray projector via one saddle-point solve (``optcont_main.py:405-408``):
    ``P^T rhsv = A sadpnt^-1([rhsv;0])[:NV]``, ``P rhsv = sadpnt^-1([A rhsv;0])[:NV]``."""
    dv.require_cuda()
    NV = amat.shape[0]
    alu = SadLU(dv.sadpnt_matrix(amat, jmat, jmatT))
    Ad = dv.DeviceCSR(amat)
    R = dv.to_dev(_dense(rhsv))
    Ufb = Vt = None
    if umat is not None:
        Ufb, Vt = dv.to_dev(_dense(umat)), dv.DeviceCSR(sps.csr_matrix(vmat))
    if transposedprj:
        sol = alu.lu.smw_solve(R, NV, Ufb=Ufb, Vt=Vt, nrows_out=NV)
        return dv.to_host(Ad.matmul(sol))
    return dv.to_host(alu.lu.smw_solve(Ad.matmul(R), NV, Ufb=Ufb, Vt=Vt, nrows_out=NV))


def _gram_fnorm_sq(a, b):
    G = dv.gram(a, b)
    return float((G*G).sum().item())


def comp_sqfnrm_factrd_diff(zone, ztwo):
    """``||Z1 Z1^T - Z2 Z2^T||_F^2`` from three small Gram products (DMMA)."""
    a, b = dv.to_dev(zone), dv.to_dev(ztwo)
    return (_gram_fnorm_sq(a, a) - 2*_gram_fnorm_sq(a, b) + _gram_fnorm_sq(b, b))


def comp_sqfnrm_factrd_sum(zone, ztwo):
    """``||Z1 Z1^T + Z2 Z2^T||_F^2``."""
    a, b = dv.to_dev(zone), dv.to_dev(ztwo)
    return (_gram_fnorm_sq(a, a) + 2*_gram_fnorm_sq(a, b) + _gram_fnorm_sq(b, b))


def comp_sqfnrm_factrd_lyap_res(A, B, C):
    """``||A B^T + B A^T + C C^T||_F^2 = tr(D G D G)`` with the stacked Gram
    ``G = [A B C]^T [A B C]`` (one DMMA Gram product) and ``D`` swapping the A/B blocks."""
    U = torch.cat([dv.to_dev(A), dv.to_dev(B), dv.to_dev(C)], dim=1).contiguous()
    ka, kb = A.shape[1], B.shape[1]
    G = dv.gram(U, U)
    DG = torch.cat([G[ka:ka+kb, :], G[:ka, :], G[ka+kb:, :]], dim=0)
    return float((DG*DG.t()).sum().item())
