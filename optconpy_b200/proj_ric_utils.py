"""CUDA-backed ``sadptprj_riclyap_adi.proj_ric_utils`` for the optconpy hot path.

Module-level functions with the names, keyword arguments and return shapes of
the reference's call sites:

* ``solve_proj_lyap_stein``   ``tests/test_units_compfacres_compress.py:62-64``
* ``proj_alg_ric_newtonadi``  ``optcont_main.py:488-492``, ``solve_dae_ric.py:152-159``
* ``compress_Zsvd``           ``solve_dae_ric.py:162``, ``optcont_main.py:498``
* ``get_mTzzTtb``             ``solve_dae_ric.py:101,183,189``, ``optcont_main.py:505-506``
* ``comp_proj_lyap_res_norm`` ``tests/test_units_compfacres_compress.py:82,104``

Host objects in, host objects out; the factor ``Z`` stays resident on the device
between the Newton steps (``*_dev`` helpers work on torch tensors and are what
the device-resident DRE loop uses).  The per-shift sparse LU factorisations are
the separately timed host setup (``device.LU``); all solves, products, norms,
the compression and the eigen-decomposition run in CUDA kernels.  No CPU fallback.
"""
import numpy as np
import scipy.sparse as sps
import torch

from . import device as dv
from . import parallel as par

__all__ = ['solve_proj_lyap_stein', 'proj_alg_ric_newtonadi', 'compress_Zsvd',
           'get_mTzzTtb', 'comp_proj_lyap_res_norm', 'factors_async', 'lookahead_thread_init',
           'tail_thread_init']

lookahead_thread_init = dv.lookahead_thread_init
# a second host thread may call get_mTzzTtb / lau.solve_sadpnt_smw while this one sits in
# proj_alg_ric_newtonadi: workspaces, staging buffers and the C library's scratch are per thread
tail_thread_init = dv.tail_thread_init

DEFAULT_SHIFTS = [-30.0, -20.0, -10.0, -5.0, -3.0, -1.0]

# the last factor handed back to the caller (host array, device tensor): compress_Zsvd of that
# very array object skips the re-upload.  Callers must not modify the array in between - the
# reference never does (solve_dae_ric.py:159-162 passes it straight on).
_LAST = dict()


def _dense(a):
    if sps.issparse(a):
        return np.asarray(a.todense(), dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    return a[:, None] if a.ndim == 1 else a


class ShiftedFactors(object):
    """Per-shift LU handles of ``[[At + mu Mt, J^T], [J, 0]]`` - the setup step that
    ``north_star`` times separately.  Reused across the Newton steps of one
    Riccati solve (the low-rank closed-loop part enters through SMW only).

    Construction only SUBMITS the host factorisations (worker processes); the handles are
    uploaded on first use of ``.lus``.  Creating the object early therefore overlaps the
    setup of the next time step with the device work of the current one
    (``factors_async`` / ``dre_stepper`` look-ahead)."""

    def __init__(self, At, Mt, jmat, ms, Mt_dev=None, k_hint=None, wide=False, shared=None):
        self.ms = [float(m) for m in ms]
        self.NV, self.NP = At.shape[0], jmat.shape[0]
        # k_hint: expected number of right-hand-side columns of the ADI blocks (picks the
        # cluster size the factor images are packed for); wide: also pack the panel program of
        # the all-columns-at-once executor; shared: a parallel.ShardComm - the shifts are dealt
        # to its ranks for the host factorisation (shift-sharded setup)
        mats = _shifted_saddle_matrices(At, Mt, jmat, self.ms)
        self._lus = None
        if shared is not None:
            self._job = None
            self._lus = par.shared_factors(shared, mats, wide=wide, k_hint=k_hint)
        else:
            self._job = dv.FactorJob(mats, k_hint=k_hint, wide=wide).start_upload()
        self._Mt, self._Mt_dev = Mt, Mt_dev

    @property
    def lus(self):
        return self._lus if self._lus is not None else self._job.result()

    @property
    def Mt_dev(self):
        if self._Mt_dev is None:
            self._Mt_dev = dv.DeviceCSR(self._Mt)
        return self._Mt_dev


_ASM = dict()


def _same_pattern(a, indptr, indices):
    return (a.nnz == len(indices) and np.array_equal(a.indptr, indptr)
            and np.array_equal(a.indices, indices))


def _shifted_saddle_matrices(At, Mt, jmat, ms):
    """``[[At + mu Mt, J^T], [J, 0]]`` for every shift.  The block structure is assembled
    once per (pattern of At, pattern of Mt, J) on the union pattern and cached; per shift and
    per time step only values are scattered (``sps.bmat`` + a sparse add per matrix otherwise)."""
    At, Mt = sps.csr_matrix(At), sps.csr_matrix(Mt)
    for m_ in (At, Mt):
        m_.sum_duplicates()
        if not m_.has_sorted_indices:
            m_.sort_indices()
    c = _ASM.get('asm')
    if not (c is not None and c['jmat'] is jmat and _same_pattern(At, c['a_ip'], c['a_ix'])
            and _same_pattern(Mt, c['m_ip'], c['m_ix'])):
        ncol = At.shape[1]
        P = sps.csr_matrix((np.ones(At.nnz), At.indices, At.indptr), shape=At.shape) \
            + sps.csr_matrix((np.ones(Mt.nnz), Mt.indices, Mt.indptr), shape=Mt.shape)
        P.sort_indices()
        rows = lambda m_: np.repeat(np.arange(m_.shape[0], dtype=np.int64), np.diff(m_.indptr))
        keyP = rows(P)*ncol + P.indices
        c = dict(jmat=jmat, a_ip=At.indptr.copy(), a_ix=At.indices.copy(),
                 m_ip=Mt.indptr.copy(), m_ix=Mt.indices.copy(), nnz=P.nnz,
                 posA=np.searchsorted(keyP, rows(At)*ncol + At.indices),
                 posM=np.searchsorted(keyP, rows(Mt)*ncol + Mt.indices),
                 asm=dv.SaddleAssembler(P, jmat))
        _ASM['asm'] = c
    out = []
    base = np.zeros(c['nnz'])
    base[c['posA']] = At.data
    for mu in ms:
        d = base.copy()
        d[c['posM']] += mu*Mt.data
        out.append(c['asm'].assemble(d))
    return out


def factors_async(mmat=None, amat=None, jmat=None, nwtn_adi_dict=None, transposed=False,
                  k_hint=None, **kw):
    """Start the per-shift factorisations of a later ``proj_alg_ric_newtonadi`` /
    ``solve_proj_lyap_stein`` call (same ``mmat, amat, jmat, transposed``) in the background;
    hand the result to that call as ``_factors=``.  Extension of the reference interface:
    callers that do not use it get the same numbers, only later."""
    dv.require_cuda()
    At, Mt = _transposed_pair(amat, mmat, transposed)
    ms = (nwtn_adi_dict or {}).get('ms', DEFAULT_SHIFTS)
    return ShiftedFactors(At, Mt, jmat, ms, k_hint=k_hint)


def _stein_dev(fac, W, adi_dict, Ufb=None, Vt=None):
    """LR-ADI on the device: W (NV x k) device block -> (Z device, rel norms)."""
    return dv.adi_run(fac.lus, fac.ms, fac.NV, fac.NP, fac.Mt_dev, W,
                      int(adi_dict['adi_max_steps']), float(adi_dict['adi_newZ_reltol']),
                      Ufb=Ufb, Vt=Vt)


def _transposed_pair(amat, mmat, transposed):
    if transposed:
        return sps.csr_matrix(amat), sps.csr_matrix(mmat)
    return sps.csr_matrix(amat.T), sps.csr_matrix(mmat.T)


def solve_proj_lyap_stein(amat=None, jmat=None, wmat=None, mmat=None,
                          umat=None, vmat=None, transposed=False,
                          adi_dict=dict(adi_max_steps=150, adi_newZ_reltol=1e-8),
                          nwtn_adi_dict=None, **kw):
    """Low-rank ADI for ``[F-UV]^T X M + M^T X [F-UV] + W W^T = 0`` on the
    divergence-free subspace, ``X = Z Z^T``.  Returns
    ``dict(zfac=Z, adi_rel_newZ_norms=[...])`` with ``Z`` a numpy array."""
    dv.require_cuda()
    if nwtn_adi_dict is not None:
        adi_dict = nwtn_adi_dict
    At, Mt = _transposed_pair(amat, mmat, transposed)
    fac = kw.get('_factors')
    if fac is None:
        fac = ShiftedFactors(At, Mt, jmat, adi_dict.get('ms', DEFAULT_SHIFTS),
                             k_hint=_dense(wmat).shape[1])
    W = dv.to_dev(_dense(wmat))
    Ufb = Vt = None
    if umat is not None and vmat is not None:
        # (F - U V)^T = F^T - V^T U^T: dense SMW factor V^T (NV x m), sparse factor U^T
        Ufb = dv.to_dev(_dense(vmat).T)
        Vt = dv.DeviceCSR(sps.csr_matrix(umat).T)
    cm = par.comm()
    if cm is not None:
        # column-sharded over the ranks (parallel.enable): same result on every rank
        Zl, widths, rel = par.sharded_stein(cm, fac, W, adi_dict, Ufb=Ufb, Vt=Vt)
        zf = par.ShardedFactor(cm, Zl, widths)
        return dict(zfac=zf if kw.get('_lazy_zfac') else np.asarray(zf), adi_rel_newZ_norms=rel)
    Z, rel = _stein_dev(fac, W, adi_dict, Ufb=Ufb, Vt=Vt)
    return dict(zfac=dv.to_host(Z), adi_rel_newZ_norms=rel)


_CSR_CACHE = dict()


def _cached_csr(mat):
    """DeviceCSR of a host matrix the caller passes again and again (``MT`` in every
    feedback product): keyed by object identity, validated by the value checksum."""
    c = _CSR_CACHE.get('m')
    sig = (id(mat), mat.shape, mat.nnz, float(mat.data.sum()) if mat.nnz else 0.0)
    if c is None or c[0] != sig:
        c = (sig, dv.DeviceCSR(mat))
        _CSR_CACHE['m'] = c
    return c[1]


def get_mTzzTtb(MT, Z, tB, output=None):
    """``M^T (Z (Z^T tB))`` -> dense ndarray (NV, m)."""
    dv.require_cuda()
    with dv.phase('feedback'):
        Mt = MT if isinstance(MT, dv.DeviceCSR) else _cached_csr(MT)
        out = dv.feedback(Mt, dv.to_dev(Z), dv.to_dev(_dense(tB)))
        return dv.to_host(out)


def _probe_vec(n, nwtn_adi_dict):
    rng = np.random.default_rng(nwtn_adi_dict.get('probe_seed', 0))
    vec = rng.standard_normal((n, 1))
    return vec/np.linalg.norm(vec)


def _fro(t):
    return float(torch.sqrt((t*t).sum()).item())


def newtonadi_dev(fac, Bd, Vt_b, W, z0, nwtn_adi_dict, mtxoldb=None):
    """Newton-Kleinman with everything resident on the device.
    fac: ShiftedFactors, Bd: dense device B (NV x m), Vt_b: DeviceCSR of B^T,
    W: device (NV x p), z0: device factor or None.  Returns (Z device, info)."""
    znc = z0
    fnorms, adi_steps, rels = [], [], []
    maxstp = int(nwtn_adi_dict['nwtn_max_steps'])
    reltol = nwtn_adi_dict.get('nwtn_upd_reltol', 0.0)
    abstol = nwtn_adi_dict.get('nwtn_upd_abstol', 0.0)
    full = nwtn_adi_dict.get('full_upd_norm_check', False)
    vec = None
    stp = 0
    while stp < maxstp:
        if znc is None:
            rhsadi, kfb = W, None
        else:
            kfb = dv.feedback(fac.Mt_dev, znc, Bd)               # M^T Z Z^T B
            rhsadi = torch.cat([kfb, W], dim=1).contiguous()
        if mtxoldb is not None:
            kfb = -mtxoldb if kfb is None else kfb - mtxoldb
        znn, rel = _stein_dev(fac, rhsadi, nwtn_adi_dict,
                              Ufb=kfb, Vt=Vt_b if kfb is not None else None)
        adi_steps.append(len(rel))
        rels.append(rel)
        if full:
            gnn = dv.gram(znn, znn)
            ref = _fro(gnn)
            if znc is None:
                upd = ref
            else:
                gnc, gcc = dv.gram(znn, znc), dv.gram(znc, znc)
                upd = np.sqrt(abs(ref**2 - 2*_fro(gnc)**2 + _fro(gcc)**2))
        else:
            if vec is None:
                vec = dv.to_dev(_probe_vec(znn.shape[0], nwtn_adi_dict))
            nv = dv.tall_gemm(znn, dv.gram(znn, vec))
            ref = _fro(nv)
            if znc is None:
                upd = ref
            else:
                upd = _fro(nv - dv.tall_gemm(znc, dv.gram(znc, vec)))
        fnorms.append(upd)
        znc = znn
        stp += 1
        if upd < abstol or upd < reltol*ref:
            break
    return znc, dict(nwtn_upd_fnorms=fnorms, adi_steps=adi_steps, adi_rel_norms=rels)


class DeviceFactor(object):
    """A low-rank factor that still lives in HBM, handed out instead of an ndarray where the
    caller asks for it (``proj_alg_ric_newtonadi(..., _lazy_zfac=True)``).  It has ``shape``,
    ``dtype`` and ``ndim``, turns into an ndarray on first use (``np.asarray``, ``np.save``,
    indexing), and ``compress_Zsvd`` takes it as is.  The DRE driver (``solve_dae_ric.py:152-163``)
    only compresses the uncompressed factor and never reads it, so the device->host copy of
    45 MB per time step is not made unless somebody looks."""

    def __init__(self, dev):
        self._dev = dev
        self._host = None

    shape = property(lambda self: tuple(self._dev.shape))
    dtype = property(lambda self: np.dtype(np.float64))
    ndim = property(lambda self: self._dev.dim())

    def __len__(self):
        return self._dev.shape[0]

    def __array__(self, dtype=None, copy=None):
        if self._host is None:
            with dv.phase('ric_d2h_factor'):
                self._host = dv.to_host(self._dev)
        a = self._host
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __getitem__(self, idx):
        return self.__array__()[idx]


def proj_alg_ric_newtonadi(mmat=None, amat=None, jmat=None,
                           bmat=None, wmat=None, z0=None, mtxoldb=None,
                           transposed=False,
                           nwtn_adi_dict=dict(adi_max_steps=150, adi_newZ_reltol=1e-5,
                                              nwtn_max_steps=14, nwtn_upd_reltol=1e-8),
                           **kw):
    """Newton-ADI for ``F^T X M + M^T X F - M^T X B B^T X M + W W^T = 0``,
    ``X = Z Z^T`` (projected).  ``transposed=True``: ``mmat=M^T``, ``amat=F^T``.
    Returns ``dict(zfac=Z, nwtn_upd_fnorms=[...], adi_steps=[...])``."""
    dv.require_cuda()
    At, Mt = _transposed_pair(amat, mmat, transposed)
    fac = kw.get('_factors')
    if fac is None:
        khint = _dense(wmat).shape[1] + (0 if z0 is None else sps.csr_matrix(bmat).shape[1])
        fac = ShiftedFactors(At, Mt, jmat, nwtn_adi_dict.get('ms', DEFAULT_SHIFTS), k_hint=khint)
    with dv.phase('ric_upload_inputs'):
        Bd = dv.to_dev(_dense(bmat))
        Vt_b = dv.DeviceCSR(sps.csr_matrix(bmat).T)
        W = dv.to_dev(_dense(wmat))
        z0d = None if z0 is None else dv.to_dev(z0)
        old = None if mtxoldb is None else dv.to_dev(_dense(mtxoldb))
    with dv.phase('ric_factor_wait_upload'):
        fac.lus
        fac.Mt_dev
    cm = par.comm()
    if cm is not None:
        # column-sharded Newton-ADI (parallel.enable): the factor stays distributed
        probe = None if nwtn_adi_dict.get('full_upd_norm_check', False) else \
            dv.to_dev(_probe_vec(W.shape[0], nwtn_adi_dict))
        with dv.phase('ric_newton_adi_device'):
            Zl, widths, info = par.sharded_newtonadi(cm, fac, Bd, Vt_b, W, z0d, nwtn_adi_dict, mtxoldb=old,
                                                     probe=probe)
            torch.cuda.current_stream().synchronize()
        zf = par.ShardedFactor(cm, Zl, widths)
        if kw.get('_lazy_zfac') or kw.get('_return_device'):
            return dict(zfac=zf, **info)
        return dict(zfac=np.asarray(zf), **info)
    with dv.phase('ric_newton_adi_device'):
        Z, info = newtonadi_dev(fac, Bd, Vt_b, W, z0d, nwtn_adi_dict, mtxoldb=old)
        torch.cuda.current_stream().synchronize()
    if kw.get('_return_device'):
        return dict(zfac=Z, **info)
    if kw.get('_lazy_zfac'):
        return dict(zfac=DeviceFactor(Z), **info)
    with dv.phase('ric_d2h_factor'):
        zh = dv.to_host(Z)
    _LAST['host'], _LAST['dev'] = zh, Z
    return dict(zfac=zh, **info)


def compress_Zsvd(Z, k=None, thresh=None, shplot=False):
    """Column compression ``Zc = Z V_k`` (singular values ``> thresh``, at most ``k``):
    rank-revealing Cholesky of the Gram matrix, Jacobi eigen-solver and the two
    tall products, all on the device."""
    dv.require_cuda()
    if isinstance(Z, par.ShardedFactor):
        with dv.phase('compress'):
            Zc, info = par.sharded_compress(Z.comm, Z.local, Z.widths, thresh=thresh, k=k)
            return dv.to_host(Zc)
    with dv.phase('compress'):
        if isinstance(Z, torch.Tensor):
            Zd = Z
        elif isinstance(Z, DeviceFactor):
            Zd = Z._dev
        elif Z is _LAST.get('host'):
            Zd = _LAST['dev']          # the factor this module just returned: still in HBM
        else:
            Zd = dv.to_dev(Z)
        _LAST.clear()
        Zc, info = dv.compress(Zd, thresh=thresh, k=k)
        if isinstance(Z, torch.Tensor):
            return Zc
        Zc = Zc.contiguous()
        zh = dv.to_host(Zc)
        dv.remember_device_copy(zh, Zc)      # the driver hands this array straight back in
        return zh


def comp_proj_lyap_res_norm(Z, amat=None, mmat=None, wmat=None, jmat=None,
                            umat=None, vmat=None):
    """SQUARED Frobenius norm of ``P^T (F^T Z Z^T M + M^T Z Z^T F + W W^T) P`` in
    factored form (positional call ``(Z, F, M, W, J)``)."""
    dv.require_cuda()
    Ft, Mt = dv.DeviceCSR(sps.csr_matrix(amat.T)), dv.DeviceCSR(sps.csr_matrix(mmat.T))
    Zd = dv.to_dev(Z)
    ftz = Ft.matmul(Zd)
    if umat is not None and vmat is not None:
        utz = dv.DeviceCSR(sps.csr_matrix(umat).T).matmul(Zd)
        ftz = ftz - dv.tall_gemm(dv.to_dev(_dense(vmat).T), utz)
    mtz = Mt.matmul(Zd)
    Wd = dv.to_dev(_dense(wmat))
    stacked = torch.cat([ftz, mtz, Wd], dim=1).contiguous()
    NV = mmat.shape[0]
    lu = dv.LU(dv.sadpnt_matrix(mmat, jmat))
    prj = dv.DeviceCSR(mmat).matmul(lu.solve(stacked, nrows_out=NV))    # P^T [..]
    ka, kb = ftz.shape[1], mtz.shape[1]
    G = dv.gram(prj, prj)
    DG = torch.cat([G[ka:ka+kb, :], G[:ka, :], G[ka+kb:, :]], dim=0)
    return float((DG*DG.t()).sum().item())
