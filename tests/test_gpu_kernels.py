"""Kernel-level parity: every C-ABI entry point against numpy/scipy on seeded inputs.
All calls go through the ctypes C ABI (optconpy_b200.device)."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return np.linalg.norm(a - b)/max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope='module')
def dv():
    from optconpy_b200 import device
    device.require_cuda()
    return device


@pytest.fixture(scope='module')
def sad(cav10):
    from optconpy_b200 import problems as pb, device
    M, A, J = cav10['M'], cav10['A'], cav10['J']
    Nc = pb.convection_matrix(cav10, pb.analytic_vortex)
    Ft = -(0.5*M.T + 0.05*(A.T + Nc.T))
    return device.sadpnt_matrix(Ft - 1.0*M.T, J)


@pytest.mark.parametrize('k', [1, 5, 32, 33, 66, 130])
def test_spmm(dv, cav10, k):
    rng = np.random.default_rng(k)
    S = cav10['J']                      # rectangular NP x NV
    X = rng.standard_normal((S.shape[1], k))
    Y0 = rng.standard_normal((S.shape[0], k))
    Sd = dv.DeviceCSR(S)
    Y = dv.to_host(Sd.matmul(dv.to_dev(X)))
    assert _relerr(Y, S @ X) < 1e-14
    out = dv.to_dev(Y0)
    Sd.matmul(dv.to_dev(X), alpha=-0.5, beta=2.0, out=out)
    assert _relerr(dv.to_host(out), -0.5*(S @ X) + 2.0*Y0) < 1e-14


def test_spmm_strided_and_empty_rows(dv):
    rng = np.random.default_rng(0)
    S = sps.random(300, 200, density=0.02, random_state=1, format='csr')  # has empty rows
    Xbig = rng.standard_normal((200, 40))
    Xd = dv.to_dev(Xbig)[:, 3:20]       # non-contiguous view, ld = 40
    Y = dv.to_host(dv.DeviceCSR(S).matmul(Xd))
    assert _relerr(Y, S @ Xbig[:, 3:20]) < 1e-14


@pytest.mark.parametrize('k', [1, 3, 4, 7, 66])
def test_lu_solve_matches_superlu(dv, sad, k):
    rng = np.random.default_rng(10 + k)
    n = sad.shape[0]
    B = rng.standard_normal((n, k))
    ref = spsla.splu(sad).solve(B)       # the reference's factorisation (COLAMD default)
    lu = dv.LU(sad)
    X = dv.to_host(lu.solve(dv.to_dev(B)))
    assert _relerr(X, ref) < 1e-11
    assert np.linalg.norm(sad @ X - B)/np.linalg.norm(B) < 1e-12
    # same factors as the oracle would use: default SuperLU options
    lu2 = dv.LU(sad, lu_options={})
    X2 = dv.to_host(lu2.solve(dv.to_dev(B)))
    assert _relerr(X2, ref) < 1e-12


def test_lu_solve_partial_rows(dv, sad, cav10):
    rng = np.random.default_rng(3)
    NV, n = cav10['NV'], sad.shape[0]
    R = rng.standard_normal((NV, 9))
    full = np.vstack([R, np.zeros((n - NV, 9))])
    ref = spsla.splu(sad).solve(full)[:NV]
    lu = dv.LU(sad)
    X = dv.to_host(lu.solve(dv.to_dev(R), nrows_out=NV))
    assert X.shape == (NV, 9)
    assert _relerr(X, ref) < 1e-11
    # in place
    Rd = dv.to_dev(R)
    lu.solve(Rd, nrows_out=NV, out=Rd)
    assert _relerr(dv.to_host(Rd), ref) < 1e-11


def test_lu_solve_wide_executor_small_n(dv, sad):
    """Image with the flat program: blocks of >= 640 columns go through the all-columns
    executor, narrower ones through the cluster kernel; both agree with SuperLU."""
    rng = np.random.default_rng(21)
    n = sad.shape[0]
    lu = dv.LU(sad, wide=True)
    ref_lu = spsla.splu(sad)
    lib = dv.require_cuda()
    for k in (7, 200, 641):
        B = rng.standard_normal((n, k))
        assert (lib.ocb_lu_solve_ws_bytes(lu.handle, k) > 0) == (k >= 640)
        X = dv.to_host(lu.solve(dv.to_dev(B)))
        assert _relerr(X, ref_lu.solve(B)) < 1e-11
    # partial rows in / out through the wide path
    NV = n - 100
    R = rng.standard_normal((NV, 700))
    full = np.vstack([R, np.zeros((n - NV, 700))])
    X = dv.to_host(lu.solve(dv.to_dev(R), nrows_out=NV))
    assert X.shape == (NV, 700) and _relerr(X, ref_lu.solve(full)[:NV]) < 1e-11


@pytest.mark.parametrize('wide', [False, True])
def test_lu_solve_one_step_supernodes(dv, sad, wide, monkeypatch):
    """OCB_MERGE=1 (flags bit 2 of ocb_lu_pack_host): the program with one-step supernodes has
    clearly fewer sub-levels and gives the same solutions through the cluster kernel (narrow and
    two-wave blocks) and through the wide executor."""
    rng = np.random.default_rng(31)
    n = sad.shape[0]
    ref_lu = spsla.splu(sad)
    monkeypatch.setenv('OCB_MERGE', '0')
    two = dv.LU(sad, wide=wide)
    monkeypatch.setenv('OCB_MERGE', '1')
    one = dv.LU(sad, wide=wide)
    lv = lambda lu: lu.info['levelsL'] + lu.info['levelsU']
    assert lv(one) <= 0.8*lv(two) and one.info['program_rows'] == two.info['program_rows']
    for k in ((3, 40, 160) if not wide else (5, 700)):
        B = rng.standard_normal((n, k))
        ref = ref_lu.solve(B)
        Bd = dv.to_dev(B)
        X1, X2 = dv.to_host(one.solve(Bd)), dv.to_host(two.solve(Bd))
        assert _relerr(X1, ref) < 1e-11 and _relerr(X1, X2) < 1e-12


def test_lu_solve_global_panel_path(dv):
    """n large enough that the column panel does not fit shared memory: wide executor."""
    from optconpy_b200 import problems as pb
    p = pb.drivcav_problem(64, 1e-2)
    S = dv.sadpnt_matrix(p['M'] + 0.01*p['A'], p['J'])
    n = S.shape[0]
    assert n*2*8 > 227*1024
    rng = np.random.default_rng(5)
    lu = dv.LU(S)
    slu = spsla.splu(S)
    for k in (11, 48, 130):          # one, two and four column tiles per warp (panel executor)
        B = rng.standard_normal((n, k))
        X = dv.to_host(lu.solve(dv.to_dev(B)))
        assert np.linalg.norm(S @ X - B)/np.linalg.norm(B) < 1e-11
        assert _relerr(X, slu.solve(B)) < 1e-10
        X2 = dv.to_host(lu.solve(dv.to_dev(B)))
        assert np.array_equal(X, X2)                  # deterministic


def test_gram_symmetric_half(dv):
    """G = Z^T Z with the same block on both sides: only the tiles on and below the diagonal are
    computed on the tensor pipe, the rest is mirrored."""
    rng = np.random.default_rng(77)
    Z = rng.standard_normal((3000, 200))
    Zd = dv.to_dev(Z)
    G = dv.to_host(dv.gram(Zd, Zd))
    assert _relerr(G, Z.T @ Z) < 1e-13
    assert np.array_equal(G, G.T)


@pytest.mark.parametrize('shape', [(722, 5, 5), (1000, 66, 66), (4802, 130, 8), (333, 7, 129)])
def test_gram(dv, shape):
    n, ka, kb = shape
    rng = np.random.default_rng(n)
    Z, W = rng.standard_normal((n, ka)), rng.standard_normal((n, kb))
    G = dv.to_host(dv.gram(dv.to_dev(Z), dv.to_dev(W)))
    assert _relerr(G, Z.T @ W) < 1e-13
    G2 = dv.to_host(dv.gram(dv.to_dev(Z), dv.to_dev(W)))
    assert np.array_equal(G, G2)         # deterministic


@pytest.mark.parametrize('shape', [(722, 5, 3), (1000, 70, 66), (4802, 300, 50), (65, 33, 1)])
def test_tall_gemm(dv, shape):
    n, k, kc = shape
    rng = np.random.default_rng(n)
    Z, T = rng.standard_normal((n, k)), rng.standard_normal((k, kc))
    Cm = dv.to_host(dv.tall_gemm(dv.to_dev(Z), dv.to_dev(T), alpha=0.5))
    assert _relerr(Cm, 0.5*Z @ T) < 1e-13


@pytest.mark.parametrize('k', [1, 2, 7, 64, 101, 256])
def test_sym_eig(dv, k):
    rng = np.random.default_rng(k)
    A = rng.standard_normal((k + 20, k))*np.logspace(0, -6, k)[None, :]
    G = A.T @ A
    lam, V, sweeps = dv.sym_eig(dv.to_dev(G))
    lam, V = dv.to_host(lam), dv.to_host(V)
    ref = np.linalg.eigvalsh(G)[::-1]
    assert np.max(np.abs(lam - ref)) < 1e-13*ref[0]
    assert np.linalg.norm(V.T @ V - np.eye(k)) < 1e-13*max(k, 10)
    assert np.linalg.norm(V @ np.diag(lam) @ V.T - G) < 1e-12*ref[0]*k
    assert np.all(np.diff(lam) <= 0)


def test_compress_matches_svd(dv):
    rng = np.random.default_rng(7)
    n, K, r = 900, 1500, 120
    Z = (rng.standard_normal((n, r))*np.logspace(-1, -9, r)[None, :]) @ rng.standard_normal((r, K))/30.0
    s = np.linalg.svd(Z, compute_uv=False)
    U, sv, Vt = np.linalg.svd(Z, full_matrices=False)
    for thresh, k in ((5e-5, 50), (1e-6, None), (None, 17)):
        keep = len(sv) if thresh is None else int(np.sum(sv > thresh))
        keep = keep if k is None else min(keep, k)
        ref = Z @ Vt[:keep].T
        Zc, info = dv.compress(dv.to_dev(Z), thresh=thresh, k=k)
        Zc = dv.to_host(Zc)
        assert Zc.shape[1] == keep
        assert _relerr(Zc @ Zc.T, ref @ ref.T) < 1e-9
        assert _relerr(dv.to_host(info['sigma'])[:10], s[:10]) < 1e-10


def test_smw_solve(dv, sad, cav10):
    rng = np.random.default_rng(2)
    NV, n = cav10['NV'], sad.shape[0]
    U = rng.standard_normal((NV, 8))*1e-2
    V = sps.random(8, NV, density=0.05, random_state=3, format='csr')
    R = rng.standard_normal((NV, 6))
    Ue = np.vstack([U, np.zeros((n-NV, 8))])
    Ve = sps.hstack([V, sps.csr_matrix((8, n-NV))]).tocsr()
    ref = np.linalg.solve(sad.toarray() - Ue @ Ve.toarray(), np.vstack([R, np.zeros((n-NV, 6))]))
    lu = dv.LU(sad)
    X = dv.to_host(lu.smw_solve(dv.to_dev(R), NV, Ufb=dv.to_dev(U), Vt=dv.DeviceCSR(V)))
    assert X.shape == (n, 6)
    assert _relerr(X, ref) < 1e-10


def test_feedback(dv, cav10):
    rng = np.random.default_rng(4)
    NV = cav10['NV']
    Z = rng.standard_normal((NV, 37))
    tB = rng.standard_normal((NV, 8))
    MT = cav10['M'].T.tocsr()
    out = dv.to_host(dv.feedback(dv.DeviceCSR(MT), dv.to_dev(Z), dv.to_dev(tB), alpha=-1.0))
    assert _relerr(out, -(MT @ (Z @ (Z.T @ tB)))) < 1e-13


def test_errors_are_loud(dv, sad):
    from optconpy_b200 import _cabi
    lu = dv.LU(sad)
    import torch
    B = torch.zeros((sad.shape[0] + 1, 2), dtype=torch.float64, device='cuda')
    with pytest.raises(_cabi.OcbError):
        lu.solve(B)


def test_fp64_peak_microbenchmark(dv):
    """The measured FP64 denominators (DMMA tensor pipe, DFMA pipe) are sane B200 numbers."""
    dm, df = dv.fp64_peak('dmma'), dv.fp64_peak('dfma')
    assert 5.0 < dm < 120.0 and 5.0 < df < 120.0, (dm, df)


def test_device_copy_cache_of_returned_factors(dv):
    """compress_Zsvd registers the device copy of the array it returns; handing that very array
    back (as the reference driver does) skips the upload, a modified array is uploaded afresh."""
    import optconpy_b200.proj_ric_utils as gpru
    rng = np.random.default_rng(3)
    Z = rng.standard_normal((900, 40)) @ rng.standard_normal((40, 120))
    zc = gpru.compress_Zsvd(Z, thresh=1e-8)
    before = dv.STATS['h2d_bytes']
    t1 = dv.to_dev(zc)
    assert dv.STATS['h2d_bytes'] == before                       # served from the cache
    assert _relerr(dv.to_host(t1), zc) == 0.0
    zc[3, 2] += 1.0                                              # caller modified it in place
    t2 = dv.to_dev(zc)
    assert dv.STATS['h2d_bytes'] > before and _relerr(dv.to_host(t2), zc) == 0.0
    other = zc.copy()                                            # another object: never served from the cache
    b2 = dv.STATS['h2d_bytes']
    dv.to_dev(other)
    assert dv.STATS['h2d_bytes'] > b2
