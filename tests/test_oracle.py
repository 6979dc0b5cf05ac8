"""CPU suite, part 1: the oracle (oracle/) against everything that can pin it.

There are no upstream golden vectors (Fact 3: parity unpinned), so the
oracle is held to
  * the five identities of the reference's only test
    (``tests/test_units_compfacres_compress.py:85-106``), restated here on the seeded
    synthetic cavity with the same kind of inputs (``F = -M - 0.1 A - sprand``,
    ``W`` random, ``adi_max_steps=50``, ``adi_newZ_reltol=1e-11``),
  * the equations the driver encodes: Lyapunov / Riccati residuals -> 0,
  * its own committed fixtures under ``tests/golden/`` (``tools/gen_golden.py``).
"""
import os

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

from oracle import lin_alg_utils as olau, proj_ric_utils as opru
from optconpy_b200 import problems as pb, scenarios as sc, dre_stepper as ds

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _explicit_projector(M, J):
    Mlu = spsla.factorized(sps.csc_matrix(M))
    MinvJt = olau.app_luinv_to_spmat(Mlu, J.T)
    Sinv = np.linalg.inv(J @ MinvJt)
    return np.eye(M.shape[0]) - MinvJt @ (Sinv @ J.toarray())


def test_reference_identities(cav10):
    """tests/test_units_compfacres_compress.py:49-106 on the N=10 cavity, seeded."""
    M, A, J = cav10['M'], cav10['A'], cav10['J']
    NV, NY = cav10['NV'], 5
    F = -M - 0.1*A - sps.random(NV, NV, density=0.03, format='csr', random_state=11)
    W = np.random.default_rng(12).standard_normal((NV, NY))
    d = dict(adi_max_steps=50, adi_newZ_reltol=1e-11, nwtn_max_steps=24,
             nwtn_upd_reltol=4e-7, nwtn_upd_abstol=4e-7, full_upd_norm_check=True)
    Z = opru.solve_proj_lyap_stein(amat=F, mmat=M, jmat=J, wmat=W, adi_dict=d)['zfac']
    MtZ = M.T @ Z
    MtXM = MtZ @ MtZ.T
    FtXM = F.T @ (Z @ (Z.T @ M.toarray()))
    P = _explicit_projector(M, J)
    PtW = P.T @ W
    ProjRes = P.T @ FtXM @ P + P.T @ FtXM.T @ P + PtW @ PtW.T
    resn = np.linalg.norm(ProjRes)
    ownresn = np.sqrt(opru.comp_proj_lyap_res_norm(Z, F, M, W, J))
    # 1. smart small-Gram norm
    assert np.allclose(np.linalg.norm(MtXM), np.linalg.norm(MtZ.T @ MtZ))
    # 2. explicit projected residual == factored residual
    assert np.allclose(resn, ownresn)
    Zred = opru.compress_Zsvd(Z, k=None, thresh=1e-6)
    assert Zred.shape[1] <= Z.shape[1]
    MtZr = M.T @ Zred
    MtXMr = MtZr @ MtZr.T
    # 3. the reduced factor stays projected
    assert np.allclose(MtXMr, P.T @ MtXMr @ P)
    # 4. the Gram norm is unchanged by the compression
    assert np.allclose(np.linalg.norm(MtZ.T @ MtZ), np.linalg.norm(MtZr.T @ MtZr))
    # 5. residual of the compressed factor == residual of the full one
    ownresr = np.sqrt(opru.comp_proj_lyap_res_norm(Zred, F, M, W, J))
    assert np.allclose(ownresr, resn)


def test_lyapunov_residual_goes_to_zero(cav6):
    """Converged ADI solves P^T(F^T X M + M^T X F + W W^T)P = 0."""
    M, A, J = cav6['M'], cav6['A'], cav6['J']
    Nc = pb.convection_matrix(cav6, pb.analytic_vortex)
    F = -(0.5*M + 0.05*(A + Nc))
    W = np.random.default_rng(0).standard_normal((cav6['NV'], 3))
    d = dict(adi_max_steps=200, adi_newZ_reltol=1e-12, ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0])
    res = opru.solve_proj_lyap_stein(amat=F, mmat=M, jmat=J, wmat=W, adi_dict=d)
    Z = res['zfac']
    P = _explicit_projector(M, J)
    PtW = P.T @ W
    r0 = np.linalg.norm(PtW @ PtW.T)
    X, Md, Fd = Z @ Z.T, M.toarray(), F.toarray()
    R = P.T @ (Fd.T @ X @ Md + Md.T @ X @ Fd) @ P + PtW @ PtW.T
    assert np.linalg.norm(R) < 1e-9*r0
    # the factored form agrees down to its cancellation floor (it is a squared norm)
    assert abs(opru.comp_proj_lyap_res_norm(Z, F, M, W, J)) < 1e-12*r0**2
    rel = res['adi_rel_newZ_norms']
    assert rel[0] == 1.0 and rel[-1] <= 1e-12 and len(rel)*3 == Z.shape[1]
    # transposed call convention (solve_dae_ric.py:152-153) gives the same factor
    res_t = opru.solve_proj_lyap_stein(amat=F.T, mmat=M.T, jmat=J, wmat=W, adi_dict=d,
                                       transposed=True)
    assert np.allclose(res_t['zfac'], Z, rtol=1e-10, atol=1e-14)


def test_riccati_residual_and_newton_convergence(cav6):
    """Newton-ADI solves the projected ARE of solve_dae_ric.py:147-158; the update norms
    contract at least quadratically towards the end."""
    M, A, J = cav6['M'], cav6['A'], cav6['J']
    NV = cav6['NV']
    cs = pb.control_setup(cav6, olau, alphau=1e-4)
    tau = 0.05
    Nc = pb.convection_matrix(cav6, pb.analytic_vortex)
    F = -(0.5*M + tau*(A + Nc))
    B = np.sqrt(tau)*cs['tb_mat']
    W = np.sqrt(tau)*cs['trct_mat']
    d = dict(adi_max_steps=200, adi_newZ_reltol=1e-11, nwtn_max_steps=12,
             nwtn_upd_reltol=1e-10, nwtn_upd_abstol=1e-14, full_upd_norm_check=True,
             ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0])
    res = opru.proj_alg_ric_newtonadi(mmat=M.T, amat=F.T, transposed=True, jmat=J,
                                      bmat=B, wmat=W, z0=None, nwtn_adi_dict=d)
    Z = res['zfac']
    X = Z @ Z.T
    Md, Fd, Bd = M.toarray(), F.toarray(), np.asarray(B.todense())
    P = _explicit_projector(M, J)
    R = Fd.T @ X @ Md + Md.T @ X @ Fd - Md.T @ X @ Bd @ Bd.T @ X @ Md + W @ W.T
    PtW = P.T @ W
    assert np.linalg.norm(P.T @ R @ P) < 1e-7*np.linalg.norm(PtW @ PtW.T)
    f = res['nwtn_upd_fnorms']
    assert len(f) >= 2 and f[-1] < 1e-6*f[0]
    # non-transposed convention (optcont_main.py:488-492) is the same equation
    res2 = opru.proj_alg_ric_newtonadi(mmat=M, amat=F, jmat=J, bmat=B, wmat=W, z0=None,
                                       nwtn_adi_dict=d)
    assert res2['adi_steps'] == res['adi_steps']
    assert np.allclose(res2['zfac'], Z, rtol=1e-9, atol=1e-13)


def test_smw_and_projection(cav6):
    M, A, J = cav6['M'], cav6['A'], cav6['J']
    NV, NP = cav6['NV'], cav6['NP']
    rng = np.random.default_rng(3)
    rhs = rng.standard_normal((NV, 2))
    U = rng.standard_normal((NV, 4))*1e-2
    V = sps.random(4, NV, density=0.05, random_state=5, format='csr')
    amat = M + 0.1*A
    sol = olau.solve_sadpnt_smw(amat=amat, jmat=J, rhsv=rhs, umat=U, vmat=V)
    assert sol.shape == (NV+NP, 2)
    full = sps.bmat([[sps.csr_matrix(amat - U @ V.toarray()), J.T], [J, None]]).toarray()
    ref = np.linalg.solve(full, np.vstack([rhs, np.zeros((NP, 2))]))
    assert np.linalg.norm(sol - ref) < 1e-10*np.linalg.norm(ref)
    P = _explicit_projector(M, J)
    assert np.allclose(olau.app_prj_via_sadpnt(amat=M, jmat=J, rhsv=rhs), P @ rhs, atol=1e-11)
    assert np.allclose(olau.app_prj_via_sadpnt(amat=M, jmat=J, rhsv=rhs, transposedprj=True),
                       P.T @ rhs, atol=1e-11)
    # factored norms
    Z1, Z2 = rng.standard_normal((NV, 5)), rng.standard_normal((NV, 3))
    assert np.isclose(olau.comp_sqfnrm_factrd_diff(Z1, Z2),
                      np.linalg.norm(Z1 @ Z1.T - Z2 @ Z2.T)**2)
    assert np.isclose(olau.comp_sqfnrm_factrd_sum(Z1, Z2),
                      np.linalg.norm(Z1 @ Z1.T + Z2 @ Z2.T)**2)
    A3, B3, C3 = (rng.standard_normal((40, 3)), rng.standard_normal((40, 3)),
                  rng.standard_normal((40, 2)))
    assert np.isclose(olau.comp_sqfnrm_factrd_lyap_res(A3, B3, C3),
                      np.linalg.norm(A3 @ B3.T + B3 @ A3.T + C3 @ C3.T)**2)
    # square roots: (R L^-T)(R L^-T)^T = R M^-1 R^T
    Rm = cav6_small_mass(4)
    X = rng.standard_normal((6, 4))
    Y = olau.apply_invsqrt_fromright(Rm, X)
    assert np.allclose(Y @ Y.T, X @ np.linalg.solve(Rm, X.T))
    Y2 = olau.apply_sqrt_fromright(Rm, X)
    assert np.allclose(Y2 @ Y2.T, X @ Rm @ X.T)


def cav6_small_mass(n):
    T = np.diag(np.full(n, 4.0)) + np.diag(np.ones(n-1), 1) + np.diag(np.ones(n-1), -1)
    return T/6.0


def test_golden_lyapunov_fixture():
    """tests/golden/lyap_cav6.npz (tools/gen_golden.py): oracle regression pin."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'gen_golden', os.path.join(os.path.dirname(GOLD), '..', 'tools', 'gen_golden.py'))
    gg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gg)
    g = np.load(os.path.join(GOLD, 'lyap_cav6.npz'))
    prob, F, W, d = gg.lyap_case()
    assert np.allclose([prob['M'].sum(), prob['A'].sum(), abs(prob['J']).sum()],
                       g['matrix_sums'], rtol=1e-12)
    res = opru.solve_proj_lyap_stein(amat=F, mmat=prob['M'], jmat=prob['J'], wmat=W, adi_dict=d)
    Z = res['zfac']
    assert Z.shape[1] == int(g['zcols'])
    assert np.allclose(res['adi_rel_newZ_norms'], g['rel_norms'], rtol=1e-6, atol=1e-14)
    assert np.allclose(np.linalg.svd(Z, compute_uv=False)[:20], g['sv'], rtol=1e-8, atol=1e-12)
    Zc = opru.compress_Zsvd(Z, thresh=1e-6)
    assert Zc.shape[1] == int(g['zc_cols'])
    G = Zc.T @ (prob['M'] @ Zc)
    assert np.allclose(np.sort(np.linalg.eigvalsh(G))[::-1][:20], g['gram_m_eigs'],
                       rtol=1e-8, atol=1e-12)


def test_golden_dre_fixture():
    """tests/golden/dre_cav6.npz: three backward DRE steps through the restated driver."""
    g = np.load(os.path.join(GOLD, 'dre_cav6.npz'))
    prob6 = pb.drivcav_problem(6, 1e-2)
    cs = pb.control_setup(prob6, olau, alphau=1e-9)
    tmesh = pb.get_tint(0.0, 1.0, 3)
    kw = sc.dre_kwargs(prob6, cs, tmesh, dict(sc.DEFAULT_NWTN_ADI, adi_max_steps=120), 1e-3,
                       sc._ystar_sin(cs['NY']))
    store, info = ds.MemStore(), []
    fb = ds.solve_flow_daeric(lau=olau, pru=opru, store=store, stepinfo=info, **kw)
    ts = sorted(fb)
    assert np.allclose(ts, g['tmesh'])
    assert [sum(i['adi_steps']) for i in info] == list(g['adi_steps'])
    assert [i['zc_cols'] for i in info] == list(g['zc_cols'])
    for i, t in enumerate(ts):
        a, b = store[fb[t]['mtxtb']], g['mtxtb'][i]
        assert np.linalg.norm(a - b) <= 1e-8*np.linalg.norm(b)
        a, b = store[fb[t]['w']], g['w'][i]
        assert np.linalg.norm(a - b) <= 1e-8*np.linalg.norm(b)
