"""Path-level parity of the CUDA modules against the CPU oracle on the same seeded
inputs: factor products Z Z^T and gains within 1e-9 relative Frobenius error, ADI
iteration counts and relative-norm histories equal, DRE trajectories within 1e-8."""
import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu

TOL_FACTOR = 1e-9      # north_star: Z Z^T and feedback gains
TOL_TRAJ = 1e-8        # north_star: DRE trajectory quantities


def _relerr(a, b):
    return np.linalg.norm(a - b)/max(np.linalg.norm(b), 1e-300)


def _zzt_relerr(Za, Zb):
    """||Za Za^T - Zb Zb^T||_F / ||Zb Zb^T||_F, formed without the cancellation of the
    three-Gram formula: [Za Zb] = Q R, then || R diag(I,-I) R^T ||_F."""
    ka = Za.shape[1]
    R = np.linalg.qr(np.hstack([Za, Zb]), mode='r')
    D = R[:, :ka] @ R[:, :ka].T - R[:, ka:] @ R[:, ka:].T
    return np.linalg.norm(D)/np.linalg.norm(Zb.T @ Zb)


@pytest.fixture(scope='module')
def mods():
    import optconpy_b200.lin_alg_utils as glau
    import optconpy_b200.proj_ric_utils as gpru
    from oracle import lin_alg_utils as olau, proj_ric_utils as opru
    return glau, gpru, olau, opru


@pytest.fixture(scope='module')
def lyap_setup(cav10):
    from optconpy_b200 import problems as pb
    M, A, J = cav10['M'], cav10['A'], cav10['J']
    Nc = pb.convection_matrix(cav10, pb.analytic_vortex)
    tau = 0.05
    F = -(0.5*M + tau*(A + Nc))
    rng = np.random.default_rng(0)
    W = rng.standard_normal((cav10['NV'], 5))
    return M, F, J, W


def test_lau_small_functions(mods, cav10):
    glau, gpru, olau, opru = mods
    M, J = cav10['M'], cav10['J']
    rng = np.random.default_rng(1)
    R = rng.standard_normal((cav10['NV'], 3))
    assert _relerr(glau.apply_massinv(M, R), olau.apply_massinv(M, R)) < 1e-12
    Rs = sps.random(cav10['NV'], 4, density=0.05, random_state=2, format='csr')
    a, b = glau.apply_massinv(M, Rs, output='sparse'), olau.apply_massinv(M, Rs, output='sparse')
    assert sps.issparse(a) and _relerr(a.toarray(), b.toarray()) < 1e-12
    for tp in (True, False):
        pa = glau.app_prj_via_sadpnt(amat=M, jmat=J, rhsv=R, transposedprj=tp)
        pb_ = olau.app_prj_via_sadpnt(amat=M, jmat=J, rhsv=R, transposedprj=tp)
        assert _relerr(pa, pb_) < 1e-11
    # projector property: J M^-1 (P^T R) = 0
    pt = glau.app_prj_via_sadpnt(amat=M, jmat=J, rhsv=R, transposedprj=True)
    assert np.linalg.norm(J @ olau.apply_massinv(M, pt)) < 1e-11*np.linalg.norm(R)
    assert _relerr(glau.mm_dnssps(M, R), M @ R) < 1e-14
    Z1, Z2 = rng.standard_normal((cav10['NV'], 7)), rng.standard_normal((cav10['NV'], 4))
    assert abs(glau.comp_sqfnrm_factrd_diff(Z1, Z2) - olau.comp_sqfnrm_factrd_diff(Z1, Z2)) \
        < 1e-11*olau.comp_sqfnrm_factrd_sum(Z1, Z2)
    assert abs(glau.comp_sqfnrm_factrd_sum(Z1, Z2) - olau.comp_sqfnrm_factrd_sum(Z1, Z2)) \
        < 1e-11*olau.comp_sqfnrm_factrd_sum(Z1, Z2)


def test_solve_sadpnt_smw(mods, cav10):
    glau, gpru, olau, opru = mods
    M, A, J = cav10['M'], cav10['A'], cav10['J']
    rng = np.random.default_rng(3)
    NV = cav10['NV']
    rhs = rng.standard_normal((NV, 1))
    U = rng.standard_normal((NV, 8))*1e-3
    V = sps.random(8, NV, density=0.03, random_state=5, format='csr')
    amat = M.T + 0.1*A.T
    ref = olau.solve_sadpnt_smw(amat=amat, jmat=J, rhsv=rhs, umat=U, vmat=V)
    got = glau.solve_sadpnt_smw(amat=amat, jmat=J, rhsv=rhs, umat=U, vmat=V)
    assert got.shape == ref.shape == (NV + cav10['NP'], 1)
    assert _relerr(got, ref) < 1e-10
    ref0 = olau.solve_sadpnt_smw(amat=amat, jmat=J, rhsv=rhs)
    got0 = glau.solve_sadpnt_smw(amat=amat, jmat=J, rhsv=rhs)
    assert _relerr(got0, ref0) < 1e-10


def test_stein_parity(mods, lyap_setup):
    glau, gpru, olau, opru = mods
    M, F, J, W = lyap_setup
    d = dict(adi_max_steps=80, adi_newZ_reltol=1e-9, ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0])
    ref = opru.solve_proj_lyap_stein(amat=F, mmat=M, jmat=J, wmat=W, adi_dict=d)
    got = gpru.solve_proj_lyap_stein(amat=F, mmat=M, jmat=J, wmat=W, adi_dict=d)
    assert got['zfac'].shape == ref['zfac'].shape            # same iteration count
    assert len(got['adi_rel_newZ_norms']) == len(ref['adi_rel_newZ_norms'])
    assert np.allclose(got['adi_rel_newZ_norms'], ref['adi_rel_newZ_norms'], rtol=1e-7, atol=0)
    assert _zzt_relerr(got['zfac'], ref['zfac']) < TOL_FACTOR
    # the five identities of the reference's test, on the CUDA result
    Z = got['zfac']
    scale = np.linalg.norm(W.T @ W)**2
    res_full = gpru.comp_proj_lyap_res_norm(Z, F, M, W, J)
    res_ref = opru.comp_proj_lyap_res_norm(Z, F, M, W, J)
    assert abs(res_full - res_ref) <= 1e-12*scale        # converged: both at the noise floor
    assert res_ref <= 1e-12*scale                        # ... and the residual is ~0
    Zpart = Z[:, :4*W.shape[1]]                          # unconverged: a substantial residual
    rp_g = gpru.comp_proj_lyap_res_norm(Zpart, F, M, W, J)
    rp_o = opru.comp_proj_lyap_res_norm(Zpart, F, M, W, J)
    assert rp_o > 1e-6*scale and abs(rp_g - rp_o) <= 1e-9*rp_o
    Zr = gpru.compress_Zsvd(Z, k=None, thresh=1e-6)
    MtZ, MtZr = M.T @ Z, M.T @ Zr
    assert np.allclose(np.linalg.norm(MtZ.T @ MtZ), np.linalg.norm(MtZr.T @ MtZr))
    Pt = olau.app_prj_via_sadpnt(amat=M, jmat=J, rhsv=MtZr, transposedprj=True)
    assert np.allclose(Pt, MtZr, atol=1e-8*np.abs(MtZr).max())


def test_stein_transposed_with_lowrank(mods, lyap_setup, cav10):
    glau, gpru, olau, opru = mods
    M, F, J, W = lyap_setup
    rng = np.random.default_rng(9)
    NV = cav10['NV']
    B = sps.random(NV, 8, density=0.02, random_state=4, format='csr')
    K = rng.standard_normal((8, NV))*1e-2
    d = dict(adi_max_steps=60, adi_newZ_reltol=1e-8, ms=[-5.0, -2.0, -1.0])
    kw = dict(amat=F.T, mmat=M.T, jmat=J, wmat=W, umat=B, vmat=K, transposed=True, adi_dict=d)
    ref, got = opru.solve_proj_lyap_stein(**kw), gpru.solve_proj_lyap_stein(**kw)
    assert got['zfac'].shape == ref['zfac'].shape
    assert _zzt_relerr(got['zfac'], ref['zfac']) < TOL_FACTOR


def test_newtonadi_parity(mods, lyap_setup, cav10):
    glau, gpru, olau, opru = mods
    from optconpy_b200 import problems as pb
    M, F, J, _ = lyap_setup
    cs = pb.control_setup(cav10, olau, alphau=1e-4)
    d = dict(adi_max_steps=120, adi_newZ_reltol=1e-9, nwtn_max_steps=12,
             nwtn_upd_reltol=1e-9, nwtn_upd_abstol=1e-12, full_upd_norm_check=False,
             ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0])
    tau = 0.05
    kw = dict(mmat=M.T, amat=F.T, transposed=True, jmat=J, bmat=np.sqrt(tau)*cs['tb_mat'],
              wmat=np.sqrt(tau)*cs['trct_mat'], z0=None, nwtn_adi_dict=d)
    ref, got = opru.proj_alg_ric_newtonadi(**kw), gpru.proj_alg_ric_newtonadi(**kw)
    assert got['adi_steps'] == ref['adi_steps']
    assert len(got['nwtn_upd_fnorms']) == len(ref['nwtn_upd_fnorms'])
    assert _zzt_relerr(got['zfac'], ref['zfac']) < TOL_FACTOR
    for full in (True,):
        d2 = dict(d, full_upd_norm_check=full, nwtn_max_steps=3)
        kw2 = dict(kw, nwtn_adi_dict=d2, z0=ref['zfac'][:, :16])
        r2, g2 = opru.proj_alg_ric_newtonadi(**kw2), gpru.proj_alg_ric_newtonadi(**kw2)
        assert g2['adi_steps'] == r2['adi_steps']
        # the full check is sqrt of a difference of squared norms: floor ~ sqrt(eps)*||X||
        assert np.allclose(g2['nwtn_upd_fnorms'], r2['nwtn_upd_fnorms'], rtol=1e-5,
                           atol=2e-7*r2['nwtn_upd_fnorms'][0])
        assert _zzt_relerr(g2['zfac'], r2['zfac']) < TOL_FACTOR
    # feedback gains
    tb = cs['tb_mat']
    ga, gb = gpru.get_mTzzTtb(M.T, got['zfac'], tb), opru.get_mTzzTtb(M.T, ref['zfac'], tb)
    assert _relerr(ga, gb) < TOL_FACTOR
    fv = np.random.default_rng(1).standard_normal((cav10['NV'], 1))
    assert _relerr(gpru.get_mTzzTtb(M.T, got['zfac'], fv), opru.get_mTzzTtb(M.T, ref['zfac'], fv)) < TOL_FACTOR
    # compressed factor products
    zc_g = gpru.compress_Zsvd(got['zfac'], thresh=5e-5, k=50)
    zc_o = opru.compress_Zsvd(ref['zfac'], thresh=5e-5, k=50)
    assert zc_g.shape == zc_o.shape
    assert _zzt_relerr(zc_g, zc_o) < TOL_FACTOR


def test_dre_trajectory_parity(mods):
    """Config 1 (driven cavity N=10, optcon_nse defaults), 3 backward steps."""
    glau, gpru, olau, opru = mods
    from optconpy_b200 import scenarios as sc, dre_stepper as ds
    prob, cs, kw = sc.config1(olau, Nts=3)
    so, sg = ds.MemStore(), ds.MemStore()
    io, ig = [], []
    fo = ds.solve_flow_daeric(lau=olau, pru=opru, store=so, stepinfo=io, **dict(kw, gtdtstrargs=dict(kw['gtdtstrargs'])))
    fg = ds.solve_flow_daeric(lau=glau, pru=gpru, store=sg, stepinfo=ig, **dict(kw, gtdtstrargs=dict(kw['gtdtstrargs'])))
    assert sorted(fo) == sorted(fg)
    for a, b in zip(io, ig):
        assert a['adi_steps'] == b['adi_steps']
        assert a['zc_cols'] == b['zc_cols']
    for t in fo:
        assert _relerr(sg[fg[t]['mtxtb']], so[fo[t]['mtxtb']]) < TOL_TRAJ
        assert _relerr(sg[fg[t]['w']], so[fo[t]['w']]) < TOL_TRAJ
        ko = fo[t]['mtxtb'].replace('__mtxtb', '__Z')
        assert _zzt_relerr(sg[ko], so[ko]) < TOL_FACTOR


def test_device_resident_loop_matches_host_api(mods):
    """dre_device (factor kept in HBM between steps) == dre_stepper through the host API."""
    glau, gpru, olau, opru = mods
    from optconpy_b200 import scenarios as sc, dre_stepper as ds, dre_device as dd, device as dv
    prob, cs, kw = sc.config1(glau, Nts=3)
    sg = ds.MemStore()
    fg = ds.solve_flow_daeric(lau=glau, pru=gpru, store=sg, **dict(kw, gtdtstrargs=dict(kw['gtdtstrargs'])))
    ctx = dd.context_from_kwargs(kw)
    setups = dd.prepare_steps(ctx, kw, 3)
    info = []
    for st in setups:
        dd.run_step(ctx, st, info)
    t0 = kw['tmesh'][0]
    assert _relerr(dv.to_host(ctx.mtxtb), sg[fg[t0]['mtxtb']]) < 1e-10
    assert _relerr(dv.to_host(ctx.wc), sg[fg[t0]['w']]) < 1e-10
    assert _zzt_relerr(dv.to_host(ctx.Zc), sg[fg[t0]['mtxtb'].replace('__mtxtb', '__Z')]) < 1e-10


def test_steady_state_branch_parity(mods):
    """Config 3 (cyl_wake_cont.py params on the synthetic channel, coarse): the steady-state
    branch optcont_main.py:488-514 - Newton-ADI from z0=None with the built-in shift list,
    gain and feed-forward - CUDA modules against the oracle."""
    glau, gpru, olau, opru = mods
    from optconpy_b200 import scenarios as sc
    prob, cs, kw = sc.config3(olau, nx=22, ny=8, nu=2e-2)
    kw['nwtn_adi_dict'] = dict(kw['nwtn_adi_dict'], adi_max_steps=120, nwtn_max_steps=8)
    ro = sc.steady_state_feedback(prob, cs, lau=olau, pru=opru, **kw)
    rg = sc.steady_state_feedback(prob, cs, lau=glau, pru=gpru, **kw)
    assert rg['info']['adi_steps'] == ro['info']['adi_steps']
    assert _zzt_relerr(rg['Z'], ro['Z']) < TOL_FACTOR
    assert _relerr(rg['mtxtb'], ro['mtxtb']) < TOL_FACTOR
    assert _relerr(rg['w'], ro['w']) < TOL_TRAJ
    # compressed variant (optcont_main.py:497-500)
    rgc = sc.steady_state_feedback(prob, cs, lau=glau, pru=gpru, compress=(5e-5, 100), **kw)
    roc = sc.steady_state_feedback(prob, cs, lau=olau, pru=opru, compress=(5e-5, 100), **kw)
    assert rgc['Z'].shape == roc['Z'].shape
    assert _relerr(rgc['mtxtb'], roc['mtxtb']) < 1e-7     # truncation at 5e-5 bounds the agreement


def test_lookahead_does_not_change_results(mods):
    """The background LU setup of later steps (dre_stepper look-ahead) is numerically inert."""
    glau, gpru, olau, opru = mods
    from optconpy_b200 import scenarios as sc, dre_stepper as ds
    prob, cs, kw = sc.config1(glau, Nts=3)
    s0, s2 = ds.MemStore(), ds.MemStore()
    f0 = ds.solve_flow_daeric(lau=glau, pru=gpru, store=s0, lookahead=0,
                              **dict(kw, gtdtstrargs=dict(kw['gtdtstrargs'])))
    f2 = ds.solve_flow_daeric(lau=glau, pru=gpru, store=s2, lookahead=2,
                              **dict(kw, gtdtstrargs=dict(kw['gtdtstrargs'])))
    for t in f0:
        assert np.array_equal(s0[f0[t]['mtxtb']], s2[f2[t]['mtxtb']])
        assert np.array_equal(s0[f0[t]['w']], s2[f2[t]['w']])


def test_edge_cases(mods, cav10):
    """Ragged / degenerate inputs through the reference-facing functions."""
    glau, gpru, olau, opru = mods
    from optconpy_b200 import device as dv
    M, A, J = cav10['M'], cav10['A'], cav10['J']
    NV = cav10['NV']
    rng = np.random.default_rng(7)
    # a single right-hand side, given as an (NV,1) column and as a sparse column
    r1 = rng.standard_normal((NV, 1))
    assert _relerr(glau.apply_massinv(M, r1), olau.apply_massinv(M, r1)) < 1e-12
    rs = sps.csr_matrix(r1)
    assert _relerr(glau.apply_massinv(M, rs), olau.apply_massinv(M, rs)) < 1e-12
    # compress: no truncation requested, k larger than the column count, rank-deficient input
    Z = rng.standard_normal((NV, 6))
    Zd = np.hstack([Z, Z[:, :3] @ rng.standard_normal((3, 4))])          # rank 6, 10 columns
    for kw in (dict(), dict(k=50), dict(thresh=1e-8), dict(k=4), dict(k=4, thresh=1e-8)):
        a, b = gpru.compress_Zsvd(Zd, **kw), opru.compress_Zsvd(Zd, **kw)
        kept = min(a.shape[1], b.shape[1])
        if 'thresh' in kw or 'k' in kw and kw['k'] < 6:
            assert a.shape == b.shape, (kw, a.shape, b.shape)
        # the GPU path drops numerically-zero directions even without a threshold
        assert kept >= min(6, kw.get('k', 6))
        assert _zzt_relerr(a, b) < 1e-9
    # one column in, one column out
    z1 = rng.standard_normal((NV, 1))
    a, b = gpru.compress_Zsvd(z1, thresh=1e-10), opru.compress_Zsvd(z1, thresh=1e-10)
    assert a.shape == b.shape == (NV, 1) and _zzt_relerr(a, b) < 1e-12
    # feedback product with a dense 1-column and a sparse multi-column tB
    tbs = sps.random(NV, 3, density=0.05, random_state=3, format='csr')
    assert _relerr(gpru.get_mTzzTtb(M.T, Z, tbs), opru.get_mTzzTtb(M.T, Z, tbs)) < 1e-11
    assert _relerr(gpru.get_mTzzTtb(M.T, Z, r1), opru.get_mTzzTtb(M.T, Z, r1)) < 1e-11
    # zero right-hand sides are legal
    lu = dv.LU(dv.sadpnt_matrix(M + 0.1*A, J))
    X = lu.solve(dv.to_dev(np.zeros((lu.n, 3))))
    assert float(X.abs().max()) == 0.0
    assert lu.solve(dv.to_dev(np.zeros((lu.n, 0)))).shape == (lu.n, 0)


def test_mtxoldb_and_nonconvergence_are_not_errors(mods, lyap_setup, cav10):
    """``mtxoldb`` (outer Newton carry, solve_dae_ric.py:150-155) and hitting ``*_max_steps``
    (non-convergence returns the last iterate)."""
    glau, gpru, olau, opru = mods
    from optconpy_b200 import problems as pb
    M, F, J, _ = lyap_setup
    cs = pb.control_setup(cav10, olau, alphau=1e-4)
    d = dict(adi_max_steps=12, adi_newZ_reltol=1e-14, nwtn_max_steps=2, nwtn_upd_reltol=1e-14,
             nwtn_upd_abstol=1e-16, full_upd_norm_check=False, ms=[-5.0, -2.0, -1.0])
    old = 1e-3*np.random.default_rng(2).standard_normal((cav10['NV'], 8))
    kw = dict(mmat=M.T, amat=F.T, transposed=True, jmat=J, bmat=cs['tb_mat'], wmat=cs['trct_mat'],
              z0=None, mtxoldb=old, nwtn_adi_dict=d)
    ref, got = opru.proj_alg_ric_newtonadi(**kw), gpru.proj_alg_ric_newtonadi(**kw)
    assert got['adi_steps'] == ref['adi_steps'] == [12, 12]
    assert _zzt_relerr(got['zfac'], ref['zfac']) < TOL_FACTOR


def test_lazy_device_factor(mods, lyap_setup, cav10):
    """``_lazy_zfac=True`` (what the DRE driver passes): the factor stays in HBM, has the
    ndarray's shape, converts on demand to exactly the eager result and compresses to the same
    columns; the driver's ``save_full_z`` path stores the converted array."""
    glau, gpru, olau, opru = mods
    from optconpy_b200 import problems as pb, device as dv, scenarios as sc, dre_stepper as ds
    M, F, J, _ = lyap_setup
    cs = pb.control_setup(cav10, olau, alphau=1e-4)
    d = dict(adi_max_steps=60, adi_newZ_reltol=1e-8, nwtn_max_steps=4, nwtn_upd_reltol=1e-8,
             nwtn_upd_abstol=1e-12, ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0])
    kw = dict(mmat=M.T, amat=F.T, transposed=True, jmat=J, bmat=np.sqrt(0.05)*cs['tb_mat'],
              wmat=np.sqrt(0.05)*cs['trct_mat'], z0=None, nwtn_adi_dict=d)
    eager = gpru.proj_alg_ric_newtonadi(**kw)['zfac']
    before = dv.STATS['d2h_bytes']
    lazy = gpru.proj_alg_ric_newtonadi(_lazy_zfac=True, **kw)['zfac']
    assert isinstance(lazy, gpru.DeviceFactor) and lazy.shape == eager.shape and lazy.ndim == 2
    zc_l = gpru.compress_Zsvd(lazy, thresh=5e-5, k=50)
    assert dv.STATS['d2h_bytes'] - before < eager.nbytes        # the big factor never crossed PCIe
    zc_e = gpru.compress_Zsvd(eager, thresh=5e-5, k=50)
    assert zc_l.shape == zc_e.shape and _zzt_relerr(zc_l, zc_e) < 1e-12
    assert _relerr(np.asarray(lazy), eager) < 1e-12 and _relerr(lazy[:, :3], eager[:, :3]) < 1e-12
    prob, cs1, kw1 = sc.config1(glau, Nts=2)
    s1, s2 = ds.MemStore(), ds.MemStore()
    f1 = ds.solve_flow_daeric(lau=glau, pru=gpru, store=s1,
                              **dict(kw1, save_full_z=True, gtdtstrargs=dict(kw1['gtdtstrargs'])))
    f2 = ds.solve_flow_daeric(lau=glau, pru=gpru, store=s2,
                              **dict(kw1, gtdtstrargs=dict(kw1['gtdtstrargs'])))
    for t in f1:
        kz = f1[t]['mtxtb'].replace('__mtxtb', '__Z')
        assert isinstance(s1[kz], np.ndarray) and s1[kz].shape[1] >= s2[kz].shape[1]
        assert _relerr(s1[f1[t]['mtxtb']], s2[f2[t]['mtxtb']]) < 1e-10
