"""Parameter sets of the reference's scenario scripts as synthetic workloads.

Each ``*_config`` returns the keyword arguments of the reference's drivers with
the FEniCS parts replaced by :mod:`optconpy_b200.problems`:

* ``config1``: ``optcon_nse`` defaults (``optcont_main.py:267-279``: N=10, Nts=10,
  nu=1e-2, alphau=1e-9, gamma=1e-3, t in [0,1] as in ``:677-678``), time-dependent
  branch, default ``nwtn_adi_dict`` (``:122-131``), compression 5e-5 / 50 (``:132-134``).
* ``config2``: ``run_optcont.py:12-41`` (N=25, Nts=128, tE=0.2, nu=5e-3,
  alphau=1e-7, gamma=1e-1, 7 shifts, y* = -/+0.1 sin(5*3.14 t)).
* ``config2b``: ``driv_cav_cont.py:8-30`` (N=25, Nts=40, nu=1e-2, alphau=1e-4, k<=60).
* ``config3``: ``cyl_wake_cont.py:8-28`` steady-state branch on the synthetic channel.
* ``config4``: the same on the fine 200 x 62 channel mesh (NV ~ 1e5), columns sharded over GPUs.
"""
import numpy as np

from . import problems as pb

DEFAULT_NWTN_ADI = dict(adi_max_steps=200, adi_newZ_reltol=1e-8, nwtn_max_steps=16,
                        nwtn_upd_reltol=5e-8, nwtn_upd_abstol=1e-7, verbose=False,
                        full_upd_norm_check=False, check_lyap_res=False)


def _ystar_sin(NY):
    def ystarvec(t):
        a = 0.1*np.sin(5*3.14*t)
        return np.vstack([np.full((NY, 1), -a), np.full((NY, 1), a)])
    return ystarvec


def _ystar_zero(NY, rows=2):
    def ystarvec(t):
        return np.zeros((rows*NY, 1))
    return ystarvec


def dre_kwargs(prob, cs, tmesh, nwtn_adi_dict, gamma, ystarvec,
               comprz_thresh=5e-5, comprz_maxc=50, conv_scale=1.0, stokes=False):
    """Keyword arguments for ``solve_flow_daeric`` as ``optcont_main.py:584-600``
    passes them; the time-dependent part is the analytic-vortex Oseen matrix
    (stand-in for the forward simulation ``snu.solve_nse``, ``:548-568``)."""
    import scipy.sparse as sps
    NV = prob['NV']
    cache = {}

    def get_tdpart(time=None, **kw):
        if stokes:
            return sps.csr_matrix((NV, NV)), np.zeros((NV, 1))
        if time not in cache:
            cache[time] = pb.convection_matrix(
                prob, lambda xy: conv_scale*pb.analytic_vortex(xy, time))
        return cache[time], np.zeros((NV, 1))

    return dict(mmat=prob['M'], amat=prob['A'], jmat=prob['J'], bmat=cs['b_mat'],
                mcmat=cs['mct_mat_reg'].T, v_is_my=True, rmat=cs['R'],
                vmat=cs['y_masmat'], rhsv=prob['fv'], gamma=gamma, rhsp=None,
                tmesh=tmesh, ystarvec=ystarvec, nwtn_adi_dict=nwtn_adi_dict,
                comprz_thresh=comprz_thresh, comprz_maxc=comprz_maxc,
                save_full_z=False, get_tdpart=get_tdpart, gttdprtargs={},
                gtdtstrargs=dict(meshp=prob['N'], nu=prob['nu'], Nts=len(tmesh)-1,
                                 data_prfx=''))


def config1(lau, Nts=10):
    prob = pb.drivcav_problem(10, 1e-2)
    cs = pb.control_setup(prob, lau, alphau=1e-9)
    tmesh = pb.get_tint(0.0, 1.0, Nts)
    return prob, cs, dre_kwargs(prob, cs, tmesh, dict(DEFAULT_NWTN_ADI), 1e-3,
                                _ystar_sin(cs['NY']))


def config2(lau, N=25, Nts=128, tE=0.2):
    nwtn_adi_dict = dict(adi_max_steps=300, adi_newZ_reltol=1e-7, nwtn_max_steps=20,
                         nwtn_upd_reltol=4e-8, nwtn_upd_abstol=1e-7, verbose=False,
                         ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0],
                         full_upd_norm_check=False, check_lyap_res=False)
    prob = pb.drivcav_problem(N, 0.5e-2)
    cs = pb.control_setup(prob, lau, alphau=1e-7)
    tmesh = pb.get_tint(0.0, tE, Nts)
    return prob, cs, dre_kwargs(prob, cs, tmesh, nwtn_adi_dict, 1e-1,
                                _ystar_sin(cs['NY']))


def config2b(lau, N=25, Nts=40):
    prob = pb.drivcav_problem(N, 1e-2)
    cs = pb.control_setup(prob, lau, alphau=1e-4)
    tmesh = pb.get_tint(0.0, 1.0, Nts)
    return prob, cs, dre_kwargs(prob, cs, tmesh, dict(DEFAULT_NWTN_ADI), 1e-3,
                                _ystar_zero(cs['NY']), comprz_maxc=60)


def config3(lau, nx=44, ny=16, nu=2.5e-3):
    """Steady-state branch (``optcont_main.py:451-514``) on the channel problem;
    returns the pieces the branch needs."""
    nwtn_adi_dict = dict(adi_max_steps=199, adi_newZ_reltol=1e-8, nwtn_max_steps=16,
                         nwtn_upd_reltol=5e-8, nwtn_upd_abstol=1e-7, verbose=False,
                         full_upd_norm_check=False, check_lyap_res=False)
    prob = pb.channel_problem(nx, ny, nu)
    cs = pb.control_setup(prob, lau, alphau=1e-4, ystar_none_x=True)
    lx, ly = prob['mesh'].lx, prob['mesh'].ly

    def base_flow(xy):
        y = xy[:, 1]
        return np.stack([4.0*y*(ly-y)/ly**2, np.zeros_like(y)], 1)
    convc = pb.convection_matrix(prob, base_flow)
    return prob, cs, dict(convc_mat=convc, nwtn_adi_dict=nwtn_adi_dict,
                          ystarvec=_ystar_zero(cs['NY'], rows=1))


def config4(lau, nx=200, ny=62, nu=2.5e-3):
    """BASELINE config[3]: "cylinder wake fine mesh (~1e5 velocity dofs)" - the parameters of
    ``cyl_wake_cont.py:8-28`` (config 3) on the 200 x 62 channel mesh: NV = 97 146, NP = 12 542,
    n = 109 688.  Steady-state branch ``optcont_main.py:451-514``."""
    return config3(lau, nx=nx, ny=ny, nu=nu)


def steady_state_feedback(prob, cs, convc_mat, nwtn_adi_dict, ystarvec, lau, pru,
                          zini=None, compress=None):
    """The steady-state branch ``optcont_main.py:488-514``: Newton-ADI for the
    projected ARE, optional compression, gain and feed-forward."""
    M, A, J = prob['M'], prob['A'], prob['J']
    NV = prob['NV']
    res = pru.proj_alg_ric_newtonadi(mmat=M, amat=-A-convc_mat, jmat=J,
                                     bmat=cs['tb_mat'], wmat=cs['trct_mat'], z0=zini,
                                     nwtn_adi_dict=nwtn_adi_dict)
    Z = res['zfac']
    if compress is not None:
        Z = pru.compress_Zsvd(Z, thresh=compress[0], k=compress[1])
    fv = prob['fv']
    mtxtb = -pru.get_mTzzTtb(M.T, Z, cs['tb_mat'])
    mtxfv = -pru.get_mTzzTtb(M.T, Z, fv)
    fl = cs['mc_mat'].T @ ystarvec(0)
    wft = lau.solve_sadpnt_smw(amat=A.T+convc_mat.T, jmat=J, rhsv=fl+mtxfv,
                               umat=mtxtb, vmat=cs['tb_mat'].T)[:NV]
    return dict(Z=Z, mtxtb=mtxtb, w=wft, info=res)
