"""Tracking cost along the closed-loop trajectory (north_star: "the DRE
trajectory of the tracking cost matches within 1e-8").

CPU part: the restated cost functional (``optcont_main.py:213-264``) and the closed-loop
simulator behave (the feedback lowers the cost).  GPU part: the cost computed from the CUDA
modules' gains equals the one from the oracle's gains to 1e-8 relative."""
import numpy as np
import pytest

from oracle import lin_alg_utils as olau, proj_ric_utils as opru
from optconpy_b200 import scenarios as sc, dre_stepper as ds, tracking as tr

TOL_COST = 1e-8


def _case(N, Nts):
    """config 1 parameters (optcon_nse defaults) on the mesh-N cavity."""
    from optconpy_b200 import problems as pb
    if N == 10:
        return sc.config1(olau, Nts=Nts)
    prob = pb.drivcav_problem(N, 1e-2)
    cs = pb.control_setup(prob, olau, alphau=1e-9)
    tmesh = pb.get_tint(0.0, 1.0, Nts)
    return prob, cs, sc.dre_kwargs(prob, cs, tmesh, dict(sc.DEFAULT_NWTN_ADI, adi_max_steps=120),
                                   1e-3, sc._ystar_sin(cs['NY']))


def _closed_loop_cost(lau, pru, Nts=4, feedback=True, N=10):
    prob, cs, kw = _case(N, Nts)                      # operators from the oracle: same inputs
    kw = dict(kw, gtdtstrargs=dict(kw['gtdtstrargs']))
    store = ds.MemStore()
    fb = ds.solve_flow_daeric(lau=lau, pru=pru, store=store, **kw) if feedback else None
    tb = olau.apply_invsqrt_fromright(cs['R'], cs['b_mat'], output='sparse')
    vel = tr.simulate_closed_loop(mmat=prob['M'], amat=prob['A'], jmat=prob['J'], rhsv=prob['fv'],
                                  tb_mat=tb, tmesh=kw['tmesh'], get_tdpart=kw['get_tdpart'],
                                  fbftdict=fb, store=store)
    cost = tr.eval_costfunc(W=cs['y_masmat'], V=kw['gamma']*cs['y_masmat'], R=None, tbmat=tb,
                            cmat=cs['c_mat'], ystar=kw['ystarvec'], tmesh=kw['tmesh'],
                            veldict=vel, fbftdict=fb, store=store, penau=feedback)
    return cost, store, vel


def test_costfunc_is_the_trapezoidal_rule():
    """eval_costfunc on a hand-made trajectory."""
    import scipy.sparse as sps
    store = ds.MemStore()
    tmesh = np.array([0.0, 0.5, 2.0])
    vel = {}
    for t in tmesh:
        store.save(np.array([[t], [2*t]]), 'v%g' % t)
        vel[t] = 'v%g' % t
    C = sps.csr_matrix(np.array([[1.0, 1.0]]))
    Wm = sps.csr_matrix(np.array([[2.0]]))
    ystar = lambda t: np.array([[1.0]])
    g = lambda t: 2.0*(1.0 - 3*t)**2
    want = 0.5*0.5*(g(0.0)+g(0.5)) + 0.5*1.5*(g(0.5)+g(2.0)) + 5.0*(1 - 6.0)**2
    got = tr.eval_costfunc(W=Wm, V=sps.csr_matrix(np.array([[5.0]])), cmat=C, ystar=ystar,
                           tmesh=tmesh, veldict=vel, fbftdict=None, store=store, penau=False)
    assert np.isclose(got, want, rtol=1e-14)


def test_feedback_lowers_the_tracking_cost():
    c_fb, store, vel = _closed_loop_cost(olau, opru, Nts=8, feedback=True, N=6)
    c_no, _, _ = _closed_loop_cost(olau, opru, Nts=8, feedback=False, N=6)
    assert np.isfinite(c_fb) and c_fb > 0
    assert c_fb < c_no                  # total cost (tracking + control) below the uncontrolled one
    # the simulated velocities stay discretely divergence free
    prob = _case(6, 8)[0]
    for t in vel:
        assert np.linalg.norm(prob['J'] @ store[vel[t]]) < 1e-10


def test_sigout_json_roundtrip(tmp_path):
    """``__sigout`` format of optcont_main.py:160-182 (read by plot_output.py)."""
    c_fb, store, vel = _closed_loop_cost(olau, opru, Nts=3, feedback=False, N=6)
    prob, cs, kw = _case(6, 3)
    ys, ystars = tr.extract_output(dictofpaths=vel, tmesh=kw['tmesh'], c_mat=cs['c_mat'],
                                   ystarvec=kw['ystarvec'], store=store)
    assert len(ys) == cs['c_mat'].shape[0] and len(ys[0]) == len(kw['tmesh'])
    f = tr.save_output_json(ys, kw['tmesh'].tolist(), ystar=ystars, fstring=str(tmp_path / 'x__sigout'))
    js = tr.load_json_dicts(f)
    assert sorted(js) == ['tmesh', 'ycomp', 'ystar']
    assert np.allclose(js['ycomp'], ys) and np.allclose(js['tmesh'], kw['tmesh'])


@pytest.mark.gpu
def test_tracking_cost_parity_gpu_vs_oracle():
    import optconpy_b200.lin_alg_utils as glau
    import optconpy_b200.proj_ric_utils as gpru
    c_o, so, vo = _closed_loop_cost(olau, opru, Nts=4)
    c_g, sg, vg = _closed_loop_cost(glau, gpru, Nts=4)
    assert abs(c_g - c_o) <= TOL_COST*abs(c_o)
    for t in vo:                                        # and the closed-loop states themselves
        assert np.linalg.norm(sg[vg[t]] - so[vo[t]]) <= TOL_COST*max(np.linalg.norm(so[vo[t]]), 1e-30)
