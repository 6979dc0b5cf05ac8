"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

There are no upstream golden vectors (Fact 3): these pin the ORACLE's own
output on seeded inputs so that regressions of the oracle, the generator or the CUDA
path are caught.  Run:  python tools/gen_golden.py
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from optconpy_b200 import problems as pb, scenarios as sc, dre_stepper as ds
from oracle import lin_alg_utils as olau, proj_ric_utils as opru


def lyap_case():
    prob = pb.drivcav_problem(6, 1e-2)
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    F = -(0.5*M + 0.05*(A + Nc))
    W = np.random.default_rng(0).standard_normal((prob['NV'], 3))
    d = dict(adi_max_steps=60, adi_newZ_reltol=1e-9, ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0])
    return prob, F, W, d


def main():
    out = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out, exist_ok=True)
    prob, F, W, d = lyap_case()
    res = opru.solve_proj_lyap_stein(amat=F, mmat=prob['M'], jmat=prob['J'], wmat=W, adi_dict=d)
    Z = res['zfac']
    Zc = opru.compress_Zsvd(Z, thresh=1e-6)
    G = Zc.T @ (prob['M'] @ Zc)
    np.savez_compressed(os.path.join(out, 'lyap_cav6.npz'),
                        rel_norms=np.array(res['adi_rel_newZ_norms']),
                        zcols=Z.shape[1], zc_cols=Zc.shape[1],
                        sv=np.linalg.svd(Z, compute_uv=False)[:20],
                        gram_m_eigs=np.sort(np.linalg.eigvalsh(G))[::-1][:20],
                        matrix_sums=np.array([prob['M'].sum(), prob['A'].sum(), abs(prob['J']).sum()]))
    # three backward DRE steps of config 1 geometry on the N=6 mesh
    prob6 = pb.drivcav_problem(6, 1e-2)
    cs = pb.control_setup(prob6, olau, alphau=1e-9)
    tmesh = pb.get_tint(0.0, 1.0, 3)
    kw = sc.dre_kwargs(prob6, cs, tmesh, dict(sc.DEFAULT_NWTN_ADI, adi_max_steps=120), 1e-3,
                       sc._ystar_sin(cs['NY']))
    store, info = ds.MemStore(), []
    fb = ds.solve_flow_daeric(lau=olau, pru=opru, store=store, stepinfo=info, **kw)
    ts = sorted(fb)
    np.savez_compressed(os.path.join(out, 'dre_cav6.npz'), tmesh=np.array(ts),
                        mtxtb=np.stack([store[fb[t]['mtxtb']] for t in ts]),
                        w=np.stack([store[fb[t]['w']] for t in ts]),
                        adi_steps=np.array([sum(i['adi_steps']) for i in info]),
                        zc_cols=np.array([i['zc_cols'] for i in info]))
    print('wrote', os.listdir(out))


if __name__ == '__main__':
    main()
