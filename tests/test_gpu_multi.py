"""Multi-GPU parity: column-sharded LR-ADI over 2 GPUs with NCCL.  Needs >= 2 CUDA
devices (``gpurun --gpus 2``); skipped on a single-GPU box.  Every rank iterates on its slice
of the right-hand sides, the stopping test is global (one all-reduced scalar per step), the
factor is re-sharded to row blocks and the Gram matrix is one all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    from optconpy_b200 import problems as pb
    prob = pb.drivcav_problem(10, 1e-2)
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    F = -(0.5*M + 0.05*(A + Nc))
    W = np.random.default_rng(4).standard_normal((prob['NV'], 9))
    d = dict(adi_max_steps=80, adi_newZ_reltol=1e-9, ms=[-5.0, -2.0, -1.0])
    return prob, M, F, J, W, d


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['OCB_LU_WORKERS'] = '0'
    import scipy.sparse as sps
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from optconpy_b200 import device as dv, parallel as par, proj_ric_utils as gpru
    prob, M, F, J, W, d = _case()
    fac = gpru.ShiftedFactors(sps.csr_matrix(F.T), sps.csr_matrix(M.T), J, d['ms'])
    Zl, rel = par.sharded_stein_dev(fac, dv.to_dev(W), d)
    Zall = par.gather_columns_dev(Zl)
    Zrows = par.reshard_columns_to_rows_dev(Zl)
    r0, r1 = par.column_slice(prob['NV'], rank, world)
    MZrows = dv.DeviceCSR(M).matmul(Zall)[r0:r1].contiguous()
    G = par.sharded_gram_dev(Zrows, MZrows)
    # the look-ahead thread and the upload thread must work on THIS rank's device,
    # not on device 0 (nothing here pre-creates the uploader on the main thread)
    import optconpy_b200.lin_alg_utils as glau
    from optconpy_b200 import scenarios as sc, dre_stepper as ds
    os.environ['OCB_LU_WORKERS'] = '2'
    prob1, cs1, kw1 = sc.config1(glau, Nts=2)
    devs = []
    orig = dv.LU.__init__

    def spy(self, *a, **k):
        orig(self, *a, **k)
        devs.append(self.arena.device.index)
    dv.LU.__init__ = spy
    s2 = ds.MemStore()
    f2 = ds.solve_flow_daeric(lau=glau, pru=gpru, store=s2, lookahead=2, **kw1)
    dv.LU.__init__ = orig
    t0 = sorted(f2)[0]
    np.savez(os.path.join(outdir, 'rank%d.npz' % rank), Zl=dv.to_host(Zl), rel=np.array(rel),
             Zall=dv.to_host(Zall), G=dv.to_host(G), lu_devices=np.array(devs),
             gain=s2[f2[t0]['mtxtb']])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_column_sharded_adi_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from oracle import proj_ric_utils as opru
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True,
                       start_method='spawn')
    prob, M, F, J, W, d = _case()
    ref = opru.solve_proj_lyap_stein(amat=F, mmat=M, jmat=J, wmat=W, adi_dict=d)
    outs = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    assert np.array_equal(outs[0]['rel'], outs[1]['rel'])            # same global history
    assert len(outs[0]['rel']) == len(ref['adi_rel_newZ_norms'])     # same iteration count
    assert np.allclose(outs[0]['rel'], ref['adi_rel_newZ_norms'], rtol=1e-5, atol=0)
    assert outs[0]['Zl'].shape[1] + outs[1]['Zl'].shape[1] == ref['zfac'].shape[1]
    Z, Zr = outs[0]['Zall'], ref['zfac']
    assert np.array_equal(outs[0]['Zall'], outs[1]['Zall'])
    R = np.linalg.qr(np.hstack([Z, Zr]), mode='r')
    ka = Z.shape[1]
    D = R[:, :ka] @ R[:, :ka].T - R[:, ka:] @ R[:, ka:].T
    assert np.linalg.norm(D) <= 1e-9*np.linalg.norm(Zr.T @ Zr)
    # every factor image of rank r was uploaded to device r (look-ahead + upload threads), and
    # both ranks computed the same DRE gains
    for r in range(world):
        assert outs[r]['lu_devices'].size > 0 and (outs[r]['lu_devices'] == r).all()
    assert np.allclose(outs[0]['gain'], outs[1]['gain'], rtol=1e-12, atol=0)
    # Gram all-reduce: identical on both ranks, equals Z^T M Z
    assert np.array_equal(outs[0]['G'], outs[1]['G'])
    assert np.allclose(outs[0]['G'], Z.T @ (M @ Z), rtol=1e-11, atol=1e-13)


def _ric_case():
    from optconpy_b200 import problems as pb
    from oracle import lin_alg_utils as olau
    prob = pb.drivcav_problem(10, 1e-2)
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    tau = 0.05
    Ft = -(0.5*M.T + tau*(A.T + Nc.T))
    cs = pb.control_setup(prob, olau, alphau=1e-4)
    d = dict(adi_max_steps=120, adi_newZ_reltol=1e-9, nwtn_max_steps=8, nwtn_upd_reltol=1e-9,
             nwtn_upd_abstol=1e-12, full_upd_norm_check=False, ms=[-5.0, -3.0, -2.0, -1.5, -1.3, -1.1, -1.0])
    kw = dict(mmat=M.T, amat=Ft, transposed=True, jmat=J, bmat=np.sqrt(tau)*cs['tb_mat'],
              wmat=np.sqrt(tau)*cs['trct_mat'], z0=None, nwtn_adi_dict=d)
    return prob, cs, kw


def _worker_public_api(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['OCB_LU_WORKERS'] = '0'
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import optconpy_b200.proj_ric_utils as gpru
    from optconpy_b200 import parallel as par
    prob, cs, kw = _ric_case()
    out = {}
    for name, use_symm in (('auto', None), ('nccl', False)):
        cm = par.enable(use_symm=use_symm)
        res = gpru.proj_alg_ric_newtonadi(_lazy_zfac=True, **kw)          # factor stays sharded
        zc = gpru.compress_Zsvd(res['zfac'], thresh=5e-5, k=50)
        full = np.asarray(res['zfac'])
        gain = -gpru.get_mTzzTtb(prob['M'].T, zc, cs['tb_mat'])
        # Newton restarted from the compressed factor (z0 replicated -> column slices), full check
        d2 = dict(kw['nwtn_adi_dict'], nwtn_max_steps=2, full_upd_norm_check=True)
        res2 = gpru.proj_alg_ric_newtonadi(**dict(kw, z0=zc, nwtn_adi_dict=d2))
        out[name] = dict(transport=cm.transport, symm_error=str(cm.symm_error), zc=zc, full=full, gain=gain,
                         adi_steps=np.array(res['adi_steps']), z2=res2['zfac'],
                         adi_steps2=np.array(res2['adi_steps']), upd2=np.array(res2['nwtn_upd_fnorms']),
                         bytes_p2p=cm.bytes_p2p, bytes_nccl=cm.bytes_nccl)
    par.disable()
    np.savez(os.path.join(outdir, 'api_rank%d.npz' % rank),
             **{'%s_%s' % (n_, k_): v for n_, o in out.items() for k_, v in o.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_newton_and_compress_behind_the_reference_signatures(tmp_path):
    """north_star (d) / SURVEY 8e through the PUBLIC functions: with ``parallel.enable()`` on two
    ranks, ``proj_alg_ric_newtonadi`` and ``compress_Zsvd`` run column-sharded (peer-to-peer
    re-shard + Gram all-reduce over symmetric memory when the box offers it, NCCL otherwise) and
    give the oracle's iteration counts, factor products and gains on every rank."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from oracle import proj_ric_utils as opru
    world = 2
    mp.start_processes(_worker_public_api, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True,
                       start_method='spawn')
    prob, cs, kw = _ric_case()
    ref = opru.proj_alg_ric_newtonadi(**kw)
    zo = opru.compress_Zsvd(ref['zfac'], thresh=5e-5, k=50)
    gain_o = -opru.get_mTzzTtb(prob['M'].T, zo, cs['tb_mat'])
    d2 = dict(kw['nwtn_adi_dict'], nwtn_max_steps=2, full_upd_norm_check=True)
    ref2 = opru.proj_alg_ric_newtonadi(**dict(kw, z0=zo, nwtn_adi_dict=d2))

    def zzt(Za, Zb):
        R = np.linalg.qr(np.hstack([Za, Zb]), mode='r')
        ka = Za.shape[1]
        D = R[:, :ka] @ R[:, :ka].T - R[:, ka:] @ R[:, ka:].T
        return np.linalg.norm(D)/np.linalg.norm(Zb.T @ Zb)
    outs = [np.load(os.path.join(str(tmp_path), 'api_rank%d.npz' % r)) for r in range(world)]
    print('transports:', [str(o['auto_transport']) for o in outs], str(outs[0]['auto_symm_error']))
    for name in ('auto', 'nccl'):
        for o in outs:
            assert list(o[name + '_adi_steps']) == ref['adi_steps']
            assert o[name + '_full'].shape == ref['zfac'].shape and zzt(o[name + '_full'], ref['zfac']) < 1e-9
            assert o[name + '_zc'].shape == zo.shape and zzt(o[name + '_zc'], zo) < 1e-9
            assert np.linalg.norm(o[name + '_gain'] - gain_o) < 1e-9*np.linalg.norm(gain_o)
            assert list(o[name + '_adi_steps2']) == ref2['adi_steps']
            assert zzt(o[name + '_z2'], ref2['zfac']) < 1e-9
        assert np.array_equal(outs[0][name + '_zc'], outs[1][name + '_zc'])      # replicated results
    assert str(outs[0]['nccl_transport']) == 'nccl'
