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
