"""CPU suite, part 4: the N>1 path with world_size 2 over ``gloo``.

Each rank owns a contiguous slice of the right-hand-side columns, runs the LR-ADI on its
slice with the GLOBAL stopping test (two all-reduced scalars per step), then the factor is
re-sharded to row blocks for the k x k Gram all-reduce.  Checked against the single-process
oracle: same iteration count, same relative-norm history, same Z Z^T and Gram matrix."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    from optconpy_b200 import problems as pb
    prob = pb.drivcav_problem(6, 1e-2)
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    F = -(0.5*M + 0.05*(A + Nc))
    W = np.random.default_rng(4).standard_normal((prob['NV'], 5))
    ms = [-5.0, -2.0, -1.0]
    return prob, M, F, J, W, ms


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import scipy.sparse as sps
    import torch
    import torch.distributed as dist
    from oracle import lin_alg_utils as olau
    from optconpy_b200 import parallel as par
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    prob, M, F, J, W, ms = _case()
    NV, NP = prob['NV'], prob['NP']
    Ft, Mt = sps.csr_matrix(F.T), sps.csr_matrix(M.T)
    lus = [olau.SadLU(olau.sadpnt_matrix(Ft + mu*Mt, J)) for mu in ms]

    def solve(i, R):
        return lus[i](np.vstack([R, np.zeros((NP, R.shape[1]))]))[:NV]

    def stein_fn(Wloc, stop):
        V = np.sqrt(-2*ms[0])*solve(0, Wloc)
        blocks = [V]
        step = 1
        while not stop(np.linalg.norm(V)**2):
            i, ip = step % len(ms), (step-1) % len(ms)
            V = np.sqrt(ms[i]/ms[ip])*(V - (ms[i]+ms[ip])*solve(i, np.asarray(Mt @ V)))
            blocks.append(V)
            step += 1
        return np.hstack(blocks)

    Zloc, rel = par.sharded_adi(stein_fn, W, rank, world, maxsteps=80, reltol=1e-9)
    # re-shard: column blocks -> row blocks (all_gather stands in for the NVLink all-to-all)
    parts = [None]*world
    dist.all_gather_object(parts, Zloc)
    Zall = np.hstack(parts)
    r0, r1 = par.column_slice(NV, rank, world)            # same balanced split, over rows
    MZ = np.asarray(M @ Zall)
    G = par.sharded_gram(torch.from_numpy(Zall[r0:r1].copy()), torch.from_numpy(MZ[r0:r1].copy()))
    np.savez(os.path.join(outdir, 'rank%d.npz' % rank), Zloc=Zloc, rel=np.array(rel), G=G.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_column_sharded_adi_world2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import proj_ric_utils as opru
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world,
                       join=True, start_method='spawn')
    prob, M, F, J, W, ms = _case()
    ref = opru.solve_proj_lyap_stein(amat=F, mmat=M, jmat=J, wmat=W,
                                     adi_dict=dict(adi_max_steps=80, adi_newZ_reltol=1e-9, ms=ms))
    outs = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    # every rank stopped at the same, global, iteration and saw the same norm history
    assert np.array_equal(outs[0]['rel'], outs[1]['rel'])
    assert len(outs[0]['rel']) == len(ref['adi_rel_newZ_norms'])
    assert np.allclose(outs[0]['rel'], ref['adi_rel_newZ_norms'], rtol=1e-5, atol=0)
    Z = np.hstack([o['Zloc'] for o in outs])
    Zr = ref['zfac']
    assert Z.shape == Zr.shape
    assert np.linalg.norm(Z @ Z.T - Zr @ Zr.T) <= 1e-12*np.linalg.norm(Zr @ Zr.T)
    # Gram all-reduce: identical on both ranks, equals Z^T M Z
    assert np.array_equal(outs[0]['G'], outs[1]['G'])
    assert np.allclose(outs[0]['G'], Z.T @ (M @ Z), rtol=1e-12, atol=1e-14)
