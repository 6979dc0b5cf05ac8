"""Column sharding of right-hand-side blocks across ranks.

The LR-ADI is linear and column-wise independent: every column of W generates its own
columns of Z through the same sequence of shifted solves (SMW included), and ``Z Z^T`` does
not depend on the column order.  Rank g owns a contiguous slice of the columns; the only
collectives are (i) two scalars per ADI step for the stopping test, (ii) the k x k Gram
all-reduce of the compression after a re-shard to row blocks, (iii) the NV x m all-reduce
of the feedback product.  This module holds the backend-agnostic plumbing (works with
``gloo`` on CPU and ``nccl`` on GPUs); the kernels are in ``csrc/``.
"""
import numpy as np


def column_slice(k, rank, world):
    """Contiguous, balanced slice [c0, c1) of k columns for ``rank`` of ``world``."""
    base, rem = divmod(int(k), int(world))
    c0 = rank*base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def allreduce_sum(t, group=None):
    """In-place sum over ranks of a torch tensor (k x k Gram, norms, feedback)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def sharded_adi(stein_fn, W, rank, world, allreduce=allreduce_sum, maxsteps=None, reltol=None):
    """Run a column-sharded LR-ADI: ``stein_fn(W_local, stop)`` must generate blocks V_i one
    at a time for the local columns and call ``stop(v_nsq_local)`` after each; ``stop``
    all-reduces the two norms and returns True when the GLOBAL criterion
    ``||V_i||_F / ||Z||_F <= reltol`` holds — every rank stops at the same iteration."""
    import torch
    c0, c1 = column_slice(W.shape[1], rank, world)
    state = dict(z=0.0, rel=[])

    def stop(v_nsq_local):
        t = torch.tensor([float(v_nsq_local)], dtype=torch.float64)
        allreduce(t)
        v = float(t[0])
        state['z'] += v
        rel = np.sqrt(v/state['z']) if state['z'] > 0 else 0.0
        state['rel'].append(rel)
        return not (rel > reltol) or (maxsteps is not None and len(state['rel']) >= maxsteps)
    Zloc = stein_fn(W[:, c0:c1], stop)
    return Zloc, state['rel']


def sharded_gram(Zloc_rows, Wloc_rows, allreduce=allreduce_sum):
    """k x k Gram from ROW-sharded blocks: local partial product + all-reduce."""
    G = Zloc_rows.T @ Wloc_rows
    return allreduce(G)


# ---------------------------------------------------------------------------------------
# device path: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing
# ---------------------------------------------------------------------------------------
def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def sharded_stein_dev(fac, W, adi_dict, Ufb=None, Vt=None, group=None):
    """Column-sharded LR-ADI on the GPUs.  ``fac`` (``proj_ric_utils.ShiftedFactors``) and the
    low-rank factors ``Ufb`` / ``Vt`` are replicated on every rank, ``W`` (device, NV x k) is
    the full right-hand-side block; rank g iterates on its column slice only.  The ONLY
    communication is the all-reduce of one scalar per ADI step for the global stopping test
    ``||V_i||_F / ||Z||_F <= adi_newZ_reltol``, so every rank stops at the same step.
    Returns (local factor block NV x (steps * k_local), global relative norms)."""
    import torch
    import torch.distributed as dist
    from . import device as dv
    rank, world = _world(group)
    c0, c1 = column_slice(W.shape[1], rank, world)
    if c1 <= c0:
        raise ValueError('fewer right-hand-side columns than ranks')
    Wl = W[:, c0:c1].contiguous()
    buf = torch.zeros(1, dtype=torch.float64, device=W.device)

    def reduce_norm(v):
        if world == 1:
            return v
        buf[0] = v
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        return float(buf.item())
    return dv.adi_run(fac.lus, fac.ms, fac.NV, fac.NP, fac.Mt_dev, Wl,
                      int(adi_dict['adi_max_steps']), float(adi_dict['adi_newZ_reltol']),
                      Ufb=Ufb, Vt=Vt, norm_reduce=reduce_norm)


def gather_columns_dev(Zl, group=None):
    """All ranks' column blocks side by side (``Z Z^T`` does not depend on the column order)."""
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return Zl
    widths = [torch.zeros(1, dtype=torch.int64, device=Zl.device) for _ in range(world)]
    dist.all_gather(widths, torch.tensor([Zl.shape[1]], dtype=torch.int64, device=Zl.device), group=group)
    parts = [torch.empty((Zl.shape[0], int(w.item())), dtype=Zl.dtype, device=Zl.device) for w in widths]
    dist.all_gather(parts, Zl.contiguous(), group=group)
    return torch.cat(parts, dim=1).contiguous()


def reshard_columns_to_rows_dev(Zl, group=None):
    """Column blocks -> row blocks: rank g ends up with rows ``column_slice(NV, g, world)`` of
    ALL columns (the re-shard of SURVEY 8e before the Gram all-reduce; point-to-point traffic
    over NVLink, n*K*8 bytes in total)."""
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return Zl
    NV = Zl.shape[0]
    widths = [torch.zeros(1, dtype=torch.int64, device=Zl.device) for _ in range(world)]
    dist.all_gather(widths, torch.tensor([Zl.shape[1]], dtype=torch.int64, device=Zl.device), group=group)
    widths = [int(w.item()) for w in widths]
    r0, r1 = column_slice(NV, rank, world)
    send = [Zl[slice(*column_slice(NV, g, world)), :].contiguous() for g in range(world)]
    recv = [torch.empty((r1 - r0, widths[g]), dtype=Zl.dtype, device=Zl.device) for g in range(world)]
    dist.all_to_all(recv, send, group=group)
    return torch.cat(recv, dim=1).contiguous()


def sharded_gram_dev(Zrows, Wrows, group=None):
    """K x K Gram matrix from ROW-sharded blocks: local partial product on the FP64 tensor pipe
    (``ocb_gram``, DMMA) + one all-reduce - the only bulk collective of the scheme."""
    import torch.distributed as dist
    from . import device as dv
    G = dv.gram(Zrows, Wrows)
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    return G
