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
