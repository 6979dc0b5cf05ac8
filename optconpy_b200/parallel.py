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


def sharded_stein_local(fac, Wl, adi_dict, Ufb=None, Vt=None, group=None):
    """As ``sharded_stein_dev`` for a caller that already holds only ITS column slice ``Wl`` (blocks
    too wide to replicate): same global stopping test, one all-reduced scalar per ADI step."""
    import torch
    import torch.distributed as dist
    from . import device as dv
    rank, world = _world(group)
    buf = torch.zeros(1, dtype=torch.float64, device=Wl.device)

    def reduce_norm(v):
        if world == 1:
            return v
        buf[0] = v
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        return float(buf.item())
    return dv.adi_run(fac.lus, fac.ms, fac.NV, fac.NP, fac.Mt_dev, Wl.contiguous(),
                      int(adi_dict['adi_max_steps']), float(adi_dict['adi_newZ_reltol']),
                      Ufb=Ufb, Vt=Vt, norm_reduce=reduce_norm)


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


# ---------------------------------------------------------------------------------------
# The column-sharded hot path behind the reference signatures (north_star (d)):
#   * LR-ADI: rank g iterates on its slice of the right-hand-side columns; per ADI step ONE
#     all-reduced scalar (global stopping test);
#   * Newton-Kleinman: the feedback product Z Z^T B and the update probe are sums over the
#     column blocks -> NV x m all-reduces; the factor itself never moves;
#   * compression: column blocks -> row blocks (peer-to-peer stores into symmetric memory),
#     local partial Gram product on the FP64 tensor pipe, the K x K all-reduce as its epilogue
#     (peer-to-peer loads, fixed order), replicated Cholesky / eigen core, local Z_rows T, rows
#     gathered by peer-to-peer stores.  Without symmetric memory the same steps run over NCCL.
# ---------------------------------------------------------------------------------------
_COMM = dict(comm=None)


class ShardComm(object):
    """Ranks of one process group (one process per GPU) + the collectives of the sharded path."""

    def __init__(self, group=None, use_symm=None):
        import os
        import torch
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = _world(group)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._bufs = {}
        self.transport = 'nccl'
        self.symm_error = None
        self.bytes_p2p = 0          # bytes this rank moved through peer memory
        self.bytes_nccl = 0         # bytes this rank handed to NCCL collectives
        if self.world > 1 and use_symm is not False and not os.environ.get('OCB_NO_SYMM_MEM'):
            ok = 1
            try:
                self._buf('probe', 4096)
            except Exception as exc:      # no CUDA VMM / fabric handle exchange on this box
                self.symm_error = repr(exc)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.transport = 'p2p'
            else:
                self._bufs.clear()

    # -- symmetric memory -----------------------------------------------------------------
    def _buf(self, name, nbytes):
        """Symmetric buffer ``name`` of at least ``nbytes`` (collective when it has to grow: all
        ranks call with the same sizes in the same order).  Returns (uint8 tensor, handle)."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        ent = self._bufs.get(name)
        if ent is None or ent[0].numel() < nbytes:
            cap = int(nbytes*1.25) + 4096
            t = symm_mem.empty(cap, dtype=torch.uint8, device=self.device)
            grp = self.group if self.group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(t, grp)
            ent = (t, hdl)
            self._bufs[name] = ent
        return ent

    # -- small collectives (NCCL) -----------------------------------------------------------
    def allreduce_(self, t):
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.bytes_nccl += t.numel()*t.element_size()
        return t

    def widths(self, kl):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return [int(kl)]
        w = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        w[self.rank] = int(kl)
        dist.all_reduce(w, op=dist.ReduceOp.SUM, group=self.group)
        return [int(v) for v in w.tolist()]

    # -- bulk movement ------------------------------------------------------------------------
    def reshard_columns_to_rows(self, Zl, widths):
        """Column block of this rank (NV x widths[rank]) -> row block (rows of this rank x K)."""
        import torch
        from . import device as dv
        NV = Zl.shape[0]
        K = int(sum(widths))
        offs = np.concatenate([[0], np.cumsum(widths)]).astype(np.int64)
        r0, r1 = column_slice(NV, self.rank, self.world)
        if self.world == 1:
            return Zl
        if self.transport != 'p2p':
            import torch.distributed as dist
            send = [Zl[slice(*column_slice(NV, g, self.world)), :].contiguous() for g in range(self.world)]
            recv = [torch.empty((r1 - r0, widths[g]), dtype=Zl.dtype, device=Zl.device) for g in range(self.world)]
            dist.all_to_all(recv, send, group=self.group)
            self.bytes_nccl += Zl.numel()*8
            return torch.cat(recv, dim=1).contiguous()
        rmax = -(-NV//self.world)
        buf, hdl = self._buf('rows', rmax*K*8)
        hdl.barrier(channel=0)                    # peers have finished with the previous contents
        for g in range(self.world):
            g0, g1 = column_slice(NV, g, self.world)
            if g1 > g0 and Zl.shape[1] > 0:
                dv.p2p_put2d(Zl[g0:g1], int(hdl.buffer_ptrs[g]) + int(offs[self.rank])*8, K)
        self.bytes_p2p += Zl.numel()*8
        hdl.barrier(channel=0)                    # everybody's stores have landed
        return buf[:(r1 - r0)*K*8].view(torch.float64).view(r1 - r0, K)

    def gram_allreduce(self, Zrows, Wrows):
        """K x K Gram matrix from ROW-sharded blocks: local partial product on the FP64 tensor
        pipe, summed over the ranks in its epilogue (peer loads in fixed rank order: bitwise
        identical on every rank)."""
        import torch
        from . import device as dv
        ka, kb = Zrows.shape[1], Wrows.shape[1]
        if self.world == 1:
            return dv.gram(Zrows, Wrows)
        if self.transport != 'p2p':
            import torch.distributed as dist
            G = dv.gram(Zrows, Wrows)
            dist.all_reduce(G, op=dist.ReduceOp.SUM, group=self.group)
            self.bytes_nccl += G.numel()*8
            return G
        buf, hdl = self._buf('gram', ka*kb*8)
        part = buf[:ka*kb*8].view(torch.float64).view(ka, kb)
        hdl.barrier(channel=1)                    # nobody still reads the previous partials
        dv.gram(Zrows, Wrows, out=part)
        hdl.barrier(channel=1)
        G = torch.empty((ka, kb), dtype=torch.float64, device=self.device)
        dv.p2p_sum_peers([int(p) for p in hdl.buffer_ptrs], ka*kb, G)
        self.bytes_p2p += (self.world - 1)*ka*kb*8
        return G

    def allgather_rows(self, Xrows, NV):
        """Row blocks (rows of rank g x kc) -> the full NV x kc block on every rank."""
        import torch
        from . import device as dv
        kc = Xrows.shape[1]
        if self.world == 1:
            return Xrows
        if self.transport != 'p2p':
            import torch.distributed as dist
            parts = [torch.empty((column_slice(NV, g, self.world)[1] - column_slice(NV, g, self.world)[0], kc),
                                 dtype=Xrows.dtype, device=Xrows.device) for g in range(self.world)]
            dist.all_gather(parts, Xrows.contiguous(), group=self.group)
            self.bytes_nccl += Xrows.numel()*8*(self.world - 1)
            return torch.cat(parts, dim=0).contiguous()
        buf, hdl = self._buf('gather', NV*max(kc, 1)*8)
        r0, r1 = column_slice(NV, self.rank, self.world)
        hdl.barrier(channel=2)
        if kc > 0 and r1 > r0:
            for g in range(self.world):
                dv.p2p_put2d(Xrows, int(hdl.buffer_ptrs[g]) + r0*kc*8, kc)
        self.bytes_p2p += Xrows.numel()*8*self.world
        hdl.barrier(channel=2)
        return buf[:NV*kc*8].view(torch.float64).view(NV, kc).clone()

    def allgather_columns(self, Zl, widths):
        """Column blocks side by side on every rank (only for callers that ask for the full
        uncompressed factor as an ndarray)."""
        import torch
        from . import device as dv
        if self.world == 1:
            return Zl
        NV, K = Zl.shape[0], int(sum(widths))
        offs = np.concatenate([[0], np.cumsum(widths)]).astype(np.int64)
        if self.transport != 'p2p':
            import torch.distributed as dist
            parts = [torch.empty((NV, w), dtype=Zl.dtype, device=Zl.device) for w in widths]
            dist.all_gather(parts, Zl.contiguous(), group=self.group)
            return torch.cat(parts, dim=1).contiguous()
        buf, hdl = self._buf('cols', NV*K*8)
        hdl.barrier(channel=3)
        if Zl.shape[1] > 0:
            for g in range(self.world):
                dv.p2p_put2d(Zl, int(hdl.buffer_ptrs[g]) + int(offs[self.rank])*8, K)
        self.bytes_p2p += Zl.numel()*8*self.world
        hdl.barrier(channel=3)
        return buf[:NV*K*8].view(torch.float64).view(NV, K).clone()


def enable(group=None, use_symm=None):
    """Switch the column-sharded path on for this process (call on every rank after
    ``torch.distributed.init_process_group``; one process per GPU).  The module functions of
    ``proj_ric_utils`` keep their signatures; they must then be called on all ranks with the
    same arguments (SPMD), and every rank gets the same results."""
    _COMM['comm'] = ShardComm(group, use_symm=use_symm)
    return _COMM['comm']


def disable():
    _COMM['comm'] = None


def comm():
    c = _COMM['comm']
    return c if (c is not None and c.world > 1) else None


class ShardedFactor(object):
    """A low-rank factor whose COLUMN blocks live on the ranks of a ShardComm; turns into the
    full ndarray on demand (all-gather), ``compress_Zsvd`` takes it as is."""

    def __init__(self, comm_, local, widths):
        self.comm, self.local, self.widths = comm_, local, list(widths)
        self._host = None

    shape = property(lambda self: (int(self.local.shape[0]), int(sum(self.widths))))
    dtype = property(lambda self: np.dtype(np.float64))
    ndim = property(lambda self: 2)

    def __len__(self):
        return int(self.local.shape[0])

    def __array__(self, dtype=None, copy=None):
        from . import device as dv
        if self._host is None:
            self._host = dv.to_host(self.comm.allgather_columns(self.local, self.widths))
        a = self._host
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __getitem__(self, idx):
        return self.__array__()[idx]


def sharded_stein(comm_, fac, W, adi_dict, Ufb=None, Vt=None):
    """Column-sharded LR-ADI (W: the full block, replicated).  Blocks narrower than the number
    of ranks are iterated on redundantly by everybody and owned by rank 0.
    Returns (local column block, widths of all ranks, relative norms)."""
    from . import device as dv
    k = W.shape[1]
    if k < comm_.world:
        Z, rel = dv.adi_run(fac.lus, fac.ms, fac.NV, fac.NP, fac.Mt_dev, W, int(adi_dict['adi_max_steps']),
                            float(adi_dict['adi_newZ_reltol']), Ufb=Ufb, Vt=Vt)
        Zl = Z if comm_.rank == 0 else Z[:, :0]
        return Zl, [Z.shape[1]] + [0]*(comm_.world - 1), rel
    Zl, rel = sharded_stein_dev(fac, W, adi_dict, Ufb=Ufb, Vt=Vt, group=comm_.group)
    return Zl, comm_.widths(Zl.shape[1]), rel


def sharded_newtonadi(comm_, fac, Bd, Vt_b, W, z0, nwtn_adi_dict, mtxoldb=None, probe=None):
    """Newton-Kleinman with the factor column-sharded over the ranks (``proj_ric_utils.
    newtonadi_dev`` is the one-GPU version).  ``z0``: full initial factor (replicated) or None.
    Returns (local block, widths, info)."""
    import torch
    from . import device as dv
    NV = W.shape[0]
    if z0 is not None:
        c0, c1 = column_slice(z0.shape[1], comm_.rank, comm_.world)
        znc = z0[:, c0:c1].contiguous()
    else:
        znc = None
    fnorms, adi_steps, rels = [], [], []
    maxstp = int(nwtn_adi_dict['nwtn_max_steps'])
    reltol = nwtn_adi_dict.get('nwtn_upd_reltol', 0.0)
    abstol = nwtn_adi_dict.get('nwtn_upd_abstol', 0.0)
    full = nwtn_adi_dict.get('full_upd_norm_check', False)
    fro = lambda t: float(torch.sqrt((t*t).sum()).item())
    widths = None
    stp = 0
    while stp < maxstp:
        if znc is None:
            rhsadi, kfb = W, None
        else:
            kfb = comm_.allreduce_(dv.feedback(fac.Mt_dev, znc, Bd))          # M^T Z Z^T B
            rhsadi = torch.cat([kfb, W], dim=1).contiguous()
        if mtxoldb is not None:
            kfb = -mtxoldb if kfb is None else kfb - mtxoldb
        znn, widths, rel = sharded_stein(comm_, fac, rhsadi, nwtn_adi_dict, Ufb=kfb,
                                         Vt=Vt_b if kfb is not None else None)
        adi_steps.append(len(rel))
        rels.append(rel)
        if full:
            # ||Z Z^T||_F etc. need the cross terms of all column blocks: row re-shard + Gram
            rows_n = comm_.reshard_columns_to_rows(znn, widths).clone()
            ref = fro(comm_.gram_allreduce(rows_n, rows_n))
            if znc is None:
                upd = ref
            else:
                wc_ = comm_.widths(znc.shape[1])
                rows_c = comm_.reshard_columns_to_rows(znc, wc_).clone()
                gnc, gcc = comm_.gram_allreduce(rows_n, rows_c), comm_.gram_allreduce(rows_c, rows_c)
                upd = np.sqrt(abs(ref**2 - 2*fro(gnc)**2 + fro(gcc)**2))
        else:
            vec = probe
            nv = comm_.allreduce_(dv.tall_gemm(znn, dv.gram(znn, vec)) if znn.shape[1] else
                                  torch.zeros((NV, 1), dtype=torch.float64, device=W.device))
            ref = fro(nv)
            if znc is None:
                upd = ref
            else:
                cv = comm_.allreduce_(dv.tall_gemm(znc, dv.gram(znc, vec)) if znc.shape[1] else
                                      torch.zeros((NV, 1), dtype=torch.float64, device=W.device))
                upd = fro(nv - cv)
        fnorms.append(upd)
        znc = znn
        stp += 1
        if upd < abstol or upd < reltol*ref:
            break
    return znc, widths, dict(nwtn_upd_fnorms=fnorms, adi_steps=adi_steps, adi_rel_norms=rels)


def sharded_compress(comm_, Zl, widths, thresh=None, k=None):
    """``compress_Zsvd`` of a column-sharded factor: re-shard to rows, partial Gram + all-reduce,
    replicated core, local ``Z_rows T``, rows gathered.  Returns (full Zc on every rank, info)."""
    from . import device as dv
    NV = Zl.shape[0]
    rows = comm_.reshard_columns_to_rows(Zl, widths)
    G = comm_.gram_allreduce(rows, rows)
    T, info = dv.compress_from_gram(G, thresh=thresh, k=k)
    zc_rows = dv.tall_gemm(rows.contiguous(), T.contiguous()) if T.shape[1] else rows[:, :0].contiguous()
    return comm_.allgather_rows(zc_rows, NV), info


def shared_factors(comm_, mats, wide=False, k_hint=None, lu_options=None):
    """Shift-sharded SETUP (north_star (d): "sharding of RHS columns and ADI shifts"): the host
    factorisations of ``mats`` (the shifted saddle-point matrices of one ADI, replicated input)
    are dealt to the ranks - rank r factorises mats[r::world] in its worker processes - and the
    finished device images are handed to the other ranks of the node through POSIX shared
    memory; every rank uploads all of them (the solves need every shift on every GPU).
    Returns the list of ``device.LU`` handles in the order of ``mats``."""
    import torch.distributed as dist
    from multiprocessing import shared_memory
    from . import device as dv, _lu_worker
    world, rank = (1, 0) if comm_ is None else (comm_.world, comm_.rank)
    if world == 1:
        return dv.FactorJob(mats, lu_options=lu_options, wide=wide, k_hint=k_hint).result()
    opts = dict(dv.LU_OPTIONS if lu_options is None else lu_options)
    so = dv.smem_optin()
    mine = list(range(rank, len(mats), world))
    pool = dv._lu_pool()
    jobs = []
    for i in mine:
        a, key = dv._with_order(dv._csc_args(mats[i], opts) + (so, dv._pack_flags(wide, k_hint)), opts)
        jobs.append((i, pool.apply_async(_lu_worker.factor_image_to_shm, (a, None)) if pool is not None
                     else _lu_worker.factor_image_to_shm(a, None)))
    local = []
    for i, j in jobs:
        name, nbytes, tf, tp, order, guard = j.get() if hasattr(j, 'get') else j
        dv.STATS['lu_factor_s'] += tf
        dv.STATS['lu_worker_pack_s'] += tp
        dv.STATS['n_factor'] += 1
        dv._record_guard(guard)
        local.append((i, name, nbytes))
    everything = [None]*world
    dist.all_gather_object(everything, local, group=comm_.group)
    lus = [None]*len(mats)
    for owner, lst in enumerate(everything):
        for i, name, nbytes in lst:
            shm = shared_memory.SharedMemory(name=name)
            try:
                img = np.frombuffer(shm.buf, dtype=np.uint8, count=nbytes)
                lus[i] = dv.LU(None, image=img)
                del img
            finally:
                shm.close()
    dist.barrier(group=comm_.group)          # everybody has uploaded: the owners release the segments
    for i, name, nbytes in local:
        try:
            seg = shared_memory.SharedMemory(name=name)
            seg.close()
            seg.unlink()
        except FileNotFoundError:
            pass
    return lus
