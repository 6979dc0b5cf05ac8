"""Device-resident backward DRE loop: the same recursion as
``dre_stepper.solve_flow_daeric`` (reference ``solve_dae_ric.py:122-211``) with the
factor ``Zc``, the gain ``mtxtb`` and the feed-forward ``w`` kept in HBM between the
time steps, and the host-side setup of a step (matrix assembly, sparse LU of the
shifted saddle-point matrices, upload) separated from its device work so that the two
can be timed apart, as ``north_star`` asks ("the per-shift LU factorisation is a setup
step and is timed separately").

``prepare_step`` does the setup for one time step; ``run_step`` is the device part:
SpMM ``M^T Zc``, Newton-ADI (multi-RHS SpTRSM + SpMM + SMW + fused update/norm per ADI
iteration), compression, the two feedback products and the single SMW feed-forward solve.
"""
import numpy as np
import scipy.sparse as sps
import torch

from . import device as dv
from . import lin_alg_utils as lau
from . import proj_ric_utils as pru


class DreContext(object):
    """Time-independent data of one DRE solve, uploaded once."""

    def __init__(self, mmat, amat, jmat, bmat, mcmat, rmat, vmat, rhsv, gamma,
                 nwtn_adi_dict, comprz_thresh, comprz_maxc, ystarvec):
        dv.require_cuda()
        self.MT = sps.csr_matrix(mmat.T)
        self.AT = sps.csr_matrix(amat.T)
        self.M = sps.csr_matrix(mmat)
        self.J = sps.csr_matrix(jmat)
        self.NV, self.NP = amat.shape[0], jmat.shape[0]
        self.gamma = gamma
        self.nwtn_adi_dict = dict(nwtn_adi_dict)
        self.shifts = list(self.nwtn_adi_dict.get('ms', pru.DEFAULT_SHIFTS))
        self.thresh, self.maxc = comprz_thresh, comprz_maxc
        self.ystarvec = ystarvec
        self.mcmatT = np.asarray(mcmat.T.todense() if sps.issparse(mcmat) else mcmat.T)
        self.rhsv = np.asarray(rhsv)
        # terminal values (solve_dae_ric.py:92-101,107-108)
        tct = lau.apply_invsqrt_fromright(vmat, mcmat.T, output='dense')
        tb = lau.apply_invsqrt_fromright(rmat, bmat, output='sparse')
        self.tb_host = sps.csr_matrix(tb)
        self.Mt_dev = dv.DeviceCSR(self.MT)
        self.tct = dv.to_dev(tct)
        self.tb = dv.to_dev(tb)                                  # dense NV x m
        self.tbT = dv.DeviceCSR(self.tb_host.T)                  # m x NV, SMW "V" factor
        # size the ADI factor buffer for the widest block the recursion can produce, so that no
        # time step pays for growing it (a 0.5 GB cudaMalloc takes 30-700 ms on a shared box)
        kmax = (comprz_maxc if comprz_maxc is not None else 64) + self.tct.shape[1] + self.tb.shape[1]
        dv.workspace('adi_Z', self.NV*kmax*int(self.nwtn_adi_dict['adi_max_steps'])*8)
        dv.workspace('compress', dv.require_cuda().ocb_compress_ws_bytes(self.NV, kmax*48, 1024))
        mlu = dv.LU(sps.csc_matrix(mmat))
        self.Zc = np.sqrt(gamma)*mlu.solve(self.tct)
        self.mtxtb = dv.feedback(self.Mt_dev, self.Zc, self.tb, alpha=-1.0)
        mtlu = dv.LU(sps.csc_matrix(self.MT))
        self.wc = mtlu.solve(dv.to_dev(gamma*np.dot(self.mcmatT, ystarvec(None))))

    def set_terminal_time(self, tE):
        mtlu = dv.LU(sps.csc_matrix(self.MT))
        self.wc = mtlu.solve(dv.to_dev(self.gamma*np.dot(self.mcmatT, self.ystarvec(tE))))


class StepSetup(object):
    """Host-side setup of one backward step: everything that depends on (t, tau) only."""

    def __init__(self, ctx, t, cts, nmattd, rhsvtd):
        NT = sps.csr_matrix(nmattd.T)
        self.t, self.cts = t, cts
        ft_mat = -(0.5*ctx.MT + cts*(ctx.AT + NT))
        khint = (ctx.maxc if ctx.maxc is not None else ctx.Zc.shape[1]) + ctx.tct.shape[1] + ctx.tb.shape[1]
        self.fac = pru.ShiftedFactors(sps.csr_matrix(ft_mat), ctx.MT, ctx.J, ctx.shifts,
                                      Mt_dev=ctx.Mt_dev, k_hint=khint)
        at_mat = ctx.MT + cts*(ctx.AT + NT)
        self._at_job = dv.FactorJob([dv.sadpnt_matrix(at_mat, ctx.J)]).start_upload()
        self.ftilde = dv.to_dev(np.asarray(rhsvtd) + ctx.rhsv)
        self.fl1 = dv.to_dev(np.dot(ctx.mcmatT, ctx.ystarvec(t)))
        sq = np.sqrt(cts)
        self.Bd = sq*ctx.tb
        self.Vt_b = dv.DeviceCSR(sq*ctx.tb_host.T)
        self.at_lu = None

    def finish(self):
        """Wait for the background factorisations and upload them (still setup, not step)."""
        if self.at_lu is None:
            self.at_lu = self._at_job.result()[0]
            self.fac.lus
        return self


class _Phases(object):
    """Optional CUDA-event phase timer (bench.py --phases): name -> summed milliseconds."""

    def __init__(self):
        self.ev, self.ms = [], {}

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append((name, e))

    def collect(self):
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(self.ev, self.ev[1:]):
            if name != 'start':
                self.ms[name] = self.ms.get(name, 0.0) + a.elapsed_time(b)
        self.ev = []
        return self.ms


def run_step(ctx, st, info=None, phases=None):
    """Device part of one backward step; updates ctx.Zc, ctx.mtxtb, ctx.wc in place."""
    cts = st.cts
    st.finish()
    mark = phases.mark if phases is not None else (lambda name: None)
    mark('start')
    w_mat = torch.cat([ctx.Mt_dev.matmul(ctx.Zc), np.sqrt(cts)*ctx.tct], dim=1).contiguous()
    mark('rhs')
    Zp, ninfo = pru.newtonadi_dev(st.fac, st.Bd, st.Vt_b, w_mat, ctx.Zc, ctx.nwtn_adi_dict)
    mark('newton_adi')
    if ctx.maxc is not None or ctx.thresh is not None:
        Zc, cinfo = dv.compress(Zp, thresh=ctx.thresh, k=ctx.maxc)
        Zc = Zc.contiguous()
    else:
        Zc, cinfo = Zp.contiguous(), {}
    mark('compress')
    # feed-forward (solve_dae_ric.py:173-194); the SMW gain is the PREVIOUS step's mtxtb
    cnsmtxtb = ctx.mtxtb
    mtxft = dv.feedback(ctx.Mt_dev, Zc, st.ftilde)
    rhswc = ctx.Mt_dev.matmul(ctx.wc) + cts*(st.fl1 - mtxft)
    mtxtb = dv.feedback(ctx.Mt_dev, Zc, ctx.tb, alpha=-1.0)
    wc = st.at_lu.smw_solve(rhswc, ctx.NV, Ufb=(cts*cnsmtxtb).contiguous(), Vt=ctx.tbT,
                            nrows_out=ctx.NV)
    ctx.Zc, ctx.mtxtb, ctx.wc = Zc, mtxtb, wc
    mark('feedforward')
    if info is not None:
        info.append(dict(t=st.t, tau=cts, adi_steps=ninfo['adi_steps'],
                         nwtn_upd_fnorms=ninfo['nwtn_upd_fnorms'], zp_cols=Zp.shape[1],
                         zc_cols=Zc.shape[1], solves=sum(ninfo['adi_steps']),
                         chol_rank=cinfo.get('chol_rank')))
    return ctx


def context_from_kwargs(kw):
    """Build a DreContext from the keyword arguments of ``solve_flow_daeric``
    (``scenarios.dre_kwargs``)."""
    tm = kw['tmesh']
    ys = kw['ystarvec']
    ctx = DreContext(kw['mmat'], kw['amat'], kw['jmat'], kw['bmat'], kw['mcmat'], kw['rmat'],
                     kw['vmat'], kw['rhsv'], kw['gamma'], kw['nwtn_adi_dict'],
                     kw['comprz_thresh'], kw['comprz_maxc'],
                     lambda t: ys(tm[-1] if t is None else t))
    return ctx


def prepare_steps(ctx, kw, nsteps):
    """Setup objects for the first ``nsteps`` backward steps from the terminal time."""
    tm = kw['tmesh']
    out = []
    for tk in range(len(tm)-2, len(tm)-2-nsteps, -1):
        t = tm[tk]
        nmattd, rhsvtd = kw['get_tdpart'](time=t, **kw['gttdprtargs'])
        out.append(StepSetup(ctx, t, tm[tk+1]-t, nmattd, rhsvtd))
    for st in out:          # all factorisations were submitted above and run concurrently
        st.finish()
    return out
