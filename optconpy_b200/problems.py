"""Seeded synthetic Oseen-type saddle-point problems with the reference's sizes.

The reference gets its matrices from FEniCS through ``dolfin_navier_scipy``
(``optcont_main.py:322-334``: ``dts.get_stokessysmats`` +
``dts.condense_sysmatsbybcs``; ``optcont_main.py:373-395``: the input/output
operators from ``distr_control_fenics``).  None of that is available here, so
this module assembles the same *kind* of matrices with a self-written
P2-P1 Taylor-Hood assembler on the structured right-diagonal triangulation
dolfin's ``UnitSquareMesh`` uses.  Sizes equal the reference's
(``NV = 2(2N-1)^2``, ``NP = (N+1)^2 - 1``; N=20 -> NV=3042 as recorded in
``debugstuff.py:30``).

Nothing here is on the hot path: it is input synthesis for tests and
``bench.py`` (numpy/scipy on the host, run once per problem).
"""
import numpy as np
import scipy.sparse as sps

__all__ = ['drivcav_problem', 'channel_problem', 'get_tint', 'convection_matrix',
           'analytic_vortex', 'control_setup', 'oseen_sweep_problem']

# 7-point degree-5 rule on the reference triangle (barycentric points, weights sum to 1)
_a1, _b1 = 0.059715871789770, 0.470142064105115
_a2, _b2 = 0.797426985353087, 0.101286507323456
_QP = np.array([[1/3, 1/3, 1/3],
                [_a1, _b1, _b1], [_b1, _a1, _b1], [_b1, _b1, _a1],
                [_a2, _b2, _b2], [_b2, _a2, _b2], [_b2, _b2, _a2]])
_QW = np.array([0.225] + [0.132394152788506]*3 + [0.125939180544827]*3)


def _p2_shape(lam):
    """P2 basis values (nq,6) and d/dlambda (nq,6,3) at barycentric points lam (nq,3).

    local numbering: 0,1,2 vertices; 3,4,5 midpoints of edges (0,1),(1,2),(2,0)."""
    l0, l1, l2 = lam[:, 0], lam[:, 1], lam[:, 2]
    phi = np.stack([l0*(2*l0-1), l1*(2*l1-1), l2*(2*l2-1),
                    4*l0*l1, 4*l1*l2, 4*l2*l0], axis=1)
    z = np.zeros_like(l0)
    dphi = np.stack([
        np.stack([4*l0-1, z, z], 1), np.stack([z, 4*l1-1, z], 1),
        np.stack([z, z, 4*l2-1], 1), np.stack([4*l1, 4*l0, z], 1),
        np.stack([z, 4*l2, 4*l1], 1), np.stack([4*l2, z, 4*l0], 1)], axis=1)
    return phi, dphi


class _Mesh(object):
    """Structured (nx x ny cells) right-diagonal triangulation of [0,lx]x[0,ly],
    optionally with a mask of removed cells (an obstacle)."""

    def __init__(self, nx, ny, lx=1.0, ly=1.0, cellmask=None):
        self.nx, self.ny, self.lx, self.ly = nx, ny, lx, ly
        w = 2*nx + 1                      # P2 lattice width
        self.w = w
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
        ii, jj = ii.ravel(), jj.ravel()
        if cellmask is not None:
            keep = ~cellmask[ii, jj]
            ii, jj = ii[keep], jj[keep]

        def nid(ix, iy):
            return iy*w + ix
        x0, y0 = 2*ii, 2*jj
        # triangle 1: (v00, v10, v11); triangle 2: (v00, v11, v01)
        t1 = np.stack([nid(x0, y0), nid(x0+2, y0), nid(x0+2, y0+2),
                       nid(x0+1, y0), nid(x0+2, y0+1), nid(x0+1, y0+1)], 1)
        t2 = np.stack([nid(x0, y0), nid(x0+2, y0+2), nid(x0, y0+2),
                       nid(x0+1, y0+1), nid(x0+1, y0+2), nid(x0, y0+1)], 1)
        self.tri = np.vstack([t1, t2])                   # (nel, 6) lattice ids
        nlat = w*(2*ny+1)
        ix = np.arange(nlat) % w
        iy = np.arange(nlat) // w
        self.xy = np.stack([ix*(lx/(2*nx)), iy*(ly/(2*ny))], 1)
        self.ix, self.iy = ix, iy
        self.active = np.zeros(nlat, dtype=bool)
        self.active[self.tri.ravel()] = True
        # boundary detection: an edge midpoint that belongs to one element only
        mids = self.tri[:, 3:].ravel()
        cnt = np.bincount(mids, minlength=nlat)
        bmid = np.where(cnt == 1)[0]
        self.bnd = np.zeros(nlat, dtype=bool)
        self.bnd[bmid] = True
        # the two vertices adjacent to a boundary midpoint
        tv = self.tri
        for (a, b, m) in ((0, 1, 3), (1, 2, 4), (2, 0, 5)):
            sel = self.bnd[tv[:, m]]
            self.bnd[tv[sel, a]] = True
            self.bnd[tv[sel, b]] = True

    def geometry(self):
        X = self.xy[self.tri[:, :3]]                       # (nel,3,2)
        d1, d2 = X[:, 1]-X[:, 0], X[:, 2]-X[:, 0]
        det = d1[:, 0]*d2[:, 1] - d1[:, 1]*d2[:, 0]       # 2*area (>0)
        # gradients of barycentric coordinates
        g1 = np.stack([d2[:, 1], -d2[:, 0]], 1)/det[:, None]
        g2 = np.stack([-d1[:, 1], d1[:, 0]], 1)/det[:, None]
        g0 = -g1 - g2
        glam = np.stack([g0, g1, g2], 1)                  # (nel,3,2)
        return X, det, glam


def _assemble(rows, cols, vals, shape):
    return sps.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())),
                          shape=shape).tocsr()


def _stokes_mats(mesh, nu):
    """Scalar P2 mass / stiffness and the P1xP2 divergence blocks on lattice ids."""
    X, det, glam = mesh.geometry()
    nel = det.size
    phi, dphi = _p2_shape(_QP)                            # (nq,6), (nq,6,3)
    gphi = np.einsum('qil,eld->eqid', dphi, glam)         # (nel,nq,6,2)
    wq = 0.5*det[:, None]*_QW[None, :]                    # (nel,nq)
    Me = np.einsum('eq,qi,qj->eij', wq, phi, phi)
    Ke = np.einsum('eq,eqid,eqjd->eij', wq, gphi, gphi)
    psi = _QP                                             # P1 basis = barycentric
    # De[e, p, j, d] = int psi_p d_d phi_j
    De = np.einsum('eq,qp,eqjd->epjd', wq, psi, gphi)
    nlat = mesh.xy.shape[0]
    t = mesh.tri
    r6 = np.repeat(t[:, :, None], 6, 2)
    c6 = np.repeat(t[:, None, :], 6, 1)
    Ms = _assemble(r6, c6, Me, (nlat, nlat))
    Ks = _assemble(r6, c6, Ke, (nlat, nlat))
    r3 = np.repeat(t[:, :3, None], 6, 2)
    c3 = np.repeat(t[:, None, :], 3, 1)
    Dx = _assemble(r3, c3, De[..., 0], (nlat, nlat))
    Dy = _assemble(r3, c3, De[..., 1], (nlat, nlat))
    return Ms, nu*Ks, Dx, Dy


def convection_matrix(prob, vfun, newton_term=True):
    """Oseen convection matrix for the linearisation point ``vfun(xy)->(n,2)``.

    ``N1[i,j] = int (v.grad phi_j) phi_i`` per component, plus (if
    ``newton_term``) ``N2[(c,i),(d,j)] = int phi_j d_d v_c phi_i``; this is the
    synthetic stand-in for ``snu.get_v_conv_conts`` (``optcont_main.py:556-568``).
    Returned on the inner (non-Dirichlet) velocity dofs, shape (NV, NV).

    The time stepper calls this once per time step (``get_tdpart``), so everything that depends
    on the mesh only - shape-function tables, and the map from the element entries of the four
    velocity blocks to the entries of the condensed CSR matrix - is computed once per problem;
    a call then costs the element integrals and one ``bincount``.  The sparsity pattern is the
    same for every linearisation point (entries that happen to cancel stay as stored zeros)."""
    cache = prob.setdefault('_conv_cache', {})
    c = cache.get(bool(newton_term))
    if c is None:
        mesh = prob['mesh']
        X, det, glam = mesh.geometry()
        phi, dphi = _p2_shape(_QP)
        nlat = mesh.xy.shape[0]
        t = mesh.tri
        inv = np.asarray(prob['invinds'])
        pos = np.full(2*nlat, -1, dtype=np.int64)
        pos[inv] = np.arange(inv.size)
        NV = inv.size
        blocks = [(0, 0), (0, 1), (1, 0), (1, 1)] if newton_term else [(0, 0), (1, 1)]
        r6 = np.repeat(t[:, :, None], 6, 2)
        c6 = np.repeat(t[:, None, :], 6, 1)
        keys = []
        for (cc, dd) in blocks:                     # interleaved dof = 2*node + component
            rr, cl = pos[2*r6 + cc], pos[2*c6 + dd]
            keys.append(np.where((rr >= 0) & (cl >= 0), rr*NV + cl, -1).ravel())
        keys = np.concatenate(keys)
        keep = np.flatnonzero(keys >= 0)
        uniq, inverse = np.unique(keys[keep], return_inverse=True)
        rows = uniq // NV
        c = dict(phi=phi, gphi=np.einsum('qil,eld->eqid', dphi, glam),
                 wq=0.5*det[:, None]*_QW[None, :], blocks=blocks, keep=keep, inverse=inverse,
                 nnz=uniq.size, NV=NV, indices=(uniq % NV).astype(np.int32),
                 indptr=np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=NV))]).astype(np.int32))
        cache[bool(newton_term)] = c
    mesh = prob['mesh']
    phi, gphi, wq = c['phi'], c['gphi'], c['wq']
    vnod = vfun(mesh.xy)                                   # (nlat,2) P2 interpolant
    vel = vnod[mesh.tri]                                   # (nel,6,2)
    vq = np.einsum('qi,eic->eqc', phi, vel)                # (nel,nq,2)
    N1e = np.einsum('eq,qi,eqd,eqjd->eij', wq, phi, vq, gphi, optimize=True)
    if newton_term:
        gv = np.einsum('eqid,eic->eqcd', gphi, vel)        # d_d v_c
        wpp = np.einsum('eq,qi,qj->eqij', wq, phi, phi)
        vals = []
        for (cc, dd) in c['blocks']:
            E = np.einsum('eqij,eq->eij', wpp, gv[:, :, cc, dd])
            vals.append((E + N1e if cc == dd else E).ravel())
    else:
        vals = [N1e.ravel(), N1e.ravel()]
    vals = np.concatenate(vals)
    data = np.bincount(c['inverse'], weights=vals[c['keep']], minlength=c['nnz'])
    out = sps.csr_matrix((data, c['indices'].copy(), c['indptr'].copy()), shape=(c['NV'], c['NV']))
    out.has_sorted_indices = True
    return out


def _interleave_blocks(blocks, nlat):
    """2x2 block matrix on (comp, node) -> interleaved dof = 2*node + comp."""
    B = sps.bmat(blocks, format='coo')
    r = 2*(B.row % nlat) + B.row // nlat
    c = 2*(B.col % nlat) + B.col // nlat
    return sps.coo_matrix((B.data, (r, c)), shape=(2*nlat, 2*nlat)).tocsr()


def _finish(mesh, nu, dirimask, bcvals, name):
    """Condense by the Dirichlet dofs and drop the last pressure row
    (``optcont_main.py:327-334``)."""
    Ms, As, Dx, Dy = _stokes_mats(mesh, nu)
    nlat = mesh.xy.shape[0]
    M = _interleave_blocks([[Ms, None], [None, Ms]], nlat)
    A = _interleave_blocks([[As, None], [None, As]], nlat)
    Dfull = sps.hstack([Dx, Dy], format='coo')            # (nlat, 2 nlat) comp-major cols
    cc = 2*(Dfull.col % nlat) + Dfull.col // nlat
    J = sps.coo_matrix((Dfull.data, (Dfull.row, cc)), shape=(nlat, 2*nlat)).tocsr()
    # pressure nodes: active even-even lattice points
    pn = np.where(mesh.active & (mesh.ix % 2 == 0) & (mesh.iy % 2 == 0))[0]
    J = J[pn, :]
    J = J[:-1, :]                                          # remove the pressure freedom
    act2 = np.repeat(mesh.active, 2)
    dir2 = np.asarray(dirimask).ravel() & act2             # interleaved (node, comp)
    invinds = np.where(act2 & ~dir2)[0]
    bcinds = np.where(dir2)[0]
    bcv = np.asarray(bcvals).ravel()[bcinds][:, None]
    Mc = M[invinds, :][:, invinds].tocsr()
    Ac = A[invinds, :][:, invinds].tocsr()
    Jc = J[:, invinds].tocsr()
    fv_bc = -(A[invinds, :][:, bcinds] @ bcv)
    fp_bc = -(J[:, bcinds] @ bcv)
    Mc.sum_duplicates(); Ac.sum_duplicates(); Jc.sum_duplicates()
    Jc.eliminate_zeros()
    return dict(name=name, mesh=mesh, nu=nu, M=Mc, A=Ac, J=Jc, JT=Jc.T.tocsr(),
                fv=np.asarray(fv_bc), fp=np.asarray(fp_bc),
                invinds=invinds, bcinds=bcinds, bcvals=bcv,
                NV=Mc.shape[0], NP=Jc.shape[0])


def drivcav_problem(N=10, nu=1e-2):
    """Driven cavity on the unit square, lid velocity (1,0) at y=1
    (stand-in for ``dnsps.drivcav_fems(N)``, ``optcont_main.py:287-289``)."""
    mesh = _Mesh(N, N)
    nlat = mesh.xy.shape[0]
    dirimask = np.zeros((nlat, 2), dtype=bool)
    dirimask[mesh.bnd, :] = True
    bcvals = np.zeros((nlat, 2))
    lid = mesh.bnd & (mesh.iy == 2*N)
    bcvals[lid, 0] = 1.0
    prob = _finish(mesh, nu, dirimask, bcvals, 'drivencavity')
    prob.update(N=N, uspacedep=0,
                cdcoo=dict(xmin=0.4, xmax=0.6, ymin=0.2, ymax=0.3),
                odcoo=dict(xmin=0.45, xmax=0.55, ymin=0.5, ymax=0.7))
    return prob


def channel_problem(nx=44, ny=16, nu=2.5e-3, lx=2.2, ly=0.4):
    """Channel with a square obstacle: the cylinder-wake-*type* configuration
    (stand-in for ``dnsps.cyl_fems``; parabolic inflow at x=0, no-slip walls and
    obstacle, do-nothing outflow at x=lx)."""
    hx, hy = lx/nx, ly/ny
    cx = (np.arange(nx)+0.5)*hx
    cy = (np.arange(ny)+0.5)*hy
    cellmask = (np.abs(cx[:, None]-0.2) < 0.05) & (np.abs(cy[None, :]-0.2) < 0.05)
    mesh = _Mesh(nx, ny, lx, ly, cellmask=cellmask)
    nlat = mesh.xy.shape[0]
    outflow = mesh.bnd & (mesh.ix == 2*nx) & (mesh.iy > 0) & (mesh.iy < 2*ny)
    dirimask = np.zeros((nlat, 2), dtype=bool)
    dirimask[mesh.bnd & ~outflow, :] = True
    bcvals = np.zeros((nlat, 2))
    inflow = mesh.bnd & (mesh.ix == 0)
    y = mesh.xy[:, 1]
    bcvals[inflow, 0] = 4.0*y[inflow]*(ly-y[inflow])/ly**2
    prob = _finish(mesh, nu, dirimask, bcvals, 'cylinderwake')
    prob.update(N=(nx, ny), uspacedep=1,
                cdcoo=dict(xmin=0.27, xmax=0.32, ymin=0.15, ymax=0.25),
                odcoo=dict(xmin=0.6, xmax=0.7, ymin=0.15, ymax=0.25))
    return prob


def oseen_sweep_problem(N, nu=1e-2, conv_scale=1.0):
    """Synthetic Oseen sweep member (BASELINE config 5): cavity mesh N with the
    analytic vortex as linearisation point.  Returns prob with key 'Nconv'."""
    prob = drivcav_problem(N, nu)
    prob['Nconv'] = convection_matrix(prob, lambda xy: conv_scale*analytic_vortex(xy))
    return prob


def analytic_vortex(xy, t=0.0):
    """Smooth divergence-free field v=(sin(pi x)cos(pi y), -cos(pi x)sin(pi y)),
    mildly modulated in time (SURVEY 8(d) synthetic inputs)."""
    x, y = np.pi*xy[:, 0], np.pi*xy[:, 1]
    amp = 1.0 + 0.25*np.sin(2*np.pi*t)
    return amp*np.stack([np.sin(x)*np.cos(y), -np.cos(x)*np.sin(y)], 1)


def get_tint(t0, tE, Nts, sqzmesh=True):
    """The squeezed time mesh of ``optcont_main.py:141-151``."""
    if sqzmesh:
        taux = np.linspace(-0.5*np.pi, 0.5*np.pi, int(Nts)+1)
        taux = (np.sin(taux) + 1)*0.5
        return (t0 + (tE-t0)*taux).flatten()
    return np.linspace(t0, tE, int(Nts)+1).flatten()


def _hat_1d(s, n):
    """values (npts, n) of the n P1 hat functions on [0,1] at s."""
    pos = np.clip(s, 0.0, 1.0)*(n-1)
    out = np.zeros((s.size, n))
    for l in range(n):
        out[:, l] = np.clip(1.0-np.abs(pos-l), 0.0, None)
    return out


def _mass_1d(n, length):
    h = length/(n-1)
    main = np.full(n, 2*h/3.0)
    main[0] = main[-1] = h/3.0
    off = np.full(n-1, h/6.0)
    return sps.diags([off, main, off], [-1, 0, 1], format='csr')


def _domain_operator(prob, dcoo, nfun, vardir):
    """G[(comp,l), dof] = int_{domain} theta_l(xi) phi_dof^comp dx  (2 nfun x NVfull)."""
    mesh = prob['mesh']
    X, det, glam = mesh.geometry()
    phi, _ = _p2_shape(_QP)
    xq = np.einsum('ql,eld->eqd', _QP, X)                  # (nel,nq,2)
    inside = ((xq[..., 0] >= dcoo['xmin']) & (xq[..., 0] <= dcoo['xmax']) &
              (xq[..., 1] >= dcoo['ymin']) & (xq[..., 1] <= dcoo['ymax']))
    lo = dcoo['ymin'] if vardir == 1 else dcoo['xmin']
    hi = dcoo['ymax'] if vardir == 1 else dcoo['xmax']
    s = (xq[..., vardir]-lo)/(hi-lo)
    th = _hat_1d(s.ravel(), nfun).reshape(s.shape + (nfun,))
    wq = 0.5*det[:, None]*_QW[None, :]*inside
    Ge = np.einsum('eq,eql,qi->eli', wq, th, phi)           # (nel,nfun,6)
    nlat = mesh.xy.shape[0]
    rows = np.repeat(np.arange(nfun)[None, :, None], Ge.shape[0], 0)
    rows = np.repeat(rows, 6, 2)
    cols = np.repeat(mesh.tri[:, None, :], nfun, 1)
    Gs = _assemble(rows, cols, Ge, (nfun, nlat)).tocoo()    # scalar
    r = np.concatenate([Gs.row, Gs.row+nfun])
    c = np.concatenate([2*Gs.col, 2*Gs.col+1])
    d = np.concatenate([Gs.data, Gs.data])
    return sps.coo_matrix((d, (r, c)), shape=(2*nfun, 2*nlat)).tocsr()


def control_setup(prob, lau, NU=4, NY=4, alphau=1e-9, ystar_none_x=False):
    """Input / output operators and their regularised forms, following
    ``optcont_main.py:373-425`` (cou.get_inp_opa / get_mout_opa, projection of
    C^T, B R^-1/2, C^T My^-1/2).  ``lau`` is the lin_alg_utils module to use
    (the oracle's in CPU tests, the CUDA one in GPU runs)."""
    inv = prob['invinds']
    cd, od = prob['cdcoo'], prob['odcoo']
    xcomp = prob['uspacedep']
    b_full = _domain_operator(prob, cd, NU, xcomp).T.tocsr()    # (2nlat, 2NU)
    ulen = (cd['ymax']-cd['ymin']) if xcomp == 1 else (cd['xmax']-cd['xmin'])
    u_masmat = sps.block_diag([_mass_1d(NU, ulen)]*2, format='csr')
    mc_full = _domain_operator(prob, od, NY, 1)                 # (2NY, 2nlat)
    y_masmat = sps.block_diag([_mass_1d(NY, od['ymax']-od['ymin'])]*2, format='csr')
    mc_mat = mc_full[:, inv][:, :].tocsr()
    b_mat = b_full[inv, :][:, :].tocsr()
    c_mat = lau.apply_massinv(y_masmat, mc_mat, output='sparse')
    if ystar_none_x:            # optcont_main.py:400-403
        c_mat = c_mat[NY:, :][:, :]
        mc_mat = mc_mat[NY:, :][:, :]
        y_masmat = y_masmat[:NY, :][:, :NY]
    mct_mat_reg = lau.app_prj_via_sadpnt(amat=prob['M'], jmat=prob['J'],
                                         rhsv=mc_mat.T, transposedprj=True)
    R = alphau*u_masmat
    tb_mat = lau.apply_invsqrt_fromright(R, b_mat, output='sparse')
    trct_mat = lau.apply_invsqrt_fromright(y_masmat, mct_mat_reg, output='dense')
    return dict(b_mat=b_mat, u_masmat=u_masmat, mc_mat=mc_mat, y_masmat=y_masmat,
                c_mat=c_mat, mct_mat_reg=mct_mat_reg, R=R, tb_mat=tb_mat,
                trct_mat=trct_mat, NU=NU, NY=NY, alphau=alphau)
