"""CPU suite: numeric-only refactorisation with symbolic reuse (csrc/refactor.cpp).

The reference factorises ``A + p_j M`` from scratch for every shift and every time step
(``proj_ric_utils.py:108-111`` inside ``solve_dae_ric.py:147-163``); all of these matrices share
one sparsity pattern.  ``ocb_refactor_*`` keeps the pivot order of a first SuperLU run and redoes
the numbers only.  Checked here without a GPU: ``L U = P A Q`` to rounding, solves against
SuperLU, fixed structure, unsymmetric patterns, the zero-pivot and the guard fall-backs of the
worker, and the image of the static-pivot path through the gather program."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

from optconpy_b200 import _cabi, _lu_worker, device as dv, problems as pb
from test_lu_program import _program, _execute


@pytest.fixture(scope='module')
def cav10():
    return pb.drivcav_problem(10, 1e-2)


def _shifted(prob, tau, p):
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    return dv.sadpnt_matrix(-(0.5*M.T + tau*(A.T + Nc.T)) + p*M.T, J)


def _csc(K):
    K = sps.csc_matrix(K, dtype=np.float64)
    K.sum_duplicates()
    K.sort_indices()
    return K


def _lu_of(rf, K):
    n = K.shape[0]
    arrs = rf.numeric(K.data)
    assert arrs is not None
    L = sps.csr_matrix((arrs[2].copy(), arrs[1], arrs[0]), shape=(n, n))
    U = sps.csr_matrix((arrs[5].copy(), arrs[4], arrs[3]), shape=(n, n))
    return L, U, arrs[6], arrs[7]


def _permuted(K, pr, pc):
    co = K.tocoo()
    return sps.csr_matrix((co.data, (pr[co.row], pc[co.col])), shape=K.shape)


def test_static_pivots_reproduce_the_matrix_and_superlu(cav10):
    """First matrix through SuperLU (ordering + pivots), the other shifts / step lengths through
    the numeric-only pass: L U = P A Q, unit lower L, solves equal to SuperLU's."""
    K0 = _csc(_shifted(cav10, 2e-3, -1.0))
    n = K0.shape[0]
    a = dv._csc_args(K0, dict(dv.LU_OPTIONS)) + (232448, 2)
    q = _lu_worker.order_only(a)
    arrs = _lu_worker.factor_arrays(a + (q,), transposed=True)
    rf = _lu_worker._Refactor(n, K0.indptr, K0.indices, arrs[6], arrs[7])
    assert rf.info['nnzL'] == rf.info['nnzU']                      # symmetric structure
    # no more fill than SuperLU's own factors of this pivot order (+ a few accidental zeros it drops)
    assert rf.info['nnzL'] + rf.info['nnzU'] <= 1.1*(len(arrs[1]) + len(arrs[4]))
    rng = np.random.default_rng(0)
    for tau, p in ((2e-3, -1.0), (3e-4, -5.0), (1e-3, -1.3), (5e-2, -2.0)):
        K = _csc(_shifted(cav10, tau, p))
        assert np.array_equal(K.indices, K0.indices) and np.array_equal(K.indptr, K0.indptr)
        L, U, pr, pc = _lu_of(rf, K)
        assert np.array_equal(L.diagonal(), np.ones(n))
        assert sps.triu(L, 1).nnz == 0 and sps.tril(U, -1).nnz == 0
        Cm = _permuted(K, pr, pc)
        assert abs(L @ U - Cm).max() <= 1e-12*abs(Cm).max()
        b = rng.standard_normal(n)
        y = np.empty(n)
        y[pr] = b
        z = spsla.spsolve_triangular(U, spsla.spsolve_triangular(L, y, lower=True, unit_diagonal=True),
                                     lower=False)
        x = z[pc]
        ref = spsla.splu(K).solve(b)
        assert np.linalg.norm(x - ref) <= 1e-11*np.linalg.norm(ref)
        assert np.linalg.norm(K @ x - b) <= 1e-12*np.linalg.norm(b)


@pytest.mark.parametrize('n,density,seed', [(1, 1.0, 0), (7, 0.4, 1), (300, 0.02, 2), (900, 0.006, 3)])
def test_unsymmetric_pattern_identity_pivots(n, density, seed):
    """General sparse pattern (not symmetric), diagonally dominant values, identity permutations:
    the structure is symmetrised internally, the product must still be exact; large supernodes
    (dense trailing fronts) exercise the blocked kernels."""
    rng = np.random.default_rng(seed)
    S = sps.random(n, n, density=density, format='csc', random_state=seed)
    S.data[:] = rng.standard_normal(S.nnz)
    K = _csc(S + sps.identity(n)*(1.0 + abs(S).sum(axis=1).max()))
    ident = np.arange(n, dtype=np.int32)
    rf = _lu_worker._Refactor(n, K.indptr, K.indices, ident, ident)
    L, U, pr, pc = _lu_of(rf, K)
    assert sorted(pr) == list(range(n)) and np.array_equal(pr, pc)      # identity composed with a postorder
    Cm = _permuted(K, pr, pc)
    assert abs(L @ U - Cm).max() <= 1e-13*abs(Cm).max()
    assert rf.info['supernodes'] >= 1 and rf.info['max_front'] <= n


def test_dense_block_goes_through_the_blocked_kernels():
    """One 150-column supernode with a 40-row border: five 32-pivot blocks, the inverse-based panel
    updates, the in-place row panel and the register-blocked trailing update."""
    rng = np.random.default_rng(5)
    n = 190
    D = rng.standard_normal((n, n))
    D[150:, 150:] = 0.0
    D += np.diag(np.full(n, 60.0))
    K = _csc(sps.csc_matrix(D))
    ident = np.arange(n, dtype=np.int32)
    rf = _lu_worker._Refactor(n, K.indptr, K.indices, ident, ident)
    assert rf.info['max_front'] == n
    L, U, pr, pc = _lu_of(rf, K)
    Cm = _permuted(K, pr, pc)
    assert abs(L @ U - Cm).max() <= 1e-13*abs(Cm).max()


def test_bad_arguments_and_zero_pivot():
    lib = _cabi.load()
    n = 4
    K = _csc(sps.csc_matrix(np.array([[2.0, 1, 0, 0], [1, 2, 1, 0], [0, 1, 2, 1], [0, 0, 1, 2]])))
    ip, ii = K.indptr.astype(np.int32), K.indices.astype(np.int32)
    bad = np.array([0, 1, 1, 3], dtype=np.int32)
    ident = np.arange(n, dtype=np.int32)
    h = C.c_void_p()
    assert lib.ocb_refactor_create(C.byref(h), n, ip.ctypes.data, ii.ctypes.data, bad.ctypes.data,
                                   ident.ctypes.data) == -1
    assert b'permutation' in lib.ocb_last_error()
    rf = _lu_worker._Refactor(n, ip, ii, ident, ident)
    assert rf.numeric(K.data) is not None
    Z = K.copy()
    Z.data[:] = K.data
    Z[0, 0] = 0.0                     # stored zero on the first static pivot
    Z = _csc(Z)
    assert Z.nnz == K.nnz
    assert rf.numeric(Z.data) is None
    assert b'pivot' in lib.ocb_last_error()
    nan = K.data.copy()
    nan[3] = np.nan
    assert rf.numeric(nan) is None
    # empty system
    e = np.zeros(1, dtype=np.int32)
    rf0 = _lu_worker._Refactor(0, e, e, e, e)
    assert rf0.info['nnzL'] == 0


def test_worker_uses_the_static_pivots_it_is_given(cav10, monkeypatch):
    """_lu_worker._build: the main process fixes ordering and pivots once per pattern (one SuperLU
    run on the first matrix, ``static_pivots``); every job then takes the numeric-only path, is
    guarded, and its program reproduces SuperLU's solve.  Without pivots, or switched off: SuperLU."""
    monkeypatch.delenv('OCB_REFACTOR', raising=False)
    monkeypatch.delenv('OCB_LU_GUARD_TOL', raising=False)
    _lu_worker._REFAC.clear()
    flags = 2 | 4 | (2 << 4)
    K0 = _shifted(cav10, 2e-3, -1.0)
    a0 = dv._csc_args(K0, dict(dv.LU_OPTIONS)) + (232448, flags)
    q = _lu_worker.order_only(a0)
    piv = _lu_worker.static_pivots(a0 + (q,))
    n = K0.shape[0]
    paths, images = [], []
    for tau, p in ((2e-3, -1.0), (3e-4, -5.0), (1e-3, -1.3)):
        K = _shifted(cav10, tau, p)
        a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, flags, q, piv)
        used, img, nbytes, tf, tp, order, guard = _lu_worker._build(a)
        paths.append(guard[2])
        images.append(img)
        assert guard[0] <= _lu_worker.GUARD_TOL and guard[1] == 0 and guard[3] == 0
        assert img is not None and img.nbytes == nbytes
    assert paths == ['static', 'static', 'static']
    assert len({im.nbytes for im in images}) == 1                  # one structure, whatever the values
    # the same job again, in a "fresh worker": bit-identical image (no dependence on job history)
    _lu_worker._REFAC.clear()
    again = _lu_worker._build(dv._csc_args(_shifted(cav10, 1e-3, -1.3), dict(dv.LU_OPTIONS))
                              + (232448, flags, q, piv))[1]
    assert np.array_equal(again, images[2])
    assert _lu_worker._build(a0 + (q,))[6][2] == 'slu' and _lu_worker._build(a0 + (q, None))[6][2] == 'slu'
    # the static-pivot factors through the program builder and the numpy executor
    rf = next(iter(_lu_worker._REFAC.values()))
    K = _csc(_shifted(cav10, 7e-4, -3.0))
    arrs = rf.numeric(K.data)
    prog = _program(arrs, n, flags=flags & ~2)
    rng = np.random.default_rng(3)
    B = rng.standard_normal((n, 3))
    X = np.zeros((prog[0]['n_ext'], 3))
    X[arrs[6]] = B
    _execute(prog, X, check_hazards=True)
    got = X[arrs[7]]
    ref = spsla.splu(K).solve(B)
    assert np.linalg.norm(got - ref) <= 1e-11*np.linalg.norm(ref)
    # switched off: SuperLU every time
    monkeypatch.setenv('OCB_REFACTOR', '0')
    a = dv._csc_args(_shifted(cav10, 3e-4, -5.0), dict(dv.LU_OPTIONS)) + (232448, flags, q, piv)
    assert _lu_worker._build(a)[6][2] == 'slu'
    _lu_worker._REFAC.clear()


def test_guard_rejects_bad_static_pivots_and_superlu_takes_over(cav10, monkeypatch):
    """A later matrix of the same pattern whose values make the static pivot order unusable (a
    velocity diagonal entry shrunk by fourteen orders of magnitude): the numeric-only image fails the
    residual guard or hits a zero pivot, and the matrix is factorised by SuperLU again."""
    monkeypatch.delenv('OCB_REFACTOR', raising=False)
    monkeypatch.delenv('OCB_LU_GUARD_TOL', raising=False)
    _lu_worker._REFAC.clear()
    flags = 2 | 4 | (2 << 4)
    K0 = _csc(_shifted(cav10, 2e-3, -1.0))
    a0 = dv._csc_args(K0, dict(dv.LU_OPTIONS)) + (232448, flags)
    q = _lu_worker.order_only(a0)
    piv = _lu_worker.static_pivots(a0 + (q,))
    assert _lu_worker._build(a0 + (q, piv))[6][2] == 'static'
    K1 = K0.copy()
    # the entry that is eliminated FIRST under the static order: nothing repairs it by fill
    rf = next(iter(_lu_worker._REFAC.values()))
    first = int(np.nonzero(rf.perm_r == 0)[0][0]), int(np.nonzero(rf.perm_c == 0)[0][0])
    assert K1[first[0], first[1]] != 0.0
    K1[first[0], first[1]] *= 1e-14
    K1 = _csc(K1)
    assert np.array_equal(K1.indices, K0.indices)
    a1 = dv._csc_args(K1, dict(dv.LU_OPTIONS)) + (232448, flags, q, piv)
    used, img, nbytes, tf, tp, order, guard = _lu_worker._build(a1)
    assert guard[2] == 'slu' and guard[3] == 1 and guard[0] <= _lu_worker.GUARD_TOL
    _lu_worker._REFAC.clear()


def test_pinned_slot_keeps_the_structure_between_images(cav10, monkeypatch):
    """ocb_lu_pack_host_into on a buffer that already holds an image of the same structure rewrites
    numbers and permutations only (header slot 21 = structure hash); whatever the buffer held
    before - the same structure with other numbers, another structure, garbage - the bytes are
    those of a build into a fresh buffer."""
    monkeypatch.delenv('OCB_NO_SLOT_REUSE', raising=False)
    lib = _cabi.load()
    flags = 4 | (2 << 4)
    K0 = _csc(_shifted(cav10, 2e-3, -1.0))
    n = K0.shape[0]
    a0 = dv._csc_args(K0, dict(dv.LU_OPTIONS)) + (232448, flags | 2)
    q = _lu_worker.order_only(a0)
    pr, pc = _lu_worker.static_pivots(a0 + (q,))
    rf = _lu_worker._Refactor(n, K0.indptr, K0.indices, pr, pc)

    def arrays(tau, p):
        return [x.copy() for x in rf.numeric(_csc(_shifted(cav10, tau, p)).data)]

    def into(arrs, buf, fl=flags):
        nb = C.c_int64(0)
        _cabi.check(lib.ocb_lu_pack_host_into(n, *[x.ctypes.data for x in arrs], 232448, fl,
                                              buf.ctypes.data, buf.nbytes, C.byref(nb)), 'pack_into')
        return buf[:nb.value].copy()
    A1, A2 = arrays(3e-4, -5.0), arrays(1e-3, -1.3)
    ref1 = _lu_worker.pack_image(A1, n, 232448, flags)
    ref2 = _lu_worker.pack_image(A2, n, 232448, flags)
    assert ref1.nbytes == ref2.nbytes and not np.array_equal(ref1, ref2)
    assert ref1[168:176].any() and np.array_equal(ref1[168:176], ref2[168:176])      # same structure hash
    buf = np.full(ref1.nbytes + 4096, 0xAB, dtype=np.uint8)
    assert np.array_equal(into(A1, buf), ref1)                       # garbage before
    assert np.array_equal(into(A2, buf), ref2)                       # same structure before: numbers only
    assert np.array_equal(into(A1, buf), ref1)
    # another structure in between (other cluster size -> other stream layout)
    other = into(A1, buf, fl=4 | (4 << 4))
    assert not np.array_equal(other[168:176], ref1[168:176])
    assert np.array_equal(into(A2, buf), ref2)
    # a buffer whose header claims the structure but whose body is not trusted when the hash is off
    buf[168:176] = 0
    buf[4096:8192] = 0x55
    assert np.array_equal(into(A1, buf), ref1)


def _nd(G, leaf=4):
    lib = _cabi.load()
    G = sps.csr_matrix(G)
    G.sort_indices()
    n = G.shape[0]
    ip, ii = G.indptr.astype(np.int32), G.indices.astype(np.int32)
    q = np.empty(n, dtype=np.int32)
    _cabi.check(lib.ocb_order_nd(n, ip.ctypes.data, ii.ctypes.data, leaf, q.ctypes.data), 'ocb_order_nd')
    return q


def _sublevels_and_fill(S, q):
    """Sub-levels of the gather program and nnz(L+U) of the matrix S under the symmetric order q."""
    n = S.shape[0]
    a = dv._csc_args(S, dict(dv.LU_OPTIONS)) + (232448, 6, np.ascontiguousarray(q, dtype=np.int32))
    arrs = _lu_worker.factor_arrays(a, transposed=True)
    info = _program(arrs, n, flags=6)[0]
    return info['nsub_L'] + info['nsub_U'], len(arrs[1]) + len(arrs[4])


def test_nested_dissection_orders_every_node_once_and_flattens_the_tree(cav10):
    """ocb_order_nd: a permutation whatever the graph (grid, path, disconnected, complete, empty
    rows); on a 2-D grid the elimination tree is far lower than under minimum degree at no more
    fill; on the cavity saddle pattern the gather program gets at most 60 % of the sub-levels."""
    nx = 40
    I = sps.identity(nx)
    T = sps.diags([np.ones(nx-1), np.ones(nx-1)], [-1, 1])
    grid = (sps.kron(I, T) + sps.kron(T, I)).tocsr()
    q = _nd(grid)
    assert sorted(q.tolist()) == list(range(nx*nx))
    S = sps.csc_matrix(grid + sps.identity(nx*nx)*10.0)
    qm = np.argsort(spsla.splu(S, permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0.0,
                               options=dict(SymmetricMode=True)).perm_c)
    (l_nd, f_nd), (l_md, f_md) = _sublevels_and_fill(S, q), _sublevels_and_fill(S, qm)
    assert l_nd <= 0.8*l_md and f_nd <= 1.25*f_md, (l_nd, l_md, f_nd, f_md)
    # degenerate graphs
    path = sps.diags([np.ones(99), np.ones(99)], [-1, 1]).tocsr()
    assert sorted(_nd(path, leaf=1).tolist()) == list(range(100))
    two = sps.block_diag([grid[:50][:, :50], path, sps.csr_matrix((3, 3))]).tocsr()   # components + isolated nodes
    assert sorted(_nd(two).tolist()) == list(range(two.shape[0]))
    full = sps.csr_matrix(np.ones((30, 30)) - np.eye(30))
    assert sorted(_nd(full).tolist()) == list(range(30))
    assert _nd(sps.csr_matrix((0, 0))).size == 0
    lib = _cabi.load()
    bad_ip, bad_ii, out = np.array([0, 1], np.int32), np.array([5], np.int32), np.empty(1, np.int32)
    assert lib.ocb_order_nd(1, bad_ip.ctypes.data, bad_ii.ctypes.data, 4, out.ctypes.data) == -1
    # the saddle-point pattern through order_only (ND + delayed zero diagonals) vs minimum degree
    K = _shifted(cav10, 2e-3, -1.0)
    n = K.shape[0]
    a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, 6)
    levels = {}
    import os
    for which in ('nd', 'mmd'):
        os.environ['OCB_ORDERING'] = which
        try:
            qq = _lu_worker.order_only(a)
        finally:
            os.environ.pop('OCB_ORDERING')
        assert sorted(qq.tolist()) == list(range(n))
        arrs = _lu_worker.factor_arrays(a + (qq,), transposed=True)
        info = _program(arrs, n, flags=6)[0]
        levels[which] = (info['nsub_L'] + info['nsub_U'], len(arrs[1]) + len(arrs[4]))
    assert levels['nd'][0] <= 0.6*levels['mmd'][0] and levels['nd'][1] <= 1.25*levels['mmd'][1], levels   # (N=10: +15 % fill; N=25: -14 %, N=50: -24 %)


def test_random_patterns_and_pivot_orders_property():
    """Property test (hypothesis): for random sparse patterns, random row / column permutations
    that put a dominant entry on every pivot position, and random values, the numeric-only
    factorisation reproduces ``P A Q`` and solves ``A x = b``; a second matrix with the same
    pattern reuses the handle."""
    hyp = pytest.importorskip('hypothesis')
    st = pytest.importorskip('hypothesis.strategies')

    @hyp.settings(max_examples=25, deadline=None)
    @hyp.given(n=st.integers(2, 60), dens=st.floats(0.02, 0.5), seed=st.integers(0, 10**6))
    def check(n, dens, seed):
        rng = np.random.default_rng(seed)
        S = sps.random(n, n, density=dens, format='coo', random_state=seed)
        pr, pc = rng.permutation(n).astype(np.int32), rng.permutation(n).astype(np.int32)
        # pivot positions: entry (i, j) of A lands on the diagonal of P A Q iff pr[i] == pc[j]
        inv_pc = np.empty(n, dtype=np.int64)
        inv_pc[pc] = np.arange(n)
        rows = np.arange(n)
        cols = inv_pc[pr[rows]]
        mats = []
        for rep in range(2):
            vals = rng.standard_normal(S.nnz)
            A = sps.coo_matrix((vals, (S.row, S.col)), shape=(n, n)).tocsc()
            A = A + sps.coo_matrix((np.full(n, 2.0 + abs(vals).sum()), (rows, cols)), shape=(n, n)).tocsc()
            mats.append(_csc(A))
        assert np.array_equal(mats[0].indices, mats[1].indices)
        rf = _lu_worker._Refactor(n, mats[0].indptr, mats[0].indices, pr, pc)
        for A in mats:
            L, U, qr, qc = _lu_of(rf, A)
            Cm = _permuted(A, qr, qc)
            assert abs(L @ U - Cm).max() <= 1e-12*abs(Cm).max()
            b = rng.standard_normal(n)
            y = np.empty(n)
            y[qr] = b
            z = spsla.spsolve_triangular(sps.csr_matrix(U), spsla.spsolve_triangular(
                sps.csr_matrix(L), y, lower=True, unit_diagonal=True), lower=False)
            assert np.linalg.norm(A @ z[qc] - b) <= 1e-10*np.linalg.norm(b)
    check()


def test_nested_dissection_random_graphs_property():
    """Property test: whatever the (symmetric) graph and the leaf size, ``ocb_order_nd`` returns a
    permutation, and it is deterministic."""
    hyp = pytest.importorskip('hypothesis')
    st = pytest.importorskip('hypothesis.strategies')

    @hyp.settings(max_examples=40, deadline=None)
    @hyp.given(n=st.integers(1, 400), dens=st.floats(0.0, 0.2), leaf=st.integers(1, 64), seed=st.integers(0, 10**6))
    def check(n, dens, leaf, seed):
        S = sps.random(n, n, density=dens, format='csr', random_state=seed)
        G = ((S + S.T) != 0).astype(np.float64).tocsr()
        q1, q2 = _nd(G, leaf), _nd(G, leaf)
        assert np.array_equal(np.sort(q1), np.arange(n)) and np.array_equal(q1, q2)
    check()


def test_zero_diagonal_delay_matches_the_python_statement(cav10):
    """``ocb_order_delay_zero_diagonals`` (used by ``order_only``) against the Python statement of
    the rule, ``_lu_worker._delay_zero_diagonals``, on the saddle-point pattern and random orders;
    every pressure node ends up right after a velocity neighbour or at the end."""
    lib = _cabi.load()
    K = sps.csr_matrix(_shifted(cav10, 2e-3, -1.0))
    n = K.shape[0]
    P = sps.csr_matrix((np.ones(K.nnz), K.indices, K.indptr), shape=K.shape)
    P = ((P + P.T) != 0).astype(np.float64).tocsr()
    P.setdiag(0.0)
    P.eliminate_zeros()
    P.sort_indices()
    ip, ii = P.indptr.astype(np.int32), P.indices.astype(np.int32)
    dz = (K.diagonal() == 0.0)
    assert dz.any() and not dz.all()
    dz8 = np.ascontiguousarray(dz, dtype=np.uint8)
    for seed in range(4):
        q = np.random.default_rng(seed).permutation(n).astype(np.int32)
        ref = _lu_worker._delay_zero_diagonals(ip, ii, dz, q)
        out = np.empty(n, dtype=np.int32)
        _cabi.check(lib.ocb_order_delay_zero_diagonals(n, ip.ctypes.data, ii.ctypes.data, dz8.ctypes.data,
                                                       q.ctypes.data, out.ctypes.data), 'delay')
        assert np.array_equal(ref, out)
        pos = np.empty(n, dtype=np.int64)
        pos[out] = np.arange(n)
        for v in np.flatnonzero(dz)[:50]:
            nb = ii[ip[v]:ip[v+1]]
            nb = nb[~dz[nb]]
            assert nb.size == 0 or pos[nb].min() < pos[v]       # a velocity neighbour goes first
    bad = np.zeros(n, dtype=np.int32)
    assert lib.ocb_order_delay_zero_diagonals(n, ip.ctypes.data, ii.ctypes.data, dz8.ctypes.data,
                                              bad.ctypes.data, out.ctypes.data) == -1
