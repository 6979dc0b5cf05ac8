"""CPU suite, part 5: the host-side analysis of the solve kernel (csrc/lu_program.cu).

``ocb_lu_program_create`` turns SuperLU's factors into the gather program the CUDA kernel
executes (supernodes with inverted diagonal blocks -> a short sequence of sub-levels).  It
is pure host code, so it is checked here without a GPU: the exported program is executed
with numpy and must reproduce ``SuperLU.solve`` to rounding, rows of a sub-level must be
independent of each other, and the depth must be far below the scalar level count."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

from optconpy_b200 import _cabi, _lu_worker, device as dv, problems as pb


def _program(arrs, n, flags=0):
    lib = _cabi.load()
    h = C.c_void_p()
    _cabi.check(lib.ocb_lu_program_create(C.byref(h), n, *[a.ctypes.data for a in arrs[:6]], flags),
                'ocb_lu_program_create')
    info = (C.c_int64*12)()
    _cabi.check(lib.ocb_lu_program_info(h, info), 'info')
    keys = ['n', 'n_ext', 'ymax', 'nsub_L', 'nsub_U', 'nsuper', 'max_w', 'nslice', 'nrows',
            'nent', 'nnzL', 'nnzU']
    info = dict(zip(keys, [int(v) for v in info]))
    nsub = info['nsub_L'] + info['nsub_U']
    sub_ptr = np.zeros(nsub+1, dtype=np.int32)
    slices = np.zeros((info['nslice'], 4), dtype=np.int32)
    dst = np.zeros(info['nrows'], dtype=np.int32)
    init = np.zeros(info['nrows'], dtype=np.int32)
    scale = np.zeros(info['nrows'], dtype=np.float64)
    col = np.zeros(info['nent'], dtype=np.int32)
    val = np.zeros(info['nent'], dtype=np.float64)
    _cabi.check(lib.ocb_lu_program_export(h, *[a.ctypes.data for a in
                                               (sub_ptr, slices, dst, init, scale, col, val)]), 'export')
    lib.ocb_lu_program_destroy(h)
    return info, sub_ptr, slices, dst, init, scale, col, val


def _slice_rows(sl, col, val):
    """(row index, cols, vals) of every row of a slice (SELL-32 layout, padding dropped)."""
    ebase, trips, gn, q0 = [int(v) for v in sl]
    g, nr = gn & 255, gn >> 8
    G = 1 << g
    assert 1 <= nr <= (32 >> g)
    blk_c = col[ebase:ebase+32*trips].reshape(trips, 32)
    blk_v = val[ebase:ebase+32*trips].reshape(trips, 32)
    # lanes of rows beyond nr are pure padding
    assert not blk_v[:, nr*G:].any()
    out = []
    for r in range(nr):
        c = blk_c[:, r*G:(r+1)*G].reshape(-1)
        v = blk_v[:, r*G:(r+1)*G].reshape(-1)
        out.append((q0 + r, c, v))
    return out


def _execute(prog, X, check_hazards=False):
    """numpy executor of the gather program on the extended block X (n_ext x k)."""
    info, sub_ptr, slices, dst, init, scale, col, val = prog
    for sb in range(len(sub_ptr)-1):
        rows = []
        for sl in slices[sub_ptr[sb]:sub_ptr[sb+1]]:
            rows += _slice_rows(sl, col, val)
        q = np.array([r[0] for r in rows])
        lens = np.array([len(r[1]) for r in rows])
        A = sps.csr_matrix((np.concatenate([r[2] for r in rows]),
                            np.concatenate([r[1] for r in rows]),
                            np.concatenate([[0], np.cumsum(lens)])), shape=(len(rows), info['n_ext']))
        ini = init[q]
        base = np.where(ini[:, None] >= 0, X[np.maximum(ini, 0)], 0.0)
        new = (base - A @ X)*scale[q, None]
        if check_hazards:
            d = dst[q]
            assert len(np.unique(d)) == len(d)                 # no two rows write one slot
            A.eliminate_zeros()
            read = np.union1d(A.indices, ini[(ini >= 0) & (ini != d)])
            assert not np.intersect1d(read, d).size           # nobody reads what a peer writes
        X[dst[q]] = new
    return X


def _saddle(prob, mu=-1.0, tau=0.05):
    M, A, J = prob['M'], prob['A'], prob['J']
    Nc = pb.convection_matrix(prob, pb.analytic_vortex)
    Ft = -(0.5*M.T + tau*(A.T + Nc.T))
    return dv.sadpnt_matrix(Ft + mu*M.T, J)


@pytest.mark.parametrize('opts', [dv.LU_OPTIONS, {}])
def test_program_reproduces_superlu(cav10, opts):
    K = _saddle(cav10)
    n = K.shape[0]
    arrs = _lu_worker.factor_arrays(dv._csc_args(K, dict(opts)))
    prog = _program(arrs, n)
    info = prog[0]
    assert info['n'] == n and info['n_ext'] == n + 2*info['ymax']
    rng = np.random.default_rng(0)
    B = rng.standard_normal((n, 3))
    X = np.zeros((info['n_ext'], 3))
    X[arrs[6]] = B                                   # x[perm_r[i]] = b[i]
    _execute(prog, X, check_hazards=True)
    got = X[arrs[7]]                                 # out[j] = x[perm_c[j]]
    ref = spsla.splu(K).solve(B)
    assert np.linalg.norm(got - ref) <= 1e-12*np.linalg.norm(ref)
    assert np.linalg.norm(K @ got - B) <= 1e-12*np.linalg.norm(B)
    # every entry of the factors is accounted for: off-block entries + inverse blocks
    assert info['nnzL'] == sps.csr_matrix((arrs[2], arrs[1], arrs[0]), shape=(n, n)).nnz - n \
        or info['nnzL'] <= len(arrs[2])
    assert info['nnzU'] == len(arrs[5])


def test_depth_is_cut_by_supernodes(cav10):
    """The point of the program: far fewer sequential steps than scalar level scheduling."""
    K = _saddle(cav10)
    n = K.shape[0]
    arrs = _lu_worker.factor_arrays(dv._csc_args(K, dict(dv.LU_OPTIONS)))
    info = _program(arrs, n)[0]
    L = sps.csr_matrix((arrs[2], arrs[1], arrs[0]), shape=(n, n))
    U = sps.csr_matrix((arrs[5], arrs[4], arrs[3]), shape=(n, n))

    def scalar_levels(T, upper):
        lev = np.zeros(n, dtype=np.int64)
        rng = range(n-1, -1, -1) if upper else range(n)
        for i in rng:
            c = T.indices[T.indptr[i]:T.indptr[i+1]]
            c = c[c != i]
            if c.size:
                lev[i] = lev[c].max() + 1
        return lev.max() + 1
    scalar = scalar_levels(L, False) + scalar_levels(U, True)
    assert info['nsub_L'] + info['nsub_U'] < scalar/3
    assert info['nsuper'] < n and info['max_w'] > 8


def test_small_and_degenerate_factors():
    lib = _cabi.load()
    # diagonal matrix: L = I (no work), U = diag -> one sub-level of n scalings
    n = 5
    Lc = sps.identity(n, format='csr')
    Uc = sps.diags(np.arange(1.0, n+1)).tocsr()
    arrs = [Lc.indptr.astype(np.int32), Lc.indices.astype(np.int32), Lc.data,
            Uc.indptr.astype(np.int32), Uc.indices.astype(np.int32), Uc.data]
    prog = _program(arrs, n)
    assert prog[0]['nsub_L'] == 0 and prog[0]['nsub_U'] == 1 and prog[0]['nrows'] == n
    X = np.ones((prog[0]['n_ext'], 1))
    _execute(prog, X)
    assert np.allclose(X[:n, 0], 1.0/np.arange(1.0, n+1))
    # dense 6x6: a single supernode, two sub-levels per sweep
    rng = np.random.default_rng(1)
    Ad = rng.standard_normal((6, 6)) + 6*np.eye(6)
    import scipy.linalg as sla
    Pm, Ld, Ud = sla.lu(Ad)
    Lc, Uc = sps.csr_matrix(Ld), sps.csr_matrix(Ud)
    Lc.sort_indices()
    Uc.sort_indices()
    arrs = [Lc.indptr.astype(np.int32), Lc.indices.astype(np.int32), Lc.data,
            Uc.indptr.astype(np.int32), Uc.indices.astype(np.int32), Uc.data]
    prog = _program(arrs, 6)
    assert prog[0]['nsuper'] == 1 and prog[0]['max_w'] == 6
    assert prog[0]['nsub_L'] == 2 and prog[0]['nsub_U'] == 2
    b = rng.standard_normal((6, 2))
    X = np.zeros((prog[0]['n_ext'], 2))
    X[:6] = Pm.T @ b
    _execute(prog, X, check_hazards=True)
    assert np.allclose(X[:6], np.linalg.solve(Ad, b))
    # zero pivot and a misplaced entry are reported, not executed
    Ubad = Uc.copy()
    Ubad.data[Ubad.indptr[2]] = 0.0
    h = C.c_void_p()
    bad = [arrs[0], arrs[1], arrs[2], Ubad.indptr.astype(np.int32), Ubad.indices.astype(np.int32),
           Ubad.data]
    rc = lib.ocb_lu_program_create(C.byref(h), 6, *[a.ctypes.data for a in bad], 0)
    assert rc == -3 and b'zero pivot' in lib.ocb_last_error()
    rc = lib.ocb_lu_program_create(C.byref(h), 6, *[a.ctypes.data for a in
                                                    (arrs[3], arrs[4], arrs[5], arrs[3], arrs[4], arrs[5])], 0)
    assert rc == -1 and b'wrong side' in lib.ocb_last_error()
    # n = 0
    z = np.zeros(1, dtype=np.int32)
    e = np.zeros(0)
    prog0 = _program([z, z, e, z, z, e], 0)
    assert prog0[0]['nrows'] == 0 and prog0[0]['nslice'] == 0


def test_reused_ordering_gives_the_same_solution(cav10):
    """Second and later factorisations of one sparsity pattern reuse the first one's ordering
    (symmetric pre-permutation + NATURAL column order, _lu_worker.factor_arrays): the composed
    permutations must still solve the ORIGINAL system, with no more fill than before."""
    K = _saddle(cav10)
    n = K.shape[0]
    a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, 0)
    arrs1, order = _lu_worker.factor_arrays(a, want_order=True)
    assert order is not None and sorted(order.tolist()) == list(range(n))
    K2 = _saddle(cav10, mu=-3.0, tau=0.02)                       # same pattern, other values
    a2 = dv._csc_args(K2, dict(dv.LU_OPTIONS)) + (232448, 0, order)
    arrs2, order2 = _lu_worker.factor_arrays(a2, want_order=True)
    assert order2 is None
    prog = _program(arrs2, n)
    rng = np.random.default_rng(5)
    B = rng.standard_normal((n, 2))
    X = np.zeros((prog[0]['n_ext'], 2))
    X[arrs2[6]] = B
    _execute(prog, X, check_hazards=True)
    got = X[arrs2[7]]
    assert np.linalg.norm(K2 @ got - B) <= 1e-12*np.linalg.norm(B)
    assert len(arrs2[2]) + len(arrs2[5]) <= 1.05*(len(arrs1[2]) + len(arrs1[5]))


@pytest.mark.parametrize('reuse', [False, True])
def test_transposed_layout(cav10, reuse):
    """Factors of A^T taken column-wise (no CSC->CSR conversion on the host): lower factor with
    the pivots, unit upper factor, swapped permutations - same solution of A x = b."""
    K = _saddle(cav10)
    n = K.shape[0]
    a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, 2)
    if reuse:
        a = a + (_lu_worker.order_only(a),)
    arrs = _lu_worker.factor_arrays(a, transposed=True)
    prog = _program(arrs, n, flags=2)
    rng = np.random.default_rng(8)
    B = rng.standard_normal((n, 3))
    X = np.zeros((prog[0]['n_ext'], 3))
    X[arrs[6]] = B
    _execute(prog, X, check_hazards=True)
    got = X[arrs[7]]
    ref = spsla.splu(K).solve(B)
    assert np.linalg.norm(got - ref) <= 1e-12*np.linalg.norm(ref)
    assert np.linalg.norm(K @ got - B) <= 1e-12*np.linalg.norm(B)
    # same amount of work as the row-wise layout (structurally symmetric pattern)
    a0 = a[:6] + (0,) + a[7:]
    arrs0 = _lu_worker.factor_arrays(a0)
    p0 = _program(arrs0, n)[0]
    assert prog[0]['nent'] <= 1.15*p0['nent'] and prog[0]['nsub_L'] + prog[0]['nsub_U'] <= 1.3*(p0['nsub_L'] + p0['nsub_U'])


def test_unsorted_upper_rows_are_sorted_by_the_builder(cav10):
    """SuperLU's columns come in supernodal order; build_lu_program sorts the rows of the upper
    factor itself (counting sort): same program as from rows sorted beforehand; a duplicate
    column index is an error."""
    K = _saddle(cav10)
    n = K.shape[0]
    a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, 2)
    arrs = _lu_worker.factor_arrays(a, transposed=True)
    rp, ci = arrs[3], arrs[4]
    assert any(np.any(np.diff(ci[rp[i]:rp[i+1]]) <= 0) for i in range(n))   # really unsorted
    up = sps.csr_matrix((arrs[5], ci, rp), shape=(n, n))
    up.sort_indices()
    arrs_s = arrs[:3] + [up.indptr.astype(np.int32), up.indices.astype(np.int32), up.data] + arrs[6:]
    p1, p2 = _program(arrs, n, flags=2), _program(arrs_s, n, flags=2)
    assert p1[0] == p2[0]
    for x, y in zip(p1[1:], p2[1:]):
        assert np.array_equal(x, y)
    lib = _cabi.load()
    dup = ci.copy()
    i = int(np.argmax(np.diff(rp) >= 3))
    dup[rp[i] + 2] = dup[rp[i] + 1]
    h = C.c_void_p()
    rc = lib.ocb_lu_program_create(C.byref(h), n, *[x.ctypes.data for x in arrs[:4] + [dup, arrs[5]]], 2)
    assert rc == -1 and b'duplicate' in lib.ocb_last_error()


@pytest.mark.parametrize('transposed', [False, True])
def test_one_step_supernodes(cav10, transposed, monkeypatch):
    """flags bit 2: small supernodes are solved in one sub-level (inverse-multiplied rows writing
    to the y region, zero-length rows copying the result home one sub-level later).  Same
    solution as the two-step program, hazard-free, clearly fewer sub-levels, bounded growth."""
    K = _saddle(cav10)
    n = K.shape[0]
    base = 2 if transposed else 0
    a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, base)
    a = a + (_lu_worker.order_only(a),)
    arrs = _lu_worker.factor_arrays(a, transposed=transposed)
    rng = np.random.default_rng(11)
    B = rng.standard_normal((n, 3))
    ref = spsla.splu(K).solve(B)
    out = {}
    for flags in (base, base | 4):
        prog = _program(arrs, n, flags=flags)
        X = np.zeros((prog[0]['n_ext'], 3))
        X[arrs[6]] = B
        _execute(prog, X, check_hazards=True)
        got = X[arrs[7]]
        assert np.linalg.norm(got - ref) <= 1e-12*np.linalg.norm(ref)
        out[flags] = prog[0]
    two, one = out[base], out[base | 4]
    assert one['nsub_L'] + one['nsub_U'] <= 0.7*(two['nsub_L'] + two['nsub_U'])
    assert one['nent'] <= 1.25*two['nent'] and one['n_ext'] == n + 2*one['ymax']
    assert one['nrows'] == two['nrows']          # A+B rows became one-step + copy rows
    # width cap 0 -> nothing is merged: the very same program as without the flag
    monkeypatch.setenv('OCB_MERGE_W', '1')
    p_off, p_ref = _program(arrs, n, flags=base | 4), _program(arrs, n, flags=base)
    assert p_off[0] == p_ref[0] and all(np.array_equal(x, y) for x, y in zip(p_off[1:], p_ref[1:]))
    # everything merged, however much the rows grow: still exact
    monkeypatch.setenv('OCB_MERGE_W', '512')
    monkeypatch.setenv('OCB_MERGE_GROWTH', '1e9')
    prog = _program(arrs, n, flags=base | 4)
    X = np.zeros((prog[0]['n_ext'], 3))
    X[arrs[6]] = B
    _execute(prog, X, check_hazards=True)
    assert np.linalg.norm(X[arrs[7]] - ref) <= 1e-12*np.linalg.norm(ref)
    assert prog[0]['nsub_L'] + prog[0]['nsub_U'] <= one['nsub_L'] + one['nsub_U']


def test_one_step_supernodes_dense_block():
    """A dense matrix is one supernode: one sub-level per sweep plus the copy level."""
    import scipy.linalg as sla
    rng = np.random.default_rng(3)
    Ad = rng.standard_normal((7, 7)) + 7*np.eye(7)
    Pm, Ld, Ud = sla.lu(Ad)
    Lc, Uc = sps.csr_matrix(Ld), sps.csr_matrix(Ud)
    Lc.sort_indices()
    Uc.sort_indices()
    arrs = [Lc.indptr.astype(np.int32), Lc.indices.astype(np.int32), Lc.data,
            Uc.indptr.astype(np.int32), Uc.indices.astype(np.int32), Uc.data]
    prog = _program(arrs, 7, flags=4)
    assert prog[0]['nsuper'] == 1 and prog[0]['nsub_L'] == 2 and prog[0]['nsub_U'] == 2
    b = rng.standard_normal((7, 2))
    X = np.zeros((prog[0]['n_ext'], 2))
    X[:7] = Pm.T @ b
    _execute(prog, X, check_hazards=True)
    assert np.allclose(X[:7], np.linalg.solve(Ad, b))


@pytest.mark.parametrize('flags', [2, 6, 0, 3, 7])
def test_structure_template_refills_numbers_only(cav10, flags, monkeypatch):
    """A second factor with the SAME index arrays (other shift: same ordering, same pivots) is
    served from the cached structure: only numbers are recomputed.  The result must be the very
    program a fresh build gives - compared array by array - and the device images must be
    identical byte for byte; a zero pivot is still reported; other structures miss."""
    lib = _cabi.load()
    transposed = bool(flags & 2)
    Ks = [_saddle(cav10, mu=mu) for mu in (-1.0, -2.5)]
    n = Ks[0].shape[0]
    a0 = dv._csc_args(Ks[0], dict(dv.LU_OPTIONS)) + (232448, flags)
    q = _lu_worker.order_only(a0)
    arrs = [_lu_worker.factor_arrays(dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, flags, q),
                                     transposed=transposed) for K in Ks]
    for x, y in zip(arrs[0][:2] + arrs[0][3:5], arrs[1][:2] + arrs[1][3:5]):
        assert np.array_equal(x, y)                    # same structure, other numbers
    monkeypatch.setenv('OCB_NO_TEMPLATE', '1')
    fresh = [_program(a, n, flags=flags) for a in arrs]
    fresh_img = [_lu_worker.pack_image(a, n, 232448, flags) for a in arrs]
    monkeypatch.delenv('OCB_NO_TEMPLATE')
    h0 = lib.ocb_lu_program_template_hits()
    first = _program(arrs[0], n, flags=flags)          # records (or hits an earlier test's entry)
    h1 = lib.ocb_lu_program_template_hits()
    second = _program(arrs[1], n, flags=flags)
    assert lib.ocb_lu_program_template_hits() == h1 + 1 and h1 - h0 in (0, 1)
    for got, ref in ((first, fresh[0]), (second, fresh[1])):
        assert got[0] == ref[0]
        for x, y in zip(got[1:], ref[1:]):
            assert np.array_equal(x, y)
    for a, ref in zip(arrs, fresh_img):
        img = _lu_worker.pack_image(a, n, 232448, flags)
        # header slot 21 = hash of the structure: set by the image-template path only (the wide
        # layout, flags bit 0, is packed without one)
        assert (img[168:176].any() or flags & 1) and not ref[168:176].any()
        img[168:176] = 0
        assert np.array_equal(img, ref)
    # the refilled program solves the second system
    rng = np.random.default_rng(2)
    B = rng.standard_normal((n, 2))
    X = np.zeros((second[0]['n_ext'], 2))
    X[arrs[1][6]] = B
    _execute(second, X, check_hazards=True)
    assert np.linalg.norm(Ks[1] @ X[arrs[1][7]] - B) <= 1e-12*np.linalg.norm(B)
    # zero pivot: reported by the refill as by the fresh build
    piv = 0 if transposed else 3                       # the factor that carries the pivots
    bad = [x.copy() for x in arrs[1]]
    rp, ci = bad[piv], bad[piv + 1]
    row = n//2
    bad[piv + 2][rp[row] + int(np.argmax(ci[rp[row]:rp[row+1]] == row))] = 0.0
    h = C.c_void_p()
    rc = lib.ocb_lu_program_create(C.byref(h), n, *[x.ctypes.data for x in bad[:6]], flags)
    assert rc == -3 and b'zero pivot' in lib.ocb_last_error()
    # another structure (coarser factor of another matrix) does not hit
    h2 = lib.ocb_lu_program_template_hits()
    K3 = _saddle(cav10, mu=-1.0) + sps.identity(n, format='csc')*1e-3     # new pattern (pressure block)
    a3 = _lu_worker.factor_arrays(dv._csc_args(K3, dict(dv.LU_OPTIONS)) + (232448, flags), transposed=transposed)
    p3 = _program(a3, n, flags=flags)
    assert lib.ocb_lu_program_template_hits() == h2
    X = np.zeros((p3[0]['n_ext'], 2))
    X[a3[6]] = B
    _execute(p3, X)
    assert np.linalg.norm(K3 @ X[a3[7]] - B) <= 1e-11*np.linalg.norm(B)


def test_structure_templates_are_evicted_not_leaked():
    """More than four structures in flight: the least recently used one goes; everything stays
    correct (diagonal systems of different sizes are different structures)."""
    lib = _cabi.load()
    for rep in range(2):
        for n in range(3, 10):
            Lc = sps.identity(n, format='csr')
            Uc = sps.diags(np.arange(1.0, n+1) + rep).tocsr()
            arrs = [Lc.indptr.astype(np.int32), Lc.indices.astype(np.int32), Lc.data,
                    Uc.indptr.astype(np.int32), Uc.indices.astype(np.int32), Uc.data]
            prog = _program(arrs, n)
            X = np.ones((prog[0]['n_ext'], 1))
            _execute(prog, X)
            assert np.allclose(X[:n, 0], 1.0/(np.arange(1.0, n+1) + rep))


@pytest.mark.parametrize('flags', [0, 2, 6])
def test_program_on_unstructured_pattern(cav10, flags):
    """The reference's own test perturbs F with a random sparse matrix of density 0.03
    (tests/test_units_compfacres_compress.py:49-52): no mesh structure, more off-diagonal pivots,
    much denser factors.  Builder, one-step supernodes and the numpy execution must not care."""
    M, A, J = cav10['M'], cav10['A'], cav10['J']
    NV = cav10['NV']
    F = -M - 0.1*A - sps.random(NV, NV, density=0.03, format='csr', random_state=11)
    K = dv.sadpnt_matrix(sps.csr_matrix(F.T - 2.0*M.T), J)
    n = K.shape[0]
    transposed = bool(flags & 2)
    a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, flags)
    a = a + (_lu_worker.order_only(a),)
    arrs = _lu_worker.factor_arrays(a, transposed=transposed)
    prog = _program(arrs, n, flags=flags)
    rng = np.random.default_rng(13)
    B = rng.standard_normal((n, 4))
    X = np.zeros((prog[0]['n_ext'], 4))
    X[arrs[6]] = B
    _execute(prog, X, check_hazards=True)
    got = X[arrs[7]]
    ref = spsla.splu(K).solve(B)
    # forward error: cond(K) * eps ~ 6e-9 is what any LU solve may show here; the program's is
    # 3e-11 .. 1.2e-10 over seeds / layouts, with row substitution or the blocked block inverses alike
    assert np.linalg.norm(got - ref) <= 5e-10*np.linalg.norm(ref)
    # cond(K) ~ 6e7 and ||x|| ~ 1e7 ||b||: the residual is measured normwise.  (Solving with the
    # explicit inverses of wide diagonal blocks is not backward stable row by row as
    # substitution is: relative to ||b|| this residual is 5e-7 where SuperLU reaches 3e-9.)
    assert np.linalg.norm(K @ got - B) <= 1e-13*sps.linalg.norm(K)*np.linalg.norm(got)
    img = _lu_worker.pack_image(arrs, n, 232448, flags | (2 << 4))
    assert img.nbytes > 0 and prog[0]['n_ext'] == n + 2*prog[0]['ymax']


def test_worker_builds_the_image_in_the_callers_segment(cav10, monkeypatch):
    """_lu_worker.factor_image_to_shm with a slot: ocb_lu_pack_host_into writes the image straight
    into the (pinned) shared-memory segment - same bytes as the malloc'ed image, fresh build and
    template hit alike; a segment that is too small falls back to a one-off segment.  (SuperLU path
    only: the static-pivot path has its own test in test_refactor.py.)"""
    from multiprocessing import shared_memory
    monkeypatch.setenv('OCB_REFACTOR', '0')
    K = _saddle(cav10)
    n = K.shape[0]
    flags = 2 | (2 << 4)
    a = dv._csc_args(K, dict(dv.LU_OPTIONS)) + (232448, flags)
    a = a + (_lu_worker.order_only(a),)
    ref = _lu_worker.pack_image(_lu_worker.factor_arrays(a, transposed=True), n, 232448, flags)
    seg = shared_memory.SharedMemory(create=True, size=ref.nbytes + 4096)
    try:
        np.frombuffer(seg.buf, dtype=np.uint8)[:] = 0xAB            # stale content must not survive
        for rep in range(2):                                         # fresh build, then template hit
            name, nbytes, tf, tp, order, guard = _lu_worker.factor_image_to_shm(a, (seg.name, seg.size))
            assert name is None and nbytes == ref.nbytes
            assert np.array_equal(np.frombuffer(seg.buf, dtype=np.uint8, count=nbytes), ref)
        name, nbytes, tf, tp, order, guard = _lu_worker.factor_image_to_shm(a, (seg.name, 1000))
        assert name is not None and nbytes == ref.nbytes
        one = shared_memory.SharedMemory(name=name)
        try:
            assert np.array_equal(np.frombuffer(one.buf, dtype=np.uint8, count=nbytes), ref)
        finally:
            one.close()
            one.unlink()
        lib = _cabi.load()
        assert b'needs' in lib.ocb_last_error()
    finally:
        _lu_worker._ADDRESS.pop(seg.name, None)
        att = _lu_worker._ATTACHED.pop(seg.name, None)
        import gc
        gc.collect()
        for s_ in (att, seg):
            try:
                if s_ is not None:
                    s_.close()
            except BufferError:
                pass
        seg.unlink()


def test_residual_guard_refactorises_in_safe_mode(monkeypatch):
    """the workers check every factor image where it is built - the finished program
    is executed on the host for one right-hand side (``ocb_lu_pack_host_checked``) and a
    backward error above the tolerance triggers a second factorisation with full partial
    pivoting and supernodes of at most 32 rows.  The reference test's randomly perturbed F
    (cond ~1e8, one 512-row block) is the case that needs it; the Oseen matrices never do."""
    prob = pb.drivcav_problem(15, 1.0)
    M, A, J = prob['M'], prob['A'], prob['J']
    NV = prob['NV']
    oseen = _saddle(prob)
    Fp = -M - 0.1*A - sps.random(NV, NV, density=0.03, format='csr', random_state=11)
    hard = dv.sadpnt_matrix(Fp.T - 1.0*M.T, J)
    monkeypatch.delenv('OCB_LU_GUARD_TOL', raising=False)
    tol = _lu_worker._guard_tol()
    flags = 2 | 4 | (2 << 4)
    img, tf, tp, order, guard = _lu_worker.factor_image(dv._csc_args(oseen, dict(dv.LU_OPTIONS)) + (232448, flags))
    assert guard[1] == 0 and guard[0] < 0.1*tol                 # fast path, far below the tolerance
    img, tf, tp, order, guard = _lu_worker.factor_image(dv._csc_args(hard, dict(dv.LU_OPTIONS)) + (232448, flags))
    assert guard[1] == 1 and guard[0] <= tol                    # re-factorised; now at substitution level
    # guard off: the fast image is handed out, with the larger error
    monkeypatch.setenv('OCB_LU_GUARD_TOL', '-1')
    img2, tf, tp, order, guard2 = _lu_worker.factor_image(dv._csc_args(hard, dict(dv.LU_OPTIONS)) + (232448, flags))
    assert guard2[:2] == (None, 0) and img2.nbytes != img.nbytes
    # the safe layout itself: no supernode wider than 32 rows, still reproduces SuperLU
    n = hard.shape[0]
    arrs = _lu_worker.factor_arrays(_lu_worker._safe_args(dv._csc_args(hard, dict(dv.LU_OPTIONS)) + (232448, flags)),
                                    transposed=True)
    prog = _program(arrs, n, flags=2 | 8)
    assert prog[0]['max_w'] <= 32
    B = np.random.default_rng(5).standard_normal((n, 2))
    X = np.zeros((prog[0]['n_ext'], 2))
    X[arrs[6]] = B
    _execute(prog, X)
    got = X[arrs[7]]
    assert np.linalg.norm(hard @ got - B) <= 1e-14*sps.linalg.norm(hard)*np.linalg.norm(got)


@pytest.mark.parametrize('flags', [2, 2 | 4])
def test_panel_form_reproduces_the_program(cav10, flags):
    """The PANEL form of the program (register-blocked layout of the all-columns-at-once
    executor: up to 8 rows of a supernode share one zero-padded column list) solves the same
    system: host execution of both forms against SuperLU, and the padding stays bounded."""
    K = _saddle(cav10)
    n = K.shape[0]
    lib = _cabi.load()
    arrs = _lu_worker.factor_arrays(dv._csc_args(K, dict(dv.LU_OPTIONS)), transposed=True)
    h = C.c_void_p()
    _cabi.check(lib.ocb_lu_program_create(C.byref(h), n, *[a.ctypes.data for a in arrs[:6]], flags), 'create')
    b = np.random.default_rng(3).standard_normal(n)
    ref = spsla.splu(K).solve(b)
    xs = []
    stats = (C.c_int64*4)()
    for mode in (0, 1):
        x = np.zeros(n)
        _cabi.check(lib.ocb_lu_program_solve_host(h, arrs[6].ctypes.data, arrs[7].ctypes.data,
                                                  b.ctypes.data, x.ctypes.data, mode, 0.0, stats), 'solve_host')
        assert np.linalg.norm(x - ref) <= 1e-12*np.linalg.norm(ref)
        xs.append(x)
    lib.ocb_lu_program_destroy(h)
    assert np.linalg.norm(xs[0] - xs[1]) <= 1e-13*np.linalg.norm(ref)
    npanels, stored, actual, nsub = [int(v) for v in stats]
    assert npanels > 0 and actual > 0
    assert stored <= 2.2*actual            # zero padding (union lists + rows up to 8) stays bounded
