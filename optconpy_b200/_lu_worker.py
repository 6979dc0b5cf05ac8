"""Host sparse-LU worker (setup step).  Imports numpy/scipy and the C library through
ctypes only, never torch and never initialises CUDA, so it can run in spawned worker
processes: SuperLU factorisation, then the host half of ``ocb_lu_create``
(``ocb_lu_pack_host``: supernodes, inverted diagonal blocks, packed batch stream)."""
import ctypes as C
import time

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla


def worker_init():
    """Pool initializer: the factorisations are throughput work - run them at a lower priority
    than the main process, whose threads drive the GPU and are latency sensitive (with as many
    busy workers as cores every host-side wait of the main thread otherwise costs milliseconds)."""
    import os
    try:
        os.nice(10)
    except OSError:
        pass
    tune_malloc()


def tune_malloc():
    """Keep large blocks on the heap instead of mmap/munmap per allocation: SuperLU and the
    program builder allocate and release tens of MB of work arrays per factorisation, and with
    glibc's default thresholds every one of them is a fresh mapping that page-faults in again
    (splu of the N=25 cavity saddle matrix: 44 -> 35 ms with this setting)."""
    try:
        libc = C.CDLL('libc.so.6')
        if not libc.mallopt(-3, 1 << 30):      # M_MMAP_THRESHOLD (older glibc caps it at 32 MB)
            libc.mallopt(-3, 32 << 20)
        libc.mallopt(-1, 1 << 30)      # M_TRIM_THRESHOLD
    except (OSError, AttributeError):
        pass


ORDER_TRIAL_MAX_N = 12000


def _delay_zero_diagonals(indptr, indices, diag_is_zero, q):
    """Constrained ordering for saddle-point patterns.  ``q[i]`` = node eliminated i-th.  A node
    with a ZERO diagonal (a pressure dof: the (2,2) block of ``[[A, J^T], [J, 0]]``) is delayed
    until right after the first of its neighbours with a nonzero diagonal has been eliminated -
    from then on its pivot is the Schur-complement entry ``-J a^-1 J^T != 0``, so the
    factorisation never meets an exact zero pivot, diagonal pivoting goes through and the fill is
    the SYMBOLIC fill of the ordering.  Without this, minimum degree eliminates pressure nodes
    first, SuperLU has to interchange rows, and with non-symmetric values (convection) the
    interchanges wreck the structure: 8.6x the entries and 3.3x the sub-levels on the channel
    Oseen matrix, 6.5x / 5.8x on the cavity one."""
    n = len(q)
    touched = np.zeros(n, dtype=bool)
    pending = np.zeros(n, dtype=bool)
    out = np.empty(n, dtype=np.int32)
    m = 0
    for v in q:
        if diag_is_zero[v]:
            if touched[v]:
                out[m] = v
                m += 1
            else:
                pending[v] = True
            continue
        out[m] = v
        m += 1
        nb = indices[indptr[v]:indptr[v+1]]
        z = nb[diag_is_zero[nb]]
        if z.size:
            rel = z[pending[z]]
            touched[z] = True
            if rel.size:
                rel = np.unique(rel)
                out[m:m+rel.size] = rel
                m += rel.size
                pending[rel] = False
    rest = np.flatnonzero(pending)          # zero-diagonal nodes without such a neighbour: last
    if rest.size:
        pos = np.empty(n, dtype=np.int64)
        pos[q] = np.arange(n)
        rest = rest[np.argsort(pos[rest])]
        out[m:m+rest.size] = rest
        m += rest.size
    assert m == n
    return out


def order_only(args):
    """The fill-reducing ordering of a sparsity pattern, computed ONCE per pattern, before any
    job of that pattern is queued, so that every queued factorisation takes the fast reuse path
    (symmetric permutation + NATURAL column order, ``factor_arrays``).

    Minimum degree on the PATTERN only - one SuperLU run on a diagonally dominant surrogate with
    the symmetrised pattern, so the ordering does not depend on the values of whichever matrix
    happens to come first - followed by the zero-diagonal constraint of
    ``_delay_zero_diagonals`` for the saddle-point block."""
    data, indices, indptr, shape, opts = args[:5]
    n = shape[0]
    K = sps.csc_matrix((data, indices, indptr), shape=shape)
    if opts.get('permc_spec', 'COLAMD') != 'MMD_AT_PLUS_A':
        slu = spsla.splu(K, **opts)
        return np.argsort(slu.perm_c).astype(np.int32)
    P = sps.csc_matrix((np.ones(K.nnz), K.indices, K.indptr), shape=shape)
    P = (P + P.T).tocsc()
    P.data[:] = 1.0
    import os
    if os.environ.get('OCB_ORDERING', 'nd') != 'mmd' and n > 0:
        # Nested dissection (ocb_order_nd): the solve kernels are bound by the number of dependent
        # sub-levels, i.e. by the height of the elimination tree.  Against minimum degree on the
        # cavity saddle matrices: N=25 90 -> 36 sub-levels with 14 % less fill, N=50 130 -> 66
        # with 24 % less.  Zero-diagonal (pressure) nodes are then delayed behind a neighbour as
        # for minimum degree.
        from optconpy_b200 import _cabi
        lib = _cabi.load()
        G = P.tocsr()
        G.setdiag(0.0)
        G.eliminate_zeros()
        G.sort_indices()
        gp = np.ascontiguousarray(G.indptr, dtype=np.int32)
        gi = np.ascontiguousarray(G.indices, dtype=np.int32)
        q = np.empty(n, dtype=np.int32)
        _cabi.check(lib.ocb_order_nd(n, gp.ctypes.data, gi.ctypes.data,
                                     int(os.environ.get('OCB_ND_LEAF', '4')), q.ctypes.data), 'ocb_order_nd')
        diag_is_zero = (K.diagonal() == 0.0)
        if diag_is_zero.any():
            # _delay_zero_diagonals in C++ (same result; 0.7 s -> 5 ms at n = 89 402)
            dz = np.ascontiguousarray(diag_is_zero, dtype=np.uint8)
            qc = np.empty(n, dtype=np.int32)
            _cabi.check(lib.ocb_order_delay_zero_diagonals(n, gp.ctypes.data, gi.ctypes.data, dz.ctypes.data,
                                                           q.ctypes.data, qc.ctypes.data),
                        'ocb_order_delay_zero_diagonals')
            q = qc
        return np.ascontiguousarray(q, dtype=np.int32)
    S = (P + sps.identity(n, format='csc')*float(2*P.getnnz(axis=0).max() + 1)).tocsc()
    slu = spsla.splu(S, permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0.0,
                     options=dict(SymmetricMode=True))
    q = np.argsort(slu.perm_c).astype(np.int32)
    diag_is_zero = (K.diagonal() == 0.0)
    if not diag_is_zero.any():
        return q
    Pr = P.tocsr()
    qc = _delay_zero_diagonals(Pr.indptr, Pr.indices, diag_is_zero, q)
    # Small systems whose values let plain minimum degree + threshold pivoting through (the
    # mass-dominated DRE matrices): keep that ordering - its factors have the SAME index arrays
    # for every shift and time step, which the structure templates of the program builder live
    # on, while the constrained ordering's factors differ in a handful of accidental zeros.
    # One trial factorisation of the actual matrix decides; it is only affordable when small.
    if n <= ORDER_TRIAL_MAX_N:
        fill_sym = slu.L.nnz + slu.U.nnz
        qo = np.argsort(spsla.splu(K, **opts).perm_c).astype(np.int32)
        o2 = dict(opts, permc_spec='NATURAL')          # the reuse path of factor_arrays
        trial = spsla.splu(K[qo][:, qo].tocsc(), **o2)
        if trial.L.nnz + trial.U.nnz <= 1.15*fill_sym:
            return qo
    return qc


_ARRANGE = dict()     # pattern -> (indices, indptr, source position of every entry) or None


def _arranged(data, indices, indptr, shape, q, transposed):
    """``A`` (CSC arrays), optionally transposed, then optionally permuted symmetrically by
    ``q``, as a canonical CSC matrix.  The matrices of one run share a handful of sparsity
    patterns, so the transposition and the fancy indexing (8 of 50 ms per matrix) are done once
    per pattern on entry TAGS; every later matrix is one gather of its values."""
    import zlib
    ib, pb = indices.tobytes(), indptr.tobytes()
    key = (shape, bool(transposed), len(ib), hash(ib), zlib.crc32(ib), hash(pb),
           None if q is None else hash(q.tobytes()))
    if key not in _ARRANGE:
        nnz = len(indices)
        tags = sps.csc_matrix((np.arange(1, nnz + 1, dtype=np.float64), indices, indptr), shape=shape)
        m = tags.T.tocsc() if transposed else tags
        if q is not None:
            m = m[q][:, q].tocsc()
        m.sum_duplicates()
        src = np.rint(m.data).astype(np.int64) - 1
        ok = m.nnz == nnz and np.array_equal(np.sort(src), np.arange(nnz))   # no duplicate entries
        if len(_ARRANGE) >= 16:
            _ARRANGE.clear()
        _ARRANGE[key] = (m.indices.copy(), m.indptr.copy(), src) if ok else None
    ent = _ARRANGE[key]
    if ent is None:
        mat = sps.csc_matrix((data, indices, indptr), shape=shape)
        if transposed:
            mat = mat.T.tocsc()
        if q is not None:
            mat = mat[q][:, q].tocsc()
        return mat
    mat = sps.csc_matrix((data[ent[2]], ent[0], ent[1]), shape=shape)
    mat.has_sorted_indices = True
    mat.has_canonical_format = True      # splu's sum_duplicates() becomes a no-op
    return mat


def factor_arrays(args, want_order=False, transposed=False):
    """(data, indices, indptr, shape, lu_options[, smem, flags, q]) of a CSC matrix ->
    int32/FP64 CSR arrays of the lower and the upper factor plus the two permutations, in the
    layout ``ocb_lu_pack_host`` expects.

    ``q`` (optional, args[7]): a fill-reducing ordering obtained from an earlier factorisation
    of a matrix with the SAME sparsity pattern.  The matrix is then permuted symmetrically,
    ``B = A[q][:, q]``, and factorised with the NATURAL column order: SuperLU skips its
    minimum-degree ordering (half of its run time here) and, because the permutation is
    symmetric, pivots on the diagonal more often (25 % fewer entries in L+U on the cavity
    matrices).  The returned permutations are composed so that they refer to ``A`` again.

    ``transposed=True`` (flags bit 1 of ``ocb_lu_pack_host``): factorise ``A^T`` instead.
    SuperLU hands its factors out column-wise; the columns of ``U`` are the rows of the lower
    factor ``U^T`` and the columns of ``L`` the rows of the upper factor ``L^T`` in
    ``A = U^T L^T``, so no CSC->CSR conversion (19 of 80 ms per matrix) is needed.  With
    ``Pr A^T Pc = L U``:  ``U^T L^T (Pr x) = Pc^T b``, i.e. the kernel's load permutation is this
    factorisation's ``perm_c`` and its store permutation ``perm_r``."""
    data, indices, indptr, shape, opts = args[:5]
    q = args[7] if len(args) > 7 else None
    n = shape[0]
    o2 = dict(opts)
    if q is not None:
        o2['permc_spec'] = 'NATURAL'
    mat = _arranged(data, indices, indptr, shape, q, transposed)
    slu = spsla.splu(mat, **o2)
    load_p, store_p = (slu.perm_c, slu.perm_r) if transposed else (slu.perm_r, slu.perm_c)
    if q is None:
        perm_r, perm_c = load_p, store_p
    else:
        perm_r = np.empty(n, dtype=np.int32)
        perm_c = np.empty(n, dtype=np.int32)
        perm_r[q] = load_p            # xe[perm_r[i]] = b[i]   with b' = b[q]
        perm_c[q] = store_p           # x[j] = xe[perm_c[j]]   with x[q] = y
    if transposed:
        lo, up = slu.U, slu.L         # CSC columns of U / L = CSR rows of U^T / L^T (rows of the
        #                               upper factor unsorted: build_lu_program sorts its own copy)
    else:
        lo = sps.csr_matrix(slu.L)
        up = sps.csr_matrix(slu.U)
        lo.sort_indices()
        up.sort_indices()
    out = [np.ascontiguousarray(lo.indptr, dtype=np.int32),
           np.ascontiguousarray(lo.indices, dtype=np.int32),
           np.ascontiguousarray(lo.data, dtype=np.float64),
           np.ascontiguousarray(up.indptr, dtype=np.int32),
           np.ascontiguousarray(up.indices, dtype=np.int32),
           np.ascontiguousarray(up.data, dtype=np.float64),
           np.ascontiguousarray(perm_r, dtype=np.int32),
           np.ascontiguousarray(perm_c, dtype=np.int32)]
    if want_order:
        # ordering for later matrices of this pattern (None if this run already used one)
        return out, (None if q is not None else np.argsort(slu.perm_c).astype(np.int32))
    return out


# Residual guard: the host LU runs with relaxed pivoting (diag_pivot_thresh=0.01,
# symmetric mode) and the device solve multiplies by explicit inverses of the supernode blocks;
# the reference's ``spsla.factorized`` does neither.  Every factor image is therefore checked
# where it is built: the finished gather program is executed on the host for one right-hand
# side and the normwise backward error against the ORIGINAL matrix must stay below GUARD_TOL,
# else the matrix is factorised again with full partial pivoting and narrow supernodes.
GUARD_TOL = 2e-15
SAFE_FLAG = 8          # ocb_lu_pack_host flags bit 3: supernodes <= 32 rows, no one-step blocks


def _guard_tol():
    import os
    v = os.environ.get('OCB_LU_GUARD_TOL')
    return GUARD_TOL if v is None else float(v)      # <= 0 switches the guard off


class _CImage(object):
    """Device image in a C buffer (``ocb_lu_pack_host_checked``); ``view`` is a uint8 numpy
    view, ``free()`` releases it; ``backerr`` is the guard's backward error (None: not checked)."""

    def __init__(self, arrs, n, smem_optin, flags=0, amat=None):
        from optconpy_b200 import _cabi
        self._lib = _cabi.load()
        img, nbytes = C.c_void_p(), C.c_int64(0)
        self.backerr = None
        if amat is None:
            _cabi.check(self._lib.ocb_lu_pack_host(n, *[a.ctypes.data for a in arrs], int(smem_optin),
                                                   int(flags), C.byref(img), C.byref(nbytes)),
                        'ocb_lu_pack_host')
        else:
            be = C.c_double(0.0)
            _cabi.check(self._lib.ocb_lu_pack_host_checked(
                n, *[a.ctypes.data for a in arrs], int(smem_optin), int(flags), None, 0,
                C.byref(img), C.byref(nbytes), *[a.ctypes.data for a in amat], C.byref(be)),
                'ocb_lu_pack_host_checked')
            self.backerr = be.value
        self._ptr = img
        self.view = np.ctypeslib.as_array(C.cast(img, C.POINTER(C.c_uint8)), shape=(nbytes.value,))

    def free(self):
        if self._ptr is not None:
            self.view = None
            self._lib.ocb_host_free(self._ptr)
            self._ptr = None


def pack_image(arrs, n, smem_optin, flags=0):
    """Host half of ``ocb_lu_create``: returns the device image as a uint8 array."""
    ci = _CImage(arrs, n, smem_optin, flags)
    try:
        return ci.view.copy()
    finally:
        ci.free()


def _amat_arrays(args):
    """int32/FP64 CSC arrays of the ORIGINAL matrix for the residual guard."""
    data, indices, indptr = args[:3]
    return (np.ascontiguousarray(indptr, dtype=np.int32), np.ascontiguousarray(indices, dtype=np.int32),
            np.ascontiguousarray(data, dtype=np.float64))


def _safe_args(args):
    """The same job with full partial pivoting (what ``spsla.factorized`` does) and the safe
    program layout; keeps the cached ordering if there is one."""
    opts = dict(args[4])
    opts['diag_pivot_thresh'] = 1.0
    opts['options'] = dict(opts.get('options', {}), SymmetricMode=False)
    flags = (args[6] if len(args) > 6 else 0) | SAFE_FLAG
    return args[:4] + (opts, args[5], flags) + tuple(args[7:])


class _Refactor(object):
    """Numeric-only factorisation of one sparsity pattern with the pivot order of a first SuperLU
    run (``ocb_refactor_*``, csrc/refactor.cpp): the index arrays of both factors and the
    permutations are fixed, ``numeric(data)`` fills in the numbers."""

    def __init__(self, n, indptr, indices, perm_r, perm_c):
        from optconpy_b200 import _cabi
        self._lib = lib = _cabi.load()
        self.n = n
        ip = np.ascontiguousarray(indptr, dtype=np.int32)
        ii = np.ascontiguousarray(indices, dtype=np.int32)
        h = C.c_void_p()
        _cabi.check(lib.ocb_refactor_create(C.byref(h), n, ip.ctypes.data, ii.ctypes.data,
                                            perm_r.ctypes.data, perm_c.ctypes.data), 'ocb_refactor_create')
        self._h = h
        info = (C.c_int64*8)()
        lib.ocb_refactor_info(h, info)
        self.info = dict(nnzL=int(info[1]), nnzU=int(info[2]), supernodes=int(info[3]),
                         max_front=int(info[4]), flops=int(info[6]))
        self.Lrp, self.Urp = np.empty(n+1, np.int32), np.empty(n+1, np.int32)
        self.Lci, self.Uci = np.empty(info[1], np.int32), np.empty(info[2], np.int32)
        self.perm_r, self.perm_c = np.empty(n, np.int32), np.empty(n, np.int32)
        lib.ocb_refactor_structure(h, self.Lrp.ctypes.data, self.Lci.ctypes.data, self.Urp.ctypes.data,
                                   self.Uci.ctypes.data, self.perm_r.ctypes.data, self.perm_c.ctypes.data)
        self.Lva, self.Uva = np.empty(info[1]), np.empty(info[2])

    def numeric(self, data):
        """The eight arrays ``ocb_lu_pack_host`` takes (layout: P A Q = L U, flags bit 1 clear),
        or None if a static pivot vanished."""
        d = np.ascontiguousarray(data, dtype=np.float64)
        rc = self._lib.ocb_refactor_numeric(self._h, d.ctypes.data, self.Lva.ctypes.data, self.Uva.ctypes.data)
        if rc != 0:
            return None
        return [self.Lrp, self.Lci, self.Lva, self.Urp, self.Uci, self.Uva, self.perm_r, self.perm_c]

    def __del__(self):
        try:
            if self._h:
                self._lib.ocb_refactor_destroy(self._h)
                self._h = None
        except Exception:
            pass


# Static-pivot refactorisation: OCB_REFACTOR=0 switches it off.  Handles by
# sparsity pattern + pivots, at most four (one run has one or two patterns).
_REFAC = dict()
REFACTOR_MAX_PATTERNS = 4


def refactor_wanted(flags):
    """Static pivots are used with the default (transposed) SuperLU layout, unless OCB_REFACTOR=0."""
    import os
    return os.environ.get('OCB_REFACTOR', '1') != '0' and bool(flags & 2) and not (flags & SAFE_FLAG)


def static_pivots(args):
    """The two permutations every later matrix of this pattern is factorised with: those of ONE
    SuperLU run on the given matrix (ordering ``args[7]`` reused).  Computed once per pattern in
    the main process and handed to every job, so that which worker runs which job - or whether
    the look-ahead is on - cannot change a single bit of the results."""
    arrs = factor_arrays(args[:8], transposed=True)
    return arrs[6], arrs[7]


def _refactor_enabled(args):
    flags = args[6] if len(args) > 6 else 0
    return (len(args) > 8 and args[8] is not None and args[7] is not None and refactor_wanted(flags)
            and args[3][0] > 0)


def _refactor_key(args):
    import zlib
    return (args[3], len(args[1]), zlib.crc32(memoryview(np.ascontiguousarray(args[1]))),
            zlib.crc32(memoryview(np.ascontiguousarray(args[2]))),
            zlib.crc32(memoryview(np.ascontiguousarray(args[8][0]))),
            zlib.crc32(memoryview(np.ascontiguousarray(args[8][1]))))


def _pack(lib, arrs, n, smem, flags, slot, amat):
    """Analyse + pack one factorisation.  Returns (slot or None, image or None, nbytes, backerr)."""
    from optconpy_b200 import _cabi
    img, backerr, nbytes = None, None, 0
    if slot is not None:
        # build the image right in the pinned segment (no intermediate buffer, no copy)
        seg = _attach(slot[0])
        if slot[0] not in _ADDRESS:    # one exported view per segment, kept for the process lifetime
            _ADDRESS[slot[0]] = C.addressof(C.c_char.from_buffer(seg.buf))
        nb = C.c_int64(0)
        if amat is None:
            rc = lib.ocb_lu_pack_host_into(n, *[a.ctypes.data for a in arrs], int(smem), int(flags),
                                           _ADDRESS[slot[0]], int(slot[1]), C.byref(nb))
        else:
            be = C.c_double(0.0)
            rc = lib.ocb_lu_pack_host_checked(n, *[a.ctypes.data for a in arrs], int(smem),
                                              int(flags), _ADDRESS[slot[0]], int(slot[1]), None,
                                              C.byref(nb), *[a.ctypes.data for a in amat], C.byref(be))
            backerr = be.value
        if rc == 0:
            nbytes = nb.value
        elif rc != -5:                    # anything but "does not fit": a real error
            _cabi.check(rc, 'ocb_lu_pack_host_checked')
        else:
            slot = None                   # too small: hand the image back another way
    if slot is None:
        ci = _CImage(arrs, n, smem, flags, amat=amat)
        img, backerr, nbytes = ci.view.copy(), ci.backerr, ci.view.nbytes
        ci.free()
    return slot, img, nbytes, backerr


def _build(args, slot=None):
    """Factorise + analyse + pack (guarded).  Returns (name-or-None, image-or-None, nbytes,
    seconds factor, seconds pack, ordering, guard) where guard = (backward error of the image
    handed out, 1 if it is the safe re-factorisation else 0, 'static' if the numbers come from
    the numeric-only refactorisation with static pivots else 'slu', 1 if a static-pivot image
    was rejected by the guard first else 0)."""
    from optconpy_b200 import _cabi
    lib = _cabi.load()
    tol = _guard_tol()
    amat = _amat_arrays(args) if tol > 0 else None
    tf = tp = 0.0
    order = None
    safe = 0
    rejected = 0
    n = args[3][0]
    rkey = _refactor_key(args) if _refactor_enabled(args) else None
    rf = _REFAC.get(rkey) if rkey is not None else None
    if rkey is not None and rkey not in _REFAC:
        # first job of this pattern in this process: symbolic analysis for the given pivots
        t0 = time.perf_counter()
        if len(_REFAC) >= REFACTOR_MAX_PATTERNS:
            _REFAC.pop(next(iter(_REFAC)))
        try:
            rf = _Refactor(n, args[2], args[1], np.ascontiguousarray(args[8][0], dtype=np.int32),
                           np.ascontiguousarray(args[8][1], dtype=np.int32))
        except RuntimeError:
            rf = None
        _REFAC[rkey] = rf
        tf += time.perf_counter() - t0
    if rf is not None:
        # numbers only, pivots as given.  The guard is what makes static pivots safe, so it
        # always runs here.
        t0 = time.perf_counter()
        arrs = rf.numeric(args[0])
        t1 = time.perf_counter()
        tf += t1 - t0
        if arrs is not None:
            am = amat if amat is not None else _amat_arrays(args)
            used, img, nbytes, backerr = _pack(lib, arrs, n, args[5], args[6] & ~2, slot, am)
            tp += time.perf_counter() - t1
            if backerr <= (tol if tol > 0 else GUARD_TOL):
                return used, img, nbytes, tf, tp, None, (backerr, 0, 'static', 0)
        rejected = 1
    while True:
        t0 = time.perf_counter()
        flags = args[6] if len(args) > 6 else 0
        arrs, o2 = factor_arrays(args, want_order=True, transposed=bool(flags & 2))
        order = o2 if order is None else order
        t1 = time.perf_counter()
        tf += t1 - t0
        used, img, nbytes, backerr = _pack(lib, arrs, n, args[5], flags, slot, amat)
        tp += time.perf_counter() - t1
        if backerr is None or backerr <= tol or safe:
            return used, img, nbytes, tf, tp, order, (backerr, safe, 'slu', rejected)
        args = _safe_args(args)
        safe = 1


def factor_image(args):
    """(data, indices, indptr, shape, lu_options, smem_optin[, flags, q, (perm_r, perm_c)]) ->
    (image, seconds factor, seconds analyse+pack, ordering, guard)."""
    _, img, _, tf, tp, order, guard = _build(args)
    return img, tf, tp, order, guard


_ATTACHED = dict()
_ADDRESS = dict()


def _attach(name):
    """Attach (once per worker) to a pinned segment of the main process's pool."""
    seg = _ATTACHED.get(name)
    if seg is None:
        from multiprocessing import shared_memory
        seg = shared_memory.SharedMemory(name=name)   # spawned workers share the parent's resource tracker
        _ATTACHED[name] = seg
    return seg


def _timeline(t0, tf, tp):
    """OCB_TIMELINE=<dir>: one line per job (wall-clock start, end, seconds factor / pack) in
    <dir>/worker_<pid>.log, merged with the main process's events by tools/e2e_timeline.py."""
    import os
    d = os.environ.get('OCB_TIMELINE')
    if d:
        with open(os.path.join(d, 'worker_%d.log' % os.getpid()), 'a') as f:
            f.write('%.6f %.6f %.6f %.6f\n' % (t0, time.time(), tf, tp))


def factor_image_to_shm(args, slot=None):
    """Pool entry point: factorise, analyse, pack and hand the image back through POSIX shared
    memory (no pickling of ~10 MB through a pipe).  ``slot = (name, capacity)`` is a segment of
    the main process's page-locked pool: if the image fits it is written there (the upload is
    then a plain DMA); otherwise a fresh segment is created.
    Returns (name or None if the slot was used, nbytes, seconds factor, seconds analyse+pack,
    ordering for later matrices of the same pattern or None, guard)."""
    from multiprocessing import shared_memory
    t0 = time.time()
    used, img, nbytes, tf, tp, order, guard = _build(args, slot)
    _timeline(t0, tf, tp)
    if used is not None:
        return None, nbytes, tf, tp, order, guard
    shm = shared_memory.SharedMemory(create=True, size=max(nbytes, 64))
    np.frombuffer(shm.buf, dtype=np.uint8, count=nbytes)[:] = img
    name = shm.name
    shm.close()
    return name, nbytes, tf, tp, order, guard
