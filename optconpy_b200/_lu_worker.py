"""Host sparse-LU worker (setup step).  Imports numpy/scipy only, so it can run in
spawned worker processes without touching CUDA or torch."""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla


def factor_arrays(args):
    """(data, indices, indptr, shape, lu_options) of a CSC matrix ->
    int32/FP64 CSR arrays of L and U plus the two permutations."""
    data, indices, indptr, shape, opts = args
    mat = sps.csc_matrix((data, indices, indptr), shape=shape)
    slu = spsla.splu(mat, **opts)
    L = sps.csr_matrix(slu.L)
    U = sps.csr_matrix(slu.U)
    L.sort_indices()
    U.sort_indices()
    return [np.ascontiguousarray(L.indptr, dtype=np.int32),
            np.ascontiguousarray(L.indices, dtype=np.int32),
            np.ascontiguousarray(L.data, dtype=np.float64),
            np.ascontiguousarray(U.indptr, dtype=np.int32),
            np.ascontiguousarray(U.indices, dtype=np.int32),
            np.ascontiguousarray(U.data, dtype=np.float64),
            np.ascontiguousarray(slu.perm_r, dtype=np.int32),
            np.ascontiguousarray(slu.perm_c, dtype=np.int32)]


def factor_to_shm(args):
    """Pool entry point: factorise and hand the arrays back through one POSIX shared-memory
    block (avoids pickling ~13 MB per factor through a pipe).  Returns (name, layout)."""
    from multiprocessing import shared_memory
    arrs = factor_arrays(args)
    layout, off = [], 0
    for a in arrs:
        off = (off + 63) & ~63
        layout.append((a.dtype.str, a.size, off))
        off += a.nbytes
    shm = shared_memory.SharedMemory(create=True, size=max(off, 64))
    for a, (_, _, o) in zip(arrs, layout):
        np.frombuffer(shm.buf, dtype=a.dtype, count=a.size, offset=o)[:] = a
    name = shm.name
    shm.close()
    return name, layout
