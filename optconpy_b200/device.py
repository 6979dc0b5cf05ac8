"""Device-side carriers for the CUDA path: torch tensors own the HBM buffers,
ctypes hands their raw pointers to ``liboptconpy_b200.so``.

There is no CPU fallback: every entry point calls :func:`require_cuda` and the
library loader raises if the extension is missing.
"""
import ctypes as C
import os
import threading
import time

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla
import torch

from . import _cabi

# Counters the bench reads (setup is timed separately, north_star).
STATS = dict(lu_factor_s=0.0,          # SuperLU seconds summed over the workers
             lu_worker_pack_s=0.0,     # host analysis + packing seconds summed over the workers
             lu_order_s=0.0,           # main process: one-off minimum-degree ordering per pattern
             lu_submit_s=0.0,          # main process: building + sending the CSC arguments
             lu_wait_s=0.0,            # main thread: blocked until the factors are on the device
             lu_collect_wait_s=0.0,    # collecting thread: blocked on a worker result
             lu_analyse_upload_s=0.0,  # main process: image upload + handle creation
             lu_arena_s=0.0,           # ... of which: device buffer from the caching allocator
             lu_unpinned_uploads=0,    # images that did not travel through a pinned pool segment
             lu_guard_refactors=0,     # residual guard: matrices factorised again in safe mode
             lu_static_pivot=0,         # ... images whose numbers came from the numeric-only refactorisation
             lu_static_rejected=0,      # ... static-pivot images the guard rejected (SuperLU took over)
             lu_guard_max_backerr=0.0,  # ... largest backward error of an image handed out
             n_factor=0, h2d_bytes=0, d2h_bytes=0)

# wall seconds of the main thread per phase of the host API (diagnostics, bench.py e2e)
PHASE = dict()


# OCB_TIMELINE=<dir>: (thread, label, wall start, wall end) of the host-side events of a run
# (diagnostics; tools/e2e_timeline.py merges them with the workers' logs)
TIMELINE = [] if os.environ.get('OCB_TIMELINE') else None


def timeline(label, t0, t1=None):
    if TIMELINE is not None:
        TIMELINE.append((threading.current_thread().name, label, t0, time.time() if t1 is None else t1))


class phase(object):
    """``with phase('name'):`` adds the wall time of the block to PHASE['name']."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.t0 = time.perf_counter()
        self.w0 = time.time()

    def __exit__(self, *a):
        PHASE[self.name] = PHASE.get(self.name, 0.0) + time.perf_counter() - self.t0
        timeline(self.name, self.w0)
        return False

# Host factorisation used for the (separately timed) setup step.  SuperLU with a
# symmetric-pattern ordering: the saddle-point matrices have symmetric structure,
# for which minimum degree on A^T+A gives ~25 % less fill and ~2.4x fewer
# dependency levels than the COLAMD default (measured on the N=25 cavity);
# solutions agree with the default-ordering factorisation to ~5e-15 relative.
# relax / panel_size: SuperLU's supernode relaxation and panel width; 4 / 10 factorises these
# matrices ~35 % faster than the library defaults (same fill, same pivots).
LU_OPTIONS = dict(permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0.01,
                  options=dict(SymmetricMode=True), relax=4, panel_size=10)


# How threads of this process wait for the device.  Spinning (the CUDA default while there are
# more cores than contexts) gives the lowest latency and is right for the device-resident loop;
# but with one process per GPU and few host cores per process - 8 GPUs on a 16-core host - the
# spinning main threads take the cores the LU workers need (e2e at 8 GPUs: 55 -> 88 steps/s with
# sleeping waits, while the device-resident loop loses 20 % to the wake-up latency).  So the
# waits sleep exactly while host factorisations are in flight.  OCB_BLOCKING_SYNC=0 / 1: never /
# always while jobs are in flight; default: when a rank has fewer than 6 cores.
_SYNC = dict(inflight=0, blocking=False, enabled=None)


def _jobs_inflight(delta):
    with _LOCK:
        if _SYNC['enabled'] is None:
            world = max(int(os.environ.get('WORLD_SIZE', '1')), 1)
            want = os.environ.get('OCB_BLOCKING_SYNC')
            _SYNC['enabled'] = (want == '1') or (want is None and (os.cpu_count() or 1) < 6*world)
        _SYNC['inflight'] = max(0, _SYNC['inflight'] + delta)
        want = _SYNC['enabled'] and _SYNC['inflight'] > 0
        if want != _SYNC['blocking']:
            _cabi.check(_cabi.load().ocb_set_sync_mode(1 if want else 0), 'ocb_set_sync_mode')
            _SYNC['blocking'] = want


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('optconpy_b200: no CUDA device visible; this package has no '
                           'CPU path (the CPU oracle lives in oracle/ and is test-only)')
    return _cabi.load()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def cur_device():
    return torch.device('cuda', torch.cuda.current_device())


def reset_stats():
    for k in STATS:
        STATS[k] = 0 if isinstance(STATS[k], int) else 0.0
    PHASE.clear()


# Device copies of host arrays this package handed out (the compressed factor Zc): the reference
# driver passes the very same array object straight back in (get_mTzzTtb twice, z0 of the next
# time step: solve_dae_ric.py:152-189), so the re-upload can be skipped.  Keyed by object identity,
# validated by shape, the sum of all entries and a strided sample (a caller that modified the array
# in place gets a fresh upload).
_DEV_CACHE = dict()
_DEV_CACHE_MAX = 4


def _fingerprint(a):
    flat = a.reshape(-1)
    step = max(1, flat.size//61)
    return (a.shape, a.dtype.str, float(a.sum()), flat[::step][:64].tobytes())


def remember_device_copy(host_arr, dev_tensor):
    """Register ``dev_tensor`` as the device copy of ``host_arr`` (a fresh array we return)."""
    import weakref
    if not isinstance(host_arr, np.ndarray) or host_arr.size == 0:
        return
    try:
        ref = weakref.ref(host_arr)
    except TypeError:
        return
    fp = _fingerprint(host_arr)
    with _LOCK:
        if len(_DEV_CACHE) >= _DEV_CACHE_MAX:
            _DEV_CACHE.pop(next(iter(_DEV_CACHE)))
        _DEV_CACHE[id(host_arr)] = (ref, fp, dev_tensor)


def _cached_device_copy(arr):
    with _LOCK:
        ent = _DEV_CACHE.get(id(arr))
    if ent is None:
        return None
    ref, fp, t = ent
    if ref() is not arr or t.device != cur_device() or fp != _fingerprint(arr):
        with _LOCK:
            _DEV_CACHE.pop(id(arr), None)
        return None
    return t


def to_dev(arr):
    """host array (dense or sparse) -> contiguous FP64 2-D device tensor."""
    if isinstance(arr, torch.Tensor):
        t = arr.to(device=cur_device(), dtype=torch.float64)
        return t if t.is_contiguous() else t.contiguous()
    if isinstance(arr, np.ndarray) and arr.ndim == 2 and arr.dtype == np.float64:
        t = _cached_device_copy(arr)
        if t is not None:
            return t
    if sps.issparse(arr):
        arr = arr.toarray()
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))
    if a.ndim == 1:
        a = a[:, None]
    STATS['h2d_bytes'] += a.nbytes
    return torch.from_numpy(a).to(cur_device())


_PINNED_STAGE = dict()      # per host thread (the stepper's tail thread copies results out too)


def to_host(t):
    """device tensor -> fresh numpy array.  Large blocks travel through ONE persistent, grow-only
    pinned staging buffer (DMA at PCIe speed) and are then copied into the array that is handed
    out: a fresh pinned allocation per call (cudaHostAlloc of 64 MB: 20-50 ms, and it
    synchronises the device) made the 45 MB ADI factor cost 57 ms per call through the plain
    reference signatures; a pageable ``.cpu()`` runs at ~2 GB/s."""
    nbytes = t.numel()*t.element_size()
    STATS['d2h_bytes'] += nbytes
    if nbytes < (1 << 20):
        return t.cpu().numpy()
    tc = t if t.is_contiguous() else t.contiguous()
    numel = tc.numel()
    me = threading.get_ident()
    st = _PINNED_STAGE.get(me)
    if st is None or st.numel() < numel or st.dtype != tc.dtype:
        cap = 1 << max(int(numel - 1).bit_length(), 20)
        st = torch.empty(cap, dtype=tc.dtype, pin_memory=True)
        _PINNED_STAGE[me] = st
    host = st[:numel].view(tc.shape)
    host.copy_(tc, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy().copy()


def ptr(t):
    return 0 if t is None else t.data_ptr()


class Workspace(object):
    """One growing byte buffer per purpose; avoids an allocation per call."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes):
        nbytes = int(max(nbytes, 256))
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != cur_device():
            t0 = time.perf_counter()
            self.buf = None
            self.buf = torch.empty(int(nbytes*1.5) + 1024, dtype=torch.uint8,
                                   device=cur_device())
            if os.environ.get('OCB_DEBUG_WS'):
                torch.cuda.synchronize()
                import sys
                sys.stderr.write('workspace grow to %.1f MB took %.1f ms\n'
                                 % (self.buf.numel()/1e6, 1e3*(time.perf_counter() - t0)))
        return self.buf


_WS = dict()
_POOL_RESERVED = dict(bytes=0)


def reserve_pool(nbytes):
    """Make torch's caching allocator hold ``nbytes`` of free device memory (allocate once,
    release to the cache): the blocks whose size changes from call to call - the ADI factor
    grows with the block width every time step - are then carved out of that cached block
    instead of triggering a cudaMalloc (20-100 ms on a busy box, and it synchronises)."""
    nbytes = int(nbytes)
    if nbytes > _POOL_RESERVED['bytes']:
        t = torch.empty(nbytes, dtype=torch.uint8, device=cur_device())
        del t
        _POOL_RESERVED['bytes'] = nbytes


def workspace(name, nbytes):
    """The growing buffer of purpose ``name`` of the CALLING host thread (two threads of one
    process may drive the device at the same time, on different streams)."""
    ws = _WS.setdefault((name, threading.get_ident()), Workspace())
    return ws.get(nbytes)


class DeviceCSR(object):
    """CSR matrix on the device (int32 indices, FP64 values)."""

    def __init__(self, mat):
        m = sps.csr_matrix(mat, dtype=np.float64)
        m.sum_duplicates()
        m.sort_indices()
        if m.nnz >= 2**31 - 1:
            raise ValueError('nnz does not fit int32')
        self.shape = m.shape
        self.nnz = m.nnz
        d = cur_device()
        self.rowptr = torch.from_numpy(m.indptr.astype(np.int32)).to(d)
        self.colidx = torch.from_numpy(m.indices.astype(np.int32)).to(d)
        self.vals = torch.from_numpy(m.data.astype(np.float64)).to(d)
        STATS['h2d_bytes'] += 4*(m.shape[0]+1) + 12*m.nnz

    def matmul(self, X, alpha=1.0, beta=0.0, out=None):
        """``alpha*S@X (+ beta*out)`` for a device block X (ncols x k)."""
        lib = require_cuda()
        assert X.dtype == torch.float64 and X.dim() == 2 and X.stride(1) == 1
        assert X.shape[0] == self.shape[1]
        k = X.shape[1]
        if out is None:
            out = torch.empty((self.shape[0], k), dtype=torch.float64, device=X.device)
            beta = 0.0
        _cabi.check(lib.ocb_spmm(self.shape[0], self.shape[1], ptr(self.rowptr),
                                 ptr(self.colidx), ptr(self.vals), ptr(X), X.stride(0),
                                 ptr(out), out.stride(0), k, alpha, beta, stream_ptr()),
                    'ocb_spmm')
        return out


_POOL = dict(pool=None, workers=0)
# the look-ahead thread of the DRE stepper and the main thread both create jobs: every lazily
# created singleton below (worker pool, pinned pool, upload thread, ordering cache) is built
# under this lock, so that two threads never spawn (and leak) a second copy
_LOCK = threading.RLock()


def _lu_pool():
    """Process pool for the host LU setup (SuperLU holds the GIL).  OCB_LU_WORKERS=0
    keeps everything in-process."""
    import multiprocessing as mp
    want = os.environ.get('OCB_LU_WORKERS')
    if want is None:
        world = int(os.environ.get('WORLD_SIZE', '1'))
        want = max(2, min(14, ((os.cpu_count() or 1) - 2)//max(world, 1)))
    want = int(want)
    if want <= 1:
        return None
    with _LOCK:
        if _POOL['pool'] is None or _POOL['workers'] != want:
            if _POOL['pool'] is not None:
                _POOL['pool'].terminate()
            from . import _lu_worker
            _POOL['pool'] = mp.get_context('spawn').Pool(want, initializer=_lu_worker.worker_init)
            _POOL['workers'] = want
        return _POOL['pool']


def _record_guard(guard):
    """Book-keeping of the workers' residual guard (``_lu_worker._build``)."""
    if guard is None or guard[0] is None:
        return
    backerr, safe = guard[:2]
    if len(guard) > 2:
        STATS['lu_static_pivot'] += int(guard[2] == 'static')
        STATS['lu_static_rejected'] += int(guard[3])
    from . import _lu_worker
    STATS['lu_guard_refactors'] += int(safe)
    STATS['lu_guard_max_backerr'] = max(STATS['lu_guard_max_backerr'], float(backerr))
    if safe and backerr > _lu_worker._guard_tol():
        import warnings
        warnings.warn('optconpy_b200: backward error %.1e of a saddle-point factorisation is above '
                      'the guard tolerance even with full partial pivoting (matrix close to '
                      'singular?)' % backerr, RuntimeWarning)


def _csc_args(mat, opts):
    m = sps.csc_matrix(mat, dtype=np.float64)
    m.sum_duplicates()
    return (m.data, m.indices, m.indptr, m.shape, opts)


# fill-reducing orderings by sparsity pattern: the first factorisation of a pattern runs
# SuperLU's minimum-degree ordering, all later ones (other shifts, other time steps) reuse it
_ORDER = dict()


# one-step (inverse-multiplied) supernodes in the gather program, flags bit 2 of
# ocb_lu_pack_host (lu_program.h); OCB_MERGE=0/1 overrides
MERGE_DEFAULT = '1'


def _pack_flags(wide, k_hint=None):
    """ocb_lu_pack_host flags: bit 0 = flat program for the wide executor, bit 1 = factorise A^T
    and take SuperLU's column-wise factors as they are (no CSC->CSR conversion on the host),
    bit 2 = small supernodes solved in one sub-level instead of two,
    bits 4..7 = cluster size of the column-panel kernel from the expected block width
    (measured on the N=25 factor, profiles/r02f_cluster_size_sweep.log + r02f_kp2_sweep.log: one
    column per cluster while one wave of clusters covers the block, two columns - one 16-byte gather
    per entry - beyond: clusters of 4 take 85 us up to 33 right-hand sides and 118-123 us up to 66,
    clusters of 3 100 us up to 45 and 138-141 us up to 90, clusters of 2 128 us up to 74 and 171 us
    up to 148)."""
    if k_hint is None:
        cl = 0
    elif k_hint <= 33:
        cl = 4
    elif k_hint <= 44:
        cl = 3
    elif k_hint <= 60:      # (clusters of 4 fall off a cliff beyond 66 columns - 232 us -: the hint
        cl = 4              # may lag the actual width by a few columns)
    elif k_hint <= 74:
        cl = 2
    elif k_hint <= 90:
        cl = 3
    else:
        cl = 2
    merge = 4 if os.environ.get('OCB_MERGE', MERGE_DEFAULT) == '1' else 0
    return (1 if wide else 0) | (0 if os.environ.get('OCB_NO_TRANSPOSED_LU') else 2) | merge | (cl << 4)


def _pattern_key(a):
    import zlib
    return (a[3], len(a[1]), zlib.crc32(a[1].tobytes()), zlib.crc32(a[2].tobytes()))


def _with_order(a, opts):
    """worker arguments (.., smem, flags, q, pivots) with the cached ordering and the static
    pivots (``_lu_worker.static_pivots``; None when switched off) of this pattern"""
    if os.environ.get('OCB_NO_ORDER_REUSE') or opts.get('permc_spec') == 'NATURAL':
        return a, None
    key = _pattern_key(a)
    with _LOCK:
        if key not in _ORDER:
            # first matrix of this pattern: get the ordering NOW (one synchronous SuperLU run), so
            # that no queued job runs the slow path or produces the larger factor; a second run
            # fixes the pivot order of the numeric-only refactorisation
            from . import _lu_worker
            t0 = time.perf_counter()
            q = _lu_worker.order_only(a)
            piv = _lu_worker.static_pivots(a + (q,)) if _lu_worker.refactor_wanted(a[6]) else None
            _ORDER[key] = (q, piv)
            STATS['lu_order_s'] += time.perf_counter() - t0
        return a + _ORDER[key], key


_SMEM_OPTIN = dict()


def smem_optin():
    """Opt-in shared memory per block of the current device (the packer sizes the ring)."""
    d = torch.cuda.current_device()
    if d not in _SMEM_OPTIN:
        _SMEM_OPTIN[d] = int(torch.cuda.get_device_properties(d).shared_memory_per_block_optin)
    return _SMEM_OPTIN[d]


class _PinnedShmPool(object):
    """Page-locked POSIX shared-memory segments that the LU worker processes write their
    device images into: the host-to-device copy is then a DMA from pinned memory instead of a
    driver-staged pageable copy (1.9 GB/s measured under load, and it serialises other CUDA
    calls of the process).  Created lazily once the image size is known."""

    def __init__(self, nseg, seg_bytes):
        from multiprocessing import shared_memory
        import threading
        self.seg_bytes = int(seg_bytes)
        self.segs, self.free = [], []
        self.lock = threading.Lock()
        rt = torch.cuda.cudart()
        try:
            for i in range(nseg):
                shm = shared_memory.SharedMemory(create=True, size=self.seg_bytes)
                addr = C.addressof(C.c_char.from_buffer(shm.buf))
                self.segs.append((shm, addr, False))
                # page-locking touches (and thereby really allocates) every page of the segment:
                # if /dev/shm cannot back it this fails here, in the parent, not as a SIGBUS in a
                # worker that writes into it later
                if int(rt.cudaHostRegister(addr, self.seg_bytes, 0)) != 0:
                    raise MemoryError('cudaHostRegister failed')
                self.segs[-1] = (shm, addr, True)
                self.free.append(i)
        except Exception:
            self.close()
            raise

    def acquire(self):
        with self.lock:
            return self.free.pop() if self.free else None

    def release(self, i):
        with self.lock:
            self.free.append(i)

    def close(self):
        rt = torch.cuda.cudart()
        for shm, addr, pinned in self.segs:
            try:
                if pinned:
                    rt.cudaHostUnregister(addr)
            except Exception:
                pass
            try:
                shm.close()
            except Exception:
                pass            # exported buffer views may still exist at interpreter exit
            try:
                shm.unlink()
            except Exception:
                pass
        self.segs, self.free = [], []


_SHM = dict(pool=None, disabled=False)


def _shm_pool(image_bytes=None):
    """The pinned pool (created after the first image told us the size), or None."""
    if _SHM['disabled'] or os.environ.get('OCB_NO_PINNED_POOL'):
        return None
    with _LOCK:
        return _shm_pool_locked(image_bytes)


def _shm_pool_locked(image_bytes):
    if _SHM['pool'] is None and image_bytes is not None:
        try:
            world = max(int(os.environ.get('WORLD_SIZE', '1')), 1)
            nseg = int(os.environ.get('OCB_PINNED_POOL_SEGMENTS', str(max(12, 48//world))))
            seg_bytes = int(image_bytes*1.3) + (1 << 20)
            try:
                st = os.statvfs('/dev/shm')
                fit = int(0.4*st.f_bavail*st.f_frsize)//seg_bytes
                if fit < nseg:
                    # fewer segments: images that find no free segment travel through a
                    # one-off unpinned segment (slower, still correct)
                    if fit < 4:
                        raise MemoryError('/dev/shm too small for the pinned pool')
                    nseg = fit
            except OSError:
                pass
            _SHM['pool'] = _PinnedShmPool(nseg, seg_bytes)
            import atexit
            atexit.register(_SHM['pool'].close)
        except Exception:
            _SHM['disabled'] = True
            return None
    return _SHM['pool']


_UPLOADER = dict()


def _uploader_init(device_index):
    torch.cuda.set_device(device_index)
    torch.cuda.set_stream(torch.cuda.Stream())      # thread-local: uploads leave the compute stream alone


def _uploader(device_index):
    """One upload thread PER DEVICE, bound to ``device_index`` - the device of the thread that
    created the job, not whatever is current in the thread that happens to get here first (a
    look-ahead thread starts on device 0 whatever the main thread selected)."""
    with _LOCK:
        if device_index not in _UPLOADER:
            from concurrent.futures import ThreadPoolExecutor
            _UPLOADER[device_index] = ThreadPoolExecutor(max_workers=1, initializer=_uploader_init,
                                                         initargs=(device_index,))
        return _UPLOADER[device_index]


_MAIN_DEVICE = dict(index=None)


def bind_thread_to_device(index=None):
    """Make the calling thread use CUDA device ``index`` (default: the device that was current
    in the thread that last called :func:`lookahead_thread_init`).  CUDA's current device is
    per thread; helper threads call this first."""
    index = _MAIN_DEVICE['index'] if index is None else index
    if index is not None:
        torch.cuda.set_device(index)


def tail_thread_init():
    """Called on the MAIN thread: returns the initializer of a helper thread that drives device
    work of its own (``dre_stepper``: the feed-forward tail of a time step runs beside the
    Riccati part of the next one): bound to the current device, on a stream of its own."""
    require_cuda()
    index = torch.cuda.current_device()

    def init():
        bind_thread_to_device(index)
        torch.cuda.set_stream(torch.cuda.Stream())
    return init


def lookahead_thread_init():
    """Called on the MAIN thread: returns the initializer for helper threads that submit
    factorisation jobs on its behalf (``dre_stepper`` look-ahead), bound to the device that is
    current now."""
    require_cuda()
    index = torch.cuda.current_device()
    _MAIN_DEVICE['index'] = index
    return lambda: bind_thread_to_device(index)


class FactorJob(object):
    """Several host factorisations in flight (worker processes: SuperLU + analysis +
    packing); ``result()`` uploads the images and returns the ``LU`` handles."""

    def __init__(self, mats, lu_options=None, wide=False, k_hint=None):
        from . import _lu_worker
        require_cuda()
        self.device = torch.cuda.current_device()      # results are uploaded to THIS device
        opts = dict(LU_OPTIONS if lu_options is None else lu_options)
        t0 = time.perf_counter()
        so = smem_optin()
        args, self._keys = [], []
        for m in mats:
            a, key = _with_order(_csc_args(m, opts) + (so, _pack_flags(wide, k_hint)), opts)
            args.append(a)
            self._keys.append(key)
        self.n = len(mats)
        pool = _lu_pool()
        self._done = None
        self._future = None
        if pool is None:
            self._sync = [_lu_worker.factor_image(a) for a in args]
            self._async = None
        else:
            self._sync = None
            self._async, self._slots = [], []
            shp = _shm_pool()
            for a in args:
                i = shp.acquire() if shp is not None else None
                slot = None if i is None else (shp.segs[i][0].name, shp.seg_bytes)
                self._slots.append(i)
                self._async.append(pool.apply_async(_lu_worker.factor_image_to_shm, (a, slot)))
            self._counted = True
            _jobs_inflight(+1)
        STATS['lu_submit_s'] += time.perf_counter() - t0
        STATS['n_factor'] += self.n
        timeline('submit %d' % self.n, time.time() - (time.perf_counter() - t0))

    def start_upload(self):
        """Collect the worker results and upload the images from a helper thread on its own
        CUDA stream (copy engine), so that the main thread does not spend its time in
        pageable host-to-device copies; ``result()`` then only joins."""
        if self._future is None and self._done is None and not os.environ.get('OCB_NO_UPLOAD_THREAD'):
            self._future = _uploader(self.device).submit(self._collect)
        return self

    def result(self):
        if self._done is not None:
            return self._done
        if self._future is not None:
            t0 = time.perf_counter()
            out = self._future.result()
            STATS['lu_wait_s'] += time.perf_counter() - t0
            timeline('result_wait', time.time() - (time.perf_counter() - t0))
            return out
        return self._collect()

    def _collect(self):
        if self._done is not None:
            return self._done
        if torch.cuda.current_device() != self.device:
            torch.cuda.set_device(self.device)
        try:
            return self._collect_inner()
        finally:
            if getattr(self, '_counted', False):
                self._counted = False
                _jobs_inflight(-1)

    def _collect_inner(self):
        out = []
        if self._async is None:
            for (img, tf, tp, order, guard), key in zip(self._sync, self._keys):
                STATS['lu_factor_s'] += tf
                STATS['lu_worker_pack_s'] += tp
                _record_guard(guard)
                if order is not None and key is not None:
                    _ORDER.setdefault(key, (order, None))
                out.append(LU(None, image=img))
        else:
            from multiprocessing import shared_memory
            shp0 = _shm_pool()
            for idx, (ar, slot, key) in enumerate(zip(self._async, self._slots, self._keys)):
                t0 = time.perf_counter()
                try:
                    name, nbytes, tf, tp, order, guard = ar.get(
                        timeout=float(os.environ.get('OCB_LU_TIMEOUT_S', '900')))
                except Exception as exc:        # a dead worker would otherwise block forever
                    # give the pinned segments of this and all later jobs back to the pool (the
                    # pool would otherwise shrink for good)
                    if shp0 is not None:
                        for sl in self._slots[idx:]:
                            if sl is not None:
                                shp0.release(sl)
                    self._slots = [None]*len(self._slots)
                    raise RuntimeError('optconpy_b200: host LU worker failed or timed out: %r' % (exc,))
                _record_guard(guard)
                if order is not None and key is not None:
                    _ORDER.setdefault(key, (order, None))
                STATS['lu_collect_wait_s'] += time.perf_counter() - t0
                timeline('collect_wait', time.time() - (time.perf_counter() - t0))
                tw0 = time.time()
                STATS['lu_factor_s'] += tf
                STATS['lu_worker_pack_s'] += tp
                shp = _shm_pool(nbytes)
                if name is None:                  # the worker wrote into our pinned segment
                    img = np.frombuffer(shp.segs[slot][0].buf, dtype=np.uint8, count=nbytes)
                    try:
                        out.append(LU(None, image=img))       # DMA from page-locked memory
                    finally:
                        del img
                        shp.release(slot)
                    timeline('upload', tw0)
                    continue
                if slot is not None:
                    shp.release(slot)
                STATS['lu_unpinned_uploads'] += 1
                shm = shared_memory.SharedMemory(name=name)
                try:
                    img = np.frombuffer(shm.buf, dtype=np.uint8, count=nbytes)
                    out.append(LU(None, image=img))           # uploads synchronously
                    del img
                finally:
                    shm.close()
                    shm.unlink()
        self._sync = self._async = None
        self._done = out
        return out


def factorize_many(mats, lu_options=None):
    """Factorise several matrices (the shifts of one ADI) on the host, in parallel worker
    processes when available, then upload each: the separately timed setup."""
    return FactorJob(mats, lu_options).result()


_ARENA_WARM = set()
_ARENA_CLASSES = []


def _arena_class(nbytes, classes=_ARENA_CLASSES):
    """Smallest known size class that holds ``nbytes`` without wasting more than a quarter, or a
    new one: ``nbytes`` + 4 % rounded up to 1/16 .. 1/32 of its magnitude."""
    nbytes = int(nbytes)
    cls = next((c for c in classes if nbytes <= c <= 1.25*nbytes + 4096), None)
    if cls is None:
        step = 1 << max(nbytes.bit_length() - 5, 9)
        cls = -(-int(1.04*nbytes)//step)*step
        classes.append(cls)
        classes.sort()
    return cls


def _new_arena(nbytes):
    """Device buffer for one factor image.  The images of one run differ by a few KB (pivoting
    changes the fill slightly), and torch's caching allocator only reuses a freed block for a
    request that is not larger.  Requests are therefore rounded up to STICKY size classes (the
    first image of a new size defines one with 4 % headroom; later images up to that size use
    it), and a new class makes the allocator cache enough blocks of it for the look-ahead
    pipeline (per stream: the uploads run on their own).  Without this, cudaMalloc calls - each
    one synchronises the device - kept turning up inside the timed steps whenever the size
    sequence happened to miss the cache (host-API steps of 37 ms vs 41-89 ms)."""
    cls = _arena_class(nbytes)
    key = (cls, torch.cuda.current_stream().cuda_stream, torch.cuda.current_device())
    if key not in _ARENA_WARM and cls >= (4 << 20):
        _ARENA_WARM.add(key)
        tw = time.time()
        count = min(int(os.environ.get('OCB_ARENA_PREWARM', '64')), (1 << 30)//cls)
        warm = [torch.empty(cls, dtype=torch.uint8, device=cur_device()) for _ in range(count)]
        del warm
        timeline('arena_prewarm %d x %d' % (count, cls), tw)
    return torch.empty(cls, dtype=torch.uint8, device=cur_device())


class LU(object):
    """Device-resident LU factorisation ``Pr A Pc = L U`` (handle of the C ABI)."""

    def __init__(self, mat, lu_options=None, image=None, wide=False):
        """``wide=True`` also uploads the flat program, so that blocks of >= 640 right-hand
        sides go through the all-columns-at-once executor (factor read once per solve)."""
        lib = require_cuda()
        if image is None:
            from . import _lu_worker
            opts = dict(LU_OPTIONS if lu_options is None else lu_options)
            a, key = _with_order(_csc_args(mat, opts) + (smem_optin(), _pack_flags(wide)), opts)
            image, tf, tp, order, guard = _lu_worker.factor_image(a)
            _record_guard(guard)
            if order is not None and key is not None:
                _ORDER.setdefault(key, (order, None))
            STATS['lu_factor_s'] += tf
            STATS['lu_worker_pack_s'] += tp
            STATS['n_factor'] += 1
        t1 = time.perf_counter()
        h = C.c_void_p()
        # the device image lives in a torch buffer: the caching allocator makes creating and
        # dropping a factorisation free of cudaMalloc / cudaFree (both synchronise the device)
        tw = time.time()
        self.arena = _new_arena(int(image.nbytes))
        STATS['lu_arena_s'] += time.perf_counter() - t1
        timeline('arena', tw)
        tw = time.time()
        _cabi.check(lib.ocb_lu_create_from_image(C.byref(h), image.ctypes.data, image.nbytes,
                                                 ptr(self.arena), stream_ptr()),
                    'ocb_lu_create_from_image')
        timeline('create_from_image', tw)
        self.handle = h
        self._lib = lib
        info = (C.c_int64*8)()
        _cabi.check(lib.ocb_lu_info(h, info), 'ocb_lu_info')
        st8 = (C.c_int64*8)()
        _cabi.check(lib.ocb_lu_stats(h, st8), 'ocb_lu_stats')
        self.n = int(info[0])
        self.info = dict(n=info[0], nnzL=info[1], nnzU=info[2], levelsL=info[3],
                         levelsU=info[4], device_bytes=info[5], stream_kp=info[6],
                         stream_batches=info[7], n_ext=st8[0], supernodes=st8[1],
                         max_supernode=st8[2], segments=st8[3], program_rows=st8[4],
                         program_entries=st8[5], stage_bytes=st8[6], stages=st8[7])
        STATS['lu_analyse_upload_s'] += time.perf_counter() - t1
        STATS['h2d_bytes'] += image.nbytes

    def __del__(self):
        try:
            if getattr(self, 'handle', None) is not None and self.handle.value:
                self._lib.ocb_lu_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def algorithmic_bytes(self, k):
        """12*(nnzL+nnzU) + 16*(n+1) + 32*n*k per n x k solve."""
        i = self.info
        return 12*(i['nnzL'] + i['nnzU']) + 16*(i['n'] + 1) + 32*i['n']*k

    def solve(self, B, nrows_out=None, out=None):
        """``(A^-1 [B; 0])[:nrows_out]`` for a device block B (rows <= n)."""
        lib = self._lib
        assert B.dtype == torch.float64 and B.dim() == 2
        k = B.shape[1]
        nrows_out = self.n if nrows_out is None else nrows_out
        if out is None:
            out = torch.empty((nrows_out, k), dtype=torch.float64, device=B.device)
        if k == 0:
            return out
        assert B.stride(1) == 1
        self.arena.record_stream(torch.cuda.current_stream())
        wsb = lib.ocb_lu_solve_ws_bytes(self.handle, k)
        ws = workspace('lu', wsb) if wsb > 0 else None
        _cabi.check(lib.ocb_lu_solve(self.handle, ptr(B), B.stride(0), B.shape[0],
                                     ptr(out), out.stride(0), nrows_out, k,
                                     ptr(ws), wsb, stream_ptr()), 'ocb_lu_solve')
        return out

    def smw_solve(self, B, NV, Ufb=None, Vt=None, nrows_out=None):
        """``((A - Ue Ve)^-1 [B; 0])[:nrows_out]`` with ``Ue=[Ufb;0]`` (dense NV x m)
        and ``Ve=[Vt, 0]`` (DeviceCSR m x NV)."""
        lib = self._lib
        k = B.shape[1]
        nrows_out = self.n if nrows_out is None else nrows_out
        m = 0 if Ufb is None else Ufb.shape[1]
        self.arena.record_stream(torch.cuda.current_stream())
        out = torch.empty((nrows_out, k), dtype=torch.float64, device=B.device)
        wsb = lib.ocb_smw_solve_ws_bytes(self.handle, k, m)
        ws = workspace('smw', wsb)
        _cabi.check(lib.ocb_smw_solve(
            self.handle, NV, ptr(B), B.stride(0), B.shape[0], k,
            ptr(Ufb), Ufb.stride(0) if m else 0, m,
            ptr(Vt.rowptr) if m else 0, ptr(Vt.colidx) if m else 0, ptr(Vt.vals) if m else 0,
            ptr(out), out.stride(0), nrows_out, ptr(ws), wsb, stream_ptr()), 'ocb_smw_solve')
        return out


class SaddleAssembler(object):
    """``[[A, J^T], [J, 0]]`` in CSC for many A of ONE sparsity pattern (the shifted
    matrices ``At + mu Mt`` of an ADI, every time step): the block matrix is assembled once
    with ``bmat``; afterwards only the values of the (1,1) block are scattered into a copy
    of the CSC data array (``sps.bmat`` costs ~2 ms per matrix, this ~0.2 ms)."""

    def __init__(self, apattern, jmat, jmatT=None):
        P = sps.csr_matrix(apattern, dtype=np.float64, copy=True)
        P.sum_duplicates()
        P.sort_indices()
        self.indptr, self.indices = P.indptr.copy(), P.indices.copy()
        tag = sps.csr_matrix((np.arange(1, P.nnz+1, dtype=np.float64), self.indices, self.indptr),
                             shape=P.shape)
        NV = P.shape[0]
        K = sadpnt_matrix(tag, jmat, jmatT)
        K.sort_indices()
        # entries of the (1,1) block carry their CSR position + 1 as value
        cols = np.repeat(np.arange(K.shape[1]), np.diff(K.indptr))
        in_a = (K.indices < NV) & (cols < NV)
        self.pos = np.flatnonzero(in_a)
        self.src = np.rint(K.data[self.pos]).astype(np.int64) - 1
        base = sadpnt_matrix(sps.csr_matrix(P.shape), jmat, jmatT)
        self.K = K
        self.base_data = np.zeros_like(K.data)
        # J / J^T values: assemble once with a zero (1,1) block on the same pattern
        Kj = K.copy()
        Kj.data[self.pos] = 0.0
        tagged = K.data.copy()
        tagged[self.pos] = 0.0
        self.base_data = tagged
        del base, Kj

    def matches(self, amat):
        return (sps.isspmatrix_csr(amat) and amat.has_sorted_indices
                and amat.nnz == len(self.indices)
                and np.array_equal(amat.indptr, self.indptr)
                and np.array_equal(amat.indices, self.indices))

    def assemble(self, adata):
        """CSC saddle-point matrix whose (1,1) block has the CSR values ``adata``."""
        data = self.base_data.copy()
        data[self.pos] = adata[self.src]
        return sps.csc_matrix((data, self.K.indices, self.K.indptr), shape=self.K.shape)


def sadpnt_matrix(amat, jmat, jmatT=None):
    """``[[A, J^T], [J, 0]]`` (host, CSC) — the coefficient of every saddle-point solve."""
    nnpp = jmat.shape[0]
    if jmatT is None:
        jmatT = jmat.T
    return sps.bmat([[sps.csr_matrix(amat), sps.csr_matrix(jmatT)],
                     [sps.csr_matrix(jmat), sps.csr_matrix((nnpp, nnpp))]], format='csc')


def gram(Z, W, out=None):
    """``Z^T W`` on the FP64 tensor pipe (device tensors, row-major)."""
    lib = require_cuda()
    n, ka = Z.shape
    kb = W.shape[1]
    assert W.shape[0] == n and Z.stride(1) == 1 and W.stride(1) == 1
    G = torch.empty((ka, kb), dtype=torch.float64, device=Z.device) if out is None else out
    assert G.shape == (ka, kb) and G.stride(1) == 1
    if ka == 0 or kb == 0 or n == 0:
        return G.zero_()
    wsb = lib.ocb_gram_ws_bytes(n, ka, kb)
    ws = workspace('gram', wsb)
    _cabi.check(lib.ocb_gram(ptr(Z), Z.stride(0), ka, ptr(W), W.stride(0), kb, n,
                             ptr(G), G.stride(0), ptr(ws), wsb, stream_ptr()), 'ocb_gram')
    return G


def tall_gemm(Z, T, alpha=1.0):
    """``alpha * Z @ T`` for a tall Z (n x k) and a small T (k x kc)."""
    lib = require_cuda()
    n, k = Z.shape
    kc = T.shape[1]
    assert T.shape[0] == k and Z.stride(1) == 1 and T.stride(1) == 1
    Cm = torch.empty((n, kc), dtype=torch.float64, device=Z.device)
    _cabi.check(lib.ocb_tall_gemm(ptr(Z), Z.stride(0), n, k, ptr(T), T.stride(0), kc,
                                  ptr(Cm), Cm.stride(0), alpha, 0.0, stream_ptr()),
                'ocb_tall_gemm')
    return Cm


def sym_eig(G):
    """Eigen-decomposition of a small symmetric device matrix (destroys a copy):
    returns (lam descending, V with eigenvectors in columns, sweeps)."""
    lib = require_cuda()
    k = G.shape[0]
    Gw = G.clone().contiguous()
    lam = torch.empty((k,), dtype=torch.float64, device=G.device)
    V = torch.empty((k, k), dtype=torch.float64, device=G.device)
    sw = C.c_int32(0)
    _cabi.check(lib.ocb_sym_eig(ptr(Gw), Gw.stride(0), k, ptr(lam), ptr(V), V.stride(0),
                                C.byref(sw), stream_ptr()), 'ocb_sym_eig')
    return lam, V, sw.value


# Singular values below COMPRESS_DELTA * sigma_max are not resolved reliably by a Gram matrix in
# FP64 (lambda = sigma^2 carries an absolute error of ~eps * sigma_max^2); a threshold below
# that is honoured by DEFLATION: the resolved part is projected out of Z and the remainder -
# whose own sigma_max is now small - goes through the same kernels again.
COMPRESS_DELTA = 1e-5
COMPRESS_NOISE = 1e-13


def _compress_once(Z, thresh, k, eta, rmax):
    lib = require_cuda()
    n, K = Z.shape
    assert Z.stride(1) == 1
    rmax = int(min(K, n, 1024)) if rmax is None else int(min(rmax, K, n, 1024))
    rmax = max(rmax, 1)
    cap = rmax if k is None else int(min(rmax, k))
    Zc = torch.empty((n, max(cap, 1)), dtype=torch.float64, device=Z.device)
    sig = torch.zeros((rmax,), dtype=torch.float64, device=Z.device)
    info = (C.c_int64*3)()
    wsb = lib.ocb_compress_ws_bytes(n, K, rmax)
    # the K x K Gram matrix grows with every DRE step (K = block width x ADI steps): ask for
    # headroom, so that the buffer is not re-allocated (cudaFree + cudaMalloc: both synchronise)
    # every few time steps
    have = _WS.get(('compress', threading.get_ident()))
    if have is None or have.buf is None or have.buf.numel() < wsb:
        workspace('compress', int(wsb*1.6))
    ws = workspace('compress', wsb)
    _cabi.check(lib.ocb_compress(ptr(Z), Z.stride(0), n, K,
                                 -1.0 if thresh is None else float(thresh),
                                 0 if k is None else int(k), float(eta), rmax,
                                 ptr(Zc), Zc.stride(0), cap, ptr(sig), info,
                                 ptr(ws), wsb, stream_ptr()), 'ocb_compress')
    keep = int(info[0])
    if int(info[1]) >= rmax and rmax < min(K, n) and (k is None or keep < int(k)):
        import warnings
        warnings.warn('optconpy_b200: compress_Zsvd reached its rank limit of %d before the stopping '
                      'rule fired; directions beyond it were dropped' % rmax, RuntimeWarning)
    return Zc[:, :keep], dict(kept=keep, chol_rank=int(info[1]), sweeps=int(info[2]),
                              sigma=sig[:int(info[1])])


_DEFLATION_WARM = set()


def _warm_deflation_ops(Z):
    """The deflation levels below run kernels, and need device blocks and workspaces, that the
    first time steps of a run do not: they start when sigma_max has grown past thresh / DELTA.
    Whichever time step needed them first paid for it - CUDA loads a kernel on its first launch
    (torch's element-wise kernels out of a library of more than a gigabyte: 50 ms on a cold page
    cache), and the factor-sized temporaries and the larger Gram workspace were fresh cudaMallocs
    (10-13 ms, synchronising) - i.e. one 40-80 ms step in every bench.py e2e run.  So the very
    first compression of a process runs one deflation step on its own input (a 32-column
    subspace; results discarded, ~1 ms) and leaves three factor-sized blocks with 50 % headroom
    in the allocator's cache."""
    _DEFLATION_WARM.add(Z.device.index)
    k1 = max(1, min(32, Z.shape[1]//2))
    z1 = Z[:, :k1].contiguous()
    sg = torch.from_numpy(np.ones(k1)).to(Z.device)
    q = (z1/sg).contiguous()
    lam, wq, _ = sym_eig(gram(q, q))
    q = tall_gemm(q, (wq/torch.sqrt(lam.abs() + 1.0)).contiguous())
    z2 = (Z - tall_gemm(q, gram(q, Z))).contiguous()
    torch.cat([z1, z2[:, :1]], dim=1).contiguous()
    torch.cat([lam[:1], sg[:1]])
    lam.cpu()
    del z2
    spare = [torch.empty(int(1.5*Z.numel()), dtype=torch.float64, device=Z.device) for _ in range(3)]
    del spare


def compress(Z, thresh=None, k=None, eta=1e-14, rmax=None, _smax0=None, _level=0):
    """Device version of ``compress_Zsvd``: returns (Zc device tensor, info dict).
    ``Zc = Z V`` with V the right singular vectors of the singular values ``> thresh`` (at most
    ``k``), computed from Gram matrices on the FP64 tensor pipe; thresholds below the resolution
    of one Gram matrix (``COMPRESS_DELTA * sigma_max``) are reached by deflation levels."""
    if _level == 0 and Z.device.index not in _DEFLATION_WARM and Z.shape[1] >= 2:
        _warm_deflation_ops(Z)
    tw = time.time()
    Zc, info = _compress_once(Z, thresh, k, eta, rmax)
    info['levels'] = _level + 1
    if thresh is None or info['chol_rank'] == 0:
        return Zc, info
    sig = info['sigma'].cpu().numpy()
    timeline('compress:level %d' % _level, tw)
    smax = float(sig[0])
    smax0 = smax if _smax0 is None else _smax0
    if (thresh >= COMPRESS_DELTA*smax or (k is not None and info['kept'] >= int(k))
            or smax <= COMPRESS_NOISE*smax0 or _level >= 3):
        return Zc, info
    # the part this level resolves: sigma > DELTA * sigma_max (a prefix: sigma is sorted)
    k1 = int((sig > COMPRESS_DELTA*smax).sum())
    if k1 == 0 or k1 > info['kept']:
        return Zc, info
    tw = time.time()
    Zc1 = Zc[:, :k1].contiguous()
    sg1 = torch.from_numpy(sig[:k1].copy()).to(Z.device)
    Q1 = (Zc1/sg1).contiguous()                       # ~ left singular vectors U_1
    timeline('deflate:scale', tw)
    tw = time.time()
    lam, Wq, _ = sym_eig(gram(Q1, Q1))                # re-orthonormalise: Q1 <- Q1 W lam^-1/2
    timeline('deflate:gram+eig', tw)
    tw = time.time()
    Q1 = tall_gemm(Q1, (Wq/torch.sqrt(lam)).contiguous())
    timeline('deflate:orth', tw)
    tw = time.time()
    Z2 = Z - tall_gemm(Q1, gram(Q1, Z))               # (I - Q1 Q1^T) Z
    timeline('deflate:project', tw)
    t2 = max(float(thresh), COMPRESS_NOISE*smax0)
    Zc2, info2 = compress(Z2.contiguous(), thresh=t2, k=None if k is None else int(k) - k1, eta=eta,
                          rmax=rmax, _smax0=smax0, _level=_level + 1)
    out = torch.cat([Zc1, Zc2], dim=1).contiguous()
    return out, dict(kept=out.shape[1], chol_rank=info['chol_rank'], sweeps=info['sweeps'],
                     sigma=torch.cat([info['sigma'][:k1], info2['sigma']]), levels=info2['levels'])


def compress_from_gram(G, thresh=None, k=None, eta=1e-14, rmax=None):
    """Replicated half of the compression from a K x K Gram matrix (column-sharded runs):
    returns (T device tensor K x keep with ``Zc = Z T``, info)."""
    lib = require_cuda()
    K = G.shape[0]
    rmax = int(min(K, 1024)) if rmax is None else int(min(rmax, K, 1024))
    cap = rmax if k is None else int(min(rmax, k))
    T = torch.empty((K, max(cap, 1)), dtype=torch.float64, device=G.device)
    sig = torch.zeros((rmax,), dtype=torch.float64, device=G.device)
    info = (C.c_int64*3)()
    wsb = lib.ocb_compress_gram_ws_bytes(K, rmax)
    ws = workspace('compress_gram', wsb)
    _cabi.check(lib.ocb_compress_from_gram(ptr(G), G.stride(0), K, -1.0 if thresh is None else float(thresh),
                                           0 if k is None else int(k), float(eta), rmax, ptr(T), T.stride(0),
                                           cap, ptr(sig), info, ptr(ws), wsb, stream_ptr()),
                'ocb_compress_from_gram')
    keep = int(info[0])
    return T[:, :keep], dict(kept=keep, chol_rank=int(info[1]), sweeps=int(info[2]), sigma=sig[:int(info[1])])


def p2p_put2d(src, dst_ptr, ldd):
    """2-D block copy of a device block into a (peer-mapped) buffer at ``dst_ptr``."""
    lib = require_cuda()
    assert src.dtype == torch.float64 and src.dim() == 2 and src.stride(1) == 1
    _cabi.check(lib.ocb_p2p_put2d(ptr(src), src.stride(0), src.shape[0], src.shape[1], int(dst_ptr), int(ldd),
                                  stream_ptr()), 'ocb_p2p_put2d')


def p2p_sum_peers(ptrs, count, out):
    lib = require_cuda()
    arr = (C.c_void_p*len(ptrs))(*[int(p) for p in ptrs])
    _cabi.check(lib.ocb_p2p_sum_peers(arr, len(ptrs), int(count), ptr(out), stream_ptr()), 'ocb_p2p_sum_peers')
    return out


def feedback(Mt, Z, tB, alpha=1.0):
    """``alpha * Mt (Z (Z^T tB))`` with Mt a DeviceCSR, Z and tB device blocks."""
    lib = require_cuda()
    NV, kz = Z.shape
    m = tB.shape[1]
    out = torch.empty((NV, m), dtype=torch.float64, device=Z.device)
    if kz == 0:
        return out.zero_()
    wsb = lib.ocb_feedback_ws_bytes(NV, kz, m)
    ws = workspace('feedback', wsb)
    _cabi.check(lib.ocb_feedback(ptr(Mt.rowptr), ptr(Mt.colidx), ptr(Mt.vals), NV,
                                 ptr(Z), Z.stride(0), kz, ptr(tB), tB.stride(0), m,
                                 ptr(out), out.stride(0), alpha, ptr(ws), wsb,
                                 stream_ptr()), 'ocb_feedback')
    return out


NORM_HOOK_T = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_void_p)


def adi_run(lus, shifts, NV, NP, Mt, W, maxsteps, reltol, Ufb=None, Vt=None, norm_reduce=None):
    """The LR-ADI loop on the device.  Returns (Z device tensor NV x (steps*k),
    list of relative norms)."""
    lib = require_cuda()
    k = W.shape[1]
    m = 0 if Ufb is None else Ufb.shape[1]
    nsh = len(shifts)
    steps_cap = int(maxsteps)
    need = NV*k*steps_cap*8
    if need > (4 << 30):            # only then is the driver asked (the query is not free)
        free, _ = torch.cuda.mem_get_info()
        if need > 0.6*free:
            steps_cap = max(1, int(0.6*free // (NV*k*8)))
    # the iteration writes into a persistent, monotonically growing buffer (a fresh
    # 0.4 GB torch.empty per call made the caching allocator cudaMalloc/cudaFree, i.e.
    # synchronise, whenever the block width changed); the used part is copied out compactly
    reserve_pool(min(12*NV*max(k, 64)*64*8, 4 << 30))
    zbuf = workspace('adi_Z', NV*k*steps_cap*8)
    Z = zbuf[:NV*k*steps_cap*8].view(torch.float64).view(NV, k*steps_cap)
    for lu in lus:
        lu.arena.record_stream(torch.cuda.current_stream())
    harr = (C.c_void_p*nsh)(*[lu.handle.value for lu in lus])
    sarr = (C.c_double*nsh)(*[float(s) for s in shifts])
    rel = (C.c_double*int(maxsteps))()
    nst = C.c_int64(0)
    wsb = lib.ocb_adi_ws_bytes(NV+NP, k, m, nsh, harr)
    ws = workspace('adi', wsb)
    hook = None
    if norm_reduce is not None:
        # column-sharded run: ``norm_reduce(local ||V_i||^2) -> global`` (an all-reduce)
        def _hook(pv, _ctx):
            pv[0] = float(norm_reduce(pv[0]))
        hook = NORM_HOOK_T(_hook)
        _cabi.check(lib.ocb_adi_set_norm_hook(C.cast(hook, C.c_void_p), None), 'set_norm_hook')
    try:
        rc = _adi_call(lib, harr, sarr, nsh, NV, NP, Mt, W, k, Ufb, m, Vt, maxsteps, steps_cap, reltol,
                       Z, rel, nst, ws, wsb)
    finally:
        if hook is not None:
            lib.ocb_adi_set_norm_hook(None, None)
    _cabi.check(rc, 'ocb_adi_run')
    steps = int(nst.value)
    if steps_cap < int(maxsteps) and steps == steps_cap and rel[steps-1] > reltol:
        import warnings
        warnings.warn('optconpy_b200: LR-ADI stopped after %d of %d allowed steps because the factor '
                      'buffer (%d columns per step) would not fit the free device memory; the factor '
                      'is not the one adi_max_steps=%d would give' % (steps, maxsteps, k, maxsteps),
                      RuntimeWarning)
    # clone, not contiguous(): when the iteration used the whole buffer the slice IS contiguous
    # and contiguous() would hand out a view of the workspace that the next call overwrites
    return Z[:, :steps*k].clone(memory_format=torch.contiguous_format), [rel[i] for i in range(steps)]


def _adi_call(lib, harr, sarr, nsh, NV, NP, Mt, W, k, Ufb, m, Vt, maxsteps, steps_cap, reltol, Z, rel,
              nst, ws, wsb):
    return (lib.ocb_adi_run(
        harr, sarr, nsh, NV, NP, ptr(Mt.rowptr), ptr(Mt.colidx), ptr(Mt.vals),
        ptr(W), W.stride(0), k, ptr(Ufb), Ufb.stride(0) if m else 0, m,
        ptr(Vt.rowptr) if m else 0, ptr(Vt.colidx) if m else 0, ptr(Vt.vals) if m else 0,
        int(min(maxsteps, steps_cap)), float(reltol), ptr(Z), Z.stride(0), Z.shape[1],
        rel, C.byref(nst), ptr(ws), wsb, stream_ptr()))


_FP64_PEAK = dict()


def fp64_peak(kind='dmma'):
    """Measured FP64 peak (TFLOP/s) of the current device: ``'dmma'`` = the FP64 tensor pipe
    (DMMA.8x8x4 chains), ``'dfma'`` = the FP64 FMA pipe.  Best of 1/2/4 CTAs per SM; cached."""
    lib = require_cuda()
    key = (torch.cuda.current_device(), kind)
    if key not in _FP64_PEAK:
        sink = torch.zeros(1, dtype=torch.float64, device=cur_device())
        best = 0.0
        for cps in (1, 2, 4):
            tf = C.c_double(0.0)
            _cabi.check(lib.ocb_fp64_peak(0 if kind == 'dmma' else 1, 20000, cps, C.byref(tf),
                                          ptr(sink), stream_ptr()), 'ocb_fp64_peak')
            best = max(best, tf.value)
        _FP64_PEAK[key] = best
    return _FP64_PEAK[key]


def launch_count():
    return int(_cabi.load().ocb_launch_count())
