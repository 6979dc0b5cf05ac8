"""CPU suite, part 3: host-side logic that needs no GPU — the synthetic problem generator
(sizes of the reference), the squeezed time mesh (``optcont_main.py:141-150``), the restated
driver's storage/memoisation semantics (``solve_dae_ric.py:143-145``), the ``.npy`` shim and
the data-string naming (``optcont_main.py:153-157``), and the column partition."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import lin_alg_utils as olau, proj_ric_utils as opru
from optconpy_b200 import problems as pb, scenarios as sc, dre_stepper as ds, parallel as par


@pytest.mark.parametrize('N,NV,NP', [(10, 722, 120), (20, 3042, 440), (25, 4802, 675)])
def test_problem_sizes_match_reference(N, NV, NP):
    # NV = 2(2N-1)^2, NP = (N+1)^2 - 1; N=20 -> NV 3042 is recorded in debugstuff.py:30
    assert 2*(2*N-1)**2 == NV and (N+1)**2 - 1 == NP
    if N > 10:
        return
    prob = pb.drivcav_problem(N, 1e-2)
    assert prob['NV'] == NV and prob['NP'] == NP
    assert prob['M'].shape == (NV, NV) and prob['J'].shape == (NP, NV)


def test_matrices_are_a_stokes_system(cav6):
    M, A, J = cav6['M'], cav6['A'], cav6['J']
    assert abs(M - M.T).max() < 1e-14 and abs(A - A.T).max() < 1e-12
    assert np.linalg.eigvalsh(M.toarray()).min() > 0
    assert np.linalg.eigvalsh(A.toarray()).min() > 0
    assert np.linalg.matrix_rank(J.toarray()) == cav6['NP']      # last pressure row dropped
    # convection matrix of a divergence-free field is (nearly) skew in the interior
    Nc = pb.convection_matrix(cav6, pb.analytic_vortex, newton_term=False)
    assert Nc.shape == M.shape and abs(Nc).max() > 0


def test_squeezed_time_mesh():
    t = pb.get_tint(0.0, 1.0, 6)
    ref = (np.sin(np.linspace(-0.5*np.pi, 0.5*np.pi, 7)) + 1)*0.5
    assert np.allclose(t, ref) and t[0] == 0.0 and np.isclose(t[-1], 1.0)
    t2 = pb.get_tint(0.0, 0.2, 128)
    d = np.diff(t2)
    assert len(t2) == 129 and d.min() > 0 and d[0] < d[64] and d[-1] < d[64]
    assert np.allclose(pb.get_tint(1.0, 3.0, 4, sqzmesh=False), np.linspace(1, 3, 5))


def test_datastr_and_npy_store(tmp_path):
    s = ds.default_datastr(time=0.5, meshp=10, nu=0.01, Nts=10, data_prfx='data/tdst_')
    assert s == 'data/tdst_time0.5_nu0.01_mesh10_Nts10'
    st = ds.NpyStore()
    a = np.arange(6.0).reshape(3, 2)
    key = str(tmp_path / 'sub' / 'x__Z')
    st.save(a, key)
    assert os.path.exists(key + '.npy')            # np.save appends .npy (optcont_main.py:231)
    assert np.array_equal(st.load(key), a)
    with pytest.raises(IOError):
        st.load(str(tmp_path / 'missing'))


def test_stepper_memoises_and_keeps_reference_order(cav6):
    cs = pb.control_setup(cav6, olau, alphau=1e-4)
    tmesh = pb.get_tint(0.0, 0.1, 2)
    nd = dict(sc.DEFAULT_NWTN_ADI, adi_max_steps=60, nwtn_max_steps=4)
    kw = sc.dre_kwargs(cav6, cs, tmesh, nd, 1e-3, sc._ystar_sin(cs['NY']))
    store, info = ds.MemStore(), []
    fb = ds.solve_flow_daeric(lau=olau, pru=opru, store=store, stepinfo=info, **kw)
    assert sorted(fb) == sorted(tmesh)
    for t in fb:
        assert store[fb[t]['mtxtb']].shape == (cav6['NV'], 2*cs['NU'])
        assert store[fb[t]['w']].shape == (cav6['NV'], 1)
    # gain definition at the terminal time: mtxtb = -M^T Zc Zc^T tB (solve_dae_ric.py:101)
    tE = tmesh[-1]
    kE = fb[tE]['mtxtb'].replace('__mtxtb', '__Z')
    tb = olau.apply_invsqrt_fromright(cs['R'], cs['b_mat'], output='sparse')
    assert np.allclose(store[fb[tE]['mtxtb']],
                       -opru.get_mTzzTtb(cav6['M'].T, store[kE], tb), rtol=1e-12, atol=1e-14)
    # second run over the same store: every __Z entry loads, no Riccati solve is repeated
    calls = []

    class CountingPru(object):
        def __getattr__(self, name):
            if name == 'proj_alg_ric_newtonadi':
                calls.append(name)
            return getattr(opru, name)
    fb2 = ds.solve_flow_daeric(lau=olau, pru=CountingPru(), store=store,
                               **dict(kw, gtdtstrargs=dict(kw['gtdtstrargs'])))
    assert not calls and fb2 == fb


def test_inconsistent_output_operator_raises(cav6):
    cs = pb.control_setup(cav6, olau, alphau=1e-4)
    tmesh = pb.get_tint(0.0, 0.1, 1)
    kw = sc.dre_kwargs(cav6, cs, tmesh, dict(sc.DEFAULT_NWTN_ADI), 1e-3, sc._ystar_sin(cs['NY']))
    bad = sps.csr_matrix(np.random.default_rng(0).standard_normal(kw['mcmat'].shape))
    with pytest.raises(Warning):                     # solve_dae_ric.py:79,83
        ds.solve_flow_daeric(lau=olau, pru=opru, store=ds.MemStore(), **dict(kw, mcmat=bad))


def test_column_slices_partition():
    for k in (1, 7, 8, 66, 1024):
        for world in (1, 2, 3, 8):
            sl = [par.column_slice(k, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == k
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            w = [b - a for a, b in sl]
            assert max(w) - min(w) <= 1


def test_outer_newton_carry(cav6):
    """Outer Newton passes (optcont_main.py:577-600): ``curnwtnsdict`` names the entries that
    accumulate w and the gain across passes (init_nwtnstps_value_dict, :200-210).  Pass 1 fills
    them; pass 2 finds them, hands ``sqrt(tau) * cnsmtxtb`` to the Riccati solver as ``mtxoldb``
    and adds the carried w to the right-hand side (solve_dae_ric.py:134-141,150,176-178)."""
    cs = pb.control_setup(cav6, olau, alphau=1e-4)
    tmesh = pb.get_tint(0.0, 0.1, 2)
    nd = dict(sc.DEFAULT_NWTN_ADI, adi_max_steps=60, nwtn_max_steps=3)
    kw = sc.dre_kwargs(cav6, cs, tmesh, nd, 1e-3, sc._ystar_sin(cs['NY']))
    cnd = {t: dict(v='cns_v_t%r' % t, mtxtb='cns_mtxtb_t%r' % t, w='cns_w_t%r' % t) for t in tmesh}
    store = ds.MemStore()
    seen = []

    class SpyPru(object):
        def __getattr__(self, name):
            f = getattr(opru, name)
            if name != 'proj_alg_ric_newtonadi':
                return f

            def wrapped(**k):
                seen.append(None if k.get('mtxoldb') is None else np.array(k['mtxoldb']))
                return f(**k)
            return wrapped

    def one_pass(cns):
        g = dict(kw['gtdtstrargs'], data_prfx='cns%d_' % cns)
        return ds.solve_flow_daeric(lau=olau, pru=SpyPru(), store=store, curnwtnsdict=cnd,
                                    **dict(kw, gtdtstrargs=g))
    fb1 = one_pass(0)
    assert all(s is None for s in seen)                       # nothing carried in the first pass
    for t in tmesh:
        assert cnd[t]['w'] in store and cnd[t]['mtxtb'] in store
    carried = {t: store[cnd[t]['mtxtb']].copy() for t in tmesh}
    n_first = len(seen)
    fb2 = one_pass(1)
    second = seen[n_first:]
    assert len(second) == len(tmesh) - 1 and all(s is not None for s in second)
    # the drift of the first backward step of pass 2 carries sqrt(tau) * (gain stored by pass 1)
    t_last = tmesh[-2]
    assert np.allclose(second[0], np.sqrt(tmesh[-1] - t_last)*carried[t_last], rtol=1e-12, atol=0)
    # the carried entries were accumulated, and both passes produced complete feedback dicts
    assert not np.allclose(store[cnd[t_last]['mtxtb']], carried[t_last])
    assert sorted(fb1) == sorted(fb2) == sorted(tmesh)


def test_overlapped_tail_gives_the_same_results(cav6):
    """dre_stepper: with a backend that offers ``tail_thread_init`` the feed-forward half of step k
    runs on a helper thread beside the Riccati half of step k-1.  Same stores, bit for bit, as the
    sequential loop - also across two outer Newton passes (``curnwtnsdict``: the tail of a step
    writes the carried entries the next pass reads) - and the callback sees every step in order."""
    import threading
    cs = pb.control_setup(cav6, olau, alphau=1e-4)
    tmesh = pb.get_tint(0.0, 0.1, 4)
    nd = dict(sc.DEFAULT_NWTN_ADI, adi_max_steps=60, nwtn_max_steps=3)
    kw = sc.dre_kwargs(cav6, cs, tmesh, nd, 1e-3, sc._ystar_sin(cs['NY']))
    tail_threads = set()

    class AsyncPru(object):               # the oracle plus the three optional entry points
        def __getattr__(self, name):
            f = getattr(opru, name)
            if name != 'get_mTzzTtb':
                return f

            def spy(*a, **k):
                tail_threads.add(threading.current_thread().name)
                return f(*a, **k)
            return spy

        def factors_async(self, **k):
            return None

        def tail_thread_init(self):
            return lambda: None

    class AsyncLau(object):
        def __getattr__(self, name):
            return getattr(olau, name)

        def sadlu_async(self, **k):
            return None

    def run(overlap):
        store, order = ds.MemStore(), []
        cnd = {t: dict(v='cns_v_t%r' % t, mtxtb='cns_mtxtb_t%r' % t, w='cns_w_t%r' % t) for t in tmesh}
        fbs = []
        for cns in (0, 1):
            g = dict(kw['gtdtstrargs'], data_prfx='cns%d_' % cns)
            fbs.append(ds.solve_flow_daeric(lau=AsyncLau(), pru=AsyncPru(), store=store, curnwtnsdict=cnd,
                                            lookahead=2, overlap_tail=overlap, step_callback=order.append,
                                            **dict(kw, gtdtstrargs=g)))
        return store, fbs, order
    tail_threads.clear()
    s_seq, fb_seq, o_seq = run(False)
    assert tail_threads == {threading.current_thread().name}
    tail_threads.clear()
    s_ovl, fb_ovl, o_ovl = run(True)
    assert any(n != threading.current_thread().name for n in tail_threads)      # really on a helper thread
    assert o_seq == o_ovl == 2*list(range(len(tmesh)-2, -1, -1)) and fb_seq == fb_ovl
    assert sorted(s_seq) == sorted(s_ovl)
    for k in s_seq:
        assert np.array_equal(s_seq[k], s_ovl[k]), k

    # an exception in the tail surfaces in the caller
    class Boom(AsyncLau):
        def __getattr__(self, name):
            if name == 'solve_sadpnt_smw':
                def boom(**k):
                    raise RuntimeError('tail failed')
                return boom
            return getattr(olau, name)
    with pytest.raises(RuntimeError, match='tail failed'):
        ds.solve_flow_daeric(lau=Boom(), pru=AsyncPru(), store=ds.MemStore(), lookahead=2,
                             **dict(kw, gtdtstrargs=dict(kw['gtdtstrargs'])))


def test_device_factor_handle_behaves_like_the_array():
    """proj_ric_utils.DeviceFactor (the lazily downloaded ADI factor): shape/dtype/len without a
    copy, ndarray on demand through the numpy protocol (np.asarray, np.save, slicing, hstack)."""
    import io
    import torch
    from optconpy_b200 import proj_ric_utils as gpru
    a = np.arange(12.0).reshape(4, 3)
    f = gpru.DeviceFactor(torch.from_numpy(a.copy()))
    assert f.shape == (4, 3) and f.ndim == 2 and len(f) == 4 and f.dtype == np.float64
    assert f._host is None                                    # nothing fetched yet
    assert np.array_equal(np.asarray(f), a) and f._host is not None
    assert np.array_equal(f[:, 1:], a[:, 1:])
    assert np.array_equal(np.hstack([f, f]), np.hstack([a, a]))
    assert np.asarray(f, dtype=np.float32).dtype == np.float32
    buf = io.BytesIO()
    np.save(buf, f)
    buf.seek(0)
    assert np.array_equal(np.load(buf), a)
    st = ds.MemStore()
    st.save(f, 'z')
    assert isinstance(st['z'], np.ndarray) and np.array_equal(st['z'], a)


def test_arena_size_classes_are_sticky():
    """device._arena_class: images of one run (sizes within a few per cent) share one class, so
    the caching allocator always finds a freed block; different magnitudes get their own."""
    from optconpy_b200 import device as dv
    classes = []
    first = dv._arena_class(9385216, classes)
    assert first >= 9385216*1.04 - 1 and first <= 9385216*1.10
    for nb in (9385216, 9300000, 9577472, 9700000):
        assert dv._arena_class(nb, classes) == first
    assert classes == [first]
    small = dv._arena_class(50000, classes)
    big = dv._arena_class(120000000, classes)
    assert small < first < big and classes == [small, first, big]
    assert dv._arena_class(51000, classes) == small and dv._arena_class(100, classes) >= 100
    # a request just above a class opens the next one instead of overflowing the buffer
    nxt = dv._arena_class(first + 1, classes)
    assert nxt > first and dv._arena_class(first, classes) == first


def test_worker_pattern_cache_matches_the_plain_arrangement(cav6):
    """_lu_worker._arranged: transposition + symmetric permutation through the cached entry map
    equals the plain scipy operations; a pattern with duplicate entries takes the plain path."""
    from optconpy_b200 import _lu_worker as w, device as dv
    K = dv.sadpnt_matrix(cav6['M'] + 0.1*cav6['A'], cav6['J']).tocsc()
    n = K.shape[0]
    rng = np.random.default_rng(4)
    q = rng.permutation(n).astype(np.int32)
    for transposed in (False, True):
        for qq in (None, q):
            w._ARRANGE.clear()
            for rep in range(2):          # second pass: served from the cache
                data = K.data*(1.0 + rep)
                got = w._arranged(data, K.indices, K.indptr, K.shape, qq, transposed)
                ref = sps.csc_matrix((data, K.indices, K.indptr), shape=K.shape)
                ref = ref.T.tocsc() if transposed else ref
                ref = ref[qq][:, qq].tocsc() if qq is not None else ref
                assert got.has_canonical_format and abs(got - ref).max() == 0.0
            assert len(w._ARRANGE) == 1 and next(iter(w._ARRANGE.values())) is not None
    # duplicates: (0,0) stored twice
    ind = np.array([0, 0, 1, 1], dtype=np.int32)
    ptr = np.array([0, 3, 4], dtype=np.int32)
    dat = np.array([1.0, 2.0, 3.0, 4.0])
    w._ARRANGE.clear()
    got = w._arranged(dat, ind, ptr, (2, 2), None, True)
    assert next(iter(w._ARRANGE.values())) is None
    assert np.array_equal(got.toarray(), np.array([[3.0, 3.0], [0.0, 4.0]]))
    w.tune_malloc()                       # must not raise anywhere
