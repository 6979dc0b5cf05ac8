"""CPU suite, part 2: the C-ABI library loads here (no GPU) and exports every symbol that
``include/optconpy_b200.h`` declares; the ctypes prototypes cover the same set; and the
product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, 'include')
    for fn in os.listdir(inc):
        if fn.endswith('.h'):
            src = open(os.path.join(inc, fn)).read()
            src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
            names |= set(re.findall(r'\b(ocb_[a-z0-9_]+)\s*\(', src))
    return names


@pytest.fixture(scope='module')
def lib():
    from optconpy_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _cabi.load()


def test_library_exports_every_declared_symbol(lib):
    from optconpy_b200 import _cabi
    declared = _declared_symbols()
    assert len(declared) >= 20
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, missing
    # the ctypes binding covers exactly the declared entry points
    assert set(_cabi.PROTOTYPES) == declared, set(_cabi.PROTOTYPES) ^ declared


def test_version_and_error_channel(lib):
    assert lib.ocb_version() >= 100
    assert lib.ocb_launch_count() >= 0
    # argument validation happens before any CUDA call: bad sizes -> OCB_ERR_ARG + message
    rc = lib.ocb_spmm(-1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0.0, 0)
    assert rc == -1
    assert b'argument' in lib.ocb_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is visible')
    from optconpy_b200 import device as dv
    import optconpy_b200.proj_ric_utils as gpru
    import optconpy_b200.lin_alg_utils as glau
    with pytest.raises(RuntimeError):
        dv.require_cuda()
    Z = np.ones((4, 2))
    with pytest.raises(RuntimeError):
        gpru.compress_Zsvd(Z, thresh=1e-3)
    with pytest.raises(RuntimeError):
        glau.apply_massinv(np.eye(4), Z)


def test_product_does_not_import_oracle():
    """Nothing under optconpy_b200/ or shim/ may reference the oracle package."""
    bad = []
    for base in ('optconpy_b200', 'shim'):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                    src = open(os.path.join(dp, fn)).read()
                    if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
