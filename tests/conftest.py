import os
import sys

import pytest

# small dense problems: a full-width BLAS thread pool only thrashes (the CPU suite ran 3x
# slower with 8-16 threads than with 2-4)
os.environ.setdefault('OMP_NUM_THREADS', '4')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '4')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device; on a GPU box
    nothing is skipped, and the package itself has no CPU path to fall back to."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def cav10():
    from optconpy_b200 import problems as pb
    return pb.drivcav_problem(10, 1e-2)


@pytest.fixture(scope='session')
def cav6():
    from optconpy_b200 import problems as pb
    return pb.drivcav_problem(6, 1e-2)


@pytest.fixture(scope='session', autouse=True)
def _limit_blas_threads():
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        yield
        return
    with threadpool_limits(limits=4):
        yield
