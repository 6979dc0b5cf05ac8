"""optconpy_b200 — B200-native (sm_100a) implementation of the LR-ADI / projected
Riccati hot path of highlando/optconpy, behind the ``sadptprj_riclyap_adi``
module interface the reference driver imports (``optcont_main.py:13-14``).

    import optconpy_b200.lin_alg_utils as lau
    import optconpy_b200.proj_ric_utils as pru

The compute modules need a CUDA device and the in-tree ``liboptconpy_b200.so``;
there is no CPU fallback (importing this top-level package alone is harmless so
that problem generation works on a CPU box).
"""
__version__ = '0.1.0'

import os as _os

# Load every kernel of the library when it is opened instead of at its first launch: with
# CUDA's default lazy loading the first launch of each template instance (e.g. the two-column
# panel variant of the solve kernel, first needed when a block gets wider than one wave of
# clusters) stalled a time step by 100-700 ms in the middle of a run.
_os.environ.setdefault('CUDA_MODULE_LOADING', 'EAGER')
