"""Drop-in alias: makes the CUDA-backed modules importable under the name the reference
imports (``optcont_main.py:13-14``, ``solve_dae_ric.py:3-4``).  Put ``shim/`` and the repo
root on ``PYTHONPATH``:

    import sadptprj_riclyap_adi.lin_alg_utils as lau
    import sadptprj_riclyap_adi.proj_ric_utils as pru
"""
