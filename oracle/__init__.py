"""CPU oracle for the LR-ADI / projected-Riccati hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``optconpy_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and only
as the checker / timed CPU baseline.

PARITY UNPINNED: the arithmetic of this path lives in the
third-party module ``sadptprj_riclyap_adi`` (github.com/highlando/
sadptprj_riclyap_adi, no version pinned by the reference, ``README.md:26-33``),
which is absent from ``/root/reference``; the reference's two callers are
Python 2 + FEniCS and cannot be imported here; and the reference's only test
(``tests/test_units_compfacres_compress.py``) asserts identities on unseeded
random data, so there are no golden vectors.  This oracle therefore restates
the published algorithm (Newton-Kleinman + low-rank ADI on the projected
Lyapunov equation, saddle-point solves through scipy/SuperLU with
Sherman-Morrison-Woodbury low-rank updates) behind the exact call signatures
the reference uses, and is validated by (i) the five identities of the
reference's test, (ii) the equations the reference driver encodes
(``solve_dae_ric.py:147-194``), (iii) residuals of the Lyapunov / Riccati
equations going to zero.
"""
