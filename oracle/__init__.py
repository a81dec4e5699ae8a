"""CPU oracle for the BiFold policy hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and there only as the checker or as the timed CPU baseline.  The product path
(``bifold_b200``) never imports this package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
restatement in ``policy_oracle.py`` is pinned against outputs of the reference itself, imported in
the build container from ``/root/reference`` by ``make_golden.py`` (shims in ``ref_shims.py``) and
committed as fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the
restatement against those fixtures on every run (CPU, no reference needed).
"""
