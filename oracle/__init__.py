"""TEST INFRASTRUCTURE ONLY — CPU restatement of the PLeaS-Merging merge hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU arm.
The product package (``pleas_merging_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md §4), so the pins are outputs of the unmodified reference itself,
imported from ``/root/reference`` in the build container by
``oracle/make_golden.py`` (committed) and stored under ``tests/golden/``; the
LAP restatement is additionally pinned against SciPy's
``linear_sum_assignment`` (the third-party routine the reference calls at
``pleas/core/solvers.py:29-31``) on the instances in
``tests/golden/lap_golden.npz``.
"""
