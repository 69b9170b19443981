"""CPU oracle for the DA-CLIP universal-restoration inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or as the timed CPU baseline), never as the shipped compute path.

The oracle is a from-scratch, functional fp32 PyTorch restatement of the
reference algorithm (the reference itself is pure Python on top of PyTorch, so
"plain torch ops on the reference state-dict" is the closest CPU statement of
its arithmetic).  Each function cites the reference file:line it follows.

Parity pinning: the reference repository ships no tests, golden vectors or
weights, so the oracle is pinned against outputs of the
reference modules themselves, generated in the build container by
``oracle/gen_golden.py`` (which imports /root/reference read-only) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` replays
them on every CPU test run.
"""
