"""CPU oracle for the RRTEncoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product path (``rrt_mil_b200``) never imports
this package and raises if its CUDA library is missing.
"""
