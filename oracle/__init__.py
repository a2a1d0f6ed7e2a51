"""CPU oracle for the forest pair-counting path: TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import this package, and only as the checker or the timed CPU baseline.  The product
(``picca_b200``) never imports it.
"""
