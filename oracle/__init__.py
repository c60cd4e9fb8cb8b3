"""CPU oracle for the DIRAL hot path -- TEST INFRASTRUCTURE, not product code.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  Parity status: pinned against outputs of the unmodified
reference (tests/golden/*.npz, produced by tests/golden/make_golden.py).
"""
