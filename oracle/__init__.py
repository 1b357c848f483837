"""CPU parity oracle for the Surface-17 hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``surface_17_iqt_b200``) never does and fails loudly without its CUDA library.
"""
