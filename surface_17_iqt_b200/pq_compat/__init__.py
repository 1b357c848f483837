"""A `projectq`-shaped package over libs17.so.

`sys.path.insert(0, surface_17_iqt_b200.pq_compat.PATH)` makes `import projectq` resolve to the package in this
directory: `MainEngine`, `Simulator`, `DummyEngine` and the gate objects the reference uses
(compiled_surface_code_arbitrary_error_model.py:5, calc_log_e_rate_arbitrary_error_model.py:6-7,
test_logical_qubit.py:2-4), executing on the B200 through the engine-level C ABI (`s17_eng_*`).
"""
import os

PATH = os.path.dirname(os.path.abspath(__file__))


def install():
    """Put the projectq-shaped package first on sys.path (idempotent)."""
    import sys
    if PATH not in sys.path:
        sys.path.insert(0, PATH)
