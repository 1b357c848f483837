"""Stand-in for the `projectq` package, used ONLY to run the unmodified reference here.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Real ProjectQ is an un-vendored,
unpinned third-party dependency of the reference (README.md:4) and is absent from this image.
This shim exposes exactly the surface the reference touches:
``MainEngine(Simulator())``, ``allocate_qureg``, ``flush``, ``int(qubit)``,
``All, Measure, X, Y, Z, H, Rx, Ry, Rz, Rxx, Rzz`` and ``cengines.DummyEngine``, executing
each gate as one CPU sweep through oracle/sv.py with ProjectQ's published conventions.
It adds two test hooks ProjectQ does not have: a forced-outcome queue and an event log
(``projectq.control``).
"""
from ._core import MainEngine, control  # noqa: F401
