"""`projectq`, as far as the Surface-17 reference uses it, on the B200 backend.

The implementation is surface_17_iqt_b200/pq_core.py (one module object, whichever name it is imported under)."""
from surface_17_iqt_b200.pq_core import MainEngine  # noqa: F401
