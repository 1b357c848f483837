from surface_17_iqt_b200.pq_core import (All, Allocate, Deallocate, H, Measure, Rx, Rxx, Ry, Rz, Rzz, X, Y, Z,  # noqa: F401
                                         BasicGate, HGate, MeasureGate, XGate, YGate, ZGate)
