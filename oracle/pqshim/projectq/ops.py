from ._core import (All, Measure, X, Y, Z, H, Rx, Ry, Rz, Rxx, Rzz,  # noqa: F401
                    XGate, YGate, ZGate, HGate, MeasureGate, AllocateQubitGate, Allocate)
