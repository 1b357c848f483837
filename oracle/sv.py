"""ctypes face of oracle/csrc/sv.c (ProjectQ-style one-sweep-per-gate CPU state vector).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Gate matrices follow ProjectQ's
published definitions; the reference reaches them through
``from projectq.ops import All, Measure, X, Y, Z, H, Rx, Ry, Rxx, Rzz``
(compiled_surface_code_arbitrary_error_model.py:5).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle_sv.so with the committed Makefile (gcc only)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_sv.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        cp = ctypes.c_void_p
        L.sv_apply_1q.argtypes = [cp, ctypes.c_int, ctypes.c_int, cp]
        L.sv_apply_2q.argtypes = [cp, ctypes.c_int, ctypes.c_int, ctypes.c_int, cp]
        L.sv_prob_one.argtypes = [cp, ctypes.c_int, ctypes.c_int]
        L.sv_prob_one.restype = ctypes.c_double
        L.sv_sample_basis.argtypes = [cp, ctypes.c_int, ctypes.c_double]
        L.sv_sample_basis.restype = ctypes.c_uint64
        L.sv_collapse.argtypes = [cp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.sv_collapse.restype = ctypes.c_double
        L.sv_init_zero.argtypes = [cp, ctypes.c_int]
        _LIB = L
    return _LIB


ANGLE_PRECISION = 12


def reduce_angle(theta):
    """ProjectQ BasicRotationGate: angle stored as round(theta mod 4pi, 12); ~4pi -> 0."""
    a = round(float(theta) % (4.0 * math.pi), ANGLE_PRECISION)
    if a > 4.0 * math.pi - 10.0 ** (-ANGLE_PRECISION):
        a = 0.0
    return a


def mat_rx(theta):
    c, s = math.cos(0.5 * theta), math.sin(0.5 * theta)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def mat_ry(theta):
    c, s = math.cos(0.5 * theta), math.sin(0.5 * theta)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def mat_rz(theta):
    return np.array([[np.exp(-0.5j * theta), 0], [0, np.exp(0.5j * theta)]], dtype=np.complex128)


def mat_rxx(theta):
    c, s = math.cos(0.5 * theta), math.sin(0.5 * theta)
    m = np.zeros((4, 4), dtype=np.complex128)
    for i in range(4):
        m[i, i] = c
        m[i, 3 - i] = -1j * s
    return m


def mat_rzz(theta):
    e0, e1 = np.exp(-0.5j * theta), np.exp(0.5j * theta)
    return np.diag([e0, e1, e1, e0]).astype(np.complex128)


MAT_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
MAT_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
MAT_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
MAT_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2.0)


class StateVector:
    """n-qubit complex128 register; every method is one full sweep, like ProjectQ's kernels."""

    def __init__(self, n):
        self.n = n
        self.psi = np.zeros(1 << n, dtype=np.complex128)
        lib().sv_init_zero(self.psi.ctypes.data, n)

    def apply_1q(self, m, q):
        m = np.ascontiguousarray(m, dtype=np.complex128)
        lib().sv_apply_1q(self.psi.ctypes.data, self.n, int(q), m.ctypes.data)

    def apply_2q(self, m, q0, q1):
        m = np.ascontiguousarray(m, dtype=np.complex128)
        lib().sv_apply_2q(self.psi.ctypes.data, self.n, int(q0), int(q1), m.ctypes.data)

    def prob_one(self, q):
        return lib().sv_prob_one(self.psi.ctypes.data, self.n, int(q))

    def sample_basis(self, rnd):
        return int(lib().sv_sample_basis(self.psi.ctypes.data, self.n, float(rnd)))

    def collapse(self, q, bit):
        return lib().sv_collapse(self.psi.ctypes.data, self.n, int(q), int(bit))
