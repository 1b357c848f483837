"""Python face of the engine-level C ABI (`s17_eng_*`, include/s17.h): ProjectQ's Simulator backend for a batch of
state vectors -- one sweep per gate, P(1) reduction + collapse per Measure.

This is what `MainEngine(Simulator())` is in the reference (calc_log_e_rate_arbitrary_error_model.py:63, 221, 299, 392);
`surface_17_iqt_b200/pq_compat/projectq` wraps it in ProjectQ's class names.  No CPU fallback: creating an engine
without a CUDA device raises.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L

SQ = 1.0 / math.sqrt(2.0)
MAT_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
MAT_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
MAT_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
MAT_H = np.array([[SQ, SQ], [SQ, -SQ]], dtype=np.complex128)
ANGLE_PRECISION = 12           # ProjectQ's BasicRotationGate keeps round(theta mod 4 pi, 12)


def reduce_angle(theta):
    a = round(float(theta) % (4.0 * math.pi), ANGLE_PRECISION)
    return 0.0 if a > 4.0 * math.pi - 10.0 ** (-ANGLE_PRECISION) else a


def rotation_matrix(name, theta):
    """Rx / Ry / Rz / Rxx / Rzz(theta) in ProjectQ's conventions: exp(-i theta/2 P)."""
    c, s = math.cos(0.5 * theta), math.sin(0.5 * theta)
    if name == "Rx":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "Ry":
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == "Rz":
        return np.array([[c - 1j * s, 0], [0, c + 1j * s]], dtype=np.complex128)
    if name == "Rxx":
        m = np.zeros((4, 4), dtype=np.complex128)
        for i in range(4):
            m[i, i] = c
            m[i, 3 - i] = -1j * s
        return m
    if name == "Rzz":
        return np.diag([c - 1j * s, c + 1j * s, c + 1j * s, c - 1j * s]).astype(np.complex128)
    raise ValueError(name)


class EngineBackend:
    """`batch` state vectors of 2^n_bits complex128 amplitudes on one GPU; bit k of the index = k-th qubit."""

    def __init__(self, n_bits=17, batch=1, device=0, seed=0):
        self._lib = L.load()
        self._h = C.c_void_p()
        self.n_bits, self.batch, self.device = int(n_bits), int(batch), int(device)
        rc = self._lib.s17_eng_create(C.byref(self._h), self.device, self.n_bits, self.batch, int(seed) & (2 ** 64 - 1))
        if rc != 0:
            msg = self._lib.s17_eng_last_error(None).decode()
            self._h = None
            raise L.S17Error(msg)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.s17_eng_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise L.S17Error(self._lib.s17_eng_last_error(self._h).decode())

    @staticmethod
    def _pred(pred):
        if pred is None:
            return None, None
        p = np.ascontiguousarray(pred, dtype=np.uint8)
        return p, p.ctypes.data

    def reset(self):
        self._ck(self._lib.s17_eng_reset(self._h))

    def gate1(self, bit, matrix, pred=None):
        m = np.ascontiguousarray(matrix, dtype=np.complex128).reshape(2, 2)
        keep, pp = self._pred(pred)
        self._ck(self._lib.s17_eng_gate1(self._h, int(bit), m.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), pp))

    def gate2(self, bit0, bit1, matrix, pred=None):
        m = np.ascontiguousarray(matrix, dtype=np.complex128).reshape(4, 4)
        keep, pp = self._pred(pred)
        self._ck(self._lib.s17_eng_gate2(self._h, int(bit0), int(bit1),
                                         m.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), pp))

    def measure(self, bit, mode=L.MEAS_SAMPLE, forced=None):
        """Measure | qubit on every shot.  Returns (outcomes uint8[batch], probability of each outcome, n_impossible)."""
        out = np.zeros(self.batch, dtype=np.uint8)
        prob = np.zeros(self.batch, dtype=np.float64)
        bad = C.c_uint32()
        f = None
        if forced is not None:
            f = np.ascontiguousarray(np.broadcast_to(np.asarray(forced, dtype=np.uint8), (self.batch,)))
            mode = L.MEAS_FORCED
        self._ck(self._lib.s17_eng_measure(self._h, int(bit), int(mode), None if f is None else f.ctypes.data,
                                           out.ctypes.data, prob.ctypes.data_as(C.POINTER(C.c_double)), C.byref(bad)))
        return out, prob, int(bad.value)

    def prob1(self, bit):
        p = np.zeros(self.batch, dtype=np.float64)
        self._ck(self._lib.s17_eng_prob1(self._h, int(bit), p.ctypes.data_as(C.POINTER(C.c_double))))
        return p

    def flush(self):
        self._ck(self._lib.s17_eng_flush(self._h))

    def state(self, shot=0):
        buf = np.zeros(2 << self.n_bits, dtype=np.float64)
        self._ck(self._lib.s17_eng_read_state(self._h, int(shot), buf.ctypes.data_as(C.POINTER(C.c_double))))
        return buf.view(np.complex128)

    def timing(self, clear=False):
        """Device time and launch counts per kernel class since the last clear, plus the algorithmic HBM bytes of one
        launch: a gate layer reads and writes the state once (4 194 304 B per shot at 17 qubits); a
        measurement layer reads it once for the P(1) reduction (2 097 152 B) and, for the collapse, reads the surviving
        half and writes both halves (3 145 728 B -- 8d budgets 4 194 304 B for it; the discarded half is never read)."""
        ms = (C.c_double * 3)()
        calls = (C.c_uint64 * 3)()
        self._ck(self._lib.s17_eng_timing(self._h, ms, calls, int(bool(clear))))
        state_bytes = self.batch * (16 << self.n_bits)
        return {"gate_ms": ms[0], "reduce_ms": ms[1], "collapse_ms": ms[2], "gate_calls": calls[0],
                "reduce_calls": calls[1], "collapse_calls": calls[2], "gate_bytes_per_call": 2 * state_bytes,
                "reduce_bytes_per_call": state_bytes, "collapse_bytes_per_call": 3 * state_bytes // 2}
