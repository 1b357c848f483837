"""Core of the projectq stand-in (see package docstring).  TEST INFRASTRUCTURE ONLY."""
import os
import random as _random
import sys
from collections import deque

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import sv as _sv  # noqa: E402


class _Control:
    """Test hooks: forced measurement outcomes and an event log shared by all engines."""

    def __init__(self):
        self.forced = deque()   # bits consumed one per Measure command, in program order
        self.log = None         # when a list: receives ('gate'|'meas'|'alloc', ...) tuples
        self.rng = _random.Random()

    def reset(self, forced=(), log=None, seed=None):
        self.forced = deque(int(b) for b in forced)
        self.log = log
        if seed is not None:
            self.rng.seed(seed)


control = _Control()


class Command:
    def __init__(self, gate, ids):
        self.gate = gate
        self.ids = list(ids)

    def __repr__(self):
        return "Command({}, {})".format(self.gate, self.ids)


class DummyEngine:
    def __init__(self, save_commands=False):
        self.save_commands = save_commands
        self.received_commands = []

    def receive(self, cmd):
        if self.save_commands:
            self.received_commands.append(cmd)


class Simulator:
    """ProjectQ Simulator() semantics: default gate_fusion=False => one sweep per gate."""

    def __init__(self, gate_fusion=False, rnd_seed=None):
        self.n = 0
        self.psi = np.ones(1, dtype=np.complex128)
        self.pos = {}           # qubit id -> bit position
        self.results = {}       # qubit id -> last measured bit

    # -- allocation: k-th live qubit = bit k; new qubit = new MSB, zeros in the new half
    def allocate(self, qid):
        new = np.zeros(self.psi.size * 2, dtype=np.complex128)
        new[:self.psi.size] = self.psi
        self.psi = new
        self.pos[qid] = self.n
        self.n += 1

    def deallocate(self, qid):
        p = self.pos.pop(qid)
        view = self.psi.reshape(1 << (self.n - p - 1), 2, 1 << p)
        w1 = float(np.sum(np.abs(view[:, 1, :]) ** 2))
        w0 = float(np.sum(np.abs(view[:, 0, :]) ** 2))
        if min(w0, w1) > 1e-10:
            raise RuntimeError("Qubit {} is not in a classical state and cannot be deallocated".format(qid))
        self.psi = np.ascontiguousarray(view[:, 1 if w1 > w0 else 0, :]).reshape(-1)
        self.n -= 1
        for k in self.pos:
            if self.pos[k] > p:
                self.pos[k] -= 1

    def _sv(self):
        s = _sv.StateVector.__new__(_sv.StateVector)
        s.n, s.psi = self.n, self.psi
        return s

    def apply(self, gate, ids):
        bits = [self.pos[i] for i in ids]
        if control.log is not None:
            control.log.append(("gate", gate.name, getattr(gate, "angle", None), tuple(bits)))
        m = gate.matrix
        if len(bits) == 1:
            self._sv().apply_1q(m, bits[0])
        elif len(bits) == 2:
            self._sv().apply_2q(m, bits[0], bits[1])
        else:
            raise ValueError("gate on {} qubits unsupported".format(len(bits)))

    def measure(self, qid):
        bit = self.pos[qid]
        s = self._sv()
        if control.forced:
            out = control.forced.popleft()
        else:
            # ProjectQ: draw one uniform, pick a basis state by cumulative |amp|^2, read its bit
            pick = s.sample_basis(control.rng.random())
            out = (pick >> bit) & 1
        p = s.collapse(bit, out)
        if p <= 0.0:
            raise RuntimeError("forced outcome {} on bit {} has zero probability".format(out, bit))
        if control.log is not None:
            control.log.append(("meas", bit, out, p))
        self.results[qid] = out
        return out


class MainEngine:
    def __init__(self, backend=None, engine_list=None):
        self.backend = backend if backend is not None else Simulator()
        self.engine_list = list(engine_list) if engine_list else []
        self._next_id = 0

    def _send(self, cmd):
        for e in self.engine_list:
            e.receive(cmd)

    def allocate_qubit(self):
        return self.allocate_qureg(1)

    def allocate_qureg(self, n):
        reg = Qureg()
        for _ in range(n):
            qid = self._next_id
            self._next_id += 1
            self.backend.allocate(qid)
            self._send(Command(Allocate, [qid]))
            reg.append(Qubit(self, qid))
        return reg

    def flush(self):
        pass  # gates are executed eagerly

    def apply(self, gate, ids):
        self._send(Command(gate, ids))
        if isinstance(gate, MeasureGate):
            for i in ids:
                self.backend.measure(i)
        else:
            self.backend.apply(gate, ids)


class Qubit:
    def __init__(self, engine, qid):
        self.engine = engine
        self.id = qid

    def __int__(self):
        return int(self.engine.backend.results[self.id])

    def __bool__(self):
        return bool(int(self))

    def __del__(self):
        try:
            be = self.engine.backend
            if self.id in be.pos:
                be.deallocate(self.id)
        except RuntimeError:
            raise
        except Exception:
            pass


class Qureg(list):
    pass


def _flatten(qubits):
    if isinstance(qubits, Qubit):
        return [qubits]
    if isinstance(qubits, tuple):
        out = []
        for q in qubits:
            out.extend(_flatten(q))
        return out
    return [q for q in qubits]


class BasicGate:
    name = "?"

    def __or__(self, qubits):
        qs = _flatten(qubits)
        qs[0].engine.apply(self, [q.id for q in qs])

    def __str__(self):
        return self.name

    __repr__ = __str__


class _Fixed(BasicGate):
    pass


class XGate(_Fixed):
    name, matrix = "X", _sv.MAT_X


class YGate(_Fixed):
    name, matrix = "Y", _sv.MAT_Y


class ZGate(_Fixed):
    name, matrix = "Z", _sv.MAT_Z


class HGate(_Fixed):
    name, matrix = "H", _sv.MAT_H


class MeasureGate(BasicGate):
    name = "Measure"

    def __or__(self, qubits):
        for q in _flatten(qubits):
            q.engine.apply(self, [q.id])


class AllocateQubitGate(BasicGate):
    name = "Allocate"


class _Rotation(BasicGate):
    _mat = None

    def __init__(self, angle):
        self.angle = _sv.reduce_angle(angle)

    @property
    def matrix(self):
        return type(self)._mat(self.angle)

    def __str__(self):
        return "{}({})".format(self.name, self.angle)

    __repr__ = __str__


class Rx(_Rotation):
    name, _mat = "Rx", staticmethod(_sv.mat_rx)


class Ry(_Rotation):
    name, _mat = "Ry", staticmethod(_sv.mat_ry)


class Rz(_Rotation):
    name, _mat = "Rz", staticmethod(_sv.mat_rz)


class Rxx(_Rotation):
    name, _mat = "Rxx", staticmethod(_sv.mat_rxx)


class Rzz(_Rotation):
    name, _mat = "Rzz", staticmethod(_sv.mat_rzz)


class All:
    def __init__(self, gate):
        self.gate = gate

    def __or__(self, qureg):
        for q in _flatten(qureg):
            self.gate | q


X, Y, Z, H = XGate(), YGate(), ZGate(), HGate()
Measure = MeasureGate()
Allocate = AllocateQubitGate()
