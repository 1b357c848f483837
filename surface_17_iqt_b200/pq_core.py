"""The projectq-shaped face (re-exported as `projectq` by pq_compat/projectq): engines, qubits and gate objects with ProjectQ's call syntax.

Surface the reference uses: `MainEngine(Simulator())` or `MainEngine(backend, engine_list)`,
`eng.allocate_qureg(n)`, `eng.flush()`, `int(qubit)`, `Gate | qubit`, `Gate | (q0, q1)`, `All(Gate) | qureg` for
X, Y, Z, H, Rx, Ry, Rz, Rxx, Rzz, Measure; tests add `DummyEngine(save_commands=True).received_commands`.

Backends:
  Simulator         the GPU: every `|` is one sweep over a batch of state vectors (surface_17_iqt_b200.engine), every
                    Measure a reduction + collapse.  batch > 1 runs N copies in lock-step by default (the outcome drawn
                    for shot 0 is imposed on all, so `int(qubit)` has one value); `lockstep=False` samples per shot and
                    `int(qubit)` raises if the shots disagree (`Simulator.outcomes(qubit)` gives the per-shot bits).
  RecordingBackend  no device: records the command stream (the tracer, surface_17_iqt_b200.tracer); Measure returns 0
                    or the next forced bit.
"""
import random as _random
from collections import deque

import numpy as np

from surface_17_iqt_b200 import _lib as _L
from surface_17_iqt_b200 import engine as _engine


class Command:
    def __init__(self, gate, ids):
        self.gate = gate
        self.ids = list(ids)

    def __repr__(self):
        return "Command({}, {})".format(self.gate, self.ids)


class DummyEngine:
    """ProjectQ's test helper: keeps the commands that pass through it."""

    def __init__(self, save_commands=False):
        self.save_commands = save_commands
        self.received_commands = []

    def receive(self, cmd):
        if self.save_commands:
            self.received_commands.append(cmd)


class NotClassicalError(RuntimeError):
    pass


class Simulator:
    """ProjectQ's `Simulator()` (gate_fusion=False: one sweep per gate) for `batch` state vectors on one B200."""

    def __init__(self, gate_fusion=False, rnd_seed=None, batch=1, n_qubits=17, device=None, lockstep=True):
        if gate_fusion:
            raise NotImplementedError("gate_fusion=True is not offered by the engine-level face (the reference never "
                                      "sets it); the fused path is surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model")
        self.batch, self.capacity, self.lockstep = int(batch), int(n_qubits), bool(lockstep)
        self.device = device
        self.seed = _random.getrandbits(64) if rnd_seed is None else int(rnd_seed)
        self._eng = None
        self.pos = {}                 # qubit id -> state-vector bit
        self.pending = []             # allocated, no bit yet (all bits in use): resolved when bits are freed / on first use
        self.free = list(range(self.capacity))
        self._out = {}                # qubit id -> per-shot outcome of its last Measure
        self.forced = deque()         # test hook: bits consumed one per Measure command, in program order
        self.impossible = 0           # Measure commands that imposed an outcome of probability zero on some shot

    # -- device ---------------------------------------------------------------------------------------------------
    @property
    def eng(self):
        if self._eng is None:
            import os
            dev = int(os.environ.get("LOCAL_RANK", "0")) if self.device is None else int(self.device)
            self._eng = _engine.EngineBackend(self.capacity, self.batch, dev, self.seed)
        return self._eng

    def close(self):
        """Free the device context; qubits that are still alive are dropped without the classical-state check."""
        if self._eng is not None:
            self._eng.close()
            self._eng = None
        self.pos.clear()
        self.pending.clear()
        self.free = list(range(self.capacity))

    # -- allocation: the k-th live qubit sits on bit k (ProjectQ: a new qubit is the new most significant bit) ------
    def allocate(self, qid):
        self.eng                       # a Simulator without a CUDA device fails here, loudly
        if self.free:
            self.pos[qid] = self.free.pop(0)
        else:
            self.pending.append(qid)

    def _place(self, qid):
        if qid in self.pos:
            return self.pos[qid]
        if qid in self.pending and self.free:
            self.pending.remove(qid)
            self.pos[qid] = self.free.pop(0)
            return self.pos[qid]
        raise _L.S17Error("qubit {} needs a state-vector bit but all {} are in use".format(qid, self.capacity))

    def deallocate(self, qid):
        if qid in self.pending:
            self.pending.remove(qid)
            return
        if qid not in self.pos:
            return
        bit = self.pos[qid]
        p1 = self.eng.prob1(bit)
        if np.any(np.minimum(p1, 1.0 - p1) > 1e-10):
            raise NotClassicalError("Qubit {} is not in a classical state and cannot be deallocated".format(qid))
        ones = p1 > 0.5
        if np.any(ones):               # return the bit to |0> before it is handed to another qubit
            self.eng.gate1(bit, _engine.MAT_X, pred=ones.astype(np.uint8))
        del self.pos[qid]
        self._out.pop(qid, None)
        self.free.append(bit)
        self.free.sort()
        while self.pending and self.free:
            self.pos[self.pending.pop(0)] = self.free.pop(0)

    # -- commands -------------------------------------------------------------------------------------------------
    def apply(self, gate, ids, pred=None):
        bits = [self._place(i) for i in ids]
        if len(bits) == 1:
            self.eng.gate1(bits[0], gate.matrix, pred)
        elif len(bits) == 2:
            self.eng.gate2(bits[0], bits[1], gate.matrix, pred)
        else:
            raise ValueError("gate on {} qubits unsupported".format(len(bits)))

    def measure(self, qid):
        bit = self._place(qid)
        if self.forced:
            out, prob, bad = self.eng.measure(bit, forced=self.forced.popleft())
        else:
            out, prob, bad = self.eng.measure(bit, _L.MEAS_LOCKSTEP if self.lockstep else _L.MEAS_SAMPLE)
        self.impossible += int(bad > 0)
        self._out[qid] = out
        return out

    def result(self, qid):
        out = self._out[qid]
        if np.any(out != out[0]):
            raise _L.S17Error("int(qubit): the shots of this batch measured different values; use "
                              "Simulator.outcomes(qubit), or lockstep=True")
        return int(out[0])

    def outcomes(self, qubit):
        return self._out[qubit.id].copy()

    def flush(self):
        if self._eng is not None:
            self._eng.flush()

    def state(self, shot=0):
        """cheat(): the full state vector of one shot, bit k = qubit on position k."""
        return self.eng.state(shot)


class RecordingBackend:
    """No device: every command is appended to `log` as ('gate', name, angle-or-None, qubit ids) / ('meas', qubit id)."""

    def __init__(self, forced=()):
        self.log = []
        self.forced = deque(int(b) for b in forced)
        self._res = {}

    def allocate(self, qid):
        pass

    def deallocate(self, qid):
        pass

    def apply(self, gate, ids, pred=None):
        self.log.append(("gate", gate.name, getattr(gate, "angle", None), tuple(ids)))

    def measure(self, qid):
        out = self.forced.popleft() if self.forced else 0
        self.log.append(("meas", qid, out))
        self._res[qid] = out
        return out

    def result(self, qid):
        return int(self._res[qid])

    def flush(self):
        pass


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
        self.backend.flush()

    def apply(self, gate, ids, pred=None):
        self._send(Command(gate, ids))
        if isinstance(gate, MeasureGate):
            for i in ids:
                self.backend.measure(i)
        else:
            self.backend.apply(gate, ids, pred)

    def release(self, qid):
        self._send(Command(Deallocate, [qid]))
        self.backend.deallocate(qid)


class Qubit:
    def __init__(self, engine, qid):
        self.engine = engine
        self.id = qid

    def __int__(self):
        return int(self.engine.backend.result(self.id))

    def __bool__(self):
        return bool(int(self))

    def __del__(self):
        try:
            self.engine.release(self.id)
        except NotClassicalError:
            raise
        except Exception:      # interpreter shutdown, closed device context
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


class XGate(BasicGate):
    name, matrix = "X", _engine.MAT_X


class YGate(BasicGate):
    name, matrix = "Y", _engine.MAT_Y


class ZGate(BasicGate):
    name, matrix = "Z", _engine.MAT_Z


class HGate(BasicGate):
    name, matrix = "H", _engine.MAT_H


class MeasureGate(BasicGate):
    name = "Measure"

    def __or__(self, qubits):
        for q in _flatten(qubits):
            q.engine.apply(self, [q.id])


class AllocateQubitGate(BasicGate):
    name = "Allocate"


class DeallocateQubitGate(BasicGate):
    name = "Deallocate"


class _Rotation(BasicGate):
    def __init__(self, angle):
        self.angle = _engine.reduce_angle(angle)

    @property
    def matrix(self):
        return _engine.rotation_matrix(self.name, self.angle)

    def __str__(self):
        return "{}({})".format(self.name, self.angle)

    __repr__ = __str__


class Rx(_Rotation):
    name = "Rx"


class Ry(_Rotation):
    name = "Ry"


class Rz(_Rotation):
    name = "Rz"


class Rxx(_Rotation):
    name = "Rxx"


class Rzz(_Rotation):
    name = "Rzz"


class All:
    def __init__(self, gate):
        self.gate = gate

    def __or__(self, qureg):
        for q in _flatten(qureg):
            self.gate | q


X, Y, Z, H = XGate(), YGate(), ZGate(), HGate()
Measure = MeasureGate()
Allocate = AllocateQubitGate()
Deallocate = DeallocateQubitGate()
