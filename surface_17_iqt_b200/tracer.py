"""Circuit tracer: run a module's OWN `stabilizer_cycle_error_index` once against a recording engine and read the
compiled cycle off the command stream -- the 24 entangling operations (qubits, signs, which single-qubit gates survived
the hand cancellation, leak-predicate indices, timestep) and, per error entry, the probability list and slot it draws
from and what it does when it fires.

The module can be the reference's compiled_surface_code_arbitrary_error_model.py (imported unmodified, with
`surface_17_iqt_b200.pq_compat` providing `projectq`) or this package's module of the same name.  The traced table is
what ``circuit.compile_program(..., ops=, slots=)`` lowers to the device program, so the reference's quirks (the leak
indices of the second operation of timestep 5, CS:559-561; the control/target swap between timesteps 1-4 and 5-8,
CS:297 vs 326; `sympathetic_cooling` ignoring stab_ind, CS:767) are observed rather than restated:
tests/test_tracer.py asserts that the trace of the reference equals ``circuit.build_cycle()`` / ``circuit.slot_index``
and that both lower to byte-identical programs.

How the per-shot decisions are made visible without touching the traced code:
  * every probability list is a `ProbeList`: `prob[index]` returns a probe whose comparison with the uniform draw
    (`random.random() < prob[index]`, CS:198) logs (list, index) and answers "fire", so every entry shows its effect;
  * the leak register is a `LeakProbe`: reads (the skip predicate, CS:293, 322) and writes (insert_leakage, CS:286) are
    logged, the stored flags stay 0 so no Rxx is skipped;
  * the module's `stabiliser_timestep_*` functions are wrapped for the duration of the trace to mark timestep borders.
No device is needed (`RecordingBackend`).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from . import circuit, engine, pq_core


class _Probe:
    def __init__(self, log, name, index):
        self.log, self.name, self.index = log, name, index

    def __gt__(self, other):            # random.random() < probe
        self.log.append(("draw", self.name, self.index))
        return True

    def __ne__(self, other):            # `prob[index] != 0.0` of a coherent entry: keep the rotation out of the trace
        self.log.append(("draw", self.name, self.index))
        return False

    __hash__ = None


class ProbeList(list):
    def __init__(self, log, name, length):
        super().__init__(length * [0])
        self.log, self.name = log, name

    def __getitem__(self, index):
        if not 0 <= index < len(self):
            raise IndexError("slot {} of {} (length {})".format(index, self.name, len(self)))
        return _Probe(self.log, self.name, index)


class LeakProbe(list):
    def __init__(self, log):
        super().__init__(17 * [0])
        self.log = log

    def __getitem__(self, index):
        self.log.append(("leak_read", index))
        return 0

    def __setitem__(self, index, value):
        self.log.append(("leak_set", index))


@dataclass
class TracedCycle:
    ops: List[circuit.EntanglingOp] = field(default_factory=list)
    # (operation index, gate position 'Ry1' | 'Rxx' | 'Rx' | 'Ry2', entry number) -> (list name, slot, effect)
    # effect: tuple of ('X' | 'Y' | 'Z', qubit) with qubit = 'd' / 'a', or ('leak', register index), or ('depol',)
    entries: Dict[Tuple[int, str, int], Tuple[str, int, tuple]] = field(default_factory=dict)
    idle: List[Tuple[int, int, str, int]] = field(default_factory=list)     # (timestep, qubit 0..16, list name, slot)
    n_gates: int = 0
    log: list = field(default_factory=list)

    def slots(self):
        return {k: v[1] for k, v in self.entries.items()}


def record(cs, model, list_len, stab_ind=0, cooling=False):
    """Run `cs.stabilizer_cycle_error_index` once on a recording engine; returns the raw event log."""
    names = list(model["probabilities"])
    if len(names) != len(list_len):
        raise ValueError("one list length per probability name")
    log = []
    backend = pq_core.RecordingBackend()
    backend.log = log
    eng = pq_core.MainEngine(backend)
    data, ancilla = eng.allocate_qureg(9), eng.allocate_qureg(8)
    probs = {n: ProbeList(log, n, length) for n, length in zip(names, list_len)}
    wrapped = {}

    def mark(name, t):
        orig = getattr(cs, name)

        def wrapper(*a, **k):
            log.append(("timestep", a[0] if t is None else t))
            return orig(*a, **k)
        wrapped[name] = orig
        setattr(cs, name, wrapper)

    for t in range(1, 9):
        if hasattr(cs, "stabiliser_timestep_{}".format(t)):
            mark("stabiliser_timestep_{}".format(t), t)
    if not wrapped and hasattr(cs, "stabiliser_timestep"):
        mark("stabiliser_timestep", None)
    try:
        cs.stabilizer_cycle_error_index(data, ancilla, LeakProbe(log), eng, model, probs, reset=True, stab_ind=stab_ind,
                                        cooling=cooling)
    finally:
        for name, orig in wrapped.items():
            setattr(cs, name, orig)
    return log


def trace_cycle(cs, model, list_len, stab_ind=0, cooling=False) -> TracedCycle:
    log = record(cs, model, list_len, stab_ind, cooling)
    plus, minus = engine.reduce_angle(math.pi / 2), engine.reduce_angle(-math.pi / 2)
    out = TracedCycle(log=log)
    gates = model["gates"]
    t_now, rx = 0, 0
    ry = 0
    open_bracket = set()            # data qubits between their Ry(+pi/2) and Ry(-pi/2)
    cur = None                      # operation being assembled
    reads = []                      # leak-register reads since the last gate
    last = None                     # (operation index, gate position, model gate name, qubit ids) entries attach to
    n_entry = 0
    pending_effect = None
    pending_idle = None

    def close_entry():
        nonlocal pending_effect, pending_idle
        if pending_effect is not None:
            key, name, slot, eff = pending_effect
            out.entries[key] = (name, slot, tuple(eff))
            pending_effect = None
        if pending_idle is not None:
            raise ValueError("an Idle entry fired without touching a qubit")

    def new_op(d):
        nonlocal cur
        cur = dict(timestep=t_now - 1, index=len(out.ops), data=d, ancilla=None, s=0, has_ry1=False, has_rx=False,
                   has_ry2=False, d_leak=None, a_leak=None, rx_ind=-1, ry1_ind=-1, ry2_ind=-1, kind=None)

    def finish_op():
        nonlocal cur
        if cur is not None:
            if cur["ancilla"] is None:
                raise ValueError("operation without an Rxx in the traced stream")
            out.ops.append(circuit.EntanglingOp(cur["timestep"], cur["index"], cur["kind"], cur["data"], cur["ancilla"],
                                                cur["s"], cur["has_ry1"], cur["has_rx"], cur["has_ry2"], cur["d_leak"],
                                                cur["a_leak"], cur["rx_ind"], cur["ry1_ind"], cur["ry2_ind"]))
            cur = None

    for ev in log:
        kind = ev[0]
        if kind == "timestep":
            close_entry()
            finish_op()
            t_now = ev[1]
            last = None
        elif kind == "leak_read":
            close_entry()
            reads.append(ev[1])
        elif kind == "meas":
            close_entry()
            finish_op()
            last = None
        elif kind == "gate":
            _, name, angle, ids = ev
            if name == "Z" and pending_idle is not None:
                out.idle.append((pending_idle[0], ids[0], pending_idle[1], pending_idle[2]))
                pending_idle = None
                continue
            if name in ("X", "Y", "Z") and pending_effect is not None:
                # the Pauli an entry applies when it fires; qubit ids 0-8 are data, 9-16 ancilla (allocation order)
                role = "d" if ids[0] == cur["data"] and ids[0] < 9 else ("a" if ids[0] == 9 + cur["ancilla"] else "?")
                pending_effect[3].append((name, role) if last is not None else (name, ids[0]))
                continue
            close_entry()
            out.n_gates += 1
            if name == "Ry" and angle == plus:
                finish_op()
                new_op(ids[0])
                cur["has_ry1"], cur["ry1_ind"] = True, ry
                ry += 1
                open_bracket.add(ids[0])
                last = (cur["index"], "Ry1", "Ry", ids)
            elif name == "Rxx":
                if cur is None or cur["ancilla"] is not None or cur["data"] != ids[0]:
                    finish_op()
                    new_op(ids[0])
                if len(reads) < 2:
                    raise ValueError("Rxx without a leak predicate in front of it")
                cur["d_leak"], cur["a_leak"] = reads[-2], reads[-1]
                cur["ancilla"] = ids[1] - 9
                cur["s"] = 1 if angle == plus else (-1 if angle == minus else 0)
                if cur["s"] == 0:
                    raise ValueError("Rxx angle {} is not +-pi/2".format(angle))
                cur["kind"] = "z" if ids[0] in open_bracket else "x"
                last = (cur["index"], "Rxx", "Rxx", ids)
            elif name == "Rx":
                if cur is None or cur["data"] != ids[0] or cur["ancilla"] is None:
                    raise ValueError("Rx outside an entangling operation")
                if angle != (minus if cur["s"] > 0 else plus):
                    raise ValueError("Rx angle does not undo the Rxx sign")
                cur["has_rx"], cur["rx_ind"] = True, rx
                rx += 1
                last = (cur["index"], "Rx", "Rx", ids)
            elif name == "Ry" and angle == minus:
                if cur is None or cur["data"] != ids[0] or cur["ancilla"] is None:
                    raise ValueError("Ry(-pi/2) outside an entangling operation")
                cur["has_ry2"], cur["ry2_ind"] = True, ry
                ry += 1
                open_bracket.discard(ids[0])
                last = (cur["index"], "Ry2", "Ry", ids)
            elif name == "X" and last is None:
                out.n_gates -= 1            # ancilla reset after the measurement (CS:859-862)
            else:
                raise ValueError("unexpected gate {} {} in the cycle".format(name, angle))
            reads = []
            n_entry = 0
        elif kind == "draw":
            close_entry()
            _, pname, slot = ev
            if last is None or n_entry >= len(gates[last[2]]):
                # not an entry of the last gate: an Idle entry of sympathetic_cooling (CS:762-772); the Z it applies
                # when it fires names the qubit
                pending_idle = (t_now - 1, pname, slot)
                continue
            op_index, pos, gname, _ids = last
            entry = gates[gname][n_entry]
            if entry[1] != pname:
                raise ValueError("entry {} of gate {} drew from {} instead of {}".format(n_entry, gname, pname, entry[1]))
            pending_effect = [(op_index, pos, n_entry), pname, slot, [("depol",)] if entry[0] == "depol" else []]
            if entry[0] == "depol":
                close_entry()
            n_entry += 1
        elif kind == "leak_set":
            if pending_effect is None:
                raise ValueError("leak flag set outside an error entry")
            pending_effect[3].append(("leak", ev[1]))
    close_entry()
    finish_op()
    # after the last timestep of the x-type half the cooling gap events of the trace carried last = None only if a
    # timestep marker preceded them: make sure the markers were there
    if cooling and not out.idle:
        raise ValueError("cooling was requested but no Idle draw was seen")
    return out


def check_effects(traced: TracedCycle, model):
    """The device program assumes what each error kind does (circuit.compile_program / k_plan): X, Y, Z and Zc on the data
    qubit, Zt on the ancilla, XX / ZZ on both, L on d_leak, Lc / Lt on the control / target index of the operation's
    kind.  Raise if the traced module does something else."""
    ops = {o.index: o for o in traced.ops}
    for (oi, pos, k), (pname, slot, eff) in traced.entries.items():
        op = ops[oi]
        gname = {"Ry1": "Ry", "Ry2": "Ry", "Rxx": "Rxx", "Rx": "Rx"}[pos]
        err = model["gates"][gname][k][0]
        c_ind, t_ind = (op.a_leak, op.d_leak) if op.kind == "x" else (op.d_leak, op.a_leak)
        want = {"X": (("X", "d"),), "Y": (("Y", "d"),), "Z": (("Z", "d"),), "Zc": (("Z", "d"),), "Zt": (("Z", "a"),),
                "XX": (("X", "d"), ("X", "a")), "ZZ": (("Z", "d"), ("Z", "a")), "L": (("leak", op.d_leak),),
                "Lc": (("leak", c_ind),), "Lt": (("leak", t_ind),), "LL": (("leak", c_ind), ("leak", t_ind)),
                "depol": (("depol",),)}.get(err)
        if want is None or tuple(eff) != want:
            raise ValueError("entry {} ({}) of {} in operation {} does {} -- the program assumes {}".format(
                k, err, pos, oi, eff, want))
    return True
