"""Host-side compiler: Surface-17 cycle + error model  ->  device program for libs17.so.

The reference describes the compiled Molmer-Sorensen syndrome-extraction cycle as Python functions
that emit ProjectQ gates one by one (compiled_surface_code_arbitrary_error_model.py [CS]
``stabiliser_timestep_1..8`` :365-650, ``x_type_entangling`` :291, ``z_type_entangling`` :312) and
decide error insertion per gate at run time (``insert_errors`` :127, ``insert_error`` :188).  Here
the same cycle is *data*: a table of 24 entangling operations that is compiled once into

* sweeps (``Layer``): which operations one pass over the state vector fuses, and
* an injection program (``PInstr`` list): the per-shot decisions of insert_errors/insert_error in
  the reference's draw order, evaluated on device by ``k_plan``.

Qubit -> state-vector bit: data[i] = bit i, ancilla[j] = bit 9 + j (ProjectQ allocation order,
CL:302-303).
"""
import math
from dataclasses import dataclass, field
from typing import List

from . import _lib as L

N_DATA, N_ANC = 9, 8
NUM_GATES_MS = {"Rx": 12, "Ry": 18, "Rxx": 24}      # CS:129
FIELD_OF_GATE = {"Ry1": 0, "Rxx": 6, "Rx": 12, "Ry2": 18}


@dataclass
class EntanglingOp:
    """One x_type_entangling / z_type_entangling call of the reference, as data."""
    timestep: int          # 0..7
    index: int             # 0..23 = running Rxx index (rxx_ind, CS:302, 331)
    kind: str              # 'x' (timesteps 1-4) or 'z' (5-8)
    data: int
    ancilla: int
    s: int                 # sign of the Rxx angle
    has_ry1: bool
    has_rx: bool
    has_ry2: bool
    d_leak: int            # leak-register indices the reference passes (CS: d_leak_ind / a_leak_ind)
    a_leak: int
    rx_ind: int = -1       # running indices of the 1q gates this op owns (-1: gate cancelled)
    ry1_ind: int = -1
    ry2_ind: int = -1

    @property
    def data_bit(self):
        return self.data

    @property
    def anc_bit(self):
        return N_DATA + self.ancilla


# (data, ancilla) pairs per timestep, sign of the Rxx angle per timestep, and which of the
# single-qubit gates survive the hand cancellation (CS:365-650).  'k' = kept, '.' = cancelled,
# letters in order Ry(before) Rx Ry(after).
_PAIRS = (
    ((4, 1), (8, 6), (6, 4)), ((1, 1), (5, 6), (3, 4)), ((3, 1), (7, 6), (5, 3)), ((0, 1), (4, 6), (2, 3)),
    ((1, 2), (3, 5), (7, 7)), ((2, 2), (4, 5), (8, 7)), ((4, 2), (6, 5), (0, 0)), ((5, 2), (7, 5), (1, 0)),
)
_SIGN = (+1, -1, +1, -1, +1, -1, +1, -1)
_KEPT = (
    ("...", ".k.", ".k."), (".k.", "...", "..."), ("...", ".k.", "..."), (".k.", "...", ".k."),
    ("k..", "kkk", "k.."), ("kkk", "k..", "kkk"), ("..k", "kkk", "kkk"), ("kkk", "..k", "..k"),
)
# leak-register indices are (data, 9 + ancilla) except for the second operation of timestep 5, where
# the reference passes the indices of the FIRST operation (CS:559-561).
_LEAK_OVERRIDE = {(4, 1): (1, 2 + 9)}


def build_cycle() -> List[EntanglingOp]:
    ops = []
    rx = ry = 0
    for t in range(8):
        for j, (d, a) in enumerate(_PAIRS[t]):
            kept = _KEPT[t][j]
            dl, al = _LEAK_OVERRIDE.get((t, j), (d, N_DATA + a))
            op = EntanglingOp(t, len(ops), "x" if t < 4 else "z", d, a, _SIGN[t], kept[0] == "k", kept[1] == "k",
                              kept[2] == "k", dl, al)
            if op.has_ry1:
                op.ry1_ind, ry = ry, ry + 1
            if op.has_rx:
                op.rx_ind, rx = rx, rx + 1
            if op.has_ry2:
                op.ry2_ind, ry = ry, ry + 1
            ops.append(op)
    assert (rx, ry, len(ops)) == (12, 18, 24)
    return ops


def gate_stream(ops=None):
    """The 54-gate stream of one noiseless cycle as (name, angle, bits) -- SURVEY Appendix A."""
    out = []
    for op in ops or build_cycle():
        if op.has_ry1:
            out.append(("Ry", +math.pi / 2, (op.data_bit,)))
        out.append(("Rxx", op.s * math.pi / 2, (op.data_bit, op.anc_bit)))
        if op.has_rx:
            out.append(("Rx", -op.s * math.pi / 2, (op.data_bit,)))
        if op.has_ry2:
            out.append(("Ry", -math.pi / 2, (op.data_bit,)))
    return out


# ---------------------------------------------------------------------------------------------------
# error-slot indexing (CS:134-182)
# ---------------------------------------------------------------------------------------------------
def slot_index(prob_name, gate, err, rx_ind, ry_ind, rxx_ind, sc_ind=-1):
    n_rx, n_ry, n_rxx = NUM_GATES_MS["Rx"], NUM_GATES_MS["Ry"], NUM_GATES_MS["Rxx"]
    one_q = {"Rx": rx_ind, "Ry": n_rx + ry_ind}
    if prob_name in ("p_deph", "p_leak"):
        if gate in one_q:
            return one_q[gate]
        if gate == "Rxx" and err in ("Zc", "Lc"):
            return n_rx + n_ry + rxx_ind
        if gate == "Rxx" and err in ("Zt", "Lt"):
            return n_rx + n_ry + n_rxx + rxx_ind
    elif prob_name == "p_idle":
        return sc_ind
    elif prob_name == "p_X_ctrl":
        return rx_ind
    elif prob_name == "p_Y_ctrl":
        return ry_ind
    elif prob_name in ("p_XX_ctrl", "p_heat"):
        return rxx_ind
    # extensions: the reference leaves these undefined (index stays None -> TypeError); SURVEY App. C-1 / C-10
    elif prob_name in ("px", "py", "pz", "eps_1q"):
        if gate in one_q:
            return one_q[gate]
    elif prob_name in ("p", "eps_2q"):
        return rxx_ind
    raise ValueError("error model entry ({}, {}) on gate {} has no slot-index rule".format(err, prob_name, gate))


def list_lengths(model, two_rounds=False):
    """Per-round length every probability list of `model` needs."""
    need = {name: 0 for name in model["probabilities"]}
    for op in build_cycle():
        for gate, ri in (("Ry", op.ry1_ind), ("Rxx", op.index), ("Rx", op.rx_ind), ("Ry", op.ry2_ind)):
            if ri < 0:
                continue
            for err, pname, *_ in model["gates"].get(gate, []):
                idx = slot_index(pname, gate, err, op.rx_ind if gate == "Rx" else 0, ri if gate == "Ry" else 0, op.index)
                need[pname] = max(need[pname], idx + 1)
    for err, pname, *_ in model["gates"].get("Idle", []):
        need[pname] = max(need[pname], 7 * 17)
    return [need[n] * (2 if two_rounds else 1) for n in model["probabilities"]]


@dataclass
class Program:
    layers: List[list] = field(default_factory=list)    # [prep | cycle 1 | cycle 2][layer] -> L.Layer
    params: List[float] = field(default_factory=list)   # flat (cos, sin) pairs
    plan: List[L.PInstr] = field(default_factory=list)
    list_len: List[int] = field(default_factory=list)
    fusion: int = 1
    gates_per_cycle: int = 54


_PAULI_ERR = {"X": L.ERR_X, "Y": L.ERR_Y, "Z": L.ERR_Z, "Zc": L.ERR_ZC, "Zt": L.ERR_ZT, "XX": L.ERR_XX,
              "ZZ": L.ERR_ZZ, "depol": L.ERR_DEPOL}
_OVER = {"Rx_over": L.U_ROT_X, "Ry_over": L.U_ROT_Y, "Rxx_over": L.U_ROT_XX}


def _uop(type_, sign=0, skippable=0, ev_shift=255, param=0):
    u = L.Uop()
    u.type, u.sign, u.skippable, u.ev_shift, u.param = type_, sign, skippable, ev_shift, param
    return u


def compile_program(model, list_len, fusion=1, cooling=False, over_angles=None, early_measure=False, ops=None,
                    slots=None, idle_slots=None) -> Program:
    """Compile `model` (one entry of error_models.json) for lists of total length `list_len`.

    ops / slots / idle_slots: the cycle table, the slot of every error entry {(operation, gate position, entry number):
    slot} and of every idle entry {(timestep, qubit): slot} as TRACED from a module's own stabilizer_cycle_error_index
    (tracer.trace_cycle); default: this module's table (build_cycle) and rule (slot_index).

    fusion: 0 = one sweep per gate (ProjectQ's execution model), 1 = one sweep per timestep,
            2 = one sweep per two timesteps (the form the half-cycle and whole-cycle kernels are lowered from).
    over_angles: {prob_name: [angles]} for the coherent ``*_over`` entries (extension, SURVEY C-10).
    early_measure: measure the four X-type ancillas right after timestep 4 (no later gate touches them, so
            this commutes with the rest of the cycle: same joint outcome statistics and the same collapsed
            state up to a global sign) -- the Z-type half of the cycle then runs on a 2^13-amplitude support.
    """
    if fusion not in (0, 1, 2):
        raise ValueError("fusion must be 0, 1 or 2")
    names = list(model["probabilities"])
    if len(names) != len(list_len):
        raise ValueError("make sure you pass enough probabilities to specify the error model - see error_model.json")
    if cooling and "Idle" not in model["gates"]:
        raise KeyError("Idle")     # the reference raises KeyError at error_model['gates'][gate] (CS:132)
    ops = list(ops) if ops is not None else build_cycle()
    prog = Program(list_len=list(list_len), fusion=fusion)
    over_angles = over_angles or {}

    def gates_of(op):
        g = []
        if op.has_ry1:
            g.append(("Ry1", "Ry", L.U_RY, +1, op.ry1_ind))
        g.append(("Rxx", "Rxx", L.U_RXX, op.s, op.index))
        if op.has_rx:
            g.append(("Rx", "Rx", L.U_RX, -op.s, op.rx_ind))
        if op.has_ry2:
            g.append(("Ry2", "Ry", L.U_RY, -1, op.ry2_ind))
        return g

    # ---- injection program ------------------------------------------------------------------
    def emit(**kw):
        ins = L.PInstr()
        for k, v in kw.items():
            setattr(ins, k, v)
        prog.plan.append(ins)

    for t in range(8):
        for op in (o for o in ops if o.timestep == t):
            emit(kind=L.P_OP_BEGIN, word=op.index)
            for pos, gate, _, _, run_ind in gates_of(op):
                if gate == "Rxx":     # the leak predicate is read here, after any Ry error of this op (CS:322)
                    emit(kind=L.P_SKIP_TEST, a=op.d_leak, b=op.a_leak, word=op.index)
                rx_i = op.rx_ind if gate == "Rx" else 0
                ry_i = run_ind if gate == "Ry" else 0
                seen_pauli = False
                for k_entry, (err, pname, *_) in enumerate(model["gates"].get(gate, [])):
                    lid = names.index(pname)
                    if err in _OVER:
                        if seen_pauli:
                            raise NotImplementedError("coherent entries must precede stochastic ones in a gate's list")
                        continue                      # deterministic: lives in the gate stream, not the plan
                    seen_pauli = True
                    slot = (slots[(op.index, pos, k_entry)] if slots is not None else
                            slot_index(pname, gate, err, rx_i, ry_i, op.index))
                    kw = dict(kind=L.P_ERR, list=lid, slot=slot, field=FIELD_OF_GATE[pos], word=op.index,
                              in_skip=int(gate == "Rxx"), use_round=1)
                    c_ind, t_ind = (op.a_leak, op.d_leak) if op.kind == "x" else (op.d_leak, op.a_leak)  # CS:297, 326
                    if err in _PAULI_ERR:
                        if gate != "Rxx" and err in ("Zc", "Zt", "XX", "ZZ", "depol"):
                            raise ValueError("two-qubit error {} on single-qubit gate {}".format(err, gate))
                        kw["err"] = _PAULI_ERR[err]
                    elif err == "L":
                        if gate == "Rxx":
                            raise ValueError("'L' on a two-qubit gate has no d_ind in the reference")
                        kw.update(err=L.ERR_LEAK_A, a=op.d_leak)
                    elif err in ("Lc", "Lt", "LL"):
                        if gate != "Rxx":
                            raise ValueError("'{}' on a single-qubit gate indexes leaked_reg[-1] in the reference".format(err))
                        if err == "LL":
                            kw.update(err=L.ERR_LEAK_AB, a=c_ind, b=t_ind)
                        else:
                            kw.update(err=L.ERR_LEAK_A, a=c_ind if err == "Lc" else t_ind)
                    else:
                        raise ValueError("unknown error kind {!r}".format(err))
                    emit(**kw)
            # a: which single-qubit gates the op has, b: sign of the Rxx angle -- lets the planner conjugate the op's
            # Pauli events through the remaining (Clifford) gates to one net Pauli at the end of the op
            # a bits 3-6 / slot: the frame in which the op's half cycle is diagonal (s17_diag.cuh): bit3 Hadamard frame
            # (x-type half), bit4 first / bit6 last operation of its half, bit5 half index; slot = data bit
            half = 0 if op.kind == "x" else 1
            same = [o.index for o in ops if o.kind == op.kind]
            emit(kind=L.P_OP_END, word=op.index, slot=op.data_bit,
                 a=int(op.has_ry1) | (int(op.has_rx) << 1) | (int(op.has_ry2) << 2) | (int(op.kind == "x") << 3) |
                 (int(op.index == min(same)) << 4) | (half << 5) | (int(op.index == max(same)) << 6),
                 b=int(op.s > 0))
        if cooling and t < 7:                          # sympathetic_cooling (CS:762-772) after timesteps 1-7
            for q in range(N_DATA + N_ANC):
                for err, pname, *_ in model["gates"]["Idle"]:
                    if err != "Z":
                        raise NotImplementedError("only Z idle errors are supported")
                    # stab_ind is not forwarded by the reference (CS:767), hence use_round=0.
                    # Frame information for the whole-cycle kernel (s17_diag.cuh), b: bit7 present, bit1 the gap belongs
                    # to the x-type (Hadamard-frame) half, bit0 data qubit inside its Ry bracket; field / pad: for an
                    # ancilla, index within its half / event word of the last operation on it so far (255: the ancilla
                    # is still |0>, or already measured: Z acts trivially)
                    half_ops = [o for o in ops if (o.timestep < 4) == (t < 4)]
                    b = 0x80 | (0x02 if t < 4 else 0)
                    fld, pad = 255, 0
                    if q < N_DATA:
                        mine = [o.timestep for o in half_ops if o.data == q]
                        if t >= 4 and mine and min(mine) <= t < max(mine):
                            b |= 0x01
                    else:
                        done = [o for o in half_ops if o.ancilla == q - N_DATA and o.timestep <= t]
                        if done:
                            fld, pad = half_ops.index(done[-1]), done[-1].index
                    emit(kind=L.P_IDLE, err=L.ERR_Z, list=names.index(pname),
                         slot=idle_slots[(t, q)] if idle_slots is not None else t * 17 + q, a=q, word=24 + t,
                         use_round=0, b=b, field=fld, pad=pad)

    # ---- gate stream ------------------------------------------------------------------------
    def param_index(c, s):
        prog.params.extend([c, s])
        return len(prog.params) // 2 - 1

    def make_op(op, cycle, only=None):
        o = L.Op()
        o.kind, o.data_bit, o.anc_bit, o.ev_word = L.OP_ENT, op.data_bit, op.anc_bit, op.index
        o.partial = int(only is not None)
        uops = []
        for pos, gate, utype, sign, run_ind in gates_of(op):
            if only is not None and pos != only:
                continue
            skippable = int(gate == "Rxx")
            group = [_uop(utype, sign, skippable)]
            rx_i = op.rx_ind if gate == "Rx" else 0
            ry_i = run_ind if gate == "Ry" else 0
            for err, pname, *_ in model["gates"].get(gate, []):
                if err in _OVER and cycle >= 0:      # the preparation cycle is noiseless (CS:116-122)
                    angles = over_angles.get(pname)
                    if angles is None:
                        raise ValueError("no angles given for coherent entry {}".format(pname))
                    slot = slot_index(pname, gate, err, rx_i, ry_i, op.index)
                    if cycle == 1:
                        slot += len(angles) // 2      # second cycle reads the second half (CS:191-193)
                        if slot >= len(angles):       # one-round list: a second cycle is never run with it
                            slot -= len(angles) // 2
                    eps = float(angles[slot])
                    if eps != 0.0:
                        pi = param_index(math.cos(0.5 * eps), math.sin(0.5 * eps))
                        if pi > 255:
                            raise ValueError("too many distinct rotation parameters")
                        group.append(_uop(_OVER[err], 0, skippable, 255, pi))
            group[-1].ev_shift = FIELD_OF_GATE[pos]
            uops.extend(group)
        if len(uops) > 8:
            raise ValueError("operation needs more than 8 micro-ops")
        o.n_uops = len(uops)
        for i, u in enumerate(uops):
            o.uops[i] = u
        return o

    def idle_op(t):
        o = L.Op()
        o.kind, o.ev_word = L.OP_IDLE, 24 + t
        return o

    def make_layer(items, anc_bits, meas_mask=0):
        lay = L.Layer()
        anc = sorted(set(anc_bits))
        n_anc = 4 if len(anc) > 3 else 3
        fill = [b for b in range(N_DATA, N_DATA + N_ANC) if b not in anc]
        while len(anc) < n_anc:
            anc.append(fill.pop(0))
        anc.sort()
        if len(anc) != n_anc:
            raise ValueError("a sweep can stage at most 4 ancilla bits")
        if len(items) > (16 if n_anc == 4 else 8):
            raise ValueError("too many operations in one sweep")
        lay.n_ops = len(items)
        lay.n_anc = n_anc
        for i, b in enumerate(anc):
            lay.anc_bits[i] = b
        lay.meas_mask = int(meas_mask)
        for i, o in enumerate(items):
            lay.ops[i] = o
        return lay

    for cycle in (-1, 0, 1):                           # preparation, stab_ind=0, stab_ind=1
        layers = []
        x_half_end = -1
        if fusion == 0:
            for t in range(8):
                t_ops = [o for o in ops if o.timestep == t]
                for op in t_ops:
                    for pos, *_ in gates_of(op):
                        layers.append(([make_op(op, cycle, only=pos)], [op.anc_bit]))
                if cooling and t < 7:
                    layers[-1][0].append(idle_op(t))
                if t == 3:
                    x_half_end = len(layers) - 1
        else:
            step = fusion
            for t0 in range(0, 8, step):
                items, anc = [], []
                for t in range(t0, t0 + step):
                    for op in (o for o in ops if o.timestep == t):
                        items.append(make_op(op, cycle))
                        anc.append(op.anc_bit)
                    if cooling and t < 7:
                        items.append(idle_op(t))
                layers.append((items, anc))
                if t0 + step == 4:
                    x_half_end = len(layers) - 1
        x_mask = sum(1 << a for a in {o.ancilla for o in ops if o.kind == "x"})          # ancillas 1, 3, 4, 6
        built = []
        for i, (it, an) in enumerate(layers):
            mask = 0
            if i == len(layers) - 1:
                mask = (0xFF & ~x_mask) if early_measure else 0xFF
            elif early_measure and i == x_half_end:
                mask = x_mask
            built.append(make_layer(it, an, meas_mask=mask))
        prog.layers.append(built)
    return prog
