"""The reference module of the same name (compiled_surface_code_arbitrary_error_model.py [CS]) on the B200 backend.

Two ways to run the compiled Surface-17 cycle:

* the batched fast path -- ``circuit.compile_program`` turns the cycle + error model into a device program once and
  ``calc_log_e_rate_arbitrary_error_model`` runs whole shot protocols per call;
* the gate-by-gate path of this file -- the reference's circuit-level entry points (`logical_prep` CS:96,
  `stabilizer_cycle_error_index` CS:773, `insert_errors` CS:127, `insert_error` CS:188, `sympathetic_cooling` CS:762,
  `lookup` CS:886, `apply_correction` CS:901, `logical_measurement` CS:52) with the same names, arguments and return
  values, emitting `Gate | qubit` commands to a ProjectQ-shaped engine (``pq_compat/projectq``: the GPU `Simulator`, or
  a recording backend).  The reference writes the cycle as eight hand-unrolled timestep functions (CS:365-650); here
  the same 24 entangling operations are one table (``circuit.build_cycle``) walked by one function, and the slot
  indices come from ``circuit.slot_index``.  ``tracer.py`` runs the reference's own functions and these on a recording
  engine and checks that they emit the same stream.

Error draws use Python's global ``random`` exactly like the reference (one ``random.random()`` per entry, CS:198), so
``random.seed`` reproduces a trajectory.
"""
import math
import random

import numpy as np

from . import circuit
from .api import load_error_models, load_lookup_table  # noqa: F401  (CS:910-913)
from .pq_core import All, H, Measure, Rx, Rxx, Ry, X, Y, Z

HALF_PI = math.pi / 2
_CYCLE = circuit.build_cycle()
_ROTATION = {"Rx_over": Rx, "Ry_over": Ry, "Rxx_over": Rxx}       # coherent over-rotation entries (extension, SURVEY C-10)
# insert_2q_error (CS:237-283): the 15 non-identity two-qubit Paulis in the reference's order, first letter -> qubit1
_TWO_QUBIT_PAULIS = [a + b for a in "IXYZ" for b in "IXYZ"][1:]
_PAULI_GATE = {"X": X, "Y": Y, "Z": Z}


def instantiate_error_model(p_set, model):
    """CS:31-49: bind the i-th probability list to the i-th name of ``model``.

    Returns ``(error_model, {name: list})``.  Deviation (SURVEY C-11): a length mismatch raises
    ValueError instead of printing and returning a half-filled dict.
    """
    error_model = load_error_models()[model]
    names = error_model["probabilities"]
    if len(names) != len(p_set):
        raise ValueError("make sure you pass enough probabilities to specify the error model - see error_model.json")
    return error_model, {key: p_set[i] for i, key in enumerate(names)}


# ---- error insertion (CS:127-288) -----------------------------------------------------------------------------------
def insert_leakage(index, leaked_reg):
    leaked_reg[index] = 1


def insert_2q_error(qubit1, qubit2):
    """CS:237-283: one uniform draw picks one of the 15 non-identity two-qubit Paulis."""
    r = random.random()
    k = min(14, int(r * 15))
    first, second = _TWO_QUBIT_PAULIS[k]
    if first != "I":
        _PAULI_GATE[first] | qubit1
    if second != "I":
        _PAULI_GATE[second] | qubit2


def insert_error(error, prob, leaked_reg, qubits, c_ind, t_ind, d_ind, index=0, stab_ind=0):
    """CS:188-234: one Bernoulli draw against ``prob[index]`` (second cycle: the second half of the list), then the
    Pauli(s) / leak flag(s) the entry names."""
    if stab_ind == 1:
        index += len(prob) // 2
    if error in _ROTATION:                         # deterministic: the list holds the over-rotation angle
        if prob[index] != 0.0:
            _ROTATION[error](prob[index]) | qubits
        return
    if not random.random() < prob[index]:
        return
    pair = qubits if isinstance(qubits, (list, tuple)) else (qubits, None)
    if error in _PAULI_GATE:
        _PAULI_GATE[error] | qubits
    elif error == "Zc":
        Z | pair[0]
    elif error == "Zt":
        Z | pair[1]
    elif error in ("XX", "ZZ"):
        _PAULI_GATE[error[0]] | pair[0]
        _PAULI_GATE[error[0]] | pair[1]
    elif error == "L":
        insert_leakage(d_ind, leaked_reg)
    elif error == "Lc":
        insert_leakage(c_ind, leaked_reg)
    elif error == "Lt":
        insert_leakage(t_ind, leaked_reg)
    elif error == "LL":
        insert_leakage(c_ind, leaked_reg)
        insert_leakage(t_ind, leaked_reg)
    elif error == "depol":
        insert_2q_error(pair[0], pair[1])


def insert_errors(gate, qubits, leaked_reg, error_model, e_model_probs, c_ind=-1, t_ind=-1, d_ind=-1, display=False,
                  rx_ind=0, ry_ind=0, rxx_ind=0, rzz_ind=0, stab_ind=0, sc_ind=-1):
    """CS:127-185: every entry of the model for ``gate``, in file order, at its slot of the named list."""
    for entry in error_model["gates"][gate]:
        err, pname = entry[0], entry[1]
        index = circuit.slot_index(pname, gate, err, rx_ind, ry_ind, rxx_ind, sc_ind)
        if display:
            print("{} error {} at slot {} of {}".format(gate, err, index, pname))
        insert_error(err, e_model_probs[pname], leaked_reg, qubits, c_ind, t_ind, d_ind, index=index, stab_ind=stab_ind)


# ---- the cycle (CS:291-344, 365-650, 773-863) -----------------------------------------------------------------------
def entangling_operation(op, data, ancilla, leaked_reg, error_model, e_model_probs, stab_ind=0):
    """One row of the cycle table: x_type_entangling (CS:291-309) / z_type_entangling (CS:312-344) with the hand
    cancellations of the timestep it sits in.  A leaked data or ancilla qubit skips the Rxx and its errors; the slot
    indices advance regardless."""
    dq, aq = data[op.data], ancilla[op.ancilla]
    ind = dict(rx_ind=max(op.rx_ind, 0), rxx_ind=op.index, stab_ind=stab_ind)
    if op.has_ry1:
        Ry(HALF_PI) | dq
        insert_errors("Ry", dq, leaked_reg, error_model, e_model_probs, d_ind=op.d_leak, ry_ind=op.ry1_ind, **ind)
    if leaked_reg[op.d_leak] == 0 and leaked_reg[op.a_leak] == 0:
        Rxx(op.s * HALF_PI) | (dq, aq)
        # timesteps 1-4 call the ancilla the control, 5-8 the data qubit (CS:297, 326)
        c_ind, t_ind = (op.a_leak, op.d_leak) if op.kind == "x" else (op.d_leak, op.a_leak)
        insert_errors("Rxx", [dq, aq], leaked_reg, error_model, e_model_probs, c_ind=c_ind, t_ind=t_ind, **ind)
    if op.has_rx:
        Rx(-op.s * HALF_PI) | dq
        insert_errors("Rx", dq, leaked_reg, error_model, e_model_probs, d_ind=op.d_leak, ry_ind=0, **ind)
    if op.has_ry2:
        Ry(-HALF_PI) | dq
        insert_errors("Ry", dq, leaked_reg, error_model, e_model_probs, d_ind=op.d_leak, ry_ind=op.ry2_ind, **ind)


def stabiliser_timestep(t, data, ancilla, leaked_reg, error_model, e_model_probs, stab_ind=0):
    """Timestep t = 1..8 of the cycle (the reference's stabiliser_timestep_1 .. _8, CS:365-650)."""
    for op in _CYCLE:
        if op.timestep == t - 1:
            entangling_operation(op, data, ancilla, leaked_reg, error_model, e_model_probs, stab_ind)


def sympathetic_cooling(data, ancilla, error_model, e_model_probs, leaked_reg, sc_ind):
    """CS:762-772: one Idle slot per qubit (data 0-8, ancilla 0-7) in the gap after timestep sc_ind + 1.  Like the
    reference it does not forward stab_ind: a second cycle indexes the first half of p_idle again."""
    qubits = list(data) + list(ancilla)
    for k, q in enumerate(qubits):
        insert_errors("Idle", q, leaked_reg, error_model, e_model_probs, sc_ind=sc_ind * len(qubits) + k)


def stabilizer_cycle_error_index(data, ancilla, leaked_reg, eng, error_model, e_model_probs, reset=True, stab_ind=0,
                                 cooling=False):
    """CS:773-863: the 54-gate cycle with running error-slot indices, All(Measure) | ancilla, the leak override
    (a leaked ancilla reads 1 and is un-leaked), and the reset of the ancillas that read 1.  Returns the syndrome."""
    if len(data) != 9:
        raise Exception('data qubit register does not correspond to the surface 17 QEC code')
    for t in range(1, 9):
        stabiliser_timestep(t, data, ancilla, leaked_reg, error_model, e_model_probs, stab_ind)
        if cooling and t < 8:
            sympathetic_cooling(data, ancilla, error_model, e_model_probs, leaked_reg, sc_ind=t - 1)
    All(Measure) | ancilla
    eng.flush()
    outcomes = [int(q) for q in ancilla]
    syndrome_t = list(outcomes)
    for j in range(len(ancilla)):
        if leaked_reg[9 + j] == 1:
            syndrome_t[j] = 1
            leaked_reg[9 + j] = 0
    if reset:
        for a, bit in zip(ancilla, outcomes):
            if bit == 1:
                X | a
    return syndrome_t


def logical_prep(data, basis, state, ancilla, leaked_q_reg, eng, e_model, e_probs, cz_compilation=False):
    """CS:96-123: optional All(X) / All(H), then one NOISELESS cycle (all-zero lists of the same lengths; the entries
    still draw) whose syndrome is the quiescent state."""
    if cz_compilation:
        raise NotImplementedError("the experimental CZ compilation is outside the accelerated path (SURVEY C-9)")
    if state == 1:
        All(X) | data
    if basis == 'X':
        All(H) | data
    errorless = {key: len(plist) * [0] for key, plist in e_probs.items()}
    return np.array(stabilizer_cycle_error_index(data, ancilla, leaked_q_reg, eng, e_model, errorless, reset=True))


# ---- decode and read-out (CS:52-93, 886-907) ------------------------------------------------------------------------
def _key(bits):
    return " ".join(str(int(b)) for b in np.asarray(bits).ravel())


def lookup(syndrome, table, display=False):
    """CS:886-892: 8-bit syndrome (np.ndarray of ints) -> 18-element correction vector."""
    error_vec = table[_key(syndrome)][0]
    if display:
        print('ft syndrome: {}'.format(syndrome))
        print('error vector: {}'.format(error_vec))
    return error_vec


def return_weight_classical_lookup(syndrome, table):
    """CS:895-898."""
    return table[_key(syndrome)][1]


def apply_correction(error_vec, data):
    """CS:901-907: X on data i if vec[i], Z on data i if vec[i + 9]."""
    for i in range(9):
        if error_vec[i] == 1:
            X | data[i]
        if error_vec[i + 9] == 1:
            Z | data[i]


# the data qubits each Z-type stabiliser reads (ancillas 0, 2, 5, 7), for the classical correction of the read-out
_Z_CHECKS = ((0, 1), (1, 2, 4, 5), (3, 4, 6, 7), (7, 8))
_Z_ANCILLAS = (0, 2, 5, 7)


def logical_measurement(data, basis, eng, classical_lookup, quiescent, leaked_q_reg):
    """CS:52-93: measure the nine data qubits (a leaked one reads 1), infer the four Z-stabiliser values from the
    read-out, correct the parity by the weight of the classical bit-flip table entry, return the logical value."""
    if basis == 'X':
        All(H) | data
    All(Measure) | data
    eng.flush()
    data_meas = [1 if leaked_q_reg[i] == 1 else int(q) for i, q in enumerate(data)]
    weight = 0
    if basis == 'Z':
        classical = [(sum(data_meas[i] for i in check) - quiescent[a]) % 2 for check, a in zip(_Z_CHECKS, _Z_ANCILLAS)]
        weight = return_weight_classical_lookup(classical, classical_lookup)
    return (sum(data_meas) + weight) % 2


def get_results_log_e(rounds, runs, incorrect_count, time):
    return {"runs": runs, "rounds": rounds, "incorrect_count": incorrect_count, "time_taken": time}


def get_results_log_e_run_til_fail(rounds, time, p1q, p2q):
    return {"rounds_til_fail": rounds, "p1q_error": p1q, "p2q_error": p2q, "time_taken": time}
