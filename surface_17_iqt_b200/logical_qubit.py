"""Object-oriented interface -- the names and the location bookkeeping of the reference's logical_qubit.py
(NoisyGate :14, EntanglingOperation :48, XTypeEntanglingOp :72, ZTypeEntanglingOp :85, StabiliserTimestep :115,
StabiliserCycle :159, LogicalQubit :262), executing on the batched B200 backend.

A location is ``[syndrome, timestep, entangling_op, gate]``.  The reference's OO circuit lists 60 gate locations
(24 Rxx + 12 Rx + 24 Ry, pinned by test_logical_qubit.py:26-39); its own arithmetic is unusable (all s = +1, wrong
sign on the second Ry, SURVEY C-14), so execution uses the functional module's 54-gate cycle: a location is mapped to
the error slot of the corresponding gate of that cycle; the six Ry locations that the functional circuit cancels by
hand have no counterpart and errors placed there are dropped (``StabiliserCycle.dropped_locations``).
"""
import random

import numpy as np

from . import _lib as L
from . import api, circuit
from .error import DephasingError, Error, XCtrlError, YCtrlError  # noqa: F401
from .pq_core import Qubit, Rx, Rxx, Ry          # the gate objects of the projectq-shaped face (`Gate | qubits`)

pdd_error_model = {Rxx: [XCtrlError, DephasingError], Rx: [XCtrlError, DephasingError], Ry: [YCtrlError, DephasingError]}

HALF_PI = np.pi / 2
_TABLE = circuit.build_cycle()                   # the 24 entangling operations: (timestep, data, ancilla, which gates)


def _on_engine(qubits):
    qs = qubits if isinstance(qubits, (list, tuple)) else (qubits,)
    return all(isinstance(q, Qubit) for q in qs)


class NoisyGate:
    """A gate, where it sits (``location``) and the error types the model allows behind it (logical_qubit.py:14-45).

    ``apply`` sends the gate to the engine of its qubits and then every error of ``error_list`` placed at this location
    (`Gate | qubits`, like the reference).  With plain qubit indices instead of engine qubits -- the batched path, where
    the whole cycle is one device program -- nothing is sent; either way the errors that sit here are returned."""

    def __init__(self, gate, qubits, location=None, error_model={}):
        self.gate, self.qubits, self.location, self.error_model = gate, qubits, location, error_model
        self.possible_errors = self.get_errors_from_error_model()

    def get_errors_from_error_model(self):
        return [kind() for kind in self.error_model.get(type(self.gate), ())]

    def __str__(self):
        fields = (("Gate", self.gate), ("Qubits", self.qubits), ("Location", self.location),
                  ("Error model", self.error_model), ("Errors", self.possible_errors))
        return "NoisyGate: " + ", ".join("{}: {}".format(k, v) for k, v in fields)

    def apply(self, error_list=[]):
        here = [e for e in error_list if e.location == self.location]
        if _on_engine(self.qubits):
            self.gate | self.qubits
            for e in here:
                e.insert_error(*(self.qubits if isinstance(self.qubits, (list, tuple)) else (self.qubits,)))
        return here


class EntanglingOperation:
    """One data qubit entangled with one ancilla (logical_qubit.py:48-69); subclasses list the gates."""

    def __init__(self, location=None, dataq=None, ancillaq=None, cancel_data_rx=None, s=None, v=None, error_model=None):
        self.location, self.error_model = location, error_model
        self.dataq, self.ancillaq = dataq, ancillaq
        self.cancel_data_rx, self.s, self.v = cancel_data_rx, s, v
        self.noisy_gates = self.build_noisy_gates()

    def gate_sequence(self):
        raise Exception("No build_noisy_gates function defined.")

    def build_noisy_gates(self):
        # a gate's location = the operation's location + its position in the operation
        return [NoisyGate(gate, qubits, self.location + [k], self.error_model)
                for k, (gate, qubits) in enumerate(self.gate_sequence())]

    def run(self, error_list=[]):
        return [e for g in self.noisy_gates for e in g.apply(error_list)]


class XTypeEntanglingOp(EntanglingOperation):
    def gate_sequence(self):
        seq = [(Rxx(self.s * HALF_PI), (self.dataq, self.ancillaq))]
        if not self.cancel_data_rx:
            seq.append((Rx(-self.s * HALF_PI), self.dataq))
        return seq


class ZTypeEntanglingOp(EntanglingOperation):
    def gate_sequence(self):
        # the reference's OO circuit keeps both Ry of every operation and gives them the same sign (SURVEY C-14)
        seq = [(Ry(self.v * HALF_PI), self.dataq), (Rxx(self.s * HALF_PI), (self.dataq, self.ancillaq))]
        if not self.cancel_data_rx:
            seq.append((Rx(-self.s * HALF_PI), self.dataq))
        seq.append((Ry(self.v * HALF_PI), self.dataq))
        return seq


class StabiliserTimestep:
    """Three entangling operations done in parallel (logical_qubit.py:115-156)."""

    def __init__(self, data=None, ancilla=None, location=None, qu_ind=None, entangling_type="X", cancel_data_rx=None,
                 error_model=None):
        self.data, self.ancilla, self.location, self.error_model = data, ancilla, location, error_model
        self.qu_ind, self.ent_type, self.cancel_data_rx = qu_ind, entangling_type, cancel_data_rx
        self.entangling_operations = self.build_entangling_operations()

    def build_entangling_operations(self):
        if self.ent_type not in ("X", "Z"):
            raise Exception("Entangling type not recognized, should be \"X\" or \"Z\". "
                            "Please check stabilizer_timestep instantiation.")
        klass = XTypeEntanglingOp if self.ent_type == "X" else ZTypeEntanglingOp
        return [klass(location=self.location + [k], dataq=self.data[d], ancillaq=self.ancilla[a],
                      cancel_data_rx=self.cancel_data_rx[k], s=1, v=1, error_model=self.error_model)
                for k, (d, a) in enumerate(self.qu_ind)]

    def run(self, error_list=[]):
        return [e for op in self.entangling_operations for e in op.run(error_list)]


# (data, ancilla) pairs and Rx cancellations per timestep, read off the compiled cycle table
QU_IND = tuple([[op.data, op.ancilla] for op in _TABLE if op.timestep == t] for t in range(8))
CANCEL_DATA_RX = tuple([not op.has_rx for op in _TABLE if op.timestep == t] for t in range(8))


class StabiliserCycle:
    """All data qubits get entangled with the appropriate ancillas (logical_qubit.py:159-259)."""

    def __init__(self, location=None, data=None, ancilla=None, circuit_type="MS", error_model=None, error_subset={}):
        self.data = data if data is not None else list(range(9))
        self.ancilla = ancilla if ancilla is not None else list(range(8))
        self.location = location if location is not None else [0]
        self.circuit_type = circuit_type
        self.error_model = error_model if error_model is not None else {}
        self.stabiliser_timesteps = self.build_stabiliser_timesteps(self.data, self.ancilla, circuit_type)
        self.all_gate_locations, self.error_locations, self.gate_locations = self.generate_gate_and_error_locations()
        self.errors_to_insert = Error.generate_error_list(error_subset, self.error_locations)
        self.dropped_locations = []

    def noisy_gates(self):
        for timestep in self.stabiliser_timesteps:
            for op in timestep.entangling_operations:
                yield from op.noisy_gates

    def generate_gate_and_error_locations(self):
        """(every gate location, {error class: locations it can occur at}, {gate class: locations})."""
        by_gate = {Rx: [], Ry: [], Rxx: []}
        by_error = {XCtrlError: [], YCtrlError: [], DephasingError: []}
        everything = []
        for g in self.noisy_gates():
            everything.append(g.location)
            by_gate[type(g.gate)].append(g.location)
            for e in g.possible_errors:
                by_error[type(e)].append(g.location)
        return everything, by_error, by_gate

    def build_stabiliser_timesteps(self, data, ancilla, circuit_type):
        if len(data) != 9:
            raise Exception('data qubit register does not correspond to the surface 17 QEC code')
        if circuit_type != "MS":
            print("Nothing done.")
            return []
        return [StabiliserTimestep(data=data, ancilla=ancilla, location=self.location + [t], qu_ind=QU_IND[t],
                                   entangling_type="X" if t < 4 else "Z", cancel_data_rx=CANCEL_DATA_RX[t],
                                   error_model=self.error_model) for t in range(8)]

    # ---- mapping onto the functional 54-gate cycle ---------------------------------------------------
    def slot_masks(self, error_list=None):
        """PDD_model p_set ([p_X_ctrl 12, p_Y_ctrl 18, p_XX_ctrl 24, p_deph 78, p_heat 24]) for these errors.
        Two equal errors at one location cancel (they are Paulis), so a slot holds the parity of its hits."""
        errors = self.errors_to_insert if error_list is None else error_list
        masks = [12 * [0], 18 * [0], 24 * [0], 78 * [0], 24 * [0]]
        ops = circuit.build_cycle()
        self.dropped_locations = []
        for e in errors:
            _synd, t, k, g = e.location[-4:]
            op = ops[3 * t + k]
            if t < 4:
                gate = ("Rxx", "Rx")[g]
                run = {"Rxx": op.index, "Rx": op.rx_ind}[gate]
            else:
                names = ["Ry1", "Rxx", "Rx", "Ry2"] if not CANCEL_DATA_RX[t][k] else ["Ry1", "Rxx", "Ry2"]
                gate = names[g]
                run = {"Ry1": op.ry1_ind, "Rxx": op.index, "Rx": op.rx_ind, "Ry2": op.ry2_ind}[gate]
            if run < 0:                                   # cancelled by hand in the functional circuit
                self.dropped_locations.append(e.location)
                continue
            if e.pauli == "X":
                if gate == "Rxx":
                    masks[2][run] ^= 1
                elif gate == "Rx":
                    masks[0][run] ^= 1
                else:
                    raise ValueError("XCtrlError is not defined on {}".format(gate))
            elif e.pauli == "Y":
                if not gate.startswith("Ry"):
                    raise ValueError("YCtrlError is only defined on Ry")
                masks[1][run] ^= 1
            elif e.pauli == "Z":
                if gate == "Rx":
                    masks[3][run] ^= 1
                elif gate.startswith("Ry"):
                    masks[3][12 + run] ^= 1
                else:                                     # Z on both qubits of the Rxx
                    masks[3][30 + run] ^= 1
                    masks[3][54 + run] ^= 1
        return masks

    def run(self, eng=None, reset=True):
        """Gate-by-gate execution has no batched equivalent; use LogicalQubit / calc_error_subset_error_rate."""
        return [e for ts in self.stabiliser_timesteps for e in ts.run(self.errors_to_insert)]


class LogicalQubit:
    """A batch of logical qubits on the device (logical_qubit.py:262-318): prepare, one noisy syndrome cycle with the
    given errors, lookup decode, correct, read out."""

    def __init__(self, correction_table=None, bitflip_table=None, max_shots=64, device=0):
        self.state, self.basis = 0, 'Z'
        self.correction_table = correction_table or api.load_lookup_table("correction_table_depolarising.json")
        self.bitflip_table = bitflip_table or api.load_lookup_table("correction_table_classical_bitflip.json")
        self.sim = api.Simulation("PDD_model", [12, 18, 24, 78, 24], max_shots, self.correction_table, self.bitflip_table,
                                  fusion=2, device=device)
        self.leaked_q_reg = 17 * [0]
        self.records = []

    def measure_syndromes(self, mask_sets, seed=None):
        """One shot per mask set: single-round protocol (prep, noisy cycle, decode flips_a, correct, read out)."""
        n = len(mask_sets)
        flat = [[b for lst in m for b in lst] for m in mask_sets]
        self.sim.backend.set_replay(flat, [[0] for _ in range(n)])
        c = self.sim.backend.run("single", n, L.NOISE_REPLAY, seed=random.getrandbits(64) if seed is None else seed)
        self.records = self.sim.backend.records(n)
        return c

    def lookup(self, syndrome, display=False):
        key = " ".join(str(int(b)) for b in np.asarray(syndrome).ravel())
        error_vec = self.correction_table[key][0]
        if display:
            print('ft syndrome: {}'.format(syndrome))
            print('error vector: {}'.format(error_vec))
        return error_vec

    def close(self):
        self.sim.close()


def load_lookup_table(filename):
    return api.load_lookup_table(filename)
