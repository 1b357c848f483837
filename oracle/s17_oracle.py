"""CPU restatement of the reference's Surface-17 hot path.  TEST INFRASTRUCTURE ONLY.

See oracle/__init__.py for who may import this.  It restates, table-driven and without
ProjectQ, what the reference computes per Monte-Carlo shot:

* the compiled 54-gate Molmer-Sorensen syndrome-extraction cycle with running error-slot
  indices (compiled_surface_code_arbitrary_error_model.py [CS]:291-344, 365-650, 773-863),
* per-gate error insertion from an ``error_models.json`` model (CS:127-288),
* ancilla measure / leak override / reset (CS:848-862), lookup decode and correction
  (CS:886-907), logical read-out with the classical bit-flip table (CS:52-93),
* the single-round, s1, s2 and run-til-fail shot protocols
  (calc_log_e_rate_arbitrary_error_model.py [CL]:183-259, 261-353, 355-457, 33-148) and the
  fixed-weight placement routines (CL:151-181).

State arithmetic goes through oracle/sv.py (one CPU sweep per gate, ProjectQ conventions,
SURVEY.md Appendix D).  The error draws consume ``rnd.random()`` in exactly the reference's
order (one per error entry, plus one for ``depol``; CS:198, 238), so with the same
``random`` seed this file and the unmodified reference running on oracle/pqshim follow the
same trajectory -- that is how oracle/make_golden.py pins it.

PARITY UNPINNED against real ProjectQ (not installable here); pinned against the reference's
own Python through the shim and against the reference's shipped decoder tables.

Two documented extensions the reference cannot express:
``depolarising_model`` slot indexing, and the ``*_over`` coherent over-rotation entries.
"""
import math
import random as _random

import numpy as np

from . import sv as _sv

N_DATA, N_ANC, N_QUBITS = 9, 8, 17
N_RX, N_RY, N_RXX = 12, 18, 24            # CS:129  num_gates_ms_comp
N_IDLE = 7 * 17                           # CS:762-772, 790-836

# ---------------------------------------------------------------------------------------------
# The eight stabiliser timesteps (CS:365-381, 495-650).  One row per entangling operation:
# (data, ancilla, s, cancel_data_rx, cancel_ry1, cancel_ry2, d_leak_ind, a_leak_ind)
# Timesteps 1-4 are x_type_entangling (no Ry), 5-8 are z_type_entangling (v = +1 everywhere).
# Row 2 of timestep 5 carries the reference's leak-index quirk (CS:559-561): its leak indices
# are data 1 / ancilla 2 although it acts on data 3 / ancilla 5.
# ---------------------------------------------------------------------------------------------
_T, _F = True, False
TIMESTEPS = (
    ("x", ((4, 1, +1, _T, _F, _F, 4, 10), (8, 6, +1, _F, _F, _F, 8, 15), (6, 4, +1, _F, _F, _F, 6, 13))),
    ("x", ((1, 1, -1, _F, _F, _F, 1, 10), (5, 6, -1, _T, _F, _F, 5, 15), (3, 4, -1, _T, _F, _F, 3, 13))),
    ("x", ((3, 1, +1, _T, _F, _F, 3, 10), (7, 6, +1, _F, _F, _F, 7, 15), (5, 3, +1, _T, _F, _F, 5, 12))),
    ("x", ((0, 1, -1, _F, _F, _F, 0, 10), (4, 6, -1, _T, _F, _F, 4, 15), (2, 3, -1, _F, _F, _F, 2, 12))),
    ("z", ((1, 2, +1, _T, _F, _T, 1, 11), (3, 5, +1, _F, _F, _F, 1, 11), (7, 7, +1, _T, _F, _T, 7, 16))),
    ("z", ((2, 2, -1, _F, _F, _F, 2, 11), (4, 5, -1, _T, _F, _T, 4, 14), (8, 7, -1, _F, _F, _F, 8, 16))),
    ("z", ((4, 2, +1, _T, _T, _F, 4, 11), (6, 5, +1, _F, _F, _F, 6, 14), (0, 0, +1, _F, _F, _F, 0, 9))),
    ("z", ((5, 2, -1, _F, _F, _F, 5, 11), (7, 5, -1, _T, _T, _F, 7, 14), (1, 0, -1, _T, _T, _F, 1, 9))),
)

HALF_PI = math.pi / 2.0


def slot_index(prob_name, gate, err, rx_ind, ry_ind, rxx_ind, sc_ind):
    """Slot index into the named probability list for one error entry (CS:134-182).

    ``depolarising_model`` (px/py/pz/p) has no branch in the reference (index stays None and
    the shot raises); the indexing used here is the SURVEY.md Appendix C-1 definition.
    """
    if prob_name in ("p_deph", "p_leak"):
        if gate == "Rx":
            return rx_ind
        if gate == "Ry":
            return N_RX + ry_ind
        if gate == "Rxx":
            if err in ("Zc", "Lc"):
                return N_RX + N_RY + rxx_ind
            if err in ("Zt", "Lt"):
                return N_RX + N_RY + N_RXX + rxx_ind
        return None
    if prob_name == "p_idle":
        return sc_ind
    if prob_name == "p_X_ctrl":
        return rx_ind
    if prob_name in ("p_XX_ctrl", "p_heat"):
        return rxx_ind
    if prob_name == "p_Y_ctrl":
        return ry_ind
    # ---- extensions (not expressible in the reference) ----
    if prob_name in ("px", "py", "pz", "eps_1q"):
        if gate == "Rx":
            return rx_ind
        if gate == "Ry":
            return N_RX + ry_ind
        return None
    if prob_name in ("p", "eps_2q"):
        return rxx_ind
    return None


# order of the 15 two-qubit Paulis chosen by insert_2q_error (CS:237-283); first letter -> qubit1
DEPOL_ORDER = ("IX", "IY", "IZ", "XI", "XX", "XY", "XZ", "YI", "YX", "YY", "YZ", "ZI", "ZX", "ZY", "ZZ")


class ForcedMeasure:
    """Measurement source that replays a fixed outcome sequence."""

    def __init__(self, bits):
        self.bits = [int(b) for b in bits]
        self.k = 0

    def __call__(self, state, bit):
        out = self.bits[self.k]
        self.k += 1
        return out


class SampledMeasure:
    """ProjectQ-style outcome choice: one uniform, cumulative walk, read the bit."""

    def __init__(self, rng):
        self.rng = rng

    def __call__(self, state, bit):
        return (state.sample_basis(self.rng.random()) >> bit) & 1


class ShotOracle:
    def __init__(self, error_models, correction_table, bitflip_table):
        self.error_models = error_models
        self.correction_table = correction_table
        self.bitflip_table = bitflip_table
        self.depol_script = None  # optional {(stab_ind, rxx_ind): k} replaying insert_2q_error's choice
        self.log = None  # optional list receiving ('gate', name, angle, bits) / ('meas', bit, out, p)

    # ---- primitive ops ----------------------------------------------------------------
    def _gate(self, st, name, bits, angle=None):
        if self.log is not None:
            self.log.append(("gate", name, None if angle is None else _sv.reduce_angle(angle), tuple(bits)))
        if name == "X":
            st.apply_1q(_sv.MAT_X, bits[0])
        elif name == "Y":
            st.apply_1q(_sv.MAT_Y, bits[0])
        elif name == "Z":
            st.apply_1q(_sv.MAT_Z, bits[0])
        elif name == "H":
            st.apply_1q(_sv.MAT_H, bits[0])
        elif name == "Rx":
            st.apply_1q(_sv.mat_rx(_sv.reduce_angle(angle)), bits[0])
        elif name == "Ry":
            st.apply_1q(_sv.mat_ry(_sv.reduce_angle(angle)), bits[0])
        elif name == "Rxx":
            st.apply_2q(_sv.mat_rxx(_sv.reduce_angle(angle)), bits[0], bits[1])
        else:
            raise ValueError(name)

    def _measure(self, st, bit, meas):
        out = meas(st, bit)
        p = st.collapse(bit, out)
        if p <= 0.0:
            raise RuntimeError("outcome {} on bit {} has zero probability".format(out, bit))
        if self.log is not None:
            self.log.append(("meas", bit, out, p))
        return out

    # ---- error insertion (CS:127-288) -----------------------------------------------
    def _insert_errors(self, st, gate, bits, leaked, gates, probs, rnd, c_ind=-1, t_ind=-1, d_ind=-1,
                       rx_ind=0, ry_ind=0, rxx_ind=0, stab_ind=0, sc_ind=-1):
        for entry in gates[gate]:
            err, pname = entry[0], entry[1]
            plist = probs[pname]
            index = slot_index(pname, gate, err, rx_ind, ry_ind, rxx_ind, sc_ind)
            if stab_ind == 1:
                index += int(len(plist) / 2)             # CS:191-193
            if err.endswith("_over"):                   # extension C-10: deterministic coherent error
                eps = plist[index]
                if eps != 0.0:
                    self._gate(st, {"Rx_over": "Rx", "Ry_over": "Ry", "Rxx_over": "Rxx"}[err], bits, eps)
                continue
            if not (rnd.random() < plist[index]):       # CS:198
                continue
            if err in ("X", "Y", "Z"):
                self._gate(st, err, bits[:1])
            elif err == "Zc":
                self._gate(st, "Z", bits[0:1])
            elif err == "Zt":
                self._gate(st, "Z", bits[1:2])
            elif err == "XX":
                self._gate(st, "X", bits[0:1])
                self._gate(st, "X", bits[1:2])
            elif err == "ZZ":
                self._gate(st, "Z", bits[0:1])
                self._gate(st, "Z", bits[1:2])
            elif err == "L":
                leaked[d_ind] = 1
            elif err == "LL":
                leaked[c_ind] = 1
                leaked[t_ind] = 1
            elif err == "Lc":
                leaked[c_ind] = 1
            elif err == "Lt":
                leaked[t_ind] = 1
            elif err == "depol":
                self._insert_2q_error(st, bits, rnd, key=(stab_ind, rxx_ind))

    def _insert_2q_error(self, st, bits, rnd, key=None):
        r = rnd.random()                                # CS:238
        if self.depol_script is not None:               # test hook: replay a recorded 15-way choice
            r = (self.depol_script[key] + 0.5) / 15.0
        for k, pp in enumerate(DEPOL_ORDER):
            lo, hi = k * 1 / 15, (k + 1) * 1 / 15
            if (k == 0 and r < hi) or (k > 0 and lo < r < hi):   # strict compares as in CS:239-283
                if pp[0] != "I":
                    self._gate(st, pp[0], bits[0:1])
                if pp[1] != "I":
                    self._gate(st, pp[1], bits[1:2])

    # ---- one stabiliser cycle (CS:773-863) ----------------------------------------------
    def stabilizer_cycle(self, st, leaked, gates, probs, rnd, meas, stab_ind=0, cooling=False, reset=True,
                         gates_only=False):
        rx_ind = ry_ind = rxx_ind = 0
        for step, (kind, ops) in enumerate(TIMESTEPS):
            for (d, a, s, c_rx, c_ry1, c_ry2, dl, al) in ops:
                db, ab = d, N_DATA + a
                kw = dict(stab_ind=stab_ind)
                if kind == "z" and not c_ry1:           # CS:314-320
                    self._gate(st, "Ry", (db,), +HALF_PI)
                    self._insert_errors(st, "Ry", (db,), leaked, gates, probs, rnd, d_ind=dl,
                                        rx_ind=rx_ind, ry_ind=ry_ind, rxx_ind=rxx_ind, **kw)
                    ry_ind += 1
                if leaked[dl] == 0 and leaked[al] == 0:  # CS:293, 322
                    self._gate(st, "Rxx", (db, ab), s * HALF_PI)
                    c_i, t_i = (al, dl) if kind == "x" else (dl, al)   # CS:297 vs CS:326
                    self._insert_errors(st, "Rxx", (db, ab), leaked, gates, probs, rnd, c_ind=c_i, t_ind=t_i,
                                        rx_ind=rx_ind, ry_ind=ry_ind, rxx_ind=rxx_ind, **kw)
                rxx_ind += 1
                if not c_rx:                            # CS:303-308, 332-337
                    self._gate(st, "Rx", (db,), -s * HALF_PI)
                    self._insert_errors(st, "Rx", (db,), leaked, gates, probs, rnd, d_ind=dl,
                                        rx_ind=rx_ind, ry_ind=ry_ind, rxx_ind=rxx_ind, **kw)
                    rx_ind += 1
                if kind == "z" and not c_ry2:           # CS:338-343
                    self._gate(st, "Ry", (db,), -HALF_PI)
                    self._insert_errors(st, "Ry", (db,), leaked, gates, probs, rnd, d_ind=dl,
                                        rx_ind=rx_ind, ry_ind=ry_ind, rxx_ind=rxx_ind, **kw)
                    ry_ind += 1
            if cooling and step < 7:                    # CS:790-836, 762-772
                # quirk: sympathetic_cooling() does not forward stab_ind (CS:767, 771), so the idle
                # slots of a second cycle index the FIRST half of p_idle again.
                for q in range(N_QUBITS):
                    self._insert_errors(st, "Idle", (q,), leaked, gates, probs, rnd,
                                        sc_ind=step * N_QUBITS + q, stab_ind=0)
        if gates_only:                                  # test hook: state just before All(Measure)
            return None
        outcomes = [self._measure(st, N_DATA + j, meas) for j in range(N_ANC)]   # CS:848-851
        syndrome = list(outcomes)
        for j in range(N_ANC):                          # CS:854-857
            if leaked[N_DATA + j] == 1:
                syndrome[j] = 1
                leaked[N_DATA + j] = 0
        if reset:                                       # CS:859-862 (tests the raw outcome, not the override)
            for j in range(N_ANC):
                if outcomes[j] == 1:
                    self._gate(st, "X", (N_DATA + j,))
        return syndrome

    # ---- prep, decode, read-out ----------------------------------------------------------
    def logical_prep(self, st, basis, state, leaked, gates, probs, rnd, meas):
        """CS:96-123: optional X/H layer, then one noiseless cycle (errors still *draw*)."""
        if state == 1:
            for i in range(N_DATA):
                self._gate(st, "X", (i,))
        if basis == "X":
            for i in range(N_DATA):
                self._gate(st, "H", (i,))
        errorless = {k: len(v) * [0] for k, v in probs.items()}
        return np.array(self.stabilizer_cycle(st, leaked, gates, errorless, rnd, meas))

    def lookup(self, syndrome):
        key = str(np.asarray(syndrome)).strip("[,]")     # CS:887
        return self.correction_table[key][0]

    def apply_correction(self, st, vec):
        for i in range(N_DATA):                         # CS:901-907
            if vec[i] == 1:
                self._gate(st, "X", (i,))
            if vec[i + N_DATA] == 1:
                self._gate(st, "Z", (i,))

    def logical_measurement(self, st, basis, quiescent, leaked, meas):
        if basis == "X":
            for i in range(N_DATA):
                self._gate(st, "H", (i,))
        dm = [self._measure(st, i, meas) for i in range(N_DATA)]     # CS:56-58
        for i in range(N_DATA):                         # CS:61-63
            if leaked[i] == 1:
                dm[i] = 1
        weight = 0
        if basis == "Z":                                # CS:69-79
            q = [quiescent[0], quiescent[2], quiescent[5], quiescent[7]]
            cs = [(dm[0] + dm[1]) % 2, (dm[1] + dm[2] + dm[4] + dm[5]) % 2,
                  (dm[3] + dm[4] + dm[6] + dm[7]) % 2, (dm[7] + dm[8]) % 2]
            csyn = (np.array(cs) - np.array(q)) % 2
            weight = self.bitflip_table[str(csyn).strip("[,]")][1]
        return (sum(dm) + weight) % 2, dm               # CS:86

    # ---- one shot of a protocol ------------------------------------------------------------
    def run_shot(self, model, p_set, protocol, cooling=False, rnd=None, meas=None, basis="Z", state=0,
                 leaked=None, want_state=False, drain=False):
        """One Monte-Carlo shot.  protocol in {'single','s1','s2'} (CL:183, 261, 355).

        p_set: one list per name in model['probabilities'] (masks 0/1 or true probabilities).
        Returns a record dict; ``counted`` says whether the shot reached decode + read-out.
        """
        rnd = rnd if rnd is not None else _random
        meas = meas if meas is not None else SampledMeasure(_random)
        m = self.error_models[model]
        if len(m["probabilities"]) != len(p_set):
            raise ValueError("make sure you pass enough probabilities to specify the error model")
        probs = {k: list(p_set[i]) for i, k in enumerate(m["probabilities"])}
        gates = m["gates"]
        leaked = N_QUBITS * [0] if leaked is None else leaked
        st = _sv.StateVector(N_QUBITS)
        rec = {}
        quiescent = self.logical_prep(st, basis, state, leaked, gates, probs, rnd, meas)
        rec["quiescent"] = [int(b) for b in quiescent]
        synd1 = np.array(self.stabilizer_cycle(st, leaked, gates, probs, rnd, meas, stab_ind=0, cooling=cooling))
        rec["synd1"] = [int(b) for b in synd1]
        flips_a = (quiescent - synd1) % 2
        rec["flips_a"] = [int(b) for b in flips_a]
        s1_stop = bool(np.all(flips_a == 0))
        ft = None
        if protocol == "single":
            ft = flips_a
        elif protocol == "s1":
            ft = flips_a if s1_stop else None
        elif protocol == "s2":
            if not s1_stop:
                synd2 = np.array(self.stabilizer_cycle(st, leaked, gates, probs, rnd, meas, stab_ind=1,
                                                       cooling=cooling))
                rec["synd2"] = [int(b) for b in synd2]
                ft = (quiescent - synd2) % 2             # CL:424
        else:
            raise ValueError(protocol)
        rec["not0"] = int(not s1_stop)
        rec["counted"] = int(ft is not None)
        if ft is not None:
            rec["ft_syndrome"] = [int(b) for b in ft]
            vec = self.lookup(ft)
            rec["correction"] = [int(b) for b in vec]
            self.apply_correction(st, vec)
            if want_state:
                rec["state"] = st.psi.copy()
            logic, dm = self.logical_measurement(st, basis, quiescent, leaked, meas)
            rec["data_meas"] = dm
            rec["logical_meas"] = int(logic)
            rec["fail"] = int(logic != state)
        else:
            rec["fail"] = 0
            if want_state:
                rec["state"] = st.psi.copy()
            if drain:                                   # CL:334-336 / 438-440: measure everything out
                for j in range(N_ANC):
                    self._measure(st, N_DATA + j, meas)
                for i in range(N_DATA):
                    self._measure(st, i, meas)
        rec["leaked"] = list(leaked)
        return rec

    def run_round_rtf(self, model, e_probs, leaked, rnd, meas, basis="Z", state=0, e_probs_b=None, want_record=False):
        """One round of run-til-fail (CL:68-113): fresh |0..0>, prep, cycle A, optional cycle B with the
        SAME (stab_ind=0) probabilities, decode (flips_a + flips_b) % 2, correct, read out.

        e_probs_b (test hook): other lists for cycle B -- a replay passes the 0/1 masks of the slots that fired in each
        cycle.  want_record: return the round's record dict instead of (fail, took_b)."""
        m = self.error_models[model]
        gates = m["gates"]
        st = _sv.StateVector(N_QUBITS)
        quiescent = self.logical_prep(st, basis, state, leaked, gates, e_probs, rnd, meas)
        synd_a = np.array(self.stabilizer_cycle(st, leaked, gates, e_probs, rnd, meas))
        flips_a = (quiescent - synd_a) % 2
        took_b = 0
        synd_b = None
        if np.all(flips_a == 0):
            ft = flips_a
        else:
            synd_b = np.array(self.stabilizer_cycle(st, leaked, gates, e_probs if e_probs_b is None else e_probs_b, rnd, meas))
            took_b = 1
            flips_b = (synd_a - synd_b) % 2
            ft = (flips_a + flips_b) % 2
        vec = self.lookup(ft)
        self.apply_correction(st, vec)
        logic, dm = self.logical_measurement(st, basis, quiescent, leaked, meas)
        if want_record:
            rec = {"quiescent": [int(b) for b in quiescent], "synd1": [int(b) for b in synd_a],
                   "flips_a": [int(b) for b in flips_a], "not0": int(took_b), "counted": 1,
                   "ft_syndrome": [int(b) for b in ft], "correction": [int(b) for b in vec], "data_meas": dm,
                   "logical_meas": int(logic), "fail": int(logic != state), "leaked": list(leaked)}
            if synd_b is not None:
                rec["synd2"] = [int(b) for b in synd_b]
            return rec
        return int(logic != state), took_b


# ---------------------------------------------------------------------------------------------
# fixed-weight placement (CL:151-181)
# ---------------------------------------------------------------------------------------------
def generate_error_locations(num_e, num_loc, rnd=_random):
    if num_e > num_loc:
        raise ValueError("cannot place {} errors on {} locations (the reference loops forever)".format(num_e, num_loc))
    mask = num_loc * [0]
    for _ in range(num_e):
        while True:
            ind = rnd.randint(0, num_loc - 1)
            if mask[ind] == 0:
                mask[ind] = 1
                break
    return mask


def generate_from_distr(probs, rnd=_random):
    r = rnd.random() * sum(probs)
    total = 0
    for i, p in enumerate(probs):
        total += p
        if total > r:
            return i


def generate_error_locations_non_uniform(num_e, num_loc, probs, rnd=_random):
    mask = num_loc * [0]
    for _ in range(num_e):
        while True:
            ind = generate_from_distr(probs, rnd)
            if mask[ind] == 0:
                mask[ind] = 1
                break
    return mask


def run_subset(oracle, protocol, num_runs, model, eset, locations, prob_lists, cooling=False, rnd=None,
               meas_rng=None):
    """The shot loop of CL:183 / CL:261 / CL:355, including the two dead `random.random() < 1`
    basis/state draws (CL:274-281) so the error-draw stream stays aligned with the reference."""
    rnd = rnd if rnd is not None else _random
    meas = SampledMeasure(meas_rng if meas_rng is not None else _random)
    failed = not0 = 0
    for _ in range(num_runs):
        rnd.random()
        rnd.random()
        p_set = []
        for i in range(len(eset)):
            if len(prob_lists[i]) == 1:
                p_set.append(generate_error_locations(eset[i], locations[i], rnd))
            else:
                p_set.append(generate_error_locations_non_uniform(eset[i], locations[i], prob_lists[i], rnd))
        rec = oracle.run_shot(model, p_set, protocol, cooling=cooling, rnd=rnd, meas=meas, drain=True)
        failed += rec["fail"]
        not0 += rec["not0"]
    return failed, not0


def run_til_fail(oracle, num_runs, model, e_probs, rnd=None, meas_rng=None, leak_reset="never", max_rounds=None):
    """The run loop of calculate_log_e_rate (CL:33-148): every run repeats fresh-start rounds (run_round_rtf) until the
    logical read-out is wrong.  Returns (rounds_til_fail_list, syndrome_b_measured_list).

    leak_reset: 'never' = the reference, whose leak register is created once for all runs (CL:49, SURVEY C-6);
    'run' / 'round' clear it at the start of every run / round (the variants the batched backend also offers).
    The two dead `random.random() < 1` basis/state draws per run (CL:54-61) are consumed so that the error-draw stream
    stays aligned with the reference for a given seed."""
    rnd = rnd if rnd is not None else _random
    meas = SampledMeasure(meas_rng if meas_rng is not None else _random)
    if leak_reset not in ("never", "run", "round"):
        raise ValueError(leak_reset)
    leaked = N_QUBITS * [0]
    rounds_list, syndb_list = [], []
    for _ in range(num_runs):
        rnd.random()
        rnd.random()
        if leak_reset == "run":
            leaked = N_QUBITS * [0]
        rounds = took_b_total = 0
        while True:
            if leak_reset == "round":
                leaked = N_QUBITS * [0]
            fail, took_b = oracle.run_round_rtf(model, e_probs, leaked, rnd, meas)
            rounds += 1
            took_b_total += took_b
            if fail or (max_rounds and rounds >= max_rounds):
                break
        rounds_list.append(rounds)
        syndb_list.append(took_b_total)
    return rounds_list, syndb_list
