"""GPU tests of the engine-level face: the projectq-shaped `MainEngine(Simulator())` over the `s17_eng_*` C ABI, and the
gate-by-gate circuit functions of surface_17_iqt_b200.compiled_surface_code_arbitrary_error_model on top of it.

The reference is not on the GPU box, so these tests drive (a) the reference's traced gate stream
(tests/golden/cycle_trace.json, minted from the unmodified reference) and (b) this package's circuit functions -- which
tests/test_tracer.py shows to emit the reference's command stream -- and compare states with the oracle (1e-10 per
amplitude) and records with the reference's golden replays (bit-exact)."""
import os
import random
from math import pi

import numpy as np
import pytest

from oracle import s17_oracle as so
from oracle import sv
from surface_17_iqt_b200 import _lib as L
from surface_17_iqt_b200 import compiled_surface_code_arbitrary_error_model as CS
from surface_17_iqt_b200 import pq_core
from surface_17_iqt_b200.logical_qubit import NoisyGate, StabiliserCycle, pdd_error_model
from surface_17_iqt_b200.pq_core import All, DummyEngine, H, MainEngine, Measure, Rx, Rxx, Ry, Simulator, X, Z

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AMP_TOL = 1e-10
GATES = {"Rx": Rx, "Ry": Ry, "Rxx": Rxx}


def test_traced_reference_gate_stream_amplitudes(golden):
    """The 54 gates the unmodified reference emits for one cycle, as `Gate | qubits` on a 3-shot engine:
    the full 2^17 state of every shot equals the oracle's after each gate, then All(Measure) | ancilla with forced
    outcomes collapses like the oracle does."""
    trace = golden("cycle_trace.json")["trace"]
    backend = Simulator(batch=3, rnd_seed=1)
    eng = MainEngine(backend)
    data, ancilla = eng.allocate_qureg(9), eng.allocate_qureg(8)
    qubit = list(data) + list(ancilla)
    st = sv.StateVector(17)
    orc = so.ShotOracle({}, {}, {})
    for k, (name, angle, bits) in enumerate(trace):
        GATES[name](angle) | tuple(qubit[b] for b in bits)
        orc._gate(st, name, bits, angle)
        if k % 6 == 5 or k == len(trace) - 1:
            for shot in (0, 2):
                err = np.max(np.abs(backend.state(shot) - st.psi))
                assert err < AMP_TOL, (k, name, bits, shot, err)
    outcomes = [0, 1, 0, 0, 1, 0, 1, 0]
    backend.forced.extend(outcomes)
    All(Measure) | ancilla
    eng.flush()
    assert [int(q) for q in ancilla] == outcomes
    for j, o in enumerate(outcomes):
        assert st.collapse(9 + j, o) > 0
    assert np.max(np.abs(backend.state(1) - st.psi)) < AMP_TOL
    assert backend.impossible == 0
    backend.forced.append(1)                      # a Z-type ancilla cannot flip without an error: probability zero
    Measure | ancilla[0]
    assert backend.impossible == 1
    backend.close()


def run_protocol(eng, backend, model, p_set, protocol, cooling, correction_table, bitflip_table):
    """The loop body of CL:261-353 / 355-457 / 183-259 with this package's circuit functions (same call sequence as
    the reference's drivers)."""
    leaked = 17 * [0]
    e_model, probs = CS.instantiate_error_model(p_set, model)
    data, ancilla = eng.allocate_qureg(9), eng.allocate_qureg(8)
    quiescent = CS.logical_prep(data, "Z", 0, ancilla, leaked, eng, e_model, probs, False)
    synd1 = np.array(CS.stabilizer_cycle_error_index(data, ancilla, leaked, eng, e_model, probs, reset=True, cooling=cooling))
    flips_a = (np.array(quiescent) - synd1) % 2
    rec = {"quiescent": [int(b) for b in quiescent], "synd1": [int(b) for b in synd1], "flips_a": [int(b) for b in flips_a]}
    stop = bool(np.all(flips_a == 0))
    ft = None
    if protocol == "single":
        ft = flips_a
    elif protocol == "s1":
        ft = flips_a if stop else None
    elif protocol == "s2" and not stop:
        synd2 = np.array(CS.stabilizer_cycle_error_index(data, ancilla, leaked, eng, e_model, probs, stab_ind=1, reset=True,
                                                         cooling=cooling))
        rec["synd2"] = [int(b) for b in synd2]
        ft = (np.array(quiescent) - synd2) % 2
    rec["not0"], rec["counted"] = int(not stop), int(ft is not None)
    eng.flush()
    rec["pre_block"] = backend.state(0)[:512].copy()
    if ft is not None:
        vec = CS.lookup(ft, correction_table)
        CS.apply_correction(vec, data)
        eng.flush()
        rec["ft_syndrome"], rec["correction"] = [int(b) for b in ft], [int(b) for b in vec]
        rec["block"] = backend.state(1)[:512].copy()
        rec["outside"] = float(np.sum(np.abs(backend.state(1)[512:]) ** 2))
        logic = CS.logical_measurement(data, "Z", eng, bitflip_table, quiescent, leaked)
        dm = [1 if leaked[i] else int(q) for i, q in enumerate(data)]
        rec.update(data_meas=dm, logical_meas=int(logic), fail=int(logic != 0))
    else:
        rec["fail"] = 0
        All(Measure) | ancilla
        All(Measure) | data
        eng.flush()
    rec["leaked"] = list(leaked)
    for q in list(data) + list(ancilla):          # the registers go out of scope in the reference (CL:70-71): release them
        eng.release(q.id)
    return rec


def test_circuit_functions_on_the_engine_reproduce_reference_replays(error_models, correction_table, bitflip_table, golden):
    """Gate by gate on the GPU (`MainEngine(Simulator(batch=2))`, one sweep per gate, one reduction + collapse per
    Measure): the reference's golden replays -- explicit 0/1 slot lists, recorded outcomes forced -- give the reference's
    records bit for bit and its 512-amplitude blocks before and after the correction within 1e-10."""
    reps = golden("replays.json")["replays"]
    npz = np.load(os.path.join(GOLDEN, "replay_blocks.npz"))
    fields = ("quiescent", "synd1", "synd2", "flips_a", "ft_syndrome", "correction", "data_meas", "logical_meas", "fail",
              "not0", "counted", "leaked")
    picked = [r for i, r in enumerate(reps) if i % 2 == 0]
    for r in picked:
        # a fresh engine per shot, like the reference's drivers (CL:299): a re-used engine carries the global phase of
        # the previous shot's final basis state into the next one
        backend = Simulator(batch=2, rnd_seed=3)
        eng = MainEngine(backend)
        backend.forced.extend(r["outcomes"])
        rec = run_protocol(eng, backend, r["model"], r["p_set"], r["protocol"], r["cooling"], correction_table, bitflip_table)
        for f in fields:
            assert rec.get(f) == r.get(f), (r["model"], r["protocol"], f, rec.get(f), r.get(f))
        assert np.max(np.abs(rec["pre_block"] - npz["pre_blocks"][r["pre_block_index"]])) < AMP_TOL
        if r["block_index"] >= 0:
            assert np.max(np.abs(rec["block"] - npz["blocks"][r["block_index"]])) < AMP_TOL and rec["outside"] < 1e-20
        assert not backend.forced and backend.impossible == 0
        assert sorted(backend.free) == list(range(17)) and not backend.pos      # every bit came back
        backend.close()
    assert len(picked) >= 18


def test_noisy_gate_and_stabiliser_cycle_like_the_reference_tests():
    """test_logical_qubit.py:11-39 of the reference, on the GPU backend."""
    dummy = DummyEngine(save_commands=True)
    eng = MainEngine(backend=Simulator(), engine_list=[dummy])
    data = eng.allocate_qureg(1)
    ancilla = eng.allocate_qureg(1)
    NoisyGate(Rxx(pi / 2), (data[0], ancilla[0])).apply()
    NoisyGate(Rx(pi / 2), (data[0])).apply()
    All(Measure) | data
    All(Measure) | ancilla
    assert len(dummy.received_commands) == 6
    eng.flush()
    assert int(data[0]) in (0, 1) and int(ancilla[0]) in (0, 1)

    backend = Simulator()
    dummy = DummyEngine(save_commands=True)
    eng = MainEngine(backend, [dummy])
    data = eng.allocate_qureg(9)
    ancilla = eng.allocate_qureg(8)
    syndrome0 = StabiliserCycle(data=data, ancilla=ancilla, location=[0], error_model=pdd_error_model)
    assert len(syndrome0.all_gate_locations) == 60
    assert len(syndrome0.gate_locations[Rxx]) == 24
    assert len(syndrome0.gate_locations[Rx]) == 12
    assert len(syndrome0.gate_locations[Ry]) == 24
    # and the cycle object really drives the engine: 60 gates arrive
    syndrome0.run()
    assert sum(1 for c in dummy.received_commands if c.gate.name in ("Rxx", "Rx", "Ry")) == 60
    All(Measure) | ancilla
    All(Measure) | data
    eng.flush()


def test_allocation_deallocation_semantics():
    """ProjectQ's rules the reference relies on: a qubit can only go once it is classical (hence CL:334-336); the
    run-til-fail loop allocates its next registers BEFORE the old ones are dropped (CL:70-71) -- the new qubits take
    the freed bits in allocation order, and a bit that held |1> is returned as |0>."""
    backend = Simulator(batch=2, rnd_seed=5)
    eng = MainEngine(backend)
    q = eng.allocate_qureg(17)
    H | q[3]
    with pytest.raises(pq_core.NotClassicalError):
        eng.release(q[3].id)
    X | q[5]
    backend.forced.append(1)
    Measure | q[3]
    All(Measure) | [x for i, x in enumerate(q) if i != 3]
    eng.flush()
    assert int(q[3]) == 1 and int(q[5]) == 1 and int(q[0]) == 0
    new = eng.allocate_qureg(9)                       # all 17 bits are in use: pending
    assert len(backend.pending) == 9
    for x in q[:9]:
        eng.release(x.id)
    assert [backend.pos[x.id] for x in new] == list(range(9)) and not backend.pending
    psi = backend.state(0)
    assert abs(abs(psi[0]) - 1.0) < 1e-12             # bits 3 and 5 were flipped back before they were handed over
    X | new[5]
    Measure | new[5]
    assert int(new[5]) == 1
    backend.close()


def test_batch_sampling_per_shot_and_predicated_gates():
    """lockstep=False: every shot draws its own outcome (Philox, counter = shot); `int(qubit)` refuses a batch that
    disagrees; a per-shot predicate applies a gate to some shots only (the conditional reset of CS:859-862)."""
    n = 4096
    backend = Simulator(batch=n, n_qubits=10, lockstep=False, rnd_seed=11)
    eng = MainEngine(backend)
    q = eng.allocate_qureg(10)
    H | q[2]
    Rxx(pi / 2) | (q[2], q[7])
    Measure | q[2]
    out = backend.outcomes(q[2])
    assert abs(out.mean() - 0.5) < 5 * 0.5 / np.sqrt(n)
    with pytest.raises(L.S17Error):
        int(q[2])
    # conditional reset: X on the shots that read 1, then every shot reads 0 with certainty
    eng.apply(X, [q[2].id], pred=out)
    Measure | q[2]
    assert int(q[2]) == 0
    # the Rxx made q[7] a copy of q[2]'s X-basis partner: its own outcomes are still 50/50 and independent per shot
    Measure | q[7]
    o7 = backend.outcomes(q[7])
    assert abs(o7.mean() - 0.5) < 5 * 0.5 / np.sqrt(n)
    t = backend.eng.timing(clear=True)
    assert t["gate_calls"] == 3 and t["reduce_calls"] == 3 and t["collapse_calls"] == 3
    assert t["gate_bytes_per_call"] == 2 * n * 16 * 1024
    backend.close()


def test_lockstep_batch_is_copies_of_one_trajectory(error_models, correction_table, bitflip_table):
    """batch = 64 in lock-step: one noisy s2-style shot through the circuit functions; every shot of the batch ends in
    the same state, and that trajectory replays bit-exactly in the oracle."""
    random.seed(12)
    backend = Simulator(batch=64, rnd_seed=21)
    eng = MainEngine(backend)
    model = "Dress_model"
    locs = [24, 36, 48, 156, 48]
    p_set = [[0] * x for x in locs]
    for t, i in ((2, 5), (3, 40), (3, 100), (1, 20)):
        p_set[t][i] = 1
    logged = []
    real_measure = backend.measure

    def spy(qid):
        out = real_measure(qid)
        logged.append(int(out[0]))
        assert np.all(out == out[0])
        return out
    backend.measure = spy
    rec = run_protocol(eng, backend, model, p_set, "s2", False, correction_table, bitflip_table)
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    ref = orc.run_shot(model, p_set, "s2", rnd=random.Random(0), meas=so.ForcedMeasure(logged), drain=True)
    for f in ("quiescent", "synd1", "synd2", "flips_a", "ft_syndrome", "correction", "data_meas", "logical_meas", "fail",
              "not0", "counted", "leaked"):
        assert rec.get(f) == ref.get(f), (f, rec.get(f), ref.get(f))
    backend.close()
