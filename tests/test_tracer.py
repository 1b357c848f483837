"""The circuit tracer (surface_17_iqt_b200/tracer.py) and the projectq-shaped face without a device.

With /root/reference present (the build container) the reference's UNMODIFIED
compiled_surface_code_arbitrary_error_model.py is imported on top of surface_17_iqt_b200.pq_compat and traced; its cycle
table, slot indices and error effects must equal what circuit.compile_program uses, and the trace is committed as
tests/golden/traced_cycle.json so that the same comparison runs where the reference is absent (the GPU box)."""
import contextlib
import ctypes
import io
import json
import os
import sys

import pytest

from surface_17_iqt_b200 import api, circuit, pq_compat, pq_core, tracer

REF = os.environ.get("S17_REFERENCE", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traced_cycle.json")
MODELS = [("PDD_model", False), ("Dress_model", False), ("PDD_cooling", True), ("Dress_cooling", True)]


def reference_module():
    if not os.path.exists(os.path.join(REF, "compiled_surface_code_arbitrary_error_model.py")):
        pytest.skip("reference checkout not present")
    pq_compat.install()
    if REF not in sys.path:
        sys.path.insert(1, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import compiled_surface_code_arbitrary_error_model as cs_ref
    finally:
        os.chdir(cwd)
    assert cs_ref.__file__.startswith(REF)
    return cs_ref


def as_json(tr):
    return {"ops": [[o.timestep, o.index, o.kind, o.data, o.ancilla, o.s, int(o.has_ry1), int(o.has_rx), int(o.has_ry2),
                     o.d_leak, o.a_leak, o.rx_ind, o.ry1_ind, o.ry2_ind] for o in tr.ops],
            "entries": [[k[0], k[1], k[2], v[0], v[1], [list(e) for e in v[2]]] for k, v in sorted(tr.entries.items())],
            "idle": [list(i) for i in tr.idle], "n_gates": tr.n_gates}


def program_bytes(prog):
    out = [bytes(ctypes.string_at(ctypes.addressof(lay), ctypes.sizeof(lay))) for cyc in prog.layers for lay in cyc]
    out += [bytes(ctypes.string_at(ctypes.addressof(p), ctypes.sizeof(p))) for p in prog.plan]
    return out, list(prog.params), list(prog.list_len)


def check_trace(tr, model, list_len, cooling, stab_ind):
    """A traced cycle against the tables of circuit.py: operations, slots, effects, and the lowered program."""
    assert tr.ops == circuit.build_cycle()
    assert tr.n_gates == 54
    tracer.check_effects(tr, model)
    ops = {o.index: o for o in tr.ops}
    for (oi, pos, k), (pname, slot, _eff) in tr.entries.items():
        op = ops[oi]
        gate = {"Ry1": "Ry", "Ry2": "Ry", "Rxx": "Rxx", "Rx": "Rx"}[pos]
        err = model["gates"][gate][k][0]
        ry = op.ry1_ind if pos == "Ry1" else op.ry2_ind
        want = circuit.slot_index(pname, gate, err, max(op.rx_ind, 0), max(ry, 0), op.index)
        length = list_len[model["probabilities"].index(pname)]
        assert slot == want + (length // 2 if stab_ind else 0), (oi, pos, k)          # CS:191-193
    if cooling:
        # 17 Idle slots per gap after timesteps 1-7, data 0-8 then ancilla 0-7; stab_ind is NOT forwarded (CS:767)
        assert [(t, q, s) for t, q, _n, s in tr.idle] == [(t, q, 17 * t + q) for t in range(7) for q in range(17)]
    else:
        assert tr.idle == []


@pytest.mark.parametrize("name,cooling", MODELS)
def test_reference_cycle_traces_to_the_compiled_program(name, cooling):
    cs_ref = reference_module()
    model = api.load_error_models()[name]
    one = circuit.list_lengths(model)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = tracer.trace_cycle(cs_ref, model, one, cooling=cooling)
        two = [2 * x for x in one]
        tr2 = tracer.trace_cycle(cs_ref, model, two, stab_ind=1, cooling=cooling)
    check_trace(tr, model, one, cooling, 0)
    check_trace(tr2, model, two, cooling, 1)
    # the program lowered from the reference's trace is byte-identical with the one compiled from circuit.py's tables
    for fusion, early in ((0, False), (1, False), (2, True)):
        a = circuit.compile_program(model, two, fusion=fusion, cooling=cooling, early_measure=early)
        b = circuit.compile_program(model, two, fusion=fusion, cooling=cooling, early_measure=early, ops=tr.ops,
                                    slots=tr.slots(), idle_slots={(t, q): s for t, q, _n, s in tr.idle} if cooling else None)
        assert program_bytes(a) == program_bytes(b)
    golden = json.load(open(GOLDEN))
    assert golden[name] == as_json(tr), "tests/golden/traced_cycle.json is stale: python tests/test_tracer.py"


@pytest.mark.parametrize("name,cooling", MODELS)
def test_own_cycle_functions_emit_the_reference_stream(name, cooling):
    """This package's gate-by-gate functions (compiled_surface_code_arbitrary_error_model.py) traced the same way give
    the trace committed from the reference: same operations, slots, effects, idle slots."""
    import surface_17_iqt_b200.compiled_surface_code_arbitrary_error_model as cs_own
    model = api.load_error_models()[name]
    one = circuit.list_lengths(model)
    tr = tracer.trace_cycle(cs_own, model, one, cooling=cooling)
    check_trace(tr, model, one, cooling, 0)
    assert json.load(open(GOLDEN))[name] == as_json(tr)
    tr2 = tracer.trace_cycle(cs_own, model, [2 * x for x in one], stab_ind=1, cooling=cooling)
    check_trace(tr2, model, [2 * x for x in one], cooling, 1)


def test_engine_face_command_stream_without_a_device():
    """The reference's test_noisy_gate (test_logical_qubit.py:11-24) against the recording backend: two allocations, two
    gates and two measurements are six commands on a DummyEngine(save_commands=True)."""
    from math import pi
    from surface_17_iqt_b200.logical_qubit import NoisyGate
    dummy = pq_core.DummyEngine(save_commands=True)
    eng = pq_core.MainEngine(backend=pq_core.RecordingBackend(), engine_list=[dummy])
    data = eng.allocate_qureg(1)
    ancilla = eng.allocate_qureg(1)
    NoisyGate(pq_core.Rxx(pi / 2), (data[0], ancilla[0])).apply()
    NoisyGate(pq_core.Rx(pi / 2), (data[0])).apply()
    pq_core.All(pq_core.Measure) | data
    pq_core.All(pq_core.Measure) | ancilla
    assert len(dummy.received_commands) == 6
    assert [c.gate.name for c in dummy.received_commands] == ["Allocate", "Allocate", "Rxx", "Rx", "Measure", "Measure"]
    eng.flush()
    assert int(data[0]) == 0


def test_simulator_without_a_device_fails_loudly():
    """No CPU fallback: a Simulator on a box without CUDA raises at the first allocation."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from surface_17_iqt_b200 import pq_core, _lib\n"
            "try:\n"
            "    pq_core.MainEngine(pq_core.Simulator()).allocate_qureg(1)\n"
            "except _lib.S17Error as e:\n"
            "    print('RAISED', e)\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert "RAISED" in out.stdout and "no CUDA device" in out.stdout, out.stdout + out.stderr


if __name__ == "__main__":          # mint the golden from the reference (build container only)
    cs = reference_module()
    out = {}
    for name_, cooling_ in MODELS:
        model_ = api.load_error_models()[name_]
        with contextlib.redirect_stdout(io.StringIO()):
            out[name_] = as_json(tracer.trace_cycle(cs, model_, circuit.list_lengths(model_), cooling=cooling_))
    json.dump(out, open(GOLDEN, "w"))
    print("wrote", GOLDEN)
