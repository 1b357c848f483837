"""Host-side logic (no GPU): the compiled cycle equals the reference's traced gate stream, slot
indexing / list lengths, program structure, table packing and the C-ABI export list."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from surface_17_iqt_b200 import _lib as L
from surface_17_iqt_b200 import backend, circuit

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reduce(theta):
    return round(float(theta) % (4.0 * np.pi), 12)


def test_gate_stream_equals_reference_trace(golden):
    trace = golden("cycle_trace.json")["trace"]
    mine = circuit.gate_stream()
    assert len(mine) == len(trace) == 54
    for (name, angle, bits), (rn, ra, rb) in zip(mine, trace):
        assert name == rn and list(bits) == rb and abs(_reduce(angle) - ra) < 1e-12


def test_list_lengths(error_models):
    assert circuit.list_lengths(error_models["PDD_model"]) == [12, 18, 24, 78, 24]
    assert circuit.list_lengths(error_models["Dress_cooling"]) == [12, 18, 24, 78, 24, 119]
    assert circuit.list_lengths(error_models["PDD_model"], two_rounds=True) == [24, 36, 48, 156, 48]
    assert circuit.list_lengths(error_models["depolarising_model"]) == [30, 30, 30, 24]
    assert circuit.list_lengths(error_models["coherent_model"]) == [30, 24, 24]


def test_leak_index_quirk_preserved():
    ops = circuit.build_cycle()
    op = [o for o in ops if o.timestep == 4][1]
    assert (op.data, op.ancilla) == (3, 5) and (op.d_leak, op.a_leak) == (1, 11)    # CS:559-561
    for o in ops:
        if o is not op:
            assert (o.d_leak, o.a_leak) == (o.data, 9 + o.ancilla)


@pytest.mark.parametrize("fusion,n_layers", [(0, 54), (1, 8), (2, 4)])
def test_program_shapes(error_models, fusion, n_layers):
    prog = circuit.compile_program(error_models["PDD_model"], [12, 18, 24, 78, 24], fusion=fusion)
    assert len(prog.layers[0]) == len(prog.layers[1]) == len(prog.layers[2]) == n_layers
    n_uops = sum(op.n_uops for lay in prog.layers[0] for op in lay.ops[:lay.n_ops])
    assert n_uops == 54
    assert [lay.meas_mask for lay in prog.layers[0]] == [0] * (n_layers - 1) + [0xFF]
    early = circuit.compile_program(error_models["PDD_model"], [12, 18, 24, 78, 24], fusion=fusion, early_measure=True)
    masks = [lay.meas_mask for lay in early.layers[1]]
    assert [m for m in masks if m] == [0x5A, 0xA5] and masks[-1] == 0xA5
    for lay in prog.layers[0]:
        for op in lay.ops[:lay.n_ops]:
            assert op.anc_bit in list(lay.anc_bits)[:lay.n_anc]
    errs = [p for p in prog.plan if p.kind == L.P_ERR]
    # 12 Rx x 2 + 18 Ry x 2 + 24 Rxx x 4 entries
    assert len(errs) == 12 * 2 + 18 * 2 + 24 * 4
    for p in errs:
        assert p.slot < prog.list_len[p.list]


def test_cooling_program(error_models):
    prog = circuit.compile_program(error_models["PDD_cooling"], [12, 18, 24, 78, 24, 119], fusion=2, cooling=True)
    idle = [p for p in prog.plan if p.kind == L.P_IDLE]
    assert len(idle) == 119 and [p.slot for p in idle] == list(range(119))
    assert all(p.use_round == 0 for p in idle)
    assert sum(1 for lay in prog.layers[0] for op in lay.ops[:lay.n_ops] if op.kind == L.OP_IDLE) == 7
    with pytest.raises(KeyError):
        circuit.compile_program(error_models["PDD_model"], [12, 18, 24, 78, 24], cooling=True)
    with pytest.raises(ValueError):
        circuit.compile_program(error_models["PDD_model"], [12, 18, 24, 78])


def test_coherent_program(error_models):
    eps1, eps2 = [0.02] * 30, [0.02] * 24
    prog = circuit.compile_program(error_models["coherent_model"], [30, 24, 24], fusion=1,
                                   over_angles={"eps_1q": eps1, "eps_2q": eps2})
    assert not [u for lay in prog.layers[0] for op in lay.ops[:lay.n_ops] for u in op.uops[:op.n_uops]
                if u.type >= L.U_ROT_X]              # preparation cycle stays noiseless
    rot = [u.type for lay in prog.layers[1] for op in lay.ops[:lay.n_ops] for u in op.uops[:op.n_uops]
           if u.type >= L.U_ROT_X]
    assert len(rot) == 54
    assert abs(prog.params[0] - np.cos(0.01)) < 1e-15 and abs(prog.params[1] - np.sin(0.01)) < 1e-15


def test_table_packing(correction_table, bitflip_table):
    lut = backend.pack_correction_table(correction_table)
    assert lut[0] == 0
    vec = correction_table["0 0 0 0 0 0 0 1"][0]       # ancilla 7 flipped -> X on data 8
    assert lut[1 << 7] == sum(v << i for i, v in enumerate(vec)) == 1 << 8
    cw = backend.pack_bitflip_table(bitflip_table)
    assert cw[0] == 0 and cw[0b0101] == 2 and cw[0b1000] == 1


def test_library_exports_every_declared_symbol():
    """libs17.so loads without a GPU and exports exactly what include/s17.h declares."""
    lib_path = os.path.join(REPO, "surface_17_iqt_b200", "libs17.so")
    if not os.path.exists(lib_path):
        import __graft_entry__ as g
        g.build_cuda()
    header = open(os.path.join(REPO, "include", "s17.h")).read()
    declared = set(re.findall(r"\b(s17_[a-z_0-9]+)\s*\(", header))
    assert declared == set(L.EXPORTS)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.s17_device_count() >= 0


def test_no_silent_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.S17Error):
        backend.Backend(4)


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "surface_17_iqt_b200")
    out = subprocess.run(["grep", "-rIl", "--include=*.py", "--include=*.cu", "--include=*.cuh", "-E",
                          r"(from|import) +oracle|oracle/", pkg], capture_output=True, text=True).stdout.strip()
    assert out == "", out


def test_host_lowering_selects_the_whole_cycle_kernel(error_models):
    """s17_describe_program (host only): the standard cycle gets the whole-cycle kernel with label maps in which every
    ancilla slot indexes its table with exactly one register bit -- the unique assignment {0,3,5,8} / {1,2,6,7} with data
    bit 4 left to the shuffle stage -- for the Pauli / leakage / cooling models, and only the noiseless preparation
    cycle for the coherent model; without the early X-ancilla measurement there are no half cycles at all."""
    from surface_17_iqt_b200 import backend
    ops = circuit.build_cycle()
    for model, cooling in (("PDD_model", False), ("Dress_model", False), ("depolarising_model", False),
                           ("PDD_cooling", True), ("Dress_cooling", True)):
        prog = circuit.compile_program(error_models[model], circuit.list_lengths(error_models[model]), fusion=2,
                                       cooling=cooling, early_measure=True)
        infos = backend.describe_program(prog)
        # the noiseless preparation cycle is the same whole-cycle program as the first noisy cycle (events aside)
        assert infos[1]["prep_rides_along"], model
        for info in infos:
            assert info["k_half"] and info["k_cycle"] and info["k_cycle1"], (model, info)
            assert sorted(info["reg_bits_z"]) == [0, 3, 5, 8] and sorted(info["reg_bits_x"]) == [1, 2, 6, 7]
            assert info["shuffle_bit"] == 4 and info["table_entries"] == [40, 40]
            for h, kind, regs in ((0, "x", info["reg_bits_x"]), (1, "z", info["reg_bits_z"])):
                assert sorted(info["slot_ancilla"][h]) == sorted({o.ancilla for o in ops if o.kind == kind})
                for k, anc in enumerate(info["slot_ancilla"][h]):
                    data = {o.data for o in ops if o.kind == kind and o.ancilla == anc}
                    assert data & set(regs) == {regs[k]}, (model, h, k, data, regs)      # slot k <-> register bit k
    coherent = circuit.compile_program(error_models["coherent_model"], [30, 24, 24], fusion=2, early_measure=True,
                                       over_angles={"eps_1q": [0.02] * 30, "eps_2q": [0.02] * 24})
    info = backend.describe_program(coherent)
    # coherent over-rotations: the noisy cycles keep the maps of their Clifford skeleton and run on the coherent form of the
    # whole-cycle kernel (k_half / the generic-map k_cycle do not apply); the preparation cycle is the plain one
    assert [i["k_cycle1"] for i in info] == [True, True, True]
    assert [i["coherent"] for i in info] == [False, True, True] and not info[1]["prep_rides_along"]
    assert [i["k_cycle"] for i in info] == [True, False, False] and [i["k_half"] for i in info] == [True, False, False]
    plain = backend.describe_program(circuit.compile_program(error_models["PDD_model"], [12, 18, 24, 78, 24], fusion=2,
                                                            early_measure=True))
    for c in (1, 2):
        for key in ("reg_bits_z", "reg_bits_x", "shuffle_bit", "slot_ancilla", "table_entries"):
            assert info[c][key] == plain[c][key]
    # ... but only while every stochastic entry stays a Pauli in the frame next to the rotations: a Z on the data qubit
    # after an over-rotated Rx is a basis flip in the Hadamard frame that does not commute with them -> sweep kernels
    import copy
    flipping = copy.deepcopy(error_models["coherent_model"])
    flipping["gates"]["Rx"].append(["Z", "p_XX_ctrl", "test: dephasing after Rx"])
    info = backend.describe_program(circuit.compile_program(flipping, [30, 24, 24], fusion=2, early_measure=True,
                                                            over_angles={"eps_1q": [0.02] * 30, "eps_2q": [0.02] * 24}))
    assert [i["k_cycle1"] for i in info] == [True, False, False] and not any(i["coherent"] for i in info)
    late = circuit.compile_program(error_models["PDD_model"], [12, 18, 24, 78, 24], fusion=2, early_measure=False)
    assert not any(i["k_half"] or i["k_cycle"] for i in backend.describe_program(late))


def test_allreduce_counts_sums_contexts():
    """s17_allreduce_counts: the host-side sum over the contexts of one process (no device needed)."""
    import ctypes as C
    from surface_17_iqt_b200 import _lib as L
    lib = L.load()
    parts = (L.Counts * 3)()
    for i, c in enumerate(parts):
        c.shots, c.failed, c.not0, c.counted, c.errors, c.sweeps, c.rounds, c.sweep_bytes = (
            100 + i, 3 * i, 7 + i, 90 + i, 0, 200 * (i + 1), i, 1 << (20 + i))
    out = L.Counts()
    assert lib.s17_allreduce_counts(parts, 3, C.byref(out)) == 0
    assert (out.shots, out.failed, out.not0, out.counted, out.sweeps, out.rounds) == (303, 9, 24, 273, 1200, 3)
    assert out.sweep_bytes == (1 << 20) + (1 << 21) + (1 << 22)
    assert lib.s17_allreduce_counts(parts, 0, C.byref(out)) != 0
