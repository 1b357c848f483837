"""The object-oriented interface: structural counts pinned by the reference's test_logical_qubit.py:26-39, the
mapping of OO locations onto the functional cycle's error slots, and (GPU) the batched driver."""
import random

import pytest

from surface_17_iqt_b200 import logical_qubit as lq
from surface_17_iqt_b200.error import DephasingError, Error, XCtrlError, YCtrlError, create_error_subset_list


def test_stabilizer_cycle_location_counts():
    cyc = lq.StabiliserCycle(data=list(range(9)), ancilla=list(range(8)), location=[0], error_model=lq.pdd_error_model)
    assert len(cyc.all_gate_locations) == 60
    assert len(cyc.gate_locations[lq.Rxx]) == 24
    assert len(cyc.gate_locations[lq.Rx]) == 12
    assert len(cyc.gate_locations[lq.Ry]) == 24
    assert len(cyc.error_locations[YCtrlError]) == 24
    assert len(cyc.error_locations[XCtrlError]) == 36
    assert len(cyc.error_locations[DephasingError]) == 60
    assert cyc.all_gate_locations[0] == [0, 0, 0, 0] and all(len(loc) == 4 for loc in cyc.all_gate_locations)


def test_noisy_gate_errors_and_error_list():
    g = lq.NoisyGate(lq.Rxx(1.57), [0, 9], [0, 0, 0, 0], error_model=lq.pdd_error_model)
    assert sorted(type(e).__name__ for e in g.possible_errors) == ["DephasingError", "XCtrlError"]
    cyc = lq.StabiliserCycle(location=[1], error_model=lq.pdd_error_model)
    random.seed(4)
    errs = Error.generate_error_list({"XCtrlError": 5, "DephasingError": 6}, cyc.error_locations)
    assert len(errs) == 11 and all(e.location in cyc.all_gate_locations for e in errs)
    assert create_error_subset_list() == [{"DephasingError": i} for i in range(6)]


def test_location_to_slot_mapping():
    cyc = lq.StabiliserCycle(location=[1], error_model=lq.pdd_error_model)
    # every location, one dephasing error each: 54 functional gates get Z (Rxx: both qubits), 6 locations are dropped
    errs = [DephasingError(location=loc) for loc in cyc.all_gate_locations]
    masks = cyc.slot_masks(errs)
    assert len(cyc.dropped_locations) == 6
    assert sum(masks[3]) == 12 + 18 + 24 + 24 and [len(m) for m in masks] == [12, 18, 24, 78, 24]
    x = cyc.slot_masks([XCtrlError(location=[1, 0, 0, 0]), XCtrlError(location=[1, 0, 1, 1])])
    assert x[2][0] == 1 and x[0][0] == 1 and sum(map(sum, x)) == 2
    twice = cyc.slot_masks([XCtrlError(location=[1, 0, 0, 0]), XCtrlError(location=[1, 0, 0, 0])])
    assert sum(map(sum, twice)) == 0                      # two equal Paulis cancel
    with pytest.raises(ValueError):
        cyc.slot_masks([YCtrlError(location=[1, 0, 0, 0])])


@pytest.mark.gpu
def test_error_subset_error_rate_driver():
    from surface_17_iqt_b200.error_subset_error_rate import calc_error_subset_error_rate
    random.seed(1)
    assert calc_error_subset_error_rate({"DephasingError": 0}, num_runs=32, verbose=False) == 0.0
    # one fault can defeat single-round decoding (that is why the s1/s2 syndrome processing exists), but rarely
    assert calc_error_subset_error_rate({"DephasingError": 1}, num_runs=256, verbose=False) < 0.15
    r = calc_error_subset_error_rate({"XCtrlError": 5, "DephasingError": 6}, num_runs=200, verbose=False)
    assert 0.05 < r < 0.95


@pytest.mark.gpu
def test_importance_sampling_pipeline_end_to_end(tmp_path):
    """All stages of the reference's importance_sampling_syndrome_processing.py on the device path."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("is_pipeline", os.path.join(here, "examples",
                                                                             "importance_sampling_syndrome_processing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.run_pipeline("PDD_model", runs=2000, cutoff=2e-3, out=str(tmp_path), verbose=False)
    assert res["s1_subsets"] >= 5 and res["s2_subsets"] >= 5
    assert 0.0 < res["p_logical"] < 0.5
    for name in ("incremental_subsets_s1_PDD_model.json", "incremental_subsets_s2_PDD_model.json",
                 "ImpSamp_2000_syndrome_process_s1subsets_PDD_model.json",
                 "ImpSamp_2000_syndrome_process_s2subsets_PDD_model.json"):
        assert os.path.exists(os.path.join(str(tmp_path), name))
    res = mod.run_pipeline("Dress_cooling", runs=1000, cutoff=5e-3, out=str(tmp_path), verbose=False)
    assert res["s1_subsets"] >= 3
