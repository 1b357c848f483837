"""Pin the CPU oracle (oracle/s17_oracle.py) to the golden vectors minted from the unmodified
reference running on oracle/pqshim (oracle/make_golden.py).  CPU only."""
import os
import random

import numpy as np
import pytest

from oracle import s17_oracle as so
from oracle import sv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("quiescent", "synd1", "synd2", "flips_a", "ft_syndrome", "correction", "data_meas",
          "logical_meas", "fail", "not0", "counted", "leaked")


def test_cycle_trace_matches_reference(oracle, golden):
    """The oracle's noiseless cycle emits the reference's 54-gate stream."""
    trace = golden("cycle_trace.json")["trace"]
    oracle.log = []
    try:
        oracle.run_shot("PDD_model", [12 * [0], 18 * [0], 24 * [0], 78 * [0], 24 * [0]], "s1",
                        rnd=random.Random(1), meas=so.SampledMeasure(random.Random(2)))
        gates = [e for e in oracle.log if e[0] == "gate"][:54]
    finally:
        oracle.log = None
    assert len(trace) == 54
    for g, t in zip(gates, trace):
        assert g[1] == t[0] and list(g[3]) == t[2]
        assert abs(g[2] - t[1]) < 1e-12
    names = [t[0] for t in trace]
    assert (names.count("Rxx"), names.count("Rx"), names.count("Ry")) == (24, 12, 18)


def test_replays_bit_exact_and_amplitudes(oracle, golden):
    reps = golden("replays.json")["replays"]
    blocks = np.load(os.path.join(GOLDEN, "replay_blocks.npz"))["blocks"]
    assert len(reps) >= 36
    for r in reps:
        rec = oracle.run_shot(r["model"], r["p_set"], r["protocol"], cooling=r["cooling"],
                              rnd=random.Random(0), meas=so.ForcedMeasure(r["outcomes"]),
                              want_state=True, drain=True)
        for f in FIELDS:
            assert rec.get(f) == r.get(f), (r["model"], r["protocol"], f)
        if r["block_index"] >= 0:
            psi = rec["state"]
            assert np.max(np.abs(psi[:512] - blocks[r["block_index"]])) < 1e-12
            assert np.sum(np.abs(psi[512:]) ** 2) < 1e-24


def test_single_fault_table(oracle, golden):
    """156 single faults of PDD_model: flips_a deterministic, no logical failure."""
    rows = golden("single_fault_pdd.json")["rows"]
    assert len(rows) == 156
    locs = [12, 18, 24, 78, 24]
    undetected = 0
    for row in rows:
        p_set = [2 * l * [0] for l in locs]
        p_set[row["type"]][row["slot"]] = 1
        rec = oracle.run_shot("PDD_model", p_set, "s2", rnd=random.Random(0),
                              meas=so.ForcedMeasure(row["outcomes"]), drain=True)
        assert rec["flips_a"] == row["flips_a"]
        assert rec.get("ft_syndrome") == row["ft"]
        assert rec["fail"] == row["fail"] == 0
        undetected += int(sum(row["flips_a"]) == 0)
    assert undetected == 58


def test_seeded_driver_counts(oracle, golden):
    """Same python `random` seed => the oracle's shot loop reproduces the reference drivers' counters."""
    for c in golden("driver_counts.json")["counts"]:
        rnd = random.Random(c["seed"])
        failed, not0 = so.run_subset(oracle, c["protocol"], c["runs"], c["model"], c["eset"], c["locations"],
                                     c["prob_lists"], cooling=c["cooling"], rnd=rnd,
                                     meas_rng=random.Random(c["meas_seed"]))
        assert failed == c["failed"], c
        if c["not0"] is not None:
            assert not0 == c["not0"], c


def test_angle_reduction_and_matrices():
    assert sv.reduce_angle(-np.pi / 2) == round(3.5 * np.pi, 12)
    assert sv.reduce_angle(4 * np.pi) == 0.0
    m = sv.mat_rxx(np.pi / 2)
    assert np.allclose(m @ m.conj().T, np.eye(4))
    assert np.allclose(m[0, 3], -1j / np.sqrt(2))
    st = sv.StateVector(3)
    st.apply_1q(sv.MAT_H, 1)
    assert abs(st.prob_one(1) - 0.5) < 1e-15
    assert st.collapse(1, 1) == pytest.approx(0.5)
    assert abs(st.psi[2] - 1.0) < 1e-15


def test_placement_routines():
    rnd = random.Random(3)
    for k in range(0, 5):
        m = so.generate_error_locations(k, 12, rnd)
        assert len(m) == 12 and sum(m) == k
    w = [0.0] * 10
    w[7] = 1.0
    assert so.generate_error_locations_non_uniform(1, 10, w, rnd)[7] == 1
    with pytest.raises(ValueError):
        so.generate_error_locations(5, 4, rnd)


def test_oracle_sampling_agrees_with_reference_statistics(oracle, golden):
    """The oracle's own sampled s2 counters against 4000 shots of the unmodified reference driver on the projectq
    stand-in (oracle/make_golden_stats.py): within 3.5 binomial sigma."""
    import random
    import numpy as np
    from oracle import s17_oracle as so
    st = golden("stats_s2.json")["stats"][0]
    m = 200
    failed, not0 = so.run_subset(oracle, "s2", m, st["model"], st["eset"], st["locations"], st["prob_lists"],
                                 rnd=random.Random(3), meas_rng=random.Random(4))
    p_ref, p = st["not0"] / st["runs"], not0 / m
    assert abs(p - p_ref) < 3.5 * np.sqrt(p_ref * (1 - p_ref) * (1 / st["runs"] + 1 / m))
    f_ref, f = st["failed"] / st["not0"], failed / max(1, not0)
    assert abs(f - f_ref) < 3.5 * np.sqrt(f_ref * (1 - f_ref) * (1 / st["not0"] + 1 / max(1, not0)))


def test_run_til_fail_matches_reference_chains(oracle, golden):
    """`oracle.run_til_fail` (CL:33-148 restated) with the seeds of oracle/make_golden_rtf.py reproduces the UNMODIFIED
    reference driver's rounds_til_fail_list and syndrome_b_measured_list run by run -- including the reference's
    never-reset leak register (CL:49) for Dress_model.  Two chains per model here (the CPU suite stays short); the GPU
    statistics test uses all of them."""
    for case in golden("rtf_ref.json")["cases"]:
        names = oracle.error_models[case["model"]]["probabilities"]
        e_probs = dict(zip(names, case["p_set"]))
        for chain in case["chains"][:2]:
            n = 6
            rounds, syndb = so.run_til_fail(oracle, n, case["model"], e_probs, rnd=random.Random(chain["seed"]),
                                            meas_rng=random.Random(chain["meas_seed"]), leak_reset="never")
            assert rounds == chain["rounds_til_fail_list"][:n], (case["model"], chain["seed"])
            assert syndb == chain["syndrome_b_measured_list"][:n], (case["model"], chain["seed"])
