"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Bars: integer / bit fields (syndromes, corrections, read-outs, pass/fail, leak flags) bit-exact;
amplitudes within 1e-10 per amplitude (BASELINE.json north_star).
"""
import os
import random

import numpy as np
import pytest

from oracle import s17_oracle as so
from oracle import sv
from surface_17_iqt_b200 import _lib as L
from surface_17_iqt_b200 import api, circuit

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AMP_TOL = 1e-10
FIELDS = ("quiescent", "synd1", "synd2", "flips_a", "ft_syndrome", "correction", "data_meas", "logical_meas",
          "fail", "not0", "counted", "leaked")


def flat(p_set):
    return [b for lst in p_set for b in lst]


def make_sim(models, ct, bt, model, list_len, n, fusion=1, cooling=False, **kw):
    return api.Simulation(model, list_len, max_shots=n, fusion=fusion, cooling=cooling, correction_table=ct,
                          bitflip_table=bt, error_models=models, **kw)


@pytest.mark.parametrize("fusion,dense,early", [(0, "0", False), (1, "0", False), (2, "0", False),
                                                (2, "1", False), (0, "1", False),
                                                (0, "0", True), (1, "0", True), (2, "0", True),
                                                (2, "1", True)])
def test_golden_replays(error_models, correction_table, bitflip_table, golden, fusion, dense, early, monkeypatch):
    """Replayed reference trajectories (masks + forced outcomes): every record field bit-exact, and the
    post-correction state equals the reference's 512-amplitude block within 1e-10."""
    monkeypatch.setenv("S17_DENSE", dense)     # 1: full-state sweeps + explicit collapse kernels
    reps = golden("replays.json")["replays"]
    blocks = np.load(os.path.join(GOLDEN, "replay_blocks.npz"))["blocks"]
    groups = {}
    for r in reps:
        groups.setdefault((r["model"], r["protocol"], r["cooling"]), []).append(r)
    for (model, protocol, cooling), rs in groups.items():
        list_len = [len(x) for x in rs[0]["p_set"]]
        sim = make_sim(error_models, correction_table, bitflip_table, model, list_len, len(rs), fusion, cooling,
                       early_measure=early)
        sim.backend.set_replay([flat(r["p_set"]) for r in rs], [r["outcomes"] for r in rs])
        c = sim.backend.run(protocol, len(rs), L.NOISE_REPLAY, forced=True, materialize=True)
        assert c.errors == 0
        recs = sim.backend.records(len(rs))
        for i, r in enumerate(rs):
            for f in FIELDS:
                assert recs[i].get(f) == r.get(f), (model, protocol, i, f, recs[i].get(f), r.get(f))
            if r["block_index"] >= 0:
                psi = sim.backend.state(i)
                err = np.max(np.abs(psi[:512] - blocks[r["block_index"]]))
                if early:   # an idle Z on an already-measured X ancilla is a global sign of the branch
                    err = min(err, np.max(np.abs(psi[:512] + blocks[r["block_index"]])))
                assert err < AMP_TOL
                assert np.sum(np.abs(psi[512:]) ** 2) < 1e-20
        assert c.failed == sum(r["fail"] for r in rs)
        assert c.not0 == sum(r["not0"] for r in rs)
        sim.close()


@pytest.mark.parametrize("dense,fusion,early,diag", [("0", 2, True, "1"), ("0", 2, True, "0"), ("1", 2, True, "1"),
                                                     ("0", 1, True, "1"), ("0", 2, False, "1"),
                                                     ("0", 0, True, "1")])
def test_golden_replays_fast_path(error_models, correction_table, bitflip_table, golden, dense, fusion, early, diag,
                                  monkeypatch):
    """Same replays through the production path (lazy collapse, correction folded into the read-out).
    diag=1: whole-cycle diagonal-frame kernel (k_cycle1, host-chosen label maps) where the cycle allows it; diag=g: its
    generic form k_cycle; diag=0: k_half everywhere."""
    monkeypatch.setenv("S17_DENSE", dense)
    monkeypatch.setenv("S17_DIAG", diag)
    reps = golden("replays.json")["replays"]
    groups = {}
    for r in reps:
        groups.setdefault((r["model"], r["protocol"], r["cooling"]), []).append(r)
    for (model, protocol, cooling), rs in groups.items():
        list_len = [len(x) for x in rs[0]["p_set"]]
        sim = make_sim(error_models, correction_table, bitflip_table, model, list_len, len(rs), fusion, cooling,
                       early_measure=early)
        sim.backend.set_replay([flat(r["p_set"]) for r in rs], [r["outcomes"] for r in rs])
        c = sim.backend.run(protocol, len(rs), L.NOISE_REPLAY, forced=True)
        assert c.errors == 0
        recs = sim.backend.records(len(rs))
        for i, r in enumerate(rs):
            for f in FIELDS:
                assert recs[i].get(f) == r.get(f), (model, protocol, i, f, recs[i].get(f), r.get(f))
        sim.close()


def test_amplitudes_after_every_gate(error_models, correction_table, bitflip_table, golden):
    """fusion=0 is ProjectQ's execution model (one sweep per gate): compare the full 2^17 state with the
    oracle after each of the 54 gates of a noiseless cycle."""
    trace = golden("cycle_trace.json")["trace"]
    outcomes = [0, 1, 0, 0, 1, 0, 1, 0]
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", [12, 18, 24, 78, 24], 1, fusion=0,
                   early_measure=False)
    zero = [[0] * n for n in (12, 18, 24, 78, 24)]
    sim.backend.set_replay([flat(zero)], [outcomes + [0] * 8])
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    st = sv.StateVector(17)
    gates = error_models["PDD_model"]["gates"]
    probs = dict(zip(error_models["PDD_model"]["probabilities"], zero))
    q = orc.logical_prep(st, "Z", 0, 17 * [0], gates, probs, random.Random(0), so.ForcedMeasure(outcomes))
    assert list(q) == outcomes
    sim.backend.run("s1", 1, L.NOISE_REPLAY, forced=True, stop_stage=L.STOP_AFTER_PREP)
    assert np.max(np.abs(sim.backend.state(0) - st.psi)) < AMP_TOL
    for k, (name, angle, bits) in enumerate(trace):
        orc._gate(st, name, bits, angle)
        sim.backend.run("s1", 1, L.NOISE_REPLAY, forced=True, stop_stage=L.STOP_CYCLE1_GATES, stop_layers=k + 1)
        err = np.max(np.abs(sim.backend.state(0) - st.psi))
        assert err < AMP_TOL, (k, name, bits, err)
    sim.close()


@pytest.mark.parametrize("fusion", [1, 2])
def test_amplitudes_noisy_cycle_before_measurement(error_models, correction_table, bitflip_table, fusion):
    """Full state after the gates of a noisy cycle (Pauli events of every kind, leak skips, idle Z)."""
    rng = random.Random(99)
    for model, cooling in (("PDD_model", False), ("Dress_cooling", True), ("depolarising_model", False)):
        list_len = circuit.list_lengths(error_models[model])
        n = 6
        p_sets, scripts = [], []
        for _ in range(n):
            p_set = [[0] * L_ for L_ in list_len]
            for t, L_ in enumerate(list_len):
                for i in rng.sample(range(L_), min(L_, 3)):
                    p_set[t][i] = 1
            p_sets.append(p_set)
        outcomes = [[rng.randint(0, 1) if j in (1, 3, 4, 6) else 0 for j in range(8)] for _ in range(n)]
        sim = make_sim(error_models, correction_table, bitflip_table, model, list_len, n, fusion, cooling,
                       early_measure=False)
        sim.backend.set_replay([flat(p) for p in p_sets], outcomes)
        sim.backend.run("single", n, L.NOISE_REPLAY, forced=True, stop_stage=L.STOP_CYCLE1_GATES, seed=3)
        ev = sim.backend.events(n)
        orc = so.ShotOracle(error_models, correction_table, bitflip_table)
        gates = error_models[model]["gates"]
        for i in range(n):
            if model == "depolarising_model":
                # the 15-way choice is drawn on device: read it back from the event word of each Rxx
                script = {}
                for k in range(24):
                    f = (int(ev[i, k]) >> 6) & 63
                    code = {0: 0, 1: 1, 3: 2, 2: 3}   # (x | z<<1) -> I X Y Z
                    p0, p1 = code[f & 3], code[(f >> 2) & 3]
                    if p0 or p1:
                        script[(0, k)] = p0 * 4 + p1 - 1
                orc.depol_script = script
            st = sv.StateVector(17)
            probs = dict(zip(error_models[model]["probabilities"], p_sets[i]))
            leaked = 17 * [0]
            orc.logical_prep(st, "Z", 0, leaked, gates, probs, random.Random(0), so.ForcedMeasure(outcomes[i]))
            orc.stabilizer_cycle(st, leaked, gates, probs, random.Random(0), None, cooling=cooling, gates_only=True)
            err = np.max(np.abs(sim.backend.state(i) - st.psi))
            assert err < AMP_TOL, (model, i, err)
        sim.close()


@pytest.mark.parametrize("model,protocol,cooling,eset", [
    ("PDD_model", "s1", False, [0, 0, 1, 1, 0]),
    ("PDD_model", "s2", False, [0, 1, 2, 1, 0]),
    ("Dress_model", "s2", False, [1, 0, 1, 3, 1]),
    ("Dress_cooling", "s2", True, [0, 0, 1, 2, 0, 3]),
    ("PDD_cooling", "single", True, [1, 1, 1, 1, 1, 2]),
])
@pytest.mark.parametrize("fusion,diag", [(2, "1"), (2, "0"), (1, "1")])
def test_sampled_trajectories_replay_bit_exact_in_oracle(error_models, correction_table, bitflip_table, model,
                                                         protocol, cooling, eset, fusion, diag, monkeypatch):
    """Device-sampled shots (Philox placement + outcomes): replaying their masks and outcomes through the
    oracle reproduces syndromes, corrections, read-outs and pass/fail bit for bit."""
    monkeypatch.setenv("S17_DIAG", diag)
    per_round = circuit.list_lengths(error_models[model])
    locs = [2 * x for x in per_round] if protocol == "s2" else per_round
    n = 32
    sim = make_sim(error_models, correction_table, bitflip_table, model, locs, n, fusion, cooling)
    sim.backend.set_subset(eset, locs)
    c = sim.backend.run(protocol, n, L.NOISE_SUBSET, seed=1234, first_shot_id=77)
    recs, masks = sim.backend.records(n), sim.backend.masks(n)
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    fails = not0 = 0
    for i in range(n):
        p_set = sim.backend.split_mask(masks[i, 0])
        assert [sum(x) for x in p_set] == eset
        ref = orc.run_shot(model, p_set, protocol, cooling=cooling, rnd=random.Random(0),
                           meas=so.ForcedMeasure(recs[i]["outcomes"]))
        for f in FIELDS:
            assert recs[i].get(f) == ref.get(f), (i, f, recs[i].get(f), ref.get(f))
        fails += ref["fail"]
        not0 += ref["not0"]
    assert (c.failed, c.not0, c.errors) == (fails, not0, 0)
    sim.close()


@pytest.mark.parametrize("diag", ["1"])
@pytest.mark.parametrize("model,eset", [("PDD_model", [2, 2, 3, 6, 2]), ("Dress_model", [1, 1, 2, 8, 1]),
                                        ("PDD_model", [0, 0, 0, 0, 0]), ("PDD_cooling", [1, 1, 2, 4, 1, 12]),
                                        ("Dress_cooling", [1, 1, 1, 6, 1, 20])])
def test_cycle_kernel_matches_half_kernel(error_models, correction_table, bitflip_table, model, eset, diag, monkeypatch):
    """The diagonal-frame cycle kernels (k_cycle1 / k_cycle) against the general amplitude kernel (k_half): the
    trajectories one samples, replayed through the other, give the same records and the same stored amplitudes
    (<= 1e-10, phase included) for every shot."""
    per_round = circuit.list_lengths(error_models[model])
    locs = [2 * x for x in per_round]
    n = 96
    cooling = model.endswith("cooling")
    monkeypatch.setenv("S17_DIAG", diag)
    sim = make_sim(error_models, correction_table, bitflip_table, model, locs, n, 2, cooling)
    sim.backend.set_subset(eset, locs)
    c1 = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=321, first_shot_id=5)
    recs1, masks = sim.backend.records(n), sim.backend.masks(n)
    states1 = [sim.backend.state(i)[:512].copy() for i in range(n)]
    sim.close()
    monkeypatch.setenv("S17_DIAG", "0")
    sim = make_sim(error_models, correction_table, bitflip_table, model, locs, n, 2, cooling)
    sim.backend.set_replay([list(masks[i, 0]) for i in range(n)], [r["outcomes"] for r in recs1])
    c2 = sim.backend.run("s2", n, L.NOISE_REPLAY, forced=True)
    recs2 = sim.backend.records(n)
    assert (c1.failed, c1.not0, c1.counted, c1.errors) == (c2.failed, c2.not0, c2.counted, 0)
    assert c1.sweeps != c2.sweeps                       # really two different kernels (one vs two sweeps per cycle)
    for i in range(n):
        for f in FIELDS + ("outcomes",):
            assert recs1[i].get(f) == recs2[i].get(f), (i, f, recs1[i].get(f), recs2[i].get(f))
        st2 = sim.backend.state(i)[:512]
        assert abs(np.linalg.norm(states1[i]) - 1.0) < 1e-12
        assert np.max(np.abs(states1[i] - st2)) < AMP_TOL, (i, np.max(np.abs(states1[i] - st2)))
    sim.close()


def test_true_probability_mode_replays_in_oracle(error_models, correction_table, bitflip_table):
    """Bernoulli-per-slot noise (run-til-fail style probabilities, high p so that many slots fire)."""
    model = "PDD_model"
    locs = circuit.list_lengths(error_models[model])
    n = 40
    sim = make_sim(error_models, correction_table, bitflip_table, model, locs, n, 1)
    sim.backend.set_slot_probs([[0.03] * x for x in locs])
    c = sim.backend.run("single", n, L.NOISE_PROB, seed=42)
    recs, masks = sim.backend.records(n), sim.backend.masks(n)
    assert masks[:, 0].sum() > n          # plenty of events fired
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    for i in range(n):
        ref = orc.run_shot(model, sim.backend.split_mask(masks[i, 0]), "single", rnd=random.Random(0),
                           meas=so.ForcedMeasure(recs[i]["outcomes"]))
        for f in FIELDS:
            assert recs[i].get(f) == ref.get(f), (i, f)
    assert c.failed == sum(r["fail"] for r in recs)
    sim.close()


def test_single_fault_table(error_models, correction_table, bitflip_table, golden):
    rows = golden("single_fault_pdd.json")["rows"]
    locs = [24, 36, 48, 156, 48]
    masks = []
    for row in rows:
        p_set = [[0] * x for x in locs]
        p_set[row["type"]][row["slot"]] = 1
        masks.append(flat(p_set))
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, len(rows), 2)
    sim.backend.set_replay(masks, [r["outcomes"] for r in rows])
    c = sim.backend.run("s2", len(rows), L.NOISE_REPLAY, forced=True)
    recs = sim.backend.records(len(rows))
    for rec, row in zip(recs, rows):
        assert rec["flips_a"] == row["flips_a"]
        assert rec.get("ft_syndrome") == row["ft"]
    assert c.failed == 0 and c.not0 == 156 - 58 and c.errors == 0
    sim.close()


def test_full_size_properties(error_models, correction_table, bitflip_table):
    """Size-independent properties at a large batch: no noise => no detection, no failure; the counters do
    not depend on how the shot range is split into batches (Philox counter = global shot id); every
    single-error subset is corrected (code distance 3)."""
    locs = [12, 18, 24, 78, 24]
    n = 2048
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n, 2, early_measure=False)
    sim.backend.set_subset([0, 0, 0, 0, 0], locs)
    c = sim.backend.run("s1", n, L.NOISE_SUBSET, seed=9)
    assert (c.failed, c.not0, c.counted, c.errors) == (0, 0, n, 0)
    assert c.sweeps == 2 * 4 * n
    quiescent = {tuple(r["quiescent"]) for r in sim.backend.records(n)}
    assert len(quiescent) == 16 and all(q[0] == q[2] == q[5] == q[7] == 0 for q in quiescent)

    sim.backend.set_subset([0, 0, 2, 1, 0], locs)
    whole = sim.backend.run("s1", n, L.NOISE_SUBSET, seed=11, first_shot_id=0)
    whole = (whole.failed, whole.not0, whole.counted)
    parts = [0, 0, 0]
    for first in range(0, n, n // 4):
        c = sim.backend.run("s1", n // 4, L.NOISE_SUBSET, seed=11, first_shot_id=first)
        parts = [parts[0] + c.failed, parts[1] + c.not0, parts[2] + c.counted]
    assert tuple(parts) == whole
    for t in range(5):
        eset = [0] * 5
        eset[t] = 1
        sim.backend.set_subset(eset, locs)
        assert sim.backend.run("s1", n, L.NOISE_SUBSET, seed=t).failed == 0
    sim.close()
    locs2 = [2 * x for x in locs]
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs2, n, 2)
    for t in range(5):
        eset = [0] * 5
        eset[t] = 1
        sim.backend.set_subset(eset, locs2)
        assert sim.backend.run("s2", n, L.NOISE_SUBSET, seed=t).failed == 0
    sim.close()


@pytest.mark.parametrize("model,protocol,cooling,eset", [("PDD_model", "s2", False, [1, 1, 2, 2, 1]),
                                                         ("Dress_cooling", "s1", True, [0, 1, 1, 3, 0, 4])])
def test_memoised_preparation_is_bit_identical(error_models, correction_table, bitflip_table, model, protocol, cooling,
                                               eset, monkeypatch):
    """k_prep replays the decisions k_cycle1 takes in the noiseless preparation cycle (same uniforms, same probability
    pairs) and hands the first noisy cycle the cached block: records, outcome logs and stored amplitudes are identical
    to running the preparation cycle per shot, sampled and replayed."""
    per_round = circuit.list_lengths(error_models[model])
    locs = [2 * x for x in per_round] if protocol == "s2" else per_round
    n = 4096
    out = {}
    for cache in ("1", "0"):
        monkeypatch.setenv("S17_PREP_CACHE", cache)
        sim = make_sim(error_models, correction_table, bitflip_table, model, locs, n, 2, cooling)
        sim.backend.set_subset(eset, locs)
        c = sim.backend.run(protocol, n, L.NOISE_SUBSET, seed=99, first_shot_id=12345)
        recs = sim.backend.records(n)
        states = np.stack([sim.backend.state(i)[:512] for i in range(0, n, 64)])
        # forced replay of the first 64 trajectories through the same configuration
        sim.backend.set_replay([list(m) for m in sim.backend.masks(64)[:, 0]], [r["outcomes"] for r in recs[:64]])
        c2 = sim.backend.run(protocol, 64, L.NOISE_REPLAY, forced=True)
        assert c2.errors == 0
        out[cache] = ((c.failed, c.not0, c.counted, c.errors, c.sweeps), recs, states, sim.backend.records(64))
        sim.close()
    assert out["1"][0] == out["0"][0]
    assert out["1"][1] == out["0"][1]
    assert np.array_equal(out["1"][2], out["0"][2])
    for i in range(64):
        for f in FIELDS + ("outcomes",):
            assert out["1"][3][i].get(f) == out["1"][1][i].get(f) == out["0"][3][i].get(f), (i, f)


def test_cycle_kernel_large_batch_properties(error_models, correction_table, bitflip_table, monkeypatch):
    """The whole-cycle kernel at a batch where every warp walks a few dozen shots (prefetch ring, barrier phases): no
    noise => no detection and no failure; every single-error subset is corrected (distance 3); the whole-cycle kernel
    and the fused half-cycle kernel (S17_DIAG=0) tally the same sampled batch alike; the tallies do not depend on how the
    shot range is split."""
    locs = [24, 36, 48, 156, 48]
    n = 65536
    monkeypatch.setenv("S17_DIAG", "1")
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n, 2)
    sim.backend.set_subset([0, 0, 0, 0, 0], locs)
    c = sim.backend.run("s1", n, L.NOISE_SUBSET, seed=3)
    assert (c.failed, c.not0, c.counted, c.errors) == (0, 0, n, 0)
    assert c.sweeps == 2 * n                                       # one fused sweep per shot and cycle (prep + cycle 1)
    for t in range(5):
        eset = [0] * 5
        eset[t] = 1
        sim.backend.set_subset(eset, locs)
        assert sim.backend.run("s2", n, L.NOISE_SUBSET, seed=t).failed == 0
    sim.backend.set_subset([0, 0, 2, 1, 1], locs)
    whole = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=17, first_shot_id=1000)
    whole = (whole.failed, whole.not0, whole.counted)
    parts = [0, 0, 0]
    for first in range(0, n, n // 4):
        c = sim.backend.run("s2", n // 4, L.NOISE_SUBSET, seed=17, first_shot_id=1000 + first)
        parts = [parts[0] + c.failed, parts[1] + c.not0, parts[2] + c.counted]
    assert tuple(parts) == whole
    sim.close()
    monkeypatch.setenv("S17_DIAG", "0")
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n, 2)
    sim.backend.set_subset([0, 0, 2, 1, 1], locs)
    c = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=17, first_shot_id=1000)
    # same placements and the same uniforms for the ancilla Measures: the not0 tally is identical; the data read-out of
    # the k_half path is drawn by k_readout in a different order, so failures agree statistically
    assert c.not0 == whole[1] and c.counted == whole[2]
    p = (c.failed + whole[0]) / (2 * whole[2])
    assert abs(c.failed - whole[0]) < 5 * np.sqrt(2 * whole[2] * p * (1 - p))
    sim.close()


def test_statistics_against_oracle_sampling(error_models, correction_table, bitflip_table):
    """Sampled s2 failure / not0 rates agree with the oracle's own sampling within binomial 3 sigma."""
    locs = [24, 36, 48, 156, 48]
    eset = [0, 0, 2, 0, 0]
    n = 4096
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n, 2)
    sim.backend.set_subset(eset, locs)
    c = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=5)
    sim.close()
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    m = 150
    failed, not0 = so.run_subset(orc, "s2", m, "PDD_model", eset, locs, [[1]] * 5, rnd=random.Random(8),
                                 meas_rng=random.Random(9))
    p_gpu, p_cpu = c.not0 / n, not0 / m
    p = (c.not0 + not0) / (n + m)
    assert abs(p_gpu - p_cpu) < 3 * np.sqrt(p * (1 - p) * (1 / n + 1 / m)) + 1e-9
    f_gpu, f_cpu = c.failed / max(1, c.not0), failed / max(1, not0)
    f = (c.failed + failed) / max(1, c.not0 + not0)
    assert abs(f_gpu - f_cpu) < 3 * np.sqrt(f * (1 - f) * (1 / max(1, c.not0) + 1 / max(1, not0))) + 1e-9


@pytest.mark.parametrize("fusion", [1, 2])
def test_coherent_overrotation_against_oracle(error_models, correction_table, bitflip_table, fusion):
    """BASELINE config 3 (extension): deterministic over-rotations + stochastic XX, full-state parity."""
    eps1, eps2 = [0.02] * 30, [0.02] * 24
    n = 4
    sim = make_sim(error_models, correction_table, bitflip_table, "coherent_model", [30, 24, 24], n, fusion,
                   over_angles={"eps_1q": eps1, "eps_2q": eps2}, early_measure=False)
    masks = []
    rng = random.Random(4)
    for _ in range(n):
        xx = [0] * 24
        xx[rng.randrange(24)] = 1
        masks.append([0] * 54 + xx)
    outcomes = [[0, rng.randint(0, 1), 0, rng.randint(0, 1), rng.randint(0, 1), 0, rng.randint(0, 1), 0] for _ in range(n)]
    sim.backend.set_replay(masks, outcomes)
    sim.backend.run("single", n, L.NOISE_REPLAY, forced=True, stop_stage=L.STOP_CYCLE1_GATES)
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    gates = error_models["coherent_model"]["gates"]
    for i in range(n):
        st = sv.StateVector(17)
        probs = {"eps_1q": eps1, "eps_2q": eps2, "p_XX_ctrl": masks[i][54:]}
        noiseless = {"eps_1q": [0.0] * 30, "eps_2q": [0.0] * 24, "p_XX_ctrl": [0] * 24}
        orc.stabilizer_cycle(st, 17 * [0], gates, noiseless, random.Random(0), so.ForcedMeasure(outcomes[i]))
        orc.stabilizer_cycle(st, 17 * [0], gates, probs, random.Random(0), None, gates_only=True)
        assert np.max(np.abs(sim.backend.state(i) - st.psi)) < AMP_TOL
    sim.close()


def test_run_til_fail_bookkeeping(error_models, correction_table, bitflip_table):
    """Run-til-fail (CL:33-148): every run ends with a failure, rounds >= 1, totals consistent."""
    locs = [12, 18, 24, 78, 24]
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, 64, 2)
    p = 0.02
    sim.backend.set_slot_probs([[p] * 12, [p] * 18, [p] * 24, [0.0] * 78, [0.0] * 24])
    runs = 200
    c = sim.backend.run("rtf", 64, L.NOISE_PROB, seed=3, rtf_runs=runs)
    rounds, syndb = sim.backend.rtf_results(runs)
    assert (rounds >= 1).all() and int(rounds.sum()) == c.rounds
    assert c.failed == runs
    assert (syndb <= rounds).all()
    mean = rounds.mean()
    assert 2.0 < mean < 60.0
    sim.close()


def test_dense_ring_kernel_stress(error_models, correction_table, bitflip_table, monkeypatch):
    """Regression for the TMA-ring sweep (k_layer2) in its HBM-bound regime (dense, one timestep per sweep): with
    two math groups sharing three buffers a group used to wait on an mbarrier phase it had not followed and could
    start on a buffer that was still being filled (launch failures / hangs at thousands of shots).  Dense and
    support-aware schedules must also agree exactly on every counter."""
    locs = [24, 36, 48, 156, 48]
    n = 4096
    counts = {}
    for dense in ("1", "0"):
        monkeypatch.setenv("S17_DENSE", dense)
        sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n, 1, early_measure=False)
        sim.backend.set_subset([0, 0, 2, 1, 1], locs)
        res = []
        for k in range(3):
            c = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=77 + k)
            assert c.errors == 0
            res.append((c.failed, c.not0, c.counted))
        counts[dense] = res
        sim.close()
    assert counts["1"] == counts["0"]


def test_depolarising_true_probability_statistics(error_models, correction_table, bitflip_table):
    """BASELINE config 1 (depolarising_model, Bernoulli per slot, single round) at p = 0.02: failure rate and detection
    rate of 2^17 device shots (geometric skipping, memoised event-free cycles) within 3.5 binomial sigma of 480 oracle
    shots."""
    import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as drv
    p = 0.02
    p_set = [30 * [p / 3], 30 * [p / 3], 30 * [p / 3], 24 * [p]]
    drv.VERBOSE = False
    drv.BATCH_SHOTS = 1 << 17
    n = 1 << 17
    rate, failed, counted, not0 = drv.calculate_log_e_rate_true_probability(n, correction_table, "depolarising_model",
                                                                           p_set, bitflip_table)
    drv.release()
    assert counted == n
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    rnd, mr = random.Random(5), random.Random(6)
    m, ofail, onot0 = 480, 0, 0
    for _ in range(m):
        r = orc.run_shot("depolarising_model", p_set, "single", rnd=rnd, meas=so.SampledMeasure(mr))
        ofail += r["fail"]
        onot0 += r["not0"]
    # failure rate and detection rate (syndrome changed in the noisy cycle): 480 oracle shots resolve a factor-two error
    # in either (sigma of the detection rate ~ 2.3 %, of the failure rate ~ 1.2 %)
    for got, ref in ((rate, ofail / m), (not0 / n, onot0 / m)):
        assert abs(got - ref) < 3.5 * np.sqrt(got * (1 - got) / m) + 1e-9, (got, ref)
    assert 0.02 < rate < 0.6 and 0.2 < not0 / n < 0.95


def test_run_til_fail_sweep_example():
    import importlib.util
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("vary_p", os.path.join(here, "examples", "vary_p.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.sweep("Dress_model", runs=300, ps=(0.02, 0.01))
    assert out["0.02"]["log_e_rate_fit"] > 0 and out["0.01"]["log_e_rate_fit"] > 0
    assert out["0.02"]["rounds"] < out["0.01"]["rounds"] * 1.5      # more noise -> runs fail sooner


def test_driver_batch_shrinks_when_full_states_do_not_fit(correction_table, bitflip_table, monkeypatch):
    """The drop-in drivers default to a batch sized for the whole-cycle kernel (8 KiB per shot).  A program that needs
    full 2 MiB states (here: coherent over-rotations kept on the sweep kernels, S17_COH=0) makes the first allocation
    fail; the driver shrinks the batch and carries on instead of surfacing the error."""
    import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as drv
    monkeypatch.setenv("S17_COH", "0")
    monkeypatch.setattr(drv, "BATCH_SHOTS", 262144)
    monkeypatch.setattr(drv, "VERBOSE", False)
    drv.release()
    p_set = [[0.02] * 30, [0.02] * 24, [5e-3] * 24]
    rate, failed, counted, not0 = drv.calculate_log_e_rate_true_probability(
        2000, correction_table, "coherent_model", p_set, bitflip_table,
        over_angles={"eps_1q": [0.02] * 30, "eps_2q": [0.02] * 24})
    assert counted == 2000 and 0 <= failed <= counted
    assert drv.BATCH_SHOTS == 262144                         # the shrink is per program, not global
    assert [v for k, v in drv._BATCH_OF.items() if k[0] == "coherent_model"] == [262144 // 16]
    drv.release()
    drv._BATCH_OF.clear()


@pytest.mark.parametrize("diag", ["1", "0"])
def test_logical_one_preparation(error_models, correction_table, bitflip_table, diag, monkeypatch):
    """state = 1 (All(X) | data before the preparation cycle, CS:103-104; dead in the reference's drivers, CL:274-281,
    but part of logical_prep): sampled trajectories replay bit-exactly in the oracle, and a context alternates between
    the two logical states (the memoised preparation cycle is rebuilt per basis state)."""
    monkeypatch.setenv("S17_DIAG", diag)
    model, locs, eset, n = "PDD_model", [24, 36, 48, 156, 48], [0, 1, 1, 2, 0], 48
    sim = make_sim(error_models, correction_table, bitflip_table, model, locs, n, 2)
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    for state in (1, 0, 1):
        sim.backend.set_subset([0] * 5, locs)
        c = sim.backend.run("s1", n, L.NOISE_SUBSET, seed=21 + state, logical_state=state)
        assert (c.failed, c.not0, c.counted, c.errors) == (0, 0, n, 0)
        assert all(r["logical_meas"] == state for r in sim.backend.records(n))
        sim.backend.set_subset(eset, locs)
        c = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=5 + state, first_shot_id=9, logical_state=state)
        recs, masks = sim.backend.records(n), sim.backend.masks(n)
        for i in range(n):
            ref = orc.run_shot(model, sim.backend.split_mask(masks[i, 0]), "s2", rnd=random.Random(0),
                               meas=so.ForcedMeasure(recs[i]["outcomes"]), state=state)
            for f in FIELDS:
                assert recs[i].get(f) == ref.get(f), (state, i, f, recs[i].get(f), ref.get(f))
        assert c.failed == sum(r["fail"] for r in recs)
    sim.close()


def test_compact_then_full_state_allocation(error_models, correction_table, bitflip_table):
    """A context that has only run the whole-cycle kernel keeps 8 KiB per shot; a later call that needs full states
    (materialised collapse + correction for a parity read-back) re-allocates and gives the same trajectories."""
    locs = [24, 36, 48, 156, 48]
    n = 64
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n, 2)
    sim.backend.set_subset([0, 1, 1, 2, 1], locs)
    c1 = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=8)
    recs1, masks = sim.backend.records(n), sim.backend.masks(n)
    compact = [sim.backend.state(i) for i in range(4)]
    assert all(np.sum(np.abs(v[512:]) ** 2) == 0.0 for v in compact)
    sim.backend.set_replay([list(masks[i, 0]) for i in range(n)], [r["outcomes"] for r in recs1])
    c2 = sim.backend.run("s2", n, L.NOISE_REPLAY, forced=True, materialize=True)
    recs2 = sim.backend.records(n)
    assert (c1.failed, c1.not0, c1.counted) == (c2.failed, c2.not0, c2.counted) and c2.errors == 0
    for i in range(n):
        for f in FIELDS + ("outcomes",):
            assert recs1[i].get(f) == recs2[i].get(f), (i, f)
    # the materialised state is the stored block with the decoder's correction applied
    for i in range(4):
        if not recs1[i]["counted"]:
            continue
        full = sim.backend.state(i)
        x = sum(b << k for k, b in enumerate(recs1[i]["correction"][:9]))
        z = sum(b << k for k, b in enumerate(recs1[i]["correction"][9:]))
        idx = np.arange(512)
        sign = (-1.0) ** np.array([bin(int(m) & z).count("1") for m in idx])
        expect = np.zeros(512, dtype=complex)
        expect[idx] = sign * compact[i][:512][idx ^ x]
        err = min(np.max(np.abs(full[:512] - expect)), np.max(np.abs(full[:512] + expect)))
        assert err < AMP_TOL and np.sum(np.abs(full[512:]) ** 2) < 1e-20
    sim.close()


def test_sampled_rates_against_reference_statistics(error_models, correction_table, bitflip_table, golden):
    """Sampled s1 / s2 rates against the UNMODIFIED reference drivers run on the projectq stand-in
    (tests/golden/stats_s2.json, oracle/make_golden_stats.py, 4000 shots per subset): PDD_model with uniform placement and
    with the linear-heating lists of importance_sampling_syndrome_processing.py:20-36 (non-uniform placement), Dress_model
    (leakage), Dress_cooling / PDD_cooling with cooling=True (ISC:19-34).  not0 / runs and failed / decoded within 3.5
    binomial sigma of the reference sample (the GPU sample of 2^20 shots adds < 2 % to the variance)."""
    n = 1 << 20
    stats = golden("stats_s2.json")["stats"]
    assert len(stats) >= 11 and {s_["model"] for s_ in stats} >= {"PDD_model", "Dress_model", "Dress_cooling", "PDD_cooling"}
    for st in stats:
        cooling = bool(st.get("cooling", False))
        sim = make_sim(error_models, correction_table, bitflip_table, st["model"], st["locations"], n, 2, cooling)
        sim.backend.set_subset(st["eset"], st["locations"], [None if len(p) == 1 else p for p in st["prob_lists"]])
        c = sim.backend.run(st["protocol"], n, L.NOISE_SUBSET, seed=20260101)
        sim.close()
        assert c.errors == 0
        p_ref, p_gpu = st["not0"] / st["runs"], c.not0 / n
        sig = np.sqrt(max(p_ref * (1 - p_ref), 1e-6) / st["runs"] + p_gpu * (1 - p_gpu) / n)
        assert abs(p_gpu - p_ref) < 3.5 * sig, (st["model"], st["eset"], "not0", p_gpu, p_ref, sig)
        # s2 decodes the shots whose first syndrome changed (CL:416), s1 the others (CL:321)
        d_ref = st["not0"] if st["protocol"] == "s2" else st["runs"] - st["not0"]
        d_gpu = c.not0 if st["protocol"] == "s2" else n - c.not0
        assert c.counted == d_gpu
        f_ref, f_gpu = st["failed"] / max(1, d_ref), c.failed / max(1, d_gpu)
        sig = np.sqrt(max(f_ref * (1 - f_ref), 1e-6) / max(1, d_ref) + f_gpu * (1 - f_gpu) / max(1, d_gpu))
        assert abs(f_gpu - f_ref) < 3.5 * sig, (st["model"], st["eset"], "failed", f_gpu, f_ref, sig)


@pytest.mark.parametrize("diag", ["1", "0"])
def test_impossible_forced_outcome_is_reported(error_models, correction_table, bitflip_table, diag, monkeypatch):
    """A replay that forces an outcome of probability zero (a z-type ancilla reading 1 in the noiseless preparation
    cycle; a flipped x-type ancilla in a noiseless second cycle) is counted in `errors` and flagged in the record on
    every path (memoised preparation, whole-cycle kernels, half-cycle kernel) instead of yielding a garbage state."""
    monkeypatch.setenv("S17_DIAG", diag)
    locs = [12, 18, 24, 78, 24]
    n = 8
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n, 2)
    zero = [0] * sum(locs)
    good = [0, 1, 0, 0, 1, 0, 1, 0]
    outs = []
    for i in range(n):
        o = good + good + [0] * 9
        if i % 4 == 1:
            o[0] = 1                      # z-type ancilla 0 cannot read 1 after a noiseless preparation
        if i % 4 == 2:
            o[8 + 1] = 0                  # x-type ancilla 1 must repeat its quiescent outcome
        outs.append(o)
    sim.backend.set_replay([zero] * n, outs)
    c = sim.backend.run("s1", n, L.NOISE_REPLAY, forced=True)
    recs = sim.backend.records(n)
    assert c.errors >= n // 2
    for i in range(n):
        assert bool(recs[i]["error"]) == (i % 4 in (1, 2)), (i, recs[i])
    sim.close()


@pytest.mark.parametrize("model,protocol,cooling,p", [("PDD_model", "s1", False, 2e-3), ("depolarising_model", "single", False, 1e-3),
                                                      ("PDD_cooling", "s2", True, 1e-3), ("Dress_model", "s1", False, 3e-3)])
def test_memoised_event_free_cycle_is_bit_identical(error_models, correction_table, bitflip_table, model, protocol, cooling,
                                                    p, monkeypatch):
    """True-probability noise at rates where most first cycles have no event: k_quiet replays those from the cache of
    k_cycle1's own results (outcomes = the preparation leaf's, the block and the read-out weights the kernel produced).
    Counters, records, outcome logs and stored blocks are identical to running every cycle (S17_QUIET=0), and the
    sampled trajectories replay bit-exactly in the oracle."""
    per_round = circuit.list_lengths(error_models[model])
    locs = [2 * x for x in per_round] if protocol == "s2" else per_round
    n = 8192
    out = {}
    for quiet in ("1", "0"):
        monkeypatch.setenv("S17_QUIET", quiet)
        sim = make_sim(error_models, correction_table, bitflip_table, model, locs, n, 2, cooling)
        sim.backend.set_slot_probs([[p] * x for x in locs])
        c = sim.backend.run(protocol, n, L.NOISE_PROB, seed=77, first_shot_id=3)
        recs, masks = sim.backend.records(n), sim.backend.masks(n)
        states = np.stack([sim.backend.state(i)[:512] for i in range(0, n, 128)])
        out[quiet] = ((c.failed, c.not0, c.counted, c.errors), recs, states, masks, c.sweeps)
        if quiet == "1" and model != "depolarising_model":          # (the 15-way depolarising choice is not in the masks)
            orc = so.ShotOracle(error_models, correction_table, bitflip_table)
            for i in range(48):
                ref = orc.run_shot(model, sim.backend.split_mask(masks[i, 0] | masks[i, 1]) if protocol == "s2" else
                                   sim.backend.split_mask(masks[i, 0]), protocol, cooling=cooling, rnd=random.Random(0),
                                   meas=so.ForcedMeasure(recs[i]["outcomes"]))
                for f in FIELDS:
                    assert recs[i].get(f) == ref.get(f), (i, f, recs[i].get(f), ref.get(f))
        sim.close()
    assert out["1"][0] == out["0"][0] and out["1"][4] == out["0"][4]
    assert out["1"][1] == out["0"][1]
    assert np.array_equal(out["1"][3], out["0"][3])
    assert np.array_equal(out["1"][2], out["0"][2])
    quiet_fraction = np.mean(out["1"][3][:, 0].sum(axis=1) == 0)
    assert quiet_fraction > 0.3                                     # the memoised path really carried a large share


@pytest.mark.parametrize("model,p_leak", [("Dress_model", 3e-3), ("PDD_model", 1e-3)])
def test_run_til_fail_memoisation_is_bit_identical(error_models, correction_table, bitflip_table, model, p_leak,
                                                   monkeypatch):
    """Run-til-fail with every memoisation (preparation leaves, event-free first cycles; with a leakage model only for
    the shots without a set leak flag, the others run their own preparation cycle) against running every cycle of
    every round: the same rounds-til-fail and second-syndrome counts.  (A freed slot takes the next run id by an atomic
    counter, so which run id labels which trajectory can differ between two executions; the trajectories themselves
    are keyed by slot and round, hence the comparison of the sorted lists.)"""
    locs = [12, 18, 24, 78, 24]
    runs, slots = 3000, 512
    res = {}
    for cache in ("1", "0"):
        monkeypatch.setenv("S17_PREP_CACHE", cache)
        sim = make_sim(error_models, correction_table, bitflip_table, model, locs, slots, 2)
        p = 0.01
        sim.backend.set_slot_probs([[p / 10] * 12, [p / 10] * 18, [p] * 24, [p_leak] * 78, [0.0] * 24])
        c = sim.backend.run("rtf", slots, L.NOISE_PROB, seed=11, rtf_runs=runs)
        rounds, syndb = sim.backend.rtf_results(runs)
        res[cache] = (rounds.copy(), syndb.copy(), c.rounds, c.failed)
        sim.close()
    assert res["1"][2:] == res["0"][2:]
    assert np.array_equal(res["1"][0][:slots], res["0"][0][:slots]) and np.array_equal(res["1"][1][:slots], res["0"][1][:slots])
    both = lambda r: sorted(zip(r[0].tolist(), r[1].tolist()))
    assert both(res["1"]) == both(res["0"])
