"""GPU parity tests, second set: run-til-fail against the oracle and the reference's chains, non-uniform placement
against its exact law, the production kernel's stored block against the reference's block.  All through the C ABI."""
import itertools
import os
import random

import numpy as np
import pytest

from oracle import s17_oracle as so
from surface_17_iqt_b200 import _lib as L
from surface_17_iqt_b200 import api, circuit

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AMP_TOL = 1e-10
RTF_FIELDS = ("quiescent", "synd1", "synd2", "flips_a", "ft_syndrome", "correction", "data_meas", "logical_meas", "fail",
              "not0", "leaked")


def make_sim(models, ct, bt, model, list_len, n, fusion=2, cooling=False, **kw):
    return api.Simulation(model, list_len, max_shots=n, fusion=fusion, cooling=cooling, correction_table=ct,
                          bitflip_table=bt, error_models=models, **kw)


def flat(p_set):
    return [b for lst in p_set for b in lst]


# ---------------------------------------------------------------------------------------------------------------------
# run-til-fail (CL:33-148)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("diag", ["1", "0"])
@pytest.mark.parametrize("model,p_leak,leak_reset", [("Dress_model", 2e-2, "never"), ("Dress_model", 2e-2, "run"),
                                                     ("Dress_model", 2e-2, "round"), ("PDD_model", 1e-2, "run")])
def test_run_til_fail_rounds_replay_in_oracle(error_models, correction_table, bitflip_table, model, p_leak, leak_reset, diag,
                                              monkeypatch):
    """Every round of every slot of a run-til-fail job, replayed through `oracle.run_round_rtf` (which is pinned to the
    reference's own chains, tests/test_oracle_golden.py): the slots that fired in cycle A and in cycle B (cycle B
    re-uses the stab_ind=0 slots, CL:92), the outcomes, the leak register carried from round to round and -- depending
    on leak_reset -- from run to run.  Syndromes, the protocol branch, correction, read-out, failure and the leak
    register after the round are bit-exact; the per-run round counts the library reports equal the rounds seen."""
    monkeypatch.setenv("S17_DIAG", diag)
    locs = [12, 18, 24, 78, 24]
    names = error_models[model]["probabilities"]
    slots, runs = 24, 60
    sim = make_sim(error_models, correction_table, bitflip_table, model, locs, slots)
    p = 0.03
    sim.backend.set_slot_probs([[p / 10] * 12, [p / 10] * 18, [p] * 24, [p_leak] * 78, [0.0] * 24])
    sim.backend.rtf_begin(slots, L.NOISE_PROB, seed=77, rtf_runs=runs, leak_reset=leak_reset)
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    leak = [17 * [0] for _ in range(slots)]
    seen_rounds, seen_b = {}, {}
    alive, n_rounds = slots, 0
    while alive:
        before = sim.backend.rtf_slots(slots)
        alive, total = sim.backend.rtf_step(1)
        recs, masks = sim.backend.records(slots), sim.backend.masks(slots)
        after = sim.backend.rtf_slots(slots)
        for s in range(slots):
            if not before[s, 3]:
                continue
            n_rounds += 1
            run_id = int(before[s, 0])
            if leak_reset == "round":
                leak[s] = 17 * [0]
            pa = dict(zip(names, sim.backend.split_mask(masks[s, 0])))
            pb = dict(zip(names, sim.backend.split_mask(masks[s, 1])))
            ref = orc.run_round_rtf(model, pa, leak[s], random.Random(0), so.ForcedMeasure(recs[s]["outcomes"]),
                                    e_probs_b=pb, want_record=True)
            for f in RTF_FIELDS:
                assert recs[s].get(f) == ref.get(f), (n_rounds, s, f, recs[s].get(f), ref.get(f))
            leak[s] = ref["leaked"]
            seen_rounds[run_id] = seen_rounds.get(run_id, 0) + 1
            seen_b[run_id] = seen_b.get(run_id, 0) + ref["not0"]
            if ref["fail"]:
                if leak_reset == "run":
                    leak[s] = 17 * [0]
                assert after[s, 0] != run_id or not after[s, 3]      # the slot moved on (or went idle)
            else:
                assert after[s, 0] == run_id and after[s, 3] == 1 and after[s, 1] == seen_rounds[run_id]
        assert total == n_rounds
    c = sim.backend.rtf_end()
    rounds, syndb = sim.backend.rtf_results(runs)
    assert c.failed == runs and c.rounds == n_rounds and c.errors == 0
    assert [seen_rounds[i] for i in range(runs)] == rounds.tolist()
    assert [seen_b[i] for i in range(runs)] == syndb.tolist()
    if model == "Dress_model" and leak_reset == "never":
        assert any(any(v[:9]) for v in leak)       # data leak flags did persist in some slot
    sim.close()


def test_run_til_fail_run_ids_are_reproducible(error_models, correction_table, bitflip_table):
    """A freed slot takes the next run id in slot order (a scan, not a race on a counter): two executions of the same
    job give the same rounds-til-fail list, element by element, and s17_run equals the stepped job."""
    locs = [12, 18, 24, 78, 24]
    out = []
    for mode in ("run", "run", "stepped"):
        sim = make_sim(error_models, correction_table, bitflip_table, "Dress_model", locs, 256)
        sim.backend.set_slot_probs([[2e-3] * 12, [2e-3] * 18, [2e-2] * 24, [2e-3] * 78, [0.0] * 24])
        if mode == "run":
            c = sim.backend.run("rtf", 256, L.NOISE_PROB, seed=5, rtf_runs=3000)
        else:
            sim.backend.rtf_begin(256, L.NOISE_PROB, seed=5, rtf_runs=3000)
            while sim.backend.rtf_step(3)[0]:
                pass
            c = sim.backend.rtf_end()
        r, b = sim.backend.rtf_results(3000)
        out.append((r.tolist(), b.tolist(), c.rounds, c.failed))
        sim.close()
    assert out[0] == out[1] == out[2]
    assert out[0][3] == 3000 and min(out[0][0]) >= 1


def test_run_til_fail_statistics_against_reference_chains(error_models, correction_table, bitflip_table, golden):
    """Rates of the run-til-fail job against the UNMODIFIED reference driver's chains (tests/golden/rtf_ref.json,
    oracle/make_golden_rtf.py: 8 calls of calculate_log_e_rate(16, ...) per model).

    PDD_model (no leakage): failures per round and second syndromes per round, binomial 3.5 sigma.
    Dress_model: the reference never clears its leak register (CL:49), so the runs of one call are one chain -- the
    library reproduces the protocol with leak_reset='never' and chains of exactly 16 runs per slot; the mean rounds per
    chain is compared with the reference's 8 chains (z-test with the GPU's own chain-to-chain variance)."""
    locs = [12, 18, 24, 78, 24]
    for case in golden("rtf_ref.json")["cases"]:
        ref_rounds = [c["rounds_til_fail_list"] for c in case["chains"]]
        ref_b = [c["syndrome_b_measured_list"] for c in case["chains"]]
        k = len(ref_rounds[0])
        slots = 4096
        sim = make_sim(error_models, correction_table, bitflip_table, case["model"], locs, slots)
        sim.backend.set_slot_probs(case["p_set"])
        c = sim.backend.run("rtf", slots, L.NOISE_PROB, seed=2027, rtf_runs=slots * k, rtf_chain_runs=k, leak_reset="never")
        r, b = sim.backend.rtf_results(slots * k)
        sim.close()
        assert c.errors == 0 and c.failed == slots * k and int(r.sum()) == c.rounds
        r, b = r.reshape(slots, k).astype(float), b.reshape(slots, k).astype(float)
        n_ref, n_gpu = sum(map(sum, ref_rounds)), r.sum()
        if case["p_leak"] == 0.0:
            for ref_events, gpu_events in ((len(ref_rounds) * k, slots * k), (sum(map(sum, ref_b)), b.sum())):
                p_ref, p_gpu = ref_events / n_ref, gpu_events / n_gpu
                sig = np.sqrt(p_gpu * (1 - p_gpu) * (1 / n_ref + 1 / n_gpu))
                assert abs(p_ref - p_gpu) < 3.5 * sig, (case["model"], p_ref, p_gpu, sig)
        chain_means = r.mean(axis=1)
        ref_mean = np.mean([np.mean(x) for x in ref_rounds])
        sig = chain_means.std() * np.sqrt(1 / len(ref_rounds) + 1 / slots)
        assert abs(ref_mean - chain_means.mean()) < 3.5 * sig, (case["model"], ref_mean, chain_means.mean(), sig)
        # the first run of a chain starts from a clean register, later runs inherit leaked data qubits: for the leakage
        # model the chains really are chains
        if case["p_leak"] > 0:
            assert r[:, 0].mean() > 1.3 * r[:, k - 1].mean()


def test_run_til_fail_caps_and_budget(error_models, correction_table, bitflip_table):
    """rtf_max_rounds closes a run at the cap; rtf_round_budget stops the job and leaves the runs in flight censored."""
    locs = [12, 18, 24, 78, 24]
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, 512)
    sim.backend.set_slot_probs([[1e-4] * 12, [1e-4] * 18, [1e-3] * 24, [0.0] * 78, [0.0] * 24])
    c = sim.backend.run("rtf", 512, L.NOISE_PROB, seed=1, rtf_runs=2048, rtf_max_rounds=5)
    r, _ = sim.backend.rtf_results(2048)
    assert r.max() == 5 and r.min() >= 1 and c.rounds == int(r.sum()) and c.failed < 2048
    c = sim.backend.run("rtf", 512, L.NOISE_PROB, seed=1, rtf_runs=10 ** 6, rtf_round_budget=20000, rtf_poll=8)
    r, _ = sim.backend.rtf_results(10 ** 6)
    assert 20000 <= c.rounds < 20000 + 8 * 512
    assert int(r.sum()) <= c.rounds and int((r > 0).sum()) == c.failed
    sim.close()


# ---------------------------------------------------------------------------------------------------------------------
# fixed-weight placement (CL:151-181)
# ---------------------------------------------------------------------------------------------------------------------
def sequential_rejection_marginals(w, k):
    """Exact P(slot i is hit) for generate_error_locations_non_uniform (CL:171-181): k sequential draws proportional to w,
    a draw that lands on an occupied slot is repeated -- i.e. draw j is proportional to w over the free slots.
    (Not the conditional-Bernoulli law of independent slot errors given k hits.)"""
    w = np.asarray(w, dtype=float)
    n = len(w)
    hit = np.zeros(n)
    for seq in itertools.permutations(range(n), k):
        pr, rest = 1.0, w.sum()
        for i in seq:
            pr *= w[i] / rest
            rest -= w[i]
        for i in seq:
            hit[i] += pr
    return hit


def test_weighted_placement_follows_the_sequential_rejection_law(error_models, correction_table, bitflip_table):
    """k_place, weighted branch: per-slot hit frequencies over 2^20 shots against the exact law of the reference's
    sequential draw-with-rejection, for the linear-heating weights of importance_sampling_syndrome_processing.py:20-36
    (both rounds, 48 slots, k = 2) and a steep 12-slot list with k = 3; the uniform types of the same call must be
    uniform.  Tolerance: 4.5 binomial sigma per slot (60 + slots tested)."""
    heat = [(i // 3 + 1) * 5e-3 for i in range(24)] + [(i // 3 + 9) * 5e-3 for i in range(24)]
    steep = [2.0 ** (-i) for i in range(12)] + [0.0] * 12
    locs = [24, 36, 48, 156, 48]
    n, batch = 1 << 20, 1 << 17
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, batch)
    sim.backend.set_subset([3, 1, 2, 1, 0], locs, [steep, None, heat, None, None])
    hits = np.zeros(sum(locs))
    for k in range(n // batch):
        sim.backend.run("s2", batch, L.NOISE_SUBSET, seed=99, first_shot_id=k * batch)
        m = np.unpackbits(sim.backend.masks_packed(batch)[:, :L.MASK_WORDS].copy().view(np.uint8), axis=1,
                          bitorder="little")[:, :sum(locs)]
        assert (m[:, :24].sum(axis=1) == 3).all() and (m[:, 60:108].sum(axis=1) == 2).all()
        assert (m[:, 24:60].sum(axis=1) == 1).all() and (m[:, 264:].sum(axis=1) == 0).all()
        hits += m.sum(axis=0)
    sim.close()
    off = np.cumsum([0] + locs)
    expect = {0: sequential_rejection_marginals(steep[:12], 3).tolist() + [0.0] * 12,
              2: sequential_rejection_marginals(heat, 2), 1: np.full(36, 1 / 36), 3: np.full(156, 1 / 156)}
    for t, e in expect.items():
        f = hits[off[t]:off[t + 1]] / n
        e = np.asarray(e)
        sig = np.sqrt(np.maximum(e * (1 - e), 1e-12) / n)
        assert np.all(np.abs(f - e) < 4.5 * sig + 1e-12), (t, np.max(np.abs(f - e) / sig))
    # and the law is NOT "k draws proportional to the weights": on the steep list the two differ by hundreds of sigma
    naive = np.minimum(1.0, 3 * np.asarray(steep[:12]) / np.sum(steep))
    f = hits[off[0]:off[0] + 12] / n
    assert np.max(np.abs(f - naive) / np.sqrt(np.maximum(naive * (1 - naive), 1e-9) / n)) > 50


def test_s2_needs_two_round_lists(error_models, correction_table, bitflip_table):
    """A second (stab_ind=1) cycle indexes slot + len/2 of every list (CS:191-193): the reference raises IndexError
    when it is given one-round lists; the library refuses the run instead of reading the next list's slots."""
    locs = [12, 18, 24, 78, 24]
    sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, 64)
    sim.backend.set_subset([0, 0, 1, 0, 0], locs)
    sim.backend.run("s1", 64, L.NOISE_SUBSET, seed=1)
    with pytest.raises(L.S17Error, match="two-round"):
        sim.backend.run("s2", 64, L.NOISE_SUBSET, seed=1)
    sim.close()


# ---------------------------------------------------------------------------------------------------------------------
# the production kernel's block against the reference's block
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prep_cache", ["1", "0"])
def test_cycle_kernel_block_equals_reference_block(error_models, correction_table, bitflip_table, golden, prep_cache,
                                                   monkeypatch):
    """k_cycle1 (the production path: diagonal frame, early X-ancilla measurement, lazy correction) on the golden
    replays: the 512-amplitude block it stores after the last executed cycle equals the block of the UNMODIFIED
    reference at that point (tests/golden/replay_blocks.npz `pre_blocks`: psi[:512] before lookup / apply_correction)
    within 1e-10 per amplitude, phase included.  With sympathetic cooling an idle Z can hit an X-type ancilla after its
    (early) measurement, which leaves a global sign: those models are compared up to that sign."""
    monkeypatch.setenv("S17_DIAG", "1")
    monkeypatch.setenv("S17_PREP_CACHE", prep_cache)
    reps = golden("replays.json")["replays"]
    pre = np.load(os.path.join(GOLDEN, "replay_blocks.npz"))["pre_blocks"]
    groups = {}
    for r in reps:
        groups.setdefault((r["model"], r["protocol"], r["cooling"]), []).append(r)
    checked = 0
    for (model, protocol, cooling), rs in groups.items():
        list_len = [len(x) for x in rs[0]["p_set"]]
        sim = make_sim(error_models, correction_table, bitflip_table, model, list_len, len(rs), 2, cooling)
        sim.backend.set_replay([flat(r["p_set"]) for r in rs], [r["outcomes"] for r in rs])
        c = sim.backend.run(protocol, len(rs), L.NOISE_REPLAY, forced=True)
        assert c.errors == 0
        assert c.sweeps <= 3 * len(rs)                  # one whole-cycle sweep per shot and cycle: really k_cycle1
        for i, r in enumerate(rs):
            psi = sim.backend.state(i)
            want = pre[r["pre_block_index"]]
            err = np.max(np.abs(psi[:512] - want))
            if cooling:
                err = min(err, np.max(np.abs(psi[:512] + want)))
            assert err < AMP_TOL, (model, protocol, i, err)
            assert r["pre_block_norm_outside"] < 1e-20 and np.sum(np.abs(psi[512:]) ** 2) == 0.0
            checked += 1
        sim.close()
    assert checked == len(reps) >= 36


def test_subset_sweep_equals_the_per_subset_drivers(error_models, correction_table, bitflip_table):
    """subset_sweep (one counter exchange for a whole sweep of subsets) returns, subset by subset, what the s2 driver
    returns for the same seeds."""
    import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as drv
    drv.VERBOSE = False
    locs = [24, 36, 48, 156, 48]
    probs = [[1e-4], [1e-4], [5e-3], [1e-3], [5e-3]]
    esets = [[0, 0, 1, 0, 0], [0, 0, 2, 0, 0], [1, 1, 1, 1, 0]]
    random.seed(2026)
    sweep = drv.subset_sweep("s2", 20000, correction_table, "PDD_model", esets, locs, bitflip_table, probs)
    random.seed(2026)
    single = [drv.s2_calculate_log_e_rate_error_subset(20000, correction_table, "PDD_model", e, locs, bitflip_table, probs)
              for e in esets]
    drv.release()
    assert [(f, n0) for f, n0, _c in sweep] == [(r[2], r[3]) for r in single]
    assert all(c == n0 for _f, n0, c in sweep) and sweep[0][0] == 0           # s2 counts the shots that went on; distance 3


def test_per_kernel_class_timing(error_models, correction_table, bitflip_table, monkeypatch):
    """s17_last_timing_detail: the full-state mode materialises one collapse per shot after the preparation cycle and one
    per shot that goes on to a second cycle, samples three All(Measure)s, and the classes' device times are part of the
    measurement time of s17_last_timing."""
    monkeypatch.setenv("S17_DENSE", "1")
    locs = [24, 36, 48, 156, 48]
    n = 64
    sim = api.Simulation("PDD_model", locs, max_shots=n, fusion=1, correction_table=correction_table,
                         bitflip_table=bitflip_table, error_models=error_models, early_measure=False)
    sim.backend.set_subset([0, 0, 2, 1, 1], locs)
    c = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=3)
    t = sim.backend.timing()
    cl = t["classes"]
    assert cl["collapse"]["launches"] == 2 and cl["collapse"]["units"] == n + c.not0
    assert cl["sample"]["launches"] == 3 and cl["plan"]["launches"] == 3
    assert 0 < cl["collapse"]["ms"] + cl["sample"]["ms"] <= t["measure_ms"] + 1e-6
    assert t["gate_launches"] == 24 and c.sweeps == 8 * (2 * n + c.not0)
    sim.close()


@pytest.mark.parametrize("switch", ["S17_TRIVIAL_PREP", "S17_FUSE_PREP", "S17_GEO"])
def test_launch_shortcuts_do_not_change_results(error_models, correction_table, bitflip_table, switch, monkeypatch):
    """S17_TRIVIAL_PREP (no planning launch and no shot list for the preparation cycle of a leak-free model; the call's
    reset rides in the next k_plan) and S17_FUSE_PREP (the preparation cycle of an un-memoised run computed inside the
    launch of the first noisy cycle, block in registers, its outcomes handed to the bookkeeping through the same words)
    give bit-identical counters, records, masks and stored blocks with every cycle computed per shot.  S17_GEO (geometric skipping over the survival products instead of one Bernoulli draw per entry)
    draws other uniforms, so the trajectories differ: the rates of 2^18 true-probability shots agree within 4 sigma."""
    monkeypatch.setenv("S17_PREP_CACHE", "0")
    locs = [24, 36, 48, 156, 48]
    out = {}
    for val in ("1", "0"):
        monkeypatch.setenv(switch, val)
        if switch != "S17_GEO":
            n = 4096
            sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", locs, n)
            sim.backend.set_subset([0, 1, 2, 1, 1], locs)
            c = sim.backend.run("s2", n, L.NOISE_SUBSET, seed=31, first_shot_id=777)
            out[val] = ((c.failed, c.not0, c.counted, c.errors, c.sweeps), sim.backend.records(n),
                        sim.backend.masks_packed(n)[:, :L.MASK_WORDS].copy(),
                        np.stack([sim.backend.state(i)[:512] for i in range(0, n, 128)]),
                        sim.backend.timing()["launches"])
        else:
            n = 1 << 18
            per_round = [12, 18, 24, 78, 24]
            sim = make_sim(error_models, correction_table, bitflip_table, "PDD_model", per_round, n)
            sim.backend.set_slot_probs([[2e-3] * 12, [2e-3] * 18, [2e-2] * 24, [2e-3] * 78, [2e-3] * 24])
            c = sim.backend.run("s1", n, L.NOISE_PROB, seed=31)
            out[val] = (c.failed, c.not0, c.counted)
        sim.close()
    if switch != "S17_GEO":
        assert out["1"][0] == out["0"][0] and out["1"][1] == out["0"][1]
        assert np.array_equal(out["1"][2], out["0"][2]) and np.array_equal(out["1"][3], out["0"][3])
        # launches of an s2 call: plan, cycle (preparation + first), plan, cycle, read-out = 5; without the fused
        # preparation cycle one more (its own cycle launch), planned like any other cycle one more again
        assert out["1"][4] == out["0"][4] - (2 if switch == "S17_TRIVIAL_PREP" else 1)
    else:
        for k in (0, 1):
            a, b = out["1"][k] / n, out["0"][k] / n
            assert abs(a - b) < 4 * np.sqrt((a * (1 - a) + b * (1 - b)) / n) + 1e-9, (k, a, b)
        assert 0.05 < out["1"][1] / n < 0.95
