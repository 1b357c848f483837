"""Mint tests/golden/* by running the UNMODIFIED reference on oracle/pqshim.  Run in the build
container only (needs /root/reference):   python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The reference's functions are imported from /root/reference and
called as they are; nothing is copied.  Per-shot records drive the reference's own
compiled_surface_code_arbitrary_error_model.py functions in the order of
calc_log_e_rate_arbitrary_error_model.py:261-353 / 355-457 / 183-259 with explicit masks, while
the `counts_*` entries call the reference's s1_/s2_/single drivers themselves with seeded RNGs.
"""
import contextlib
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("S17_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")

sys.path.insert(0, os.path.join(HERE, "pqshim"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

LOC = {"PDD_model": [12, 18, 24, 78, 24], "Dress_model": [12, 18, 24, 78, 24],
       "PDD_cooling": [12, 18, 24, 78, 24, 119], "Dress_cooling": [12, 18, 24, 78, 24, 119]}


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ref_shot(CS, pq, model, p_set, protocol, cooling, forced=None, seed=None, want_block=False):
    """The loop body of CL:261-353 (s1), CL:355-457 (s2) or CL:183-259 (single), one shot,
    calling the reference's own functions; returns a record in the oracle's format."""
    from projectq import MainEngine
    from projectq.backends import Simulator
    log = []
    pq.control.reset(forced=forced or (), log=log, seed=seed)
    leaked = 17 * [0]
    e_model, probs = CS.instantiate_error_model(p_set, model)
    eng = MainEngine(Simulator())
    data = eng.allocate_qureg(9)
    ancilla = eng.allocate_qureg(8)
    quiescent = CS.logical_prep(data, "Z", 0, ancilla, leaked, eng, e_model, probs, False)
    synd1 = np.array(CS.stabilizer_cycle_error_index(data, ancilla, leaked, eng, e_model, probs, reset=True,
                                                     cooling=cooling))
    flips_a = (np.array(quiescent) - synd1) % 2
    rec = {"quiescent": [int(b) for b in quiescent], "synd1": [int(b) for b in synd1],
           "flips_a": [int(b) for b in flips_a]}
    stop = bool(np.all(flips_a == 0))
    ft = None
    if protocol == "single":
        ft = flips_a
    elif protocol == "s1":
        ft = flips_a if stop else None
    elif protocol == "s2" and not stop:
        synd2 = np.array(CS.stabilizer_cycle_error_index(data, ancilla, leaked, eng, e_model, probs, stab_ind=1,
                                                         reset=True, cooling=cooling))
        rec["synd2"] = [int(b) for b in synd2]
        ft = (np.array(quiescent) - synd2) % 2
    rec["not0"] = int(not stop)
    rec["counted"] = int(ft is not None)
    if want_block:
        # the state the last executed cycle leaves behind (ancillas measured and reset: only psi[:512] is populated),
        # BEFORE lookup / apply_correction -- what a whole-cycle kernel stores
        eng.flush()
        rec["_pre_block"] = eng.backend.psi[:512].copy()
        rec["pre_block_norm_outside"] = float(np.sum(np.abs(eng.backend.psi[512:]) ** 2))
    if ft is not None:
        vec = CS.lookup(ft, correction_table)
        CS.apply_correction(vec, data)
        eng.flush()
        rec["ft_syndrome"] = [int(b) for b in ft]
        rec["correction"] = [int(b) for b in vec]
        if want_block:
            psi = eng.backend.psi
            rec["block_norm_outside"] = float(np.sum(np.abs(psi[512:]) ** 2))
            rec["_block"] = psi[:512].copy()
        n_before = len(log)
        logic = CS.logical_measurement(data, "Z", eng, bitflip_table, quiescent, leaked)
        dm = [e[2] for e in log[n_before:] if e[0] == "meas"]
        dm = [1 if leaked[i] == 1 else dm[i] for i in range(9)]
        rec["data_meas"] = dm
        rec["logical_meas"] = int(logic)
        rec["fail"] = int(logic != 0)
    else:
        rec["fail"] = 0
        from projectq.ops import All, Measure
        All(Measure) | ancilla
        All(Measure) | data
        eng.flush()
    rec["leaked"] = list(leaked)
    rec["outcomes"] = [int(e[2]) for e in log if e[0] == "meas"]
    rec["meas_probs"] = [float(e[3]) for e in log if e[0] == "meas"]
    rec["n_gates"] = sum(1 for e in log if e[0] == "gate")
    return rec, log


def random_masks(rng, locs, eset, two_round):
    out = []
    for L, k in zip(locs, eset):
        n = 2 * L if two_round else L
        m = n * [0]
        for i in rng.sample(range(n), k):
            m[i] = 1
        out.append(m)
    return out


def main():
    global correction_table, bitflip_table
    os.makedirs(OUT, exist_ok=True)
    os.chdir(REF)  # the reference opens error_models.json relative to the cwd (CS:38)
    with quiet():
        import compiled_surface_code_arbitrary_error_model as CS
        import calc_log_e_rate_arbitrary_error_model as CL
    import projectq as pq
    correction_table = CS.load_lookup_table("correction_table_depolarising.json")
    bitflip_table = CS.load_lookup_table("correction_table_classical_bitflip.json")

    # 1. program trace of one noiseless cycle + prep (SURVEY.md Appendix A is the expected answer)
    rec, log = ref_shot(CS, pq, "PDD_model", [12 * [0], 18 * [0], 24 * [0], 78 * [0], 24 * [0]], "s1", False, seed=7)
    first_meas = next(i for i, e in enumerate(log) if e[0] == "meas")
    trace = [[e[1], e[2], list(e[3])] for e in log[:first_meas]]
    json.dump({"trace": trace}, open(os.path.join(OUT, "cycle_trace.json"), "w"))
    assert len(trace) == 54

    # 2. replay trajectories: masks + recorded outcomes -> records (bit-exact fields) + 512-amp blocks
    rng = random.Random(20261019)
    replays, blocks, pre_blocks = [], [], []
    cases = []
    for model in ("PDD_model", "Dress_model", "PDD_cooling", "Dress_cooling"):
        nt = len(LOC[model])
        for protocol in ("s1", "s2", "single"):
            for rep in range(4 if model == "PDD_model" else 3):
                k = rng.choice([1, 1, 2, 2, 3, 4])
                eset = nt * [0]
                for _ in range(k):
                    eset[rng.randrange(nt)] += 1
                if protocol == "s1" and rep == 0:
                    eset = nt * [0]
                cases.append((model, protocol, eset))
    for ci, (model, protocol, eset) in enumerate(cases):
        cooling = model.endswith("cooling")
        p_set = random_masks(rng, LOC[model], eset, protocol == "s2")
        with quiet():
            rec, log = ref_shot(CS, pq, model, p_set, protocol, cooling, seed=1000 + ci, want_block=True)
            # replay with the recorded outcomes forced: must reproduce itself
            rec2, _ = ref_shot(CS, pq, model, p_set, protocol, cooling, forced=rec["outcomes"], want_block=True)
        blk = rec.pop("_block", None)
        blk2 = rec2.pop("_block", None)
        pre, pre2 = rec.pop("_pre_block"), rec2.pop("_pre_block")
        assert np.allclose(pre, pre2, atol=1e-14)
        rec["pre_block_index"] = len(pre_blocks)
        rec2["pre_block_index"] = len(pre_blocks)
        pre_blocks.append(pre)
        assert {k: v for k, v in rec.items() if k != "meas_probs"} == {k: v for k, v in rec2.items() if k != "meas_probs"}
        if blk is not None:
            assert np.allclose(blk, blk2, atol=1e-14)
        rec.update({"model": model, "protocol": protocol, "cooling": cooling, "eset": eset, "p_set": p_set,
                    "block_index": len(blocks) if blk is not None else -1})
        if blk is not None:
            blocks.append(blk)
        replays.append(rec)
    json.dump({"replays": replays}, open(os.path.join(OUT, "replays.json"), "w"))
    np.savez_compressed(os.path.join(OUT, "replay_blocks.npz"), blocks=np.array(blocks), pre_blocks=np.array(pre_blocks))

    # 3. single-fault table for PDD_model under the s2 protocol (two-round masks one-hot in round 1)
    table = []
    locs = LOC["PDD_model"]
    for t, L in enumerate(locs):
        for i in range(L):
            p_set = [2 * l * [0] for l in locs]
            p_set[t][i] = 1
            with quiet():
                rec, _ = ref_shot(CS, pq, "PDD_model", p_set, "s2", False, seed=5000 + 100 * t + i)
            table.append({"type": t, "slot": i, "flips_a": rec["flips_a"], "ft": rec.get("ft_syndrome"),
                          "fail": rec["fail"], "outcomes": rec["outcomes"], "quiescent": rec["quiescent"]})
    json.dump({"rows": table}, open(os.path.join(OUT, "single_fault_pdd.json"), "w"))

    # 4. the reference's own drivers, seeded (python `random` for placement/errors, shim rng for outcomes)
    counts = []
    for (fn, proto, model, eset, two, cooling, runs, seed) in [
        ("s1_calculate_log_e_rate_error_subset", "s1", "PDD_model", [0, 0, 1, 0, 0], False, False, 24, 11),
        ("s1_calculate_log_e_rate_error_subset", "s1", "PDD_model", [1, 0, 0, 1, 0], False, False, 24, 12),
        ("s2_calculate_log_e_rate_error_subset", "s2", "PDD_model", [0, 0, 2, 0, 0], True, False, 24, 13),
        ("s2_calculate_log_e_rate_error_subset", "s2", "Dress_model", [0, 1, 1, 2, 0], True, False, 24, 14),
        ("s2_calculate_log_e_rate_error_subset", "s2", "Dress_cooling", [0, 0, 1, 1, 0, 2], True, True, 16, 15),
        ("s1_calculate_log_e_rate_error_subset", "s1", "PDD_cooling", [0, 0, 0, 0, 1, 1], False, True, 16, 16),
    ]:
        locs = [(2 * l if two else l) for l in LOC[model]]
        prob_lists = [[1e-3] for _ in locs]
        if model == "PDD_model" and proto == "s2":   # exercise the non-uniform placement path (CL:171)
            prob_lists[2] = [(i // 3 + 1) * 5e-3 for i in range(locs[2])]
        random.seed(seed)
        pq.control.reset(seed=seed + 100)
        with quiet():
            r = getattr(CL, fn)(runs, correction_table, model, eset, locs, bitflip_table, prob_lists,
                                cooling=cooling)
        counts.append({"fn": fn, "protocol": proto, "model": model, "eset": eset, "locations": locs,
                       "prob_lists": prob_lists, "cooling": cooling, "runs": runs, "seed": seed,
                       "meas_seed": seed + 100, "failed": int(r[2]), "not0": int(r[3]),
                       "rate": r[0], "not0_rate": r[1]})
    # single-round driver CL:183 (uniform placement branch)
    random.seed(21)
    pq.control.reset(seed=121)
    with quiet():
        rate = CL.calculate_log_e_rate_error_subset(24, correction_table, "PDD_model", [0, 0, 2, 1, 0],
                                                    LOC["PDD_model"], bitflip_table, [1e-3] * 5)
    counts.append({"fn": "calculate_log_e_rate_error_subset", "protocol": "single", "model": "PDD_model",
                   "eset": [0, 0, 2, 1, 0], "locations": LOC["PDD_model"], "prob_lists": [[1e-3]] * 5,
                   "cooling": False, "runs": 24, "seed": 21, "meas_seed": 121, "failed": int(round(rate * 24)),
                   "not0": None, "rate": rate, "not0_rate": None})
    json.dump({"counts": counts}, open(os.path.join(OUT, "driver_counts.json"), "w"))

    # 5. weights KATs (CL:460-500)
    kat = {
        "subset_weight": CL.subset_weight([(12, 1e-4, 0), (18, 1e-4, 0), (24, 5e-3, 1), (78, 1e-3, 0), (24, 5e-3, 0)]),
        "subset_weight_mixed": CL.subset_weight_mixed([(24, [5e-3], 2)]),
        "weight_contribution_variable": CL.weight_contribution_variable_e_rate(
            8, [1e-3 * (i + 1) for i in range(8)], 2),
    }
    json.dump(kat, open(os.path.join(OUT, "weights_kat.json"), "w"))
    print("golden written to", OUT, "replays:", len(replays), "blocks:", len(blocks), "fault rows:", len(table))


if __name__ == "__main__":
    main()
