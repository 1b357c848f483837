"""Mint tests/golden/stats_s2.json: sampled counters of the UNMODIFIED reference drivers
(calc_log_e_rate_arbitrary_error_model.py:261-353 s1, :355-457 s2) on oracle/pqshim, a few thousand shots per error
subset, for the statistical parity tests (rates within binomial 3 sigma).  Build container only (needs /root/reference):

    python -m oracle.make_golden_stats [runs_per_subset]

Entries already in the file (same model / protocol / eset / locations / prob_lists / cooling) are kept, new ones are
sampled and appended, so the file only grows.

TEST INFRASTRUCTURE ONLY.  Nothing of the reference is copied: its functions are imported and called as they are.
The parameter sets are the reference scripts' own: importance_sampling_syndrome_processing.py:18-36 (PDD_model, linear
motional heating of p_XX_ctrl: non-uniform placement, CL:171-181) and
importance_sampling_cooling_syndrome_processing.py:19-34 (Dress_cooling, cooling=True).
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("S17_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden", "stats_s2.json")

LOC1 = [12, 18, 24, 78, 24]
LOC2 = [2 * x for x in LOC1]
LOC1C = LOC1 + [119]
LOC2C = [2 * x for x in LOC1C]
HEAT1 = [(i // 3 + 1) * 5e-3 for i in range(24)]             # IS:25-27: (i + 1) * p per timestep, three gates each
HEAT2 = [(i // 3 + 1 + 8) * 5e-3 for i in range(24)]
PDD = [[1e-4], [1e-4], [5e-3], [1e-3], [5e-3]]
PDD_HEAT1 = [[1e-4], [1e-4], HEAT1, [1e-3], [5e-3]]
PDD_HEAT2 = [[1e-4], [1e-4], HEAT1 + HEAT2, [1e-3], [5e-3]]
ISC = [[1e-4], [1e-4], [1e-3], [1e-3], [1e-3], [1e-3]]       # ISC:21-34
DRESS = [[1e-4], [1e-4], [1e-3], [1e-3], [1e-3]]

CONFIGS = [
    # (model, protocol, eset, locations, prob_lists, cooling)
    ("PDD_model", "s2", [0, 0, 2, 1, 1], LOC2, PDD, False),
    ("PDD_model", "s2", [0, 0, 2, 0, 0], LOC2, PDD, False),
    ("PDD_model", "s2", [1, 1, 1, 1, 0], LOC2, PDD, False),
    ("PDD_model", "s2", [0, 0, 2, 1, 0], LOC2, PDD_HEAT2, False),
    ("PDD_model", "s2", [0, 0, 3, 0, 0], LOC2, PDD_HEAT2, False),
    ("PDD_model", "s1", [0, 0, 2, 0, 0], LOC1, PDD_HEAT1, False),
    ("Dress_model", "s2", [0, 0, 1, 2, 0], LOC2, DRESS, False),
    ("Dress_model", "s2", [0, 1, 1, 1, 1], LOC2, DRESS, False),
    ("Dress_cooling", "s2", [0, 0, 1, 1, 0, 2], LOC2C, ISC, True),
    ("PDD_cooling", "s2", [0, 0, 1, 0, 1, 2], LOC2C, ISC, True),
    ("Dress_cooling", "s1", [0, 0, 0, 1, 0, 2], LOC1C, ISC, True),
]


def _worker(args):
    seed, runs, (model, protocol, eset, locations, prob_lists, cooling) = args
    sys.path.insert(0, os.path.join(HERE, "pqshim"))
    sys.path.insert(0, REF)
    os.chdir(REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import calc_log_e_rate_arbitrary_error_model as CL
        import compiled_surface_code_arbitrary_error_model as CS
        import projectq as pq
        ct = CS.load_lookup_table("correction_table_depolarising.json")
        bt = CS.load_lookup_table("correction_table_classical_bitflip.json")
        random.seed(seed)                                   # error placement (the reference's global `random`)
        pq.control.reset(forced=(), log=None, seed=seed + 100000)      # measurement outcomes of the stand-in backend
        fn = CL.s2_calculate_log_e_rate_error_subset if protocol == "s2" else CL.s1_calculate_log_e_rate_error_subset
        r = fn(runs, ct, model, eset, locations, bt, prob_lists, cooling=cooling)
    return int(r[2]), int(r[3]), runs


def same(entry, cfg):
    model, protocol, eset, locations, prob_lists, cooling = cfg
    return (entry["model"] == model and entry["protocol"] == protocol and entry["eset"] == eset and
            entry["locations"] == locations and entry["prob_lists"] == prob_lists and
            bool(entry.get("cooling", False)) == cooling)


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    procs = os.cpu_count() or 1
    per = (runs + procs - 1) // procs
    out = json.load(open(OUT))["stats"] if os.path.exists(OUT) else []
    with mp.get_context("fork").Pool(procs) as pool:
        for k, cfg in enumerate(CONFIGS):
            if any(same(e, cfg) for e in out):
                continue
            model, protocol, eset, locations, prob_lists, cooling = cfg
            res = pool.map(_worker, [(1000 * k + i, per, cfg) for i in range(procs)])
            out.append({"model": model, "protocol": protocol, "eset": eset, "locations": locations,
                        "prob_lists": prob_lists, "cooling": cooling,
                        "runs": sum(r[2] for r in res), "failed": sum(r[0] for r in res), "not0": sum(r[1] for r in res),
                        "seeds": [1000 * k + i for i in range(procs)]})
            print({k_: v for k_, v in out[-1].items() if k_ != "prob_lists"}, flush=True)
            json.dump({"source": "reference s1_/s2_calculate_log_e_rate_error_subset on oracle/pqshim", "stats": out},
                      open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
