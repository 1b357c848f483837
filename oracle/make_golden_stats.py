"""Mint tests/golden/stats_s2.json: sampled s2 counters of the UNMODIFIED reference drivers
(calc_log_e_rate_arbitrary_error_model.py:355-457) on oracle/pqshim, a few thousand shots per error subset, for the
statistical parity test (rates within binomial 3 sigma).  Build container only (needs /root/reference):

    python -m oracle.make_golden_stats [runs_per_subset]

TEST INFRASTRUCTURE ONLY.  Nothing of the reference is copied: its functions are imported and called as they are.
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

MODEL = "PDD_model"
LOCATIONS = [24, 36, 48, 156, 48]
PROB_LISTS = [[1e-4], [1e-4], [5e-3], [1e-3], [5e-3]]
SUBSETS = [[0, 0, 2, 1, 1], [0, 0, 2, 0, 0], [1, 1, 1, 1, 0]]


def _worker(args):
    seed, runs, eset = args
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
        r = CL.s2_calculate_log_e_rate_error_subset(runs, ct, MODEL, eset, LOCATIONS, bt, PROB_LISTS)
    return int(r[2]), int(r[3]), runs


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    procs = os.cpu_count() or 1
    per = (runs + procs - 1) // procs
    out = []
    with mp.get_context("fork").Pool(procs) as pool:
        for k, eset in enumerate(SUBSETS):
            res = pool.map(_worker, [(1000 * k + i, per, eset) for i in range(procs)])
            out.append({"model": MODEL, "protocol": "s2", "eset": eset, "locations": LOCATIONS, "prob_lists": PROB_LISTS,
                        "runs": sum(r[2] for r in res), "failed": sum(r[0] for r in res), "not0": sum(r[1] for r in res),
                        "seeds": [1000 * k + i for i in range(procs)]})
            print(out[-1], flush=True)
    json.dump({"source": "reference s2_calculate_log_e_rate_error_subset on oracle/pqshim", "stats": out}, open(OUT, "w"),
              indent=1)


if __name__ == "__main__":
    main()
