"""Mint tests/golden/rtf_ref.json: rounds-til-fail lists of the UNMODIFIED reference run-til-fail driver
(calc_log_e_rate_arbitrary_error_model.py:33-148, `calculate_log_e_rate`) on oracle/pqshim.  Build container only (needs
/root/reference):

    python -m oracle.make_golden_rtf [runs_per_seed]

TEST INFRASTRUCTURE ONLY.  Nothing of the reference is copied: its functions are imported and called as they are, in a
scratch working directory (the driver writes data/<f>.json and figs/<f>.png; matplotlib is the no-op stub of the shim).
One chain = one call of the driver: `random.seed(seed)` fixes the error draws, the shim's own generator (seed + 100) the
measurement outcomes -- `oracle.s17_oracle.run_til_fail` with the same two seeds must return the same lists
(tests/test_oracle_golden.py).  Note the reference keeps ONE leak register for all runs of a call (CL:49, SURVEY C-6).
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("S17_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden", "rtf_ref.json")


def p_set_for(model, p, p_leak):
    """vary_p.py:14-20 with a 78-long fourth list (SURVEY C-2): px = py = p/10, pxx = p, dephasing / leakage, heating 0."""
    return [12 * [p / 10], 18 * [p / 10], 24 * [p], 78 * [p_leak], 24 * [0]]


CASES = [("PDD_model", 0.03, 0.0), ("Dress_model", 0.03, 4e-3)]


def _worker(args):
    model, p, p_leak, seed, runs = args
    sys.path.insert(0, os.path.join(HERE, "pqshim"))
    sys.path.insert(0, REF)
    os.chdir(REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import calc_log_e_rate_arbitrary_error_model as CL
        import compiled_surface_code_arbitrary_error_model as CS
        import projectq as pq
        ct = CS.load_lookup_table("correction_table_depolarising.json")
        bt = CS.load_lookup_table("correction_table_classical_bitflip.json")
        e_model, probs = CS.instantiate_error_model(p_set_for(model, p, p_leak), model)
        tmp = tempfile.mkdtemp(prefix="s17_rtf_")
        os.makedirs(os.path.join(tmp, "data"))
        os.makedirs(os.path.join(tmp, "figs"))
        os.chdir(tmp)
        random.seed(seed)
        pq.control.reset(forced=(), log=None, seed=seed + 100)
        try:
            fit = float(CL.calculate_log_e_rate(runs, "chain", ct, e_model, probs, bt, num_bins=5))
        except Exception:                    # the fit may not converge on a handful of runs; the lists are what we pin
            fit = None
        res = json.load(open(os.path.join(tmp, "data", "chain.json")))
    return {"seed": seed, "meas_seed": seed + 100, "runs": runs, "rounds_til_fail_list": res["rounds_til_fail_list"],
            "syndrome_b_measured_list": res["syndrome_b_measured_list"], "reference_fit": fit}


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    procs = os.cpu_count() or 1
    out = []
    with mp.get_context("fork").Pool(procs) as pool:
        for k, (model, p, p_leak) in enumerate(CASES):
            chains = pool.map(_worker, [(model, p, p_leak, 7000 + 100 * k + i, runs) for i in range(procs)])
            out.append({"model": model, "p": p, "p_leak": p_leak, "p_set": p_set_for(model, p, p_leak), "chains": chains})
            print(model, [c["rounds_til_fail_list"] for c in chains], flush=True)
    json.dump({"source": "reference calculate_log_e_rate (CL:33-148) on oracle/pqshim; one leak register per chain (CL:49)",
               "cases": out}, open(OUT, "w"))


if __name__ == "__main__":
    main()
