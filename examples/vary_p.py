#!/usr/bin/env python
"""Pseudo-threshold sweep with run-til-fail (the reference's vary_p.py:9-27, with the 78-entry dephasing / leakage
list the cycle really indexes -- the shipped script passes 30 entries and raises IndexError, SURVEY C-2).

    python examples/vary_p.py [--model PDD_model|Dress_model] [--runs 2000] [--points 3]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as CL          # noqa: E402


def sweep(model="PDD_model", runs=2000, ps=(0.01, 0.007, 0.005), p_second=1e-4, num_bins=20, save=False):
    CL.VERBOSE = False
    correction_table = CL.load_lookup_table("correction_table_depolarising.json")
    bitflip_table = CL.load_lookup_table("correction_table_classical_bitflip.json")
    out = {}
    for p in ps:
        p_set = [12 * [p / 10], 18 * [p / 10], 24 * [p], 78 * [p_second], 24 * [0.0]]   # BASELINE config 5 shape
        e_model, e_probs = CL.instantiate_error_model(p_set, model)
        rate = CL.calculate_log_e_rate(runs, "{}_p={}".format(model, p), correction_table, e_model, e_probs, bitflip_table,
                                       num_bins=num_bins, model_name=model, save=save)
        out[str(p)] = {"log_e_rate_fit": float(rate), "rounds": CL.LAST_TIMING.get("rounds")}
    CL.release()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="PDD_model")
    ap.add_argument("--runs", type=int, default=2000)
    ap.add_argument("--points", type=int, default=3)
    a = ap.parse_args()
    ps = [0.01, 0.007, 0.005] if a.points == 3 else list(np.geomspace(1e-2, 1e-3, a.points))
    print(json.dumps(sweep(a.model, a.runs, ps)))
