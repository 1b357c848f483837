#!/usr/bin/env python
"""End-to-end s1/s2 importance-sampling pipeline -- the stages of the reference's
importance_sampling_syndrome_processing.py (PDD_model with linear motional heating, :8-137) and
importance_sampling_cooling_syndrome_processing.py (Dress_cooling, cooling=True), with every stage enabled,
portable paths, and the shot loops on the B200 backend.

    python examples/importance_sampling_syndrome_processing.py [--model PDD_model|Dress_cooling] [--runs 100000]
                                                                [--cutoff 1e-6] [--out imp_samp]

Stage outputs keep the reference's JSON keys (error_subset_list, subset_weights, cutoff_weight, locations,
probabilities; s1_log_e_rate_dict, s2_rate_dict, failcount, s2count; s2_log_e_rate_dict, ...).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as CL          # noqa: E402
from surface_17_iqt_b200 import importance_sampling as isamp                     # noqa: E402


def parameters(model):
    if model == "PDD_model":             # importance_sampling_syndrome_processing.py:18-36
        locations = [12, 18, 24, 78, 24]
        prob_list = [1e-4, 1e-4, 5e-3, 1e-3, 5e-3]
        heat1, heat2 = [], []
        for i in range(8):
            heat1 += 3 * [(i + 1) * prob_list[2]]
            heat2 += 3 * [(i + 1 + 8) * prob_list[2]]
        s1 = [[prob_list[0]], [prob_list[1]], heat1, [prob_list[3]], [prob_list[4]]]
        s2 = [[prob_list[0]], [prob_list[1]], heat2, [prob_list[3]], [prob_list[4]]]
        both = [[prob_list[0]], [prob_list[1]], heat1 + heat2, [prob_list[3]], [prob_list[4]]]
        return locations, s1, s2, both, False
    if model == "Dress_cooling":         # importance_sampling_cooling_syndrome_processing.py:19-34
        locations = [12, 18, 24, 78, 24, 119]
        prob_list = [1e-4, 1e-4, 1e-3, 1e-3, 1e-3, 1e-3]
        s = [[p] for p in prob_list]
        return locations, s, s, s, True
    raise SystemExit("unknown model")


def run_pipeline(model="PDD_model", runs=10000, cutoff=1e-6, out="imp_samp", verbose=True):
    os.makedirs(out, exist_ok=True)
    CL.VERBOSE = False
    correction_table = CL.load_lookup_table("correction_table_depolarising.json")
    bitflip_table = CL.load_lookup_table("correction_table_classical_bitflip.json")
    locations, pl_s1, pl_s2, pl_both, cooling = parameters(model)
    both_rounds = [2 * x for x in locations]

    def dump(name, obj):
        with open(os.path.join(out, name + ".json"), "w") as f:
            json.dump(obj, f)

    t0 = time.perf_counter()
    subsets_s1 = isamp.calculate_significant_subsets_incrementally(locations, pl_s1, cutoff_weight=cutoff)
    subsets_s12 = isamp.calculate_significant_subsets_incrementally(both_rounds, pl_both, cutoff_weight=cutoff)
    dump("incremental_subsets_s1_" + model, subsets_s1)
    dump("incremental_subsets_s1ands2_" + model, subsets_s12)
    if verbose:
        print("{} s1 subsets, {} two-round subsets ({:.2f} s)".format(len(subsets_s1["error_subset_list"]),
                                                                      len(subsets_s12["error_subset_list"]),
                                                                      time.perf_counter() - t0))
    # stage: one round per significant subset (IS:78-92)
    s1_rate, s2_rate, failcount, s2count = {}, {}, {}, {}
    # (one small untimed call first: library load, CUDA context, program upload -- seconds, once per process)
    CL.s1_calculate_log_e_rate_error_subset(64, correction_table, model, subsets_s1["error_subset_list"][0], locations,
                                            bitflip_table, pl_s1, cooling=cooling)
    t0 = time.perf_counter()
    for eset in subsets_s1["error_subset_list"]:
        r = CL.s1_calculate_log_e_rate_error_subset(runs, correction_table, model, eset, locations, bitflip_table, pl_s1,
                                                    cooling=cooling)
        s1_rate[str(eset)], s2_rate[str(eset)], failcount[str(eset)], s2count[str(eset)] = r
    t_s1 = time.perf_counter() - t0
    dump("ImpSamp_{}_syndrome_process_s1subsets_{}".format(runs, model),
         {"s1_log_e_rate_dict": s1_rate, "s2_rate_dict": s2_rate, "failcount": failcount, "s2count": s2count})
    # stage: which two-round subsets matter, given how often each s1 subset goes to a second syndrome (IS:95-105)
    subsets_s2 = isamp.s2sim_calculate_significant_subsets_incrementally(
        locations, pl_s1, pl_s2, s2_rate, subsets_s1["error_subset_list"], subsets_s12["error_subset_list"],
        cutoff_weight=cutoff)
    dump("incremental_subsets_s2_" + model, subsets_s2)
    # stage: two-round simulations (IS:120-137)
    s2_log_e, s2_fail, s2_cnt = {}, {}, {}
    t0 = time.perf_counter()
    for eset in subsets_s2["error_subset_list"]:
        r = CL.s2_calculate_log_e_rate_error_subset(runs, correction_table, model, eset, both_rounds, bitflip_table,
                                                    pl_both, cooling=cooling)
        s2_log_e[str(eset)], _, s2_fail[str(eset)], s2_cnt[str(eset)] = r
    t_s2 = time.perf_counter() - t0
    dump("ImpSamp_{}_syndrome_process_s2subsets_{}".format(runs, model),
         {"s2_log_e_rate_dict": s2_log_e, "failcount": s2_fail, "s2count": s2_cnt})
    p_l = isamp.pl_from_file(subsets_s1["error_subset_list"], subsets_s1["subset_weights"], s1_rate,
                             subsets_s2["error_subset_list"], subsets_s2["subset_weights"], s2_fail, s2_cnt)
    n1, n2 = len(subsets_s1["error_subset_list"]), len(subsets_s2["error_subset_list"])
    result = {"model": model, "runs_per_subset": runs, "p_logical": p_l, "s1_subsets": n1, "s2_subsets": n2,
              "s1_shots_per_s": n1 * runs / t_s1 if t_s1 > 0 else None,
              "s2_shots_per_s": n2 * runs / t_s2 if n2 and t_s2 > 0 else None}
    dump("pl_" + model, result)
    if verbose:
        print(json.dumps(result))
    CL.release()
    return result


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="PDD_model")
    ap.add_argument("--runs", type=int, default=10000)
    ap.add_argument("--cutoff", type=float, default=1e-6)
    ap.add_argument("--out", default="imp_samp")
    a = ap.parse_args()
    run_pipeline(a.model, a.runs, a.cutoff, a.out)
