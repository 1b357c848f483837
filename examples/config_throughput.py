#!/usr/bin/env python
"""Throughput of the five workload shapes of BASELINE.json through the drop-in drivers (wall clock around the calls,
host arguments in, host results out), one JSON line per shape.  Needs a B200.

    python examples/config_throughput.py [--scale 1.0]

1  depolarising_model, true-probability noise p = 1e-3, single round            (config 1; the reference's CPU case)
2  PDD_model, fixed-weight subset [0,0,2,1,1], s2 protocol, uniform placement     (config 2; the bench workload)
3  coherent_model: over-rotations eps = 0.02 on every gate + stochastic XX        (config 3; coherent form of the cycle kernel)
4a PDD_model with linear motional heating, non-uniform placement, s2              (config 4, importance_sampling_...py:20-36)
4b Dress_cooling with sympathetic cooling (119 idle slots per round), s2          (config 4, ..._cooling_...py:19-21)
5  Dress_model (leakage), run-til-fail, p = 1e-2                                   (config 5, vary_p.py)
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as CL          # noqa: E402


def timed(name, shots, unit, fn):
    fn(min(shots, 4096))                       # builds the context, the preparation cache, warms the kernels
    t0 = time.perf_counter()
    res = fn(shots)
    dt = time.perf_counter() - t0
    line = {"config": name, unit: shots, "seconds": dt, unit + "_per_s": shots / dt, "result": res}
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    CL.VERBOSE = False
    ct = CL.load_lookup_table("correction_table_depolarising.json")
    bt = CL.load_lookup_table("correction_table_classical_bitflip.json")
    n = lambda x: max(1024, int(x * a.scale))
    out = []

    p = 1e-3
    depol = [30 * [p / 3], 30 * [p / 3], 30 * [p / 3], 24 * [p]]
    out.append(timed("1 depolarising p=1e-3 single round", n(4e6), "shots",
                     lambda k: CL.calculate_log_e_rate_true_probability(k, ct, "depolarising_model", depol, bt)[:3]))

    locs2 = [24, 36, 48, 156, 48]
    pl = [[1e-4], [1e-4], [5e-3], [1e-3], [5e-3]]
    out.append(timed("2 PDD s2 subset [0,0,2,1,1]", n(8e6), "shots",
                     lambda k: CL.s2_calculate_log_e_rate_error_subset(k, ct, "PDD_model", [0, 0, 2, 1, 1], locs2, bt, pl)[:2]))

    coh = [[0.0] * 30, [0.0] * 24, [5e-3] * 24]
    out.append(timed("3 coherent eps=0.02 + XX 5e-3", n(1e5), "shots",
                     lambda k: CL.calculate_log_e_rate_true_probability(
                         k, ct, "coherent_model", coh, bt, over_angles={"eps_1q": [0.02] * 30, "eps_2q": [0.02] * 24})[:3]))
    CL.release()

    heat1, heat2 = [], []
    for i in range(8):
        heat1 += 3 * [(i + 1) * 5e-3]
        heat2 += 3 * [(i + 1 + 8) * 5e-3]
    pl_heat = [[1e-4], [1e-4], heat1 + heat2, [1e-3], [5e-3]]
    out.append(timed("4a PDD linear heating, non-uniform placement, s2 [0,0,2,1,0]", n(8e6), "shots",
                     lambda k: CL.s2_calculate_log_e_rate_error_subset(k, ct, "PDD_model", [0, 0, 2, 1, 0], locs2, bt, pl_heat)[:2]))

    locs_c = [24, 36, 48, 156, 48, 238]
    pl_c = [[1e-4], [1e-4], [1e-3], [1e-3], [1e-3], [1e-3]]
    out.append(timed("4b Dress_cooling s2 [0,0,1,1,0,2]", n(8e6), "shots",
                     lambda k: CL.s2_calculate_log_e_rate_error_subset(k, ct, "Dress_cooling", [0, 0, 1, 1, 0, 2], locs_c, bt, pl_c,
                                                                       False, True)[:2]))

    def rtf(k):
        p = 1e-2
        p_set = [12 * [p / 10], 18 * [p / 10], 24 * [p], 78 * [1e-4], 24 * [0.0]]
        e_model, e_probs = CL.instantiate_error_model(p_set, "Dress_model")
        CL.calculate_log_e_rate(k, "cfg5", ct, e_model, e_probs, bt, model_name="Dress_model", save=False)
        return CL.LAST_TIMING.get("rounds")
    t0 = time.perf_counter()
    rtf(256)
    t0 = time.perf_counter()
    runs = n(2e5)
    rounds = rtf(runs)
    dt = time.perf_counter() - t0
    line = {"config": "5 Dress_model run-til-fail p=1e-2", "runs": runs, "rounds": rounds, "seconds": dt,
            "rounds_per_s": rounds / dt}
    print(json.dumps(line), flush=True)
    out.append(line)
    CL.release()
    return out


if __name__ == "__main__":
    main()
