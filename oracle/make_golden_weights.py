"""Mint tests/golden/subsets_small.json with the reference's own enumeration code (CL:567-753) for the
parameters of importance_sampling_syndrome_processing.py:18-36.  Build container only:
    python -m oracle.make_golden_weights
TEST INFRASTRUCTURE ONLY."""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("S17_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "pqshim"))
sys.path.insert(0, REF)


def main():
    os.chdir(REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import calc_log_e_rate_arbitrary_error_model as CL
    # a scaled-down version of importance_sampling_syndrome_processing.py:18-36 (same structure: constant rates for
    # four error types, a per-location "heating" list for the third); the reference's own enumeration of all
    # C(L, k) placements (CL:473) makes the original sizes take hours
    locations = [3, 4, 6, 5, 2]
    both = [2 * x for x in locations]
    p = [0.05, 0.02, 0.01, 0.03, 0.04]
    heat1 = [(i // 2 + 1) * p[2] for i in range(6)]
    heat2 = [(i // 2 + 4) * p[2] for i in range(6)]
    pl1 = [[p[0]], [p[1]], heat1, [p[3]], [p[4]]]
    pl2 = [[p[0]], [p[1]], heat2, [p[3]], [p[4]]]
    plb = [[p[0]], [p[1]], heat1 + heat2, [p[3]], [p[4]]]
    cutoff = 1e-3
    with contextlib.redirect_stdout(io.StringIO()):
        s1 = CL.calculate_significant_subsets_incrementally(locations, pl1, cutoff_weight=cutoff)
        s12 = CL.calculate_significant_subsets_incrementally(both, plb, cutoff_weight=cutoff)
        rates = {str(e): 0.25 + 0.5 * (sum(e) % 3) / 3 for e in s1['error_subset_list']}
        s2 = CL.s2sim_calculate_significant_subsets_incrementally(locations, pl1, pl2, rates, s1['error_subset_list'],
                                                                  s12['error_subset_list'], cutoff_weight=cutoff)
        more = CL.find_significant_subsets_with_more_error([0, 0, 1, 0, 0], s12['error_subset_list'])
        wl = CL.weighted_logical_error_rate(p, s1['error_subset_list'], rates, locations)
    out_extra = {'weighted': [wl[0], wl[1], wl[2]]}
    out = {"locations": locations, "pl1": pl1, "pl2": pl2, "plb": plb, "cutoff": cutoff, "s1": s1, "s12": s12,
           "s2_rates": rates, "s2": s2, "more": more, "weighted": out_extra["weighted"]}
    json.dump(out, open(os.path.join(REPO, "tests", "golden", "subsets_small.json"), "w"))
    print(len(s1['error_subset_list']), len(s12['error_subset_list']), len(s2['error_subset_list']))


if __name__ == "__main__":
    main()
