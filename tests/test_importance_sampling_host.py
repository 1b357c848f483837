"""Host-side importance-sampling bookkeeping against the reference's own outputs (tests/golden/subsets_small.json,
weights_kat.json -- minted by oracle/make_golden*.py from the unmodified reference) and brute force."""
import itertools
import json
import os

import pytest

from surface_17_iqt_b200 import importance_sampling as isamp
from surface_17_iqt_b200 import generate_lookup_table as glt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def brute_variable(n, probs, k):
    tot = 0.0
    for locs in itertools.combinations(range(n), k):
        pr = 1.0
        for i in range(n):
            pr *= probs[i] if i in locs else 1 - probs[i]
        tot += pr
    return tot


def test_weight_kats_from_reference():
    kat = json.load(open(os.path.join(GOLDEN, "weights_kat.json")))
    assert isamp.subset_weight([(12, 1e-4, 0), (18, 1e-4, 0), (24, 5e-3, 1), (78, 1e-3, 0), (24, 5e-3, 0)]) == \
        pytest.approx(kat["subset_weight"], rel=1e-14)
    assert kat["subset_weight"] == pytest.approx(0.08743215323971552, rel=1e-12)
    assert isamp.subset_weight_mixed([(24, [5e-3], 2)]) == pytest.approx(kat["subset_weight_mixed"], rel=1e-14)
    assert isamp.weight_contribution_variable_e_rate(8, [1e-3 * (i + 1) for i in range(8)], 2) == \
        pytest.approx(kat["weight_contribution_variable"], rel=1e-12)


def test_variable_rate_recurrence_equals_enumeration():
    probs = [0.01, 0.2, 0.05, 0.3, 0.07, 0.11, 0.02]
    for k in range(0, 8):
        assert isamp.weight_contribution_variable_e_rate(7, probs, k) == pytest.approx(brute_variable(7, probs, k), rel=1e-12, abs=1e-18)
    assert sum(isamp.weight_contribution_variable_e_rate(7, probs, k) for k in range(8)) == pytest.approx(1.0, rel=1e-13)
    # constant list == binomial
    assert isamp.weight_contribution_variable_e_rate(24, 24 * [5e-3], 2) == pytest.approx(isamp.subset_weight_mixed([(24, [5e-3], 2)]), rel=1e-13)
    assert isamp.subset_weight_variable_e_rate([(7, probs, 2), (3, [0.1, 0.2, 0.3], 1)]) == \
        pytest.approx(brute_variable(7, probs, 2) * brute_variable(3, [0.1, 0.2, 0.3], 1), rel=1e-12)


def test_significant_subset_enumeration_matches_reference_order_and_weights():
    g = json.load(open(os.path.join(GOLDEN, "subsets_small.json")))
    s1 = isamp.calculate_significant_subsets_incrementally(g["locations"], g["pl1"], cutoff_weight=g["cutoff"])
    assert s1["error_subset_list"] == g["s1"]["error_subset_list"]
    assert s1["subset_weights"] == pytest.approx(g["s1"]["subset_weights"], rel=1e-12)
    both = [2 * x for x in g["locations"]]
    s12 = isamp.calculate_significant_subsets_incrementally(both, g["plb"], cutoff_weight=g["cutoff"])
    assert s12["error_subset_list"] == g["s12"]["error_subset_list"]
    assert s12["subset_weights"] == pytest.approx(g["s12"]["subset_weights"], rel=1e-12)
    assert set(s1) == set(g["s1"])          # same JSON schema keys (CL:602-606)
    more = isamp.find_significant_subsets_with_more_error([0, 0, 1, 0, 0], g["s12"]["error_subset_list"])
    assert [list(x) for x in more[0]] == g["more"][0] and [list(x) for x in more[1]] == g["more"][1]
    s2 = isamp.s2sim_calculate_significant_subsets_incrementally(g["locations"], g["pl1"], g["pl2"], g["s2_rates"],
                                                                 g["s1"]["error_subset_list"],
                                                                 g["s12"]["error_subset_list"], cutoff_weight=g["cutoff"])
    assert s2["error_subset_list"] == g["s2"]["error_subset_list"]
    assert s2["subset_weights"] == pytest.approx(g["s2"]["subset_weights"], rel=1e-12)
    assert set(s2) == set(g["s2"])          # CL:742-747
    wl = isamp.weighted_logical_error_rate([0.05, 0.02, 0.01, 0.03, 0.04], g["s1"]["error_subset_list"], g["s2_rates"],
                                           g["locations"])
    assert wl[0] == pytest.approx(g["weighted"][0], rel=1e-12)
    assert wl[1] == pytest.approx(g["weighted"][1], rel=1e-12)


def test_pl_assembly():
    g = json.load(open(os.path.join(GOLDEN, "subsets_small.json")))
    s1, s2 = g["s1"], g["s2"]
    s1_rates = {str(e): (None if sum(e) == 0 else 0.01 * sum(e)) for e in s1["error_subset_list"]}
    fails = {str(e): 3 for e in s2["error_subset_list"]}
    counts = {str(e): 30 for e in s2["error_subset_list"]}
    p_file = isamp.pl_from_file(s1["error_subset_list"], s1["subset_weights"],
                                {k: (-1 if v is None else v) for k, v in s1_rates.items()},
                                s2["error_subset_list"], s2["subset_weights"], fails, counts)
    expect = sum(0.01 * sum(e) * w for e, w in zip(s1["error_subset_list"], s1["subset_weights"]) if sum(e)) + \
        sum(0.1 * w for w in s2["subset_weights"])
    assert p_file == pytest.approx(expect, rel=1e-12)
    # recomputing the weights at the same physical rates reproduces the stored ones
    p_new = isamp.pl(g["pl1"], g["pl2"], s1["error_subset_list"], s1_rates, s2["error_subset_list"], fails, counts,
                     g["s2_rates"], g["s12"]["error_subset_list"], g["locations"])
    assert p_new == pytest.approx(p_file, rel=1e-10)


def test_lookup_table_generation_properties():
    table = glt.generate_correction_table()
    assert len(table) == 256
    hist = {}
    T = glt.parity_check_matrix()
    for key, (vec, w) in table.items():
        hist[w] = hist.get(w, 0) + 1
        synd = [int(sum(T[r][c] * vec[c] for c in range(18)) % 2) for r in range(8)]
        assert " ".join(map(str, synd)) == key
    assert hist == {0: 1, 1: 23, 2: 120, 3: 96, 4: 16}
    assert len(glt.generate_classical_bitflip_table()) == 16
    # and the generated tables ARE the shipped ones (the reference's files, byte for byte in the JSON format it writes)
    data = os.path.join(os.path.dirname(GOLDEN), os.pardir, "surface_17_iqt_b200", "data")
    shipped = json.load(open(os.path.join(data, "correction_table_depolarising.json")))
    assert {k: [list(v[0]), v[1]] for k, v in table.items()} == {k: [list(v[0]), v[1]] for k, v in shipped.items()}
    shipped_c = json.load(open(os.path.join(data, "correction_table_classical_bitflip.json")))
    mine_c = glt.generate_classical_bitflip_table()
    assert {k: [list(v[0]), v[1]] for k, v in mine_c.items()} == {k: [list(v[0]), v[1]] for k, v in shipped_c.items()}


def test_run_til_fail_fit_equals_the_reference_fit():
    """fit_log_e_rate (histogram + scipy curve_fit of A exp(-r t), CL:546-565 without the plotting) on the seeded
    rounds-til-fail lists of the reference's own driver (tests/golden/rtf_ref.json, minted by oracle/make_golden_rtf.py:
    `reference_fit` is what the unmodified calculate_log_e_rate returned for the same list and num_bins = 5): equal to
    1e-12 relative; where the reference's fit did not converge, ours raises too."""
    import warnings
    from surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model import fit_log_e_rate
    cases = json.load(open(os.path.join(GOLDEN, "rtf_ref.json")))["cases"]
    checked = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for case in cases:
            for chain in case["chains"]:
                if chain["reference_fit"] is None:
                    with pytest.raises(RuntimeError):
                        fit_log_e_rate(chain, 5)
                    continue
                got = fit_log_e_rate(chain, 5)[0]
                assert got == pytest.approx(chain["reference_fit"], rel=1e-12)
                checked += 1
    assert checked >= 12
