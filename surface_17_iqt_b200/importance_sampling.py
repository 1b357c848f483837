"""Importance-sampling bookkeeping on the host: subset weights, significant-subset enumeration and the
assembly of the logical error rate from the counters the device path returns.

Same names, arguments and return values as the reference (calc_log_e_rate_arbitrary_error_model.py [CL]
:460-525, 528-543, 567-607, 635-753; plotting_imp_samp.py [PL]:56-98), re-implemented:

* location-dependent weights use the elementary-symmetric-polynomial recurrence (O(L k)) instead of the
  reference's enumeration of all C(L, k) placements (CL:468-487) -- same value to rounding;
* the enumerations keep the reference's traversal semantics (a list that is mutated while it is iterated,
  CL:576-578, 658-660), so the ORDER of the returned subset lists matches the reference's JSON artefacts.
"""
from math import comb


def binomial_weight(num_gates, p, num_error):
    return comb(num_gates, num_error) * p ** num_error * (1 - p) ** (num_gates - num_error)


def subset_weight(params):
    """CL:460-466: product of binomial weights; params = [(num_gates, prob_error, num_error), ...]."""
    weight = 1
    for num_gates, prob_error, num_error in params:
        weight *= binomial_weight(num_gates, prob_error, num_error)
    return weight


def weight_contribution_variable_e_rate(num_gates, prob_error_list, num_error):
    """CL:468-487: P(exactly num_error of the first num_gates locations fire), one probability per location."""
    if num_error < 0 or num_error > num_gates:
        return 0
    dp = [1.0] + num_error * [0.0]                   # dp[j] = P(j errors among the locations seen so far)
    for i in range(num_gates):
        p = prob_error_list[i]
        for j in range(min(i + 1, num_error), 0, -1):
            dp[j] = dp[j] * (1 - p) + dp[j - 1] * p
        dp[0] *= (1 - p)
    return dp[num_error]


def subset_weight_mixed(params):
    """CL:489-500: a one-element probability list means a constant rate, a longer one a rate per location."""
    weight = 1
    for num_gates, prob_error, num_error in params:
        if len(prob_error) == 1:
            weight *= binomial_weight(num_gates, prob_error[0], num_error)
        else:
            weight *= weight_contribution_variable_e_rate(num_gates, prob_error, num_error)
    return weight


def subset_weight_variable_e_rate(params):
    """CL:502-525."""
    weight = 1
    for num_gates, prob_error_list, num_error in params:
        weight *= weight_contribution_variable_e_rate(num_gates, prob_error_list, num_error)
    return weight


def weighted_logical_error_rate(probs, esets, log_e_dict, error_locations, variable_e_rate=False):
    """CL:528-543."""
    sum_log_e, weights, subset_e_rates = 0, [], []
    for eset in esets:
        params = [(error_locations[ind], probs[ind], eset[ind]) for ind in range(len(eset))]
        weight = subset_weight_variable_e_rate(params) if variable_e_rate else subset_weight(params)
        log_e_subset = weight * log_e_dict[str(eset)]
        sum_log_e += log_e_subset
        weights.append(weight)
        subset_e_rates.append(log_e_subset)
    return sum_log_e, weights, subset_e_rates


def increment(eset_a):
    """CL:635-639: the subsets with one more error of one type."""
    out = []
    for j in range(len(eset_a)):
        e = list(eset_a)
        e[j] += 1
        out.append(e)
    return out


def _drain(work, visit):
    """The reference's `while len(l): for x in l: l.remove(x) ...` traversal: a list iterator advances its index
    after every step while `remove` deletes the first equal element, so every other element is skipped per sweep."""
    while len(work) != 0:
        i = 0
        while i < len(work):
            x = work[i]
            work.remove(x)
            visit(x)
            i += 1


def calculate_significant_subsets_incrementally(locations, prob_lists, cutoff_weight=1e-6, display_checked=False):
    """CL:567-607: grow error subsets from the all-zero one while their weight exceeds the cut-off."""
    error_subset_list, subset_weight_list, already_checked = [], [], []
    esets = [len(locations) * [0]]

    def visit(eset):
        if eset in already_checked:
            return
        params = [(locations[i], prob_lists[i], eset[i]) for i in range(len(eset))]
        w = subset_weight_mixed(params)
        if w > cutoff_weight:
            error_subset_list.append(eset)
            subset_weight_list.append(w)
            esets.extend(increment(eset))
        already_checked.append(eset)

    _drain(esets, visit)
    if display_checked:
        print('checked')
        print(already_checked)
        print(len(already_checked))
    return {'error_subset_list': error_subset_list, 'subset_weights': subset_weight_list,
            'cutoff_weight': cutoff_weight, 'locations': locations, 'probabilities': prob_lists}


def find_significant_subsets_with_more_error(eset_a, error_subsets_list):
    """CL:642-675: the listed subsets reachable from eset_a by adding errors, and the added errors."""
    s2_esets, s2_subtracted = [], []
    work = increment(eset_a)
    work.append(eset_a)
    known = {tuple(e) for e in error_subsets_list}

    def visit(new_eset):
        if new_eset in s2_esets:
            return
        if tuple(new_eset) in known:
            s2_esets.append(new_eset)
            s2_subtracted.append([new_eset[i] - eset_a[i] for i in range(len(new_eset))])
            work.extend(increment(new_eset))

    _drain(work, visit)
    return s2_esets, s2_subtracted


def prob_s2_AND_eset_a_total_errors(eset_a, s2_rate_dict, significant_subsets_s1, significant_subsets_s1_and_s2,
                                    locations1round, prob_lists_s1, prob_lists_s2, _reach_cache=None):
    """CL:678-697: P(second syndrome needed AND eset_a errors in total over both rounds).
    `_reach_cache` memoises find_significant_subsets_with_more_error per s1 subset (it does not depend on eset_a)."""
    sump, ratelist, sw1list, sw2list = 0, [], [], []
    for eset_s1 in significant_subsets_s1:
        key = tuple(eset_s1)
        if _reach_cache is not None and key in _reach_cache:
            s2, s2_subtracted = _reach_cache[key]
        else:
            s2, s2_subtracted = find_significant_subsets_with_more_error(eset_s1, significant_subsets_s1_and_s2)
            if _reach_cache is not None:
                _reach_cache[key] = (s2, s2_subtracted)
        if eset_a in s2:
            eset_subtracted = s2_subtracted[s2.index(eset_a)]
            params1 = [(locations1round[i], prob_lists_s1[i], eset_s1[i]) for i in range(len(eset_s1))]
            params2 = [(locations1round[i], prob_lists_s2[i], eset_subtracted[i]) for i in range(len(eset_subtracted))]
            sw1, sw2 = subset_weight_mixed(params1), subset_weight_mixed(params2)
            rate = s2_rate_dict[str(eset_s1)]
            sump += rate * sw1 * sw2
            sw1list.append(sw1)
            sw2list.append(sw2)
            ratelist.append(rate)
    return sump, ratelist, sw1list, sw2list


def s2sim_calculate_significant_subsets_incrementally(locations, prob_lists_s1, prob_lists_s2, s2_rate_dict,
                                                      significant_subsets_s1, significant_subsets_s1_and_s2,
                                                      cutoff_weight=1e-6, display_checked=False):
    """CL:700-753: significant subsets of the two-round (s2) simulations."""
    error_subset_list, subset_weight_list, already_checked = [], [], []
    zero = len(locations) * [0]
    esets = increment(zero)
    esets.append(zero)
    reach = {}

    def visit(eset):
        if eset in already_checked:
            return
        weight = prob_s2_AND_eset_a_total_errors(eset, s2_rate_dict, significant_subsets_s1,
                                                 significant_subsets_s1_and_s2, locations, prob_lists_s1,
                                                 prob_lists_s2, _reach_cache=reach)[0]
        if weight > cutoff_weight:
            error_subset_list.append(eset)
            subset_weight_list.append(weight)
            esets.extend(increment(eset))
        already_checked.append(eset)

    _drain(esets, visit)
    if display_checked:
        print('checked')
        print(already_checked)
        print(len(already_checked))
    return {'error_subset_list': error_subset_list, 'subset_weights': subset_weight_list,
            'cutoff_weight': cutoff_weight, 'locations': locations,
            'error probabilities s1': prob_lists_s1, 'error probabilities s2': prob_lists_s2}


def pl_from_file(s1_significant_subsets, s1_weights, s1_log_e_rates, s2_significant_subsets, s2_weights, s2_fails,
                 s2_counts):
    """PL:56-73: P_L from stored subset weights and the s1 / s2 counters (the README formula)."""
    p = 0
    for i, eset in enumerate(s1_significant_subsets):
        erate = s1_log_e_rates[str(eset)]
        if erate == -1 or erate is None:
            continue
        p += erate * s1_weights[i]
    for i, eset in enumerate(s2_significant_subsets):
        if s2_counts[str(eset)] == 0:
            continue
        p += s2_fails[str(eset)] / s2_counts[str(eset)] * s2_weights[i]
    return p


def pl(prob_lists_s1, prob_lists_s2, s1_significant_subsets, s1_log_e_rates, s2_significant_subsets, s2_fails,
       s2_counts, s2_rate_dict, s1_and_s2_significant_subsets, locations_1round):
    """PL:76-98: P_L with the subset weights recomputed for new physical error rates.  Like the reference the s1
    terms are weighted by P(subset) rather than P(subset AND stop after one syndrome)."""
    p = 0
    for eset in s1_significant_subsets:
        erate = s1_log_e_rates[str(eset)]
        if erate is None or erate == -1:
            continue
        params = [(locations_1round[i], prob_lists_s1[i], eset[i]) for i in range(len(eset))]
        p += erate * subset_weight_mixed(params)
    for eset in s2_significant_subsets:
        if s2_counts[str(eset)] == 0:
            continue
        erate = s2_fails[str(eset)] / s2_counts[str(eset)]
        weight = prob_s2_AND_eset_a_total_errors(eset, s2_rate_dict, s1_significant_subsets,
                                                 s1_and_s2_significant_subsets, locations_1round, prob_lists_s1,
                                                 prob_lists_s2)[0]
        p += erate * weight
    return p
