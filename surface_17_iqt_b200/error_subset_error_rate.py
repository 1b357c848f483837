"""Object-oriented driver -- ``calc_error_subset_error_rate`` of the reference's error_subset_error_rate.py:12-54
(50 runs of: place the subset's errors at random OO locations, one noisy cycle, decode, read out), batched.
The reference version cannot run (it calls Error.generate_error_list with one argument, SURVEY C-14); the semantics
are those of the single-round protocol (calc_log_e_rate_arbitrary_error_model.py:183-259).  Nothing executes on import.
"""
import time

from .error import Error, create_error_subset_list  # noqa: F401
from .logical_qubit import LogicalQubit, StabiliserCycle, pdd_error_model

NUM_RUNS = 50


def calc_error_subset_error_rate(error_subset, num_runs=NUM_RUNS, logical_qubit=None, verbose=True):
    t_begin = time.perf_counter()
    cycle = StabiliserCycle(location=[1], error_model=pdd_error_model)
    mask_sets = []
    for _ in range(num_runs):
        error_list = Error.generate_error_list(error_subset, cycle.error_locations)
        mask_sets.append(cycle.slot_masks(error_list))
    lq = logical_qubit or LogicalQubit(max_shots=max(num_runs, 1))
    counts = lq.measure_syndromes(mask_sets)
    if logical_qubit is None:
        lq.close()
    log_e_rate = counts.failed / num_runs
    if verbose:
        print(log_e_rate)
        print('total_time_taken_{}_runs_{}'.format(num_runs, time.perf_counter() - t_begin))
    return log_e_rate


def calc_many_error_subset_error_rates(error_subset_list, num_runs=NUM_RUNS):
    log_e_rate_dict = {}
    for error_subset in error_subset_list:
        log_e_rate_dict[error_subset["DephasingError"]] = calc_error_subset_error_rate(error_subset, num_runs)
    return log_e_rate_dict
