"""Drop-in Monte-Carlo drivers with the reference's names, positional signatures and return values
(calc_log_e_rate_arbitrary_error_model.py [CL]), executing whole batches of shots on the B200 backend.

    s1_calculate_log_e_rate_error_subset   CL:261-353
    s2_calculate_log_e_rate_error_subset   CL:355-457
    calculate_log_e_rate_error_subset      CL:183-259
    calculate_log_e_rate                   CL:33-148   (run-til-fail)
    generate_error_locations[_non_uniform] CL:151-181  (host versions, same contract)

Placement and error draws use device Philox streams keyed by (seed, global shot id); the seed of each call
is taken from Python's global ``random`` (``random.getrandbits(64)``), so ``random.seed(k)`` makes a run
reproducible just as it does for the reference.  When torch.distributed is initialised the shot range is
split over the ranks and the counters are all-reduced (distributed.py).
"""
import json
import os
import random
import time

import numpy as np

from . import _lib as L
from . import api, circuit, distributed

# knobs (module-level like the reference's scripts; no CLI)
# shots resident per device context: 8 KiB each when every cycle runs on the whole-cycle kernel, 2 MiB each otherwise
# (the first out-of-memory error of a context shrinks the batch and retries, see _with_batch_retry)
BATCH_SHOTS = int(os.environ.get("S17_BATCH", "262144"))
FUSION = int(os.environ.get("S17_FUSION", "2"))           # timesteps per sweep: 0 = one gate per sweep, 1, 2 or 4
DEVICE = None                                             # None: LOCAL_RANK or 0
VERBOSE = True                                            # print the reference's summary lines

_SIMS = {}
_BATCH_OF = {}          # program key -> batch that fits (set when BATCH_SHOTS full-size states did not)
LAST_TIMING = {}
LAST_RESULTS = {}       # calculate_log_e_rate, rank 0: the per-run arrays of the last call (numpy; the JSON holds them as lists)


def _log(*a):
    if VERBOSE:
        print(*a)


def _device():
    return int(os.environ.get("LOCAL_RANK", "0")) if DEVICE is None else int(DEVICE)


def _simulation(model, list_len, cooling, correction_table, bitflip_table, over_angles=None):
    pkey = (model, tuple(list_len), bool(cooling), FUSION,
            None if over_angles is None else json.dumps(over_angles, sort_keys=True))
    batch = _BATCH_OF.get(pkey, BATCH_SHOTS)
    key = (model, tuple(list_len), bool(cooling), FUSION, _device(), batch, pkey[4])
    sim = _SIMS.get(key)
    if sim is None:
        for k in list(_SIMS):            # one resident context per device: free the previous program's state
            if k[4] == _device():
                _SIMS.pop(k).close()
        sim = api.Simulation(model, list_len, batch, correction_table, bitflip_table, fusion=FUSION,
                             cooling=cooling, device=_device(), over_angles=over_angles)
        _SIMS[key] = sim
        sim._tables = (correction_table, bitflip_table)
        sim._pkey = pkey
    elif sim._tables[0] is not correction_table or sim._tables[1] is not bitflip_table:
        sim.backend.load_tables(correction_table, bitflip_table)      # other table objects than last time: re-pack
        sim._tables = (correction_table, bitflip_table)
    return sim


def _with_batch_retry(fn):
    """Run fn(); if the device cannot hold a batch of state vectors of the size this program needs (2 MiB per shot on
    the general amplitude kernels instead of 8 KiB), drop the context, cut the batch of THIS program to a sixteenth and
    try again."""
    while True:
        try:
            return fn()
        except L.S17Error as exc:
            if "out of device memory" not in str(exc):
                raise
            shrunk = False
            for k in list(_SIMS):
                sim = _SIMS.pop(k)
                if sim.max_shots > 256:
                    _BATCH_OF[sim._pkey] = max(256, sim.max_shots // 16)
                    shrunk = True
                    _log("device memory: batch of {} reduced to {} shots".format(sim.model_name, _BATCH_OF[sim._pkey]))
                sim.close()
            if not shrunk:
                raise


def release():
    """Free every cached device context."""
    for k in list(_SIMS):
        _SIMS.pop(k).close()


def generate_error_locations(num_e, num_e_locations):
    """CL:151-161 (host version): uniform fixed-weight 0/1 mask.  Raises instead of looping forever if k > L."""
    if num_e > num_e_locations:
        raise ValueError("cannot place {} errors on {} locations".format(num_e, num_e_locations))
    mask = num_e_locations * [0]
    for ind in random.sample(range(num_e_locations), num_e):
        mask[ind] = 1
    return mask


def generate_from_distr(probs):
    """CL:163-169."""
    r = random.random() * sum(probs)
    total = 0
    for i, p in enumerate(probs):
        total += p
        if total > r:
            return i
    return len(probs) - 1


def generate_error_locations_non_uniform(num_e, num_e_locations, probs):
    """CL:171-181: sequential sampling proportional to probs, rejecting repeats."""
    if num_e > sum(1 for p in probs if p > 0):
        raise ValueError("more errors than locations with non-zero weight")
    mask = num_e_locations * [0]
    for _ in range(num_e):
        while True:
            ind = generate_from_distr(probs)
            if mask[ind] == 0:
                mask[ind] = 1
                break
    return mask


def _run_subset(*args):
    return _with_batch_retry(lambda: _run_subset_once(*args))


def subset_sweep(protocol, num_runs, correction_table, model, esets, locations, bitflip_table, prob_lists,
                 cooling=False):
    """Every error subset of `esets` through the loop of the s1 / s2 / single driver (CL:261-353, 355-457, 183-259) with
    ONE exchange for the whole sweep: the ranks' int64[n_subsets x 4] counter array (failed, not0, counted, errors) is
    all-reduced once at the end instead of once per subset.  Returns one (failed, not0, counted) per subset;
    the rates are the callers' (failed / (runs - not0) for s1, failed / not0 for s2, failed / runs for single)."""
    local = []
    for eset in esets:
        local.extend(_with_batch_retry(lambda e=eset: _run_subset_once(
            protocol, num_runs, correction_table, model, list(e), locations, bitflip_table, prob_lists, False, cooling,
            False)))
    total = distributed.allreduce_counts(local)
    out = []
    for k in range(len(esets)):
        failed, not0, counted, errors = total[4 * k:4 * k + 4]
        if errors:
            raise L.S17Error("{} shots of subset {} hit a zero-probability outcome".format(errors, list(esets[k])))
        out.append((failed, not0, counted))
    return out


def _run_subset_once(protocol, num_runs, correction_table, model, eset, locations, bitflip_table, prob_lists, cz_comp,
                     cooling, reduce=True):
    if cz_comp:
        raise NotImplementedError("the experimental CZ compilation is outside the accelerated path (SURVEY C-9)")
    if len(eset) != len(locations) or len(prob_lists) != len(eset):
        raise ValueError("eset, locations and prob_lists must have one entry per error type")
    sim = _simulation(model, list(locations), cooling, correction_table, bitflip_table)
    weights = [None if len(p) == 1 else list(p) for p in prob_lists]     # CL:287: one entry => uniform placement
    sim.backend.set_subset(eset, locations, weights)
    seed = distributed.broadcast_seed(random.getrandbits(64))
    rank, world_size = distributed.world()
    first, count = distributed.shard_range(num_runs, rank, world_size)
    failed = not0 = counted = errors = 0
    gate_ms = span_ms = 0.0
    sweeps = launches = sweep_bytes = 0
    done = 0
    while done < count:
        n = min(sim.max_shots, count - done)
        c = sim.backend.run(protocol, n, L.NOISE_SUBSET, seed=seed, first_shot_id=first + done)
        t = sim.backend.timing()
        failed, not0, counted, errors = failed + c.failed, not0 + c.not0, counted + c.counted, errors + c.errors
        gate_ms, span_ms = gate_ms + t["gate_ms"], span_ms + t["span_ms"]
        sweeps, launches = sweeps + c.sweeps, launches + t["launches"]
        sweep_bytes += c.sweep_bytes
        done += n
    if not reduce:                       # subset_sweep: the caller reduces the counters of the whole sweep at once
        LAST_TIMING.update(gate_ms=gate_ms, span_ms=span_ms, sweeps=sweeps, sweep_bytes=sweep_bytes, launches=launches,
                           shots=count)
        return [failed, not0, counted, errors]
    # the error count rides in the reduced vector: every rank leaves the collective before any of them raises
    failed, not0, counted, errors = distributed.allreduce_counts([failed, not0, counted, errors])
    if errors:
        raise L.S17Error("{} shots hit a zero-probability outcome".format(errors))
    LAST_TIMING.update(gate_ms=gate_ms, span_ms=span_ms, sweeps=sweeps, sweep_bytes=sweep_bytes, launches=launches,
                       shots=count)
    return failed, not0, counted


def calculate_log_e_rate_error_subset(num_runs, correction_table, model, eset, locations, bitflip_table,
                                      prob_lists, cz_comp=False):
    """CL:183-259: single round, always decode flips_a.  Returns failed_count / num_runs.

    Deviation: the reference chooses non-uniform placement for EVERY error type as soon as prob_lists[0] is a list
    (CL:207-214), which for a one-element list always places the error on location 0; here a one-element list means
    uniform placement, as in the s1 / s2 drivers (CL:287)."""
    t_begin = time.perf_counter()
    if not isinstance(prob_lists[0], list):      # CL:207: a flat list of floats means uniform placement
        prob_lists = [[p] for p in prob_lists]
    failed, _, _ = _run_subset("single", num_runs, correction_table, model, eset, locations, bitflip_table,
                               prob_lists, cz_comp, False)
    log_e_rate = failed / num_runs
    _log(log_e_rate)
    _log('total_time_taken_{}_runs_{}'.format(num_runs, time.perf_counter() - t_begin))
    return log_e_rate


def s1_calculate_log_e_rate_error_subset(num_runs, correction_table, model, eset, locations, bitflip_table,
                                         prob_lists, cz_comp=False, cooling=False):
    """CL:261-353.  Returns (failed/(runs-not0) or None, not0/runs, failed_count, flipsa_not0_count)."""
    t_begin = time.perf_counter()
    failed_count, flipsa_not0_count, _ = _run_subset("s1", num_runs, correction_table, model, eset, locations,
                                                     bitflip_table, prob_lists, cz_comp, cooling)
    try:
        log_e_rate = failed_count / (num_runs - flipsa_not0_count)
    except ZeroDivisionError:
        _log('cant assign error rate, 0 runs ending after first stab, placeholder e rate -1')
        log_e_rate = None                        # CL:343 (the message says -1, the value is None)
    flipsa_not0_rate = flipsa_not0_count / num_runs
    _summary(failed_count, flipsa_not0_count, log_e_rate, flipsa_not0_rate, num_runs, t_begin)
    return log_e_rate, flipsa_not0_rate, failed_count, flipsa_not0_count


def s2_calculate_log_e_rate_error_subset(num_runs, correction_table, model, eset, locations, bitflip_table,
                                         prob_lists, cz_comp=False, cooling=False):
    """CL:355-457.  `locations` / masks span two rounds.  Returns (failed/not0 or -1, not0/runs, failed, not0)."""
    t_begin = time.perf_counter()
    failed_count, flipsa_not0_count, _ = _run_subset("s2", num_runs, correction_table, model, eset, locations,
                                                     bitflip_table, prob_lists, cz_comp, cooling)
    try:
        log_e_rate = failed_count / flipsa_not0_count
    except ZeroDivisionError:
        _log('cant assign error rate, 0 runs ending after first stab, placeholder e rate -1')
        log_e_rate = -1                          # CL:447
    flipsa_not0_rate = flipsa_not0_count / num_runs
    _summary(failed_count, flipsa_not0_count, log_e_rate, flipsa_not0_rate, num_runs, t_begin)
    return log_e_rate, flipsa_not0_rate, failed_count, flipsa_not0_count


def _summary(failed_count, flipsa_not0_count, log_e_rate, flipsa_not0_rate, num_runs, t_begin):
    _log('counts')
    _log('failed {} times'.format(failed_count))
    _log('wouldnt stop at s1 {} '.format(flipsa_not0_count))
    _log('rates')
    _log(log_e_rate)
    _log(flipsa_not0_rate)
    _log('total_time_taken_{}_runs_{}'.format(num_runs, time.perf_counter() - t_begin))


def calculate_log_e_rate_true_probability(num_runs, correction_table, model, p_set, bitflip_table, protocol="single",
                                          cooling=False, over_angles=None):
    """True Bernoulli-per-slot noise with a one-round (or s1 / s2) protocol: the shot loop of CL:183-259 with the
    probability lists bound directly (instantiate_error_model, CS:31-49) instead of 0/1 masks.  This is BASELINE
    config 1 (`depolarising_model`, px = py = pz = p/3 on the 30 single-qubit slots, `depol` p on the 24 Rxx slots;
    slot indexing per SURVEY App. C-1) and config 3 (`coherent_model`, angles via `over_angles`).
    Returns (failed_count / counted, failed_count, counted, flipsa_not0_count)."""
    return _with_batch_retry(lambda: _true_probability_once(num_runs, correction_table, model, p_set, bitflip_table,
                                                            protocol, cooling, over_angles))


def _true_probability_once(num_runs, correction_table, model, p_set, bitflip_table, protocol, cooling, over_angles):
    p_set = [list(p) for p in p_set]
    sim = _simulation(model, [len(p) for p in p_set], cooling, correction_table, bitflip_table, over_angles=over_angles)
    names = sim.model["probabilities"]
    sim.backend.set_slot_probs([[0.0] * len(p) if (over_angles and names[i] in over_angles) else p
                                for i, p in enumerate(p_set)])
    seed = distributed.broadcast_seed(random.getrandbits(64))
    rank, world_size = distributed.world()
    first, count = distributed.shard_range(num_runs, rank, world_size)
    failed = not0 = counted = errors = done = 0
    while done < count:
        n = min(sim.max_shots, count - done)
        c = sim.backend.run(protocol, n, L.NOISE_PROB, seed=seed, first_shot_id=first + done)
        failed, not0, counted, done = failed + c.failed, not0 + c.not0, counted + c.counted, done + n
        errors += c.errors
    failed, not0, counted, errors = distributed.allreduce_counts([failed, not0, counted, errors])
    if errors:
        raise L.S17Error("{} shots hit a zero-probability outcome".format(errors))
    return (failed / counted if counted else None), failed, counted, not0


def f(t, A, r):
    """CL:756."""
    return A * np.exp(-r * t)


def fit_log_e_rate(results, num_bins, include_b=False):
    """The histogram + exponential fit of plot_log_e_rate_graph (CL:546-565) without the plotting."""
    from scipy.optimize import curve_fit
    data = np.array(results['rounds_til_fail_list'], dtype='float64')
    if include_b:
        data = data + np.array(results['syndrome_b_measured_list'], dtype='float64')
    vals, bins = np.histogram(data, bins=num_bins)
    popt, _pcov = curve_fit(f, bins[:-1], vals.astype('float64'))
    return popt[1], popt[0], vals, bins


def calculate_log_e_rate(num_runs, filename, correction_table, e_model, e_probs, bitflip_table, num_bins=20,
                         cz_compilation=False, model_name=None, save=True, leak_reset="run", max_rounds=0,
                         round_budget=0, fit=True):
    """CL:33-148: run-til-fail with true Bernoulli noise, JSON dump and exponential fit; returns log_e_rate.

    `e_model` is the model dict returned by instantiate_error_model (or its name), `e_probs` the {name: list} dict.
    leak_reset: when the leak register is cleared -- 'run' (default; documented deviation, SURVEY C-6), 'round', or
    'never', the reference's literal behaviour (CL:49: one register for the whole call; here one per device slot, a slot
    being one sequential chain of runs).
    Under torch.distributed the runs are sharded over the ranks (each rank's slots are independent chains), the lists
    are concatenated in rank order, and rank 0 alone writes data/<filename>.json.
    max_rounds / round_budget (extensions): cap the rounds of one run / the total number of rounds of the call; runs the
    caps cut short are reported with their censored length in 'censored_runs' and excluded from the fit.
    """
    if cz_compilation:
        raise NotImplementedError("the experimental CZ compilation is outside the accelerated path (SURVEY C-9)")
    models = api.load_error_models()
    if isinstance(e_model, str):
        model_name, e_model = e_model, models[e_model]
    if model_name is None:
        model_name = next((k for k, v in models.items() if v == e_model), None)
        if model_name is None:
            raise ValueError("e_model is not one of the models in error_models.json; pass model_name= for a custom model "
                             "that the file in the working directory defines")
    names = e_model['probabilities']
    p_set = [list(e_probs[k]) for k in names]
    need = circuit.list_lengths(e_model)
    for k, plist, n in zip(names, p_set, need):
        if len(plist) < n:
            raise IndexError("probability list {} has {} entries, the cycle indexes {}".format(k, len(plist), n))
    seed = distributed.broadcast_seed(random.getrandbits(64))
    rank, world_size = distributed.world()
    _first, my_runs = distributed.shard_range(num_runs, rank, world_size)
    t_begin = time.perf_counter()

    def run_once():
        sim_ = _simulation(model_name, [len(p) for p in p_set], False, correction_table, bitflip_table)
        sim_.backend.set_slot_probs(p_set)
        if my_runs == 0:
            return sim_, None
        slots = min(sim_.max_shots, my_runs)
        # the Philox counter is (first_shot_id + slot): rank r's slots continue where rank r-1's end
        return sim_, sim_.backend.run("rtf", slots, L.NOISE_PROB, seed=seed, first_shot_id=rank * sim_.max_shots,
                                      rtf_runs=my_runs, leak_reset=leak_reset, rtf_max_rounds=max_rounds,
                                      rtf_round_budget=round_budget)

    sim, c = _with_batch_retry(run_once)
    if c is not None:
        rounds, syndb = sim.backend.rtf_results(my_runs)
        my_rounds, my_syndb = np.asarray(rounds, dtype=np.int64), np.asarray(syndb, dtype=np.int64)
        my_total, my_errors = int(c.rounds), int(c.errors)
        span_ms = sim.backend.timing()["span_ms"]
    else:
        my_rounds, my_syndb, my_total, my_errors, span_ms = np.zeros(0, np.int64), np.zeros(0, np.int64), 0, 0, 0.0
    # the per-run lists go to rank 0 only (it writes the file and fits); every rank gets the fit and the totals
    rounds_arr = distributed.gather_to_root(my_rounds)
    syndb_arr = distributed.gather_to_root(my_syndb)
    total_rounds, errors = distributed.allreduce_counts([my_total, my_errors])
    if errors:
        raise L.S17Error("{} rounds hit a zero-probability outcome".format(errors))
    total_time_taken = time.perf_counter() - t_begin
    _log('total time taken {}'.format(total_time_taken))
    LAST_TIMING.update(rounds=int(total_rounds), span_ms=span_ms)
    log_e_rate, fit_error = None, ""
    if rank == 0:
        cens = rounds_arr == 0                                                # cut short by round_budget
        if max_rounds:
            cens |= rounds_arr >= max_rounds
        LAST_RESULTS.clear()
        LAST_RESULTS.update(rounds_til_fail=rounds_arr, syndrome_b_measured=syndb_arr, censored=cens,
                            total_rounds=int(total_rounds))
        if save:
            round_count_list, syndrome_b_list = rounds_arr.tolist(), syndb_arr.tolist()
            results = {
                'total_time_taken_{}_runs'.format(num_runs): total_time_taken,
                'probability_set': e_probs,
                'rounds_til_fail_list': round_count_list,
                'syndrome_b_measured_list': syndrome_b_list,
                'rounds_til_fail_list_X0': [], 'rounds_til_fail_list_X1': [],
                'rounds_til_fail_list_Z0': round_count_list, 'rounds_til_fail_list_Z1': [],
                'error_model': e_model,
            }
            if cens.any():
                results['censored_runs'] = np.nonzero(cens)[0].tolist()
            results['total_rounds'] = int(total_rounds)
            os.makedirs("data", exist_ok=True)
            with open(os.path.join("data", filename + '.json'), 'w') as file:
                json.dump(results, file)
        if fit:
            done = {'rounds_til_fail_list': rounds_arr[~cens], 'syndrome_b_measured_list': syndb_arr[~cens]}
            try:
                log_e_rate = float(fit_log_e_rate(done, num_bins)[0])
            except RuntimeError as exc:            # scipy: "Optimal parameters not found" -- every rank must hear of it
                log_e_rate, fit_error = float('nan'), str(exc)
    if not fit:
        return None
    log_e_rate = distributed.broadcast_float(log_e_rate)
    if log_e_rate != log_e_rate:
        raise RuntimeError("exponential fit of the rounds-til-fail histogram failed" + (": " + fit_error if fit_error else ""))
    return log_e_rate


# the host-side importance-sampling bookkeeping of the reference module (CL:460-753), same names
from .importance_sampling import (  # noqa: E402,F401
    subset_weight, weight_contribution_variable_e_rate, subset_weight_mixed, subset_weight_variable_e_rate,
    weighted_logical_error_rate, increment, calculate_significant_subsets_incrementally,
    find_significant_subsets_with_more_error, prob_s2_AND_eset_a_total_errors,
    s2sim_calculate_significant_subsets_incrementally)
from .compiled_surface_code_arbitrary_error_model import (  # noqa: E402,F401
    load_lookup_table, instantiate_error_model, lookup, return_weight_classical_lookup)
