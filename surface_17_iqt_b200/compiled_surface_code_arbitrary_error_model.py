"""Host-side mirror of the reference module of the same name (the parts the batched path keeps).

The gate-by-gate functions of the reference (stabiliser_timestep_*, insert_error, ...) have no
per-call equivalent here: they are compiled once by ``circuit.compile_program`` and executed for a
whole batch on device.  What stays on the host is the model / table plumbing, with the reference's
names and argument meaning (compiled_surface_code_arbitrary_error_model.py [CS]).
"""
import numpy as np

from .api import load_error_models, load_lookup_table  # noqa: F401  (CS:910-913)


def instantiate_error_model(p_set, model):
    """CS:31-49: bind the i-th probability list to the i-th name of ``model``.

    Returns ``(error_model, {name: list})``.  Deviation (SURVEY C-11): a length mismatch raises
    ValueError instead of printing and returning a half-filled dict.
    """
    error_model = load_error_models()[model]
    names = error_model["probabilities"]
    if len(names) != len(p_set):
        raise ValueError("make sure you pass enough probabilities to specify the error model - see error_model.json")
    return error_model, {key: p_set[i] for i, key in enumerate(names)}


def lookup(syndrome, table, display=False):
    """CS:886-892: 8-bit syndrome (np.ndarray of ints) -> 18-element correction vector."""
    key = " ".join(str(int(b)) for b in np.asarray(syndrome).ravel())
    error_vec = table[key][0]
    if display:
        print('ft syndrome: {}'.format(syndrome))
        print('error vector: {}'.format(error_vec))
    return error_vec


def return_weight_classical_lookup(syndrome, table):
    """CS:895-898."""
    key = " ".join(str(int(b)) for b in np.asarray(syndrome).ravel())
    return table[key][1]


def get_results_log_e(rounds, runs, incorrect_count, time):
    return {"runs": runs, "rounds": rounds, "incorrect_count": incorrect_count, "time_taken": time}


def get_results_log_e_run_til_fail(rounds, time, p1q, p2q):
    return {"rounds_til_fail": rounds, "p1q_error": p1q, "p2q_error": p2q, "time_taken": time}
