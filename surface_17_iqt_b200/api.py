"""`Simulation`: a programmed device context for one (error model, protocol shape) pair."""
import json
import os

from . import _lib as L  # noqa: F401  (re-exported: api.L holds the constants)
from . import circuit
from .backend import Backend

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_error_models(path=None):
    """error_models.json in the reference's format (compiled_surface_code_arbitrary_error_model.py:38-40).
    Looks in the current directory first, like the reference, then in the packaged copy."""
    if path is None:
        path = "error_models.json" if os.path.exists("error_models.json") else os.path.join(DATA_DIR, "error_models.json")
    with open(path, "r") as infile:
        return json.load(infile)


def load_lookup_table(filename):
    """Same contract as the reference's load_lookup_table (CS:910-913), with a packaged fallback."""
    if not os.path.exists(filename) and os.path.exists(os.path.join(DATA_DIR, os.path.basename(filename))):
        filename = os.path.join(DATA_DIR, os.path.basename(filename))
    with open(filename, "r") as infile:
        return json.load(infile)


class Simulation:
    """Compiles `model` for probability lists of the given total lengths and owns the device context."""

    def __init__(self, model, list_len, max_shots, correction_table, bitflip_table, fusion=1, cooling=False,
                 device=0, error_models=None, over_angles=None, early_measure=None):
        models = error_models if error_models is not None else load_error_models()
        self.model_name = model
        self.model = models[model]
        if early_measure is None:
            early_measure = os.environ.get("S17_EARLY", "1") == "1"
        self.program = circuit.compile_program(self.model, list_len, fusion=fusion, cooling=cooling,
                                               over_angles=over_angles, early_measure=early_measure)
        self.backend = Backend(max_shots, device=device)
        self.backend.load_program(self.program)
        self.backend.load_tables(correction_table, bitflip_table)
        self.max_shots = max_shots

    def close(self):
        self.backend.close()
