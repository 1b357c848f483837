"""Error objects of the object-oriented interface: the names of the reference's error.py (:4-69) -- ``Error``,
``XCtrlError``, ``DephasingError``, ``YCtrlError``, ``Error.generate_error_list``, ``create_error_subset_list`` --
with errors as *data*.

In the reference an error object issues a ProjectQ Pauli gate when the circuit reaches its location.  Here an error
is a (Pauli, location) record; ``logical_qubit.StabiliserCycle.slot_masks`` turns a list of them into the error-slot
bitmasks of the batched device path.  A location is ``[syndrome, timestep, entangling_op, gate]``.
"""
import random

_REGISTRY = {}


class Error:
    """Base record: ``name`` (the error type) and ``location``.  ``pauli`` is the single-qubit Pauli the type stands
    for; the base class has none and cannot be inserted."""

    pauli = None

    def __init__(self, name=None, location=None):
        self.name = name if name is not None else (type(self).__name__ if self.pauli else "Unnamed error")
        self.location = location

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        _REGISTRY[cls.__name__] = cls

    def __repr__(self):
        return "{}(location={})".format(type(self).__name__, self.location)

    # -- what the reference does with gates, as data -------------------------------------------------------------
    def insert_error(self, *qubits):
        """error.py:39-61: ``X | q`` / ``Z | q`` / ``Y | q`` on each qubit when they are engine qubits (gate-by-gate
        path); returns the (Pauli, qubit) pairs either way (the batched path only needs those)."""
        if self.pauli is None:
            raise Exception("This error does not have a defined insert_error function.")
        from . import pq_core
        gate = {"X": pq_core.X, "Y": pq_core.Y, "Z": pq_core.Z}[self.pauli]
        for q in qubits:
            if isinstance(q, pq_core.Qubit):
                gate | q
        return [(self.pauli, q) for q in qubits]

    def generate_location_list(self):
        raise Exception("This error does not have a defined generate_location_list function.")

    @staticmethod
    def generate_random_location(location_list):
        """Uniform over the list, WITH replacement between calls (error.py:17-19 uses ``randint``)."""
        return location_list[random.randint(0, len(location_list) - 1)]

    @classmethod
    def generate_error_list(cls, error_subset, error_locations):
        """error.py:21-33.  ``error_subset``: counts per type, keyed by class name (or class), e.g.
        ``{"XCtrlError": 5, "DephasingError": 6}``; ``error_locations``: ``{class: [locations]}``.  Returns one
        error object per requested error, each at an independently drawn location of its type."""
        drawn = []
        for key, count in error_subset.items():
            kind = _REGISTRY[key] if isinstance(key, str) else key
            places = error_locations[kind]
            drawn.extend(kind(location=list(cls.generate_random_location(places))) for _ in range(count))
        return drawn


def _pauli_error(class_name, pauli, doc):
    return type(class_name, (Error,), {"pauli": pauli, "__doc__": doc, "__module__": __name__})


XCtrlError = _pauli_error("XCtrlError", "X", "Control error of an Rx gate: an X on the data qubit (p_X_ctrl).")
DephasingError = _pauli_error("DephasingError", "Z", "Dephasing during any gate: a Z on the qubit (p_deph).")
YCtrlError = _pauli_error("YCtrlError", "Y", "Control error of an Ry gate: a Y on the data qubit (p_Y_ctrl).")

ERROR_CLASSES = dict(_REGISTRY)


def create_error_subset_list():
    """error.py:65-69: the subsets with 0..5 dephasing errors."""
    return [{"DephasingError": n} for n in range(6)]
