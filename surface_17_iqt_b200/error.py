"""Error objects of the object-oriented interface -- same names as the reference's error.py (:4-69).

In the reference an Error inserts ProjectQ gates when its location is reached; here errors are *data* that
``logical_qubit.StabiliserCycle`` turns into error-slot masks for the batched device path.
"""
from random import randint


class Error:
    """Error object: a name (type) and a location ``[syndrome, timestep, entangling_op, gate]``."""
    pauli = None

    def __init__(self, name="Unnamed error", location=None):
        self.name = name
        self.location = location

    def insert_error(self, *qubits):
        raise Exception("This error does not have a defined insert_error function.")

    def generate_location_list(self):
        raise Exception("This error does not have a defined generate_location_list function.")

    def generate_random_location(self, location_list):
        return location_list[randint(0, len(location_list) - 1)]      # uniform, WITH replacement (error.py:17-19)

    @classmethod
    def generate_error_list(cls, error_subset, error_locations):
        """error.py:21-33.  error_subset: {"XCtrlError": 5, ...}; error_locations: {XCtrlError: [locations]}."""
        error_list = []
        for error in error_subset.keys():
            klass = ERROR_CLASSES[error] if isinstance(error, str) else error
            for _ in range(error_subset[error]):
                error_i = klass(name=klass.__name__)
                error_i.location = list(error_i.generate_random_location(error_locations[klass]))
                error_list.append(error_i)
        return error_list


class XCtrlError(Error):
    pauli = "X"

    def __init__(self, name="XCtrlError", location=None):
        super().__init__(name=name, location=location)

    def insert_error(self, *qubits):
        return [("X", q) for q in qubits]


class DephasingError(Error):
    pauli = "Z"

    def __init__(self, name="DephasingError", location=None):
        super().__init__(name=name, location=location)

    def insert_error(self, *qubits):
        return [("Z", q) for q in qubits]


class YCtrlError(Error):
    pauli = "Y"

    def __init__(self, name="YCtrlError", location=None):
        super().__init__(name=name, location=location)

    def insert_error(self, *qubits):
        return [("Y", q) for q in qubits]


ERROR_CLASSES = {"XCtrlError": XCtrlError, "DephasingError": DephasingError, "YCtrlError": YCtrlError}


def create_error_subset_list():
    """error.py:65-69."""
    return [{"DephasingError": i} for i in range(6)]
