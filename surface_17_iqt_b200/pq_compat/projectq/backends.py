from surface_17_iqt_b200.pq_core import RecordingBackend, Simulator  # noqa: F401
