from ._core import Simulator  # noqa: F401
