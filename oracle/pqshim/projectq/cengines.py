from ._core import DummyEngine  # noqa: F401
