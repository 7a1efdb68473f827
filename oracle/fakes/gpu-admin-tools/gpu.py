from _state import GpuError  # noqa: F401  (reference main.py:40: `from gpu import GpuError`)
