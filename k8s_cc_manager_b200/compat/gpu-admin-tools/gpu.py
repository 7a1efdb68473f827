"""`from gpu import GpuError` (reference main.py:40) -> the shim's error type."""
from k8s_cc_manager_b200.devices import GpuError  # noqa: F401
