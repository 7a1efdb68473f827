"""`from pci.devices import find_gpus` (reference main.py:39) -> libccm.so enumeration."""
from k8s_cc_manager_b200.devices import find_gpus  # noqa: F401
