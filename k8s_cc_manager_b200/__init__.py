"""B200-native CC-mode manager: host side of the drop-in for NVIDIA/k8s-cc-manager.

Only what the hot path needs lives here:
  csrc/        sm_100a scrub/verify kernels + the C ABI (include/ccm.h) -> libccm.so
  _native.py   ctypes binding (no fallback: raises if libccm.so is missing)
  devices.py   duck-typed device objects the reference manager expects
  manager.py   CCManager — reference main.py's transition engine, concurrent + scrub gate
  drain_gate.py operator-component pause / restore (reference gpu_operator_eviction.py)
"""

__version__ = "0.1.0"
