"""ctypes binding of libccm.so (include/ccm.h).

This is the ONLY way the Python host code reaches the device shim.  There is no
pure-Python fallback: if the shared library is missing or does not export the ABI
declared in include/ccm.h, importing this module's `lib()` raises — the product
path must fail loudly rather than silently skipping the HBM scrub.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("CCM_LIB", PKG_DIR / "libccm.so"))

# ---- status codes (include/ccm.h: ccm_status) --------------------------------
OK = 0
ERR_INVALID = -1
ERR_NO_DEVICE = -2
ERR_UNSUPPORTED = -3
ERR_IO = -4
ERR_TIMEOUT = -5
ERR_CUDA = -6
ERR_NOMEM = -7
ERR_DIRTY = -8
ERR_NO_CUDA = -9
ERR_STATE = -10
ERR_FAULT = -11
ERR_NOT_BOOTED = -12

CC_MODES = {"off": 0, "on": 1, "devtools": 2}
CC_MODE_NAMES = {v: k for k, v in CC_MODES.items()}
PPCIE_MODES = {"off": 0, "on": 1}
PPCIE_MODE_NAMES = {v: k for k, v in PPCIE_MODES.items()}

KIND_GPU, KIND_NVSWITCH = 0, 1
BACKEND_SIM, BACKEND_CUDASIM, BACKEND_SYSFS = 0, 1, 2
BACKENDS = {"sim": BACKEND_SIM, "cudasim": BACKEND_CUDASIM, "sysfs": BACKEND_SYSFS}

SCRUB_AUTO, SCRUB_ST128, SCRUB_ST256, SCRUB_TMA, SCRUB_MEMSET = 0, 1, 2, 3, 4
VERIFY_AUTO, VERIFY_LD128, VERIFY_LD256 = 0, 1, 2
SCRUB_VARIANTS = {"auto": 0, "st128": 1, "st256": 2, "tma": 3, "memset": 4}
VERIFY_VARIANTS = {"auto": 0, "ld128": 1, "ld256": 2}

ABI_VERSION = 2  # include/ccm.h: CCM_ABI_VERSION

OP_QUERY_CC, OP_SET_CC, OP_QUERY_PPCIE, OP_SET_PPCIE, OP_RESET, OP_WAIT_BOOT, OP_SCRUB = 1, 2, 4, 8, 16, 32, 64


class DevInfo(C.Structure):
    _fields_ = [
        ("index", C.c_int32),
        ("kind", C.c_int32),
        ("cc_query_supported", C.c_int32),
        ("ppcie_query_supported", C.c_int32),
        ("cuda_ordinal", C.c_int32),
        ("reserved0", C.c_int32),
        ("hbm_total_bytes", C.c_uint64),
        ("bdf", C.c_char * 32),
        ("name", C.c_char * 96),
    ]


class LaunchCfg(C.Structure):
    _fields_ = [
        ("ctas_per_sm", C.c_int32),
        ("threads_per_cta", C.c_int32),
        ("tile_bytes", C.c_int32),
        ("unroll", C.c_int32),
        ("cache_policy", C.c_int32),
        ("schedule", C.c_int32),
    ]


class ScrubResult(C.Structure):
    _fields_ = [
        ("bytes_requested", C.c_uint64),
        ("bytes_scrubbed", C.c_uint64),
        ("device_total_bytes", C.c_uint64),
        ("nonzero_bytes", C.c_uint64),
        ("ms_acquire", C.c_double),
        ("ms_scrub", C.c_double),
        ("ms_verify", C.c_double),
        ("ms_release", C.c_double),
        ("ms_total", C.c_double),
        ("segments", C.c_int32),
        ("scrub_variant", C.c_int32),
        ("verify_variant", C.c_int32),
        ("sm_count", C.c_int32),
        ("status", C.c_int32),
        ("release_deferred", C.c_int32),
        ("device_free_before", C.c_uint64),
        ("bytes_unreached", C.c_uint64),
        ("ms_release_wait", C.c_double),
        ("ms_gpu_span", C.c_double),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


class ArenaInfo(C.Structure):
    _fields_ = [
        ("bytes", C.c_uint64),
        ("device_total_bytes", C.c_uint64),
        ("device_free_before", C.c_uint64),
        ("segments", C.c_int32),
        ("reserved", C.c_int32),
        ("ms_acquire", C.c_double),
    ]


_P = C.POINTER
_SIGNATURES = {
    # name: (restype, argtypes)
    "ccm_abi_version": (C.c_int, []),
    "ccm_strerror": (C.c_char_p, [C.c_int]),
    "ccm_last_error": (C.c_int, [C.c_char_p, C.c_size_t]),
    "ccm_init": (C.c_int, [C.c_int]),
    "ccm_backend_in_use": (C.c_int, []),
    "ccm_enumerate": (C.c_int, [_P(DevInfo), C.c_int, _P(C.c_int)]),
    "ccm_query_cc_mode": (C.c_int, [C.c_int, _P(C.c_int)]),
    "ccm_set_cc_mode": (C.c_int, [C.c_int, C.c_int]),
    "ccm_query_ppcie_mode": (C.c_int, [C.c_int, _P(C.c_int)]),
    "ccm_set_ppcie_mode": (C.c_int, [C.c_int, C.c_int]),
    "ccm_reset": (C.c_int, [C.c_int]),
    "ccm_wait_for_boot": (C.c_int, [C.c_int, C.c_int]),
    "ccm_scrub_verify": (C.c_int, [C.c_int, C.c_uint64, _P(ScrubResult)]),
    "ccm_scrub_verify_many": (C.c_int, [C.c_int, _P(C.c_int), C.c_uint64, _P(ScrubResult), _P(C.c_double)]),
    "ccm_scrub_release_wait": (C.c_int, [C.c_int, _P(C.c_double), _P(C.c_double)]),
    "ccm_arena_acquire": (C.c_int, [C.c_int, C.c_uint64, _P(ArenaInfo)]),
    "ccm_arena_release": (C.c_int, [C.c_int]),
    "ccm_arena_scrub": (C.c_int, [C.c_int, C.c_int, _P(LaunchCfg), C.c_void_p, _P(C.c_float)]),
    "ccm_arena_verify": (C.c_int, [C.c_int, C.c_int, _P(LaunchCfg), C.c_void_p, _P(C.c_uint64), _P(C.c_float)]),
    "ccm_arena_scrub_verify_async": (C.c_int, [C.c_int, C.c_int, C.c_int, _P(LaunchCfg), _P(LaunchCfg), C.c_void_p]),
    "ccm_arena_fetch_count": (C.c_int, [C.c_int, C.c_void_p, _P(C.c_uint64)]),
    "ccm_arena_step_times": (C.c_int, [C.c_int, C.c_int, _P(C.c_float), _P(C.c_float), _P(C.c_int)]),
    "ccm_arena_fill": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "ccm_arena_fill_random": (C.c_int, [C.c_int, C.c_uint64, C.c_void_p]),
    "ccm_arena_write": (C.c_int, [C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]),
    "ccm_arena_read": (C.c_int, [C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]),
    "ccm_region_scrub": (C.c_int, [C.c_int, C.c_void_p, C.c_uint64, C.c_int, _P(LaunchCfg), C.c_void_p, _P(C.c_float)]),
    "ccm_region_verify": (C.c_int, [C.c_int, C.c_void_p, C.c_uint64, C.c_int, _P(LaunchCfg), C.c_void_p, _P(C.c_uint64), _P(C.c_float)]),
    "ccm_host_roundtrip": (C.c_int, [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, _P(C.c_uint64), _P(C.c_uint64)]),
    "ccm_device_release": (C.c_int, [C.c_int]),
    "ccm_device_release_many": (C.c_int, [C.c_int, _P(C.c_int), _P(C.c_double)]),
    "ccm_kernel_launches": (C.c_uint64, []),
    "ccm_default_kernels": (C.c_char_p, []),
    "ccm_sim_topology": (C.c_int, [C.c_int, C.c_int]),
    "ccm_sim_set": (C.c_int, [C.c_int, C.c_char_p, C.c_int64]),
    "ccm_sim_get": (C.c_int, [C.c_int, C.c_char_p, _P(C.c_int64)]),
    "ccm_sim_trace": (C.c_int, [C.c_char_p, C.c_size_t]),
    "ccm_sim_trace_clear": (C.c_int, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lib_lock = threading.Lock()


class NativeLibraryError(ImportError):
    """libccm.so is missing or does not match include/ccm.h."""


def lib() -> C.CDLL:
    """Loads libccm.so once and types every entry point.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -m k8s_cc_manager_b200.build` "
                "(the CC manager has no host fallback for the device shim)"
            )
        handle = C.CDLL(str(LIB_PATH), mode=C.RTLD_LOCAL)
        for name, (restype, argtypes) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as exc:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.ccm_abi_version() != ABI_VERSION:
            raise NativeLibraryError(f"{LIB_PATH}: ABI version {handle.ccm_abi_version()} != {ABI_VERSION}")
        _lib = handle
    return _lib


def default_kernels() -> tuple:
    """(scrub kernel, verify kernel) AUTO launches, named the way ncu prints them."""
    return tuple(lib().ccm_default_kernels().decode().split(";"))


def strerror(status: int) -> str:
    return lib().ccm_strerror(status).decode()


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().ccm_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def launch_cfg(ctas_per_sm: int = 0, threads: int = 0, tile_bytes: int = 0, unroll: int = 0,
               cache_policy: int = 0, schedule: int = 0) -> LaunchCfg:
    """cache_policy: 0 library default, 1 plain, 2 evict_first, 3 streaming, 4 evict_last.
    schedule: 0 library default, 1 static grid-stride, 2 dynamic (atomic chunk grabs)."""
    return LaunchCfg(ctas_per_sm, threads, tile_bytes, unroll, cache_policy, schedule)
