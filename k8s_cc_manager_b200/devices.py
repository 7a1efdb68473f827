"""Duck-typed device objects over libccm.so.

The reference manager never sees a C ABI: it calls methods on objects handed out by
`pci.devices.find_gpus()` of NVIDIA/gpu-admin-tools (reference main.py:38-40,155).
This module rebuilds exactly the attributes the reference touches
on top of include/ccm.h, and adds the one new capability, `scrub_and_verify()`.

    find_gpus() -> (devices, count)          reference main.py:155,164,174,203,275,473
    dev.bdf, dev.name                        main.py:187,191,210,239,280
    dev.is_gpu(), dev.is_nvswitch()          main.py:165,175
    dev.is_cc_query_supported  (attribute)   main.py:186
    dev.is_ppcie_query_supported (attribute) main.py:205,477
    dev.query_cc_mode() / set_cc_mode(m)     main.py:441,505,511,524
    dev.query_ppcie_mode() / set_ppcie_mode  main.py:310,340,345,353,359,373,479,482,495
    dev.reset_with_os() / wait_for_boot()    main.py:346-347,368,372,490,494,519,523
    GpuError                                 main.py:40,380,531

Errors: every non-zero ccm_status becomes a GpuError carrying `.status`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

from . import _native as N


class GpuError(Exception):
    """Raised for any failed device operation (reference: gpu.GpuError)."""

    def __init__(self, message: str, status: int = N.ERR_IO):
        super().__init__(message)
        self.status = status


def _check(rc: int, what: str, bdf: str = "") -> None:
    if rc != N.OK:
        detail = N.last_error()
        where = f" on {bdf}" if bdf else ""
        raise GpuError(f"{what}{where}: {N.strerror(rc)}" + (f" ({detail})" if detail else ""), rc)


@dataclass
class ScrubReport:
    """Outcome of one GPU's HBM scrub-and-verify (include/ccm.h: ccm_scrub_result)."""
    bdf: str
    bytes_requested: int
    bytes_scrubbed: int
    device_total_bytes: int
    nonzero_bytes: int
    ms_acquire: float
    ms_scrub: float
    ms_verify: float
    ms_release: float
    ms_total: float
    segments: int
    status: int
    release_deferred: int = 0        # 1: the HBM is handed back by libccm's background reaper
    device_free_before: int = 0      # free HBM when the call started
    bytes_unreached: int = 0         # free HBM that could not be mapped (hence not scrubbed)
    ms_release_wait: float = 0.0     # waited for the previous call's deferred release
    ms_gpu_span: float = 0.0         # first scrub launch .. last verify done

    @property
    def clean(self) -> bool:
        return self.status == N.OK and self.nonzero_bytes == 0

    @property
    def coverage(self) -> float:
        """Scrubbed fraction of the DEVICE's memory (the rest is the CUDA context's own
        footprint, driver reservations and whatever other contexts hold)."""
        return self.bytes_scrubbed / self.device_total_bytes if self.device_total_bytes else 0.0

    @property
    def coverage_of_free(self) -> float:
        """Scrubbed fraction of what was free when the call started (1.0 = every reachable byte)."""
        return self.bytes_scrubbed / self.device_free_before if self.device_free_before else 0.0

    @property
    def scrub_gbs(self) -> float:
        return self.bytes_scrubbed / self.ms_scrub / 1e6 if self.ms_scrub > 0 else 0.0

    @property
    def verify_gbs(self) -> float:
        return self.bytes_scrubbed / self.ms_verify / 1e6 if self.ms_verify > 0 else 0.0

    @classmethod
    def from_native(cls, bdf: str, r: N.ScrubResult) -> "ScrubReport":
        return cls(bdf, r.bytes_requested, r.bytes_scrubbed, r.device_total_bytes, r.nonzero_bytes,
                   r.ms_acquire, r.ms_scrub, r.ms_verify, r.ms_release, r.ms_total, r.segments, r.status,
                   r.release_deferred, r.device_free_before, r.bytes_unreached, r.ms_release_wait, r.ms_gpu_span)


class NvidiaDevice:
    """One NVIDIA PCI function (GPU or NVSwitch) behind the C ABI."""

    def __init__(self, info: N.DevInfo):
        self.index = int(info.index)
        self.bdf = info.bdf.decode()
        self.name = info.name.decode()
        self.kind = int(info.kind)
        self.is_cc_query_supported = bool(info.cc_query_supported)
        self.is_ppcie_query_supported = bool(info.ppcie_query_supported)
        self.cuda_ordinal = int(info.cuda_ordinal)
        self.hbm_total_bytes = int(info.hbm_total_bytes)

    def __repr__(self) -> str:
        return f"<{type(self).__name__} {self.bdf} {self.name!r}>"

    def is_gpu(self) -> bool:
        return self.kind == N.KIND_GPU

    def is_nvswitch(self) -> bool:
        return self.kind == N.KIND_NVSWITCH

    # -- CC mode -------------------------------------------------------------
    def query_cc_mode(self) -> str:
        mode = C.c_int(-1)
        _check(N.lib().ccm_query_cc_mode(self.index, C.byref(mode)), "query_cc_mode", self.bdf)
        return N.CC_MODE_NAMES[mode.value]

    def set_cc_mode(self, mode: str) -> None:
        if mode not in N.CC_MODES:
            raise GpuError(f"invalid CC mode {mode!r} for {self.bdf}", N.ERR_INVALID)
        _check(N.lib().ccm_set_cc_mode(self.index, N.CC_MODES[mode]), "set_cc_mode", self.bdf)

    # -- PPCIe mode ------------------------------------------------------------
    def query_ppcie_mode(self) -> str:
        mode = C.c_int(-1)
        _check(N.lib().ccm_query_ppcie_mode(self.index, C.byref(mode)), "query_ppcie_mode", self.bdf)
        return N.PPCIE_MODE_NAMES[mode.value]

    def set_ppcie_mode(self, mode: str) -> None:
        if mode not in N.PPCIE_MODES:
            raise GpuError(f"invalid PPCIe mode {mode!r} for {self.bdf}", N.ERR_INVALID)
        _check(N.lib().ccm_set_ppcie_mode(self.index, N.PPCIE_MODES[mode]), "set_ppcie_mode", self.bdf)

    # -- reset / boot ----------------------------------------------------------
    def reset_with_os(self) -> None:
        _check(N.lib().ccm_reset(self.index), "reset_with_os", self.bdf)

    def wait_for_boot(self, timeout_ms: int = 0) -> None:
        _check(N.lib().ccm_wait_for_boot(self.index, timeout_ms), "wait_for_boot", self.bdf)

    # -- NEW: HBM scrub-and-verify ---------------------------------------------
    def scrub_and_verify(self, nbytes: int = 0) -> ScrubReport:
        """Zero-fills and reads back this GPU's HBM (nbytes=0: all the context can map).

        Raises GpuError if the scrub cannot run (no CUDA device, no HBM, CUDA error)
        or if any byte reads back non-zero.
        """
        res = N.ScrubResult()
        rc = N.lib().ccm_scrub_verify(self.index, nbytes, C.byref(res))
        _check(rc, "scrub_and_verify", self.bdf)
        return ScrubReport.from_native(self.bdf, res)

    def wait_scrub_released(self) -> Tuple[float, float]:
        """Blocks until the HBM of the last scrub_and_verify() is back with the driver (libccm
        hands it back on a background thread so the verdict is not held up by cuMemUnmap /
        cuMemRelease).  Returns (ms the release took, ms this call waited)."""
        rel, waited = C.c_double(0.0), C.c_double(0.0)
        _check(N.lib().ccm_scrub_release_wait(self.index, C.byref(rel), C.byref(waited)), "wait_scrub_released", self.bdf)
        return rel.value, waited.value

    def release_cuda_context(self) -> None:
        """Give back everything libccm holds on this GPU, including its CUDA primary context
        (include/ccm.h: ccm_device_release).  Call it when the scrub gate is done."""
        _check(N.lib().ccm_device_release(self.index), "release_cuda_context", self.bdf)


class Gpu(NvidiaDevice):
    pass


class NvSwitch(NvidiaDevice):
    pass


def find_gpus() -> Tuple[List[NvidiaDevice], int]:
    """All NVIDIA PCI functions (GPUs AND NVSwitches), freshly enumerated.

    Mirrors pci.devices.find_gpus() as the reference uses it (main.py:144-155): the
    second element is the device count.
    """
    lib = N.lib()
    n = C.c_int(0)
    _check(lib.ccm_enumerate(None, 0, C.byref(n)), "enumerate")
    infos = (N.DevInfo * max(1, n.value))()
    _check(lib.ccm_enumerate(infos, n.value, C.byref(n)), "enumerate")
    devices: List[NvidiaDevice] = []
    for i in range(n.value):
        cls = Gpu if infos[i].kind == N.KIND_GPU else NvSwitch
        devices.append(cls(infos[i]))
    return devices, len(devices)


def scrub_and_verify_many(devices: Sequence[NvidiaDevice], nbytes: int = 0) -> Tuple[List[ScrubReport], float]:
    """Concurrent multi-context scrub of several GPUs through ONE native call
    (ccm_scrub_verify_many: one host thread + primary context + stream per GPU, no
    peer access, no collective).  Returns (reports, wall_ms); never raises for a
    per-device failure — inspect report.status / report.clean."""
    n = len(devices)
    if n == 0:
        return [], 0.0
    idx = (C.c_int * n)(*[d.index for d in devices])
    res = (N.ScrubResult * n)()
    wall = C.c_double(0.0)
    N.lib().ccm_scrub_verify_many(n, idx, nbytes, res, C.byref(wall))
    return [ScrubReport.from_native(d.bdf, res[i]) for i, d in enumerate(devices)], wall.value


def release_cuda_contexts(devices: Sequence[NvidiaDevice]) -> float:
    """Concurrent ccm_device_release for several GPUs (one native thread each): joins each
    engine's deferred HBM release, then resets its CUDA primary context.  Returns wall ms."""
    n = len(devices)
    if n == 0:
        return 0.0
    idx = (C.c_int * n)(*[d.index for d in devices])
    wall = C.c_double(0.0)
    _check(N.lib().ccm_device_release_many(n, idx, C.byref(wall)), "release_cuda_contexts")
    return wall.value


def select_backend(name: str) -> None:
    """Explicit backend choice ('sim' | 'cudasim' | 'sysfs'); rebuilds the device table."""
    if name not in N.BACKENDS:
        raise ValueError(f"unknown backend {name!r}")
    _check(N.lib().ccm_init(N.BACKENDS[name]), f"ccm_init({name})")


class ScrubbingProxy:
    """Wraps a device object from ANOTHER device library (e.g. the real gpu-admin-tools `Gpu`)
    and adds the one capability it lacks: `scrub_and_verify()`, served by libccm.so on the CUDA
    device with the same PCI address.  Every other attribute is forwarded untouched, so the
    manager keeps using that library's register access for query/set/reset/wait_for_boot
    (arrangement (1) of INTEGRATION.md §D)."""

    def __init__(self, foreign, native: "NvidiaDevice | None"):
        object.__setattr__(self, "_foreign", foreign)
        object.__setattr__(self, "_native", native)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_foreign"), name)

    def __setattr__(self, name, value):
        setattr(object.__getattribute__(self, "_foreign"), name, value)

    def __repr__(self):
        return f"<ScrubbingProxy {self._foreign!r} scrub={'libccm' if self._native else 'unavailable'}>"

    def release_cuda_context(self) -> None:
        native = object.__getattribute__(self, "_native")
        if native is not None:
            native.release_cuda_context()

    def wait_scrub_released(self):
        native = object.__getattribute__(self, "_native")
        return native.wait_scrub_released() if native is not None else (0.0, 0.0)

    def scrub_and_verify(self, nbytes: int = 0) -> ScrubReport:
        native = object.__getattribute__(self, "_native")
        if native is None:
            raise GpuError(f"no CUDA device with PCI address {self._foreign.bdf}: HBM scrub cannot run", N.ERR_NO_CUDA)
        return native.scrub_and_verify(nbytes)


def _norm_bdf(bdf: str) -> str:
    bdf = bdf.strip().lower()
    return bdf if bdf.count(":") == 2 else "0000:" + bdf


def with_scrub(foreign_find_gpus):
    """Turns another library's `find_gpus()` into a device source for CCManager whose GPUs can be
    scrubbed: `CCManager(..., device_source=with_scrub(pci.devices.find_gpus))`."""
    def source():
        devices, _ = foreign_find_gpus()
        by_bdf = {_norm_bdf(d.bdf): d for d in find_gpus()[0] if d.is_gpu() and d.cuda_ordinal >= 0}
        wrapped = [ScrubbingProxy(d, by_bdf.get(_norm_bdf(d.bdf))) if d.is_gpu() else d for d in devices]
        return wrapped, len(wrapped)
    return source
