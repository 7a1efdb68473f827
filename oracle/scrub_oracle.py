"""numpy statement of the scrub contract + loader for the C oracle.

ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/scrub_oracle.c for the contract and
its parity status).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libscrub_oracle.so"

MASK64 = (1 << 64) - 1


# ---- independent numpy / pure-Python statements --------------------------------
def scrub_np(buf: np.ndarray) -> None:
    buf.view(np.uint8)[...] = 0


def count_nonzero_np(buf: np.ndarray) -> int:
    return int(np.count_nonzero(buf.view(np.uint8)))


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK64
    return x ^ (x >> 31)


def pattern_word_py(seed: int, j: int) -> int:
    """Pure-Python statement of the seeded sparse pattern (small cases only)."""
    r = _splitmix64((seed + j) & MASK64)
    if r & 7:
        return 0
    keep = 0
    for b in range(8):
        if (r >> (8 + b)) & 1:
            keep |= 0xFF << (8 * b)
    return ((r >> 3) | 0x0101010101010101) & keep


def pattern_np(nbytes: int, seed: int, word_index0: int = 0) -> np.ndarray:
    """Vectorised numpy statement of the same pattern."""
    nwords = nbytes // 8
    j = np.arange(word_index0, word_index0 + nwords, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = j + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        r = x ^ (x >> np.uint64(31))
    keep = np.zeros(nwords, dtype=np.uint64)
    for b in range(8):
        bit = (r >> np.uint64(8 + b)) & np.uint64(1)
        keep |= (bit * np.uint64(0xFF)) << np.uint64(8 * b)
    w = ((r >> np.uint64(3)) | np.uint64(0x0101010101010101)) & keep
    w[(r & np.uint64(7)) != 0] = 0
    out = np.zeros(nbytes, dtype=np.uint8)
    out[: nwords * 8] = w.astype("<u8").view(np.uint8)
    return out


# ---- C oracle --------------------------------------------------------------------
def build() -> Path:
    if not LIB.exists() or LIB.stat().st_mtime < (HERE / "scrub_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    return LIB


_lib = None


def clib() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        lib.ccm_oracle_scrub.argtypes = [C.c_void_p, C.c_uint64]
        lib.ccm_oracle_scrub.restype = None
        lib.ccm_oracle_count_nonzero.argtypes = [C.c_void_p, C.c_uint64]
        lib.ccm_oracle_count_nonzero.restype = C.c_uint64
        lib.ccm_oracle_pattern_word.argtypes = [C.c_uint64, C.c_uint64]
        lib.ccm_oracle_pattern_word.restype = C.c_uint64
        lib.ccm_oracle_fill_pattern.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        lib.ccm_oracle_fill_pattern.restype = None
        lib.ccm_oracle_pattern_count.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        lib.ccm_oracle_pattern_count.restype = C.c_uint64
        lib.ccm_oracle_scrub_verify_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        lib.ccm_oracle_scrub_verify_mt.restype = C.c_uint64
        lib.ccm_oracle_scrub_verify_mt_pinned.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        lib.ccm_oracle_scrub_verify_mt_pinned.restype = C.c_uint64
        lib.ccm_oracle_fill_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        lib.ccm_oracle_fill_mt.restype = None
        lib.ccm_oracle_pool_create.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        lib.ccm_oracle_pool_create.restype = C.c_void_p
        lib.ccm_oracle_pool_pass.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ccm_oracle_pool_pass.restype = C.c_uint64
        lib.ccm_oracle_pool_destroy.argtypes = [C.c_void_p]
        lib.ccm_oracle_pool_destroy.restype = None
        _lib = lib
    return _lib


def scrub_c(buf: np.ndarray) -> None:
    clib().ccm_oracle_scrub(buf.ctypes.data, buf.nbytes)


def count_nonzero_c(buf: np.ndarray) -> int:
    return int(clib().ccm_oracle_count_nonzero(buf.ctypes.data, buf.nbytes))


def fill_pattern_c(buf: np.ndarray, seed: int, word_index0: int = 0) -> None:
    clib().ccm_oracle_fill_pattern(buf.ctypes.data, buf.nbytes, seed, word_index0)


def pattern_count_c(nbytes: int, seed: int, word_index0: int = 0) -> int:
    return int(clib().ccm_oracle_pattern_count(nbytes, seed, word_index0))


def scrub_verify_mt_c(buf: np.ndarray, threads: int, scrub: bool = True, verify: bool = True,
                      pin: bool = False) -> int:
    """memset + byte-wise count over `threads` contiguous slices.  pin=True: thread i runs on the
    i-th allowed CPU (use with fill_mt_c so every slice is NUMA-local to its thread)."""
    mode = (1 if scrub else 0) | (2 if verify else 0)
    if pin:
        return int(clib().ccm_oracle_scrub_verify_mt_pinned(buf.ctypes.data, buf.nbytes, threads, mode, 1))
    return int(clib().ccm_oracle_scrub_verify_mt(buf.ctypes.data, buf.nbytes, threads, mode))


def fill_mt_c(buf: np.ndarray, threads: int, byte: int, pin: bool = True) -> None:
    """Parallel first touch: thread i fills slice i with `byte` (same partition / pinning as
    scrub_verify_mt_c(pin=True))."""
    clib().ccm_oracle_fill_mt(buf.ctypes.data, buf.nbytes, threads, byte & 0xFF, 1 if pin else 0)


def numa_layout() -> str:
    """'2 nodes: node0 cpus 0-63 ...' from sysfs, or 'unknown'."""
    import glob
    nodes = []
    for path in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        try:
            nodes.append(f"{path.rsplit('/', 1)[1]} cpus {open(path + '/cpulist').read().strip()}")
        except OSError:
            pass
    return f"{len(nodes)} NUMA node(s): " + "; ".join(nodes) if nodes else "NUMA layout unknown"


class Pool:
    """Persistent pinned worker pool over one host buffer (bench.py's CPU arm): thread i owns slice i
    for its first touch and for every pass.  stream=True scrubs with non-temporal stores instead of
    libc memset (same bytes, no write-allocate reads)."""

    def __init__(self, buf: np.ndarray, threads: int, pin: bool = True):
        self.buf, self.threads = buf, threads
        self._h = clib().ccm_oracle_pool_create(buf.ctypes.data, buf.nbytes, threads, 1 if pin else 0)

    def poison(self) -> int:
        """first touch: fill 0xA5, return the count (== nbytes)"""
        return int(clib().ccm_oracle_pool_pass(self._h, 4 | 2, 0))

    def scrub_verify(self, stream: bool = False) -> int:
        return int(clib().ccm_oracle_pool_pass(self._h, 1 | 2, 1 if stream else 0))

    def close(self) -> None:
        if self._h:
            clib().ccm_oracle_pool_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
