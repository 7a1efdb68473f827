#!/usr/bin/env python3
"""Small, complete pass over every kernel variant for compute-sanitizer (memcheck / racecheck /
synccheck): ragged host round trips, a 96 MiB arena with static and dynamic schedules, the default
kernels' self-resetting grab counters over back-to-back launches (incl. the scrub that clears the verify
counter), and the product call in all three shapes (pipelined, mapped-first, with the dirt drill).
Run once more with CCM_PDL=1 to cover the programmatic-dependent-launch path."""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from k8s_cc_manager_b200 import _native as N
L = N.lib(); assert L.ccm_init(1) == 0
rng = np.random.default_rng(3)
for sv in (0, 1, 2, 3):
    for vv in (0, 1, 2):
        for nbytes, off in ((0, 0), (1, 1), (4099, 3), (65536 + 7, 16), ((1 << 20) + 5, 100)):
            host = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
            want = int(np.count_nonzero(host))
            pre, post = C.c_uint64(), C.c_uint64()
            rc = L.ccm_host_roundtrip(0, host.ctypes.data if nbytes else None, nbytes, off, sv, vv, C.byref(pre), C.byref(post))
            assert rc == 0 and pre.value == want and post.value == 0 and not host.any(), (sv, vv, nbytes, rc, N.last_error())
ai = N.ArenaInfo(); assert L.ccm_arena_acquire(0, (96 << 20) + 48, C.byref(ai)) == 0
nz = C.c_uint64()
for sched in (1, 2, 3):
    for sv in (1, 2, 3):
        cfg = N.launch_cfg(schedule=sched)
        assert L.ccm_arena_fill(0, 0xA5, None) == 0
        assert L.ccm_arena_scrub(0, sv, C.byref(cfg), None, None) == 0
        for vv in (1, 2):
            vcfg = N.launch_cfg(schedule=sched)
            assert L.ccm_arena_verify(0, vv, C.byref(vcfg), None, C.byref(nz), None) == 0 and nz.value == 0
# default kernels, back to back: every launch must find its grab/done words zeroed by the previous one
assert L.ccm_arena_fill(0, 0xA5, None) == 0
for _ in range(6):
    assert L.ccm_arena_scrub_verify_async(0, 0, 0, None, None, None) == 0
assert L.ccm_arena_fetch_count(0, None, C.byref(nz)) == 0 and nz.value == 0
assert L.ccm_arena_write(0, 12345, (C.c_uint8 * 1)(7), 1) == 0
assert L.ccm_arena_verify(0, 0, None, None, C.byref(nz), None) == 0 and nz.value == 1
assert L.ccm_arena_verify(0, 0, None, None, C.byref(nz), None) == 0 and nz.value == 1
assert L.ccm_arena_scrub_verify_async(0, 0, 0, None, None, None) == 0
assert L.ccm_arena_fetch_count(0, None, C.byref(nz)) == 0 and nz.value == 0
assert L.ccm_arena_fill_random(0, 5, None) == 0
assert L.ccm_arena_verify(0, 0, None, None, C.byref(nz), None) == 0 and nz.value > 0
L.ccm_arena_release(0)
r = N.ScrubResult()
for env in ({}, {"CCM_MAP_FIRST": "1"}, {"CCM_INTERLEAVE_VERIFY": "0"}, {"CCM_ASYNC_RELEASE": "0"}):
    os.environ.update(env)
    assert L.ccm_scrub_verify(0, (3 << 30) + (2 << 20), C.byref(r)) == 0 and r.nonzero_bytes == 0, N.last_error()
    assert L.ccm_scrub_release_wait(0, None, None) == 0
    for k in env:
        del os.environ[k]
L.ccm_sim_set(0, b"scrub_inject", 6)
assert L.ccm_scrub_verify(0, 2 << 30, C.byref(r)) == N.ERR_DIRTY and r.nonzero_bytes == 6
L.ccm_sim_set(0, b"scrub_inject", 0)
assert L.ccm_device_release(0) == 0
print("sanitize target ok", L.ccm_kernel_launches(), "launches")
