#!/usr/bin/env python3
"""Small, complete pass over every kernel variant for compute-sanitizer (memcheck / racecheck /
synccheck): ragged host round trips + a 96 MiB arena with static and dynamic schedules."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from k8s_cc_manager_b200 import _native as N
L = N.lib(); assert L.ccm_init(1) == 0
rng = np.random.default_rng(3)
for sv in (1, 2, 3):
    for vv in (1, 2, 3):
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
        for vv in (1, 2, 3):
            vcfg = N.launch_cfg(schedule=sched if vv != 3 else 0)
            assert L.ccm_arena_verify(0, vv, C.byref(vcfg), None, C.byref(nz), None) == 0 and nz.value == 0
assert L.ccm_arena_fill_random(0, 5, None) == 0
assert L.ccm_arena_verify(0, 0, None, None, C.byref(nz), None) == 0 and nz.value > 0
L.ccm_arena_release(0)
r = N.ScrubResult(); assert L.ccm_scrub_verify(0, 256 << 20, C.byref(r)) == 0 and r.nonzero_bytes == 0
print("sanitize target ok", L.ccm_kernel_launches(), "launches")
