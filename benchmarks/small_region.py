#!/usr/bin/env python3
"""Small end of the region sweep (BASELINE configs[4], 1-8 GB): where do the microseconds go?
Times ONE launch of the default scrub / verify kernel with CUDA events (best and median of 30 after
5 warm passes), per launch shape (CCM_FAST_*_SHAPE = threads x vectors x CTAs/SM), plus the scrub+verify
PAIR (one event bracket around both launches) with and without programmatic dependent launch.

  python benchmarks/small_region.py --gb 1,4 (round-2 result: launch shape is irrelevant — profiles/r2_small_region_shapes.log — so the library keeps ONE instantiation; the shape switches of this script only work on a build that carries more)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def worker(args):
    from k8s_cc_manager_b200 import _native as N
    L = N.lib()
    assert L.ccm_init(N.BACKEND_CUDASIM) == 0
    out = []
    ms, nz = C.c_float(), C.c_uint64()
    for gb in [float(x) for x in args.gb.split(",")]:
        want = int(gb * 1e9) // (2 << 20) * (2 << 20)
        ai = N.ArenaInfo()
        assert L.ccm_arena_acquire(0, want, C.byref(ai)) == 0, N.last_error()
        R = ai.bytes
        s_t, v_t, p_t = [], [], []
        for i in range(35):
            assert L.ccm_arena_fill(0, 0xA5, None) == 0
            assert L.ccm_arena_scrub(0, N.SCRUB_AUTO, None, None, C.byref(ms)) == 0, N.last_error()
            if i >= 5:
                s_t.append(ms.value)
            assert L.ccm_arena_verify(0, N.VERIFY_AUTO, None, None, C.byref(nz), C.byref(ms)) == 0, N.last_error()
            assert nz.value == 0
            if i >= 5:
                v_t.append(ms.value)
        # pair: the bench's step (two launches back to back, no host sync in between)
        L.ccm_arena_step_times(0, 0, None, None, None)
        for i in range(35):
            assert L.ccm_arena_scrub_verify_async(0, N.SCRUB_AUTO, N.VERIFY_AUTO, None, None, None) == 0
        sm, vm, n = (C.c_float * 64)(), (C.c_float * 64)(), C.c_int()
        assert L.ccm_arena_step_times(0, 64, sm, vm, C.byref(n)) == 0
        p_t = [sm[i] + vm[i] for i in range(5, n.value)]
        assert L.ccm_arena_fetch_count(0, None, C.byref(nz)) == 0 and nz.value == 0
        assert L.ccm_arena_release(0) == 0
        out.append({"gb": gb, "bytes": R,
                    "scrub_gbs_best": R / min(s_t) / 1e6, "scrub_gbs_median": R / statistics.median(s_t) / 1e6,
                    "verify_gbs_best": R / min(v_t) / 1e6, "verify_gbs_median": R / statistics.median(v_t) / 1e6,
                    "scrub_us_median": statistics.median(s_t) * 1e3, "verify_us_median": statistics.median(v_t) * 1e3,
                    "pair_us_median": statistics.median(p_t) * 1e3, "pair_gbs_median": 2 * R / statistics.median(p_t) / 1e6})
    print("ROW " + json.dumps({"scrub_shape": os.environ.get("CCM_FAST_SCRUB_SHAPE", "default"),
                               "verify_shape": os.environ.get("CCM_FAST_VERIFY_SHAPE", "default"),
                               "pdl": os.environ.get("CCM_PDL", "0"), "sizes": out}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", default="1,4")
    ap.add_argument("--shapes-scrub", default="")
    ap.add_argument("--shapes-verify", default="")
    ap.add_argument("--worker", action="store_true")
    args = ap.parse_args()
    if args.worker:
        return worker(args)
    runs = [({}, "default")]
    for sh in filter(None, args.shapes_scrub.split(",")):
        runs.append(({"CCM_FAST_SCRUB_SHAPE": sh}, f"scrub {sh}"))
    for sh in filter(None, args.shapes_verify.split(",")):
        runs.append(({"CCM_FAST_VERIFY_SHAPE": sh}, f"verify {sh}"))
    runs.append(({"CCM_PDL": "1"}, "default + PDL"))
    for env, tag in runs:
        proc = subprocess.run([sys.executable, __file__, "--worker", "--gb", args.gb], capture_output=True, text=True,
                              env=dict(os.environ, **env), timeout=600)
        rows = [ln for ln in proc.stdout.splitlines() if ln.startswith("ROW ")]
        if not rows:
            print(f"FAILED {tag}: {proc.stderr[-400:]}", flush=True)
            continue
        r = json.loads(rows[-1][4:])
        for s in r["sizes"]:
            print(f"{tag:18s} {s['gb']:5.1f} GB  scrub {s['scrub_gbs_median']:7.0f} (best {s['scrub_gbs_best']:7.0f}) {s['scrub_us_median']:7.1f} us | "
                  f"verify {s['verify_gbs_median']:7.0f} (best {s['verify_gbs_best']:7.0f}) {s['verify_us_median']:7.1f} us | "
                  f"pair {s['pair_gbs_median']:7.0f} {s['pair_us_median']:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
