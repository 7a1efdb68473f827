#!/usr/bin/env python3
"""Small, fixed launch sequence for ncu (run under gpurun, one GPU).

  python benchmarks/profile_target.py --gib 4 --seq memset,st256,st256:4:512:4,tma,ld256,vauto

Each item is variant[:ctas_per_sm:threads:unroll[:policy[:tile[:schedule]]]]; scrub variants
zero the arena, verify variants (ld128, ld256, vauto) count it.  Prints CUDA-event
times so the same command can be timed outside the profiler.
"""
from __future__ import annotations

import argparse
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from k8s_cc_manager_b200 import _native as N  # noqa: E402

SCRUB = {"memset": N.SCRUB_MEMSET, "st128": N.SCRUB_ST128, "st256": N.SCRUB_ST256, "tma": N.SCRUB_TMA,
         "auto": N.SCRUB_AUTO}
VERIFY = {"ld128": N.VERIFY_LD128, "ld256": N.VERIFY_LD256, "vauto": N.VERIFY_AUTO}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--seq", default="memset,auto,vauto")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--warm", type=int, default=0)
    args = ap.parse_args()
    L = N.lib()
    ai = N.ArenaInfo()
    rc = L.ccm_arena_acquire(0, int(args.gib * 2**30) if args.gib > 0 else 0, C.byref(ai))
    assert rc == 0, N.last_error()
    ms = C.c_float()
    nz = C.c_uint64()
    print(f"arena {ai.bytes} bytes in {ai.segments} segment(s)", flush=True)
    for item in args.seq.split(","):
        parts = item.split(":")
        name = parts[0]
        nums = [int(x) for x in parts[1:]] + [0] * 6
        cfg = N.launch_cfg(nums[0], nums[1], nums[4], nums[2], nums[3], nums[5]) if len(parts) > 1 else None
        cfgp = C.byref(cfg) if cfg else None
        for i in range(args.warm + args.reps):
            if name in SCRUB:
                rc = L.ccm_arena_scrub(0, SCRUB[name], cfgp, None, C.byref(ms))
            else:
                rc = L.ccm_arena_verify(0, VERIFY[name], cfgp, None, C.byref(nz), C.byref(ms))
            assert rc == 0, N.last_error()
            if i >= args.warm:
                print(f"{item:28s} {ms.value:9.4f} ms  {ai.bytes / ms.value / 1e6:8.1f} GB/s  nz={nz.value}", flush=True)
    L.ccm_arena_release(0)


if __name__ == "__main__":
    main()
