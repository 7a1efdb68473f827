#!/usr/bin/env python3
"""Launch-shape sweep of the scrub / verify kernels on ONE B200 (run under gpurun).

For every variant and launch shape: correctness first (0xA5 poison -> verify ==
bytes, scrub -> verify == 0), then best-of-N CUDA-event time over a region far
larger than L2.  Writes gpurun_out/tune_<tag>.json; the winners become the library
defaults in csrc/ccm_scrub.cu (recorded in DESIGN.md §5 and profiles/).
"""
from __future__ import annotations

import argparse
import ctypes as C
import itertools
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from k8s_cc_manager_b200 import _native as N  # noqa: E402


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {N.strerror(rc)}: {N.last_error()}")


def time_scrub(L, dev, variant, cfg, reps):
    ms = C.c_float()
    best = 1e30
    for _ in range(reps):
        check(L.ccm_arena_scrub(dev, variant, C.byref(cfg) if cfg else None, None, C.byref(ms)), "scrub")
        best = min(best, ms.value)
    return best


def time_verify(L, dev, variant, cfg, reps):
    ms = C.c_float()
    nz = C.c_uint64()
    best = 1e30
    for _ in range(reps):
        check(L.ccm_arena_verify(dev, variant, C.byref(cfg) if cfg else None, None, C.byref(nz), C.byref(ms)), "verify")
        best = min(best, ms.value)
    return best, nz.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=32.0, help="sweep region size in GiB (0 = max)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out")
    args = ap.parse_args()

    L = N.lib()
    n = C.c_int()
    infos = (N.DevInfo * 16)()
    check(L.ccm_enumerate(infos, 16, C.byref(n)), "enumerate")
    dev = 0
    print(f"backend={L.ccm_backend_in_use()} devices={n.value} dev0={infos[0].name.decode()} "
          f"{infos[0].bdf.decode()} total={infos[0].hbm_total_bytes/2**30:.1f} GiB", flush=True)

    out = {"device": infos[0].name.decode(), "hbm_total_bytes": infos[0].hbm_total_bytes, "results": []}

    # ---- acquire / release cost at max size
    ai = N.ArenaInfo()
    t0 = time.perf_counter()
    check(L.ccm_arena_acquire(dev, 0, C.byref(ai)), "acquire max")
    t1 = time.perf_counter()
    out["max_arena"] = {"bytes": ai.bytes, "segments": ai.segments, "free_before": ai.device_free_before,
                        "total": ai.device_total_bytes, "ms_acquire": ai.ms_acquire,
                        "coverage": ai.bytes / ai.device_total_bytes}
    print("max arena:", out["max_arena"], flush=True)
    # full-size correctness + default-shape timing
    check(L.ccm_arena_fill(dev, 0xA5, None), "fill")
    ms_v, nz = time_verify(L, dev, N.VERIFY_AUTO, None, 1)
    assert nz == ai.bytes, f"poisoned arena: expected {ai.bytes} nonzero bytes, got {nz}"
    ms_s = time_scrub(L, dev, N.SCRUB_AUTO, None, 1)
    ms_v0, nz0 = time_verify(L, dev, N.VERIFY_AUTO, None, 1)
    assert nz0 == 0, f"after scrub {nz0} nonzero bytes"
    out["max_default"] = {"ms_scrub": ms_s, "ms_verify_dirty": ms_v, "ms_verify_clean": ms_v0,
                          "scrub_gbs": ai.bytes / ms_s / 1e6, "verify_gbs": ai.bytes / ms_v0 / 1e6}
    print("max default:", out["max_default"], flush=True)
    t2 = time.perf_counter()
    check(L.ccm_arena_release(dev), "release")
    out["max_arena"]["ms_release"] = (time.perf_counter() - t2) * 1e3
    print(f"release {out['max_arena']['ms_release']:.1f} ms", flush=True)

    # ---- sweep region
    want = int(args.gib * 2**30) if args.gib > 0 else 0
    check(L.ccm_arena_acquire(dev, want, C.byref(ai)), "acquire sweep")
    R = ai.bytes
    out["sweep_bytes"] = R
    print(f"sweep region {R/2**30:.1f} GiB", flush=True)

    def rec(kind, variant, cfg, ms, extra=None):
        row = {"kind": kind, "variant": variant, "ctas_per_sm": cfg.ctas_per_sm if cfg else 0,
               "threads": cfg.threads_per_cta if cfg else 0, "unroll": cfg.unroll if cfg else 0,
               "policy": cfg.cache_policy if cfg else 0, "tile": cfg.tile_bytes if cfg else 0,
               "ms": ms, "gbs": R / ms / 1e6}
        if extra:
            row.update(extra)
        out["results"].append(row)
        print(json.dumps(row), flush=True)

    # library bar
    rec("scrub", "memset", None, time_scrub(L, dev, N.SCRUB_MEMSET, None, args.reps))

    pol_all = [1, 2, 3] if not args.quick else [1]
    # ST variants
    for vname, v in (("st256", N.SCRUB_ST256), ("st128", N.SCRUB_ST128)):
        combos = itertools.product([1, 2, 4, 8], [128, 256, 512, 1024], [1, 2, 4, 8], pol_all)
        for cps, th, un, pol in combos:
            if cps * th > 2048 or cps * th < 256:
                continue
            if args.quick and (un not in (2, 4) or th not in (256, 512)):
                continue
            cfg = N.launch_cfg(cps, th, 0, un, pol)
            rec("scrub", vname, cfg, time_scrub(L, dev, v, cfg, args.reps))
    # TMA
    for cps, th, tile, grp, pol in itertools.product([1, 2, 4], [32, 128], [8192, 16384, 32768, 65536],
                                                     [1, 4], [1, 2]):
        if cps * tile > 200 * 1024:
            continue
        cfg = N.launch_cfg(cps, th, tile, grp, pol)
        rec("scrub", "tma", cfg, time_scrub(L, dev, N.SCRUB_TMA, cfg, args.reps))

    # verify: must count exactly on a dirty region too
    check(L.ccm_arena_fill(dev, 0xA5, None), "fill")
    ms, nz = time_verify(L, dev, N.VERIFY_LD256, None, 1)
    assert nz == R, (nz, R)
    rec("verify_dirty", "ld256", None, ms)
    ms, nz = time_verify(L, dev, N.VERIFY_LD128, None, 1)
    assert nz == R, (nz, R)
    rec("verify_dirty", "ld128", None, ms)
    check(L.ccm_arena_scrub(dev, N.SCRUB_AUTO, None, None, None), "scrub")

    for vname, v in (("ld256", N.VERIFY_LD256), ("ld128", N.VERIFY_LD128)):
        for cps, th, un, pol in itertools.product([1, 2, 4, 8], [128, 256, 512, 1024], [1, 2, 4, 8], [1, 2, 3]):
            if cps * th > 2048 or cps * th < 256:
                continue
            if vname == "ld256" and un == 8 and th * cps > 1024:
                continue  # register budget: 64 regs of payload per thread
            if args.quick and (un not in (2, 4) or th not in (256, 512)):
                continue
            cfg = N.launch_cfg(cps, th, 0, un, pol)
            ms, nz = time_verify(L, dev, v, cfg, args.reps)
            assert nz == 0, (vname, cps, th, un, pol, nz)
            rec("verify", vname, cfg, ms)
    check(L.ccm_arena_release(dev), "release")
    out["max_arena"]["ms_release"] = (time.perf_counter() - t2) * 1e3
    print(f"release {out['max_arena']['ms_release']:.1f} ms", flush=True)

    # ---- sweep region
    want = int(args.gib * 2**30) if args.gib > 0 else 0
    check(L.ccm_arena_acquire(dev, want, C.byref(ai)), "acquire sweep")
    R = ai.bytes
    out["sweep_bytes"] = R
    print(f"sweep region {R/2**30:.1f} GiB", flush=True)

    def rec(kind, variant, cfg, ms, extra=None):
        row = {"kind": kind, "variant": variant, "ctas_per_sm": cfg.ctas_per_sm if cfg else 0,
               "threads": cfg.threads_per_cta if cfg else 0, "unroll": cfg.unroll if cfg else 0,
               "policy": cfg.cache_policy if cfg else 0, "tile": cfg.tile_bytes if cfg else 0,
               "ms": ms, "gbs": R / ms / 1e6}
        if extra:
            row.update(extra)
        out["results"].append(row)
        print(json.dumps(row), flush=True)

    # library bar
    rec("scrub", "memset", None, time_scrub(L, dev, N.SCRUB_MEMSET, None, args.reps))

    pol_all = [1, 2, 3] if not args.quick else [1]
    # ST variants
    for vname, v in (("st256", N.SCRUB_ST256), ("st128", N.SCRUB_ST128)):
        combos = itertools.product([1, 2, 4, 8], [128, 256, 512, 1024], [1, 2, 4, 8], pol_all)
        for cps, th, un, pol in combos:
            if cps * th > 2048 or cps * th < 256:
                continue
            if args.quick and (un not in (2, 4) or th not in (256, 512)):
                continue
            cfg = N.launch_cfg(cps, th, 0, un, pol)
            rec("scrub", vname, cfg, time_scrub(L, dev, v, cfg, args.reps))
    # TMA
    for cps, th, tile, grp, pol in itertools.product([1, 2, 4], [32, 128], [8192, 16384, 32768, 65536],
                                                     [1, 4], [1, 2]):
        if cps * tile > 200 * 1024:
            continue
        cfg = N.launch_cfg(cps, th, tile, grp, pol)
        rec("scrub", "tma", cfg, time_scrub(L, dev, N.SCRUB_TMA, cfg, args.reps))

    # verify: must count exactly on a dirty region too
    check(L.ccm_arena_fill(dev, 0xA5, None), "fill")
    ms, nz = time_verify(L, dev, N.VERIFY_LD256, None, 1)
    assert nz == R, (nz, R)
    rec("verify_dirty", "ld256", None, ms)
    ms, nz = time_verify(L, dev, N.VERIFY_LD128, None, 1)
    assert nz == R, (nz, R)
    rec("verify_dirty", "ld128", None, ms)
    check(L.ccm_arena_scrub(dev, N.SCRUB_AUTO, None, None, None), "scrub")

    for vname, v in (("ld256", N.VERIFY_LD256), ("ld128", N.VERIFY_LD128)):
        for cps, th, un, pol in itertools.product([1, 2, 4, 8], [128, 256, 512, 1024], [1, 2, 4, 8], [1, 2, 3]):
            if cps * th > 2048 or cps * th < 256:
                continue
            if vname == "ld256" and un == 8 and th * cps > 1024:
                continue  # register budget: 64 regs of payload per thread
            if args.quick and (un not in (2, 4) or th not in (256, 512)):
                continue
            cfg = N.launch_cfg(cps, th, 0, un, pol)
            ms, nz = time_verify(L, dev, v, cfg, args.reps)
            assert nz == 0, (vname, cps, th, un, pol, nz)
            rec("verify", vname, cfg, ms)
    check(L.ccm_arena_release(dev), "release")

    for kind in ("scrub", "verify"):
        rows = sorted((r for r in out["results"] if r["kind"] == kind), key=lambda r: r["ms"])
        print(f"--- top {kind}")
        for r in rows[:8]:
            print(json.dumps(r))
        out[f"best_{kind}"] = rows[:8]
    out["kernel_launches"] = L.ccm_kernel_launches()
    os.makedirs(args.out, exist_ok=True)
    path = Path(args.out) / f"tune_{args.tag}.json"
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path)


if __name__ == "__main__":
    main()
