#!/usr/bin/env python3
"""Cold scrub gate, N GPUs from ONE process (the manager's launcher: ccm_scrub_verify_many):
time to verdict, its phases, and the deferred HBM release, per call.

  python benchmarks/cold_gate_probe.py [--gpus N] [--calls K] [--gap-s S] [--sync-release]

--gap-s: idle time between calls (a daemon has seconds to minutes between transitions; the
bench's back-to-back loop is the worst case for the driver's free/alloc backlog)."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, max(0, round(q * (len(xs) - 1))))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--calls", type=int, default=10)
    ap.add_argument("--gap-s", type=float, default=0.0)
    ap.add_argument("--sync-release", action="store_true")
    ap.add_argument("--bytes", type=int, default=0)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    if args.sync_release:
        os.environ["CCM_ASYNC_RELEASE"] = "0"
    from k8s_cc_manager_b200 import _native as N
    from k8s_cc_manager_b200 import devices as D
    assert N.lib().ccm_init(N.BACKEND_CUDASIM) == 0
    gpus = [d for d in D.find_gpus()[0] if d.is_gpu()]
    if args.gpus:
        gpus = gpus[:args.gpus]
    D.scrub_and_verify_many(gpus, 1 << 30)          # contexts, kernels loaded
    for g in gpus:
        g.wait_scrub_released()
    rows = []
    for i in range(args.calls):
        time.sleep(args.gap_s)
        t0 = time.perf_counter()
        reps, wall = D.scrub_and_verify_many(gpus, args.bytes)
        verdict_ms = (time.perf_counter() - t0) * 1e3
        rel = [g.wait_scrub_released() for g in gpus]
        cycle_ms = (time.perf_counter() - t0) * 1e3
        assert all(r.clean for r in reps), [r.status for r in reps]
        rows.append({"verdict_ms": verdict_ms, "cycle_ms": cycle_ms,
                     "acquire": [round(r.ms_acquire, 1) for r in reps], "span": [round(r.ms_gpu_span, 1) for r in reps],
                     "scrub": [round(r.ms_scrub, 1) for r in reps], "verify": [round(r.ms_verify, 1) for r in reps],
                     "total": [round(r.ms_total, 1) for r in reps], "wait_prev": [round(r.ms_release_wait, 1) for r in reps],
                     "release": [round(x[0], 1) for x in rel], "chunks": reps[0].segments,
                     "unreached_mib": [r.bytes_unreached >> 20 for r in reps], "coverage": round(reps[0].coverage, 5)})
        print(json.dumps(rows[-1]), flush=True)
    v = [r["verdict_ms"] for r in rows]
    c = [r["cycle_ms"] for r in rows]
    summary = {"tag": args.tag, "gpus": len(gpus), "calls": len(rows), "gap_s": args.gap_s, "sync_release": args.sync_release,
               "bytes_per_gpu": reps[0].bytes_scrubbed,
               "verdict_ms": {"median": statistics.median(v), "p10": pct(v, 0.1), "p90": pct(v, 0.9), "min": min(v), "max": max(v)},
               "cycle_ms_incl_release": {"median": statistics.median(c), "p10": pct(c, 0.1), "p90": pct(c, 0.9)}}
    print("SUMMARY " + json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
