#!/usr/bin/env python3
"""Node-level cold gate on N GPUs: which launcher structure reaches the verdict first?

  A  one process, one thread per GPU, pipelined (map chunk i+1 while chunk i is scrubbed/read back)
  B  one process, threads, CCM_MAP_FIRST=1 (map everything, then two launches)
  C  one process, threads, pipelined scrub but ONE verify at the end (CCM_INTERLEAVE_VERIFY=0)
  F  one process, threads, COARSE pipeline: 64 GiB chunks (3 + tail instead of 23), CCM_MAP_FIRST=0
  G  one process, threads, coarse pipeline: 32 GiB chunks
  D  one fresh worker process PER GPU (native ccm-scrub --bdf X), all started together
  E  one fresh worker process for all GPUs (threads inside)

A-C reuse warm CUDA contexts; D/E pay cuInit + context creation in every run (what a daemon that drops
its contexts after each gate pays anyway).  Prints one JSON line per mode."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def spread(xs):
    xs = sorted(xs)
    return {"median": statistics.median(xs), "min": xs[0], "max": xs[-1], "n": len(xs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--modes", default="A,B,C,D,E")
    args = ap.parse_args()
    from k8s_cc_manager_b200 import _native as N
    from k8s_cc_manager_b200 import devices as D
    assert N.lib().ccm_init(N.BACKEND_CUDASIM) == 0
    gpus = [d for d in D.find_gpus()[0] if d.is_gpu()]
    if args.gpus:
        gpus = gpus[:args.gpus]
    bdfs = [g.bdf for g in gpus]
    cli = ROOT / "k8s_cc_manager_b200" / "ccm-scrub"
    modes = args.modes.split(",")

    def in_process(tag, env):
        for k, v in env.items():
            os.environ[k] = v
        try:
            D.scrub_and_verify_many(gpus, 1 << 30)
            [g.wait_scrub_released() for g in gpus]
            v, c, acq, span = [], [], [], []
            for _ in range(args.reps):
                t0 = time.perf_counter()
                reps, _ = D.scrub_and_verify_many(gpus, 0)
                v.append(time.perf_counter() - t0)
                [g.wait_scrub_released() for g in gpus]
                c.append(time.perf_counter() - t0)
                assert all(r.clean for r in reps)
                acq.append(max(r.ms_acquire for r in reps))
                span.append(max(r.ms_gpu_span for r in reps))
            print(json.dumps({"mode": tag, "gpus": len(gpus), "verdict_s": spread(v), "cycle_s": spread(c),
                              "acquire_host_ms_max": spread(acq), "gpu_span_ms_max": spread(span), "env": env}), flush=True)
        finally:
            for k in env:
                os.environ.pop(k, None)

    if "A" in modes:
        in_process("A in-process threads, pipelined + interleaved verify (23 chunks)", {"CCM_MAP_FIRST": "0"})
    if "B" in modes:
        in_process("B in-process threads, map first", {"CCM_MAP_FIRST": "1"})
    if "C" in modes:
        in_process("C in-process threads, pipelined scrub, one verify", {"CCM_INTERLEAVE_VERIFY": "0"})
    if "F" in modes:
        in_process("F in-process threads, coarse pipeline 64 GiB chunks",
                   {"CCM_MAP_FIRST": "0", "CCM_VMM_FIRST_CHUNK_MB": "65536", "CCM_VMM_CHUNK_MB": "65536"})
    if "G" in modes:
        in_process("G in-process threads, coarse pipeline 32 GiB chunks",
                   {"CCM_MAP_FIRST": "0", "CCM_VMM_FIRST_CHUNK_MB": "32768", "CCM_VMM_CHUNK_MB": "32768"})
    # contexts of THIS process must not sit on the GPUs while the workers measure "all free HBM"
    D.release_cuda_contexts(gpus)
    env = dict(os.environ, CCM_BACKEND="cudasim")
    if "D" in modes:
        v, x = [], []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            procs = [subprocess.Popen([str(cli), "--bdf", b], stdout=subprocess.PIPE, text=True, env=env) for b in bdfs]
            lines = [p.stdout.readline() for p in procs]          # the JSON line is printed right after the verdict
            v.append(time.perf_counter() - t0)
            rcs = [p.wait() for p in procs]
            x.append(time.perf_counter() - t0)
            assert all(rc == 0 for rc in rcs), rcs
            reps = [json.loads(ln)["reports"][0] for ln in lines]
        print(json.dumps({"mode": "D one fresh worker process per GPU", "gpus": len(gpus), "verdict_s": spread(v),
                          "all_exited_s": spread(x), "last_total_ms": [round(r["ms_total"], 1) for r in reps],
                          "last_acquire_ms": [round(r["ms_acquire"], 1) for r in reps]}), flush=True)
    if "E" in modes:
        v, x = [], []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            cmd = [str(cli)]
            for b in bdfs:
                cmd += ["--bdf", b]
            p = subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True, env=env)
            line = p.stdout.readline()
            v.append(time.perf_counter() - t0)
            rc = p.wait()
            x.append(time.perf_counter() - t0)
            assert rc == 0
        print(json.dumps({"mode": "E one fresh worker process, threads inside", "gpus": len(gpus), "verdict_s": spread(v),
                          "all_exited_s": spread(x)}), flush=True)


if __name__ == "__main__":
    main()
