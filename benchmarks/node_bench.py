#!/usr/bin/env python3
"""Whole-node measurement in ONE process (the product's launcher shape), BASELINE.json
configs[2], [3]/[4] and the region sweep, next to the constructed baselines of BASELINE.md §3.

  NEW  ccm_scrub_verify_many: one host thread + primary context + stream per GPU, all
       GPUs concurrently, no collective.
  B2   serial library path: GPUs one after another, cudaMemsetAsync + torch.count_nonzero.
  B1   "serial CPU-driven" straw man north_star names: GPUs one after another,
       cudaMemset, then verify on the HOST: D2H in 256 MiB pinned chunks + np.count_nonzero.
       Bounded sample (PCIe + one core: a full 190 GB GPU would take ~a minute each).
  B0/T node transition through CCManager.set_cc_mode with the eviction gate on:
       off -> on -> devtools -> off; CC registers / reset / boot / k8s API SIMULATED
       (reported as such), HBM scrub gate real.

Neither baseline exists in the reference (it has no scrub, no CUDA: SURVEY.md §0); they
mirror its serial per-GPU loop structure (reference main.py:504-529) and are labelled
"constructed baseline" everywhere.  Writes gpurun_out/node_bench_<tag>.json.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "fakes"))

from k8s_cc_manager_b200 import _native as N  # noqa: E402
from k8s_cc_manager_b200 import devices as D  # noqa: E402


def new_concurrent(gpus, nbytes, reps):
    best = None
    for _ in range(reps):
        reports, wall_ms = D.scrub_and_verify_many(gpus, nbytes)
        assert all(r.clean for r in reports), [(r.bdf, r.status, r.nonzero_bytes) for r in reports]
        total = sum(r.bytes_scrubbed for r in reports)
        row = {
            "wall_ms": wall_ms,
            "bytes_per_gpu": [r.bytes_scrubbed for r in reports],
            "coverage": [r.coverage for r in reports],
            "scrub_gbs": [r.scrub_gbs for r in reports],
            "verify_gbs": [r.verify_gbs for r in reports],
            "ms_total_per_gpu": [r.ms_total for r in reports],
            "ms_acquire": [r.ms_acquire for r in reports], "ms_release": [r.ms_release for r in reports],
            "aggregate_e2e_gbs": 2.0 * total / wall_ms / 1e6,
            "aggregate_kernel_gbs": sum(2.0 * r.bytes_scrubbed / (r.ms_scrub + r.ms_verify) / 1e6 for r in reports),
        }
        if best is None or row["wall_ms"] < best["wall_ms"]:
            best = row
    best["scrub_gbs_min_med_max"] = [min(best["scrub_gbs"]), statistics.median(best["scrub_gbs"]), max(best["scrub_gbs"])]
    best["verify_gbs_min_med_max"] = [min(best["verify_gbs"]), statistics.median(best["verify_gbs"]), max(best["verify_gbs"])]
    return best


def resident_kernels(gpus, nbytes, reps=3):
    """Kernel-only GB/s with the region resident (arena API), all GPUs concurrently from
    one thread each: best-of-`reps` CUDA-event time of the scrub and of the verify launch."""
    import threading
    L = N.lib()
    res = [None] * len(gpus)

    def work(i, g):
        ai = N.ArenaInfo()
        rc = L.ccm_arena_acquire(g.index, nbytes, C.byref(ai))
        assert rc == 0, N.last_error()
        ms, nz = C.c_float(), C.c_uint64()
        best_s = best_v = 1e30
        for k in range(reps + 1):
            assert L.ccm_arena_scrub(g.index, N.SCRUB_AUTO, None, None, C.byref(ms)) == 0, N.last_error()
            if k:
                best_s = min(best_s, ms.value)
            assert L.ccm_arena_verify(g.index, N.VERIFY_AUTO, None, None, C.byref(nz), C.byref(ms)) == 0
            assert nz.value == 0
            if k:
                best_v = min(best_v, ms.value)
        L.ccm_arena_release(g.index)
        res[i] = (ai.bytes / best_s / 1e6, ai.bytes / best_v / 1e6, ai.bytes)

    th = [threading.Thread(target=work, args=(i, g)) for i, g in enumerate(gpus)]
    [t.start() for t in th]
    [t.join() for t in th]
    s = [r[0] for r in res]
    v = [r[1] for r in res]
    return {"kernel_scrub_gbs_min_med_max": [min(s), statistics.median(s), max(s)],
            "kernel_verify_gbs_min_med_max": [min(v), statistics.median(v), max(v)],
            "kernel_scrub_gbs_sum": sum(s), "kernel_verify_gbs_sum": sum(v)}


def b2_serial_library(n_gpus, nbytes):
    """GPUs one after another: cudaMemsetAsync + torch.count_nonzero (library-only path)."""
    import torch
    L = N.lib()
    t0 = time.perf_counter()
    total = 0
    for g in range(n_gpus):
        torch.cuda.set_device(g)
        free, _ = torch.cuda.mem_get_info(g)
        want = nbytes if nbytes else (free - (4 << 30)) // (2 << 20) * (2 << 20)
        buf = torch.empty(want, dtype=torch.uint8, device=f"cuda:{g}")
        rc = L.ccm_region_scrub(g, C.c_void_p(buf.data_ptr()), want, N.SCRUB_MEMSET, None, None, None)
        assert rc == 0, N.last_error()
        torch.cuda.synchronize(g)
        nz = 0
        step = 256 << 20  # torch.count_nonzero on uint8 materialises an 8x temporary
        for off in range(0, want, step):
            nz += int(torch.count_nonzero(buf[off:off + step]))
        assert nz == 0
        total += want
        del buf
        torch.cuda.empty_cache()
    dt = time.perf_counter() - t0
    return {"wall_ms": dt * 1e3, "bytes_total": total, "aggregate_e2e_gbs": 2.0 * total / dt / 1e9,
            "what": "constructed baseline B2: serial per GPU, cudaMemsetAsync + torch.count_nonzero (256 MiB slices), torch allocator"}


def b1_serial_cpu_driven(n_gpus, sample_bytes):
    """GPUs one after another: cudaMemset, then HOST verify via 256 MiB pinned D2H chunks."""
    import numpy as np
    import torch
    L = N.lib()
    chunk = 256 << 20
    pinned = torch.empty(chunk, dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    t0 = time.perf_counter()
    total = 0
    for g in range(n_gpus):
        torch.cuda.set_device(g)
        buf = torch.empty(sample_bytes, dtype=torch.uint8, device=f"cuda:{g}")
        buf.fill_(0xA5)
        torch.cuda.synchronize(g)  # the library launches on its own stream
        rc = L.ccm_region_scrub(g, C.c_void_p(buf.data_ptr()), sample_bytes, N.SCRUB_MEMSET, None, None, None)
        assert rc == 0, N.last_error()
        torch.cuda.synchronize(g)
        nz = 0
        for off in range(0, sample_bytes, chunk):
            n = min(chunk, sample_bytes - off)
            pinned[:n].copy_(buf[off:off + n])
            torch.cuda.synchronize(g)
            nz += int(np.count_nonzero(host[:n]))
        assert nz == 0
        total += sample_bytes
        del buf
    dt = time.perf_counter() - t0
    return {"wall_ms": dt * 1e3, "bytes_total": total, "aggregate_e2e_gbs": 2.0 * total / dt / 1e9,
            "sample_bytes_per_gpu": sample_bytes,
            "what": "constructed baseline B1: serial per GPU, cudaMemset + host verify (256 MiB pinned D2H + "
                    "np.count_nonzero, 1 thread); bounded sample, throughput extrapolates linearly"}


def node_transition(n_gpus, max_parallel, reset_ms, boot_ms):
    import logging

    import kubernetes
    from k8s_cc_manager_b200 import manager
    from k8s_cc_manager_b200.drain_gate import COMPONENT_LABELS
    logging.disable(logging.CRITICAL)
    L = N.lib()
    L.ccm_sim_set(-1, b"cc_mode", 0)
    L.ccm_sim_set(-1, b"reset_ms", reset_ms)
    L.ccm_sim_set(-1, b"boot_ms", boot_ms)
    c = kubernetes.reset_cluster()
    c.add_node("node", {k: "true" for k in COMPONENT_LABELS})
    os.environ["EVICT_OPERATOR_COMPONENTS"] = "true"
    os.environ.setdefault("CC_RELEASE_CUDA_CONTEXT", "false")   # contexts are shared with the sweep in this process
    devs = D.find_gpus()[0][:n_gpus]
    mgr = manager.CCManager("node", "on", True, max_parallel=max_parallel, device_source=lambda: (devs, len(devs)))
    out = {}
    for mode in ("on", "devtools", "off"):
        t0 = time.perf_counter()
        ok = mgr.set_cc_mode(mode)
        dt = time.perf_counter() - t0
        reps = mgr.last_transition.get("scrub") or []
        out[f"to_{mode}"] = {"ok": bool(ok), "wall_s": dt, "label": c.labels("node").get("nvidia.com/cc.mode.state"),
                             "ready": c.labels("node").get("nvidia.com/cc.ready.state"),
                             "phase_seconds": mgr.last_transition.get("phase_seconds"),
                             "scrubbed_gb_per_gpu": [r.bytes_scrubbed / 1e9 for r in reps] if isinstance(reps, list) else reps}
    logging.disable(logging.NOTSET)
    L.ccm_sim_set(-1, b"reset_ms", 0)
    L.ccm_sim_set(-1, b"boot_ms", 0)
    out["component_labels_restored"] = all(c.labels("node")[k] == "true" for k in COMPONENT_LABELS)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--sweep", default="1,2,4,8,16,32,64,128,0", help="region sizes in GB (0 = max)")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--b1-sample-gib", type=float, default=4.0)
    ap.add_argument("--skip-baselines", action="store_true")
    ap.add_argument("--out", default="gpurun_out")
    args = ap.parse_args()

    L = N.lib()
    assert L.ccm_init(N.BACKEND_CUDASIM) == 0, N.last_error()
    all_gpus = [d for d in D.find_gpus()[0] if d.is_gpu()]
    n_max = len(all_gpus) if args.gpus <= 0 else min(args.gpus, len(all_gpus))
    counts = [n for n in (1, 2, 4, 8) if n <= n_max]
    out = {"host_cpus": os.cpu_count(), "gpus_visible": len(all_gpus), "gpu_names": [g.name for g in all_gpus],
           "sweep": [], "baselines": {}, "transition": {}}
    print(f"{len(all_gpus)} GPUs, host cpus {os.cpu_count()}", flush=True)

    # warm every context once (context creation is not part of a transition on a running daemon)
    D.scrub_and_verify_many(all_gpus[:n_max], 1 << 30)

    for n in counts:
        for gb in [float(x) for x in args.sweep.split(",")]:
            nbytes = int(gb * 1e9) // (2 << 20) * (2 << 20)
            row = new_concurrent(all_gpus[:n], nbytes, args.reps)
            row.update(n_gpus=n, region_gb=gb or "max")
            row.update(resident_kernels(all_gpus[:n], nbytes))
            out["sweep"].append(row)
            print(json.dumps({k: row[k] for k in ("n_gpus", "region_gb", "wall_ms", "aggregate_e2e_gbs",
                                                  "kernel_scrub_gbs_min_med_max", "kernel_verify_gbs_min_med_max")}), flush=True)
    if not args.skip_baselines:
        for n in counts:
            out["baselines"][f"B2_n{n}_max"] = b2_serial_library(n, 0)
            print("B2", n, json.dumps(out["baselines"][f"B2_n{n}_max"]), flush=True)
        out["baselines"][f"B1_n{n_max}"] = b1_serial_cpu_driven(n_max, int(args.b1_sample_gib * 2**30))
        print("B1", json.dumps(out["baselines"][f"B1_n{n_max}"]), flush=True)

    # node transition: scrub gate real; registers/API simulated.  Also with a non-zero
    # SIMULATED reset+boot latency to show what the concurrent phases buy (labelled).
    for label, par, rst, boot in (("concurrent_sim0", 0, 0, 0), ("serial_sim0", 1, 0, 0),
                                  ("concurrent_sim_reset200_boot1000", 0, 200, 1000),
                                  ("serial_sim_reset200_boot1000", 1, 200, 1000)):
        out["transition"][label] = node_transition(n_max, par, rst, boot)
        print("transition", label, json.dumps({k: (v["wall_s"] if isinstance(v, dict) else v)
                                               for k, v in out["transition"][label].items()}), flush=True)
    out["notes"] = ["CC registers, reset/boot latency and the k8s API are SIMULATED in 'transition' (the box's GPUs are "
                    "bound to the driver and cannot be reset); only the HBM scrub gate is real.",
                    "B1/B2 are constructed baselines (the reference has no scrub path)."]
    os.makedirs(args.out, exist_ok=True)
    path = Path(args.out) / f"node_bench_{args.tag}.json"
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path)


if __name__ == "__main__":
    main()
