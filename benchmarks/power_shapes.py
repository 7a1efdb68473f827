#!/usr/bin/env python3
"""Throughput AND board power of scrub_st256 launch shapes with few warps per SM (the store path
needs only ~1 store instruction per 32 clk per SM to saturate the 32 B/clk crossbar egress)."""
import ctypes as C, json, statistics, sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pynvml
from k8s_cc_manager_b200 import _native as N

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
L = N.lib(); assert L.ccm_init(1) == 0
ai = N.ArenaInfo(); assert L.ccm_arena_acquire(0, 0, C.byref(ai)) == 0
out = {}


def sample(stop, acc):
    while not stop.is_set():
        acc.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
        time.sleep(0.02)


def run(name, variant, cfg, seconds=2.0, verify=False):
    ms, nz = C.c_float(), C.c_uint64()
    call = (lambda: L.ccm_arena_verify(0, variant, C.byref(cfg) if cfg else None, None, C.byref(nz), C.byref(ms))) if verify \
        else (lambda: L.ccm_arena_scrub(0, variant, C.byref(cfg) if cfg else None, None, C.byref(ms)))
    assert call() == 0, N.last_error()
    stop, acc = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, acc)); th.start()
    t0, n, tot = time.perf_counter(), 0, 0.0
    while time.perf_counter() - t0 < seconds:
        assert call() == 0
        n += 1; tot += ms.value
    stop.set(); th.join()
    acc = acc[len(acc) // 3:]
    out[name] = {"gbs": round(ai.bytes * n / tot / 1e6, 1), "power_w": round(statistics.median(acc), 1)}
    print(name, out[name], flush=True)


run("tma default", N.SCRUB_TMA, None)
run("st256 default 1x512x4", N.SCRUB_ST256, None)
for th_, un, chunk in ((32, 8, 65536), (32, 16, 65536), (64, 8, 65536), (64, 16, 131072), (128, 8, 131072),
                       (128, 16, 131072), (256, 8, 131072), (256, 4, 131072)):
    run(f"st256 1x{th_}x{un} chunk{chunk}", N.SCRUB_ST256, N.launch_cfg(1, th_, chunk, un, 1, 2))
for th_, un, chunk in ((256, 4, 131072), (512, 4, 131072), (256, 8, 131072), (512, 2, 65536)):
    run(f"ld256 1x{th_}x{un} chunk{chunk}", N.VERIFY_LD256, N.launch_cfg(1, th_, chunk, un, 3, 2), verify=True)
run("ld256 default 1x1024x4", N.VERIFY_LD256, None, verify=True)
L.ccm_arena_release(0)
Path("gpurun_out/power_shapes_r1.json").write_text(json.dumps(out, indent=1))
