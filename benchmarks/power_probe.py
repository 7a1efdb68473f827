#!/usr/bin/env python3
"""Board power while each scrub / verify variant streams the full arena for ~3 s (NVML)."""
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
        acc.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.02)


def run(name, fn, seconds=3.0):
    ms = C.c_float()
    fn(ms)
    stop, acc = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, acc)); th.start()
    t0, n, tot = time.perf_counter(), 0, 0.0
    while time.perf_counter() - t0 < seconds:
        fn(ms); n += 1; tot += ms.value
    stop.set(); th.join()
    acc = acc[len(acc) // 3:]   # drop the ramp
    out[name] = {"gbs": ai.bytes * n / tot / 1e6, "power_w_median": statistics.median(a[0] for a in acc),
                 "power_w_max": max(a[0] for a in acc), "sm_mhz_median": statistics.median(a[1] for a in acc), "launches": n}
    print(name, json.dumps(out[name]), flush=True)


nz = C.c_uint64()
time.sleep(1.0)
idle = [pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0 for _ in range(10)]
out["idle_w"] = statistics.median(idle)
for name, v in (("scrub_st256", N.SCRUB_ST256), ("scrub_tma", N.SCRUB_TMA), ("scrub_st128", N.SCRUB_ST128), ("scrub_memset", N.SCRUB_MEMSET)):
    run(name, lambda ms, v=v: L.ccm_arena_scrub(0, v, None, None, C.byref(ms)))
for name, v in (("verify_ld256", N.VERIFY_LD256), ("verify_ld128", N.VERIFY_LD128)):
    run(name, lambda ms, v=v: L.ccm_arena_verify(0, v, None, None, C.byref(nz), C.byref(ms)))
L.ccm_arena_release(0)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/power_probe_r1.json").write_text(json.dumps(out, indent=1))
