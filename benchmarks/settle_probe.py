#!/usr/bin/env python3
"""Does the cold call get cheaper when the driver is given time between calls?
(cudaFree / cuMemRelease of ~190 GB leave deferred work that an immediate re-allocation waits for.)"""
import ctypes as C, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from k8s_cc_manager_b200 import _native as N
from k8s_cc_manager_b200 import devices as D
L = N.lib(); L.ccm_init(1)
gpus = [d for d in D.find_gpus()[0] if d.is_gpu()]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
gpus = gpus[:n]
D.scrub_and_verify_many(gpus, 1 << 30)
for settle in (0.0, 0.0, 0.25, 1.0, 3.0, 3.0, 0.0):
    time.sleep(settle)
    reps, wall = D.scrub_and_verify_many(gpus, 0)
    print(f"n={n} settle {settle:4.2f}s wall {wall:7.1f} ms acquire {[round(r.ms_acquire) for r in reps]} "
          f"scrub {[round(r.ms_scrub) for r in reps]} verify {[round(r.ms_verify) for r in reps]} "
          f"release {[round(r.ms_release) for r in reps]}", flush=True)
