#!/usr/bin/env python3
"""A/B on N GPUs in one process: cold concurrent gate with per-GPU turns for the driver's VMM
calls (CCM_VMM_TURNS=1) vs free-for-all (0).  8 reps each, twice, interleaved."""
import os, statistics, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from k8s_cc_manager_b200 import _native as N, devices as D
L = N.lib(); L.ccm_init(1)
gpus = [d for d in D.find_gpus()[0] if d.is_gpu()]
D.scrub_and_verify_many(gpus, 1 << 30)
for turns in ("1", "0", "1", "0"):
    os.environ["CCM_VMM_TURNS"] = turns
    walls = []
    for i in range(8):
        reps, wall = D.scrub_and_verify_many(gpus, 0)
        assert all(r.clean for r in reps)
        walls.append(wall)
    print("gpus", len(gpus), "turns", turns, "walls", [round(w) for w in walls], "median", round(statistics.median(walls)),
          "min", round(min(walls)), flush=True)
