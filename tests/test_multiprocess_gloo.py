"""N>1 host logic of bench.py on CPU: world_size 2 over gloo (127.0.0.1).  The data path
has no collective; ranks only agree on max(time) and sum(bytes), and the reference arm
lets rank 0 alone print."""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

WORKER = r"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from k8s_cc_manager_b200.aggregate import aggregate_job
dist.init_process_group("gloo")
rank = dist.get_rank()
# rank r pretends: region (r+1)*1000 bytes, 4 steps in (10 + 5r) ms, 8 launches
out = aggregate_job(dist, device="cpu", region_bytes=(rank + 1) * 1000, steps=4,
                    elapsed_ms=10.0 + 5.0 * rank, launches=8)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
"""


def test_two_rank_aggregation_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=str(ROOT), MASTER_ADDR="127.0.0.1")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = [l for l in proc.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2
    assert out["ms"] == 15.0                        # MAX over ranks, never the mean
    assert out["total_region_bytes"] == 3000        # SUM over ranks (weak scaling: per-GPU work fixed)
    assert out["launches"] == 16
    # whole-job value: (bytes zeroed + bytes read back) * steps / max time
    assert abs(out["value_gbs"] - 2 * 3000 * 4 / 15e-3 / 1e9) < 1e-12


def test_reference_arm_prints_once_under_torchrun(tmp_path):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29572", str(ROOT / "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--cpu-sample-gib", "0.125"],
        capture_output=True, text=True, env=env, timeout=240)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [json.loads(l) for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    ref = lines[0]
    assert ref["impl"] == "reference" and ref["n_gpus"] == 2 and ref["unit"] == "GB/s"
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "port"
    assert ref["value"] > 0
