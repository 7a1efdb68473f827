"""CC_SCRUB_ISOLATION=process: the gate runs in a short-lived worker process."""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

import scenarios as SC
from helpers import build_native_world

ROOT = Path(__file__).resolve().parents[1]


def worker(args, env_extra):
    env = dict(os.environ, PYTHONPATH=str(ROOT), **env_extra)
    return subprocess.run([sys.executable, "-m", "k8s_cc_manager_b200.scrub_worker", *args], capture_output=True,
                          text=True, env=env, timeout=300)


def test_worker_reports_no_cuda_and_exits_3():
    proc = worker(["--bdf", SC.GPU_BDFS[0], "--bdf", SC.GPU_BDFS[3], "--bytes", "1048576"],
                  {"CCM_BACKEND": "sim", "CCM_SIM_BIND_CUDA": "0"})
    assert proc.returncode == 3, proc.stderr[-2000:]
    out = json.loads(proc.stdout.splitlines()[-1])
    assert [r["bdf"] for r in out["reports"]] == [SC.GPU_BDFS[0], SC.GPU_BDFS[3]]
    assert all(r["status"] == -9 and r["bytes_scrubbed"] == 0 for r in out["reports"])


def test_worker_usage_errors():
    assert worker([], {"CCM_BACKEND": "sim"}).returncode == 2
    proc = worker(["--bdf", "0000:ff:00.0"], {"CCM_BACKEND": "sim"})
    assert proc.returncode == 1 and "unknown GPU" in proc.stderr


def test_manager_process_isolation_fails_closed_without_cuda(monkeypatch):
    import kubernetes
    from k8s_cc_manager_b200 import manager
    build_native_world(SC.scenario("w", gpus_=SC.gpus(2), modes=[]))
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, {})
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "false")
    monkeypatch.setenv("CC_SCRUB_ISOLATION", "process")
    monkeypatch.setenv("CCM_BACKEND", "sim")
    monkeypatch.setenv("CCM_SIM_GPUS", "2")
    monkeypatch.setenv("CCM_SIM_BIND_CUDA", "0")
    mgr = manager.CCManager(SC.NODE, "on", True)
    assert mgr.scrub_isolation == "process"
    assert mgr.set_cc_mode("on") is False
    assert c.labels(SC.NODE)["nvidia.com/cc.mode.state"] == "failed"
    reports = mgr.last_transition["scrub"]
    assert len(reports) == 2 and all(r.status == -9 for r in reports)
    monkeypatch.setenv("CC_SCRUB_ISOLATION", "container")
    with pytest.raises(ValueError):
        manager.CCManager(SC.NODE, "on", True)


@pytest.mark.gpu
def test_manager_process_isolation_with_real_scrub(monkeypatch):
    import ctypes as C
    import kubernetes
    from k8s_cc_manager_b200 import _native as N, manager
    assert N.lib().ccm_init(N.BACKEND_CUDASIM) == 0
    N.lib().ccm_sim_set(-1, b"cc_mode", 0)
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, {})
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "false")
    monkeypatch.setenv("CC_SCRUB_ISOLATION", "process")
    monkeypatch.setenv("CCM_BACKEND", "cudasim")
    mgr = manager.CCManager(SC.NODE, "on", True, scrub_bytes=2 << 30)
    assert mgr.set_cc_mode("on") is True
    assert c.labels(SC.NODE)["nvidia.com/cc.mode.state"] == "on"
    reports = mgr.last_transition["scrub"]
    assert reports and all(r.clean and r.bytes_scrubbed == 2 << 30 for r in reports)
    print(f"\nprocess-isolated gate: {mgr.last_transition['phase_seconds']['scrub']*1e3:.0f} ms for {len(reports)} GPU(s)")


def test_native_cli_matches_python_worker():
    """ccm-scrub (C++) and scrub_worker (Python) print the same report fields."""
    cli = ROOT / "k8s_cc_manager_b200" / "ccm-scrub"
    assert cli.exists(), "build.py must produce the native CLI next to libccm.so"
    env = dict(os.environ, CCM_BACKEND="sim", CCM_SIM_BIND_CUDA="0")
    args = ["--bdf", SC.GPU_BDFS[1].upper(), "--bytes", "4096"]
    a = subprocess.run([str(cli), *args], capture_output=True, text=True, env=env, timeout=60)
    b = worker(["--bdf", SC.GPU_BDFS[1], "--bytes", "4096"], {"CCM_BACKEND": "sim", "CCM_SIM_BIND_CUDA": "0"})
    assert a.returncode == 3 and b.returncode == 3
    ra, rb = json.loads(a.stdout.splitlines()[-1])["reports"][0], json.loads(b.stdout.splitlines()[-1])["reports"][0]
    assert set(ra) == set(rb)
    assert {k: ra[k] for k in ("bdf", "status", "bytes_requested", "nonzero_bytes")} == \
           {k: rb[k] for k in ("bdf", "status", "bytes_requested", "nonzero_bytes")}
    assert subprocess.run([str(cli)], capture_output=True).returncode == 2
    assert subprocess.run([str(cli), "--bdf", "0000:ff:00.0"], capture_output=True, env=env).returncode == 1
    allrun = subprocess.run([str(cli), "--all", "--backend", "sim"], capture_output=True, text=True,
                            env=dict(env, CCM_SIM_GPUS="3"), timeout=60)
    assert len(json.loads(allrun.stdout.splitlines()[-1])["reports"]) == 3


@pytest.mark.parametrize("flavour", ["native", "python"])
def test_manager_uses_either_worker(flavour, monkeypatch):
    import kubernetes
    from k8s_cc_manager_b200 import manager
    build_native_world(SC.scenario("w", gpus_=SC.gpus(2), modes=[]))
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, {})
    for k, v in dict(EVICT_OPERATOR_COMPONENTS="false", CC_SCRUB_ISOLATION="process", CC_SCRUB_WORKER=flavour,
                     CCM_BACKEND="sim", CCM_SIM_GPUS="2", CCM_SIM_BIND_CUDA="0").items():
        monkeypatch.setenv(k, v)
    mgr = manager.CCManager(SC.NODE, "on", True)
    assert mgr.set_cc_mode("on") is False
    assert [r.status for r in mgr.last_transition["scrub"]] == [-9, -9]
