"""Process-level smoke of the image entrypoint (`python main.py`, reference
Dockerfile.distroless:70): real argv/env parsing, sim register backend, in-memory API
server, scripted label changes, readiness file, scrub gate policy from the environment."""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

BOOTSTRAP = r"""
import json, os, sys
from types import SimpleNamespace
sys.path[:0] = [os.environ["REPO"], os.path.join(os.environ["REPO"], "tests", "fakes")]
import kubernetes
from kubernetes.watch import WatchScriptExhausted
c = kubernetes.cluster()
c.add_node("node-a", {"nvidia.com/cc.mode": "on", "nvidia.com/gpu.deploy.vfio-manager": "true"})
def ev(label, rv):
    node = SimpleNamespace(metadata=SimpleNamespace(name="node-a", labels={"nvidia.com/cc.mode": label}, resource_version=rv))
    return {"type": "MODIFIED", "object": node}
c.watch_script.append([ev("devtools", "50"), ev("off", "51")])
import main as entry
sys.argv = ["main.py"] + json.loads(os.environ["ARGV"])
try:
    entry.main()
except WatchScriptExhausted:
    pass
from k8s_cc_manager_b200 import _native as N
import ctypes as C
modes = []
for i in range(8):
    m = C.c_int(); N.lib().ccm_query_cc_mode(i, C.byref(m)); modes.append(m.value)
print("RESULT " + json.dumps({"labels": c.labels("node-a"), "modes": modes, "patches": c.verbs().count("patch_node")}))
"""


def run(argv, extra_env, tmp_path):
    env = dict(os.environ, REPO=str(ROOT), ARGV=json.dumps(argv), CCM_BACKEND="sim", CCM_SIM_GPUS="8",
               CC_READINESS_FILE=str(tmp_path / "ready" / ".cc-manager-ctr-ready"), **extra_env)
    env.pop("NODE_NAME", None)
    if "NODE_NAME" in extra_env:
        env["NODE_NAME"] = extra_env["NODE_NAME"]
    return subprocess.run([sys.executable, "-c", BOOTSTRAP], capture_output=True, text=True, env=env, timeout=120)


def test_entrypoint_follows_label_changes(tmp_path, monkeypatch):
    # host CC detection reads /sys: on this box it is off, so the DEFAULT mode is forced to
    # 'off' (main.py:736-742) but an explicit label still wins.
    proc = run(["--debug"], {"NODE_NAME": "node-a", "CC_SCRUB_MODE": "skip", "EVICT_OPERATOR_COMPONENTS": "true"}, tmp_path)
    assert proc.returncode == 0, proc.stderr[-3000:]
    result = json.loads(next(l for l in proc.stdout.splitlines() if l.startswith("RESULT "))[7:])
    assert result["modes"] == [0] * 8                       # on -> devtools -> off
    assert result["labels"]["nvidia.com/cc.mode.state"] == "off"
    assert result["labels"]["nvidia.com/cc.ready.state"] == "false"
    assert result["labels"]["nvidia.com/gpu.deploy.vfio-manager"] == "true"   # restored after each gate
    assert (tmp_path / "ready" / ".cc-manager-ctr-ready").exists()
    assert "Label changed: 'on' -> 'devtools'" in proc.stderr and "Label changed: 'devtools' -> 'off'" in proc.stderr


def test_entrypoint_requires_node_name(tmp_path):
    proc = run([], {}, tmp_path)
    assert proc.returncode == 1
    assert "NODE_NAME environment variable must be set" in proc.stderr


def test_entrypoint_scrub_gate_fails_closed_without_cuda(tmp_path):
    """Default CC_SCRUB_MODE=require on a box with no CUDA device: the transition must NOT
    be reported as done — state label 'failed', components still restored."""
    proc = run(["--node-name", "node-a"], {"EVICT_OPERATOR_COMPONENTS": "true", "CCM_SIM_BIND_CUDA": "0"}, tmp_path)
    assert proc.returncode == 0, proc.stderr[-3000:]
    result = json.loads(next(l for l in proc.stdout.splitlines() if l.startswith("RESULT "))[7:])
    assert result["labels"]["nvidia.com/cc.mode.state"] == "failed"
    assert result["labels"]["nvidia.com/gpu.deploy.vfio-manager"] == "true"
    assert "no CUDA device" in proc.stderr or "HBM scrub" in proc.stderr


def test_entrypoint_refuses_a_simulated_register_backend_unless_told(tmp_path):
    """with nothing configured libccm picks sim/cudasim and the manager would publish
    cc.mode.state=on without touching hardware.  The production entrypoint must refuse."""
    proc = run(["--node-name", "node-a"], {"CCM_ALLOW_SIM": "0", "CC_SCRUB_MODE": "skip"}, tmp_path)
    assert proc.returncode == 1
    assert "SIMULATED register backend" in proc.stderr and "CCM_ALLOW_SIM=1" in proc.stderr
    assert "RESULT" not in proc.stdout                       # nothing was labelled, nothing was staged
    proc = run(["--node-name", "node-a"], {"CCM_ALLOW_SIM": "1", "CC_SCRUB_MODE": "skip"}, tmp_path)
    assert proc.returncode == 0, proc.stderr[-2000:]


def test_entrypoint_prefers_gpu_admin_tools_when_the_image_ships_it(tmp_path, monkeypatch):
    """Unset CC_DEVICE_LIBRARY: <app dir>/gpu-admin-tools (reference main.py:30-31) wins over libccm's registers."""
    from k8s_cc_manager_b200 import manager
    tools = tmp_path / "gpu-admin-tools" / "pci"
    tools.mkdir(parents=True)
    (tools / "__init__.py").write_text("")
    (tools / "devices.py").write_text("def find_gpus():\n    return [], 0\n")
    monkeypatch.setenv("GPU_ADMIN_TOOLS_PATH", str(tmp_path / "gpu-admin-tools"))
    monkeypatch.delenv("CC_DEVICE_LIBRARY", raising=False)
    import sys as _sys
    for m in ("pci", "pci.devices"):
        _sys.modules.pop(m, None)
    try:
        source = manager._device_source_from_env()
        assert source is not None and source() == ([], 0)    # the foreign find_gpus, wrapped with the scrub
    finally:
        for m in ("pci", "pci.devices"):
            _sys.modules.pop(m, None)
        _sys.path.remove(str(tmp_path / "gpu-admin-tools"))
    monkeypatch.setenv("CC_DEVICE_LIBRARY", "libccm")
    assert manager._device_source_from_env() is None         # CCM_ALLOW_SIM=1 in the test environment
