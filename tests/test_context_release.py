"""The manager must not sit on a CUDA context after the scrub gate (it pins HBM, blocks a vfio
re-bind, dies with the next device reset).  CPU: policy with duck-typed devices.  GPU: a
subprocess scrubs, releases, and is no longer a compute client of the GPU."""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

import scenarios as SC
from test_scrub_gate_policy import FakeGpu, make_manager, state

ROOT = Path(__file__).resolve().parents[1]


class ReleasableGpu(FakeGpu):
    def __init__(self, bdf, fail_release=False, **kw):
        super().__init__(bdf, **kw)
        self.fail_release = fail_release

    def release_cuda_context(self):
        self.ops.append("release")
        if self.fail_release:
            raise RuntimeError("context busy")


def test_contexts_are_released_after_the_gate(monkeypatch):
    monkeypatch.setenv("CC_RELEASE_CUDA_CONTEXT", "true")
    devs = [ReleasableGpu(b) for b in SC.GPU_BDFS[:3]]
    mgr, c = make_manager(devs, monkeypatch)
    assert mgr.set_cc_mode("on") is True and state(c) == "on"
    assert all(d.ops == ["reset", "boot", "scrub", "release"] for d in devs)


def test_release_also_happens_when_the_gate_fails_and_never_masks_the_verdict(monkeypatch):
    from k8s_cc_manager_b200.devices import GpuError
    monkeypatch.setenv("CC_RELEASE_CUDA_CONTEXT", "true")
    devs = [ReleasableGpu(SC.GPU_BDFS[0], fail_release=True),
            ReleasableGpu(SC.GPU_BDFS[1], error=GpuError("CUDA call failed", -6))]
    mgr, c = make_manager(devs, monkeypatch)
    assert mgr.set_cc_mode("on") is False and state(c) == "failed"
    assert devs[0].ops[-1] == "release" and devs[1].ops[-1] == "release"
    # a failing release alone does not fail a clean gate
    devs = [ReleasableGpu(SC.GPU_BDFS[0], fail_release=True)]
    mgr, c = make_manager(devs, monkeypatch)
    assert mgr.set_cc_mode("on") is True and state(c) == "on"


def test_release_can_be_switched_off(monkeypatch):
    monkeypatch.setenv("CC_RELEASE_CUDA_CONTEXT", "false")
    devs = [ReleasableGpu(SC.GPU_BDFS[0])]
    mgr, c = make_manager(devs, monkeypatch)
    assert mgr.set_cc_mode("on") is True and devs[0].ops == ["reset", "boot", "scrub"]


def test_native_release_is_a_noop_without_cuda(native):
    lib = native.lib()
    assert lib.ccm_sim_topology(2, 1) == 0
    assert lib.ccm_sim_set(-1, b"cuda_ordinal", -1) == 0
    assert lib.ccm_device_release(0) == 0 and lib.ccm_device_release(2) == 0
    assert lib.ccm_device_release(99) == native.ERR_NO_DEVICE


WORKER = r"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.environ["REPO"])
import pynvml
from k8s_cc_manager_b200 import _native as N, devices as D
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
def mine():
    return any(p.pid == os.getpid() for p in pynvml.nvmlDeviceGetComputeRunningProcesses(h))
assert N.lib().ccm_init(N.BACKEND_CUDASIM) == 0
gpu = [d for d in D.find_gpus()[0] if d.is_gpu()][0]
out = {"before": mine()}
t0 = time.perf_counter(); r1 = gpu.scrub_and_verify(4 << 30); out["first_call_s"] = time.perf_counter() - t0
out["holding"] = mine()
gpu.release_cuda_context()
out["after_release"] = mine()
gpu.wait_scrub_released()                                   # asking must not bring the context back
out["after_release_and_wait"] = mine()
t0 = time.perf_counter(); r2 = gpu.scrub_and_verify(4 << 30); out["call_after_release_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); r3 = gpu.scrub_and_verify(4 << 30); out["warm_call_s"] = time.perf_counter() - t0
out["clean"] = [r1.clean, r2.clean, r3.clean]
gpu.release_cuda_context(); gpu.release_cuda_context()      # idempotent
out["after_second_release"] = mine()
print("RESULT " + json.dumps(out))
"""


@pytest.mark.gpu
def test_release_drops_the_cuda_context_and_scrub_comes_back(tmp_path):
    proc = subprocess.run([sys.executable, "-c", WORKER], capture_output=True, text=True,
                          env=dict(os.environ, REPO=str(ROOT)), timeout=300)
    assert proc.returncode == 0, proc.stderr[-3000:]
    out = json.loads(next(l for l in proc.stdout.splitlines() if l.startswith("RESULT "))[7:])
    assert out["before"] is False and out["holding"] is True
    assert out["after_release"] is False and out["after_second_release"] is False
    assert out["after_release_and_wait"] is False
    assert out["clean"] == [True, True, True]
    print(f"\ncontext re-creation: first {out['first_call_s']*1e3:.0f} ms, after release "
          f"{out['call_after_release_s']*1e3:.0f} ms, warm {out['warm_call_s']*1e3:.0f} ms")
