"""Arrangement (1) of INTEGRATION.md §D: another device library does the register work, libccm
only adds the HBM scrub.  The 'other library' here is the oracle's fake gpu-admin-tools."""
from __future__ import annotations

import sys
from pathlib import Path

import pytest

import scenarios as SC
from helpers import build_native_world

FAKE = Path(__file__).resolve().parents[1] / "oracle" / "fakes" / "gpu-admin-tools"


@pytest.fixture()
def foreign(monkeypatch):
    for m in ("_state", "pci", "pci.devices", "gpu", "nvidia_gpu_tools"):
        sys.modules.pop(m, None)
    monkeypatch.syspath_prepend(str(FAKE))
    import _state
    from pci.devices import find_gpus
    w = _state.world()
    w.reset()
    yield w, find_gpus
    for m in ("_state", "pci", "pci.devices", "gpu", "nvidia_gpu_tools"):
        sys.modules.pop(m, None)


def test_proxy_forwards_everything_but_scrub(foreign):
    from k8s_cc_manager_b200 import devices as D
    w, find = foreign
    for bdf in SC.GPU_BDFS[:2]:
        w.add_gpu(bdf)
    w.add_nvswitch(SC.SWITCH_BDFS[0])
    build_native_world(SC.scenario("p", gpus_=SC.gpus(2), modes=[]))      # same BDFs, no CUDA behind them
    devs, n = D.with_scrub(find)()
    assert n == 3 and [type(d).__name__ for d in devs] == ["ScrubbingProxy", "ScrubbingProxy", "FakeDevice"]
    g = devs[0]
    assert g.bdf == SC.GPU_BDFS[0] and g.is_gpu() and g.is_cc_query_supported is True
    g.set_cc_mode("on"); g.reset_with_os(); g.wait_for_boot()
    assert g.query_cc_mode() == "on" and w.devices[0].cc_mode == "on"       # the FOREIGN registers moved
    assert w.trace_lines()[-1].endswith("query_cc_mode on")
    with pytest.raises(D.GpuError) as e:                                      # no CUDA device: fails loudly
        g.scrub_and_verify()
    assert e.value.status == D.N.ERR_NO_CUDA


def test_manager_with_foreign_registers_and_missing_scrub_fails_closed(foreign, monkeypatch):
    import kubernetes
    from k8s_cc_manager_b200 import devices as D, manager
    w, find = foreign
    for bdf in SC.GPU_BDFS[:4]:
        w.add_gpu(bdf)
    build_native_world(SC.scenario("p", gpus_=SC.gpus(4), modes=[]))
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, {})
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "false")
    mgr = manager.CCManager(SC.NODE, "on", True, device_source=D.with_scrub(find))
    assert mgr.set_cc_mode("on") is False                       # registers flipped, scrub impossible -> not released
    assert c.labels(SC.NODE)["nvidia.com/cc.mode.state"] == "failed"
    assert all(d.cc_mode == "on" for d in w.devices)
    skip = manager.CCManager(SC.NODE, "on", True, device_source=D.with_scrub(find), scrub_mode="skip")
    assert skip.set_cc_mode("off") is True and c.labels(SC.NODE)["nvidia.com/cc.mode.state"] == "off"


@pytest.mark.gpu
def test_foreign_registers_with_real_scrub(foreign, monkeypatch):
    """On the GPU box: foreign (fake) register library + libccm scrub matched by PCI address."""
    import ctypes as C
    import kubernetes
    from k8s_cc_manager_b200 import _native as N, devices as D, manager
    w, find = foreign
    assert N.lib().ccm_init(N.BACKEND_CUDASIM) == 0
    native = [d for d in D.find_gpus()[0] if d.is_gpu()]
    for d in native:
        w.add_gpu(d.bdf.upper() if d.bdf.startswith("0000:") else d.bdf)   # case differences must not matter
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, {})
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "false")
    mgr = manager.CCManager(SC.NODE, "on", True, device_source=D.with_scrub(find), scrub_bytes=1 << 30)
    assert mgr.set_cc_mode("on") is True
    assert c.labels(SC.NODE)["nvidia.com/cc.mode.state"] == "on"
    assert all(r.clean and r.bytes_scrubbed == 1 << 30 for r in mgr.last_transition["scrub"])
    assert all(d.cc_mode == "on" for d in w.devices)


def test_entrypoint_selects_foreign_library(foreign, monkeypatch):
    from k8s_cc_manager_b200 import manager
    monkeypatch.setenv("CC_DEVICE_LIBRARY", "gpu-admin-tools")
    monkeypatch.setenv("GPU_ADMIN_TOOLS_PATH", str(FAKE))
    w, _ = foreign
    w.add_gpu(SC.GPU_BDFS[0])
    build_native_world(SC.scenario("p", gpus_=SC.gpus(1), modes=[]))
    devs, n = manager._device_source_from_env()()
    assert n == 1 and type(devs[0]).__name__ == "ScrubbingProxy"
    monkeypatch.setenv("CC_DEVICE_LIBRARY", "libccm")
    assert manager._device_source_from_env() is None
    monkeypatch.setenv("CC_DEVICE_LIBRARY", "other")
    with pytest.raises(ValueError):
        manager._device_source_from_env()


# --------------------------------------------------------------------------------------------------
# a8, the deliverable that IS possible without hardware: the product, configured the way a real CC node
# runs it (CC_DEVICE_LIBRARY=gpu-admin-tools -> every register op is DELEGATED to the foreign library,
# libccm only adds the scrub), must drive that library through exactly the call transcript the
# unmodified reference produced on the same library for the same scenario (tests/golden/transitions.json,
# recorded by oracle/gen_golden.py).  Same ops, same order, same arguments, same end registers, same
# labels and return values — for all 31 scenarios, faults and PPCIe included.
import json as _json

_GOLDEN = {s["name"]: s for s in _json.loads(
    (Path(__file__).parent / "golden" / "transitions.json").read_text())["scenarios"]}


@pytest.mark.parametrize("name", sorted(_GOLDEN))
def test_delegating_adapter_replays_the_reference_call_transcript(name, foreign, monkeypatch):
    import os
    import kubernetes  # noqa: F401
    from helpers import build_cluster
    from k8s_cc_manager_b200 import drain_gate, manager
    w, _find = foreign
    sc = {s["name"]: s for s in SC.transition_scenarios()}[name]
    exc_types = {"GpuError": sys.modules["_state"].GpuError, "RuntimeError": RuntimeError}
    for g in sc["gpus"]:
        d = w.add_gpu(g["bdf"], cc=g["cc"], ppcie=g["ppcie"], cc_supported=g["cc_supported"],
                      ppcie_supported=g["ppcie_supported"])
        d.stuck = g["stuck"]
        d.fail = {op: exc_types[t](f"injected {op} failure on {g['bdf']}") for op, t in g["fail"].items()}
    for s in sc["switches"]:
        d = w.add_nvswitch(s["bdf"], ppcie=s["ppcie"], ppcie_supported=s["ppcie_supported"])
        d.stuck = s["stuck"]
        d.fail = {op: exc_types[t](f"injected {op} failure on {s['bdf']}") for op, t in s["fail"].items()}
    build_native_world(SC.scenario("p", gpus_=SC.gpus(max(1, len(sc["gpus"]))), modes=[]))   # libccm side: no CUDA
    c = build_cluster(sc)
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "true" if sc["evict"] else "false")
    monkeypatch.setenv("OPERATOR_NAMESPACE", SC.NAMESPACE)
    monkeypatch.setenv("CC_DEVICE_LIBRARY", "gpu-admin-tools")
    monkeypatch.setenv("GPU_ADMIN_TOOLS_PATH", str(FAKE))
    monkeypatch.setattr(drain_gate, "_now", c.clock.time)
    monkeypatch.setattr(drain_gate, "_pause", c.clock.sleep)
    source = manager._device_source_from_env()            # the production wiring, not a hand-made lambda
    assert source is not None
    if str(FAKE) in sys.path[1:]:
        sys.path.remove(str(FAKE))                        # _device_source_from_env prepends it once more
    mgr = manager.CCManager(SC.NODE, "on", sc["host_cc"], device_source=source, max_parallel=1, scrub_mode="skip")
    want = _GOLDEN[name]
    for i, mode in enumerate(sc["modes"]):
        n_dev = len(w.trace)
        got = {}
        try:
            got["result"] = mgr.set_cc_mode(mode)
        except SystemExit as exc:
            got["exit"] = exc.code
        w_step = want["steps"][i]
        assert w.trace_lines()[n_dev:] == w_step["device_trace"], f"{name} step {i}: foreign call transcript"
        assert got.get("result") == w_step.get("result") and got.get("exit") == w_step.get("exit")
        assert {d.bdf: {"cc": d.cc_mode, "ppcie": d.ppcie_mode} for d in w.devices} == w_step["registers"]
        assert c.labels(SC.NODE) == w_step["labels"]
        if "exit" in got:
            break
    assert os.environ["CC_DEVICE_LIBRARY"] == "gpu-admin-tools"
