"""Scrub-gate policy of the manager, with duck-typed devices (no CUDA needed): the gate
must fail CLOSED — no scrub capability, a dirty read-back, a CUDA failure or a partial
coverage all leave the node labelled 'failed' and the GPU unreleased."""
from __future__ import annotations

import pytest

import scenarios as SC
from k8s_cc_manager_b200.devices import GpuError, ScrubReport


class FakeGpu:
    """Minimal duck-typed device (what a foreign device library would hand the manager)."""

    def __init__(self, bdf, report=None, scrubber=True, error=None):
        self.bdf, self.name = bdf, "fake"
        self.is_cc_query_supported = self.is_ppcie_query_supported = True
        self.cc, self.staged, self.pp = "off", "off", "off"
        self.ops = []
        self._report, self._error = report, error
        if scrubber:
            self.scrub_and_verify = self._scrub

    def is_gpu(self): return True
    def is_nvswitch(self): return False
    def query_cc_mode(self): return self.cc
    def set_cc_mode(self, m): self.staged = m
    def query_ppcie_mode(self): return self.pp
    def set_ppcie_mode(self, m): pass
    def reset_with_os(self): self.cc = self.staged; self.ops.append("reset")
    def wait_for_boot(self): self.ops.append("boot")

    def _scrub(self, nbytes=0):
        self.ops.append("scrub")
        if self._error:
            raise self._error
        return self._report or ScrubReport(self.bdf, nbytes, 190 << 30, 191 << 30, 0, 1, 25, 25, 1, 52, 1, 0)


def make_manager(devs, monkeypatch, **kw):
    import kubernetes
    from k8s_cc_manager_b200 import manager
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, {})
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "false")
    return manager.CCManager(SC.NODE, "on", True, device_source=lambda: (devs, len(devs)), **kw), c


def state(c):
    return c.labels(SC.NODE).get("nvidia.com/cc.mode.state")


def test_gate_passes_and_runs_after_boot(monkeypatch):
    devs = [FakeGpu(b) for b in SC.GPU_BDFS[:4]]
    mgr, c = make_manager(devs, monkeypatch)
    assert mgr.set_cc_mode("on") is True and state(c) == "on"
    assert all(d.ops == ["reset", "boot", "scrub"] for d in devs)
    assert len(mgr.last_transition["scrub"]) == 4
    # nothing to transition -> nothing to scrub
    for d in devs:
        d.ops.clear()
    assert mgr.set_cc_mode("on") is True and all(d.ops == [] for d in devs)


def test_only_reset_gpus_are_scrubbed(monkeypatch):
    devs = [FakeGpu(b) for b in SC.GPU_BDFS[:3]]
    devs[1].cc = devs[1].staged = "on"
    mgr, c = make_manager(devs, monkeypatch)
    assert mgr.set_cc_mode("on") is True
    assert [d.ops for d in devs] == [["reset", "boot", "scrub"], [], ["reset", "boot", "scrub"]]


@pytest.mark.parametrize("case", ["no_scrubber", "dirty", "cuda_error", "low_coverage", "bad_status"])
def test_gate_fails_closed(case, monkeypatch):
    bdf = SC.GPU_BDFS[1]
    bad = {
        "no_scrubber": FakeGpu(bdf, scrubber=False),
        "dirty": FakeGpu(bdf, report=ScrubReport(bdf, 0, 190 << 30, 191 << 30, 3, 1, 25, 25, 1, 52, 1, 0)),
        "cuda_error": FakeGpu(bdf, error=GpuError("scrub_and_verify: CUDA call failed", -6)),
        "low_coverage": FakeGpu(bdf, report=ScrubReport(bdf, 0, 100 << 30, 191 << 30, 0, 1, 13, 13, 1, 28, 1, 0)),
        "bad_status": FakeGpu(bdf, report=ScrubReport(bdf, 0, 190 << 30, 191 << 30, 0, 1, 25, 25, 1, 52, 1, -7)),
    }[case]
    devs = [FakeGpu(SC.GPU_BDFS[0]), bad, FakeGpu(SC.GPU_BDFS[2])]
    mgr, c = make_manager(devs, monkeypatch)
    assert mgr.set_cc_mode("on") is False
    assert state(c) == "failed" and c.labels(SC.NODE)["nvidia.com/cc.ready.state"] == ""


def test_explicit_byte_count_skips_coverage_check(monkeypatch):
    bdf = SC.GPU_BDFS[0]
    dev = FakeGpu(bdf, report=ScrubReport(bdf, 1 << 30, 1 << 30, 191 << 30, 0, 1, 1, 1, 1, 4, 1, 0))
    mgr, c = make_manager([dev], monkeypatch, scrub_bytes=1 << 30)
    assert mgr.set_cc_mode("on") is True and state(c) == "on"


def test_skip_mode_releases_without_scrub_and_says_so(monkeypatch, caplog):
    devs = [FakeGpu(b, scrubber=False) for b in SC.GPU_BDFS[:2]]
    mgr, c = make_manager(devs, monkeypatch, scrub_mode="skip")
    with caplog.at_level("WARNING"):
        assert mgr.set_cc_mode("on") is True
    assert state(c) == "on" and mgr.last_transition["scrub"] == "skipped"
    assert any("WITHOUT an HBM scrub" in r.message for r in caplog.records)


def test_invalid_scrub_mode_rejected(monkeypatch):
    with pytest.raises(ValueError):
        make_manager([], monkeypatch, scrub_mode="maybe")
