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


# ------------------------------------------------------------------ round 2 policy
def test_default_coverage_bar_is_just_under_a_clean_device(monkeypatch):
    """0.95 of the device used to pass (default was 0.90): 9.5 GB unscrubbed is not 'full-HBM'."""
    bdf = SC.GPU_BDFS[0]
    total = 191 << 30
    for cov, ok in ((0.9966, True), (0.991, True), (0.95, False), (0.9899, False)):
        rep = ScrubReport(bdf, 0, int(cov * total), total, 0, 1, 25, 25, 0, 52, 23, 0,
                          release_deferred=1, device_free_before=int(0.9966 * total), bytes_unreached=0)
        mgr, c = make_manager([FakeGpu(bdf, report=rep)], monkeypatch)
        assert mgr.scrub_min_coverage == 0.99
        assert mgr.set_cc_mode("on") is ok, cov
        assert state(c) == ("on" if ok else "failed")


def test_partial_scrub_from_the_environment_needs_an_explicit_opt_in(monkeypatch):
    monkeypatch.setenv("CC_SCRUB_BYTES", str(1 << 30))
    with pytest.raises(ValueError, match="CC_SCRUB_ALLOW_PARTIAL"):
        make_manager([], monkeypatch)
    monkeypatch.setenv("CC_SCRUB_ALLOW_PARTIAL", "true")
    mgr, _ = make_manager([], monkeypatch)
    assert mgr.scrub_bytes == 1 << 30
    monkeypatch.delenv("CC_SCRUB_ALLOW_PARTIAL")
    mgr, _ = make_manager([], monkeypatch, scrub_mode="skip")           # no gate: the byte count is moot
    assert mgr.scrub_mode == "skip"


def test_a_failed_gate_is_not_laundered_by_a_restart(monkeypatch):
    """gate fails -> GPUs are already in the target mode, label 'failed'.  A restarted
    manager sees mode_is_set() == True; it must re-run the gate before it publishes `mode`."""
    bdf = SC.GPU_BDFS[0]
    dirty = ScrubReport(bdf, 0, 190 << 30, 191 << 30, 3, 1, 25, 25, 1, 52, 1, 0)
    dev = FakeGpu(bdf, report=dirty)
    mgr, c = make_manager([dev], monkeypatch)
    assert mgr.set_cc_mode("on") is False and state(c) == "failed" and dev.cc == "on"
    # "restart": a fresh manager, same registers, same node labels — and the memory is still dirty
    from k8s_cc_manager_b200 import manager
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "false")
    mgr2 = manager.CCManager(SC.NODE, "on", True, device_source=lambda: ([dev], 1))
    dev.ops.clear()
    assert mgr2.set_cc_mode("on") is False
    assert dev.ops == ["scrub"] and state(c) == "failed"                 # re-gated, still refused, no reset
    # once the HBM reads back clean the same call succeeds — still without touching the registers
    dev._report = None
    dev.ops.clear()
    assert mgr2.set_cc_mode("on") is True and state(c) == "on" and dev.ops == ["scrub"]
    # and a healthy already-set node costs no scrub at all — and exactly the reference's two API verbs
    dev.ops.clear()
    c.calls.clear()
    assert mgr2.set_cc_mode("on") is True and dev.ops == []
    assert c.verbs() == ["read_node", "patch_node"]


def test_skip_mode_keeps_the_reference_early_out_verbatim(monkeypatch):
    import kubernetes  # noqa: F401
    dev = FakeGpu(SC.GPU_BDFS[0])
    dev.cc = dev.staged = "on"
    mgr, c = make_manager([dev], monkeypatch, scrub_mode="skip")
    c.nodes[SC.NODE].metadata.labels["nvidia.com/cc.mode.state"] = "failed"
    c.calls.clear()
    assert mgr.set_cc_mode("on") is True and state(c) == "on"
    assert c.verbs() == ["read_node", "patch_node"]                        # exactly the reference's two verbs


def test_auto_mode_releases_gpus_that_no_cuda_context_can_reach_and_says_so(monkeypatch, caplog):
    """on a real CC node the GPUs are vfio-bound (or CC-on under a driver that cannot run
    CUDA): 'require' fails every transition there.  'auto' scrubs what can be scrubbed and marks the rest."""
    from k8s_cc_manager_b200 import devices as D
    reachable = FakeGpu(SC.GPU_BDFS[0])
    blind = D.ScrubbingProxy(FakeGpu(SC.GPU_BDFS[1], scrubber=False), None)    # foreign GPU, no CUDA device behind it
    mgr, c = make_manager([reachable, blind], monkeypatch, scrub_mode="auto")
    with caplog.at_level("WARNING"):
        assert mgr.set_cc_mode("on") is True
    assert state(c) == "on" and reachable.ops == ["reset", "boot", "scrub"]
    assert mgr.last_transition["scrub_skipped"] == [SC.GPU_BDFS[1]]
    assert c.nodes[SC.NODE].metadata.annotations["nvidia.com/cc-manager.scrub-skipped"] == SC.GPU_BDFS[1]
    assert any("WITHOUT an HBM scrub" in r.message for r in caplog.records)
    # the same node under 'require' fails closed
    for d in (reachable, object.__getattribute__(blind, "_foreign")):
        d.cc = d.staged = "off"
    mgr, c = make_manager([reachable, blind], monkeypatch)
    assert mgr.set_cc_mode("on") is False and state(c) == "failed"
    # auto + everything reachable: the annotation is cleared
    reachable.cc = reachable.staged = "off"
    mgr, c = make_manager([reachable], monkeypatch, scrub_mode="auto")
    c.nodes[SC.NODE].metadata.annotations = {"nvidia.com/cc-manager.scrub-skipped": "stale"}
    assert mgr.set_cc_mode("on") is True
    assert "nvidia.com/cc-manager.scrub-skipped" not in (c.nodes[SC.NODE].metadata.annotations or {})


def test_verdict_is_published_before_the_gate_gives_its_gpus_back(monkeypatch):
    """The label patch must not wait for unmap / release / context teardown."""
    import kubernetes  # noqa: F401
    order = []

    class Gpu(FakeGpu):
        def release_cuda_context(self):
            order.append("release")

    monkeypatch.setenv("CC_RELEASE_CUDA_CONTEXT", "true")
    devs = [Gpu(b) for b in SC.GPU_BDFS[:2]]
    mgr, c = make_manager(devs, monkeypatch)
    c.on_patch = lambda cluster, node, labels: order.append("label:" + labels.get("nvidia.com/cc.mode.state", ""))
    assert mgr.set_cc_mode("on") is True
    assert order == ["label:on", "release", "release"]
    assert 0 < mgr.last_transition["seconds_to_verdict"] <= mgr.last_transition["seconds"]
    assert "release" in mgr.last_transition["phase_seconds"]
