"""Drain gate vs the reference's label algebra (tests/golden/labels.json recorded from
reference gpu_operator_eviction.py) + the two opt-in hardening features."""
from __future__ import annotations

import json
from pathlib import Path

import pytest

import scenarios as SC

LABELS = json.loads((Path(__file__).parent / "golden" / "labels.json").read_text())


def test_constants_match_reference():
    from k8s_cc_manager_b200 import drain_gate as G, manager
    assert G.COMPONENT_LABELS == LABELS["component_labels"]
    assert G.COMPONENT_APP_LABELS == LABELS["component_app_labels"]
    assert G.PAUSED_STR == LABELS["paused_str"]
    assert manager.CC_MODE_CONFIG_LABEL == LABELS["cc_mode_config_label"]


@pytest.mark.parametrize("row", LABELS["pause_table"], ids=lambda r: repr(r["input"]))
def test_pause_unpause_table(row):
    from k8s_cc_manager_b200 import drain_gate as G
    assert G._maybe_set_paused(row["input"]) == row["paused"]
    assert G._maybe_set_unpaused(row["input"]) == row["unpaused"]


@pytest.mark.parametrize("state", sorted(LABELS["state_labels"]))
def test_state_label_mapping(state, cluster):
    import kubernetes
    from k8s_cc_manager_b200 import drain_gate as G
    cluster.add_node(SC.NODE, {"keep": "me"})
    ok = G.set_cc_state_label(kubernetes.client.CoreV1Api(), SC.NODE, state)
    assert ok == LABELS["state_labels"][state]["ok"]
    assert cluster.labels(SC.NODE) == LABELS["state_labels"][state]["labels"]


def test_state_label_api_failure_returns_false(cluster):
    import kubernetes
    from k8s_cc_manager_b200 import drain_gate as G
    cluster.add_node(SC.NODE, {})
    cluster.fail_next("patch_node", 500)
    assert G.set_cc_state_label(kubernetes.client.CoreV1Api(), SC.NODE, "on") is False


def test_fetch_raises_on_api_error(cluster):
    import kubernetes
    from kubernetes.client.rest import ApiException
    from k8s_cc_manager_b200 import drain_gate as G
    cluster.add_node(SC.NODE, {})
    cluster.fail_next("read_node", 503)
    with pytest.raises(ApiException):
        G.fetch_current_component_labels(kubernetes.client.CoreV1Api(), SC.NODE)


def test_pod_poll_survives_api_errors(cluster):
    import kubernetes
    from k8s_cc_manager_b200 import drain_gate as G
    v1 = kubernetes.client.CoreV1Api()
    cluster.add_node(SC.NODE, {"nvidia.com/gpu.deploy.vfio-manager": "true"})
    cluster.add_pod("nvidia-vfio-manager", SC.NODE, gone_after=3.0)
    cluster.fail_next("list_namespaced_pod", 500, times=2)
    current = G.fetch_current_component_labels(v1, SC.NODE)
    assert G.evict_gpu_operator_components(v1, SC.NODE, SC.NAMESPACE, current, timeout=300,
                                           clock=cluster.clock.time, sleep=cluster.clock.sleep)
    assert cluster.clock.sleeps == [2.0, 2.0]          # two failed polls, third sees no pods (t=4 s > 3 s)
    assert cluster.labels(SC.NODE)["nvidia.com/gpu.deploy.vfio-manager"] == G.PAUSED_STR


def test_concurrent_wait_overlaps_component_timeouts(cluster):
    """Reference waits component after component (up to 5 x 300 s); opt-in parallel wait
    bounds the gate by ONE timeout."""
    import threading
    import kubernetes
    from k8s_cc_manager_b200 import drain_gate as G
    v1 = kubernetes.client.CoreV1Api()
    cluster.add_node(SC.NODE, SC.all_true_labels())
    for app in SC.COMPONENTS.values():
        cluster.add_pod(app, SC.NODE, gone_after=None)   # never terminate
    current = G.fetch_current_component_labels(v1, SC.NODE)
    # per-thread virtual clocks: each waiter burns its own 10 s budget
    local = threading.local()

    def clock():
        return getattr(local, "t", 0.0)

    def sleep(s):
        local.t = clock() + s

    assert G.evict_gpu_operator_components(v1, SC.NODE, SC.NAMESPACE, current, timeout=10, concurrent_wait=True,
                                           clock=clock, sleep=sleep)
    polls = [c for c in cluster.calls if c[0] == "list_namespaced_pod"]
    assert len(polls) == 5 * 5                            # 5 components x (10 s / 2 s) polls each
    assert {c[1][2] for c in polls} == {f"app={a}" for a in SC.COMPONENTS.values()}


def test_journal_annotation_roundtrip(cluster):
    import kubernetes
    from k8s_cc_manager_b200 import drain_gate as G
    v1 = kubernetes.client.CoreV1Api()
    original = dict(SC.all_true_labels(), **{"nvidia.com/gpu.deploy.vgpu-manager": "custom"})
    cluster.add_node(SC.NODE, original)
    current = G.fetch_current_component_labels(v1, SC.NODE)
    assert G.evict_gpu_operator_components(v1, SC.NODE, SC.NAMESPACE, current, journal_annotation=True,
                                           clock=cluster.clock.time, sleep=cluster.clock.sleep)
    # "crash" here: a new manager instance can still find the original values
    assert G.recover_journaled_labels(v1, SC.NODE) == current
    assert cluster.labels(SC.NODE)["nvidia.com/gpu.deploy.vgpu-manager"] == "custom_" + G.PAUSED_STR
    assert G.reschedule_gpu_operator_components(v1, SC.NODE, current, journal_annotation=True)
    assert G.recover_journaled_labels(v1, SC.NODE) is None
    assert {k: v for k, v in cluster.labels(SC.NODE).items() if k in original} == original


def test_root_shims_expose_reference_names():
    """Drop-in file names of the reference image (/app/main.py, /app/gpu_operator_eviction.py)."""
    import importlib
    shim = importlib.import_module("gpu_operator_eviction")
    for name in ("fetch_current_component_labels", "evict_gpu_operator_components",
                 "reschedule_gpu_operator_components", "set_cc_state_label", "COMPONENT_LABELS",
                 "COMPONENT_APP_LABELS", "PAUSED_STR", "_maybe_set_paused", "_maybe_set_unpaused"):
        assert hasattr(shim, name), name
    entry = importlib.import_module("main")
    for name in ("CCManager", "main", "create_readiness_file", "is_host_cc_enabled", "CC_MODE_CONFIG_LABEL",
                 "READINESS_FILE"):
        assert hasattr(entry, name), name


def test_manager_recovers_from_a_crash_between_evict_and_reschedule(cluster, monkeypatch):
    """SURVEY §8f N1 end to end on the fake API: the manager dies after the components were paused and
    before they were restored (the reference loses the original values here: main.py:556-576 keeps them
    in a local variable).  The next start finds the journal, restores the labels, and only then reconciles."""
    import kubernetes
    from kubernetes.watch import WatchScriptExhausted
    from helpers import build_native_world
    from k8s_cc_manager_b200 import drain_gate as G, manager
    original = dict(SC.all_true_labels(), **{"nvidia.com/gpu.deploy.vgpu-manager": "custom", "nvidia.com/cc.mode": "on"})
    build_native_world(SC.scenario("crash", gpus_=SC.gpus(4), modes=[]))
    cluster.add_node(SC.NODE, original)
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "true")
    monkeypatch.setenv("CC_JOURNAL_COMPONENT_LABELS", "true")
    monkeypatch.setattr(G, "_now", cluster.clock.time)
    monkeypatch.setattr(G, "_pause", cluster.clock.sleep)

    class Died(BaseException):
        pass

    mgr = manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")

    def die(*a, **k):
        raise Died()                                       # SIGKILL stand-in: no except/finally of ours runs after it
    monkeypatch.setattr(mgr, "_set_cc_mode_direct", die)
    with pytest.raises(Died):
        mgr.set_cc_mode("on")
    paused = cluster.labels(SC.NODE)
    assert paused["nvidia.com/gpu.deploy.vfio-manager"] == G.PAUSED_STR
    assert paused["nvidia.com/gpu.deploy.vgpu-manager"] == "custom_" + G.PAUSED_STR
    assert G.JOURNAL_ANNOTATION in cluster.nodes[SC.NODE].metadata.annotations

    # ---- restart ------------------------------------------------------------------------------
    mgr2 = manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")
    seen_at_reconcile = {}
    real = mgr2.set_cc_mode

    def spy(mode):
        seen_at_reconcile.update(cluster.labels(SC.NODE))
        return real(mode)
    mgr2.set_cc_mode = spy
    monkeypatch.setattr(manager, "create_readiness_file", lambda: None)
    with pytest.raises(WatchScriptExhausted):
        mgr2.watch_and_apply()
    # the labels were back to their ORIGINAL values before the reconcile touched anything
    assert {k: seen_at_reconcile[k] for k in G.COMPONENT_LABELS} == {k: original[k] for k in G.COMPONENT_LABELS}
    final = cluster.labels(SC.NODE)
    assert {k: final[k] for k in G.COMPONENT_LABELS} == {k: original[k] for k in G.COMPONENT_LABELS}
    assert final["nvidia.com/cc.mode.state"] == "on"
    assert G.JOURNAL_ANNOTATION not in (cluster.nodes[SC.NODE].metadata.annotations or {})
    assert G.recover_journaled_labels(kubernetes.client.CoreV1Api(), SC.NODE) is None

    # without the journal (reference behaviour) the same crash leaves the components paused for good
    cluster2 = kubernetes.reset_cluster()
    cluster2.add_node(SC.NODE, original)
    build_native_world(SC.scenario("crash", gpus_=SC.gpus(4), modes=[]))
    monkeypatch.setenv("CC_JOURNAL_COMPONENT_LABELS", "false")
    mgr3 = manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")
    monkeypatch.setattr(mgr3, "_set_cc_mode_direct", die)
    with pytest.raises(Died):
        mgr3.set_cc_mode("on")
    mgr4 = manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")
    assert mgr4.recover_interrupted_transition() is False
    assert cluster2.labels(SC.NODE)["nvidia.com/gpu.deploy.vfio-manager"] == G.PAUSED_STR
