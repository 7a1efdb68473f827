"""Control-plane surface: watch loop (reference main.py:600-684), CLI / env handling
(main.py:698-763), readiness file, host CC detection — against tests/golden/watch.json
recorded from the unmodified reference (with `time` injected, see oracle/gen_golden.py)."""
from __future__ import annotations

import json
import os
from pathlib import Path
from types import SimpleNamespace

import pytest

import scenarios as SC
from helpers import build_native_world

GOLDEN = json.loads((Path(__file__).parent / "golden" / "watch.json").read_text())
WATCH = {s["name"]: s for s in GOLDEN["scenarios"]}


def install_watch_script(cluster, script):
    from kubernetes.client.rest import ApiException

    def make_event(e):
        if e == "ERROR":
            return {"type": "ERROR", "object": {"code": 500}, "raw_object": {}}
        labels = {} if e["label"] is None else {"nvidia.com/cc.mode": e["label"]}
        node = SimpleNamespace(metadata=SimpleNamespace(name=SC.NODE, labels=labels, resource_version=e["rv"]))
        return {"type": e["type"], "object": node}

    for batch in script:
        if isinstance(batch, dict):
            status, relabel = batch["raise"], batch.get("relabel")

            def raiser(c, status=status, relabel=relabel):
                if relabel is not None:
                    c.nodes[SC.NODE].metadata.labels["nvidia.com/cc.mode"] = relabel
                raise ApiException(status=status, reason="injected")
            cluster.watch_script.append([raiser])
        else:
            cluster.watch_script.append([make_event(e) for e in batch])


def run_watch(wsc, monkeypatch):
    import kubernetes
    from kubernetes.watch import WatchScriptExhausted
    from k8s_cc_manager_b200 import manager

    build_native_world(SC.scenario("w", gpus_=SC.gpus(8, cc="on"), modes=[]))
    c = kubernetes.reset_cluster()
    labels = {} if wsc["initial_label"] is None else {"nvidia.com/cc.mode": wsc["initial_label"]}
    c.add_node(SC.NODE, labels)
    install_watch_script(c, wsc["script"])
    monkeypatch.setenv("EVICT_OPERATOR_COMPONENTS", "false")
    applied = []
    monkeypatch.setattr(manager, "create_readiness_file", lambda: applied.append("<readiness>"))
    mgr = manager.CCManager(node_name=SC.NODE, default_mode=wsc["default"], host_cc=True, scrub_mode="skip")
    mgr._sleep = c.clock.sleep
    real = mgr.set_cc_mode
    mgr.set_cc_mode = lambda mode: (applied.append(mode), real(mode))[1]
    end = None
    try:
        mgr.watch_and_apply()
    except WatchScriptExhausted:
        end = "script-exhausted"
    except BaseException as exc:  # noqa: BLE001
        end = type(exc).__name__
    return {"applied": applied, "end": end,
            "watch_resource_versions": [k.get("resource_version") for k in c.watch_calls],
            "watch_kwargs_keys": sorted(c.watch_calls[0]) if c.watch_calls else [],
            "sleeps": list(c.clock.sleeps), "labels": c.labels(SC.NODE),
            "read_node_calls": c.verbs().count("read_node")}


@pytest.mark.parametrize("wsc", SC.watch_scenarios(), ids=lambda w: w["name"])
def test_watch_loop_matches_reference(wsc, monkeypatch):
    got = run_watch(wsc, monkeypatch)
    want = WATCH[wsc["name"]]
    for key in ("applied", "end", "watch_resource_versions", "watch_kwargs_keys", "sleeps", "labels",
                "read_node_calls"):
        assert got[key] == want[key], key


def test_reference_as_shipped_crashes_where_we_sleep(monkeypatch):
    """The reference raises NameError on its first reconnect (`time` never imported,
    main.py:684); the golden records that, and the product sleeps 5 s and recovers."""
    assert WATCH["watch_500_then_recover__as_shipped"]["end"] == "NameError"
    got = run_watch(SC.watch_scenarios()[4], monkeypatch)
    assert got["end"] == "script-exhausted" and got["sleeps"] == [5, 5]


# ----------------------------------------------------------------------- CLI
def test_main_requires_node_name(monkeypatch):
    from k8s_cc_manager_b200 import manager
    monkeypatch.delenv("NODE_NAME", raising=False)
    with pytest.raises(SystemExit) as e:
        manager.main([])
    assert e.value.code == 1


def test_arg_parser_defaults_follow_env(monkeypatch):
    from k8s_cc_manager_b200 import manager
    monkeypatch.setenv("NODE_NAME", "n1")
    monkeypatch.setenv("DEFAULT_CC_MODE", "devtools")
    monkeypatch.setenv("KUBECONFIG", "/tmp/kc")
    args = manager.build_arg_parser().parse_args([])
    assert (args.node_name, args.default_cc_mode, args.kubeconfig, args.debug) == ("n1", "devtools", "/tmp/kc", False)
    args = manager.build_arg_parser().parse_args(["-m", "off", "--node-name", "n2", "--debug"])
    assert (args.node_name, args.default_cc_mode, args.debug) == ("n2", "off", True)


def test_main_overrides_default_when_host_has_no_cc(monkeypatch):
    """reference main.py:736-742: host not CC-capable => default mode forced to 'off'."""
    from k8s_cc_manager_b200 import manager
    seen = {}

    class Spy:
        def __init__(self, **kw):
            seen.update(kw)

        def run(self):
            pass

    monkeypatch.setattr(manager, "CCManager", Spy)
    monkeypatch.setattr(manager, "is_host_cc_enabled", lambda: False)
    manager.main(["--node-name", "n", "-m", "on"])
    assert seen["default_mode"] == "off" and seen["host_cc"] is False
    monkeypatch.setattr(manager, "is_host_cc_enabled", lambda: True)
    manager.main(["--node-name", "n", "-m", "on", "--kubeconfig", "/k"])
    assert seen["default_mode"] == "on" and seen["host_cc"] is True and seen["kubeconfig"] == "/k"


def test_main_exit_codes(monkeypatch):
    from k8s_cc_manager_b200 import manager

    class Boom:
        def __init__(self, **kw):
            pass

        def run(self):
            raise RuntimeError("x")

    monkeypatch.setattr(manager, "CCManager", Boom)
    with pytest.raises(SystemExit) as e:
        manager.main(["--node-name", "n"])
    assert e.value.code == 1

    class Ctrl(Boom):
        def run(self):
            raise KeyboardInterrupt

    monkeypatch.setattr(manager, "CCManager", Ctrl)
    with pytest.raises(SystemExit) as e:
        manager.main(["--node-name", "n"])
    assert e.value.code == 0


def test_is_host_cc_enabled(tmp_path):
    from k8s_cc_manager_b200.manager import is_host_cc_enabled
    assert is_host_cc_enabled(str(tmp_path)) is False
    p = tmp_path / "module/kvm_intel/parameters"
    p.mkdir(parents=True)
    (p / "tdx").write_text("N\n")
    assert is_host_cc_enabled(str(tmp_path)) is False
    (p / "tdx").write_text("Y\n")
    assert is_host_cc_enabled(str(tmp_path)) is True
    (p / "tdx").write_text("0\n")
    q = tmp_path / "module/kvm_amd/parameters"
    q.mkdir(parents=True)
    (q / "sev_snp").write_text("1\n")
    assert is_host_cc_enabled(str(tmp_path)) is True


def test_readiness_file(tmp_path, monkeypatch):
    from k8s_cc_manager_b200 import manager
    target = tmp_path / "a" / "b" / ".ready"
    monkeypatch.setattr(manager, "READINESS_FILE", str(target))
    manager.create_readiness_file()
    assert target.exists()
    monkeypatch.setattr(manager, "READINESS_FILE", "/proc/nope/.ready")
    manager.create_readiness_file()  # must not raise


def test_kube_config_loading_order(monkeypatch):
    """in-cluster first, then kubeconfig; an explicit --kubeconfig path is honoured
    (the reference parses the flag and then ignores it: main.py:703-707 vs 129-138)."""
    import kubernetes
    from k8s_cc_manager_b200 import manager
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, {})
    manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")
    assert c.loaded == ["incluster"]
    c.incluster_ok = False
    manager.CCManager(SC.NODE, "on", True, kubeconfig="/tmp/kubeconfig", scrub_mode="skip")
    assert c.loaded[-1] == ("kubeconfig", "/tmp/kubeconfig")
    manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")
    assert c.loaded[-1] == ("kubeconfig", None)
    c.kubeconfig_ok = False
    with pytest.raises(kubernetes.config.ConfigException):
        manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")


def test_invalid_mode_is_refused_before_touching_hardware(monkeypatch):
    import kubernetes
    from helpers import sim_trace
    from k8s_cc_manager_b200 import manager
    build_native_world(SC.scenario("w", modes=[]))
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, SC.all_true_labels())
    mgr = manager.CCManager(SC.NODE, "on", True, scrub_mode="skip")
    assert mgr.set_cc_mode("bogus") is False
    assert sim_trace() == []
    assert c.labels(SC.NODE)["nvidia.com/cc.mode.state"] == "failed"
    assert c.labels(SC.NODE)["nvidia.com/gpu.deploy.vfio-manager"] == "true"  # nothing was paused
