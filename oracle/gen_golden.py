#!/usr/bin/env python3
"""Generates tests/golden/*.json by running the UNMODIFIED reference.

ORACLE / TEST INFRASTRUCTURE.  Runs only in the development container, where
/root/reference exists; the GPU box never needs it — the fixtures it writes are
committed under tests/golden/.

How the reference is driven without touching its source:
  * `kubernetes`            -> tests/fakes/kubernetes           (in-memory API server)
  * `nvidia_gpu_tools`, `pci.devices`, `gpu`
                            -> oracle/fakes/gpu-admin-tools     (in-memory register file)
  both simply placed on sys.path before /root/reference/main.py is imported by path.
  * gpu_operator_eviction.time is pointed at the fake cluster's virtual clock so the
    2 s pod polls (reference gpu_operator_eviction.py:186-204) cost no wall time.
  * the reference forgets `import time` (main.py:24-28 vs main.py:684).  For the watch
    scenarios a `time` object is INJECTED into the imported module's globals so the
    intended behaviour (sleep 5 s, retry, fatal after 10) can be observed; the scenario
    `watch_500_then_recover__as_shipped` records the crash the shipped code really has.

Usage:  python oracle/gen_golden.py [--check]     (--check: diff against committed files)
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import logging
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REFERENCE = Path(os.environ.get("CCM_REFERENCE_DIR", "/root/reference"))
GOLDEN_DIR = ROOT / "tests" / "golden"

sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "oracle" / "fakes" / "gpu-admin-tools"))
sys.path.insert(0, str(ROOT / "tests" / "fakes"))

import scenarios as SC  # noqa: E402


def load_reference():
    if not (REFERENCE / "main.py").exists():
        raise SystemExit(f"{REFERENCE}/main.py not found: golden fixtures can only be regenerated "
                         "where the reference checkout exists")
    sys.path.append(str(REFERENCE))  # for `from gpu_operator_eviction import ...`
    spec = importlib.util.spec_from_file_location("reference_main", REFERENCE / "main.py")
    mod = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    spec.loader.exec_module(mod)  # runs the reference's sys.path.insert + imports
    sys.path[:] = saved
    logging.getLogger().setLevel(logging.CRITICAL)
    logging.disable(logging.CRITICAL)
    return mod


# ------------------------------------------------------------------ world setup
def build_world(sc):
    """Fake devices + fake cluster for one scenario.  Returns (world, cluster)."""
    import kubernetes
    from _state import GpuError, world

    w = world()
    w.reset()
    exc_types = {"GpuError": GpuError, "RuntimeError": RuntimeError}
    for g in sc["gpus"]:
        d = w.add_gpu(g["bdf"], cc=g["cc"], ppcie=g["ppcie"], cc_supported=g["cc_supported"],
                      ppcie_supported=g["ppcie_supported"])
        d.stuck = g["stuck"]
        d.fail = {op: exc_types[t](f"injected {op} failure on {g['bdf']}") for op, t in g["fail"].items()}
    for s in sc["switches"]:
        d = w.add_nvswitch(s["bdf"], ppcie=s["ppcie"], ppcie_supported=s["ppcie_supported"])
        d.stuck = s["stuck"]
        d.fail = {op: exc_types[t](f"injected {op} failure on {s['bdf']}") for op, t in s["fail"].items()}

    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, dict(sc.get("labels", {})))
    for p in sc.get("pods", []):
        c.add_pod(p["app"], SC.NODE, namespace=SC.NAMESPACE, gone_after=p["gone_after"])
    for verb, statuses in sc.get("k8s_fail", {}).items():
        from kubernetes.client.rest import ApiException
        c.fail[verb] = [None if s is None else ApiException(status=s, reason="injected") for s in statuses]
    return w, c


def k8s_trace(cluster):
    out = []
    for verb, args in cluster.calls:
        if verb == "patch_node":
            labels = args[1]
            out.append(["patch_node", labels if isinstance(labels, dict) else None])
        elif verb == "list_namespaced_pod":
            out.append([verb, args[2]])
        else:
            out.append([verb])
    return out


def run_transition_on_reference(ref, sc):
    import gpu_operator_eviction as ref_evict

    w, c = build_world(sc)
    ref_evict.time = c.clock  # virtual time.time()/time.sleep()
    os.environ["EVICT_OPERATOR_COMPONENTS"] = "true" if sc["evict"] else "false"
    os.environ["OPERATOR_NAMESPACE"] = SC.NAMESPACE
    mgr = ref.CCManager(node_name=SC.NODE, default_mode="on", host_cc=sc["host_cc"])
    steps = []
    for mode in sc["modes"]:
        n_dev, n_k8s = len(w.trace), len(c.calls)
        step = {"mode": mode}
        try:
            step["result"] = mgr.set_cc_mode(mode)
        except SystemExit as exc:
            step["exit"] = exc.code
        step["device_trace"] = w.trace_lines()[n_dev:]
        step["k8s"] = k8s_trace(c)[n_k8s:]
        step["labels"] = c.labels(SC.NODE)
        step["registers"] = {d.bdf: {"cc": d.cc_mode, "ppcie": d.ppcie_mode} for d in w.devices}
        step["virtual_sleep_s"] = sum(c.clock.sleeps)
        steps.append(step)
        if "exit" in step:
            break
    return {"name": sc["name"], "note": sc.get("note", ""), "steps": steps}


# ------------------------------------------------------------------- watch loop
def install_watch_script(cluster, script):
    from types import SimpleNamespace
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


def run_watch_on_reference(ref, wsc, inject_time=True):
    import kubernetes
    from kubernetes.watch import WatchScriptExhausted
    from _state import world

    w = world()
    w.reset()
    for bdf in SC.GPU_BDFS:
        w.add_gpu(bdf, cc="on")
    c = kubernetes.reset_cluster()
    labels = {} if wsc["initial_label"] is None else {"nvidia.com/cc.mode": wsc["initial_label"]}
    c.add_node(SC.NODE, labels)
    install_watch_script(c, wsc["script"])
    os.environ["EVICT_OPERATOR_COMPONENTS"] = "false"
    if inject_time:
        ref.time = c.clock
    elif hasattr(ref, "time"):
        del ref.time
    applied = []
    mgr = ref.CCManager(node_name=SC.NODE, default_mode=wsc["default"], host_cc=True)
    real = mgr.set_cc_mode
    mgr.set_cc_mode = lambda mode: (applied.append(mode), real(mode))[1]
    saved_ready = ref.create_readiness_file
    ref.create_readiness_file = lambda: applied.append("<readiness>")
    end = None
    try:
        mgr.watch_and_apply()
    except WatchScriptExhausted:
        end = "script-exhausted"
    except BaseException as exc:  # noqa: BLE001 - recorded, that is the point
        end = type(exc).__name__
    finally:
        ref.create_readiness_file = saved_ready
    return {"name": wsc["name"] + ("" if inject_time else "__as_shipped"), "applied": applied, "end": end,
            "watch_resource_versions": [k.get("resource_version") for k in c.watch_calls],
            "watch_kwargs_keys": sorted(c.watch_calls[0]) if c.watch_calls else [],
            "sleeps": list(c.clock.sleeps), "labels": c.labels(SC.NODE),
            "read_node_calls": c.verbs().count("read_node")}


# ------------------------------------------------------------- label algebra
def label_algebra(ref):
    import gpu_operator_eviction as ref_evict
    inputs = [None, "", "false", "true", "paused-for-cc-mode-change", "foo", "foo_paused-for-cc-mode-change",
              "a_b", "_x_", "true_paused-for-cc-mode-change", "paused-for-cc-mode-change_tail"]
    table = [{"input": v, "paused": ref_evict._maybe_set_paused(v), "unpaused": ref_evict._maybe_set_unpaused(v)}
             for v in inputs]
    states = {}
    import kubernetes
    for state in ["on", "off", "devtools", "ppcie", "failed", "weird"]:
        c = kubernetes.reset_cluster()
        c.add_node(SC.NODE, {"keep": "me"})
        ok = ref_evict.set_cc_state_label(kubernetes.client.CoreV1Api(), SC.NODE, state)
        states[state] = {"ok": ok, "labels": c.labels(SC.NODE)}
    return {"pause_table": table, "state_labels": states,
            "component_labels": list(ref_evict.COMPONENT_LABELS),
            "component_app_labels": dict(ref_evict.COMPONENT_APP_LABELS),
            "paused_str": ref_evict.PAUSED_STR,
            "cc_mode_config_label": ref.CC_MODE_CONFIG_LABEL}


def generate():
    ref = load_reference()
    out = {
        "transitions.json": {"generated_by": "oracle/gen_golden.py", "reference": "NVIDIA/k8s-cc-manager@09cd768c main.py",
                             "scenarios": [run_transition_on_reference(ref, sc) for sc in SC.transition_scenarios()]},
        "watch.json": {"generated_by": "oracle/gen_golden.py",
                       "scenarios": [run_watch_on_reference(ref, w) for w in SC.watch_scenarios()]
                       + [run_watch_on_reference(ref, SC.watch_scenarios()[4], inject_time=False)]},
        "labels.json": dict(label_algebra(ref), generated_by="oracle/gen_golden.py"),
    }
    return {name: json.dumps(doc, indent=1, sort_keys=True) + "\n" for name, doc in out.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="fail if committed fixtures differ")
    args = ap.parse_args()
    docs = generate()
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    bad = 0
    for name, text in docs.items():
        path = GOLDEN_DIR / name
        if args.check:
            if not path.exists() or path.read_text() != text:
                print(f"DIFFERS: {path}")
                bad += 1
        else:
            path.write_text(text)
            print(f"wrote {path} ({len(text)} bytes)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
