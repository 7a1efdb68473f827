"""Shared helpers: run a scenario (oracle/scenarios.py) through the PRODUCT."""
from __future__ import annotations

import ctypes as C
import os

import scenarios as SC

from k8s_cc_manager_b200 import _native as N

OP_BITS = {"query_cc_mode": N.OP_QUERY_CC, "set_cc_mode": N.OP_SET_CC, "query_ppcie_mode": N.OP_QUERY_PPCIE,
           "set_ppcie_mode": N.OP_SET_PPCIE, "reset_with_os": N.OP_RESET, "wait_for_boot": N.OP_WAIT_BOOT}


def sim_set(dev, key, value):
    rc = N.lib().ccm_sim_set(dev, key.encode(), int(value))
    assert rc == 0, N.last_error()


def sim_get(dev, key):
    v = C.c_int64()
    rc = N.lib().ccm_sim_get(dev, key.encode(), C.byref(v))
    assert rc == 0, N.last_error()
    return v.value


def sim_trace():
    buf = C.create_string_buffer(1 << 22)
    N.lib().ccm_sim_trace(buf, len(buf))
    return [line.split(" ", 1)[1] for line in buf.value.decode().splitlines()]  # drop the seq number


def sim_trace_clear():
    N.lib().ccm_sim_trace_clear()


def build_native_world(sc):
    """Configure libccm's sim backend from a scenario (mirrors oracle/gen_golden.build_world)."""
    lib = N.lib()
    assert lib.ccm_sim_topology(len(sc["gpus"]), len(sc["switches"])) == 0, N.last_error()
    devs = list(sc["gpus"]) + list(sc["switches"])
    for i, d in enumerate(devs):
        if "cc" in d:
            sim_set(i, "cc_mode", N.CC_MODES[d["cc"]])
            sim_set(i, "cc_supported", d["cc_supported"])
        sim_set(i, "ppcie_mode", N.PPCIE_MODES[d["ppcie"]])
        sim_set(i, "ppcie_supported", d["ppcie_supported"])
        sim_set(i, "stuck", d["stuck"])
        mask = 0
        for op in d["fail"]:
            mask |= OP_BITS[op]
        sim_set(i, "fail_op", mask)
        sim_set(i, "cuda_ordinal", -1)  # control-plane tests never touch CUDA
    sim_trace_clear()


def build_cluster(sc):
    import kubernetes
    from kubernetes.client.rest import ApiException
    c = kubernetes.reset_cluster()
    c.add_node(SC.NODE, dict(sc.get("labels", {})))
    for p in sc.get("pods", []):
        c.add_pod(p["app"], SC.NODE, namespace=SC.NAMESPACE, gone_after=p["gone_after"])
    for verb, statuses in sc.get("k8s_fail", {}).items():
        c.fail[verb] = [None if s is None else ApiException(status=s, reason="injected") for s in statuses]
    return c


def k8s_trace(cluster):
    out = []
    for verb, args in cluster.calls:
        if verb == "patch_node":
            out.append(["patch_node", args[1] if isinstance(args[1], dict) else None])
        elif verb == "list_namespaced_pod":
            out.append([verb, args[2]])
        else:
            out.append([verb])
    return out


def registers():
    from k8s_cc_manager_b200.devices import find_gpus
    out = {}
    for d in find_gpus()[0]:
        out[d.bdf] = {"cc": N.CC_MODE_NAMES[sim_get(d.index, "cc_mode")],
                      "ppcie": N.PPCIE_MODE_NAMES[sim_get(d.index, "ppcie_mode")]}
    return out


def run_scenario_on_product(sc, *, max_parallel=1, scrub_mode="skip", monkeypatch=None):
    """Same record layout as oracle/gen_golden.run_transition_on_reference."""
    from k8s_cc_manager_b200 import drain_gate, manager

    build_native_world(sc)
    c = build_cluster(sc)
    os.environ["EVICT_OPERATOR_COMPONENTS"] = "true" if sc["evict"] else "false"
    os.environ["OPERATOR_NAMESPACE"] = SC.NAMESPACE
    # virtual clock for the 2 s pod polls
    saved = (drain_gate._now, drain_gate._pause)
    drain_gate._now, drain_gate._pause = c.clock.time, c.clock.sleep
    try:
        mgr = manager.CCManager(node_name=SC.NODE, default_mode="on", host_cc=sc["host_cc"],
                                max_parallel=max_parallel, scrub_mode=scrub_mode)
        steps = []
        for mode in sc["modes"]:
            n_k8s = len(c.calls)
            sim_trace_clear()
            step = {"mode": mode}
            try:
                step["result"] = mgr.set_cc_mode(mode)
            except SystemExit as exc:
                step["exit"] = exc.code
            step["device_trace"] = sim_trace()
            step["k8s"] = k8s_trace(c)[n_k8s:]
            step["labels"] = c.labels(SC.NODE)
            step["registers"] = registers()
            step["virtual_sleep_s"] = sum(c.clock.sleeps)
            steps.append(step)
            if "exit" in step:
                break
        return {"name": sc["name"], "note": sc.get("note", ""), "steps": steps}
    finally:
        drain_gate._now, drain_gate._pause = saved
