"""Transition / watch-loop scenarios shared by the golden generator, the oracle and
the parity tests (TEST INFRASTRUCTURE — nothing in the product imports this).

A scenario is plain data: the node's devices and their register state, the node's
labels, the operator pods, injected faults, and the sequence of modes to apply.  The
same scenario is run through
  (1) the UNMODIFIED reference main.py            -> oracle/gen_golden.py -> tests/golden/
  (2) the CPU restatement oracle/transition_oracle.py
  (3) the product (k8s_cc_manager_b200.manager + libccm.so sim backend)
and the observable results are compared: set_cc_mode return values / exit codes,
the ordered device-op trace, the k8s verbs, the node labels after every step.
"""
from __future__ import annotations

GPU_BDFS = ["0000:1b:00.0", "0000:43:00.0", "0000:52:00.0", "0000:61:00.0",
            "0000:9d:00.0", "0000:c3:00.0", "0000:d1:00.0", "0000:df:00.0"]
SWITCH_BDFS = ["0000:05:00.0", "0000:06:00.0", "0000:07:00.0", "0000:08:00.0"]

NODE = "gpu-node-0"
NAMESPACE = "gpu-operator"

COMPONENTS = {
    "nvidia.com/gpu.deploy.vfio-manager": "nvidia-vfio-manager",
    "nvidia.com/gpu.deploy.vgpu-manager": "nvidia-vgpu-manager",
    "nvidia.com/gpu.deploy.sandbox-validator": "nvidia-sandbox-validator",
    "nvidia.com/gpu.deploy.sandbox-device-plugin": "nvidia-sandbox-device-plugin-daemonset",
    "nvidia.com/gpu.deploy.vgpu-device-manager": "nvidia-vgpu-device-manager",
}


def gpus(n=8, cc="off", ppcie="off", **kw):
    return [dict(bdf=GPU_BDFS[i], cc=cc, ppcie=ppcie, cc_supported=True, ppcie_supported=True,
                 fail={}, stuck=False, **kw) for i in range(n)]


def switches(n=4, ppcie="off"):
    return [dict(bdf=SWITCH_BDFS[i], ppcie=ppcie, ppcie_supported=True, fail={}, stuck=False)
            for i in range(n)]


def all_true_labels():
    return {k: "true" for k in COMPONENTS}


def pods_for_all(gone_after=0.0):
    return [dict(app=app, gone_after=gone_after) for app in COMPONENTS.values()]


def scenario(name, *, gpus_=None, switches_=None, modes, evict=True, labels=None, pods=None,
             k8s_fail=None, host_cc=True, note=""):
    return dict(name=name, gpus=gpus_ if gpus_ is not None else gpus(), switches=switches_ or [],
                modes=list(modes), evict=evict, labels=dict(labels or {}), pods=list(pods or []),
                k8s_fail=dict(k8s_fail or {}), host_cc=host_cc, note=note)


def _with(devs, index, **changes):
    out = [dict(d, fail=dict(d["fail"])) for d in devs]
    for k, v in changes.items():
        if k == "fail":
            out[index]["fail"].update(v)
        else:
            out[index][k] = v
    return out


def transition_scenarios():
    S = []
    # config 1 (BASELINE.json configs[0]): get-only, no transition — reference main.py:232-258
    S.append(scenario("get_only_8gpu_on", gpus_=gpus(8, cc="on"), modes=["on"], labels=all_true_labels(),
                      note="mode_is_set short-circuit + state label"))
    S.append(scenario("get_only_8gpu_off", gpus_=gpus(8, cc="off"), modes=["off"]))
    # config 2/3: off -> on, gated by eviction, pods already gone
    S.append(scenario("off_to_on_evict_pods_gone", modes=["on"], labels=all_true_labels()))
    # pods take 3 s to terminate: the 2 s poll loop of gpu_operator_eviction.py:186-204
    S.append(scenario("off_to_on_evict_pods_3s", modes=["on"], labels=all_true_labels(),
                      pods=pods_for_all(3.0)))
    # one component's pods never go away: timeout is logged and ignored (…eviction.py:205-207)
    S.append(scenario("off_to_on_evict_timeout", modes=["on"], labels=all_true_labels(),
                      pods=[dict(app="nvidia-vfio-manager", gone_after=None)]))
    # config 4: on -> devtools -> off round trip, eviction gated
    S.append(scenario("roundtrip_on_devtools_off", gpus_=gpus(8, cc="on"), modes=["devtools", "off"],
                      labels=all_true_labels(), pods=pods_for_all(0.0)))
    S.append(scenario("roundtrip_off_on_devtools_off_direct", modes=["on", "devtools", "off"], evict=False))
    S.append(scenario("direct_single_gpu", gpus_=gpus(1), modes=["on"], evict=False))
    # partially applied: only GPUs not yet in the mode are staged and reset
    S.append(scenario("partial_3_of_8_already_on",
                      gpus_=[dict(g, cc="on" if i in (0, 4, 7) else "off") for i, g in enumerate(gpus())],
                      modes=["on"], evict=False))
    # PPCIe currently on everywhere: phase 1 of _set_cc_mode_direct turns it off first
    S.append(scenario("cc_on_with_ppcie_active", gpus_=gpus(8, ppcie="on"), switches_=switches(4, "on"),
                      modes=["on"], evict=False))
    S.append(scenario("cc_on_with_switches_idle", switches_=switches(4), modes=["on"], evict=False))
    # PPCIe mode itself (main.py:265-391)
    S.append(scenario("ppcie_from_off", switches_=switches(4), modes=["ppcie"], evict=False))
    S.append(scenario("ppcie_from_off_evict", switches_=switches(4), modes=["ppcie"],
                      labels=all_true_labels(), pods=pods_for_all(0.0)))
    S.append(scenario("ppcie_already_set", gpus_=gpus(8, ppcie="on"), switches_=switches(4, "on"),
                      modes=["ppcie"]))
    S.append(scenario("ppcie_partial", gpus_=[dict(g, ppcie="on" if i < 2 else "off") for i, g in enumerate(gpus())],
                      switches_=switches(2), modes=["ppcie"], evict=False))
    S.append(scenario("ppcie_unsupported_switch_exits",
                      switches_=_with(switches(2), 1, ppcie_supported=False), modes=["ppcie"], evict=False))
    # faults: GpuError while staging -> 'failed', reschedule still runs
    S.append(scenario("fault_set_gpuerror_gpu3", gpus_=_with(gpus(), 3, fail={"set_cc_mode": "GpuError"}),
                      modes=["on"], labels=all_true_labels()))
    S.append(scenario("fault_reset_runtimeerror_gpu0", gpus_=_with(gpus(), 0, fail={"reset_with_os": "RuntimeError"}),
                      modes=["on"], evict=False))
    S.append(scenario("fault_wait_gpuerror_gpu7", gpus_=_with(gpus(), 7, fail={"wait_for_boot": "GpuError"}),
                      modes=["on"], evict=False))
    S.append(scenario("fault_stuck_mode_gpu5", gpus_=_with(gpus(), 5, stuck=True), modes=["on"], evict=False,
                      note="read-back mismatch -> RuntimeError -> failed"))
    S.append(scenario("fault_query_error_in_mode_is_set",
                      gpus_=_with(gpus(8, cc="on"), 2, fail={"query_cc_mode": "GpuError"}),
                      modes=["on"], evict=False,
                      note="mode_is_set treats the error as 'not set' (main.py:444-446); staging then fails"))
    S.append(scenario("fault_ppcie_set_gpuerror", switches_=_with(switches(4), 1, fail={"set_ppcie_mode": "GpuError"}),
                      modes=["ppcie"], evict=False))
    # capability checks (main.py:237-253)
    S.append(scenario("noncapable_gpu_mode_on_exits", gpus_=_with(gpus(), 6, cc_supported=False),
                      modes=["on"], evict=False))
    S.append(scenario("noncapable_gpu_mode_off", gpus_=_with(gpus(8, cc="on"), 6, cc_supported=False),
                      modes=["off"], evict=False))
    S.append(scenario("no_gpus", gpus_=[], modes=["on"], evict=False))
    S.append(scenario("empty_mode_is_noop", modes=[""], evict=False))
    S.append(scenario("no_cc_capable_gpus_mode_off",
                      gpus_=[dict(g, cc_supported=False) for g in gpus(2)], modes=["off"], evict=False))
    S.append(scenario("host_without_cc_mode_on", modes=["on"], evict=False, host_cc=False))
    # drain-gate corner cases
    S.append(scenario("evict_patch_fails", modes=["on"], labels=all_true_labels(),
                      k8s_fail={"patch_node": [500]}, note="eviction fails -> False, no device op"))
    S.append(scenario("evict_custom_label_values", modes=["on"],
                      labels={"nvidia.com/gpu.deploy.vfio-manager": "true",
                              "nvidia.com/gpu.deploy.vgpu-manager": "false",
                              "nvidia.com/gpu.deploy.sandbox-validator": "custom",
                              "nvidia.com/gpu.deploy.sandbox-device-plugin": "paused-for-cc-mode-change",
                              "unrelated/label": "keep"},
                      pods=pods_for_all(0.0)))
    S.append(scenario("reschedule_fails", modes=["on"], labels=all_true_labels(),
                      k8s_fail={"patch_node": [None, None, 500]},
                      note="3rd patch (restore labels) fails -> result False although GPUs are on"))
    return S


def watch_scenarios():
    """Scripts for CCManager.watch_and_apply (reference main.py:600-684).

    `script` is a list of watch batches.  A batch is a list of events
    {type, label, rv} (label None = label absent) or the string 'ERROR'; a batch given
    as {'raise': status} makes stream() raise ApiException(status).  After the script
    is exhausted the fake Watch raises WatchScriptExhausted, which ends the loop.
    `relabel` entries change the node's label on the API server before a batch
    (what a 410 re-sync then reads back).
    """
    def ev(kind, label, rv):
        return dict(type=kind, label=label, rv=rv)

    W = []
    W.append(dict(name="watch_label_change_and_noop", initial_label="on", default="on",
                  script=[[ev("MODIFIED", "on", "11"), ev("MODIFIED", "off", "12"), ev("MODIFIED", "off", "13"),
                           ev("DELETED", "off", "14"), ev("ADDED", "devtools", "15")]]))
    W.append(dict(name="watch_label_removed_applies_default", initial_label="off", default="on",
                  script=[[ev("MODIFIED", None, "21")], [ev("MODIFIED", "off", "22")]]))
    W.append(dict(name="watch_error_event_reconnects", initial_label="on", default="on",
                  script=[["ERROR"], [ev("MODIFIED", "off", "31")]]))
    W.append(dict(name="watch_410_resync", initial_label="on", default="on",
                  script=[{"raise": 410, "relabel": "devtools"}, [ev("MODIFIED", "devtools", "41")]]))
    W.append(dict(name="watch_500_then_recover", initial_label="on", default="on",
                  script=[{"raise": 500}, {"raise": 500}, [ev("MODIFIED", "off", "51")]]))
    W.append(dict(name="watch_ten_errors_fatal", initial_label="on", default="on",
                  script=[{"raise": 500}] * 10))
    return W
