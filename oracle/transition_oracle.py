"""CPU restatement of the reference's CC-transition algorithm as a pure function.

ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under k8s_cc_manager_b200/ imports this
file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.

What it restates (reference @ 09cd768c):
  set_cc_mode dispatcher ............ main.py:214-263
  set_ppcie_mode .................... main.py:265-296
  mode_is_set / ppcie_mode_is_set ... main.py:428-447 / 298-315
  _set_cc_mode_direct ............... main.py:449-542
  _set_ppcie_mode_direct ............ main.py:317-391
  _set_*_with_eviction .............. main.py:544-578 / 393-426
  fetch/evict/reschedule/state label  gpu_operator_eviction.py:98-128,131-214,217-259,262-295
  pause / un-pause value mapping .... gpu_operator_eviction.py:43-95

Pinned: tests/test_oracle_golden.py checks it step by step against
tests/golden/transitions.json + labels.json, which oracle/gen_golden.py produced by
running the UNMODIFIED reference main.py.  The device register arithmetic itself
(bit layout of the CC-mode register) lives in NVIDIA/gpu-admin-tools v2025.11.21,
which is not in the reference tree: for that layer parity is UNPINNED and only the
stage -> reset -> read-back contract visible at the reference's call sites is modelled.

Model: state is plain dicts; every device op appends "<bdf> <op> <arg>" to a trace,
every API call appends [verb, ...] to a k8s trace, exactly the shape gen_golden.py
records, so outputs compare with ==.
"""
from __future__ import annotations

import copy

PAUSED = "paused-for-cc-mode-change"
STATE_LABEL = "nvidia.com/cc.mode.state"
READY_LABEL = "nvidia.com/cc.ready.state"
POLL_S = 2.0
EVICT_TIMEOUT_S = 300.0


class _Exit(Exception):
    def __init__(self, code):
        self.code = code


class _DeviceFault(Exception):
    pass


class _ApiFault(Exception):
    pass


# -------------------------------------------------- label algebra (eviction.py:43-95)
def pause_value(v):
    if v is None or v == "":
        return ""
    if v in ("false",):
        return v
    if v == "true":
        return PAUSED
    return v if PAUSED in v else f"{v}_{PAUSED}"


def unpause_value(v):
    if v == "false":
        return v
    if v == PAUSED:
        return "true"
    if v and PAUSED in v:
        return v.replace("_" + PAUSED, "").replace(PAUSED, "").strip("_")
    return v or ""


def ready_for(state):
    return {"on": "true", "ppcie": "true", "off": "false"}.get(state, "")


# ----------------------------------------------------------------- world model
class World:
    def __init__(self, sc, components):
        self.components = components  # deploy label -> app label, in reference order
        self.devs = []
        for g in sc["gpus"]:
            self.devs.append(dict(bdf=g["bdf"], gpu=True, cc=g["cc"], cc_st=g["cc"], pp=g["ppcie"], pp_st=g["ppcie"],
                                  cc_ok=g["cc_supported"], pp_ok=g["ppcie_supported"], fail=dict(g["fail"]),
                                  stuck=g["stuck"]))
        for s in sc["switches"]:
            self.devs.append(dict(bdf=s["bdf"], gpu=False, cc="off", cc_st="off", pp=s["ppcie"], pp_st=s["ppcie"],
                                  cc_ok=False, pp_ok=s["ppcie_supported"], fail=dict(s["fail"]), stuck=s["stuck"]))
        self.labels = dict(sc.get("labels", {}))
        self.pods = [dict(app=p["app"], after=p["gone_after"], gone_at=None) for p in sc.get("pods", [])]
        self.fail_q = {verb: list(q) for verb, q in sc.get("k8s_fail", {}).items()}
        self.now = 0.0
        self.slept = 0.0
        self.dev_trace = []
        self.k8s = []
        self.evict = sc["evict"]

    # ---- device ops ---------------------------------------------------------
    def _op(self, d, op, arg="-"):
        self.dev_trace.append(f"{d['bdf']} {op} {arg}")
        if op in d["fail"]:
            raise _DeviceFault(op)

    def query(self, d, reg):
        op = "query_cc_mode" if reg == "cc" else "query_ppcie_mode"
        if op in d["fail"]:
            self._op(d, op, "error")
        val = d["cc"] if reg == "cc" else d["pp"]
        self._op(d, op, val)
        return val

    def stage(self, d, reg, val):
        self._op(d, "set_cc_mode" if reg == "cc" else "set_ppcie_mode", val)
        d["cc_st" if reg == "cc" else "pp_st"] = val

    def reset(self, d):
        self._op(d, "reset_with_os")
        if d["stuck"]:
            d["cc_st"], d["pp_st"] = d["cc"], d["pp"]
        else:
            d["cc"], d["pp"] = d["cc_st"], d["pp_st"]

    def wait(self, d):
        self._op(d, "wait_for_boot")

    # ---- API server ----------------------------------------------------------
    def _api(self, verb, *rec):
        self.k8s.append([verb, *rec])
        q = self.fail_q.get(verb)
        if q:
            status = q.pop(0)
            if status is not None:
                raise _ApiFault(status)

    def read_node(self):
        self._api("read_node")
        return dict(self.labels)

    def patch_labels(self, updates):
        """read-modify-write of the whole label map (eviction.py:165-170,247-252,282-288)."""
        labels = self.read_node()
        labels.update(updates)
        self._api("patch_node", dict(labels))
        self.labels = labels
        for p in self.pods:
            if p["gone_at"] is None and p["after"] is not None:
                p["gone_at"] = self.now + p["after"]

    def pods_left(self, app):
        self._api("list_namespaced_pod", f"app={app}")
        return sum(1 for p in self.pods if p["app"] == app and not (p["gone_at"] is not None and self.now >= p["gone_at"]))

    def sleep(self, s):
        self.now += s
        self.slept += s


# --------------------------------------------------------------- drain gate
def set_state_label(w, state):
    """eviction.py:262-295 (API errors are swallowed there)."""
    try:
        w.patch_labels({STATE_LABEL: state, READY_LABEL: ready_for(state)})
    except _ApiFault:
        pass


def with_eviction(w, transition):
    """main.py:544-578: fetch, evict (pause + wait), transition, reschedule (always)."""
    labels = w.read_node()  # fetch_current_component_labels; ApiException would propagate
    current = {name: labels.get(name, "") for name in w.components}
    try:
        w.patch_labels({name: pause_value(v) for name, v in current.items()})
        for name, v in current.items():
            if not v:
                continue
            app = w.components[name]
            start = w.now
            while w.now - start < EVICT_TIMEOUT_S:
                try:
                    if w.pods_left(app) == 0:
                        break
                except _ApiFault:
                    pass
                w.sleep(POLL_S)
    except _ApiFault:
        return False  # eviction failed: no device is touched, nothing is restored
    result = transition()
    try:
        w.patch_labels({name: unpause_value(v) for name, v in current.items()})
    except _ApiFault:
        result = False
    return result


# ------------------------------------------------------------ transition engine
def stage_reset_verify(w, devs, reg, target):
    """stage all -> reset all staged -> wait + read back each (main.py:502-529, 349-378, 471-500)."""
    staged = []
    for d in devs:
        if w.query(d, reg) != target:
            w.stage(d, reg, target)
            staged.append(d)
    for d in staged:
        w.reset(d)
    for d in staged:
        w.wait(d)
        if w.query(d, reg) != target:
            raise _DeviceFault("verify")


def cc_direct(w, gpus, mode):
    try:
        stage_reset_verify(w, [d for d in w.devs if d["pp_ok"]], "pp", "off")  # main.py:471-500
        stage_reset_verify(w, gpus, "cc", mode)                                # main.py:502-529
    except _DeviceFault:
        set_state_label(w, "failed")
        return False
    set_state_label(w, mode)
    return True


def ppcie_direct(w, devs):
    try:
        for d in devs:  # main.py:339-347: each device on its own, reset right away
            if w.query(d, "pp") != "off":
                w.stage(d, "pp", "off")
                w.reset(d)
                w.wait(d)
        stage_reset_verify(w, devs, "pp", "on")
    except _DeviceFault:
        set_state_label(w, "failed")
        return False
    set_state_label(w, "ppcie")
    return True


def all_report(w, devs, reg, target):
    """mode_is_set / ppcie_mode_is_set: first mismatch or error ends the scan."""
    for d in devs:
        try:
            if w.query(d, reg) != target:
                return False
        except _DeviceFault:
            return False
    return True


def set_cc_mode(w, mode):
    """main.py:214-263 (+ 265-296 for 'ppcie')."""
    if mode == "ppcie":
        devs = list(w.devs)
        capable = [d for d in devs if d["pp_ok"]]
        if len(devs) != len(capable):
            raise _Exit(1)
        if not devs:
            return True
        if all_report(w, devs, "pp", "on"):
            set_state_label(w, "ppcie")
            return True
        run = lambda: ppcie_direct(w, devs)  # noqa: E731
        return with_eviction(w, run) if w.evict else run()

    gpus = [d for d in w.devs if d["gpu"]]
    capable = [d for d in gpus if d["cc_ok"]]
    if mode != "off" and len(gpus) != len(capable):
        raise _Exit(1)
    if not gpus or not mode:
        return True
    if not capable:
        set_state_label(w, "off")
        return True
    if all_report(w, capable, "cc", mode):
        set_state_label(w, mode)
        return True
    run = lambda: cc_direct(w, capable, mode)  # noqa: E731
    return with_eviction(w, run) if w.evict else run()


def run_scenario(sc, components):
    """Same record layout as oracle/gen_golden.py:run_transition_on_reference."""
    w = World(copy.deepcopy(sc), components)
    steps = []
    for mode in sc["modes"]:
        n_dev, n_k8s = len(w.dev_trace), len(w.k8s)
        step = {"mode": mode}
        try:
            step["result"] = set_cc_mode(w, mode)
        except _Exit as e:
            step["exit"] = e.code
        step["device_trace"] = w.dev_trace[n_dev:]
        step["k8s"] = w.k8s[n_k8s:]
        step["labels"] = dict(w.labels)
        step["registers"] = {d["bdf"]: {"cc": d["cc"], "ppcie": d["pp"]} for d in w.devs}
        step["virtual_sleep_s"] = w.slept
        steps.append(step)
        if "exit" in step:
            break
    return {"name": sc["name"], "note": sc.get("note", ""), "steps": steps}
