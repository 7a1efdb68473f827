#!/usr/bin/env python3
"""BASELINE.json configs[0]: CC-mode query (get-only, no transition) — the CPU-only plumbing
path the reference already runs (reference main.py:232-258: enumerate, mode_is_set over 8
GPUs, write the state label).

Times, on the same host and the same in-memory API server:
  reference   UNMODIFIED /root/reference/main.py + fake gpu-admin-tools (pure Python devices)
  reference+libccm  UNMODIFIED reference main.py, device layer = compat shim -> libccm.so
  product     k8s_cc_manager_b200.manager (concurrent and CC_MAX_PARALLEL=1), libccm.so sim
Only runs where the reference checkout exists (development container); CPU only.
"""
from __future__ import annotations

import importlib.util
import json
import logging
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("CCM_REFERENCE_DIR", "/root/reference"))
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "fakes", ROOT / "oracle"):
    sys.path.insert(0, str(p))
logging.disable(logging.CRITICAL)

import kubernetes  # noqa: E402
import scenarios as SC  # noqa: E402


def load_reference(device_dir: Path, name: str):
    for mod in ("nvidia_gpu_tools", "pci", "pci.devices", "gpu", "gpu_operator_eviction", "_state"):
        sys.modules.pop(mod, None)
    saved = list(sys.path)
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(device_dir))
    spec = importlib.util.spec_from_file_location(name, REF / "main.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.path[:] = saved
    logging.disable(logging.CRITICAL)
    return mod


def timeit(fn, reps=300, warm=20):
    for _ in range(warm):
        fn()
    xs = []
    for _ in range(reps):
        t0 = time.perf_counter()
        assert fn() is True
        xs.append((time.perf_counter() - t0) * 1e6)
    xs.sort()
    return {"median_us": statistics.median(xs), "p10_us": xs[len(xs) // 10], "p90_us": xs[9 * len(xs) // 10], "reps": reps}


def main():
    out = {"host_cpus": os.cpu_count(), "gpus": 8, "mode": "on", "what": "set_cc_mode('on') with all 8 GPUs already 'on'"}
    os.environ["EVICT_OPERATOR_COMPONENTS"] = "true"

    # --- reference + pure-Python fake devices
    ref = load_reference(ROOT / "oracle" / "fakes" / "gpu-admin-tools", "ref_fake")
    from _state import world
    w = world(); w.reset()
    for bdf in SC.GPU_BDFS:
        w.add_gpu(bdf, cc="on")
    c = kubernetes.reset_cluster(); c.add_node(SC.NODE, SC.all_true_labels())
    mgr = ref.CCManager(node_name=SC.NODE, default_mode="on", host_cc=True)

    def run_ref():
        w.trace.clear(); c.calls.clear()
        return mgr.set_cc_mode("on")
    out["reference_fake_devices"] = timeit(run_ref)

    # --- reference + libccm through the compat shim
    from helpers import build_native_world
    from k8s_cc_manager_b200 import _native as N
    build_native_world(SC.scenario("c1", gpus_=SC.gpus(8, cc="on"), modes=[]))
    ref2 = load_reference(ROOT / "k8s_cc_manager_b200" / "compat" / "gpu-admin-tools", "ref_libccm")
    c = kubernetes.reset_cluster(); c.add_node(SC.NODE, SC.all_true_labels())
    mgr2 = ref2.CCManager(node_name=SC.NODE, default_mode="on", host_cc=True)

    def run_ref2():
        N.lib().ccm_sim_trace_clear(); c.calls.clear()
        return mgr2.set_cc_mode("on")
    out["reference_on_libccm"] = timeit(run_ref2)

    # --- product
    from k8s_cc_manager_b200 import manager
    for label, par in (("product_concurrent", 0), ("product_serial", 1)):
        c = kubernetes.reset_cluster(); c.add_node(SC.NODE, SC.all_true_labels())
        m = manager.CCManager(SC.NODE, "on", True, max_parallel=par, scrub_mode="skip")

        def run_prod():
            N.lib().ccm_sim_trace_clear(); c.calls.clear()
            return m.set_cc_mode("on")
        out[label] = timeit(run_prod)
        assert c.labels(SC.NODE)["nvidia.com/cc.mode.state"] == "on"
    print(json.dumps(out, indent=1))
    (ROOT / "profiles" / "r1_config1_get_only.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
