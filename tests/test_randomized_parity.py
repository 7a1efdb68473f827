"""Beyond the fixed goldens: seeded random scenarios, product (serial engine, libccm sim
backend, fake API server) vs the pinned oracle restatement — every observable compared."""
from __future__ import annotations

import json
import random

import pytest

import scenarios as SC
import transition_oracle as TO
from helpers import run_scenario_on_product

LABEL_VALUES = ["true", "false", "", "custom", "paused-for-cc-mode-change", "x_paused-for-cc-mode-change"]


def random_scenario(seed: int) -> dict:
    rng = random.Random(seed)
    n_gpu = rng.choice([0, 1, 2, 4, 8, 8, 8])
    n_sw = rng.choice([0, 0, 0, 2, 4])
    gpus = []
    uniform_cc = rng.choice(["off", "on", "devtools", None])
    for i in range(n_gpu):
        cc = uniform_cc or rng.choice(["off", "on", "devtools"])
        g = dict(bdf=SC.GPU_BDFS[i], cc=cc, ppcie=rng.choice(["off", "off", "off", "on"]),
                 cc_supported=rng.random() > 0.06, ppcie_supported=rng.random() > 0.1, fail={}, stuck=rng.random() < 0.04)
        if rng.random() < 0.08:
            g["fail"][rng.choice(["query_cc_mode", "set_cc_mode", "reset_with_os", "wait_for_boot", "query_ppcie_mode",
                                  "set_ppcie_mode"])] = rng.choice(["GpuError", "RuntimeError"])
        gpus.append(g)
    sw = []
    for i in range(n_sw):
        s = dict(bdf=SC.SWITCH_BDFS[i], ppcie=rng.choice(["off", "off", "on"]), ppcie_supported=rng.random() > 0.08,
                 fail={}, stuck=rng.random() < 0.04)
        if rng.random() < 0.06:
            s["fail"][rng.choice(["query_ppcie_mode", "set_ppcie_mode", "reset_with_os", "wait_for_boot"])] = "GpuError"
        sw.append(s)
    evict = rng.random() < 0.6
    labels = {name: rng.choice(LABEL_VALUES) for name in SC.COMPONENTS if rng.random() < 0.8}
    if rng.random() < 0.5:
        labels["unrelated.io/thing"] = "keep"
    pods = [dict(app=app, gone_after=rng.choice([0.0, 1.0, 3.0, 7.0])) for app in SC.COMPONENTS.values()
            if rng.random() < 0.5]
    k8s_fail = {}
    if evict and rng.random() < 0.15:
        k8s_fail["patch_node"] = [None] * rng.randrange(0, 3) + [500]
    if evict and rng.random() < 0.1:
        k8s_fail["list_namespaced_pod"] = [503, None, 503]
    modes = [rng.choice(["on", "off", "devtools", "ppcie", "on", "off", ""]) for _ in range(rng.randrange(1, 4))]
    return SC.scenario(f"random_{seed}", gpus_=gpus, switches_=sw, modes=modes, evict=evict, labels=labels, pods=pods,
                       k8s_fail=k8s_fail, host_cc=rng.random() < 0.9)


@pytest.mark.parametrize("seed", range(120))
def test_product_matches_oracle_on_random_scenarios(seed):
    sc = random_scenario(seed)
    want = json.loads(json.dumps(TO.run_scenario(sc, SC.COMPONENTS)))
    got = run_scenario_on_product(sc, max_parallel=1)
    assert len(got["steps"]) == len(want["steps"])
    for g, w in zip(got["steps"], want["steps"]):
        for key in ("result", "exit"):
            assert g.get(key) == w.get(key), (key, sc["modes"])
        for key in ("device_trace", "k8s", "labels", "registers", "virtual_sleep_s"):
            assert g[key] == w[key], (key, g["mode"])
