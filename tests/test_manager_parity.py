"""Control-plane parity: the product's transition engine vs the golden traces recorded
from the UNMODIFIED reference (tests/golden/*.json, made by oracle/gen_golden.py).

Serial mode (CC_MAX_PARALLEL=1) must reproduce the reference's device-op order, k8s
verbs, label maps, return values and exit codes EXACTLY.  Concurrent mode must reach
the same end state with the same multiset of device ops per step.
"""
from __future__ import annotations

import json
from collections import Counter
from pathlib import Path

import pytest

import scenarios as SC
from helpers import run_scenario_on_product

GOLDEN = Path(__file__).parent / "golden"
TRANSITIONS = {s["name"]: s for s in json.loads((GOLDEN / "transitions.json").read_text())["scenarios"]}
SCENARIOS = {s["name"]: s for s in SC.transition_scenarios()}


def test_every_scenario_has_a_golden():
    assert set(TRANSITIONS) == set(SCENARIOS)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_serial_engine_matches_reference_trace(name):
    got = run_scenario_on_product(SCENARIOS[name], max_parallel=1)
    want = TRANSITIONS[name]
    assert len(got["steps"]) == len(want["steps"])
    for g, w in zip(got["steps"], want["steps"]):
        assert g.get("result") == w.get("result"), "return value of set_cc_mode"
        assert g.get("exit") == w.get("exit"), "sys.exit code"
        assert g["device_trace"] == w["device_trace"], "ordered device-op trace"
        assert g["k8s"] == w["k8s"], "ordered k8s API calls incl. patched label maps"
        assert g["labels"] == w["labels"]
        assert g["registers"] == w["registers"]
        assert g["virtual_sleep_s"] == w["virtual_sleep_s"]


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_concurrent_engine_reaches_reference_state(name):
    sc = SCENARIOS[name]
    got = run_scenario_on_product(sc, max_parallel=0)
    want = TRANSITIONS[name]
    faulty = any(d["fail"] or d["stuck"] for d in sc["gpus"] + sc["switches"])
    assert len(got["steps"]) == len(want["steps"])
    for g, w in zip(got["steps"], want["steps"]):
        assert g.get("result") == w.get("result")
        assert g.get("exit") == w.get("exit")
        assert g["labels"] == w["labels"]
        assert g["k8s"] == w["k8s"]
        if not faulty:
            # same work, any interleaving inside a phase; the only extra ops allowed are
            # read-only queries (the concurrent mode_is_set does not short-circuit)
            missing = Counter(w["device_trace"]) - Counter(g["device_trace"])
            extra = Counter(g["device_trace"]) - Counter(w["device_trace"])
            assert not missing
            assert all(line.split()[1].startswith("query_") for line in extra), extra
            assert g["registers"] == w["registers"]
        else:
            # a concurrent phase runs to completion before the error is raised, so it may
            # touch MORE devices than the reference's first-error-stops loop — never fewer
            assert not (Counter(w["device_trace"]) - Counter(g["device_trace"]))


def test_concurrent_phases_keep_reference_ordering():
    """stage-all precedes any reset; every reset precedes any wait (main.py:455-459)."""
    got = run_scenario_on_product(SCENARIOS["roundtrip_off_on_devtools_off_direct"], max_parallel=0)
    for step in got["steps"]:
        ops = [line.split()[1] for line in step["device_trace"] if "ppcie" not in line]
        last_set = max(i for i, o in enumerate(ops) if o == "set_cc_mode")
        first_reset = min(i for i, o in enumerate(ops) if o == "reset_with_os")
        last_reset = max(i for i, o in enumerate(ops) if o == "reset_with_os")
        first_wait = min(i for i, o in enumerate(ops) if o == "wait_for_boot")
        assert last_set < first_reset and last_reset < first_wait
