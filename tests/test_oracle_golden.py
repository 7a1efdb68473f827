"""Pins the ORACLE before it is trusted: the CPU restatements under oracle/ against the
fixtures in tests/golden/ (produced from the unmodified reference by
oracle/gen_golden.py) and against independent numpy statements."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

import scenarios as SC
import scrub_oracle as SO
import transition_oracle as TO

GOLDEN = Path(__file__).parent / "golden"
TRANSITIONS = {s["name"]: s for s in json.loads((GOLDEN / "transitions.json").read_text())["scenarios"]}
LABELS = json.loads((GOLDEN / "labels.json").read_text())
VECTORS = json.loads((GOLDEN / "scrub_vectors.json").read_text())


@pytest.mark.parametrize("sc", SC.transition_scenarios(), ids=lambda s: s["name"])
def test_transition_oracle_reproduces_reference(sc):
    got = json.loads(json.dumps(TO.run_scenario(sc, SC.COMPONENTS)))
    assert got == TRANSITIONS[sc["name"]]


def test_oracle_label_algebra_matches_reference_table():
    for row in LABELS["pause_table"]:
        assert TO.pause_value(row["input"]) == row["paused"], row
        assert TO.unpause_value(row["input"]) == row["unpaused"], row
    for state, rec in LABELS["state_labels"].items():
        assert rec["labels"]["nvidia.com/cc.ready.state"] == TO.ready_for(state)
    assert SC.COMPONENTS == LABELS["component_app_labels"]
    assert list(SC.COMPONENTS) == LABELS["component_labels"]
    assert TO.PAUSED == LABELS["paused_str"]


def test_survey_known_answers():
    """SURVEY.md §4 tables (derived by hand from the reference's pure functions)."""
    paused = "paused-for-cc-mode-change"
    table = {None: ("", ""), "": ("", ""), "false": ("false", "false"), "true": (paused, "true"),
             paused: (paused, "true"), "foo": (f"foo_{paused}", "foo"), f"foo_{paused}": (f"foo_{paused}", "foo")}
    for v, (p, u) in table.items():
        assert TO.pause_value(v) == p
        assert TO.unpause_value(v) == u
    assert [TO.ready_for(s) for s in ("on", "ppcie", "off", "devtools", "failed")] == ["true", "true", "false", "", ""]


# ------------------------------------------------------------------ scrub oracle
@pytest.mark.parametrize("vec", VECTORS["vectors"], ids=lambda v: v["name"])
def test_scrub_oracle_known_answer_vectors(vec):
    buf = np.zeros(vec["nbytes"], dtype=np.uint8)
    if vec["fill"] is not None:
        buf[:] = vec["fill"]
    for off, val in vec["poke"]:
        buf[off] = val
    assert SO.count_nonzero_c(buf) == vec["nonzero"]
    assert SO.count_nonzero_np(buf) == vec["nonzero"]
    assert SO.scrub_verify_mt_c(buf.copy(), 3, scrub=False) == vec["nonzero"]
    SO.scrub_c(buf)
    assert not buf.any() and SO.count_nonzero_c(buf) == 0


@pytest.mark.parametrize("nbytes,seed,word0", [(0, 1, 0), (7, 1, 0), (8, 2, 3), (4096 + 5, 1234, 0), (1 << 20, 99, 1 << 33)])
def test_pattern_statements_agree(nbytes, seed, word0):
    c = np.zeros(nbytes, dtype=np.uint8)
    SO.fill_pattern_c(c, seed, word0)
    n = SO.pattern_np(nbytes, seed, word0)
    assert np.array_equal(c, n)
    assert SO.pattern_count_c(nbytes, seed, word0) == SO.count_nonzero_np(n)
    for j in range(min(nbytes // 8, 64)):
        assert SO.clib().ccm_oracle_pattern_word(seed, word0 + j) == SO.pattern_word_py(seed, word0 + j)


def test_pattern_vector_is_pinned():
    """The pattern itself is a fixture: a silent change of the formula must fail here."""
    v = VECTORS["pattern"]
    buf = np.zeros(v["nbytes"], dtype=np.uint8)
    SO.fill_pattern_c(buf, v["seed"], v["word_index0"])
    assert SO.count_nonzero_c(buf) == v["nonzero"]
    assert buf[: len(v["first_bytes"])].tolist() == v["first_bytes"]
    import hashlib
    assert hashlib.sha256(buf.tobytes()).hexdigest() == v["sha256"]


def test_mt_oracle_matches_single_thread():
    rng = np.random.default_rng(7)
    buf = rng.integers(0, 4, size=(1 << 20) + 13, dtype=np.uint8)
    want = SO.count_nonzero_np(buf)
    for threads in (1, 2, 5, 16):
        assert SO.scrub_verify_mt_c(buf.copy(), threads, scrub=False) == want
    assert SO.scrub_verify_mt_c(buf, 4, scrub=True) == 0
