"""Property tests (hypothesis) of the pause / un-pause label algebra: product == oracle on
arbitrary strings, plus the algebraic facts the drain gate relies on."""
from __future__ import annotations

from hypothesis import given, settings, strategies as st

import transition_oracle as TO
from k8s_cc_manager_b200 import drain_gate as G

P = G.PAUSED_STR
label_text = st.one_of(
    st.none(), st.sampled_from(["", "true", "false", P, f"x_{P}", f"{P}_y", f"_{P}_", "a_b", "TRUE", "False"]),
    st.text(alphabet="abcXYZ019_-.", max_size=12),
    st.builds(lambda a, b: a + P + b, st.text(alphabet="ab_", max_size=4), st.text(alphabet="ab_", max_size=4)),
)


@settings(max_examples=400, deadline=None)
@given(label_text)
def test_product_equals_oracle(v):
    assert G._maybe_set_paused(v) == TO.pause_value(v)
    assert G._maybe_set_unpaused(v) == TO.unpause_value(v)


@settings(max_examples=400, deadline=None)
@given(label_text)
def test_pausing_is_idempotent_and_never_enables(v):
    once = G._maybe_set_paused(v)
    assert G._maybe_set_paused(once) == once
    assert once != "true"                       # a paused component is never left enabled
    if v in (None, "", "false"):
        assert once == (v or "")                # user-disabled components stay exactly as they were


@settings(max_examples=400, deadline=None)
@given(st.text(alphabet="abcXYZ019-.", min_size=1, max_size=12).filter(lambda s: s not in ("true", "false")))
def test_custom_values_round_trip(v):
    assert G._maybe_set_unpaused(G._maybe_set_paused(v)) == v
    assert G._maybe_set_unpaused(v) == v        # restore is fed ORIGINAL values: identity (main.py:556,570-574)


def test_true_round_trips():
    assert G._maybe_set_unpaused(G._maybe_set_paused("true")) == "true"
