"""Compatibility module: the reference ships its drain gate under this name
(reference gpu_operator_eviction.py, copied to /app/ by Dockerfile.distroless:40).
Everything is implemented in k8s_cc_manager_b200/drain_gate.py."""
from k8s_cc_manager_b200.drain_gate import (  # noqa: F401
    COMPONENT_APP_LABELS,
    COMPONENT_LABELS,
    PAUSED_STR,
    _maybe_set_paused,
    _maybe_set_unpaused,
    evict_gpu_operator_components,
    fetch_current_component_labels,
    recover_journaled_labels,
    reschedule_gpu_operator_components,
    set_cc_state_label,
)
