"""Drain gate: pause the GPU Operator's node-local components around a CC transition.

Behavioural mirror of the reference module gpu_operator_eviction.py (same public
function names, arguments, return values and label algebra), because the manager
and the GPU Operator only ever observe its effects on node labels:

  fetch_current_component_labels   reference gpu_operator_eviction.py:98-128
  evict_gpu_operator_components    reference gpu_operator_eviction.py:131-214
  reschedule_gpu_operator_components  reference gpu_operator_eviction.py:217-259
  set_cc_state_label               reference gpu_operator_eviction.py:262-295
  pause / un-pause value mapping   reference gpu_operator_eviction.py:43-95

The data path (HBM scrub) never passes through here; this is k8s-API-latency-bound
control plane and is kept functionally identical on purpose.

Two opt-in hardening features beyond the reference, both off
unless asked for so default observable behaviour stays the reference's:
  * concurrent_wait=True waits for the five components' pods in parallel threads
    (the reference waits for them one after another, each up to `timeout` seconds);
  * journal_annotation persists the ORIGINAL label values in a node annotation
    before pausing, so a manager that crashes between evict and reschedule can
    recover them (the reference keeps them only in memory, main.py:556).
"""
from __future__ import annotations

import json
import logging
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Optional

from kubernetes.client.rest import ApiException

logger = logging.getLogger(__name__)

# component deploy label -> app label of the DaemonSet pods to wait for
COMPONENT_APP_LABELS: Dict[str, str] = {
    "nvidia.com/gpu.deploy.vfio-manager": "nvidia-vfio-manager",
    "nvidia.com/gpu.deploy.vgpu-manager": "nvidia-vgpu-manager",
    "nvidia.com/gpu.deploy.sandbox-validator": "nvidia-sandbox-validator",
    "nvidia.com/gpu.deploy.sandbox-device-plugin": "nvidia-sandbox-device-plugin-daemonset",
    "nvidia.com/gpu.deploy.vgpu-device-manager": "nvidia-vgpu-device-manager",
}
COMPONENT_LABELS = list(COMPONENT_APP_LABELS)

PAUSED_STR = "paused-for-cc-mode-change"
CC_MODE_STATE_LABEL = "nvidia.com/cc.mode.state"
CC_READY_STATE_LABEL = "nvidia.com/cc.ready.state"
JOURNAL_ANNOTATION = "nvidia.com/cc-manager.paused-component-labels"

POD_POLL_SECONDS = 2.0


def _now() -> float:       # indirection points: tests swap these for a virtual clock
    return time.time()


def _pause(seconds: float) -> None:
    time.sleep(seconds)


_READY_FOR_STATE = {"on": "true", "ppcie": "true", "off": "false"}


def _maybe_set_paused(current_value: Optional[str]) -> str:
    """'true' -> paused marker; custom values get the marker appended; disabled
    ('' / None / 'false') and already-paused values are left alone."""
    if not current_value:
        return ""
    if current_value == "false":
        return "false"
    if current_value == "true":
        return PAUSED_STR
    if PAUSED_STR in current_value:
        return current_value
    return current_value + "_" + PAUSED_STR


def _maybe_set_unpaused(current_value: Optional[str]) -> str:
    """Inverse of _maybe_set_paused; identity on values that were never paused."""
    if current_value == "false":
        return "false"
    if current_value == PAUSED_STR:
        return "true"
    if current_value and PAUSED_STR in current_value:
        stripped = current_value.replace("_" + PAUSED_STR, "").replace(PAUSED_STR, "")
        return stripped.strip("_")
    return current_value or ""


def _patch_labels(v1, node_name: str, updates: Dict[str, str]) -> None:
    """read-modify-write of the node's label map, the way the reference does it
    (read_node, mutate metadata.labels, patch_node with the whole object)."""
    node = v1.read_node(node_name)
    if node.metadata.labels is None:
        node.metadata.labels = {}
    node.metadata.labels.update(updates)
    v1.patch_node(node_name, node)


def fetch_current_component_labels(v1, node_name: str) -> Dict[str, str]:
    """Current value ('' when absent) of each operator component deploy label."""
    logger.info("Fetching GPU operator component labels from node '%s'", node_name)
    try:
        labels = v1.read_node(node_name).metadata.labels or {}
    except ApiException as exc:
        logger.error("Failed to fetch node labels: %s", exc)
        raise
    found = {name: labels.get(name, "") for name in COMPONENT_LABELS}
    for name, value in found.items():
        logger.info("  %s=%s", name, value)
    return found


def _wait_for_pods_gone(v1, node_name: str, namespace: str, app_label: str, timeout: float,
                        clock: Callable[[], float], sleep: Callable[[float], None]) -> bool:
    started = clock()
    while clock() - started < timeout:
        try:
            pods = v1.list_namespaced_pod(namespace=namespace,
                                          field_selector=f"spec.nodeName={node_name}",
                                          label_selector=f"app={app_label}")
            if not pods.items:
                logger.info("  %s pods deleted", app_label)
                return True
            logger.debug("  Still waiting for %d %s pod(s)...", len(pods.items), app_label)
        except ApiException as exc:
            logger.warning("Error checking pod status: %s", exc)
        sleep(POD_POLL_SECONDS)
    logger.warning("Timeout waiting for %s pods to be deleted", app_label)
    return False


def evict_gpu_operator_components(v1, node_name: str, operator_namespace: str,
                                  current_labels: Dict[str, str], timeout: int = 300, *,
                                  concurrent_wait: bool = False,
                                  journal_annotation: bool = False,
                                  clock: Optional[Callable[[], float]] = None,
                                  sleep: Optional[Callable[[float], None]] = None) -> bool:
    """Pause every enabled component label, then wait until its pods left the node.

    A wait that times out is logged and ignored (as in the reference); only a k8s API
    failure while writing the labels makes this return False.
    """
    clock = clock or _now
    sleep = sleep or _pause
    logger.info("Evicting GPU operator components by setting deployment labels to 'paused'")
    try:
        paused = {name: _maybe_set_paused(value) for name, value in current_labels.items()}
        for name, value in paused.items():
            logger.debug("  %s: '%s' -> '%s'", name, current_labels[name], value)
        if journal_annotation:
            v1.patch_node(node_name, {"metadata": {"annotations": {
                JOURNAL_ANNOTATION: json.dumps(current_labels, sort_keys=True)}}})
        _patch_labels(v1, node_name, paused)
        logger.info("Successfully set deployment labels to 'paused' values")

        waits = [(name, COMPONENT_APP_LABELS[name]) for name, value in current_labels.items()
                 if value and name in COMPONENT_APP_LABELS]
        for _, app in waits:
            logger.info("Waiting for %s pods to be deleted...", app)
        if concurrent_wait and len(waits) > 1:
            with ThreadPoolExecutor(max_workers=len(waits), thread_name_prefix="cc-evict") as pool:
                list(pool.map(lambda w: _wait_for_pods_gone(v1, node_name, operator_namespace, w[1],
                                                            timeout, clock, sleep), waits))
        else:
            for _, app in waits:
                _wait_for_pods_gone(v1, node_name, operator_namespace, app, timeout, clock, sleep)
        logger.info("All GPU operator components evicted")
        return True
    except ApiException as exc:
        logger.error("Failed to evict GPU operator components: %s", exc)
        return False


def reschedule_gpu_operator_components(v1, node_name: str, original_labels: Dict[str, str], *,
                                       journal_annotation: bool = False) -> bool:
    """Write the un-paused form of `original_labels` back (identity on labels that
    were captured before pausing, which is how the manager calls it)."""
    logger.info("Rescheduling GPU operator components by restoring deployment labels")
    try:
        restored = {name: _maybe_set_unpaused(value) for name, value in original_labels.items()}
        for name, value in restored.items():
            logger.debug("  %s: '%s' -> '%s'", name, original_labels[name], value)
        _patch_labels(v1, node_name, restored)
        if journal_annotation:
            v1.patch_node(node_name, {"metadata": {"annotations": {JOURNAL_ANNOTATION: None}}})
        logger.info("Successfully restored deployment labels")
        return True
    except ApiException as exc:
        logger.error("Failed to reschedule GPU operator components: %s", exc)
        return False


def recover_journaled_labels(v1, node_name: str) -> Optional[Dict[str, str]]:
    """Original component labels journaled by an interrupted transition, or None."""
    try:
        annotations = getattr(v1.read_node(node_name).metadata, "annotations", None) or {}
    except ApiException as exc:
        logger.warning("Could not read node annotations: %s", exc)
        return None
    raw = annotations.get(JOURNAL_ANNOTATION)
    if not raw:
        return None
    try:
        data = json.loads(raw)
    except ValueError:
        logger.warning("Ignoring malformed %s annotation", JOURNAL_ANNOTATION)
        return None
    return {k: str(v) for k, v in data.items() if k in COMPONENT_APP_LABELS}


def set_cc_state_label(v1, node_name: str, state: str, *,
                       regate: Optional[Callable[[], bool]] = None) -> bool:
    """Publish the outcome: cc.mode.state=<state>, cc.ready.state = true for on/ppcie,
    false for off, empty for anything else (devtools, failed).

    regate (optional, not in the reference): called when the node read that precedes the patch shows
    cc.mode.state=failed, i.e. the LAST transition did not pass; it must return True for `state` to be
    published, otherwise 'failed' stays.  The check rides on the read the reference does anyway
    (read_node, mutate, patch_node), so the healthy path keeps exactly the reference's two API verbs;
    only after a (slow) regate is the node read again before it is patched."""
    try:
        node = v1.read_node(node_name)
        if regate is not None and (node.metadata.labels or {}).get(CC_MODE_STATE_LABEL) == "failed":
            if not regate():
                state = "failed"
            node = v1.read_node(node_name)          # the gate took a while: patch a fresh copy
        ready = _READY_FOR_STATE.get(state, "")
        if node.metadata.labels is None:
            node.metadata.labels = {}
        node.metadata.labels.update({CC_MODE_STATE_LABEL: state, CC_READY_STATE_LABEL: ready})
        v1.patch_node(node_name, node)
        logger.info("Set %s=%s, %s=%s", CC_MODE_STATE_LABEL, state, CC_READY_STATE_LABEL, ready)
        return True
    except ApiException as exc:
        logger.error("Failed to set cc.mode.state label: %s", exc)
        return False
