"""State of the fake API server."""
from __future__ import annotations

import copy
import threading
import time as _time
from types import SimpleNamespace


class ApiException(Exception):
    def __init__(self, status=0, reason=None, http_resp=None):
        super().__init__(f"({status})\nReason: {reason}\n")
        self.status = status
        self.reason = reason
        self.body = None
        self.headers = None


class VirtualClock:
    """time.time()/time.sleep() replacement: sleeping advances the clock at once."""

    def __init__(self, start=1_000_000.0):
        self.now = start
        self.sleeps = []
        self._lock = threading.Lock()

    def time(self):
        with self._lock:
            return self.now

    def sleep(self, seconds):
        with self._lock:
            self.sleeps.append(seconds)
            self.now += max(0.0, seconds)

    def monotonic(self):
        return self.time()

    def perf_counter(self):
        return self.time()


def make_node(name, labels=None, resource_version="1", annotations=None):
    meta = SimpleNamespace(name=name, labels=dict(labels) if labels is not None else None,
                           resource_version=resource_version,
                           annotations=dict(annotations) if annotations else None)
    return SimpleNamespace(metadata=meta, kind="Node")


class FakeCluster:
    def __init__(self):
        self.lock = threading.RLock()
        self.nodes = {}
        self.pods = []            # dicts: name, namespace, node, app, gone_at (virtual time or None)
        self.clock = VirtualClock()
        self.calls = []           # (verb, args) in order
        self.fail = {}            # verb -> list of ApiException|None consumed per call
        self.watch_script = []    # list of batches; each batch is a list of events or an Exception
        self.watch_calls = []     # kwargs of each stream() call
        self.rv = 1
        self.incluster_ok = True
        self.kubeconfig_ok = True
        self.loaded = []
        self.on_patch = None      # callback(cluster, node_name, labels) after each patch

    # -- scripting helpers -------------------------------------------------
    def add_node(self, name, labels=None):
        with self.lock:
            self.rv += 1
            self.nodes[name] = make_node(name, labels, str(self.rv))
            return self.nodes[name]

    def add_pod(self, app, node, namespace="gpu-operator", gone_after=None, name=None):
        with self.lock:
            self.pods.append({
                "name": name or f"{app}-{len(self.pods)}", "namespace": namespace, "node": node,
                "app": app, "gone_after": gone_after, "gone_at": None, "paused_seen": False,
            })

    def fail_next(self, verb, status, times=1):
        self.fail.setdefault(verb, []).extend([ApiException(status=status, reason="injected")] * times)

    def labels(self, name):
        return dict(self.nodes[name].metadata.labels or {})

    def verbs(self):
        return [c[0] for c in self.calls]

    def _record(self, verb, *args):
        self.calls.append((verb, args))
        q = self.fail.get(verb)
        if q:
            exc = q.pop(0)
            if exc is not None:
                raise exc

    # -- API surface ---------------------------------------------------------
    def read_node(self, name):
        with self.lock:
            self._record("read_node", name)
            if name not in self.nodes:
                raise ApiException(status=404, reason="Not Found")
            return copy.deepcopy(self.nodes[name])

    def patch_node(self, name, body):
        with self.lock:
            self._record("patch_node", name, copy.deepcopy(getattr(body.metadata, "labels", None))
                         if hasattr(body, "metadata") else copy.deepcopy(body))
            if name not in self.nodes:
                raise ApiException(status=404, reason="Not Found")
            node = self.nodes[name]
            if hasattr(body, "metadata"):
                new_labels = body.metadata.labels or {}
                new_ann = getattr(body.metadata, "annotations", None)
            else:  # dict-style strategic merge patch
                md = body.get("metadata", {})
                new_labels = md.get("labels") or {}
                new_ann = md.get("annotations")
            if node.metadata.labels is None:
                node.metadata.labels = {}
            for k, v in new_labels.items():
                if v is None:
                    node.metadata.labels.pop(k, None)
                else:
                    node.metadata.labels[k] = v
            if new_ann:
                if node.metadata.annotations is None:
                    node.metadata.annotations = {}
                for k, v in new_ann.items():
                    if v is None:
                        node.metadata.annotations.pop(k, None)
                    else:
                        node.metadata.annotations[k] = v
            self.rv += 1
            node.metadata.resource_version = str(self.rv)
            # pods of paused components start terminating now
            now = self.clock.time()
            for p in self.pods:
                if p["node"] == name and p["gone_at"] is None and p["gone_after"] is not None:
                    p["gone_at"] = now + p["gone_after"]
            if self.on_patch:
                self.on_patch(self, name, dict(node.metadata.labels))
            return copy.deepcopy(node)

    def list_namespaced_pod(self, namespace, field_selector=None, label_selector=None, **_):
        with self.lock:
            self._record("list_namespaced_pod", namespace, field_selector, label_selector)
            node = None
            if field_selector and field_selector.startswith("spec.nodeName="):
                node = field_selector.split("=", 1)[1]
            app = None
            if label_selector and label_selector.startswith("app="):
                app = label_selector.split("=", 1)[1]
            now = self.clock.time()
            items = []
            for p in self.pods:
                if p["namespace"] != namespace:
                    continue
                if node is not None and p["node"] != node:
                    continue
                if app is not None and p["app"] != app:
                    continue
                if p["gone_at"] is not None and now >= p["gone_at"]:
                    continue
                items.append(SimpleNamespace(metadata=SimpleNamespace(name=p["name"], labels={"app": p["app"]})))
            return SimpleNamespace(items=items)

    def list_node(self, **kwargs):  # only ever used as the watch target
        with self.lock:
            self._record("list_node", kwargs)
            return SimpleNamespace(items=[copy.deepcopy(n) for n in self.nodes.values()])


_cluster = FakeCluster()


def cluster() -> FakeCluster:
    return _cluster


def reset_cluster() -> FakeCluster:
    global _cluster
    _cluster = FakeCluster()
    return _cluster
