"""In-memory stand-in for the `kubernetes` Python client (TEST INFRASTRUCTURE).

The real package is not installed in this image and there is no network.  This
stub exposes exactly the surface the reference touches (reference main.py:33-34,
130-140,589,632; gpu_operator_eviction.py:15-16,115,165,170,189-193):

    client.CoreV1Api().read_node / patch_node / list_namespaced_pod / list_node
    client.rest.ApiException(status=...)
    config.load_incluster_config / load_kube_config / ConfigException
    watch.Watch().stream(func, **kwargs)

All state lives in one process-wide `FakeCluster` (kubernetes.cluster()), which
tests script: nodes with labels, pods that disappear after a delay on a virtual
clock, queued watch events and injected API errors.
"""
from . import client, config, watch  # noqa: F401
from ._cluster import FakeCluster, cluster, reset_cluster  # noqa: F401
