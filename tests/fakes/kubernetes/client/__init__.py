from . import rest  # noqa: F401
from .rest import ApiException  # noqa: F401


class CoreV1Api:
    """Bound to the process-wide FakeCluster at call time (so reset_cluster() works)."""

    def __init__(self, api_client=None):
        self.api_client = api_client

    def _c(self):
        from .._cluster import cluster
        return cluster()

    def read_node(self, name, **kw):
        return self._c().read_node(name)

    def patch_node(self, name, body, **kw):
        return self._c().patch_node(name, body)

    def list_namespaced_pod(self, namespace, **kw):
        return self._c().list_namespaced_pod(namespace, **kw)

    def list_node(self, **kw):
        return self._c().list_node(**kw)
