from .._cluster import ApiException  # noqa: F401
