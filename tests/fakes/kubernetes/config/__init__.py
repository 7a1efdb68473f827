class ConfigException(Exception):
    pass


def load_incluster_config():
    from .._cluster import cluster
    c = cluster()
    if not c.incluster_ok:
        raise ConfigException("Service host/port is not set.")
    c.loaded.append("incluster")


def load_kube_config(config_file=None, **kw):
    from .._cluster import cluster
    c = cluster()
    if not c.kubeconfig_ok:
        raise ConfigException("Invalid kube-config file. No configuration found.")
    c.loaded.append(("kubeconfig", config_file))
