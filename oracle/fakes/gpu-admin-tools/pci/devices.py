from _state import world


def find_gpus():
    """(devices, count) — every NVIDIA PCI function, GPUs and NVSwitches
    (reference main.py:144-155)."""
    devs = list(world().devices)
    return devs, len(devs)
