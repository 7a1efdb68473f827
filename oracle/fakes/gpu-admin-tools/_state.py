"""Pure-Python stand-in for NVIDIA/gpu-admin-tools (TEST / ORACLE INFRASTRUCTURE).

The real library (pinned v2025.11.21, reference versions.mk:22) is not vendored in
the reference tree and is not installed here.  This fake implements the device
object contract the reference manager relies on (reference main.py:38-40 imports;
call sites listed in SURVEY.md §8b) over an in-memory register file, so that the
UNMODIFIED reference main.py can be imported and driven by oracle/gen_golden.py.

Semantics (the only ones the reference can observe):
  set_*_mode stages a value; reset_with_os applies staged values; queries read the
  applied value.  Every op is appended to `world().trace` as (bdf, op, arg).
"""
from __future__ import annotations


class GpuError(Exception):
    pass


class World:
    def __init__(self):
        self.devices = []
        self.trace = []

    def reset(self):
        self.devices = []
        self.trace = []

    def add_gpu(self, bdf, name="NVIDIA B200 (sim)", cc="off", ppcie="off", cc_supported=True,
                ppcie_supported=True):
        d = FakeDevice(self, bdf, name, True, cc, ppcie, cc_supported, ppcie_supported)
        self.devices.append(d)
        return d

    def add_nvswitch(self, bdf, name="NVIDIA NVSwitch (sim)", ppcie="off", ppcie_supported=True):
        d = FakeDevice(self, bdf, name, False, "off", ppcie, False, ppcie_supported)
        self.devices.append(d)
        return d

    def trace_lines(self):
        return [f"{b} {op} {arg}" for b, op, arg in self.trace]


class FakeDevice:
    def __init__(self, world, bdf, name, is_gpu, cc, ppcie, cc_supported, ppcie_supported):
        self._world = world
        self.bdf = bdf
        self.name = name
        self._is_gpu = is_gpu
        self.is_cc_query_supported = cc_supported
        self.is_ppcie_query_supported = ppcie_supported
        self.cc_mode = self.cc_staged = cc
        self.ppcie_mode = self.ppcie_staged = ppcie
        self.fail = {}       # op name -> exception instance raised on every call
        self.stuck = False   # reset does not apply staged values

    def _op(self, op, arg="-"):
        self._world.trace.append((self.bdf, op, arg))
        exc = self.fail.get(op)
        if exc is not None:
            raise exc

    def is_gpu(self):
        return self._is_gpu

    def is_nvswitch(self):
        return not self._is_gpu

    def query_cc_mode(self):
        if "query_cc_mode" in self.fail:
            self._op("query_cc_mode", "error")
        self._op("query_cc_mode", self.cc_mode)
        return self.cc_mode

    def set_cc_mode(self, mode):
        self._op("set_cc_mode", mode)
        self.cc_staged = mode

    def query_ppcie_mode(self):
        if "query_ppcie_mode" in self.fail:
            self._op("query_ppcie_mode", "error")
        self._op("query_ppcie_mode", self.ppcie_mode)
        return self.ppcie_mode

    def set_ppcie_mode(self, mode):
        self._op("set_ppcie_mode", mode)
        self.ppcie_staged = mode

    def reset_with_os(self):
        self._op("reset_with_os")
        if self.stuck:
            self.cc_staged, self.ppcie_staged = self.cc_mode, self.ppcie_mode
        else:
            self.cc_mode, self.ppcie_mode = self.cc_staged, self.ppcie_staged

    def wait_for_boot(self):
        self._op("wait_for_boot")


_world = World()


def world() -> World:
    return _world
