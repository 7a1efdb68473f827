from _state import FakeDevice as Gpu  # noqa: F401  (reference main.py:38: imported, unused)
