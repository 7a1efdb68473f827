"""The C-ABI library loads and exports every symbol include/ccm.h declares; struct
layouts seen through ctypes match the header; the scrub fails LOUDLY without CUDA
(no compute call is made here: this file runs in the CPU-only suite)."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "ccm.h").read_text()


def declared_functions():
    return sorted(set(re.findall(r"^\s*(?:const char\*|int|uint64_t)\s+(ccm_[a-z0-9_]+)\s*\(", HEADER, re.M)))


def test_header_declares_the_documented_surface():
    names = declared_functions()
    assert len(names) >= 30
    for must in ("ccm_enumerate", "ccm_query_cc_mode", "ccm_set_cc_mode", "ccm_reset", "ccm_wait_for_boot",
                 "ccm_scrub_verify", "ccm_scrub_verify_many", "ccm_scrub_release_wait", "ccm_device_release_many",
                 "ccm_strerror"):
        assert must in names


def test_library_exports_every_declared_symbol(native):
    lib = native.lib()
    nm = subprocess.run(["nm", "-D", "--defined-only", str(native.LIB_PATH)], capture_output=True, text=True,
                        check=True).stdout
    exported = set(re.findall(r" T (ccm_[a-z0-9_]+)", nm))
    declared = set(declared_functions())
    assert declared <= exported, declared - exported
    assert exported <= declared, f"exported but undeclared: {exported - declared}"
    assert set(native.EXPORTED_SYMBOLS) == declared
    assert lib.ccm_abi_version() == native.ABI_VERSION == 2


def test_every_entry_point_cites_the_reference_or_says_new():
    """include/ccm.h must tie each device op to the reference call site it replaces."""
    for fn in ("ccm_query_cc_mode", "ccm_set_cc_mode", "ccm_query_ppcie_mode", "ccm_set_ppcie_mode", "ccm_reset",
               "ccm_wait_for_boot", "ccm_enumerate"):
        idx = re.search(rf"^int {fn}\(", HEADER, re.M).start()
        assert "main.py:" in HEADER[max(0, idx - 700):idx], fn


def test_struct_layouts_match_header(native, tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ccm.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ccm_dev_info), sizeof(ccm_launch_cfg),'
                   'sizeof(ccm_scrub_result), sizeof(ccm_arena_info), offsetof(ccm_dev_info,bdf),'
                   'offsetof(ccm_scrub_result,ms_acquire), offsetof(ccm_scrub_result,status),'
                   'offsetof(ccm_scrub_result,device_free_before), offsetof(ccm_scrub_result,ms_gpu_span));return 0;}')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    N = native
    assert got == [C.sizeof(N.DevInfo), C.sizeof(N.LaunchCfg), C.sizeof(N.ScrubResult), C.sizeof(N.ArenaInfo),
                   N.DevInfo.bdf.offset, N.ScrubResult.ms_acquire.offset, N.ScrubResult.status.offset,
                   N.ScrubResult.device_free_before.offset, N.ScrubResult.ms_gpu_span.offset]


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c89ish.c"
    src.write_text('#include "ccm.h"\nint main(void){return CCM_ABI_VERSION - 1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"), "-c", str(src),
                    "-o", str(tmp_path / "o.o")], check=True)


def test_strerror_covers_all_codes(native):
    lib = native.lib()
    for code in range(0, -13, -1):
        assert lib.ccm_strerror(code).decode() != "unknown ccm status"
    assert lib.ccm_strerror(-99).decode() == "unknown ccm status"


def test_scrub_fails_loudly_without_cuda(native):
    """No host fallback: a device without a CUDA ordinal cannot be scrubbed."""
    lib = native.lib()
    assert lib.ccm_sim_topology(2, 0) == 0
    assert lib.ccm_sim_set(-1, b"cuda_ordinal", -1) == 0
    res = native.ScrubResult()
    rc = lib.ccm_scrub_verify(0, 1 << 20, C.byref(res))
    assert rc == native.ERR_NO_CUDA and res.status == native.ERR_NO_CUDA and res.bytes_scrubbed == 0
    assert "no CUDA device" in native.last_error()
    ai = native.ArenaInfo()
    assert lib.ccm_arena_acquire(0, 1 << 20, C.byref(ai)) == native.ERR_NO_CUDA


def test_missing_library_is_an_import_error(monkeypatch, tmp_path):
    import importlib
    from k8s_cc_manager_b200 import _native
    monkeypatch.setenv("CCM_LIB", str(tmp_path / "nope.so"))
    fresh = importlib.reload(_native)
    try:
        with pytest.raises(ImportError):
            fresh.lib()
    finally:
        monkeypatch.delenv("CCM_LIB")
        importlib.reload(_native)


def test_traffic_json_is_keyed_by_the_kernels_auto_launches(native):
    """bench.py copies roofline.traffic from profiles/traffic.json (an ncu capture, never measured under
    the bench): the capture must be of the kernels the library launches TODAY, or the field goes stale
    silently."""
    import json
    scrub, verify = native.default_kernels()
    assert scrub.startswith("scrub_st256_fast_kernel<") and verify.startswith("verify_ld256_fast_kernel<")
    tr = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    assert set(tr["kernels"]) == {scrub, verify}
    for k in tr["kernels"].values():
        assert 0.98 * tr["region_bytes"] < k["dram_bytes_read"] + k["dram_bytes_write"] < 1.02 * tr["region_bytes"]
    src = (ROOT / "k8s_cc_manager_b200" / "csrc" / "ccm_scrub.cu").read_text()
    assert "launch_fast(CCM_FAST_SCRUB" in src and "launch_fast(CCM_FAST_VERIFY" in src


def test_library_asks_for_one_cuda_connection_unless_the_host_chose(native):
    """libccm drives ONE stream per GPU; a context with 1 hardware queue is created and reset in half the time
    of the default 8 (profiles/r2_ctx_connections.log).  The host's own choice is never overridden."""
    import subprocess, sys
    code = ("import ctypes as C, os, sys; sys.path.insert(0, %r); from k8s_cc_manager_b200 import _native as N; "
            "n = C.c_int(); N.lib().ccm_enumerate(None, 0, C.byref(n)); "
            "g = C.CDLL(None).getenv; g.restype = C.c_char_p; print(g(b'CUDA_DEVICE_MAX_CONNECTIONS'))" % str(ROOT))
    base = {k: v for k, v in os.environ.items() if k not in ("CUDA_DEVICE_MAX_CONNECTIONS", "CCM_CUDA_MAX_CONNECTIONS")}
    run = lambda env: subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, check=True).stdout.strip()  # noqa: E731
    assert run(base) == "b'1'"
    assert run(dict(base, CUDA_DEVICE_MAX_CONNECTIONS="4")) == "b'4'"            # the host's choice stands
    assert run(dict(base, CCM_CUDA_MAX_CONNECTIONS="0")) == "None"               # opt out: CUDA's default
    assert run(dict(base, CCM_CUDA_MAX_CONNECTIONS="2")) == "b'2'"
