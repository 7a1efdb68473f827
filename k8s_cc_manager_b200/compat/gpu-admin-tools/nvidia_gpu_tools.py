#!/usr/bin/env python3
"""Drop-in for the gpu-admin-tools module the reference imports (`from nvidia_gpu_tools
import Gpu`, reference main.py:38) and for the CLI its legacy shell engine calls
(reference scripts/cc-manager.sh:389,437):

    nvidia_gpu_tools.py --query-cc-mode --gpu-bdf=<bdf>          -> prints "CC mode is <mode>"
    nvidia_gpu_tools.py --set-cc-mode=<mode> --reset-after-cc-mode-switch --gpu-bdf=<bdf>

Placing this directory next to the UNMODIFIED reference main.py as `gpu-admin-tools/`
(main.py:30-31 puts exactly that path first on sys.path) makes the reference run on
libccm.so.  New flag: --scrub-after-cc-mode-switch runs the HBM scrub gate as well.
"""
from __future__ import annotations

import argparse
import logging
import sys

from k8s_cc_manager_b200.devices import Gpu, GpuError, find_gpus  # noqa: F401

logger = logging.getLogger("nvidia_gpu_tools")


def _find(bdf: str):
    for dev in find_gpus()[0]:
        if dev.bdf == bdf.lower():
            return dev
    raise GpuError(f"no NVIDIA device with BDF {bdf}")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="gpu-admin-tools compatible CLI over libccm.so")
    ap.add_argument("--gpu-bdf", required=True)
    ap.add_argument("--query-cc-mode", action="store_true")
    ap.add_argument("--set-cc-mode", choices=("on", "off", "devtools"))
    ap.add_argument("--reset-after-cc-mode-switch", action="store_true")
    ap.add_argument("--scrub-after-cc-mode-switch", action="store_true")
    args = ap.parse_args(argv)
    try:
        gpu = _find(args.gpu_bdf)
        if args.set_cc_mode:
            gpu.set_cc_mode(args.set_cc_mode)
            print(f"GPU {gpu.bdf} CC mode staged to {args.set_cc_mode}")
            if args.reset_after_cc_mode_switch:
                gpu.reset_with_os()
                gpu.wait_for_boot()
                print(f"GPU {gpu.bdf} reset")
                if args.scrub_after_cc_mode_switch:
                    rep = gpu.scrub_and_verify()
                    print(f"GPU {gpu.bdf} scrubbed {rep.bytes_scrubbed} bytes, {rep.nonzero_bytes} non-zero")
        if args.query_cc_mode or args.set_cc_mode:
            print(f"GPU {gpu.bdf} CC mode is {gpu.query_cc_mode()}")
        return 0
    except GpuError as exc:
        print(f"error: {exc}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
