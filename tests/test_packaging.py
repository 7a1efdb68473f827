"""N3 packaging: the image must build the SAME native library the tests exercise, from
the versions pinned in versions.mk, behind the reference's entrypoint and /app layout.  No docker in the
sandbox: these tests pin the contract between Dockerfile, Makefiles, versions.mk and build.py."""
from __future__ import annotations

import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DOCKERFILE = (ROOT / "deployments" / "container" / "Dockerfile.distroless").read_text()


def make_vars():
    out = subprocess.run(["make", "-s", "-f", "deployments/container/Makefile", "print-build-args"], cwd=ROOT,
                         capture_output=True, text=True, check=True).stdout
    return dict(kv.split("=", 1) for kv in out.split())


def test_makefile_passes_every_pinned_version_as_a_build_arg():
    mk = (ROOT / "deployments" / "container" / "Makefile").read_text()
    v = make_vars()
    assert v["GPU_ADMIN_TOOLS_VERSION"] == "v2025.11.21"           # same pin as the reference (versions.mk:22)
    assert tuple(int(x) for x in v["CUDA_VERSION"].split(".")[:2]) >= (12, 8), "sm_100a needs nvcc >= 12.8"
    assert v["IMAGE"].endswith(f":{v['VERSION']}-distroless")
    for arg in ("VERSION", "CUDA_VERSION", "GPU_ADMIN_TOOLS_VERSION", "RUNTIME_VERSION", "GIT_COMMIT"):
        assert f'--build-arg {arg}="$({arg})"' in mk, arg
        assert re.search(rf"^ARG {arg}\b", DOCKERFILE, re.M), f"Dockerfile ignores --build-arg {arg}"


def test_dockerfile_builds_the_library_with_build_py_flags():
    from k8s_cc_manager_b200 import build
    run = DOCKERFILE[DOCKERFILE.index("RUN nvcc"):DOCKERFILE.index("# Stage 3")]
    run = run.replace("\\\n", " ")
    flags = list(build.NVCC_FLAGS)
    i = 0
    while i < len(flags):                                             # every flag (and its value) appears verbatim
        tok = flags[i]
        if tok in ("-gencode", "-Xcompiler", "-cudart"):
            want = f"{tok} {flags[i + 1]}"
            if tok == "-Xcompiler":                                   # -Wall is a developer nicety, not part of the artefact
                want = want.replace(",-Wall", "")
            assert want in run, want
            i += 2
        else:
            assert tok in run.split(), tok
            i += 1
    for src in build.SOURCES + [build.CLI_SOURCE]:
        assert str(src.relative_to(ROOT)) in run, src
    assert "-o k8s_cc_manager_b200/libccm.so" in run and "-o k8s_cc_manager_b200/ccm-scrub" in run


def test_runtime_stage_keeps_the_reference_layout():
    final = DOCKERFILE[DOCKERFILE.rindex("FROM "):]
    assert "nvcr.io/nvidia/distroless/python:${RUNTIME_VERSION}" in final
    for needed in ("COPY main.py gpu_operator_eviction.py /app/", "/app/k8s_cc_manager_b200", "/app/gpu-admin-tools",
                   "/bin/rm", 'ENTRYPOINT ["python3", "/app/main.py"]', "WORKDIR /app"):
        assert needed in final, needed
    assert "CCM_ALLOW_SIM" not in re.sub(r"#.*", "", final), "the image must never opt in to simulated registers"


def test_top_level_makefile_targets_exist():
    out = subprocess.run(["make", "-n", "native", "oracle"], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "k8s_cc_manager_b200.build" in out.stdout


def test_the_dockerfile_build_stage_command_really_builds_the_library(tmp_path):
    """No docker here — but stage 2 of the image is one RUN line on top of the same CUDA toolkit this
    container has: execute exactly that line on a copy of what the stage COPYs, and check the artefacts."""
    import re
    import shutil
    import subprocess as sp
    stage = DOCKERFILE[DOCKERFILE.index("AS ccm"):DOCKERFILE.index("# Stage 3")]
    copies = re.findall(r"^COPY (\S+) (\S+)$", stage, re.M)
    assert copies == [("include/", "include/"), ("k8s_cc_manager_b200/", "k8s_cc_manager_b200/")]
    for src, dst in copies:
        shutil.copytree(ROOT / src, tmp_path / dst,
                        ignore=shutil.ignore_patterns("*.so", "ccm-scrub", "__pycache__"))
    run = stage[stage.index("RUN ") + 4:].replace("\\\n", " ").strip()
    proc = sp.run(["bash", "-c", run], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-3000:]
    lib, cli = tmp_path / "k8s_cc_manager_b200" / "libccm.so", tmp_path / "k8s_cc_manager_b200" / "ccm-scrub"
    assert lib.exists() and cli.exists()
    nm = sp.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ccm_[a-z0-9_]+)", nm))
    header = (ROOT / "include" / "ccm.h").read_text()
    declared = set(re.findall(r"^\s*(?:const char\*|int|uint64_t)\s+(ccm_[a-z0-9_]+)\s*\(", header, re.M))
    assert exported == declared
    sass = sp.run(["cuobjdump", "-lelf", str(lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass and "sm_90" not in sass, "the image must carry sm_100a code only"
    # the CLI finds the library next to itself ($ORIGIN rpath) and fails closed without a GPU
    out = sp.run([str(cli), "--all", "--backend", "sim"], capture_output=True, text=True,
                 env={"CCM_SIM_GPUS": "2", "CCM_SIM_BIND_CUDA": "0", "PATH": "/usr/bin:/bin"}, timeout=60)
    assert out.returncode == 3 and '"status": -9' in out.stdout
