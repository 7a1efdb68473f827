#!/usr/bin/env python3
"""Node transition (off->on->devtools->off, eviction-gated, registers/API simulated) under the
three context policies: keep the CUDA contexts (benchmark mode), release them after every gate
(daemon default), run the gate in a worker process.  No torch in this process."""
import json, logging, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests" / "fakes")]
import kubernetes
logging.disable(logging.CRITICAL)
from k8s_cc_manager_b200 import _native as N, devices as D, manager
from k8s_cc_manager_b200.drain_gate import COMPONENT_LABELS

L = N.lib(); assert L.ccm_init(N.BACKEND_CUDASIM) == 0
n = len([d for d in D.find_gpus()[0] if d.is_gpu()])
out = {"gpus": n}
for label, env in (("keep_context", {"CC_RELEASE_CUDA_CONTEXT": "false", "CC_SCRUB_ISOLATION": "thread"}),
                   ("release_context", {"CC_RELEASE_CUDA_CONTEXT": "true", "CC_SCRUB_ISOLATION": "thread"}),
                   ("worker_process", {"CC_RELEASE_CUDA_CONTEXT": "true", "CC_SCRUB_ISOLATION": "process"})):
    os.environ.update(env)
    os.environ["EVICT_OPERATOR_COMPONENTS"] = "true"
    L.ccm_sim_set(-1, b"cc_mode", 0)
    c = kubernetes.reset_cluster(); c.add_node("node", {k: "true" for k in COMPONENT_LABELS})
    mgr = manager.CCManager("node", "on", True)
    walls = []
    for mode in ("on", "devtools", "off", "on", "off"):
        t0 = time.perf_counter(); ok = mgr.set_cc_mode(mode); walls.append(round(time.perf_counter() - t0, 3))
        assert ok and c.labels("node")["nvidia.com/cc.mode.state"] == mode
    out[label] = walls
    print(label, walls, flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/transition_modes_n{n}.json").write_text(json.dumps(out, indent=1))
