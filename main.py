#!/usr/bin/env python3
"""Image entrypoint — same path and surface as the reference (`ENTRYPOINT ["python3",
"/app/main.py"]`, reference deployments/container/Dockerfile.distroless:70): same CLI
flags, env vars, node labels, readiness file and exit codes.  The implementation lives
in k8s_cc_manager_b200/manager.py."""
from k8s_cc_manager_b200.manager import (  # noqa: F401
    CC_MODE_CONFIG_LABEL,
    READINESS_FILE,
    CCManager,
    create_readiness_file,
    is_host_cc_enabled,
    main,
)

if __name__ == "__main__":
    main()
