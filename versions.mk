# Versions the image build pins (counterpart of the reference's versions.mk:19-23; Go toolchain
# dropped with the legacy Go manager, CUDA moved to the first toolkit that knows sm_100a).
VERSION ?= v1.0.0
vVERSION := v$(VERSION:v%=%)

# nvcc >= 12.8 is required for -gencode arch=compute_100a,code=sm_100a (B200)
CUDA_VERSION := 12.9.1
# register access on real CC nodes (CC_DEVICE_LIBRARY=gpu-admin-tools); same pin as the reference
GPU_ADMIN_TOOLS_VERSION := v2025.11.21
# nvcr.io/nvidia/distroless/python tag of the runtime stage
RUNTIME_VERSION := 3.13-v4.0.4
