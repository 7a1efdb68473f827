// Host-only stand-in for csrc/ccm_scrub.cu so that csrc/ccm_core.cpp (device table, register
// backends, concurrent launchers) can be built with plain g++ under -fsanitize=thread / address.
// TEST INFRASTRUCTURE: every scrub entry point reports "no CUDA" exactly like a GPU-less box.
#include "ccm_internal.h"

namespace ccm {
int cuda_device_count() { return 0; }
int cuda_describe(int, char*, size_t, char*, size_t, uint64_t*) { return CCM_ERR_NO_CUDA; }
ScrubEngine* engine_for(int) { set_error("stub build: no CUDA"); return nullptr; }
ScrubEngine* engine_lookup(int) { return nullptr; }
int engine_arena_acquire(ScrubEngine*, uint64_t, ccm_arena_info*) { return CCM_ERR_NO_CUDA; }
int engine_arena_release(ScrubEngine*, double*) { return CCM_ERR_NO_CUDA; }
int engine_arena_scrub(ScrubEngine*, int, const ccm_launch_cfg*, void*, float*) { return CCM_ERR_NO_CUDA; }
int engine_arena_verify(ScrubEngine*, int, const ccm_launch_cfg*, void*, uint64_t*, float*) { return CCM_ERR_NO_CUDA; }
int engine_arena_scrub_verify_async(ScrubEngine*, int, int, const ccm_launch_cfg*, const ccm_launch_cfg*, void*) { return CCM_ERR_NO_CUDA; }
int engine_arena_fetch_count(ScrubEngine*, void*, uint64_t*) { return CCM_ERR_NO_CUDA; }
int engine_arena_step_times(ScrubEngine*, int, float*, float*, int*) { return CCM_ERR_NO_CUDA; }
int engine_arena_fill(ScrubEngine*, int, void*) { return CCM_ERR_NO_CUDA; }
int engine_arena_fill_random(ScrubEngine*, uint64_t, void*) { return CCM_ERR_NO_CUDA; }
int engine_arena_rw(ScrubEngine*, uint64_t, void*, uint64_t, bool) { return CCM_ERR_NO_CUDA; }
int engine_scrub_verify(ScrubEngine*, uint64_t, uint64_t, bool, ccm_scrub_result*) { return CCM_ERR_NO_CUDA; }
int engine_release_wait(ScrubEngine*, double*, double*) { return CCM_ERR_NO_CUDA; }
int engine_region_scrub(ScrubEngine*, void*, uint64_t, int, const ccm_launch_cfg*, void*, float*) { return CCM_ERR_NO_CUDA; }
int engine_region_verify(ScrubEngine*, const void*, uint64_t, int, const ccm_launch_cfg*, void*, uint64_t*, float*) { return CCM_ERR_NO_CUDA; }
int engine_host_roundtrip(ScrubEngine*, void*, uint64_t, uint64_t, int, int, uint64_t*, uint64_t*) { return CCM_ERR_NO_CUDA; }
uint64_t kernel_launches() { return 0; }
const char* default_kernel_names() { return ""; }
int engine_teardown(int) { return CCM_OK; }
}  // namespace ccm
