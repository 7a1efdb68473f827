// ccm_internal.h — glue between the register backends (ccm_core.cpp) and the CUDA
// scrub engine (ccm_scrub.cu).  Nothing here crosses the public ABI (include/ccm.h).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "ccm.h"

namespace ccm {

// thread-local detailed error text behind ccm_last_error()
void set_error(const char* fmt, ...);
const std::string& last_error();

// ---- implemented in ccm_scrub.cu ------------------------------------------
int cuda_device_count();  // 0 when no driver / no device (never throws)
int cuda_describe(int ordinal, char* bdf, size_t bdf_cap, char* name, size_t name_cap,
                  uint64_t* total_bytes);

struct ScrubEngine;  // one per CUDA ordinal, created lazily
ScrubEngine* engine_for(int ordinal);  // nullptr + error text if unusable
ScrubEngine* engine_lookup(int ordinal);  // the engine if it exists; never creates a context

int engine_arena_acquire(ScrubEngine*, uint64_t bytes, ccm_arena_info* out);
int engine_arena_release(ScrubEngine*, double* ms);
int engine_arena_scrub(ScrubEngine*, int variant, const ccm_launch_cfg*, void* stream, float* ms);
int engine_arena_verify(ScrubEngine*, int variant, const ccm_launch_cfg*, void* stream,
                        uint64_t* nonzero, float* ms);
int engine_arena_scrub_verify_async(ScrubEngine*, int sv, int vv, const ccm_launch_cfg*,
                                    const ccm_launch_cfg*, void* stream);
int engine_arena_fetch_count(ScrubEngine*, void* stream, uint64_t* nonzero);
int engine_arena_step_times(ScrubEngine*, int cap, float* scrub_ms, float* verify_ms, int* n);
int engine_arena_fill(ScrubEngine*, int byte_value, void* stream);
int engine_arena_fill_random(ScrubEngine*, uint64_t seed, void* stream);
int engine_arena_rw(ScrubEngine*, uint64_t offset, void* host, uint64_t bytes, bool write);
// inject: fault drill (sim key "scrub_inject") — that many bytes are poisoned AFTER the scrub and
// BEFORE the read-back (first / unaligned / middle / last byte of each chunk), so the verdict must be DIRTY.
// node_fanout: the call is one of several running concurrently in THIS process (ccm_scrub_verify_many).
int engine_scrub_verify(ScrubEngine*, uint64_t bytes, uint64_t inject, bool node_fanout, ccm_scrub_result* out);
// Joins the background release of the last product call (no-op when none is pending).
int engine_release_wait(ScrubEngine*, double* ms_release, double* ms_waited);
int engine_region_scrub(ScrubEngine*, void* dptr, uint64_t bytes, int variant,
                        const ccm_launch_cfg*, void* stream, float* ms);
int engine_region_verify(ScrubEngine*, const void* dptr, uint64_t bytes, int variant,
                         const ccm_launch_cfg*, void* stream, uint64_t* nonzero, float* ms);
int engine_host_roundtrip(ScrubEngine*, void* host_buf, uint64_t bytes, uint64_t dev_offset,
                          int sv, int vv, uint64_t* pre, uint64_t* post);
uint64_t kernel_launches();
const char* default_kernel_names();
// Tears down the engine of a CUDA ordinal (if any) and resets its primary context.
int engine_teardown(int ordinal);

// ---- implemented in ccm_core.cpp -------------------------------------------
// Maps a device-table index to its CUDA ordinal (or a negative ccm_status).
int cuda_ordinal_of(int dev);
// sim fault hook for the scrub op (returns CCM_OK or CCM_ERR_FAULT) + trace line; *inject = the
// device's "scrub_inject" drill value.
int sim_scrub_hook(int dev, uint64_t* inject);

}  // namespace ccm
