/*
 * ccm.h — C ABI of libccm.so, the device shim of the B200-native CC-mode manager.
 *
 * This is the drop-in boundary for the ONE hot path: the per-GPU
 * stage -> reset -> wait-for-boot -> verify-mode sequence of the reference
 * (reference main.py:502-529), plus the NEW full-HBM scrub-and-verify stage that
 * gates a GPU's release after every CC transition (insertion point: after
 * reference main.py:529, before the state label is written at main.py:541).
 *
 * The reference has no C ABI at this boundary: it calls a duck-typed Python device
 * object from the external NVIDIA/gpu-admin-tools library (pinned v2025.11.21,
 * reference versions.mk:22; not vendored under /root/reference). Each entry point
 * below names the reference call site it replaces. k8s_cc_manager_b200/devices.py
 * rebuilds the duck-typed object on top of this ABI, so a reference-shaped manager
 * runs against it unchanged (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *  - plain C types only: ints, pointers, sizes.  No torch / C++ types cross the ABI.
 *  - every function returns CCM_OK (0) or a negative ccm_status; nothing throws.
 *  - the caller owns every out-struct / buffer; the library owns nothing the caller
 *    must free, except the scrub arena, which is released with ccm_arena_release().
 *  - thread safety: all per-device entry points may be called concurrently for
 *    DIFFERENT device indices (no global lock is held across reset / wait / scrub).
 *    Calls for the SAME device index are serialised by a per-device mutex.
 *  - `dev` is an index into the table returned by ccm_enumerate().
 *  - `stream` arguments are CUDA stream handles (cudaStream_t / CUstream) passed as
 *    void*; NULL means the library's own per-device stream.  Launches on ONE device
 *    must be ordered with respect to each other (same stream, or event-ordered
 *    streams): they share that device's result counter and tile-grab counter.
 */
#ifndef CCM_H_
#define CCM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built -fvisibility=hidden */
#endif

#define CCM_ABI_VERSION 2 /* v2: deferred release, coverage fields; ccm_transition_many and the
                             TMA verify variant removed (one phase engine, one verify family) */

/* ---- status codes ------------------------------------------------------- */
typedef enum ccm_status {
  CCM_OK = 0,
  CCM_ERR_INVALID = -1,     /* bad argument                                    */
  CCM_ERR_NO_DEVICE = -2,   /* device index out of range                       */
  CCM_ERR_UNSUPPORTED = -3, /* op not supported by this device / backend       */
  CCM_ERR_IO = -4,          /* register / sysfs access failed                  */
  CCM_ERR_TIMEOUT = -5,     /* wait_for_boot timed out                         */
  CCM_ERR_CUDA = -6,        /* a CUDA runtime call failed (see ccm_last_error) */
  CCM_ERR_NOMEM = -7,       /* could not obtain the requested scrub arena      */
  CCM_ERR_DIRTY = -8,       /* verify found non-zero bytes after the scrub     */
  CCM_ERR_NO_CUDA = -9,     /* no usable CUDA device behind this entry         */
  CCM_ERR_STATE = -10,      /* call out of order (e.g. no arena held)          */
  CCM_ERR_FAULT = -11,      /* fault injected by the sim backend               */
  CCM_ERR_NOT_BOOTED = -12  /* device was reset and has not been waited for    */
} ccm_status;

/* ---- modes (string values of the reference: 'off' 'on' 'devtools') ------ */
typedef enum ccm_cc_mode { CCM_CC_OFF = 0, CCM_CC_ON = 1, CCM_CC_DEVTOOLS = 2 } ccm_cc_mode;
typedef enum ccm_ppcie_mode { CCM_PPCIE_OFF = 0, CCM_PPCIE_ON = 1 } ccm_ppcie_mode;

typedef enum ccm_dev_kind { CCM_KIND_GPU = 0, CCM_KIND_NVSWITCH = 1 } ccm_dev_kind;

/* Register backends.  The gpurun box cannot unbind / reset its GPUs, so the
 * CC-mode register file is simulated there; see DESIGN.md "what is simulated". */
typedef enum ccm_backend {
  CCM_BACKEND_SIM = 0,     /* synthetic topology, simulated registers           */
  CCM_BACKEND_CUDASIM = 1, /* real CUDA devices (BDF/name/HBM), sim registers   */
  CCM_BACKEND_SYSFS = 2    /* real PCI enumeration via sysfs, NVML CC query;
                              set/stage are UNSUPPORTED (register map lives in
                              gpu-admin-tools, not in the reference tree)       */
} ccm_backend;

/* One row per NVIDIA PCI function.
 * Replaces: pci.devices.find_gpus() -> (devices, count), reference main.py:155;
 * fields mirror what the reference reads from each device object:
 *   .bdf / .name                      main.py:187,191,210,239,280
 *   .is_gpu() / .is_nvswitch()        main.py:165,175
 *   .is_cc_query_supported            main.py:186
 *   .is_ppcie_query_supported         main.py:205,477                         */
typedef struct ccm_dev_info {
  int32_t index;               /* index to pass as `dev`                        */
  int32_t kind;                /* ccm_dev_kind                                  */
  int32_t cc_query_supported;  /* bool                                          */
  int32_t ppcie_query_supported; /* bool                                        */
  int32_t cuda_ordinal;        /* CUDA device ordinal backing the scrub, or -1  */
  int32_t reserved0;
  uint64_t hbm_total_bytes;    /* cudaMemGetInfo total, 0 when no CUDA device   */
  char bdf[32];                /* "0000:1b:00.0"                                */
  char name[96];               /* "NVIDIA B200"                                 */
} ccm_dev_info;

/* ---- library / backend -------------------------------------------------- */
int ccm_abi_version(void);
const char* ccm_strerror(int status);
/* Copies this thread's last detailed error message; returns its length. */
int ccm_last_error(char* buf, size_t cap);

/* Selects the backend and (re)builds the device table.  Called implicitly with
 * the environment's choice (CCM_BACKEND = sim | cudasim | sysfs; default cudasim
 * when a CUDA device is visible, else sim) on first use. */
int ccm_init(int backend);
int ccm_backend_in_use(void);

/* find_gpus(): fills up to `cap` rows, writes the total count to *n.
 * Re-enumeration is cheap and keeps register state (reference calls find_gpus()
 * repeatedly: main.py:164,174,203,275,473). */
int ccm_enumerate(ccm_dev_info* out, int cap, int* n);

/* ---- register-level device ops (reference boundary) --------- */
/* dev.query_cc_mode()      reference main.py:441,505,524 */
int ccm_query_cc_mode(int dev, int* mode);
/* dev.set_cc_mode(mode)    reference main.py:511 — STAGES only; applied by reset */
int ccm_set_cc_mode(int dev, int mode);
/* dev.query_ppcie_mode()   reference main.py:310,340,353,373,479,495 */
int ccm_query_ppcie_mode(int dev, int* mode);
/* dev.set_ppcie_mode(m)    reference main.py:345,359,482 — stages only */
int ccm_set_ppcie_mode(int dev, int mode);
/* dev.reset_with_os()      reference main.py:346,368,490,519 */
int ccm_reset(int dev);
/* dev.wait_for_boot()      reference main.py:347,372,494,523; timeout_ms<=0 = default */
int ccm_wait_for_boot(int dev, int timeout_ms);

/* ---- HBM scrub-and-verify (NEW stage) ------------------ */
typedef enum ccm_scrub_variant {
  CCM_SCRUB_AUTO = 0,
  CCM_SCRUB_ST128 = 1,   /* st.global.v4.b32, persistent CTAs                   */
  CCM_SCRUB_ST256 = 2,   /* st.global.v8.b32 (256-bit, sm_100+)                 */
  CCM_SCRUB_TMA = 3,     /* cp.async.bulk.global.shared::cta from a zero tile   */
  CCM_SCRUB_MEMSET = 4   /* cudaMemsetAsync — library bar, baseline only        */
} ccm_scrub_variant;

typedef enum ccm_verify_variant {
  CCM_VERIFY_AUTO = 0,
  CCM_VERIFY_LD128 = 1,  /* ld.global.nc.v4.b32                                 */
  CCM_VERIFY_LD256 = 2   /* ld.global.nc.v8.b32 (256-bit)                       */
} ccm_verify_variant;

/* Launch shape override; zeros mean "library default for this variant". */
typedef struct ccm_launch_cfg {
  int32_t ctas_per_sm;
  int32_t threads_per_cta;
  int32_t tile_bytes;   /* TMA variants: bytes per bulk op (multiple of 16);
                           ST/LD variants with a dynamic schedule: bytes per grab */
  int32_t unroll;       /* ST/LD variants: vectors in flight per thread (2,4,8) */
  int32_t cache_policy; /* 0 library default, 1 plain, 2 L2::evict_first,
                           3 streaming (.cs / L1::no_allocate), 4 L2::evict_last */
  int32_t schedule;     /* 0 library default, 1 static grid-stride, 2 dynamic
                           (persistent CTAs grab chunks off an atomic counter),
                           3 dynamic per WARP (ST/LD variants; no block barrier) */
} ccm_launch_cfg;

typedef struct ccm_scrub_result {
  uint64_t bytes_requested;    /* as passed (0 = max)                           */
  uint64_t bytes_scrubbed;     /* bytes actually zeroed and read back           */
  uint64_t device_total_bytes; /* cudaMemGetInfo total — coverage denominator   */
  uint64_t nonzero_bytes;      /* EXACT count of bytes != 0 found by verify     */
  double ms_acquire;           /* host time inside create/map/set-access; it
                                  overlaps the GPU work (chunk i is scrubbed
                                  while chunk i+1 is being mapped)              */
  double ms_scrub;             /* sum of the scrub kernels' CUDA-event times    */
  double ms_verify;            /* sum of the verify kernels' CUDA-event times   */
  double ms_release;           /* unmap + release when done inside the call; 0
                                  when release_deferred (ask
                                  ccm_scrub_release_wait for the real figure)   */
  double ms_total;             /* host wall-clock of the call = time to verdict */
  int32_t segments;            /* physical chunks mapped                        */
  int32_t scrub_variant;       /* variant that actually ran                     */
  int32_t verify_variant;
  int32_t sm_count;
  int32_t status;              /* ccm_status of this device (batched calls)     */
  int32_t release_deferred;    /* 1: the HBM is being handed back by the
                                  engine's background reaper (see below)        */
  /* ---- ABI v2 ---- */
  uint64_t device_free_before; /* cudaMemGetInfo free when the call started     */
  uint64_t bytes_unreached;    /* free HBM that could NOT be mapped and was
                                  therefore not scrubbed (0 on a clean device)  */
  double ms_release_wait;      /* time this call waited for the PREVIOUS call's
                                  deferred release before it could start        */
  double ms_gpu_span;          /* first scrub launch .. last verify done on the
                                  stream (includes waiting for mappings)        */
} ccm_scrub_result;

/* The product call.  Obtains `bytes` of HBM on device `dev` (0 = everything the
 * context can map, down to the last 2 MiB granule), zero-fills it, reads it back
 * and counts non-zero bytes.  The verdict is returned as soon as the count is on the
 * host; handing the HBM back to the driver (cuMemUnmap / cuMemRelease — the most
 * expensive part of the call, 0.33 ms/GiB, serialised node-wide by the driver)
 * runs on a per-device background thread.  That "reaper" is joined by the next
 * scrub / arena call on the device, by ccm_scrub_release_wait() and by
 * ccm_device_release(); CCM_ASYNC_RELEASE=0 keeps the release inside the call.
 * Returns CCM_ERR_DIRTY if nonzero_bytes != 0, CCM_ERR_NO_CUDA when the device has
 * no CUDA ordinal.  Never falls back to a host path. */
int ccm_scrub_verify(int dev, uint64_t bytes, ccm_scrub_result* out);

/* Blocks until the HBM of the last ccm_scrub_verify() on `dev` is back with the
 * driver.  *ms_release (optional) = duration of that background release,
 * *ms_waited (optional) = how long THIS call blocked.  CCM_OK when nothing is pending. */
int ccm_scrub_release_wait(int dev, double* ms_release, double* ms_waited);

/* Concurrent multi-context launcher: one host thread + primary context + stream
 * per device, no peer access, no collective.  With n > 1 each device maps its whole
 * range before its first launch: driver VMM calls interleaved with kernel launches on
 * several devices of ONE process contend badly (8 x B200: 0.72 s to the verdict
 * pipelined, 0.22 s mapped-first), whereas a single device per process is fastest
 * pipelined (ccm_scrub_verify).  out[i].status is per device;
 * *wall_ms is the host wall-clock of the whole fan-out (max over devices).
 * Returns CCM_OK iff every device returned CCM_OK. */
int ccm_scrub_verify_many(int n, const int* devs, uint64_t bytes,
                          ccm_scrub_result* out, double* wall_ms);

/* -- arena: the device-resident region, held across calls (bench / tests) -- */
typedef struct ccm_arena_info {
  uint64_t bytes;              /* total bytes held                              */
  uint64_t device_total_bytes;
  uint64_t device_free_before; /* cudaMemGetInfo free before acquiring          */
  int32_t segments;
  int32_t reserved;
  double ms_acquire;
} ccm_arena_info;

int ccm_arena_acquire(int dev, uint64_t bytes, ccm_arena_info* out);
int ccm_arena_release(int dev);
/* Zero the whole arena.  *ms (optional) = CUDA-event duration on `stream`. */
int ccm_arena_scrub(int dev, int variant, const ccm_launch_cfg* cfg, void* stream, float* ms);
/* Count non-zero bytes in the whole arena (synchronises to return the count). */
int ccm_arena_verify(int dev, int variant, const ccm_launch_cfg* cfg, void* stream,
                     uint64_t* nonzero, float* ms);
/* Enqueue only (no sync): scrub then verify into the arena's device counter;
 * used inside timed regions.  Fetch the count with ccm_arena_fetch_count(). */
int ccm_arena_scrub_verify_async(int dev, int scrub_variant, int verify_variant,
                                 const ccm_launch_cfg* scrub_cfg,
                                 const ccm_launch_cfg* verify_cfg, void* stream);
int ccm_arena_fetch_count(int dev, void* stream, uint64_t* nonzero);
/* Per-step kernel durations (CUDA events recorded on the launching stream by
 * ccm_arena_scrub_verify_async, up to 64 steps since the last call); waits for
 * the recorded steps to finish, then clears the ring. */
int ccm_arena_step_times(int dev, int cap, float* scrub_ms, float* verify_ms, int* n);
/* Test scaffolding at full size: fill with a byte / a seeded xorshift pattern,
 * poke and peek host bytes at an arena offset. */
int ccm_arena_fill(int dev, int byte_value, void* stream);
int ccm_arena_fill_random(int dev, uint64_t seed, void* stream);
int ccm_arena_write(int dev, uint64_t offset, const void* host_src, uint64_t bytes);
int ccm_arena_read(int dev, uint64_t offset, void* host_dst, uint64_t bytes);

/* -- raw regions: caller-owned device memory (any alignment, any length) --- */
int ccm_region_scrub(int cuda_ordinal, void* dptr, uint64_t bytes, int variant,
                     const ccm_launch_cfg* cfg, void* stream, float* ms);
int ccm_region_verify(int cuda_ordinal, const void* dptr, uint64_t bytes, int variant,
                      const ccm_launch_cfg* cfg, void* stream, uint64_t* nonzero, float* ms);
/* Host-buffer round trip through the public path (tests, e2e): copies `bytes`
 * from host_buf to the device, counts non-zero bytes BEFORE (pre_nonzero, may be
 * NULL), scrubs, verifies, copies the region back into host_buf. */
int ccm_host_roundtrip(int cuda_ordinal, void* host_buf, uint64_t bytes, uint64_t dev_offset,
                       int scrub_variant, int verify_variant,
                       uint64_t* pre_nonzero, uint64_t* post_nonzero);

/* Drops everything the library holds on the CUDA device behind `dev` — arena, stream, events,
 * counters — and resets that device's primary context (cudaDeviceReset).  A daemon must not sit
 * on a CUDA context between transitions: the context pins ~0.5 GB of HBM, keeps the GPU "in use"
 * (it cannot be unbound for vfio) and does not survive the device reset of the next transition
 * (reference main.py:519).  Joins the background release of the last scrub first.  The next scrub
 * call re-creates what it needs.  The sysfs backend
 * calls this implicitly before it resets a device.  No-op (CCM_OK) when nothing is held. */
int ccm_device_release(int dev);
/* Concurrent form for `n` devices (one host thread each); *wall_ms = host wall-clock. */
int ccm_device_release_many(int n, const int* devs, double* wall_ms);

/* Names of the two kernels AUTO launches, as ncu prints them, separated by ';' —
 * "scrub_st256_fast_kernel<512, 8, 0>;verify_ld256_fast_kernel<1024, 4, 2>".  Built from the
 * same compile-time constants as the launches, so tools that key measurements by kernel name
 * (bench.py's roofline.traffic <- profiles/traffic.json) notice when the default changes. */
const char* ccm_default_kernels(void);

/* Number of kernels this library has launched since load (all threads). */
uint64_t ccm_kernel_launches(void);

/* ---- sim backend controls (tests, bench; no-ops on sysfs) ---------------- */
/* Rebuild a synthetic topology: n_gpus GPUs then n_switches NVSwitches. */
int ccm_sim_topology(int n_gpus, int n_switches);
/* key: "cc_mode" "ppcie_mode" "cc_supported" "ppcie_supported" "reset_ms"
 *      "boot_ms" "fail_op" (bitmask of ccm_sim_op) "stuck" (reset ignores staged)
 *      "cuda_ordinal" "scrub_inject" (fault drill: ccm_scrub_verify poisons that many bytes
 *      between the scrub and the read-back — first, unaligned, middle and last byte of each
 *      chunk — so the call must come back CCM_ERR_DIRTY with nonzero_bytes == min(value, 4 x
 *      chunks)).   dev = -1 applies to every device. */
int ccm_sim_set(int dev, const char* key, int64_t value);
int ccm_sim_get(int dev, const char* key, int64_t* value);
typedef enum ccm_sim_op {
  CCM_OP_QUERY_CC = 1, CCM_OP_SET_CC = 2, CCM_OP_QUERY_PPCIE = 4, CCM_OP_SET_PPCIE = 8,
  CCM_OP_RESET = 16, CCM_OP_WAIT_BOOT = 32, CCM_OP_SCRUB = 64
} ccm_sim_op;
/* Ordered trace of register-level ops since the last clear: lines
 * "<seq> <bdf> <op> <arg>\n".  Returns bytes needed (excluding NUL). */
int ccm_sim_trace(char* buf, size_t cap);
int ccm_sim_trace_clear(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CCM_H_ */
