// ccm_scrub.cu — host side of the HBM scrub-and-verify stage: per-device engines
// (primary context + stream + events + counter), the arena that holds "all the HBM
// a context can map", launch-shape selection and the dispatch onto the sm_100a
// kernels in scrub_kernels.cuh.
//
// No reference counterpart: SURVEY.md §0 / §8a row S (the reference stops at
// verify-mode, reference main.py:521-529).  There is NO host fallback anywhere in
// this file: without a usable CUDA device every entry point returns
// CCM_ERR_NO_CUDA / CCM_ERR_CUDA.
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>  // driver-API TYPES only: entry points are resolved at run time (no -lcuda)
#include <cuda_runtime.h>

#include "ccm_internal.h"
#include "scrub_kernels.cuh"

namespace ccm {

static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launches() { return g_launches.load(); }

#define CCM_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return e__ == cudaErrorMemoryAllocation ? CCM_ERR_NOMEM : CCM_ERR_CUDA;       \
    }                                                                               \
  } while (0)

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// A CUDA context is as expensive as the number of hardware work queues ("connections") it sets up, and
// the scrub engine drives exactly ONE stream per GPU.  Measured on B200 (benchmarks/vmm_probe.cu ctx,
// profiles/r2_ctx_connections.log): primary-context creation 138 ms -> 70 ms and reset 170 ms -> 76 ms with
// CUDA_DEVICE_MAX_CONNECTIONS=1 instead of the default 8 (32: 0.7-1 s and 0.8-3 s).  Context creation and
// teardown are what a transition on 8 GPUs mostly consists of, and the driver serialises
// them across GPUs — so the library asks for one connection unless the host process chose a value itself
// (the variable is only read when CUDA initialises; CCM_CUDA_MAX_CONNECTIONS=0 leaves CUDA's default,
// any other number is passed through).
static void apply_cuda_env_defaults() {
  static std::once_flag once;
  std::call_once(once, [] {
    const char* want = getenv("CCM_CUDA_MAX_CONNECTIONS");
    if (want && !strcmp(want, "0")) return;
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", want && *want ? want : "1", 0 /* never override the host's own choice */);
  });
}

int cuda_device_count() {
  apply_cuda_env_defaults();
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int cuda_describe(int ordinal, char* bdf, size_t bdf_cap, char* name, size_t name_cap,
                  uint64_t* total_bytes) {
  cudaDeviceProp prop;
  CCM_CUDA(cudaGetDeviceProperties(&prop, ordinal));
  char bus[64] = {0};
  CCM_CUDA(cudaDeviceGetPCIBusId(bus, sizeof bus, ordinal));
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  snprintf(bdf, bdf_cap, "%s", bus);
  snprintf(name, name_cap, "%s", prop.name);
  if (total_bytes) *total_bytes = prop.totalGlobalMem;
  return CCM_OK;
}

// --------------------------------------------------------------------- engine
struct Segment { uint8_t* ptr; uint64_t bytes; };

static constexpr int kMaxSteps = 64;
static constexpr int kMaxPipeChunks = 32;

// Word indices inside the engine's device control block (each on its own 64-byte line):
//   [0] non-zero byte count of the current verify        [8] grab counter of the GENERIC kernels
//   [16]/[24] grab + done of scrub_st256_fast_kernel      [32]/[40] grab + done of verify_ld256_fast_kernel
// The fast kernels hand their pair back zeroed (GrabCtl), so they never share words with the
// generic kernels, whose counter is memset by the host before every launch.
static constexpr int kCtlWords = 64;
static constexpr int kGenericGrab = 8, kScrubGrab = 16, kScrubDone = 24, kVerifyGrab = 32, kVerifyDone = 40;

// One contiguous virtual range backed by VMM chunks (cuMemCreate / cuMemMap).
struct VmmMapping {
  CUdeviceptr base = 0;
  uint64_t va_bytes = 0, mapped = 0;
  std::vector<CUmemGenericAllocationHandle> handles;
};

struct ScrubEngine {
  int ordinal = -1;
  std::mutex mu;
  bool ready = false;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;   // the engine's own non-blocking stream
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t step_ev[kMaxSteps][3] = {};
  cudaEvent_t pipe_ev[kMaxPipeChunks][3] = {};
  int step_count = 0;
  unsigned long long* d_counter = nullptr;
  unsigned long long* h_counter = nullptr;  // pinned
  std::vector<Segment> segs;
  uint64_t arena_bytes = 0;
  size_t total_bytes = 0;
  // deferred give-back of the last gate's HBM: runs off the caller's critical path and is
  // joined by whoever needs the memory (or the context) next
  std::thread reaper;
  double reaper_ms = 0;        // duration of the last background release

  ~ScrubEngine() { if (reaper.joinable()) reaper.join(); }

  int init();
  cudaStream_t pick(void* s) const { return s ? (cudaStream_t)s : stream; }
  GrabCtl scrub_ctl() const { return GrabCtl{d_counter + kScrubGrab, d_counter + kScrubDone}; }
  GrabCtl verify_ctl() const { return GrabCtl{d_counter + kVerifyGrab, d_counter + kVerifyDone}; }
  // Waits for the background release of the previous gate (no-op when none is pending).
  double join_reaper() {
    if (!reaper.joinable()) return 0.0;
    const double t0 = now_ms();
    reaper.join();
    return now_ms() - t0;
  }
};

// default launch shapes of the two AUTO kernels
static constexpr int kFastScrubThreads = 512, kFastScrubPer = 8;
static constexpr int kFastVerifyThreads = 1024, kFastVerifyPer = 4;
#define CCM_FAST_SCRUB scrub_st256_fast_kernel<kFastScrubThreads, kFastScrubPer, kPolDefault>
#define CCM_FAST_VERIFY verify_ld256_fast_kernel<kFastVerifyThreads, kFastVerifyPer, kPolStreaming>

const char* default_kernel_names() {
  static const std::string names = [] {
    char b[160];
    snprintf(b, sizeof b, "scrub_st256_fast_kernel<%d, %d, %d>;verify_ld256_fast_kernel<%d, %d, %d>",
             kFastScrubThreads, kFastScrubPer, (int)kPolDefault, kFastVerifyThreads, kFastVerifyPer, (int)kPolStreaming);
    return std::string(b);
  }();
  return names.c_str();
}

int ScrubEngine::init() {
  if (ready) return CCM_OK;
  CCM_CUDA(cudaSetDevice(ordinal));
  // force primary-context creation here.  (Taking turns inside the process — one context at a time — was
  // measured and changes nothing: 8 GPUs, 4.19 s vs 4.17 s per transition; the driver queues them anyway.)
  CCM_CUDA(cudaFree(0));
  int v = 0;
  CCM_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, ordinal));
  sm_count = v;
  CCM_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, ordinal));
  smem_optin = (size_t)v;
  CCM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  for (auto& e : ev) CCM_CUDA(cudaEventCreate(&e));
  for (auto& s : step_ev) for (auto& e : s) CCM_CUDA(cudaEventCreate(&e));
  for (auto& s : pipe_ev) for (auto& e : s) CCM_CUDA(cudaEventCreate(&e));
  CCM_CUDA(cudaMalloc(&d_counter, kCtlWords * sizeof(unsigned long long)));
  CCM_CUDA(cudaMemset(d_counter, 0, kCtlWords * sizeof(unsigned long long)));
  CCM_CUDA(cudaHostAlloc(&h_counter, 64, cudaHostAllocDefault));
  // The product call maps HBM down to the last granule; CUDA loads kernel code lazily at the
  // first launch and that needs device memory — make the two default kernels resident now.
  cudaFuncAttributes fa;
  CCM_CUDA(cudaFuncGetAttributes(&fa, CCM_FAST_SCRUB));
  CCM_CUDA(cudaFuncGetAttributes(&fa, CCM_FAST_VERIFY));
  size_t fr = 0;
  CCM_CUDA(cudaMemGetInfo(&fr, &total_bytes));
  ready = true;
  return CCM_OK;
}

static std::mutex g_engines_mu;
static std::vector<std::unique_ptr<ScrubEngine>> g_engines;

ScrubEngine* engine_for(int ordinal) {
  if (ordinal < 0) { set_error("device has no CUDA ordinal"); return nullptr; }
  apply_cuda_env_defaults();
  ScrubEngine* e = nullptr;
  {
    std::lock_guard<std::mutex> g(g_engines_mu);
    if ((size_t)ordinal >= g_engines.size()) {
      int n = cuda_device_count();
      if (ordinal >= n) { set_error("CUDA ordinal %d not present (%d devices)", ordinal, n); return nullptr; }
      while (g_engines.size() < (size_t)n) {
        g_engines.emplace_back(new ScrubEngine());
        g_engines.back()->ordinal = (int)g_engines.size() - 1;
      }
    }
    e = g_engines[ordinal].get();
  }
  std::lock_guard<std::mutex> g(e->mu);
  if (e->init() != CCM_OK) return nullptr;
  return e;
}

// The engine of an ordinal if one was ever created — WITHOUT (re)creating a CUDA context for it.
ScrubEngine* engine_lookup(int ordinal) {
  std::lock_guard<std::mutex> g(g_engines_mu);
  if (ordinal < 0 || (size_t)ordinal >= g_engines.size()) return nullptr;
  return g_engines[ordinal].get();
}

// Every engine_* entry point calls this right after taking e->mu: a concurrent
// ccm_device_release / sysfs reset may have torn the engine down between engine_for() and the
// lock, in which case it is rebuilt here instead of running on dead handles.
#define CCM_ENSURE_READY(e)                        \
  do {                                             \
    int rc__ = (e)->init();                        \
    if (rc__ != CCM_OK) return rc__;               \
  } while (0)

int engine_teardown(int ordinal) {
  ScrubEngine* e = nullptr;
  {
    std::lock_guard<std::mutex> g(g_engines_mu);
    if (ordinal < 0 || (size_t)ordinal >= g_engines.size()) return CCM_OK;  // never created: nothing held
    e = g_engines[ordinal].get();
  }
  std::lock_guard<std::mutex> g(e->mu);
  e->join_reaper();  // the last gate's HBM must be back with the driver before the context goes
  if (!e->ready) return CCM_OK;
  CCM_CUDA(cudaSetDevice(ordinal));
  cudaStreamSynchronize(e->stream);
  for (auto& s : e->segs) cudaFree(s.ptr);
  e->segs.clear();
  e->arena_bytes = 0;
  for (auto& ev : e->ev) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
  for (auto& st : e->step_ev) for (auto& ev : st) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
  for (auto& st : e->pipe_ev) for (auto& ev : st) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
  e->step_count = 0;
  if (e->d_counter) cudaFree(e->d_counter);
  if (e->h_counter) cudaFreeHost(e->h_counter);
  e->d_counter = nullptr;
  e->h_counter = nullptr;
  if (e->stream) cudaStreamDestroy(e->stream);
  e->stream = nullptr;
  e->ready = false;
  CCM_CUDA(cudaDeviceReset());  // destroys the primary context of the current device
  return CCM_OK;
}

// ------------------------------------------------------------- launch shapes
struct Shape { int ctas_per_sm, threads, unroll, policy, tile_bytes, schedule; };  // schedule: 1 static, 2 dynamic

// Defaults = winners of the round-1 sweeps on B200 (profiles/r1_sweep_schedule.md):
// persistent CTAs + dynamic chunk grabs beat every static split by 10-20 %.
static Shape scrub_shape(int variant, const ccm_launch_cfg* c) {
  Shape s;
  if (variant == CCM_SCRUB_TMA) s = Shape{1, 32, 1, kPolDefault, 65536, 2};
  else s = Shape{1, 512, 4, kPolDefault, 131072, 2};
  if (c) {
    if (c->ctas_per_sm > 0) s.ctas_per_sm = c->ctas_per_sm;
    if (c->threads_per_cta > 0) s.threads = c->threads_per_cta;
    if (c->unroll > 0) s.unroll = c->unroll;
    if (c->cache_policy >= 1 && c->cache_policy <= 4) s.policy = c->cache_policy - 1;
    if (c->tile_bytes > 0) s.tile_bytes = c->tile_bytes;
    if (c->schedule >= 1 && c->schedule <= 3) s.schedule = c->schedule;
  }
  return s;
}
static Shape verify_shape(int variant, const ccm_launch_cfg* c) {
  (void)variant;
  Shape s = Shape{1, 1024, 4, kPolStreaming, 131072, 2};
  if (c) {
    if (c->ctas_per_sm > 0) s.ctas_per_sm = c->ctas_per_sm;
    if (c->threads_per_cta > 0) s.threads = c->threads_per_cta;
    if (c->unroll > 0) s.unroll = c->unroll;
    if (c->cache_policy >= 1 && c->cache_policy <= 4) s.policy = c->cache_policy - 1;
    if (c->tile_bytes > 0) s.tile_bytes = c->tile_bytes;
    if (c->schedule >= 1 && c->schedule <= 3) s.schedule = c->schedule;
  }
  return s;
}

// AUTO = the fastest variant measured on B200 (profiles/README.md).  ST256 and TMA tie on
// throughput (7 618 vs 7 607 GB/s); TMA draws ~5 % less board power (601 vs 631 W,
// profiles/r1_power_probe.json), so a deployment that prefers joules can pin it with
// CCM_SCRUB_VARIANT=tma (st128 | st256 | tma | memset; CCM_VERIFY_VARIANT=ld128 | ld256).
static int env_variant(const char* name, const char* const* names, int n, int dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  for (int i = 0; i < n; ++i)
    if (!strcmp(v, names[i])) return i + 1;
  return dflt;
}
static int resolve_scrub_variant(int v) {
  static const char* const names[] = {"st128", "st256", "tma", "memset"};
  return v == CCM_SCRUB_AUTO ? env_variant("CCM_SCRUB_VARIANT", names, 4, CCM_SCRUB_ST256) : v;
}
static int resolve_verify_variant(int v) {
  static const char* const names[] = {"ld128", "ld256"};
  return v == CCM_VERIFY_AUTO ? env_variant("CCM_VERIFY_VARIANT", names, 2, CCM_VERIFY_LD256) : v;
}

template <int VB, int UNROLL>
static cudaError_t launch_scrub_st_pol(const RegionSplit& s, int grid, int threads, int pol, Sched sc, cudaStream_t st) {
  switch (pol) {
    case kPolEvictFirst: scrub_st_kernel<VB, UNROLL, kPolEvictFirst><<<grid, threads, 0, st>>>(s, sc); break;
    case kPolStreaming:  scrub_st_kernel<VB, UNROLL, kPolStreaming><<<grid, threads, 0, st>>>(s, sc); break;
    case kPolEvictLast:  scrub_st_kernel<VB, UNROLL, kPolEvictLast><<<grid, threads, 0, st>>>(s, sc); break;
    default:             scrub_st_kernel<VB, UNROLL, kPolDefault><<<grid, threads, 0, st>>>(s, sc); break;
  }
  return cudaGetLastError();
}
template <int VB>
static cudaError_t launch_scrub_st(const RegionSplit& s, int grid, const Shape& sh, Sched sc, cudaStream_t st) {
  switch (sh.unroll) {
    case 1: return launch_scrub_st_pol<VB, 1>(s, grid, sh.threads, sh.policy, sc, st);
    case 2: return launch_scrub_st_pol<VB, 2>(s, grid, sh.threads, sh.policy, sc, st);
    case 8: return launch_scrub_st_pol<VB, 8>(s, grid, sh.threads, sh.policy, sc, st);
    case 16: return launch_scrub_st_pol<VB, 16>(s, grid, sh.threads, sh.policy, sc, st);
    default: return launch_scrub_st_pol<VB, 4>(s, grid, sh.threads, sh.policy, sc, st);
  }
}

template <int VB, int UNROLL>
static cudaError_t launch_verify_ld_pol(const RegionSplit& s, int grid, int threads, int pol,
                                        unsigned long long* ctr, Sched sc, cudaStream_t st) {
  switch (pol) {
    case kPolEvictFirst: verify_ld_kernel<VB, UNROLL, kPolEvictFirst><<<grid, threads, 0, st>>>(s, ctr, sc); break;
    case kPolStreaming:  verify_ld_kernel<VB, UNROLL, kPolStreaming><<<grid, threads, 0, st>>>(s, ctr, sc); break;
    case kPolEvictLast:  verify_ld_kernel<VB, UNROLL, kPolEvictLast><<<grid, threads, 0, st>>>(s, ctr, sc); break;
    default:             verify_ld_kernel<VB, UNROLL, kPolDefault><<<grid, threads, 0, st>>>(s, ctr, sc); break;
  }
  return cudaGetLastError();
}
template <int VB>
static cudaError_t launch_verify_ld(const RegionSplit& s, int grid, const Shape& sh,
                                    unsigned long long* ctr, Sched sc, cudaStream_t st) {
  switch (sh.unroll) {
    case 1: return launch_verify_ld_pol<VB, 1>(s, grid, sh.threads, sh.policy, ctr, sc, st);
    case 2: return launch_verify_ld_pol<VB, 2>(s, grid, sh.threads, sh.policy, ctr, sc, st);
    case 8: return launch_verify_ld_pol<VB, 8>(s, grid, sh.threads, sh.policy, ctr, sc, st);
    default: return launch_verify_ld_pol<VB, 4>(s, grid, sh.threads, sh.policy, ctr, sc, st);
  }
}

static int clamp_threads(int t) {
  if (t < 32) t = 32;
  if (t > 1024) t = 1024;
  return (t / 32) * 32;
}

// Work-distribution descriptor for one launch of a GENERIC kernel.  Dynamic: the grab counter
// lives at d_counter[kGenericGrab] (its own 64-byte line) and is zeroed on the launching stream
// first.  (The default fast kernels reset their own counters: GrabCtl.)
static int make_sched(ScrubEngine* e, const Shape& sh, uint64_t body_bytes, uint64_t tile_bytes, bool tma,
                      cudaStream_t st, Sched* out) {
  out->counter = nullptr;
  out->chunk_tiles = 1;
  out->per_warp = 0;
  if (sh.schedule == 3 && !tma) {  // warp-granular grabs: a tile is what ONE warp covers
    out->per_warp = 1;
    tile_bytes = tile_bytes / (uint64_t)sh.threads * 32;
  }
  out->ntiles = tile_bytes ? body_bytes / tile_bytes : 0;
  out->nchunks = out->ntiles;
  if (sh.schedule != 2 && sh.schedule != 3) return CCM_OK;
  CCM_CUDA(cudaMemsetAsync(e->d_counter + kGenericGrab, 0, sizeof(unsigned long long), st));
  out->counter = e->d_counter + kGenericGrab;
  if (tma) {
    out->chunk_tiles = sh.unroll > 0 ? (uint32_t)sh.unroll : 4;
  } else {
    const uint64_t chunk_bytes = sh.tile_bytes > 0 ? (uint64_t)sh.tile_bytes : 256 * 1024;
    uint64_t ct = chunk_bytes / (tile_bytes ? tile_bytes : 1);
    out->chunk_tiles = (uint32_t)(ct < 1 ? 1 : ct);
  }
  out->nchunks = (out->ntiles + out->chunk_tiles - 1) / out->chunk_tiles;
  return CCM_OK;
}

// Launch with the programmatic-stream-serialization attribute (CCM_PDL=1): the kernel may be
// scheduled while the previous kernel on the stream drains; it blocks in griddepcontrol.wait
// until that kernel has completed and flushed, so ordering is unchanged.
static bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("CCM_PDL"); return v && *v && strcmp(v, "0") != 0; }();
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_fast(void (*kernel)(KArgs...), int grid, int threads, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof lc);
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3((unsigned)threads);
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&lc, kernel, KArgs(args)...);
}

// True when (variant, cfg) resolves to the compile-time-shape default kernel.
static bool is_fast_scrub(int variant, const ccm_launch_cfg* cfg) {
  return cfg == nullptr && resolve_scrub_variant(variant) == CCM_SCRUB_ST256;
}

// Scrub one contiguous range on `st` (no sync).  zero_on_exit: see scrub_st256_fast_kernel
// (honoured by the default kernel only — callers check is_fast_scrub()).
static int scrub_range(ScrubEngine* e, void* p, uint64_t n, int variant, const ccm_launch_cfg* cfg,
                       cudaStream_t st, unsigned long long* zero_on_exit = nullptr) {
  if (n == 0) return CCM_OK;
  variant = resolve_scrub_variant(variant);
  Shape sh = scrub_shape(variant, cfg);
  sh.threads = clamp_threads(sh.threads);
  // the host computes the tile count, so it must use the UNROLL that is actually instantiated
  if (variant != CCM_SCRUB_TMA && sh.unroll != 1 && sh.unroll != 2 && sh.unroll != 8 && sh.unroll != 16) sh.unroll = 4;
  const int grid = e->sm_count * (sh.ctas_per_sm > 0 ? sh.ctas_per_sm : 1);
  cudaError_t err = cudaSuccess;
  switch (variant) {
    case CCM_SCRUB_MEMSET:
      err = cudaMemsetAsync(p, 0, n, st);
      break;
    case CCM_SCRUB_ST128: {
      RegionSplit s = split_region(p, n, 16, 128);
      Sched sc;
      if (int rc = make_sched(e, sh, s.body_vecs * 16, (uint64_t)sh.threads * sh.unroll * 16, false, st, &sc)) return rc;
      err = launch_scrub_st<16>(s, grid, sh, sc, st);
      g_launches++;
      break;
    }
    case CCM_SCRUB_ST256: {
      RegionSplit s = split_region(p, n, 32, 128);
      if (cfg == nullptr) {
        // library default: compile-time shape (148 persistent CTAs x 512 threads, 8 x STG.256 per
        // thread per 128 KiB grab) — scrub_st256_fast_kernel
        const uint64_t nchunks = s.body_vecs * 32 / ((uint64_t)kFastScrubThreads * kFastScrubPer * 32);
        err = launch_fast(CCM_FAST_SCRUB, e->sm_count, kFastScrubThreads, st, s, nchunks, e->scrub_ctl(), zero_on_exit);
        g_launches++;
        break;
      }
      Sched sc;
      if (int rc = make_sched(e, sh, s.body_vecs * 32, (uint64_t)sh.threads * sh.unroll * 32, false, st, &sc)) return rc;
      err = launch_scrub_st<32>(s, grid, sh, sc, st);
      g_launches++;
      break;
    }
    case CCM_SCRUB_TMA: {
      RegionSplit s = split_region(p, n, 16, 128);
      uint32_t tile = (uint32_t)sh.tile_bytes & ~15u;
      if (tile < 1024) tile = 1024;
      if (tile > e->smem_optin - 1024) tile = (uint32_t)((e->smem_optin - 1024) & ~127ull);
      const int threads = sh.threads > 256 ? 256 : sh.threads;
      const uint32_t ops_per_group = sh.unroll > 0 ? (uint32_t)sh.unroll : 4;
      Sched sc;
      if (int rc = make_sched(e, sh, s.body_vecs * 16, tile, true, st, &sc)) return rc;
#define CCM_TMA_LAUNCH(POL)                                                                       \
      do {                                                                                        \
        err = cudaFuncSetAttribute(scrub_tma_kernel<POL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile); \
        if (err == cudaSuccess) {                                                                 \
          scrub_tma_kernel<POL><<<grid, threads, tile, st>>>(s, tile, ops_per_group, sc);         \
          err = cudaGetLastError();                                                               \
        }                                                                                         \
      } while (0)
      if (sh.policy == kPolEvictFirst) CCM_TMA_LAUNCH(kPolEvictFirst);
      else if (sh.policy == kPolEvictLast) CCM_TMA_LAUNCH(kPolEvictLast);
      else CCM_TMA_LAUNCH(kPolDefault);
#undef CCM_TMA_LAUNCH
      g_launches++;
      break;
    }
    default:
      set_error("unknown scrub variant %d", variant);
      return CCM_ERR_INVALID;
  }
  if (err != cudaSuccess) {
    set_error("scrub launch (variant %d) failed: %s", variant, cudaGetErrorString(err));
    return CCM_ERR_CUDA;
  }
  return CCM_OK;
}

// Accumulate the non-zero byte count of one range into e->d_counter (no sync).
static int verify_range(ScrubEngine* e, const void* p, uint64_t n, int variant, const ccm_launch_cfg* cfg,
                        cudaStream_t st) {
  if (n == 0) return CCM_OK;
  variant = resolve_verify_variant(variant);
  Shape sh = verify_shape(variant, cfg);
  sh.threads = clamp_threads(sh.threads);
  if (sh.unroll != 1 && sh.unroll != 2 && sh.unroll != 8) sh.unroll = 4;
  const int grid = e->sm_count * (sh.ctas_per_sm > 0 ? sh.ctas_per_sm : 1);
  cudaError_t err = cudaSuccess;
  switch (variant) {
    case CCM_VERIFY_LD128: {
      RegionSplit s = split_region(p, n, 16, 128);
      Sched sc;
      if (int rc = make_sched(e, sh, s.body_vecs * 16, (uint64_t)sh.threads * sh.unroll * 16, false, st, &sc)) return rc;
      err = launch_verify_ld<16>(s, grid, sh, e->d_counter, sc, st);
      break;
    }
    case CCM_VERIFY_LD256: {
      RegionSplit s = split_region(p, n, 32, 128);
      if (cfg == nullptr) {
        // library default: 148 persistent CTAs x 1024 threads, 4 x LDG.256 in flight per thread
        // per 128 KiB grab — verify_ld256_fast_kernel
        const uint64_t nchunks = s.body_vecs * 32 / ((uint64_t)kFastVerifyThreads * kFastVerifyPer * 32);
        err = launch_fast(CCM_FAST_VERIFY, e->sm_count, kFastVerifyThreads, st, s, nchunks, e->verify_ctl(), e->d_counter);
        break;
      }
      Sched sc;
      if (int rc = make_sched(e, sh, s.body_vecs * 32, (uint64_t)sh.threads * sh.unroll * 32, false, st, &sc)) return rc;
      err = launch_verify_ld<32>(s, grid, sh, e->d_counter, sc, st);
      break;
    }
    default:
      set_error("unknown verify variant %d", variant);
      return CCM_ERR_INVALID;
  }
  g_launches++;
  if (err != cudaSuccess) {
    set_error("verify launch (variant %d) failed: %s", variant, cudaGetErrorString(err));
    return CCM_ERR_CUDA;
  }
  return CCM_OK;
}

// ----------------------------------------------------------------------- arena
static constexpr uint64_t kMiB = 1ull << 20;

static uint64_t env_u64(const char* name, uint64_t dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return strtoull(v, nullptr, 10);
}

int engine_arena_release(ScrubEngine* e, double* ms) {
  std::lock_guard<std::mutex> g(e->mu);
  if (!e->ready) { e->segs.clear(); e->arena_bytes = 0; if (ms) *ms = 0; return CCM_OK; }  // torn down: nothing held
  const double t0 = now_ms();
  CCM_CUDA(cudaSetDevice(e->ordinal));
  int rc = CCM_OK;
  for (auto& s : e->segs) {
    cudaError_t err = cudaFree(s.ptr);
    if (err != cudaSuccess) { set_error("cudaFree failed: %s", cudaGetErrorString(err)); rc = CCM_ERR_CUDA; }
  }
  e->segs.clear();
  e->arena_bytes = 0;
  if (ms) *ms = now_ms() - t0;
  return rc;
}

int engine_arena_acquire(ScrubEngine* e, uint64_t bytes, ccm_arena_info* out) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  if (!e->segs.empty()) { set_error("arena already held on CUDA device %d", e->ordinal); return CCM_ERR_STATE; }
  e->join_reaper();  // HBM of the previous gate must be back before "all free HBM" is measured
  const double t0 = now_ms();
  CCM_CUDA(cudaSetDevice(e->ordinal));
  size_t fr = 0, tot = 0;
  CCM_CUDA(cudaMemGetInfo(&fr, &tot));
  e->total_bytes = tot;
  const bool want_max = (bytes == 0);
  // Leave a little HBM for the driver (launch-time local memory, event pools).
  const uint64_t reserve = env_u64("CCM_ARENA_RESERVE_MB", 256) * kMiB;
  uint64_t want = bytes;
  if (want_max) want = fr > reserve ? ((fr - reserve) & ~(2 * kMiB - 1)) : 0;
  if (want == 0) { set_error("no free HBM to scrub (free=%zu)", fr); return CCM_ERR_NOMEM; }

  // First choice: ONE contiguous virtual range.  Fall back to a segment list
  // (largest-first, halving) when HBM is fragmented.
  uint8_t* p = nullptr;
  // CCM_ARENA_MAX_SEGMENT_MB forces the segmented path (tests / fragmented HBM drills).
  const uint64_t max_seg = env_u64("CCM_ARENA_MAX_SEGMENT_MB", 0) * kMiB;
  cudaError_t err = max_seg ? cudaErrorMemoryAllocation : cudaMalloc(&p, want);
  if (err == cudaSuccess) {
    e->segs.push_back({p, want});
    e->arena_bytes = want;
  } else {
    cudaGetLastError();
    uint64_t got = 0, chunk = max_seg ? max_seg : 16ull << 30;
    const uint64_t min_chunk = max_seg && max_seg < 64 * kMiB ? max_seg : 64 * kMiB;
    while (got < want && chunk >= min_chunk) {
      uint64_t ask = want - got < chunk ? want - got : chunk;
      err = cudaMalloc(&p, ask);
      if (err == cudaSuccess) { e->segs.push_back({p, ask}); got += ask; }
      else { cudaGetLastError(); chunk >>= 1; }
    }
    if (!want_max && got < want) {
      for (auto& s : e->segs) cudaFree(s.ptr);
      e->segs.clear();
      set_error("could not obtain %llu bytes of HBM on CUDA device %d (free %zu)",
                (unsigned long long)want, e->ordinal, fr);
      return CCM_ERR_NOMEM;
    }
    if (got == 0) { set_error("could not obtain any HBM on CUDA device %d", e->ordinal); return CCM_ERR_NOMEM; }
    e->arena_bytes = got;
  }
  if (out) {
    out->bytes = e->arena_bytes;
    out->device_total_bytes = tot;
    out->device_free_before = fr;
    out->segments = (int)e->segs.size();
    out->reserved = 0;
    out->ms_acquire = now_ms() - t0;
  }
  return CCM_OK;
}

static int need_arena(ScrubEngine* e) {
  if (e->segs.empty()) { set_error("no arena held on CUDA device %d", e->ordinal); return CCM_ERR_STATE; }
  return CCM_OK;
}

int engine_arena_scrub(ScrubEngine* e, int variant, const ccm_launch_cfg* cfg, void* stream, float* ms) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  int rc = need_arena(e); if (rc) return rc;
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->pick(stream);
  if (ms) CCM_CUDA(cudaEventRecord(e->ev[0], st));
  for (auto& s : e->segs) { rc = scrub_range(e, s.ptr, s.bytes, variant, cfg, st); if (rc) return rc; }
  if (ms) {
    CCM_CUDA(cudaEventRecord(e->ev[1], st));
    CCM_CUDA(cudaEventSynchronize(e->ev[1]));
    CCM_CUDA(cudaEventElapsedTime(ms, e->ev[0], e->ev[1]));
  }
  return CCM_OK;
}

static int fetch_count_locked(ScrubEngine* e, cudaStream_t st, uint64_t* nonzero) {
  CCM_CUDA(cudaMemcpyAsync(e->h_counter, e->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CCM_CUDA(cudaStreamSynchronize(st));
  if (nonzero) *nonzero = (uint64_t)*e->h_counter;
  return CCM_OK;
}

int engine_arena_verify(ScrubEngine* e, int variant, const ccm_launch_cfg* cfg, void* stream,
                        uint64_t* nonzero, float* ms) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  int rc = need_arena(e); if (rc) return rc;
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->pick(stream);
  CCM_CUDA(cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), st));
  CCM_CUDA(cudaEventRecord(e->ev[0], st));
  for (auto& s : e->segs) { rc = verify_range(e, s.ptr, s.bytes, variant, cfg, st); if (rc) return rc; }
  CCM_CUDA(cudaEventRecord(e->ev[1], st));
  rc = fetch_count_locked(e, st, nonzero); if (rc) return rc;
  if (ms) CCM_CUDA(cudaEventElapsedTime(ms, e->ev[0], e->ev[1]));
  return CCM_OK;
}

int engine_arena_scrub_verify_async(ScrubEngine* e, int sv, int vv, const ccm_launch_cfg* scfg,
                                    const ccm_launch_cfg* vcfg, void* stream) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  int rc = need_arena(e); if (rc) return rc;
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->pick(stream);
  cudaEvent_t* evs = e->step_count < kMaxSteps ? e->step_ev[e->step_count] : nullptr;
  if (evs) CCM_CUDA(cudaEventRecord(evs[0], st));
  // Default kernels: the scrub's last CTA also clears the verify counter, so one step is
  // exactly two kernel nodes per segment and no memset nodes.
  const bool fast = is_fast_scrub(sv, scfg);
  // With programmatic dependent launch the two kernels must be ADJACENT on the stream (an event record
  // in between serialises them fully): the split event is then recorded up front and the whole pair
  // shows up as "verify" time.
  const bool chained = fast && pdl_enabled();
  if (evs && chained) CCM_CUDA(cudaEventRecord(evs[1], st));
  for (auto& s : e->segs) {
    rc = scrub_range(e, s.ptr, s.bytes, sv, scfg, st, fast ? e->d_counter : nullptr);
    if (rc) return rc;
  }
  if (!fast) CCM_CUDA(cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), st));
  if (evs && !chained) CCM_CUDA(cudaEventRecord(evs[1], st));
  for (auto& s : e->segs) { rc = verify_range(e, s.ptr, s.bytes, vv, vcfg, st); if (rc) return rc; }
  if (evs) { CCM_CUDA(cudaEventRecord(evs[2], st)); e->step_count++; }
  return CCM_OK;
}

int engine_arena_step_times(ScrubEngine* e, int cap, float* scrub_ms, float* verify_ms, int* n) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  CCM_CUDA(cudaSetDevice(e->ordinal));
  int k = e->step_count < cap ? e->step_count : cap;
  for (int i = 0; i < k; ++i) {
    CCM_CUDA(cudaEventSynchronize(e->step_ev[i][2]));
    if (scrub_ms) CCM_CUDA(cudaEventElapsedTime(&scrub_ms[i], e->step_ev[i][0], e->step_ev[i][1]));
    if (verify_ms) CCM_CUDA(cudaEventElapsedTime(&verify_ms[i], e->step_ev[i][1], e->step_ev[i][2]));
  }
  if (n) *n = k;
  e->step_count = 0;
  return CCM_OK;
}

int engine_arena_fetch_count(ScrubEngine* e, void* stream, uint64_t* nonzero) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  CCM_CUDA(cudaSetDevice(e->ordinal));
  return fetch_count_locked(e, e->pick(stream), nonzero);
}

__global__ void fill_byte_kernel(uint4* p, uint64_t nvec, uint32_t word) {
  const uint4 v = make_uint4(word, word, word, word);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

int engine_arena_fill(ScrubEngine* e, int byte_value, void* stream) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  int rc = need_arena(e); if (rc) return rc;
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->pick(stream);
  const uint32_t b = (uint32_t)(byte_value & 0xff);
  const uint32_t word = b | (b << 8) | (b << 16) | (b << 24);
  for (auto& s : e->segs) {
    const uint64_t nvec = s.bytes / 16;
    if (nvec) { fill_byte_kernel<<<e->sm_count * 8, 256, 0, st>>>((uint4*)s.ptr, nvec, word); g_launches++; }
    const uint64_t rem = s.bytes - nvec * 16;
    if (rem) CCM_CUDA(cudaMemsetAsync(s.ptr + nvec * 16, (int)b, rem, st));
    CCM_CUDA(cudaGetLastError());
  }
  CCM_CUDA(cudaStreamSynchronize(st));
  return CCM_OK;
}

int engine_arena_fill_random(ScrubEngine* e, uint64_t seed, void* stream) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  int rc = need_arena(e); if (rc) return rc;
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->pick(stream);
  uint64_t word0 = 0;
  for (auto& s : e->segs) {
    const uint64_t nwords = s.bytes / 8;
    if (nwords) { fill_pattern_kernel<<<e->sm_count * 8, 256, 0, st>>>((uint64_t*)s.ptr, nwords, word0, seed); g_launches++; }
    const uint64_t rem = s.bytes - nwords * 8;  // trailing bytes of the pattern are zero
    if (rem) CCM_CUDA(cudaMemsetAsync(s.ptr + nwords * 8, 0, rem, st));
    CCM_CUDA(cudaGetLastError());
    word0 += nwords;
  }
  CCM_CUDA(cudaStreamSynchronize(st));
  return CCM_OK;
}

int engine_arena_rw(ScrubEngine* e, uint64_t offset, void* host, uint64_t bytes, bool write) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  int rc = need_arena(e); if (rc) return rc;
  if (offset + bytes > e->arena_bytes || offset + bytes < offset) {
    set_error("arena access [%llu,+%llu) outside %llu bytes", (unsigned long long)offset,
              (unsigned long long)bytes, (unsigned long long)e->arena_bytes);
    return CCM_ERR_INVALID;
  }
  CCM_CUDA(cudaSetDevice(e->ordinal));
  // Pokes and peeks are synchronous copies on the legacy stream; the scrub/verify launches they
  // interleave with run on NON-BLOCKING streams (the engine's, or the caller's) and are not ordered
  // against it — a 25 ms full-arena scrub still in flight would overwrite a byte poked "after" it.
  CCM_CUDA(cudaDeviceSynchronize());
  uint8_t* h = (uint8_t*)host;
  uint64_t seg_start = 0;
  for (auto& s : e->segs) {
    const uint64_t seg_end = seg_start + s.bytes;
    if (bytes && offset < seg_end && offset + bytes > seg_start) {
      const uint64_t a = offset > seg_start ? offset : seg_start;
      const uint64_t b = offset + bytes < seg_end ? offset + bytes : seg_end;
      if (write) CCM_CUDA(cudaMemcpy(s.ptr + (a - seg_start), h + (a - offset), b - a, cudaMemcpyHostToDevice));
      else CCM_CUDA(cudaMemcpy(h + (a - offset), s.ptr + (a - seg_start), b - a, cudaMemcpyDeviceToHost));
    }
    seg_start = seg_end;
  }
  return CCM_OK;
}

// ------------------------------------------------------- VMM (driver API, lazy)
// The product call maps HBM in chunks through the virtual-memory-management API so
// that (a) the chunks sit in ONE contiguous virtual range (one verify launch) and (b)
// the GPU scrubs chunk i while the host is still creating chunk i+1.  Entry points come from
// cudaGetDriverEntryPoint, so libccm.so has no link-time dependency on libcuda and
// still loads on a CPU-only box.  (profiles/r1_alloc_probe.log: cudaMalloc/cudaFree
// of 190 GB cost 110-380 ms + 60 ms when called back to back; the kernels need 50 ms.)
struct VmmApi {
  bool ok = false;
  CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};

static const VmmApi& vmm() {
  static VmmApi api = [] {
    VmmApi a;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult st;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess &&
             st == cudaDriverEntryPointSuccess && *fn != nullptr;
    };
    a.ok = get("cuMemAddressReserve", (void**)&a.AddressReserve) && get("cuMemAddressFree", (void**)&a.AddressFree) &&
           get("cuMemCreate", (void**)&a.Create) && get("cuMemRelease", (void**)&a.Release) &&
           get("cuMemMap", (void**)&a.Map) && get("cuMemUnmap", (void**)&a.Unmap) &&
           get("cuMemSetAccess", (void**)&a.SetAccess) &&
           get("cuMemGetAllocationGranularity", (void**)&a.GetGranularity);
    if (!a.ok) cudaGetLastError();
    return a;
  }();
  return api;
}

// Hands a mapping back to the driver: unmap, release every physical chunk, free the range.
// (One cuMemUnmap over the whole range: per-chunk unmaps are 10x slower, and so is unmapping
// under a running kernel — benchmarks/vmm_probe.cu, profiles/r2_vmm_probe_1gpu.log.)
static double give_back(const VmmApi& api, VmmMapping& m) {
  const double t0 = now_ms();
  if (m.mapped) api.Unmap(m.base, m.mapped);
  for (auto h : m.handles) api.Release(h);
  if (m.base) api.AddressFree(m.base, m.va_bytes);
  m = VmmMapping();
  return now_ms() - t0;
}

static bool async_release_enabled() { return env_u64("CCM_ASYNC_RELEASE", 1) != 0; }

// Pipelined scrub-and-verify over freshly mapped HBM.  Returns CCM_ERR_UNSUPPORTED when
// the VMM path cannot be used at all (caller falls back to the arena path).
//
//   host    : create+map+access chunk 0 | chunk 1 | chunk 2 | ...            | D2H count, verdict
//   stream  :                    scrub 0, verify 0 | scrub 1, verify 1 | ...
//   reaper  :                                                                  unmap + release
//
// * chunk sizes grow 1, 2, 4, 8, 16, 16, ... GiB so the first stores are issued ~1 ms into the
//   call; a few large chunks keep the driver's per-mapping costs (unmap!) down;
// * with bytes == 0 the last CCM_VMM_TAIL_MB (256) of free HBM (the "tail zone") are taken in
//   32 MiB -> 2 MiB granules until the driver says out-of-memory, so the region is everything
//   the context can reach, not "free minus a safety margin";
// * every chunk is read back right after it was zeroed, while the host is still mapping the
//   next one (the GPU would otherwise idle: mapping 16 GiB takes as long as scrubbing it);
// * the verdict is returned as soon as the 8-byte count is on the host; cuMemUnmap/cuMemRelease
//   (0.33 ms/GiB, serialised node-wide by the driver) run on the engine's reaper thread.
static int scrub_verify_pipelined(ScrubEngine* e, uint64_t bytes, uint64_t inject, bool node_fanout, ccm_scrub_result* r) {
  const VmmApi& api = vmm();
  if (!api.ok) return CCM_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  if (!e->segs.empty()) { set_error("arena already held on CUDA device %d", e->ordinal); return CCM_ERR_STATE; }
  r->ms_release_wait = e->join_reaper();
  CCM_CUDA(cudaSetDevice(e->ordinal));
  size_t fr = 0, tot = 0;
  CCM_CUDA(cudaMemGetInfo(&fr, &tot));
  r->device_total_bytes = tot;
  r->device_free_before = fr;
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = e->ordinal;
  size_t gran = 0;
  if (api.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
    return CCM_ERR_UNSUPPORTED;
  const bool want_max = (bytes == 0);
  // physical memory comes in `gran` units; an exact request is rounded UP and the
  // surplus bytes are scrubbed too (harmless, and they are ours).
  const uint64_t va_bytes = want_max ? (uint64_t)fr / gran * gran : (bytes + gran - 1) / gran * gran;
  if (va_bytes == 0) { set_error("no free HBM to scrub (free=%zu)", fr); return CCM_ERR_NOMEM; }
  // tail zone: where running out of memory is expected and simply ends the region
  uint64_t tail_zone = want_max ? env_u64("CCM_VMM_TAIL_MB", 256) * kMiB / gran * gran : 0;
  if (tail_zone > va_bytes) tail_zone = va_bytes;
  uint64_t main_end = va_bytes - tail_zone;
  uint64_t max_chunk = env_u64("CCM_VMM_CHUNK_MB", 16384) * kMiB / gran * gran;
  if (max_chunk < gran) max_chunk = gran;
  uint64_t chunk = env_u64("CCM_VMM_FIRST_CHUNK_MB", 1024) * kMiB / gran * gran;
  if (chunk < gran) chunk = gran;
  if (chunk > max_chunk) chunk = max_chunk;
  uint64_t tail_chunk = 32 * kMiB / gran * gran;
  if (tail_chunk < gran) tail_chunk = gran;
  const bool interleave = env_u64("CCM_INTERLEAVE_VERIFY", 1) != 0;
  // Map the whole range before the first launch (no overlap of mapping and kernels)?  One GPU per
  // process: no — pipelining wins (verdict after 55 ms instead of ~75 ms).  Several GPUs driven from
  // ONE process (ccm_scrub_verify_many): yes — VMM calls interleaved with launches on 8 devices of one
  // process take 0.7 s to the verdict, mapping first takes 0.22 s (profiles/r2_node_gate_modes_8gpu.log).
  // CCM_MAP_FIRST=0/1 overrides.
  const bool map_first = env_u64("CCM_MAP_FIRST", node_fanout ? 1 : 0) != 0;

  VmmMapping m;
  if (api.AddressReserve(&m.base, va_bytes, 0, 0, 0) != CUDA_SUCCESS) return CCM_ERR_UNSUPPORTED;
  m.va_bytes = va_bytes;
  cudaStream_t st = e->stream;
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  int rc = CCM_OK;
  double host_acquire = 0;
  int timed = 0;  // chunks with an event triple (the first kMaxPipeChunks; a badly fragmented HBM may produce more)

  // enqueue scrub (+ read-back) of [off, off+n) with CUDA events around each kernel
  auto enqueue = [&](uint64_t off, uint64_t n) -> int {
    cudaEvent_t* evs = timed < kMaxPipeChunks ? e->pipe_ev[timed] : nullptr;
    if (evs) cudaEventRecord(evs[0], st);
    int rc2 = scrub_range(e, (void*)(m.base + off), n, CCM_SCRUB_AUTO, nullptr, st);
    if (evs) cudaEventRecord(evs[1], st);
    if (inject) {  // fault drill: dirty bytes the read-back below MUST find
      const uint64_t cand[4] = {off, off + 17 < off + n ? off + 17 : off, off + n / 2 + 3 < off + n ? off + n / 2 + 3 : off, off + n - 1};
      for (int i = 0; i < 4 && inject; ++i) {
        bool dup = false;
        for (int j = 0; j < i; ++j) dup |= cand[j] == cand[i];
        if (dup) continue;
        cudaMemsetAsync((void*)(m.base + cand[i]), 0xA5, 1, st);
        --inject;
      }
    }
    if (rc2 == CCM_OK && interleave) rc2 = verify_range(e, (const void*)(m.base + off), n, CCM_VERIFY_AUTO, nullptr, st);
    if (evs) { cudaEventRecord(evs[2], st); ++timed; }
    return rc2;
  };

  if (cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), st) != cudaSuccess) rc = CCM_ERR_CUDA;
  cudaEventRecord(e->ev[0], st);
  uint64_t off = 0, tail_start = 0;
  bool in_tail = false;
  while (off < va_bytes && rc == CCM_OK) {
    if (!in_tail && off >= main_end) { in_tail = true; tail_start = off; }
    const uint64_t limit = in_tail ? va_bytes : main_end;
    uint64_t& ck = in_tail ? tail_chunk : chunk;
    const uint64_t n = limit - off < ck ? limit - off : ck;
    const double t0 = now_ms();
    CUmemGenericAllocationHandle h;
    CUresult cr = api.Create(&h, n, &prop, 0);
    if (cr == CUDA_ERROR_OUT_OF_MEMORY) {
      const uint64_t floor_bytes = in_tail ? gran : (64 * kMiB > gran ? 64 * kMiB : gran);
      if (ck > floor_bytes) { ck = (ck / 2) / gran * gran; if (ck < gran) ck = gran; continue; }  // fragmented / less free than reported
      if (want_max && !in_tail) { main_end = off; continue; }  // less free than reported: finish in small granules
      if (want_max && off > 0) break;  // take what there is
    }
    if (cr != CUDA_SUCCESS) {
      set_error("cuMemCreate(%llu) failed: %d", (unsigned long long)n, (int)cr);
      rc = cr == CUDA_ERROR_OUT_OF_MEMORY ? CCM_ERR_NOMEM : CCM_ERR_CUDA;
      break;
    }
    CUresult mr = api.Map(m.base + off, n, 0, h, 0);
    if (mr == CUDA_SUCCESS) {
      mr = api.SetAccess(m.base + off, n, &acc, 1);
      if (mr != CUDA_SUCCESS) api.Unmap(m.base + off, n);
    }
    if (mr != CUDA_SUCCESS) {
      api.Release(h);
      if (mr == CUDA_ERROR_OUT_OF_MEMORY && in_tail) break;  // no room left for page tables: the region ends here
      set_error("cuMemMap/cuMemSetAccess failed: %d", (int)mr);
      rc = CCM_ERR_CUDA;
      break;
    }
    host_acquire += now_ms() - t0;
    m.handles.push_back(h);
    m.mapped = off + n;
    // tail-zone granules are scrubbed with ONE launch pair once the zone is mapped
    if (!in_tail && !map_first) rc = enqueue(off, n);
    off += n;
    if (!in_tail && chunk < max_chunk) { chunk *= 2; if (chunk > max_chunk) chunk = max_chunk; }
  }
  if (rc == CCM_OK && map_first) { if (m.mapped) rc = enqueue(0, m.mapped); }
  else if (rc == CCM_OK && in_tail && m.mapped > tail_start) rc = enqueue(tail_start, m.mapped - tail_start);
  if (rc == CCM_OK && !interleave) {
    cudaEventRecord(e->ev[1], st);
    rc = verify_range(e, (const void*)m.base, m.mapped, CCM_VERIFY_AUTO, nullptr, st);
  }
  cudaEventRecord(e->ev[2], st);
  if (rc == CCM_OK &&
      cudaMemcpyAsync(e->h_counter, e->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) != cudaSuccess)
    rc = CCM_ERR_CUDA;
  cudaError_t serr = cudaStreamSynchronize(st);  // nothing may still touch the mapping below
  if (rc == CCM_OK && serr != cudaSuccess) { set_error("stream sync failed: %s", cudaGetErrorString(serr)); rc = CCM_ERR_CUDA; }

  const uint64_t mapped = m.mapped;
  const int nchunks = (int)m.handles.size();
  if (rc == CCM_OK) {
    float span = 0, ms_s = 0, ms_v = 0;
    cudaEventElapsedTime(&span, e->ev[0], e->ev[2]);
    for (int i = 0; i < timed; ++i) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, e->pipe_ev[i][0], e->pipe_ev[i][1]);
      cudaEventElapsedTime(&b, e->pipe_ev[i][1], e->pipe_ev[i][2]);
      ms_s += a;
      ms_v += b;
    }
    if (!interleave) cudaEventElapsedTime(&ms_v, e->ev[1], e->ev[2]);
    r->bytes_scrubbed = mapped;
    r->bytes_unreached = want_max && (uint64_t)fr > mapped ? (uint64_t)fr - mapped : 0;
    r->segments = nchunks;
    r->ms_acquire = host_acquire;
    r->ms_scrub = ms_s;
    r->ms_verify = ms_v;
    r->ms_gpu_span = span;
    r->nonzero_bytes = (uint64_t)*e->h_counter;
  }
  // ---- give the HBM back: on the reaper (default) or right here -----------------------
  if (rc == CCM_OK && async_release_enabled()) {
    r->release_deferred = 1;
    r->ms_release = 0;
    const int ordinal = e->ordinal;
    e->reaper = std::thread([e, ordinal, mm = std::move(m)]() mutable {
      cudaSetDevice(ordinal);
      e->reaper_ms = give_back(vmm(), mm);  // read only after join()
    });
  } else {
    r->ms_release = give_back(api, m);
  }
  if (rc != CCM_OK) return rc;
  if (!want_max && mapped < bytes) { set_error("mapped only %llu of %llu bytes", (unsigned long long)mapped, (unsigned long long)bytes); return CCM_ERR_NOMEM; }
  return CCM_OK;
}

int engine_release_wait(ScrubEngine* e, double* ms_release, double* ms_waited) {
  std::lock_guard<std::mutex> g(e->mu);
  const double w = e->join_reaper();
  if (ms_release) *ms_release = e->reaper_ms;
  if (ms_waited) *ms_waited = w;
  return CCM_OK;
}

// --------------------------------------------------------------- product call
int engine_scrub_verify(ScrubEngine* e, uint64_t bytes, uint64_t inject, bool node_fanout, ccm_scrub_result* out) {
  const double t0 = now_ms();
  ccm_scrub_result r;
  memset(&r, 0, sizeof r);
  r.bytes_requested = bytes;
  r.sm_count = e->sm_count;
  r.scrub_variant = resolve_scrub_variant(CCM_SCRUB_AUTO);
  r.verify_variant = resolve_verify_variant(CCM_VERIFY_AUTO);
  int rc = CCM_ERR_UNSUPPORTED;
  if (env_u64("CCM_PIPELINE", 1) != 0) rc = scrub_verify_pipelined(e, bytes, inject, node_fanout, &r);
  if (rc == CCM_ERR_UNSUPPORTED) {
    // plain path: one cudaMalloc'ed arena, whole-arena scrub, whole-arena verify
    ccm_arena_info ai;
    rc = engine_arena_acquire(e, bytes, &ai);
    if (rc == CCM_OK) {
      r.ms_acquire = ai.ms_acquire;
      r.bytes_scrubbed = ai.bytes;
      r.device_total_bytes = ai.device_total_bytes;
      r.device_free_before = ai.device_free_before;
      r.bytes_unreached = bytes == 0 && ai.device_free_before > ai.bytes ? ai.device_free_before - ai.bytes : 0;
      r.segments = ai.segments;
      float ms_s = 0, ms_v = 0;
      uint64_t nz = 0;
      rc = engine_arena_scrub(e, CCM_SCRUB_AUTO, nullptr, nullptr, &ms_s);
      if (rc == CCM_OK) rc = engine_arena_verify(e, CCM_VERIFY_AUTO, nullptr, nullptr, &nz, &ms_v);
      r.ms_scrub = ms_s;
      r.ms_verify = ms_v;
      r.nonzero_bytes = nz;
      int rc2 = engine_arena_release(e, &r.ms_release);
      if (rc == CCM_OK) rc = rc2;
    }
  }
  if (rc == CCM_OK && r.nonzero_bytes != 0) {
    set_error("scrub verify found %llu non-zero bytes on CUDA device %d", (unsigned long long)r.nonzero_bytes, e->ordinal);
    rc = CCM_ERR_DIRTY;
  }
  r.ms_total = now_ms() - t0;
  r.sm_count = e->sm_count;
  r.status = rc;
  if (out) *out = r;
  return rc;
}

// ----------------------------------------------------------------- raw regions
int engine_region_scrub(ScrubEngine* e, void* dptr, uint64_t bytes, int variant, const ccm_launch_cfg* cfg,
                        void* stream, float* ms) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->pick(stream);
  if (ms) CCM_CUDA(cudaEventRecord(e->ev[0], st));
  int rc = scrub_range(e, dptr, bytes, variant, cfg, st);
  if (rc) return rc;
  if (ms) {
    CCM_CUDA(cudaEventRecord(e->ev[1], st));
    CCM_CUDA(cudaEventSynchronize(e->ev[1]));
    CCM_CUDA(cudaEventElapsedTime(ms, e->ev[0], e->ev[1]));
  }
  return CCM_OK;
}

int engine_region_verify(ScrubEngine* e, const void* dptr, uint64_t bytes, int variant,
                         const ccm_launch_cfg* cfg, void* stream, uint64_t* nonzero, float* ms) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->pick(stream);
  CCM_CUDA(cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), st));
  CCM_CUDA(cudaEventRecord(e->ev[0], st));
  int rc = verify_range(e, dptr, bytes, variant, cfg, st);
  if (rc) return rc;
  CCM_CUDA(cudaEventRecord(e->ev[1], st));
  rc = fetch_count_locked(e, st, nonzero);
  if (rc) return rc;
  if (ms) CCM_CUDA(cudaEventElapsedTime(ms, e->ev[0], e->ev[1]));
  return CCM_OK;
}

int engine_host_roundtrip(ScrubEngine* e, void* host_buf, uint64_t bytes, uint64_t dev_offset,
                          int sv, int vv, uint64_t* pre, uint64_t* post) {
  std::lock_guard<std::mutex> g(e->mu);
  CCM_ENSURE_READY(e);
  CCM_CUDA(cudaSetDevice(e->ordinal));
  cudaStream_t st = e->stream;
  uint8_t* d = nullptr;
  // dev_offset shifts the region off the allocation's natural alignment so ragged
  // heads are exercised; the guard bytes around it must stay untouched.
  const uint64_t guard = 256;
  const uint64_t alloc = guard + dev_offset + bytes + guard;
  CCM_CUDA(cudaMalloc(&d, alloc));
  int rc = CCM_OK;
  auto fail = [&](cudaError_t err, const char* what) {
    set_error("%s failed: %s", what, cudaGetErrorString(err));
    rc = CCM_ERR_CUDA;
  };
  cudaError_t err;
  uint8_t* region = d + guard + dev_offset;
  std::vector<uint8_t> g0(guard + dev_offset, 0xEE), g1(guard, 0xEE), chk;
  do {
    if ((err = cudaMemsetAsync(d, 0xEE, alloc, st)) != cudaSuccess) { fail(err, "memset guard"); break; }
    if (bytes && (err = cudaMemcpyAsync(region, host_buf, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) { fail(err, "H2D"); break; }
    if (pre) {
      if ((err = cudaMemsetAsync(e->d_counter, 0, 8, st)) != cudaSuccess) { fail(err, "counter"); break; }
      if ((rc = verify_range(e, region, bytes, vv, nullptr, st))) break;
      if ((rc = fetch_count_locked(e, st, pre))) break;
    }
    if ((rc = scrub_range(e, region, bytes, sv, nullptr, st))) break;
    if ((err = cudaMemsetAsync(e->d_counter, 0, 8, st)) != cudaSuccess) { fail(err, "counter"); break; }
    if ((rc = verify_range(e, region, bytes, vv, nullptr, st))) break;
    if ((rc = fetch_count_locked(e, st, post))) break;
    if (bytes && (err = cudaMemcpy(host_buf, region, bytes, cudaMemcpyDeviceToHost)) != cudaSuccess) { fail(err, "D2H"); break; }
    // guards: the scrub must not have written outside [region, region+bytes)
    chk.resize(g0.size());
    if ((err = cudaMemcpy(chk.data(), d, chk.size(), cudaMemcpyDeviceToHost)) != cudaSuccess) { fail(err, "D2H guard"); break; }
    if (memcmp(chk.data(), g0.data(), g0.size()) != 0) { set_error("scrub wrote below the region"); rc = CCM_ERR_DIRTY; break; }
    chk.resize(g1.size());
    if ((err = cudaMemcpy(chk.data(), region + bytes, chk.size(), cudaMemcpyDeviceToHost)) != cudaSuccess) { fail(err, "D2H guard"); break; }
    if (memcmp(chk.data(), g1.data(), g1.size()) != 0) { set_error("scrub wrote past the region"); rc = CCM_ERR_DIRTY; break; }
  } while (0);
  cudaFree(d);
  return rc;
}

}  // namespace ccm
