// vmm_probe.cu — where does the cold scrub gate's driver time go, and does it serialise
// across GPUs?  (VERDICT r1 "next round" item 1: the cold call is 93 % cuMemCreate / cuMemMap /
// cuMemSetAccess / cuMemUnmap / cuMemRelease; kernels are 50 ms of 120-800 ms.)
//
// Every driver call is logged with CLOCK_MONOTONIC timestamps (system-wide, so the logs of
// several PROCESSES can be merged on one time axis).
//
//   vmm_probe single [gib]      one GPU: context creation, create/map/access per chunk size and
//                               granularity, unmap/release, settle times, creator threads,
//                               cudaDeviceReset with everything mapped, re-create after reset
//   vmm_probe multi N [mode]    N GPUs at once, mode = procs (fork, one process per GPU) or
//                               threads (one process): per-call timestamps -> how much of the
//                               driver work overlaps between GPUs
//
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a vmm_probe.cu -o _bin/vmm_probe -lcuda
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include <cuda.h>

static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static double g_t0 = 0;

#define DR(x)                                                                 \
  do {                                                                        \
    CUresult e__ = (x);                                                       \
    if (e__ != CUDA_SUCCESS) {                                                \
      const char* s__ = nullptr;                                              \
      cuGetErrorString(e__, &s__);                                            \
      printf("ERR %s: %s (line %d)\n", #x, s__ ? s__ : "?", __LINE__);        \
      fflush(stdout);                                                         \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

__global__ void touch(uint4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(0, 0, 0, 0);
}

struct Call { double t0, t1; int gpu; const char* what; size_t bytes; };

struct Mapping {
  CUdeviceptr base = 0;
  size_t va = 0, mapped = 0;
  std::vector<CUmemGenericAllocationHandle> hs;
  std::vector<size_t> sz;
};

static CUmemAllocationProp prop_for(int dev) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof p);
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  return p;
}

// create + map + set-access chunk by chunk (what the library does); logs every call
static void acquire(int dev, size_t total, size_t chunk, Mapping* m, std::vector<Call>* log, bool access_per_chunk = true) {
  CUmemAllocationProp prop = prop_for(dev);
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  double a = now_ms();
  DR(cuMemAddressReserve(&m->base, total, 0, 0, 0));
  log->push_back({a, now_ms(), dev, "reserve", total});
  m->va = total;
  size_t off = 0;
  while (off < total) {
    size_t n = std::min(chunk, total - off);
    CUmemGenericAllocationHandle h;
    a = now_ms();
    CUresult r = cuMemCreate(&h, n, &prop, 0);
    log->push_back({a, now_ms(), dev, "create", n});
    if (r != CUDA_SUCCESS) { printf("gpu%d cuMemCreate(%zu MiB) failed at %.1f GiB (%d)\n", dev, n >> 20, off / 1073741824.0, (int)r); break; }
    a = now_ms();
    DR(cuMemMap(m->base + off, n, 0, h, 0));
    log->push_back({a, now_ms(), dev, "map", n});
    if (access_per_chunk) {
      a = now_ms();
      DR(cuMemSetAccess(m->base + off, n, &acc, 1));
      log->push_back({a, now_ms(), dev, "access", n});
    }
    m->hs.push_back(h);
    m->sz.push_back(n);
    off += n;
  }
  m->mapped = off;
  if (!access_per_chunk) {
    a = now_ms();
    DR(cuMemSetAccess(m->base, off, &acc, 1));
    log->push_back({a, now_ms(), dev, "access", off});
  }
}

static void release(int dev, Mapping* m, std::vector<Call>* log, bool per_chunk_unmap = false) {
  double a;
  if (!per_chunk_unmap) {
    a = now_ms();
    DR(cuMemUnmap(m->base, m->mapped));
    log->push_back({a, now_ms(), dev, "unmap", m->mapped});
  }
  size_t off = 0;
  for (size_t i = 0; i < m->hs.size(); ++i) {
    if (per_chunk_unmap) {
      a = now_ms();
      DR(cuMemUnmap(m->base + off, m->sz[i]));
      log->push_back({a, now_ms(), dev, "unmap", m->sz[i]});
    }
    a = now_ms();
    DR(cuMemRelease(m->hs[i]));
    log->push_back({a, now_ms(), dev, "release", m->sz[i]});
    off += m->sz[i];
  }
  a = now_ms();
  DR(cuMemAddressFree(m->base, m->va));
  log->push_back({a, now_ms(), dev, "addrfree", m->va});
  *m = Mapping();
}

static double sum_of(const std::vector<Call>& log, const char* what, size_t from = 0) {
  double s = 0;
  for (size_t i = from; i < log.size(); ++i)
    if (!strcmp(log[i].what, what)) s += log[i].t1 - log[i].t0;
  return s;
}
static double max_of(const std::vector<Call>& log, const char* what, size_t from = 0) {
  double s = 0;
  for (size_t i = from; i < log.size(); ++i)
    if (!strcmp(log[i].what, what)) s = std::max(s, log[i].t1 - log[i].t0);
  return s;
}

static size_t usable(size_t gran) {
  size_t fr = 0, tot = 0;
  DR(cuMemGetInfo(&fr, &tot));
  size_t reserve = 256ull << 20;
  return fr > reserve ? (fr - reserve) / gran * gran : 0;
}

static void touch_all(const Mapping& m) {
  touch<<<1184, 256>>>((uint4*)m.base, m.mapped / 16);
  DR(cuCtxSynchronize());
}

static void cycle(const char* tag, int dev, size_t total, size_t chunk, bool touch_it, bool per_chunk_unmap = false) {
  std::vector<Call> log;
  Mapping m;
  double a = now_ms();
  acquire(dev, total, chunk, &m, &log);
  double b = now_ms();
  double t_touch = 0;
  if (touch_it) { double c = now_ms(); touch_all(m); t_touch = now_ms() - c; }
  double c = now_ms();
  size_t rel_from = log.size();
  release(dev, &m, &log, per_chunk_unmap);
  double d = now_ms();
  printf("%-28s chunk %6zu MiB n=%3zu | acquire %7.1f ms (create %7.1f max %6.1f, map %5.1f, access %6.1f) touch %5.1f | "
         "release %7.1f ms (unmap %6.1f, release %6.1f max %6.1f)\n",
         tag, chunk >> 20, (total + chunk - 1) / chunk, b - a, sum_of(log, "create"), max_of(log, "create"), sum_of(log, "map"),
         sum_of(log, "access"), t_touch, d - c, sum_of(log, "unmap", rel_from), sum_of(log, "release", rel_from),
         max_of(log, "release", rel_from));
  fflush(stdout);
}

static int run_single(size_t cap_gib) {
  DR(cuInit(0));
  CUdevice dev;
  DR(cuDeviceGet(&dev, 0));
  CUcontext ctx;
  double a = now_ms();
  DR(cuDevicePrimaryCtxRetain(&ctx, dev));
  DR(cuCtxSetCurrent(ctx));
  printf("primary context create: %.1f ms\n", now_ms() - a);
  CUmemAllocationProp prop = prop_for(0);
  size_t gmin = 0, grec = 0;
  DR(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  DR(cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  size_t fr = 0, tot = 0;
  DR(cuMemGetInfo(&fr, &tot));
  printf("granularity min %zu recommended %zu | free %.2f GiB total %.2f GiB\n", gmin, grec, fr / 1073741824.0, tot / 1073741824.0);
  const size_t G = 1ull << 30;
  size_t total = usable(std::max(gmin, grec));
  if (cap_gib && cap_gib * G < total) total = cap_gib * G;

  // ---- 1. chunk sizes, first cycle in the process is the "fresh" one ------------------
  cycle("fresh 16GiB", 0, total, 16 * G, true);
  cycle("again 16GiB", 0, total, 16 * G, true);
  cycle("again 16GiB no-touch", 0, total, 16 * G, false);
  cycle("again 16GiB no-touch", 0, total, 16 * G, false);
  cycle("whole", 0, total, total, true);
  cycle("whole", 0, total, total, true);
  cycle("4GiB", 0, total, 4 * G, true);
  cycle("4GiB", 0, total, 4 * G, true);
  cycle("1GiB", 0, total, 1 * G, true);
  cycle("1GiB", 0, total, 1 * G, true);
  cycle("64GiB", 0, total, 64 * G, true);
  cycle("64GiB", 0, total, 64 * G, true);
  cycle("16GiB per-chunk-unmap", 0, total, 16 * G, true, true);
  cycle("16GiB per-chunk-unmap", 0, total, 16 * G, true, true);

  // ---- 2. settle: release, wait s, create again ---------------------------------------
  for (double settle : {0.0, 0.02, 0.05, 0.1, 0.2, 0.5, 1.0, 2.0, 0.0}) {
    usleep((useconds_t)(settle * 1e6));
    char tag[64];
    snprintf(tag, sizeof tag, "settle %.2fs", settle);
    cycle(tag, 0, total, 16 * G, true);
  }

  // ---- 3. creator threads (each thread creates+maps+accesses its own chunks) ----------
  for (int T : {1, 2, 4, 8}) {
    for (int rep = 0; rep < 2; ++rep) {
      CUdeviceptr base;
      DR(cuMemAddressReserve(&base, total, 0, 0, 0));
      const size_t chunk = 8 * G;
      const size_t nchunks = (total + chunk - 1) / chunk;
      std::vector<CUmemGenericAllocationHandle> hs(nchunks);
      std::vector<char> ok(nchunks, 0);
      std::atomic<size_t> next{0};
      double t0 = now_ms();
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t)
        th.emplace_back([&] {
          cuCtxSetCurrent(ctx);
          CUmemAccessDesc acc;
          memset(&acc, 0, sizeof acc);
          acc.location = prop.location;
          acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
          for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= nchunks) break;
            size_t off = i * chunk, n = std::min(chunk, total - off);
            if (cuMemCreate(&hs[i], n, &prop, 0) != CUDA_SUCCESS) continue;
            cuMemMap(base + off, n, 0, hs[i], 0);
            cuMemSetAccess(base + off, n, &acc, 1);
            ok[i] = 1;
          }
        });
      for (auto& t : th) t.join();
      double t1 = now_ms();
      // release with T threads too
      DR(cuCtxSynchronize());
      next = 0;
      th.clear();
      for (int t = 0; t < T; ++t)
        th.emplace_back([&] {
          cuCtxSetCurrent(ctx);
          for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= nchunks) break;
            if (!ok[i]) continue;
            size_t off = i * chunk, n = std::min(chunk, total - off);
            cuMemUnmap(base + off, n);
            cuMemRelease(hs[i]);
          }
        });
      for (auto& t : th) t.join();
      double t2 = now_ms();
      DR(cuMemAddressFree(base, total));
      printf("creator threads T=%d (8 GiB chunks): acquire %.1f ms, release %.1f ms\n", T, t1 - t0, t2 - t1);
      fflush(stdout);
    }
  }

  // ---- 4. release while a kernel runs on OTHER memory; release early by refcount --------
  {
    // cuMemRelease right after cuMemMap: the handle stays alive through its mapping, the
    // free then happens inside cuMemUnmap.  Does that move the cost?
    std::vector<Call> log;
    Mapping m;
    acquire(0, total, 16 * G, &m, &log);
    double a0 = now_ms();
    for (auto h : m.hs) DR(cuMemRelease(h));
    double a1 = now_ms();
    touch_all(m);
    double a2 = now_ms();
    DR(cuMemUnmap(m.base, m.mapped));
    double a3 = now_ms();
    DR(cuMemAddressFree(m.base, m.va));
    printf("early-release: release-after-map %.1f ms, touch %.1f ms, unmap (does the free) %.1f ms\n", a1 - a0, a2 - a1, a3 - a2);
    fflush(stdout);
  }

  // ---- 5. cudaDeviceReset-style teardown with everything mapped -------------------------
  {
    std::vector<Call> log;
    Mapping m;
    acquire(0, total, 16 * G, &m, &log);
    touch_all(m);
    double a0 = now_ms();
    DR(cuDevicePrimaryCtxReset(dev));
    double a1 = now_ms();
    printf("cuDevicePrimaryCtxReset with %.1f GiB still mapped: %.1f ms\n", m.mapped / 1073741824.0, a1 - a0);
    // VMM allocations are process-wide, not context-owned: the reset does NOT free them
    size_t fr2 = 0, tot2 = 0;
    a0 = now_ms();
    DR(cuDevicePrimaryCtxRetain(&ctx, dev));
    DR(cuCtxSetCurrent(ctx));
    a1 = now_ms();
    printf("  new primary context: %.1f ms\n", a1 - a0);
    DR(cuMemGetInfo(&fr2, &tot2));
    printf("  free after reset, mapping still held: %.2f GiB\n", fr2 / 1073741824.0);
    double r0 = now_ms();
    std::vector<Call> log2;
    release(0, &m, &log2);
    double r1 = now_ms();
    DR(cuMemGetInfo(&fr2, &tot2));
    printf("  unmap+release after the reset: %.1f ms -> free %.2f GiB\n", r1 - r0, fr2 / 1073741824.0);
    fflush(stdout);
    cycle("after ctx reset 16GiB", 0, std::min(total, usable(gmin)), 16 * G, true);
    cycle("again", 0, std::min(total, usable(gmin)), 16 * G, true);
  }
  // explicit teardown then reset, for comparison
  {
    std::vector<Call> log;
    Mapping m;
    size_t t2 = std::min(total, usable(gmin));
    acquire(0, t2, 16 * G, &m, &log);
    touch_all(m);
    double a0 = now_ms();
    release(0, &m, &log);
    double a1 = now_ms();
    DR(cuDevicePrimaryCtxReset(dev));
    double a2 = now_ms();
    printf("explicit release %.1f ms + cuDevicePrimaryCtxReset %.1f ms\n", a1 - a0, a2 - a1);
  }
  return 0;
}

// ------------------------------------------------------------------------------ multi
struct Shared {
  std::atomic<int> arrived[16];
  int n;
};
static void barrier(Shared* sh, int phase) {
  sh->arrived[phase].fetch_add(1);
  while (sh->arrived[phase].load() < sh->n) usleep(200);
}

static void worker(int dev, Shared* sh, std::vector<Call>* log, size_t cap_gib, bool stagger) {
  CUdevice d;
  DR(cuDeviceGet(&d, dev));
  CUcontext ctx;
  barrier(sh, 0);
  double a = now_ms();
  DR(cuDevicePrimaryCtxRetain(&ctx, d));
  DR(cuCtxSetCurrent(ctx));
  log->push_back({a, now_ms(), dev, "ctx", 0});
  CUmemAllocationProp prop = prop_for(dev);
  size_t gmin = 0;
  DR(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  size_t total = usable(gmin);
  if (cap_gib && (cap_gib << 30) < total) total = cap_gib << 30;
  for (int rep = 0; rep < 3; ++rep) {
    barrier(sh, 1 + rep * 3);
    if (stagger) usleep(dev * 30000);
    Mapping m;
    a = now_ms();
    acquire(dev, total, 16ull << 30, &m, log);
    log->push_back({a, now_ms(), dev, "ACQUIRE", total});
    a = now_ms();
    touch<<<1184, 256>>>((uint4*)m.base, m.mapped / 16);
    DR(cuCtxSynchronize());
    log->push_back({a, now_ms(), dev, "touch", m.mapped});
    barrier(sh, 2 + rep * 3);
    a = now_ms();
    release(dev, &m, log);
    log->push_back({a, now_ms(), dev, "RELEASE", total});
    barrier(sh, 3 + rep * 3);
  }
  a = now_ms();
  DR(cuDevicePrimaryCtxRelease(d));
  log->push_back({a, now_ms(), dev, "ctxrelease", 0});
}

static void dump(const std::vector<Call>& log) {
  for (auto& c : log)
    printf("CALL gpu%d %-10s %10.2f %10.2f %8.2f ms %6zu MiB\n", c.gpu, c.what, c.t0 - g_t0, c.t1 - g_t0, c.t1 - c.t0, c.bytes >> 20);
  fflush(stdout);
}

static int run_multi(int n, const char* mode, size_t cap_gib) {
  Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  memset(sh, 0, sizeof *sh);
  sh->n = n;
  const bool stagger = strstr(mode, "stagger") != nullptr;
  if (!strncmp(mode, "procs", 5)) {
    for (int i = 0; i < n; ++i) {
      pid_t p = fork();
      if (p == 0) {
        DR(cuInit(0));
        std::vector<Call> log;
        worker(i, sh, &log, cap_gib, stagger);
        dump(log);
        _exit(0);
      }
    }
    int st;
    while (wait(&st) > 0) {}
  } else {
    DR(cuInit(0));
    std::vector<std::vector<Call>> logs(n);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) th.emplace_back([&, i] { worker(i, sh, &logs[i], cap_gib, stagger); });
    for (auto& t : th) t.join();
    for (auto& l : logs) dump(l);
  }
  return 0;
}

// ctx: how long do primary-context creation and reset take (warm process), and do environment knobs that
// shrink a context (CUDA_DEVICE_MAX_CONNECTIONS=1 -> fewer channels) make them cheaper?
static int run_ctx(int reps) {
  DR(cuInit(0));
  CUdevice dev;
  DR(cuDeviceGet(&dev, 0));
  const char* mc = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
  printf("CUDA_DEVICE_MAX_CONNECTIONS=%s\n", mc ? mc : "(default 8)");
  for (int i = 0; i < reps; ++i) {
    CUcontext ctx;
    double a = now_ms();
    DR(cuDevicePrimaryCtxRetain(&ctx, dev));
    DR(cuCtxSetCurrent(ctx));
    double b = now_ms();
    touch<<<148, 256>>>(nullptr, 0);  // first launch: module load
    DR(cuCtxSynchronize());
    double c = now_ms();
    size_t fr = 0, tot = 0;
    DR(cuMemGetInfo(&fr, &tot));
    DR(cuDevicePrimaryCtxRelease(dev));
    DR(cuDevicePrimaryCtxReset(dev));
    double d = now_ms();
    printf("ctx rep %d: create %.1f ms, first launch %.1f ms, release+reset %.1f ms, context footprint %.0f MiB\n", i, b - a, c - b,
           d - c, (tot - fr) / 1048576.0);
  }
  return 0;
}

int main(int argc, char** argv) {
  g_t0 = now_ms();
  setvbuf(stdout, nullptr, _IOLBF, 0);
  if (argc >= 2 && !strcmp(argv[1], "ctx")) return run_ctx(argc >= 3 ? atoi(argv[2]) : 5);
  if (argc >= 2 && !strcmp(argv[1], "single")) return run_single(argc >= 3 ? strtoull(argv[2], nullptr, 10) : 0);
  if (argc >= 3 && !strcmp(argv[1], "multi")) {
    const char* g0 = getenv("VMM_PROBE_T0");
    if (g0) g_t0 = atof(g0);
    return run_multi(atoi(argv[2]), argc >= 4 ? argv[3] : "procs", argc >= 5 ? strtoull(argv[4], nullptr, 10) : 0);
  }
  printf("usage: vmm_probe single [gib] | multi N [procs|threads|procs-stagger] [gib] | ctx [reps]\n");
  return 2;
}
