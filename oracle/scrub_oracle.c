/*
 * scrub_oracle.c — CPU restatement of the HBM scrub-and-verify contract.
 *
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this.  The product
 * (k8s_cc_manager_b200, libccm.so) never links, loads or calls it.
 *
 * What it restates: the reference has NO scrub (reference
 * main.py:502-529 ends at verify-mode), so there is no reference arithmetic to
 * follow.  The contract is the one BASELINE.json:north_star states and SURVEY.md
 * §8a row S / §8c spell out:
 *     scrub : every byte of the region becomes 0x00               (memset)
 *     verify: nonzero_bytes = #{ i : buf[i] != 0 }, exact, as u64 (byte loop)
 * Parity status: pinned against the independent numpy statement
 * (oracle/scrub_oracle.py: np.count_nonzero) and the known-answer vectors in
 * tests/golden/scrub_vectors.json (k = 0, 1, 7, 4096 injected bytes incl. first
 * byte, last byte and a non-16-aligned offset — SURVEY.md §8d); there is no
 * reference-side golden for it because the reference has no such function.
 *
 * The seeded sparse test pattern (ccm_oracle_pattern_word) restates, independently,
 * the formula the device fill kernel uses (csrc/scrub_kernels.cuh: pattern_word) so
 * the GPU count can be checked at sizes the host never materialises.
 *
 * The *_mt entry points split the buffer over pthreads: they are the "serial
 * CPU-driven path" baseline (BASELINE.md B1 spirit) timed by bench.py.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

API void ccm_oracle_scrub(uint8_t* buf, uint64_t n) { memset(buf, 0, (size_t)n); }

/* Deliberately the plainest statement: one compare per byte.  target_clones lets
 * gcc emit an AVX2 body next to the baseline one and pick at load time, so the CPU
 * baseline is not handicapped on the GPU box's host without assuming its ISA. */
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("arch=x86-64-v4", "avx2", "default")))
#endif
API uint64_t ccm_oracle_count_nonzero(const uint8_t* buf, uint64_t n) {
  uint64_t c = 0;
  for (uint64_t i = 0; i < n; ++i) c += (buf[i] != 0);
  return c;
}

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* 8-byte little-endian word j of the seeded pattern:
 *   r = splitmix64(seed + j); 7 words out of 8 are zero ((r & 7) != 0);
 *   otherwise byte b is kept iff bit (8+b) of r is set, and a kept byte is
 *   ((r >> 3) | 0x01..01) >> 8b, i.e. never zero.                                  */
API uint64_t ccm_oracle_pattern_word(uint64_t seed, uint64_t j) {
  uint64_t r = splitmix64(seed + j);
  if ((r & 7) != 0) return 0;
  uint64_t keep = 0;
  for (int b = 0; b < 8; ++b)
    if ((r >> (8 + b)) & 1) keep |= 0xFFull << (8 * b);
  return ((r >> 3) | 0x0101010101010101ull) & keep;
}

API void ccm_oracle_fill_pattern(uint8_t* buf, uint64_t nbytes, uint64_t seed, uint64_t word_index0) {
  uint64_t nwords = nbytes / 8;
  for (uint64_t j = 0; j < nwords; ++j) {
    uint64_t w = ccm_oracle_pattern_word(seed, word_index0 + j);
    memcpy(buf + 8 * j, &w, 8); /* host is little-endian, as is the GPU */
  }
  memset(buf + 8 * nwords, 0, (size_t)(nbytes - 8 * nwords));
}

/* Expected verify result for a pattern-filled region, without materialising it. */
API uint64_t ccm_oracle_pattern_count(uint64_t nbytes, uint64_t seed, uint64_t word_index0) {
  uint64_t nwords = nbytes / 8, c = 0;
  for (uint64_t j = 0; j < nwords; ++j) {
    uint64_t w = ccm_oracle_pattern_word(seed, word_index0 + j);
    for (int b = 0; b < 8; ++b) c += ((w >> (8 * b)) & 0xFF) != 0;
  }
  return c;
}

/* ---- multi-threaded scrub + verify over a host buffer (CPU baseline) ---------- */
/* The buffer is cut into one contiguous slice per thread.  With pin != 0 thread i runs on
 * the i-th CPU this process may use (sched_getaffinity order) for EVERY pass, including the
 * first-touch fill (ccm_oracle_fill_mt) — so each slice is allocated on, and later streamed
 * from, the NUMA node of the core that owns it.  (Round 1 first-touched the whole buffer from
 * one Python thread: all 128 workers then hammered one node's DRAM, 26 GB/s on one box and
 * 112 GB/s on another — VERDICT r1 weak #3.) */
#include <sys/mman.h>

typedef struct { uint8_t* p; uint64_t n; uint64_t nz; int do_fill; int fill_byte; int do_scrub; int do_verify; int cpu; } job_t;

static int nth_allowed_cpu(int i) {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof set, &set) != 0) return -1;
  int n = CPU_COUNT(&set);
  if (n <= 0) return -1;
  i %= n;
  for (int c = 0; c < CPU_SETSIZE; ++c)
    if (CPU_ISSET(c, &set) && i-- == 0) return c;
  return -1;
}

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  if (j->cpu >= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(j->cpu, &set);
    pthread_setaffinity_np(pthread_self(), sizeof set, &set); /* best effort */
  }
  if (j->do_fill) memset(j->p, j->fill_byte, (size_t)j->n);
  if (j->do_scrub) ccm_oracle_scrub(j->p, j->n);
  if (j->do_verify) j->nz = ccm_oracle_count_nonzero(j->p, j->n);
  return NULL;
}

static uint64_t run_mt(uint8_t* buf, uint64_t n, int threads, int pin, int do_fill, int fill_byte, int mode) {
  if (threads < 1) threads = 1;
  if (threads > 1024) threads = 1024;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)threads);
  /* slices are multiples of 2 MiB when the buffer is large, so a transparent huge page is
   * never shared between two threads (two NUMA nodes) */
  const uint64_t align = n >= ((uint64_t)threads << 22) ? (2ull << 20) : 64;
  uint64_t per = (n / (uint64_t)threads + align - 1) & ~(align - 1), off = 0, total = 0;
  for (int i = 0; i < threads; ++i) {
    uint64_t len = off >= n ? 0 : (n - off < per || i == threads - 1 ? n - off : per);
    jobs[i] = (job_t){buf + off, len, 0, do_fill, fill_byte, mode & 1, (mode >> 1) & 1, pin ? nth_allowed_cpu(i) : -1};
    off += len;
    pthread_create(&th[i], NULL, worker, &jobs[i]);
  }
  for (int i = 0; i < threads; ++i) { pthread_join(th[i], NULL); total += jobs[i].nz; }
  free(th);
  free(jobs);
  return total;
}

/* Parallel FIRST TOUCH: thread i writes `byte` over slice i (same partition and pinning as the
 * timed passes).  Also asks for transparent huge pages on the buffer (best effort). */
API void ccm_oracle_fill_mt(uint8_t* buf, uint64_t n, int threads, int byte, int pin) {
  uintptr_t a = ((uintptr_t)buf + 4095) & ~(uintptr_t)4095, b = ((uintptr_t)buf + n) & ~(uintptr_t)4095;
  if (b > a) madvise((void*)a, b - a, MADV_HUGEPAGE);
  run_mt(buf, n, threads, pin, 1, byte, 0);
}

/* mode bit 0: scrub, bit 1: verify.  Returns the non-zero byte count (0 if no verify). */
API uint64_t ccm_oracle_scrub_verify_mt_pinned(uint8_t* buf, uint64_t n, int threads, int mode, int pin) {
  return run_mt(buf, n, threads, pin, 0, 0, mode);
}

API uint64_t ccm_oracle_scrub_verify_mt(uint8_t* buf, uint64_t n, int threads, int mode) {
  return run_mt(buf, n, threads, 0, 0, 0, mode);
}

/* ---- persistent pinned pool (bench.py's CPU arm) -------------------------------------------
 * Same partition and pinning as above, but the workers live across passes (no 128 x
 * pthread_create per pass) and the scrub can use non-temporal stores (scrub_kind 1): a plain
 * memset of a region that is not in cache first READS every line (write-allocate), which halves
 * the useful DRAM bandwidth of the scrub — glibc switches to NT stores only above a
 * cache-size-dependent threshold that differs from host to host.  scrub_kind 0 = memset (libc's
 * choice), 1 = explicit streaming stores.  The result is the same bytes either way. */
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2")))
static void scrub_stream_avx2(uint8_t* p, uint64_t n) {
  uint64_t head = (32 - ((uintptr_t)p & 31)) & 31;
  if (head > n) head = n;
  memset(p, 0, head);
  p += head; n -= head;
  const __m256i z = _mm256_setzero_si256();
  uint64_t i = 0;
  for (; i + 128 <= n; i += 128) {
    _mm256_stream_si256((__m256i*)(p + i), z);
    _mm256_stream_si256((__m256i*)(p + i + 32), z);
    _mm256_stream_si256((__m256i*)(p + i + 64), z);
    _mm256_stream_si256((__m256i*)(p + i + 96), z);
  }
  _mm_sfence();
  memset(p + i, 0, n - i);
}
static int have_avx2(void) { return __builtin_cpu_supports("avx2"); }
#else
static void scrub_stream_avx2(uint8_t* p, uint64_t n) { memset(p, 0, n); }
static int have_avx2(void) { return 0; }
#endif

typedef struct pool_s {
  int threads, pin, stop, mode, scrub_kind;
  pthread_barrier_t start, done;
  pthread_t* th;
  struct pool_job { struct pool_s* pool; uint8_t* p; uint64_t n; uint64_t nz; int idx; } * jobs;
} pool_t;

static void* pool_worker(void* arg) {
  struct pool_job* j = (struct pool_job*)arg;
  pool_t* P = j->pool;
  if (P->pin) {
    int cpu = nth_allowed_cpu(j->idx);
    if (cpu >= 0) { cpu_set_t set; CPU_ZERO(&set); CPU_SET(cpu, &set); pthread_setaffinity_np(pthread_self(), sizeof set, &set); }
  }
  for (;;) {
    pthread_barrier_wait(&P->start);
    if (P->stop) break;
    if (P->mode & 4) memset(j->p, 0xA5, (size_t)j->n); /* first touch / poison */
    if (P->mode & 1) { if (P->scrub_kind == 1 && have_avx2()) scrub_stream_avx2(j->p, j->n); else ccm_oracle_scrub(j->p, j->n); }
    if (P->mode & 2) j->nz = ccm_oracle_count_nonzero(j->p, j->n);
    pthread_barrier_wait(&P->done);
  }
  return NULL;
}

API void* ccm_oracle_pool_create(uint8_t* buf, uint64_t n, int threads, int pin) {
  if (threads < 1) threads = 1;
  if (threads > 1024) threads = 1024;
  pool_t* P = (pool_t*)calloc(1, sizeof(pool_t));
  P->threads = threads; P->pin = pin;
  pthread_barrier_init(&P->start, NULL, (unsigned)threads + 1);
  pthread_barrier_init(&P->done, NULL, (unsigned)threads + 1);
  P->th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  P->jobs = (struct pool_job*)calloc((size_t)threads, sizeof(struct pool_job));
  uintptr_t a = ((uintptr_t)buf + 4095) & ~(uintptr_t)4095, b = ((uintptr_t)buf + n) & ~(uintptr_t)4095;
  if (b > a) madvise((void*)a, b - a, MADV_HUGEPAGE);
  const uint64_t align = n >= ((uint64_t)threads << 22) ? (2ull << 20) : 64;
  uint64_t per = (n / (uint64_t)threads + align - 1) & ~(align - 1), off = 0;
  for (int i = 0; i < threads; ++i) {
    uint64_t len = off >= n ? 0 : (n - off < per || i == threads - 1 ? n - off : per);
    P->jobs[i].pool = P; P->jobs[i].p = buf + off; P->jobs[i].n = len; P->jobs[i].idx = i;
    off += len;
    pthread_create(&P->th[i], NULL, pool_worker, &P->jobs[i]);
  }
  return P;
}

/* mode bit 0 scrub, bit 1 verify, bit 2 poison-first (0xA5; the first-touch pass). */
API uint64_t ccm_oracle_pool_pass(void* pool, int mode, int scrub_kind) {
  pool_t* P = (pool_t*)pool;
  P->mode = mode; P->scrub_kind = scrub_kind;
  for (int i = 0; i < P->threads; ++i) P->jobs[i].nz = 0;
  pthread_barrier_wait(&P->start);
  pthread_barrier_wait(&P->done);
  uint64_t total = 0;
  for (int i = 0; i < P->threads; ++i) total += P->jobs[i].nz;
  return total;
}

API void ccm_oracle_pool_destroy(void* pool) {
  pool_t* P = (pool_t*)pool;
  P->stop = 1;
  pthread_barrier_wait(&P->start);
  for (int i = 0; i < P->threads; ++i) pthread_join(P->th[i], NULL);
  pthread_barrier_destroy(&P->start);
  pthread_barrier_destroy(&P->done);
  free(P->th); free(P->jobs); free(P);
}
