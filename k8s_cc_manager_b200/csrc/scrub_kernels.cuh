// scrub_kernels.cuh — sm_100a kernels of the HBM scrub-and-verify stage.
//
// the reference has NO scrub (reference main.py:502-529 goes
// stage -> reset -> wait -> verify-mode only); these kernels are the new stage
// inserted after main.py:529.  Contract: after scrub every byte of the region is
// 0x00; verify returns the EXACT number of bytes != 0 (u64).
//
// Bound: HBM bandwidth.  Integer / byte work only — no FP, no tensor cores (there
// is no contraction).  Algorithmic bytes: scrub = R written, verify = R read.
//
// Region decomposition (any pointer, any length):
//   [ head bytes | body: whole VB-byte vectors, VB-aligned | tail bytes ]
// head/tail are < 128+VB bytes and are handled with byte accesses by CTA 0; the
// body is cut into tiles (tile = threads * UNROLL vectors; every warp instruction
// covers whole 128-byte lines -> only full-sector writes) that are handed to
// PERSISTENT CTAs (grid = ctas_per_sm * #SMs) either statically (grid-stride) or —
// the default — dynamically: a CTA grabs the next chunk of tiles off an atomic
// counter, so SMs that drain faster take more work (see "work distribution").
//
// Kernels in this file:
//   scrub_st256_fast_kernel / verify_ld256_fast_kernel   what AUTO launches (compile-time shape)
//   scrub_st_kernel / verify_ld_kernel                   any shape / width / policy / schedule
//   scrub_tma_kernel                                     cp.async.bulk stores from one zero tile
//   fill_pattern_kernel                                  seeded test pattern (tests only)
// (Round 1 also carried a TMA bulk-LOAD verify ring; it never beat LDG.256 in speed or power and
// was the only kernel with an open racecheck finding, so it was removed — DESIGN.md §5.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ccm {

enum CachePolicy : int { kPolDefault = 0, kPolEvictFirst = 1, kPolStreaming = 2, kPolEvictLast = 3 };

// ---------------------------------------------------------------- PTX wrappers
// 128-bit stores/loads take an L2 eviction priority only through a createpolicy
// cache hint (the bare .L2::evict_* qualifier is accepted on 256-bit forms only).
template <int POL>
__device__ __forceinline__ uint64_t make_l2_policy() {
  uint64_t pol = 0;
  if constexpr (POL == kPolEvictFirst)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if constexpr (POL == kPolEvictLast)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

template <int POL>
__device__ __forceinline__ void st_zero16(void* p, uint64_t l2pol) {
  if constexpr (POL == kPolEvictFirst || POL == kPolEvictLast)
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%1,%1,%1}, %2;" ::"l"(p), "r"(0), "l"(l2pol) : "memory");
  else if constexpr (POL == kPolStreaming)
    asm volatile("st.global.cs.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(0) : "memory");
  else
    asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(0) : "memory");
}

// 256-bit store: PTX 8.8, sm_100+ only (SASS: STG.E.256).
template <int POL>
__device__ __forceinline__ void st_zero32(void* p) {
  if constexpr (POL == kPolEvictFirst)
    asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(0) : "memory");
  else if constexpr (POL == kPolStreaming)
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(0) : "memory");
  else if constexpr (POL == kPolEvictLast)
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(0) : "memory");
  else
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(0) : "memory");
}

template <int VB, int POL>
__device__ __forceinline__ void st_zero(void* p, uint64_t l2pol) {
  if constexpr (VB == 32) st_zero32<POL>(p); else st_zero16<POL>(p, l2pol);
}

struct Vec16 { uint32_t w[4]; };
struct Vec32 { uint32_t w[8]; };

template <int POL>
__device__ __forceinline__ Vec16 ld16(const void* p, uint64_t l2pol) {
  Vec16 v;
  if constexpr (POL == kPolEvictFirst || POL == kPolEvictLast)
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p), "l"(l2pol));
  else if constexpr (POL == kPolStreaming)
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p));
  else
    asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p));
  return v;
}

template <int POL>
__device__ __forceinline__ Vec32 ld32(const void* p) {
  Vec32 v;
  if constexpr (POL == kPolEvictFirst)
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]),
                   "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7]) : "l"(p));
  else if constexpr (POL == kPolStreaming)
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]),
                   "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7]) : "l"(p));
  else if constexpr (POL == kPolEvictLast)
    asm volatile("ld.global.nc.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]),
                   "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7]) : "l"(p));
  else
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]),
                   "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7]) : "l"(p));
  return v;
}

// Exact number of non-zero BYTES in a 32-bit word, branch-free:
// bit 7 of each byte of m is set iff that byte != 0.
__device__ __forceinline__ uint32_t nonzero_bytes_in_word(uint32_t w) {
  uint32_t m = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
  return __popc(m);
}

// ------------------------------------------------------------------- layout
struct RegionSplit {
  uint8_t* base;       // region start
  uint64_t head;       // bytes before the aligned body
  uint64_t body_vecs;  // number of VB-byte vectors in the body
  uint64_t tail;       // bytes after the body
};

__host__ __device__ inline RegionSplit split_region(const void* p, uint64_t n, int vb, int align) {
  RegionSplit s;
  s.base = (uint8_t*)p;
  uint64_t mis = (uint64_t)(uintptr_t)p & (uint64_t)(align - 1);
  uint64_t head = mis ? (uint64_t)align - mis : 0;
  if (head > n) head = n;
  s.head = head;
  s.body_vecs = (n - head) / (uint64_t)vb;
  s.tail = (n - head) - s.body_vecs * (uint64_t)vb;
  return s;
}

// ------------------------------------------------------------ work distribution
// Tiles are handed to the persistent CTAs either statically (grid-stride) or
// DYNAMICALLY: a CTA grabs the next chunk of `chunk_tiles` consecutive tiles with
// one atomicAdd on a device counter (zeroed by the host before the launch).  The
// grab for chunk k+1 is issued before the stores of chunk k, so its latency is
// hidden.  Why: an SM can push at most 32 B/clk into the crossbar (ncu:
// l1tex__m_l1tex2xbar_write_bytes), the chip-wide store ceiling is therefore only
// ~1.2x the HBM write ceiling, and SMs do not all drain at the same rate; with a
// static split the slowest SM sets the kernel time (profiles/r1_scrub_st256.md).
struct Sched {
  unsigned long long* counter;  // nullptr = static grid-stride
  uint32_t chunk_tiles;         // tiles per grab (dynamic only)
  uint64_t ntiles;              // whole tiles in the body   } computed on the host: no 64-bit
  uint64_t nchunks;             // ceil(ntiles / chunk_tiles) } divides in the kernel prologue
  uint32_t per_warp;            // 1: every WARP grabs its own chunks (tile = 32 lanes x UNROLL vectors);
                                //    no block barrier, no shared memory in the hot loop
};

// ------------------------------------------------------------- scrub (stores)
template <int VB, int UNROLL, int POL>
__device__ __forceinline__ void scrub_tile(uint8_t* body, uint64_t tile, uint64_t tile_vecs, uint32_t idx,
                                           uint64_t stride_bytes, uint64_t l2pol) {
  uint8_t* p = body + (tile * tile_vecs + idx) * VB;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) st_zero<VB, POL>(p + u * stride_bytes, l2pol);
}

template <int VB, int UNROLL, int POL>
__global__ void __launch_bounds__(1024)
scrub_st_kernel(RegionSplit s, Sched sched) {
  uint8_t* body = s.base + s.head;
  const uint32_t group = sched.per_warp ? 32u : blockDim.x;   // threads that share one tile
  const uint32_t idx = sched.per_warp ? (threadIdx.x & 31u) : threadIdx.x;
  const uint64_t tile_vecs = (uint64_t)group * UNROLL;
  const uint64_t ntiles = sched.ntiles;
  const uint64_t stride_bytes = (uint64_t)group * VB;
  const uint64_t l2pol = make_l2_policy<POL>();

  if (sched.counter == nullptr) {
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      scrub_tile<VB, UNROLL, POL>(body, tile, tile_vecs, idx, stride_bytes, l2pol);
  } else if (sched.per_warp) {
    const uint64_t nchunks = sched.nchunks;
    unsigned long long c = 0;
    if (idx == 0) c = atomicAdd(sched.counter, 1ull);
    c = __shfl_sync(0xffffffffu, c, 0);
    while (c < nchunks) {
      unsigned long long nxt = 0;
      if (idx == 0) nxt = atomicAdd(sched.counter, 1ull);  // prefetch the next grab
      const uint64_t t0 = c * sched.chunk_tiles;
      const uint64_t t1 = t0 + sched.chunk_tiles < ntiles ? t0 + sched.chunk_tiles : ntiles;
      for (uint64_t tile = t0; tile < t1; ++tile)
        scrub_tile<VB, UNROLL, POL>(body, tile, tile_vecs, idx, stride_bytes, l2pol);
      c = __shfl_sync(0xffffffffu, nxt, 0);
    }
  } else {
    __shared__ unsigned long long s_chunk[2];
    const uint64_t nchunks = sched.nchunks;
    if (threadIdx.x == 0) s_chunk[0] = atomicAdd(sched.counter, 1ull);
    __syncthreads();
    int buf = 0;
    for (;;) {
      const uint64_t c = s_chunk[buf];
      if (c >= nchunks) break;
      unsigned long long nxt = 0;
      if (threadIdx.x == 0) nxt = atomicAdd(sched.counter, 1ull);  // prefetch the next grab
      const uint64_t t0 = c * sched.chunk_tiles;
      const uint64_t t1 = t0 + sched.chunk_tiles < ntiles ? t0 + sched.chunk_tiles : ntiles;
      for (uint64_t tile = t0; tile < t1; ++tile)
        scrub_tile<VB, UNROLL, POL>(body, tile, tile_vecs, idx, stride_bytes, l2pol);
      if (threadIdx.x == 0) s_chunk[buf ^ 1] = nxt;
      __syncthreads();
      buf ^= 1;
    }
  }
  // remainder vectors (< one tile), spread over the grid
  for (uint64_t i = ntiles * tile_vecs + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
       i < s.body_vecs; i += (uint64_t)gridDim.x * blockDim.x)
    st_zero<VB, POL>(body + i * VB, l2pol);
  // ragged head / tail bytes
  if (blockIdx.x == 0) {
    for (uint64_t i = threadIdx.x; i < s.head; i += blockDim.x) s.base[i] = 0;
    uint8_t* t = body + s.body_vecs * VB;
    for (uint64_t i = threadIdx.x; i < s.tail; i += blockDim.x) t[i] = 0;
  }
}

// ----------------------------------------------------- default shapes, tight loops
// The two kernels below are what AUTO launches.  Same algorithm as scrub_st_kernel /
// verify_ld_kernel with a per-CTA dynamic schedule, but block size and vectors per thread
// are compile-time constants, so one chunk = THREADS * PER_THREAD * 32 bytes is addressed
// from ONE 64-bit base with immediate offsets: PER_THREAD STG.256 / LDG.256 per thread per
// chunk and ~10 instructions of bookkeeping, instead of ~13 instructions per store in the
// generic kernel (runtime stride => a 64-bit add pair per access).  Same GB/s — the kernels
// are egress / DRAM bound — fewer issued instructions, less power.
//
// Launch bookkeeping lives IN the kernel (round 2): the grab counter is handed back zeroed by
// the last CTA to finish (GrabCtl below), so a step is two kernel nodes and nothing else — no
// cudaMemsetAsync nodes in front of each launch, which cost ~2-4 us apiece and showed at the
// 1 GB end of the region sweep.  griddepcontrol.wait / launch_dependents let the host chain the
// two kernels with programmatic dependent launch (no-ops when launched the ordinary way).
struct GrabCtl {
  unsigned long long* grab;  // next chunk index; 0 before and after every launch
  unsigned long long* done;  // CTAs that have stopped grabbing; 0 before and after every launch
};

// One thread per CTA calls this after the CTA's last grab returned: the CTA that arrives last
// knows nobody will touch `grab` again and resets both words for the next launch on the stream.
__device__ __forceinline__ void grab_release(const GrabCtl& g, unsigned long long* also_zero) {
  __threadfence();
  if (atomicAdd(g.done, 1ull) == (unsigned long long)gridDim.x - 1) {
    *g.grab = 0;
    *g.done = 0;
    if (also_zero) *also_zero = 0;
    __threadfence();
  }
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// zero_on_exit (may be null): a word the last CTA clears as well — the arena step passes the
// verify counter, so "scrub" also resets the verdict of the previous read-back.
template <int THREADS, int PER_THREAD, int POL>
__global__ void __launch_bounds__(THREADS)
scrub_st256_fast_kernel(RegionSplit s, uint64_t nchunks, GrabCtl ctl, unsigned long long* zero_on_exit) {
  constexpr uint64_t kChunk = (uint64_t)THREADS * PER_THREAD * 32;
  uint8_t* body = s.base + s.head;
  __shared__ unsigned long long s_chunk[2];
  pdl_wait();
  if (threadIdx.x == 0) s_chunk[0] = atomicAdd(ctl.grab, 1ull);
  __syncthreads();
  int buf = 0;
  for (;;) {
    const uint64_t c = s_chunk[buf];
    if (c >= nchunks) break;
    unsigned long long nxt = 0;
    // next grab in flight while we store.  The raw result goes straight to shared memory at the end of
    // the iteration: any arithmetic on it here makes warp 0 wait out the ~1 us atomic before its stores
    // (a "static first chunk + offset" variant lost 10 % of the read-back rate that way, r2h).
    if (threadIdx.x == 0) nxt = atomicAdd(ctl.grab, 1ull);
    uint8_t* p = body + c * kChunk + threadIdx.x * 32u;
#pragma unroll
    for (int u = 0; u < PER_THREAD; ++u) st_zero32<POL>(p + (uint32_t)u * THREADS * 32u);
    if (threadIdx.x == 0) s_chunk[buf ^ 1] = nxt;
    __syncthreads();
    buf ^= 1;  // two slots: a slow warp may still be reading slot k when thread 0 already writes slot k+1
  }
  pdl_launch_dependents();
  if (threadIdx.x == 0) grab_release(ctl, zero_on_exit);
  // vectors after the last whole chunk (< kChunk bytes), then the ragged head / tail bytes
  for (uint64_t i = nchunks * (kChunk / 32) + (uint64_t)blockIdx.x * THREADS + threadIdx.x;
       i < s.body_vecs; i += (uint64_t)gridDim.x * THREADS)
    st_zero32<POL>(body + i * 32);
  if (blockIdx.x == 0) {
    for (uint64_t i = threadIdx.x; i < s.head; i += THREADS) s.base[i] = 0;
    uint8_t* t = body + s.body_vecs * 32;
    for (uint64_t i = threadIdx.x; i < s.tail; i += THREADS) t[i] = 0;
  }
}

// ------------------------------------------------------ scrub (TMA bulk store)
// One zeroed shared-memory tile per CTA is the source of EVERY bulk store: it
// never changes, so there is no WAR hazard and no per-op wait — the issuing lane
// streams cp.async.bulk.global.shared::cta ops (SASS: UBLKCP.G.S) grid-strided
// over the body and drains them once at exit.
template <int POL>
__device__ __forceinline__ void bulk_store(uint8_t* dst, uint32_t src_smem, uint32_t bytes, uint64_t policy) {
  if constexpr (POL == kPolEvictFirst || POL == kPolEvictLast)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst), "r"(src_smem), "r"(bytes), "l"(policy) : "memory");
  else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

template <int POL>
__global__ void __launch_bounds__(256)
scrub_tma_kernel(RegionSplit s /* VB = 16 */, uint32_t tile_bytes, uint32_t ops_per_group, Sched sched) {
  extern __shared__ __align__(128) uint8_t zero_tile[];
  for (uint32_t i = threadIdx.x; i < tile_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(zero_tile)[i] = make_uint4(0, 0, 0, 0);
  // make the generic-proxy smem writes visible to the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  uint8_t* body = s.base + s.head;
  const uint64_t body_bytes = s.body_vecs * 16;
  const uint64_t ntiles = sched.ntiles;  // = body_bytes / tile_bytes
  const uint32_t last_bytes = (uint32_t)(body_bytes - ntiles * (uint64_t)tile_bytes);  // multiple of 16

  // each WARP's lane 0 issues; warps take interleaved tiles / their own chunks.
  const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(zero_tile);
    const uint64_t policy = make_l2_policy<POL>();
    if (sched.counter == nullptr) {
      uint32_t in_group = 0;
      const uint64_t first = (uint64_t)blockIdx.x * nwarps + warp;
      const uint64_t step = (uint64_t)gridDim.x * nwarps;
      for (uint64_t tile = first; tile < ntiles; tile += step) {
        bulk_store<POL>(body + tile * (uint64_t)tile_bytes, src, tile_bytes, policy);
        if (++in_group == ops_per_group) {
          in_group = 0;
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else {
      const uint64_t nchunks = sched.nchunks;
      unsigned long long c = atomicAdd(sched.counter, 1ull);
      while (c < nchunks) {
        const unsigned long long nxt = atomicAdd(sched.counter, 1ull);
        const uint64_t t0 = c * sched.chunk_tiles;
        const uint64_t t1 = t0 + sched.chunk_tiles < ntiles ? t0 + sched.chunk_tiles : ntiles;
        for (uint64_t tile = t0; tile < t1; ++tile)
          bulk_store<POL>(body + tile * (uint64_t)tile_bytes, src, tile_bytes, policy);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        c = nxt;
      }
    }
    if (last_bytes && blockIdx.x == 0 && warp == 0)
      bulk_store<kPolDefault>(body + ntiles * (uint64_t)tile_bytes, src, last_bytes, 0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if (blockIdx.x == 0) {
    for (uint64_t i = threadIdx.x; i < s.head; i += blockDim.x) s.base[i] = 0;
    uint8_t* t = body + body_bytes;
    for (uint64_t i = threadIdx.x; i < s.tail; i += blockDim.x) t[i] = 0;
  }
}

// ----------------------------------------------------------------- reductions
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One atomicAdd(u64) per CTA into the single device counter.
__device__ __forceinline__ void block_accumulate(uint64_t cnt, unsigned long long* counter) {
  __shared__ uint64_t warp_part[32];
  cnt = warp_sum_u64(cnt);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) warp_part[warp] = cnt;
  __syncthreads();
  if (warp == 0) {
    uint64_t v = lane < ((blockDim.x + 31) >> 5) ? warp_part[lane] : 0;
    v = warp_sum_u64(v);
    if (lane == 0 && v != 0) atomicAdd(counter, (unsigned long long)v);
  }
}

// --------------------------------------------------------------- verify (loads)
// All UNROLL loads are issued before any use (UNROLL*VB bytes in flight per
// thread).  The expected answer is "all zero", so the hot loop only ORs the words
// together; the exact per-byte count runs on the (rare) batches whose OR != 0.
template <int VB, int UNROLL, int POL>
__device__ __forceinline__ uint32_t verify_tile(const uint8_t* body, uint64_t tile, uint64_t tile_vecs, uint32_t idx,
                                                uint64_t stride_bytes, uint64_t l2pol) {
  constexpr int W = VB / 4;
  const uint8_t* p = body + (tile * tile_vecs + idx) * VB;
  uint32_t w[UNROLL][W];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if constexpr (VB == 32) {
      Vec32 v = ld32<POL>(p + u * stride_bytes);
#pragma unroll
      for (int k = 0; k < W; ++k) w[u][k] = v.w[k];
    } else {
      Vec16 v = ld16<POL>(p + u * stride_bytes, l2pol);
#pragma unroll
      for (int k = 0; k < W; ++k) w[u][k] = v.w[k];
    }
  }
  uint32_t any = 0;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
#pragma unroll
    for (int k = 0; k < W; ++k) any |= w[u][k];
  uint32_t c = 0;
  if (any != 0) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int k = 0; k < W; ++k) c += nonzero_bytes_in_word(w[u][k]);
  }
  return c;
}

template <int VB, int UNROLL, int POL>
__global__ void __launch_bounds__(1024)
verify_ld_kernel(RegionSplit s, unsigned long long* counter, Sched sched) {
  const uint8_t* body = s.base + s.head;
  const uint32_t group = sched.per_warp ? 32u : blockDim.x;
  const uint32_t idx = sched.per_warp ? (threadIdx.x & 31u) : threadIdx.x;
  const uint64_t tile_vecs = (uint64_t)group * UNROLL;
  const uint64_t ntiles = sched.ntiles;
  const uint64_t stride_bytes = (uint64_t)group * VB;
  constexpr int W = VB / 4;
  uint64_t cnt = 0;
  const uint64_t l2pol = make_l2_policy<POL>();
  (void)l2pol;

  if (sched.counter == nullptr) {
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      cnt += verify_tile<VB, UNROLL, POL>(body, tile, tile_vecs, idx, stride_bytes, l2pol);
  } else if (sched.per_warp) {
    const uint64_t nchunks = sched.nchunks;
    unsigned long long c = 0;
    if (idx == 0) c = atomicAdd(sched.counter, 1ull);
    c = __shfl_sync(0xffffffffu, c, 0);
    while (c < nchunks) {
      unsigned long long nxt = 0;
      if (idx == 0) nxt = atomicAdd(sched.counter, 1ull);
      const uint64_t t0 = c * sched.chunk_tiles;
      const uint64_t t1 = t0 + sched.chunk_tiles < ntiles ? t0 + sched.chunk_tiles : ntiles;
      for (uint64_t tile = t0; tile < t1; ++tile)
        cnt += verify_tile<VB, UNROLL, POL>(body, tile, tile_vecs, idx, stride_bytes, l2pol);
      c = __shfl_sync(0xffffffffu, nxt, 0);
    }
  } else {
    __shared__ unsigned long long s_chunk[2];
    const uint64_t nchunks = sched.nchunks;
    if (threadIdx.x == 0) s_chunk[0] = atomicAdd(sched.counter, 1ull);
    __syncthreads();
    int buf = 0;
    for (;;) {
      const uint64_t c = s_chunk[buf];
      if (c >= nchunks) break;
      unsigned long long nxt = 0;
      if (threadIdx.x == 0) nxt = atomicAdd(sched.counter, 1ull);
      const uint64_t t0 = c * sched.chunk_tiles;
      const uint64_t t1 = t0 + sched.chunk_tiles < ntiles ? t0 + sched.chunk_tiles : ntiles;
      for (uint64_t tile = t0; tile < t1; ++tile)
        cnt += verify_tile<VB, UNROLL, POL>(body, tile, tile_vecs, idx, stride_bytes, l2pol);
      if (threadIdx.x == 0) s_chunk[buf ^ 1] = nxt;
      __syncthreads();
      buf ^= 1;
    }
  }
  for (uint64_t i = ntiles * tile_vecs + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
       i < s.body_vecs; i += (uint64_t)gridDim.x * blockDim.x) {
    if constexpr (VB == 32) {
      Vec32 v = ld32<POL>(body + i * VB);
#pragma unroll
      for (int k = 0; k < W; ++k) cnt += nonzero_bytes_in_word(v.w[k]);
    } else {
      Vec16 v = ld16<POL>(body + i * VB, l2pol);
#pragma unroll
      for (int k = 0; k < W; ++k) cnt += nonzero_bytes_in_word(v.w[k]);
    }
  }
  if (blockIdx.x == 0) {
    for (uint64_t i = threadIdx.x; i < s.head; i += blockDim.x) cnt += (s.base[i] != 0);
    const uint8_t* t = body + s.body_vecs * VB;
    for (uint64_t i = threadIdx.x; i < s.tail; i += blockDim.x) cnt += (t[i] != 0);
  }
  block_accumulate(cnt, counter);
}

template <int THREADS, int PER_THREAD, int POL>
__global__ void __launch_bounds__(THREADS)
verify_ld256_fast_kernel(RegionSplit s, uint64_t nchunks, GrabCtl ctl, unsigned long long* counter) {
  constexpr uint64_t kChunk = (uint64_t)THREADS * PER_THREAD * 32;
  const uint8_t* body = s.base + s.head;
  uint64_t cnt = 0;
  __shared__ unsigned long long s_chunk[2];
  pdl_wait();
  if (threadIdx.x == 0) s_chunk[0] = atomicAdd(ctl.grab, 1ull);
  __syncthreads();
  int buf = 0;
  for (;;) {
    const uint64_t c = s_chunk[buf];
    if (c >= nchunks) break;
    unsigned long long nxt = 0;
    if (threadIdx.x == 0) nxt = atomicAdd(ctl.grab, 1ull);
    // DESCENDING chunk order.  The scrub ran in ascending order and left the ~126 MB it wrote last
    // dirty in L2.  Reading upwards would stream the whole region from DRAM and, on top of that, evict
    // those dirty lines (extra write-back traffic = +17 us per launch at any size: 157 vs 140 us at
    // 1 GB, profiles/r2_small_region_shapes.log); reading downwards takes them out of L2 first, so the
    // DRAM traffic of a read-back is R again, not R + L2.
    const uint8_t* p = body + (nchunks - 1 - c) * kChunk + threadIdx.x * 32u;
    Vec32 v[PER_THREAD];
#pragma unroll
    for (int u = 0; u < PER_THREAD; ++u) v[u] = ld32<POL>(p + (uint32_t)u * THREADS * 32u);  // all loads first
    uint32_t any = 0;
#pragma unroll
    for (int u = 0; u < PER_THREAD; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) any |= v[u].w[k];
    if (any != 0) {  // rare: exact per-byte count of this batch
      uint32_t n = 0;
#pragma unroll
      for (int u = 0; u < PER_THREAD; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) n += nonzero_bytes_in_word(v[u].w[k]);
      cnt += n;
    }
    if (threadIdx.x == 0) s_chunk[buf ^ 1] = nxt;
    __syncthreads();
    buf ^= 1;
  }
  pdl_launch_dependents();
  if (threadIdx.x == 0) grab_release(ctl, nullptr);
  for (uint64_t i = nchunks * (kChunk / 32) + (uint64_t)blockIdx.x * THREADS + threadIdx.x;
       i < s.body_vecs; i += (uint64_t)gridDim.x * THREADS) {
    Vec32 v = ld32<POL>(body + i * 32);
#pragma unroll
    for (int k = 0; k < 8; ++k) cnt += nonzero_bytes_in_word(v.w[k]);
  }
  if (blockIdx.x == 0) {
    for (uint64_t i = threadIdx.x; i < s.head; i += THREADS) cnt += (s.base[i] != 0);
    const uint8_t* t = body + s.body_vecs * 32;
    for (uint64_t i = threadIdx.x; i < s.tail; i += THREADS) cnt += (t[i] != 0);
  }
  block_accumulate(cnt, counter);
}

// ------------------------------------------------------------ test scaffolding
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Deterministic sparse pattern: 8-byte word j (arena-global index) is
//   r = splitmix64(seed + j);  word = (r & 7) == 0 ? r & mask(r) : 0
// where mask keeps a pseudo-random subset of the bytes.  The C oracle
// (oracle/scrub_oracle.c: ccm_oracle_fill_random) restates the same formula.
__host__ __device__ inline uint64_t pattern_word(uint64_t seed, uint64_t j) {
  uint64_t x = seed + j;
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x = x ^ (x >> 31);
  if ((x & 7) != 0) return 0;
  uint64_t keep = 0;
  for (int b = 0; b < 8; ++b)
    if ((x >> (8 + b)) & 1) keep |= 0xFFull << (8 * b);
  return (x >> 3 | 0x0101010101010101ull) & keep;  // kept bytes are guaranteed non-zero
}

__global__ void fill_pattern_kernel(uint64_t* words, uint64_t nwords, uint64_t word_index0, uint64_t seed) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords;
       i += (uint64_t)gridDim.x * blockDim.x)
    words[i] = pattern_word(seed, word_index0 + i);
}

}  // namespace ccm
