// alloc_probe.cu — how expensive is it to OBTAIN and RETURN "all of HBM"?
// The cold product call (ccm_scrub_verify) spends ~50 ms in kernels and ~200 ms in
// cudaMalloc/cudaFree of a 190 GB arena; this probe times the alternatives:
//   A cudaMalloc / cudaFree                      (what the library does today)
//   B cudaMallocAsync / cudaFreeAsync + trim     (stream-ordered pool)
//   C VMM: cuMemAddressReserve + N x (cuMemCreate + cuMemMap) + cuMemSetAccess,
//          then cuMemUnmap + cuMemRelease        (chunk = 2 GiB or one handle)
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a alloc_probe.cu -lcuda
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

static double now() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define RT(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
#define DR(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("ERR %s: %s\n", #x, s); return 1; } } while (0)

__global__ void touch(uint4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(0, 0, 0, 0);
}

int main(int argc, char** argv) {
  RT(cudaSetDevice(0));
  RT(cudaFree(0));
  size_t fr, tot;
  RT(cudaMemGetInfo(&fr, &tot));
  size_t want = (fr - (256ull << 20)) & ~((2ull << 20) - 1);
  printf("free %.2f GiB total %.2f GiB want %.2f GiB\n", fr / 1073741824.0, tot / 1073741824.0, want / 1073741824.0);

  for (int rep = 0; rep < 3; ++rep) {  // A
    void* p;
    double t0 = now();
    RT(cudaMalloc(&p, want));
    double t1 = now();
    touch<<<1184, 256>>>((uint4*)p, want / 16);
    RT(cudaDeviceSynchronize());
    double t2 = now();
    RT(cudaFree(p));
    double t3 = now();
    printf("A cudaMalloc %.1f ms  touch %.1f ms  cudaFree %.1f ms\n", t1 - t0, t2 - t1, t3 - t2);
  }
  {  // B
    cudaMemPool_t pool;
    RT(cudaDeviceGetDefaultMemPool(&pool, 0));
    cudaStream_t st;
    RT(cudaStreamCreate(&st));
    for (int rep = 0; rep < 3; ++rep) {
      void* p;
      double t0 = now();
      RT(cudaMallocAsync(&p, want, st));
      RT(cudaStreamSynchronize(st));
      double t1 = now();
      touch<<<1184, 256, 0, st>>>((uint4*)p, want / 16);
      RT(cudaStreamSynchronize(st));
      double t2 = now();
      RT(cudaFreeAsync(p, st));
      RT(cudaStreamSynchronize(st));
      double t3 = now();
      RT(cudaMemPoolTrimTo(pool, 0));
      double t4 = now();
      printf("B mallocAsync %.1f ms  touch %.1f ms  freeAsync %.1f ms  trim %.1f ms\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3);
    }
  }
  for (size_t chunk : {(size_t)0, (size_t)2ull << 30, (size_t)16ull << 30}) {  // C
    for (int rep = 0; rep < 2; ++rep) {
      CUmemAllocationProp prop = {};
      prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id = 0;
      size_t gran;
      DR(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
      size_t total = want / gran * gran;
      size_t ck = chunk ? chunk : total;
      CUdeviceptr base;
      double t0 = now();
      DR(cuMemAddressReserve(&base, total, 0, 0, 0));
      std::vector<CUmemGenericAllocationHandle> hs;
      std::vector<size_t> sz;
      size_t off = 0;
      double t_create = 0, t_map = 0;
      while (off < total) {
        size_t n = total - off < ck ? total - off : ck;
        CUmemGenericAllocationHandle h;
        double a = now();
        CUresult r = cuMemCreate(&h, n, &prop, 0);
        double b = now();
        if (r != CUDA_SUCCESS) { printf("cuMemCreate failed at %.1f GiB\n", off / 1073741824.0); break; }
        DR(cuMemMap(base + off, n, 0, h, 0));
        double c = now();
        t_create += b - a; t_map += c - b;
        hs.push_back(h); sz.push_back(n); off += n;
      }
      CUmemAccessDesc acc = {};
      acc.location = prop.location;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      double t1 = now();
      DR(cuMemSetAccess(base, off, &acc, 1));
      double t2 = now();
      touch<<<1184, 256>>>((uint4*)base, off / 16);
      RT(cudaDeviceSynchronize());
      double t3 = now();
      DR(cuMemUnmap(base, off));
      double t4 = now();
      for (auto h : hs) DR(cuMemRelease(h));
      double t5 = now();
      DR(cuMemAddressFree(base, total));
      double t6 = now();
      printf("C chunk=%.0f GiB n=%zu: reserve+create+map %.1f ms (create %.1f map %.1f) setaccess %.1f ms touch %.1f ms unmap %.1f ms release %.1f ms addrfree %.1f ms | gran %zu\n",
             ck / 1073741824.0, hs.size(), t1 - t0, t_create, t_map, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, gran);
    }
  }
  return 0;
}
