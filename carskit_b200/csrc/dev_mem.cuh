// dev_mem.cuh -- device allocations through the device's stream-ordered memory pool.
//
// cudaMalloc / cudaFree of the multi-GB record stream, schedule arena and model cost 5-1000 ms depending on the box
// and on what the process freed just before (round 1 measured 79-215 ms for the same cars_create; round 2 saw 1 054 ms
// once).  cudaMallocAsync from a pool whose release threshold is raised keeps freed memory cached in the process, so the
// second handle a caller creates (K cross-validation folds, the next buildModel()) does not pay the OS again.
// tuning "pool=0" falls back to cudaMalloc / cudaFree.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace cars {

inline cudaError_t pool_setup(int device) {
  cudaMemPool_t pool;
  cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
  if (e != cudaSuccess) return e;
  uint64_t keep = UINT64_MAX;  // never hand cached memory back while the process lives (cudaMemPoolTrimTo undoes it)
  return cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
}

struct DevMem {
  bool pooled = true;
  cudaStream_t stream = nullptr;
  cudaError_t alloc(void** p, size_t bytes) const {
    if (bytes == 0) bytes = 1;
    return pooled ? cudaMallocAsync(p, bytes, stream) : cudaMalloc(p, bytes);
  }
  void free(void* p) const {
    if (!p) return;
    if (pooled) cudaFreeAsync(p, stream);
    else cudaFree(p);
  }
};

// Compute capability and SM count by cudaDeviceGetAttribute (microseconds).  cudaGetDeviceProperties fills ~80 fields through
// the driver and was measured at 3-88 ms per call on the B200 boxes (sched_trace) -- the largest jitter of cars_create.
struct DeviceFacts {
  int major = 0, minor = 0, sm_count = 0;
};
inline cudaError_t device_facts(int device, DeviceFacts* out) {
  cudaError_t e = cudaDeviceGetAttribute(&out->major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&out->minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&out->sm_count, cudaDevAttrMultiProcessorCount, device);
  return e;
}

}  // namespace cars
