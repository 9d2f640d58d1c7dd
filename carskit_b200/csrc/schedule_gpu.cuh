// schedule_gpu.cuh -- K7f on the device: building the level-sorted record stream of the flagged wavefront
// without a host-side sort.
//
// Host part (inherently sequential, one pass in reference order, a few ns per rating): for every rating its
// dependency level  level(n) = 1 + max(level(prev rating of u), level(prev rating of j))  and its positions
// ku / kj in the user's / item's chain.  The pass is chunked; each chunk is written straight into pinned
// staging buffers and copied to the device while the next chunk is being computed.  The caller's own
// u / j / ctx / r arrays are copied by a second host thread on a second stream at the same time.
// Device part: stable LSD radix sort of (level, n) pairs (CUB), then one gather kernel that packs the
// 32-byte RatingRec stream in level order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cub/device/device_radix_sort.cuh>
#include <thread>
#include <vector>

#include "sgd_kernels.cuh"

namespace cars {

struct RatingSoA {  // device arrays in reference order
  int32_t *u = nullptr, *j = nullptr, *ctx = nullptr, *level = nullptr, *ku = nullptr, *kj = nullptr;
  double* r = nullptr;
};

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* v, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) gather_recs_kernel(RatingSoA s, const uint32_t* __restrict__ order, int64_t n,
                                                          RatingRec* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t k = order[i];
    RatingRec x;
    x.u = s.u[k]; x.j = s.j[k]; x.ctx = s.ctx ? s.ctx[k] : 0; x.ku = s.ku[k];
    x.kj = s.kj[k]; x.pad = 0; x.r = s.r[k];
    out[i] = x;
  }
}

// Returns cudaSuccess, or the failing CUDA error; *bad_index >= 0 when an id is out of range.
struct FlaggedBuild {
  int64_t num_levels = 0, max_level_size = 0, bad_index = -1;
  int64_t h2d_bytes = 0;
  double host_ms = 0.0;
};

inline cudaError_t build_flagged_on_device(int32_t num_users, int32_t num_items, int32_t num_contexts, int64_t nnz,
                                           const int32_t* u, const int32_t* j, const int32_t* ctx, const double* r,
                                           cudaStream_t stream, int sm_count, RatingRec* d_rec, FlaggedBuild* info) {
  if (nnz == 0) return cudaSuccess;
  const int64_t CH = 1 << 22;  // ratings per staging chunk
  struct Stage {
    int32_t* i32 = nullptr;  // [3 x CH] level ku kj
    cudaEvent_t ev = nullptr;
    bool used = false;
  } st[2];
  RatingSoA d;
  uint32_t *d_idx_in = nullptr, *d_idx_out = nullptr, *d_key_out = nullptr;
  void* d_temp = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaError_t e = cudaSuccess;
  struct Chain {
    int32_t level;
    uint32_t count;
  };
  std::vector<Chain> tu, tj;  // per user / per item: level of its last rating, ratings seen so far
  std::vector<int64_t> level_count;
  std::thread copier;
  cudaError_t copy_err = cudaSuccess;
  auto cleanup = [&]() {
    if (copier.joinable()) copier.join();
    for (auto& s : st) {
      if (s.i32) cudaFreeHost(s.i32);
      if (s.ev) cudaEventDestroy(s.ev);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    cudaFree(d.u); cudaFree(d.j); cudaFree(d.ctx); cudaFree(d.level); cudaFree(d.ku); cudaFree(d.kj); cudaFree(d.r);
    cudaFree(d_idx_in); cudaFree(d_idx_out); cudaFree(d_key_out); cudaFree(d_temp);
  };
#define SG_TRY(x)              \
  do {                         \
    e = (x);                   \
    if (e != cudaSuccess) {    \
      cleanup();               \
      return e;                \
    }                          \
  } while (0)

  const int64_t ch = nnz < CH ? nnz : CH;
  for (auto& s : st) {
    SG_TRY(cudaMallocHost((void**)&s.i32, (size_t)ch * 3 * sizeof(int32_t)));
    SG_TRY(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  }
  SG_TRY(cudaMalloc((void**)&d.u, (size_t)nnz * 4));
  SG_TRY(cudaMalloc((void**)&d.j, (size_t)nnz * 4));
  if (ctx) SG_TRY(cudaMalloc((void**)&d.ctx, (size_t)nnz * 4));
  SG_TRY(cudaMalloc((void**)&d.level, (size_t)nnz * 4));
  SG_TRY(cudaMalloc((void**)&d.ku, (size_t)nnz * 4));
  SG_TRY(cudaMalloc((void**)&d.kj, (size_t)nnz * 4));
  SG_TRY(cudaMalloc((void**)&d.r, (size_t)nnz * 8));
  SG_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  try {
    tu.assign((size_t)num_users, Chain{0, 0u});
    tj.assign((size_t)num_items, Chain{0, 0u});
    level_count.assign(1024, 0);
  } catch (...) {
    cleanup();
    return cudaErrorMemoryAllocation;
  }
  // second host thread: the caller's arrays go to the device as they are (pageable copies are staged by the
  // driver and block the issuing thread, hence a thread of their own)
  int dev = 0;
  SG_TRY(cudaGetDevice(&dev));
  copier = std::thread([&, dev]() {
    cudaError_t ce = cudaSetDevice(dev);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d.u, u, (size_t)nnz * 4, cudaMemcpyHostToDevice, copy_stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d.j, j, (size_t)nnz * 4, cudaMemcpyHostToDevice, copy_stream);
    if (ce == cudaSuccess && ctx) ce = cudaMemcpyAsync(d.ctx, ctx, (size_t)nnz * 4, cudaMemcpyHostToDevice, copy_stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d.r, r, (size_t)nnz * 8, cudaMemcpyHostToDevice, copy_stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(copy_stream);
    copy_err = ce;
  });
  info->h2d_bytes += nnz * (ctx ? 20 : 16);

  int32_t num_levels = 0;
  for (int64_t base = 0, c = 0; base < nnz; base += ch, c++) {
    Stage& s = st[c & 1];
    if (s.used) SG_TRY(cudaEventSynchronize(s.ev));
    const int64_t len = (nnz - base < ch) ? nnz - base : ch;
    int32_t *sl = s.i32, *sku = s.i32 + ch, *skj = s.i32 + 2 * ch;
    for (int64_t i = 0; i < len; i++) {
      const int64_t n = base + i;
      const int32_t uu = u[n], jj = j[n];
      if ((uint32_t)uu >= (uint32_t)num_users || (uint32_t)jj >= (uint32_t)num_items ||
          (ctx && (uint32_t)ctx[n] >= (uint32_t)num_contexts)) {
        info->bad_index = n;
        cudaStreamSynchronize(stream);
        cleanup();
        return cudaSuccess;
      }
      Chain& a = tu[(size_t)uu];
      Chain& b = tj[(size_t)jj];
      const int32_t l = 1 + (a.level > b.level ? a.level : b.level);
      a.level = l;
      b.level = l;
      if (l > num_levels) {
        num_levels = l;
        if ((size_t)l >= level_count.size()) level_count.resize((size_t)l * 2, 0);
      }
      level_count[l]++;
      sl[i] = l;
      sku[i] = (int32_t)a.count++;
      skj[i] = (int32_t)b.count++;
    }
    SG_TRY(cudaMemcpyAsync(d.level + base, sl, (size_t)len * 4, cudaMemcpyHostToDevice, stream));
    SG_TRY(cudaMemcpyAsync(d.ku + base, sku, (size_t)len * 4, cudaMemcpyHostToDevice, stream));
    SG_TRY(cudaMemcpyAsync(d.kj + base, skj, (size_t)len * 4, cudaMemcpyHostToDevice, stream));
    SG_TRY(cudaEventRecord(s.ev, stream));
    s.used = true;
    info->h2d_bytes += len * 12;
  }
  copier.join();
  if (copy_err != cudaSuccess) {
    cleanup();
    return copy_err;
  }
  info->num_levels = num_levels;
  for (int32_t l = 1; l <= num_levels; l++)
    if (level_count[l] > info->max_level_size) info->max_level_size = level_count[l];

  // stable sort of (level, n) on the device, then pack the records in that order
  int bits = 1;
  while ((1ll << bits) <= num_levels) bits++;
  SG_TRY(cudaMalloc((void**)&d_idx_in, (size_t)nnz * 4));
  SG_TRY(cudaMalloc((void**)&d_idx_out, (size_t)nnz * 4));
  SG_TRY(cudaMalloc((void**)&d_key_out, (size_t)nnz * 4));
  iota_kernel<<<sm_count * 8, 256, 0, stream>>>(d_idx_in, nnz);
  SG_TRY(cudaGetLastError());
  size_t temp_bytes = 0;
  const uint32_t* keys_in = reinterpret_cast<const uint32_t*>(d.level);
  SG_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys_in, d_key_out, d_idx_in, d_idx_out, nnz, 0, bits, stream));
  SG_TRY(cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 1));
  SG_TRY(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_in, d_key_out, d_idx_in, d_idx_out, nnz, 0, bits, stream));
  gather_recs_kernel<<<sm_count * 8, 256, 0, stream>>>(d, d_idx_out, nnz, d_rec);
  SG_TRY(cudaGetLastError());
  SG_TRY(cudaStreamSynchronize(stream));
  cleanup();
#undef SG_TRY
  return cudaSuccess;
}

}  // namespace cars
