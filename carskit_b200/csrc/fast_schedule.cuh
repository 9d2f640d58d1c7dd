// fast_schedule.cuh -- K7h: the record stream, chunk table and damping factors of the FAST schedule, on the device.
//
//   1. the caller's arrays cross PCIe (StagedCopier); validate_ids_kernel range-checks them
//   2. CUB stable radix sort of (u, n): a user's ratings become contiguous, in reference order inside the user
//   3. pack_recs_kernel + gather_recs_kernel move whole 32-byte records into that order
//   4. user_ptr[u] = first sorted position of user u (binary search over the sorted keys);
//      chunk c starts at the first user boundary at or after c * chunk_len (binary search over user_ptr), so every
//      chunk is a whole number of users and a user longer than chunk_len is one chunk
//   5. degree histograms of the items (warp-aggregated atomics) and, for CAMF_C, of the conditions
//      (shared-memory histogram); scale[row] = min(1, max_conc / (degree * groups_in_flight / nnz)) -- the expected
//      number of in-flight ratings that share the row (fast_kernels.cuh).
#pragma once
#include <algorithm>

#include "schedule_gpu.cuh"

namespace cars {

__global__ void __launch_bounds__(256) fast_user_ptr_kernel(const uint32_t* __restrict__ skey, int64_t n, int32_t num_users,
                                                            int64_t* __restrict__ user_ptr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u <= num_users; u += stride) {
    int64_t lo = 0, hi = n;  // first position with skey >= u
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)skey[mid] < u) lo = mid + 1;
      else hi = mid;
    }
    user_ptr[u] = lo;
  }
}

__global__ void __launch_bounds__(256) fast_chunk_kernel(const int64_t* __restrict__ user_ptr, int32_t num_users, int64_t n,
                                                         int64_t chunk_len, int64_t num_chunks, int64_t* __restrict__ chunk_start,
                                                         unsigned long long* __restrict__ max_chunk) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= num_chunks; c += stride) {
    auto start_of = [&](int64_t cc) -> int64_t {
      if (cc >= num_chunks) return n;
      const int64_t t = cc * chunk_len;
      int64_t lo = 0, hi = num_users;  // first user whose first position is >= t (user_ptr[num_users] = n >= t)
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (user_ptr[mid] < t) lo = mid + 1;
        else hi = mid;
      }
      return user_ptr[lo];
    };
    const int64_t a = start_of(c);
    chunk_start[c] = a;
    if (c < num_chunks) {
      const int64_t b = start_of(c + 1);
      if (b > a) atomicMax(max_chunk, (unsigned long long)(b - a));
    }
  }
}

// deg[key[n]]++ with one atomic per distinct key in a warp (a Zipf-head item is most of a warp's keys)
__global__ void __launch_bounds__(256) fast_degree_kernel(const int32_t* __restrict__ key, int64_t n, unsigned* __restrict__ deg) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
    const int64_t i = i0 + threadIdx.x;
    const bool ok = i < n;
    const unsigned active = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int k = key[i];
      const unsigned same = __match_any_sync(active, k);
      if ((threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(deg + k, (unsigned)__popc(same));
    }
  }
}

// CAMF_C: ratings per condition.  C is small (hundreds): per-CTA shared-memory histogram, then one atomic per bin.
__global__ void __launch_bounds__(256) fast_cond_degree_kernel(const int32_t* __restrict__ ctx, int64_t n,
                                                               const int32_t* __restrict__ ctx_tab, int Dmax, int C,
                                                               bool use_smem, unsigned long long* __restrict__ deg) {
  extern __shared__ unsigned sh_hist[];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (!use_smem) {  // a condition table too large for shared memory: global atomics
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int32_t* row = ctx_tab + (int64_t)ctx[i] * Dmax;
      for (int d = 0; d < Dmax; d++) {
        const int cond = __ldg(row + d);
        if (cond >= 0) atomicAdd(deg + cond, 1ull);
      }
    }
    return;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) sh_hist[c] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t* row = ctx_tab + (int64_t)ctx[i] * Dmax;
    for (int d = 0; d < Dmax; d++) {
      const int cond = __ldg(row + d);
      if (cond >= 0) atomicAdd(sh_hist + cond, 1u);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    if (sh_hist[c]) atomicAdd(deg + c, (unsigned long long)sh_hist[c]);
}

// scale[k] = min(1, max_conc * nnz / (deg[k] * in_flight)); *min_bits = bit pattern of the smallest scale
template <typename DegT>
__global__ void __launch_bounds__(256) fast_scale_kernel(const DegT* __restrict__ deg, int64_t count, double nnz, double in_flight,
                                                         double max_conc, double* __restrict__ scale,
                                                         unsigned long long* __restrict__ min_bits,
                                                         unsigned long long* __restrict__ max_deg) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const double d = (double)deg[k];
    double sc = 1.0;
    if (d > 0.0 && max_conc > 0.0) {
      const double conc = d * in_flight / nnz;
      if (conc > max_conc) sc = max_conc / conc;
    }
    scale[k] = sc;
    if (sc < 1.0) atomicMin(min_bits, (unsigned long long)__double_as_longlong(sc));  // positive doubles order like ints
    atomicMax(max_deg, (unsigned long long)deg[k]);
  }
}

// items rated at least min_degree times: (degree << 32 | item) appended to `out` (capacity `cap`)
__global__ void __launch_bounds__(256) fast_hot_candidates_kernel(const unsigned* __restrict__ deg, int32_t num_items,
                                                                  int64_t min_degree, unsigned long long* __restrict__ out,
                                                                  unsigned* __restrict__ count, unsigned cap) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < num_items; j += stride) {
    const unsigned d = deg[j];
    if ((int64_t)d >= min_degree && d > 0) {
      const unsigned k = atomicAdd(count, 1u);
      if (k < cap) out[k] = ((unsigned long long)d << 32) | (unsigned long long)(uint32_t)j;
    }
  }
}

constexpr int kHotCandCap = 512;
// A row is accumulated per CTA only if every CTA flushes it at least kHotMinFlushes times per epoch: a row that is
// merely popular would sit in shared memory until the end of the kernel -- one stale batch step per epoch -- and the
// staleness bound the damping is computed from (hot_flush updates held back by every CTA) would not describe it.
// The bar is stated per 32 groups (a CTA of 8-lane groups), whatever the kernel's lanes per rating: a hot row's step is
// damped by what every CTA holds back (hot_flush x CTAs), far more than the same row would be as an ordinary row, so a
// launch shape with fewer, wider CTAs must not admit MORE rows (measured on Zipf(1.0), 600 K ratings, BiasedMF F = 10:
// 16 hot rows instead of 6 cost +0.04 held-out RMSE after 8 epochs).
constexpr int64_t kHotMinFlushes = 8;

struct FastBuild {
  int num_hot = 0;
  int64_t num_chunks = 0, max_chunk = 0, bad_index = -1;
  int64_t h2d_bytes = 0, kernel_launches = 0;
  int64_t max_item_degree = 0;
  double min_item_scale = 1.0, min_cond_scale = 1.0;
  double copy_ms = 0.0, sort_ms = 0.0, pack_ms = 0.0;
};

// d_rec [nnz], d_chunk_start [nnz / chunk_len + 2], d_item_scale [num_items], d_cond_scale [C] (CAMF_C, else nullptr)
// are the caller's (handle-owned) allocations.  Returns the failing CUDA error, or cudaSuccess with info->bad_index >= 0
// when an id is out of range.
inline cudaError_t build_fast_on_device(int32_t num_users, int32_t num_items, int32_t num_contexts, int64_t nnz, const int32_t* u,
                                        const int32_t* j, const int32_t* ctx, const double* r, cudaStream_t stream, int sm_count,
                                        StagedCopier& copier, int64_t chunk_len, double groups_in_flight, double max_conc,
                                        const int32_t* d_ctx_tab, int Dmax, int C, RatingRec* d_rec, int64_t* d_chunk_start,
                                        double* d_item_scale, double* d_cond_scale, int hot_max, int hot_flush, int grid_ctas, int64_t hot_min_degree,
                                        signed char* d_hot_slot, int32_t* d_hot_items, const DevMem& mem, FastBuild* info) {
  if (nnz == 0) return cudaSuccess;
  const size_t N = (size_t)nnz;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  char* arena = nullptr;
  cudaError_t e = cudaSuccess;
  auto cleanup = [&]() {
    mem.free(arena);
    for (auto& x : ev)
      if (x) cudaEventDestroy(x);
  };
#define FB_TRY(x)           \
  do {                      \
    e = (x);                \
    if (e != cudaSuccess) { \
      cleanup();            \
      return e;             \
    }                       \
  } while (0)
  for (auto& x : ev) FB_TRY(cudaEventCreate(&x));
  auto bits_for = [](int64_t n_values) {
    int b = 1;
    while ((1ll << b) < n_values) b++;
    return b;
  };
  size_t temp_bytes = 0;
  FB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                         (uint32_t*)nullptr, nnz, 0, 32, stream));
  size_t total = 0;
  auto reserve = [&](size_t bytes) {
    const size_t off = total;
    total += (bytes + 255) & ~(size_t)255;
    return off;
  };
  const size_t o_u = reserve(N * 4), o_j = reserve(N * 4), o_ctx = reserve(ctx ? N * 4 : 0), o_r = reserve(N * 8),
               o_idx = reserve(N * 4), o_ord = reserve(N * 4), o_skey = reserve(N * 4), o_temp = reserve(temp_bytes),
               o_tmp = reserve(N * sizeof(RatingRec)), o_uptr = reserve(((size_t)num_users + 1) * 8),
               o_ideg = reserve((size_t)num_items * 4), o_cdeg = reserve((size_t)(C > 0 ? C : 1) * 8), o_scal = reserve(64),
               o_hot = reserve((size_t)kHotCandCap * 8 + 64);
  FB_TRY(mem.alloc((void**)&arena, total));
  FB_TRY(cudaStreamSynchronize(stream));  // the copier fills the arena from its own streams
  RatingSoA d;
  d.u = (int32_t*)(arena + o_u); d.j = (int32_t*)(arena + o_j); d.ctx = ctx ? (int32_t*)(arena + o_ctx) : nullptr;
  d.r = (double*)(arena + o_r);
  uint32_t* d_idx = (uint32_t*)(arena + o_idx);
  uint32_t* d_ord = (uint32_t*)(arena + o_ord);
  uint32_t* d_skey = (uint32_t*)(arena + o_skey);
  void* d_temp = arena + o_temp;
  RatingRec* d_tmp_rec = (RatingRec*)(arena + o_tmp);
  int64_t* d_user_ptr = (int64_t*)(arena + o_uptr);
  unsigned* d_ideg = (unsigned*)(arena + o_ideg);
  unsigned long long* d_cdeg = (unsigned long long*)(arena + o_cdeg);
  unsigned long long* d_hot_cand = (unsigned long long*)(arena + o_hot);
  unsigned* d_hot_count = (unsigned*)(arena + o_hot + (size_t)kHotCandCap * 8);
  unsigned long long* d_scal = (unsigned long long*)(arena + o_scal);  // [0] bad, [1] max chunk, [2] min item scale bits,
                                                                       // [3] max item degree, [4] min cond scale bits, [5] max cond degree

  FB_TRY(cudaEventRecord(ev[0], stream));
  {
    CopySeg segs[4] = {{d.u, (void*)u, N * 4}, {d.j, (void*)j, N * 4}, {d.ctx, (void*)ctx, ctx ? N * 4 : 0}, {d.r, (void*)r, N * 8}};
    FB_TRY(copier.run(segs, 4, true));
  }
  info->h2d_bytes += nnz * (ctx ? 20 : 16);
  FB_TRY(cudaEventRecord(ev[1], stream));
  const int blocks = sm_count * 8;
  const unsigned long long one_bits = (unsigned long long)0x3ff0000000000000ull;  // 1.0
  const unsigned long long init[6] = {~0ull, 0ull, one_bits, 0ull, one_bits, 0ull};
  FB_TRY(cudaMemcpyAsync(d_scal, init, sizeof init, cudaMemcpyHostToDevice, stream));
  validate_ids_kernel<<<blocks, 256, 0, stream>>>(d.u, d.j, d.ctx, nnz, (uint32_t)num_users, (uint32_t)num_items,
                                                  (uint32_t)num_contexts, d_scal);
  FB_TRY(cudaGetLastError());
  unsigned long long bad = 0;
  FB_TRY(cudaMemcpyAsync(&bad, d_scal, 8, cudaMemcpyDeviceToHost, stream));
  FB_TRY(cudaStreamSynchronize(stream));
  info->kernel_launches += 1;
  if (bad != ~0ull) {
    info->bad_index = (int64_t)bad;
    cleanup();
    return cudaSuccess;
  }

  // ---- user-sorted record stream --------------------------------------------------------------------------
  iota_kernel<<<blocks, 256, 0, stream>>>(d_idx, nnz);
  FB_TRY(cudaGetLastError());
  {
    size_t tb = temp_bytes;
    FB_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tb, (const uint32_t*)d.u, d_skey, d_idx, d_ord, nnz, 0, bits_for(num_users), stream));
  }
  FB_TRY(cudaEventRecord(ev[2], stream));
  pack_recs_kernel<<<blocks, 256, 0, stream>>>(d, nnz, d_tmp_rec);
  FB_TRY(cudaGetLastError());
  gather_recs_kernel<<<blocks, 256, 0, stream>>>(d_tmp_rec, d_ord, nnz, d_rec);
  FB_TRY(cudaGetLastError());
  info->kernel_launches += 4;

  // ---- chunks -----------------------------------------------------------------------------------------------
  const int64_t num_chunks = (nnz + chunk_len - 1) / chunk_len;
  fast_user_ptr_kernel<<<blocks, 256, 0, stream>>>(d_skey, nnz, num_users, d_user_ptr);
  FB_TRY(cudaGetLastError());
  fast_chunk_kernel<<<blocks, 256, 0, stream>>>(d_user_ptr, num_users, nnz, chunk_len, num_chunks, d_chunk_start, d_scal + 1);
  FB_TRY(cudaGetLastError());
  info->kernel_launches += 2;

  // ---- damping factors ---------------------------------------------------------------------------------------
  FB_TRY(cudaMemsetAsync(d_ideg, 0, (size_t)num_items * 4, stream));
  fast_degree_kernel<<<blocks, 256, 0, stream>>>(d.j, nnz, d_ideg);
  FB_TRY(cudaGetLastError());
  fast_scale_kernel<unsigned><<<blocks, 256, 0, stream>>>(d_ideg, num_items, (double)nnz, groups_in_flight, max_conc, d_item_scale,
                                                           d_scal + 2, d_scal + 3);
  FB_TRY(cudaGetLastError());
  info->kernel_launches += 2;
  if (d_cond_scale) {
    FB_TRY(cudaMemsetAsync(d_cdeg, 0, (size_t)C * 8, stream));
    const bool use_smem = (size_t)C * 4 <= 48 * 1024;
    fast_cond_degree_kernel<<<sm_count * 2, 256, use_smem ? (size_t)C * 4 : 0, stream>>>(d.ctx, nnz, d_ctx_tab, Dmax, C, use_smem,
                                                                                         d_cdeg);
    FB_TRY(cudaGetLastError());
    // CAMF_C: a rating moves Dmax condBias cells at once and all of them move the SAME prediction, so the shared
    // vector as a whole sees Dmax times the concurrency of one cell: the cap is divided by Dmax
    fast_scale_kernel<unsigned long long><<<blocks, 256, 0, stream>>>(d_cdeg, C, (double)nnz, groups_in_flight,
                                                                      max_conc > 0.0 ? max_conc / (Dmax > 0 ? Dmax : 1) : max_conc,
                                                                      d_cond_scale, d_scal + 4, d_scal + 5);
    FB_TRY(cudaGetLastError());
    info->kernel_launches += 2;
  }
  // ---- hot rows: the most popular items, accumulated per CTA in shared memory (fast_kernels.cuh) -------------------
  std::vector<unsigned long long> cand;
  if (hot_max > 0 && max_conc > 0.0 && d_hot_slot) {
    FB_TRY(cudaMemsetAsync(d_hot_count, 0, 4, stream));
    fast_hot_candidates_kernel<<<blocks, 256, 0, stream>>>(d_ideg, num_items, hot_min_degree, d_hot_cand,
                                                           d_hot_count, kHotCandCap);
    FB_TRY(cudaGetLastError());
    unsigned ncand = 0;
    FB_TRY(cudaMemcpyAsync(&ncand, d_hot_count, 4, cudaMemcpyDeviceToHost, stream));
    FB_TRY(cudaStreamSynchronize(stream));
    info->kernel_launches += 1;
    if (ncand > (unsigned)kHotCandCap) ncand = kHotCandCap;
    if (ncand) {
      cand.resize(ncand);
      FB_TRY(cudaMemcpy(cand.data(), d_hot_cand, (size_t)ncand * 8, cudaMemcpyDeviceToHost));
      std::sort(cand.begin(), cand.end(), [](unsigned long long a, unsigned long long b) {
        return (a >> 32) != (b >> 32) ? (a >> 32) > (b >> 32) : (uint32_t)a < (uint32_t)b;  // degree desc, item id asc
      });
      if ((int)cand.size() > hot_max) cand.resize((size_t)hot_max);
    }
  }
  unsigned long long scal[6];
  FB_TRY(cudaMemcpyAsync(scal, d_scal, sizeof scal, cudaMemcpyDeviceToHost, stream));
  FB_TRY(cudaStreamSynchronize(stream));
  if (!cand.empty()) {
    // a hot row's staleness: the in-flight ratings that share it PLUS what the CTAs hold back between flushes
    std::vector<int32_t> items(cand.size());
    FB_TRY(cudaMemsetAsync(d_hot_slot, 0xff, (size_t)num_items, stream));
    double min_scale = 1.0;
    memcpy(&min_scale, &scal[2], 8);
    for (size_t k = 0; k < cand.size(); k++) {
      const int32_t item = (int32_t)(uint32_t)cand[k];
      const double deg = (double)(cand[k] >> 32);
      items[k] = item;
      const double stale = deg * groups_in_flight / (double)nnz + (double)hot_flush * grid_ctas;
      const double sc = stale > max_conc ? max_conc / stale : 1.0;
      if (sc < min_scale) min_scale = sc;
      const signed char slot = (signed char)k;
      FB_TRY(cudaMemcpyAsync(d_hot_slot + item, &slot, 1, cudaMemcpyHostToDevice, stream));
      FB_TRY(cudaMemcpyAsync(d_item_scale + item, &sc, 8, cudaMemcpyHostToDevice, stream));
    }
    memcpy(&scal[2], &min_scale, 8);
    FB_TRY(cudaMemcpyAsync(d_hot_items, items.data(), items.size() * 4, cudaMemcpyHostToDevice, stream));
    FB_TRY(cudaStreamSynchronize(stream));  // `items`, `slot`, `sc` are host temporaries
    info->num_hot = (int)cand.size();
  }
  FB_TRY(cudaEventRecord(ev[3], stream));
  FB_TRY(cudaStreamSynchronize(stream));
  info->num_chunks = num_chunks;
  info->max_chunk = (int64_t)scal[1];
  info->max_item_degree = (int64_t)scal[3];
  memcpy(&info->min_item_scale, &scal[2], 8);
  memcpy(&info->min_cond_scale, &scal[4], 8);
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) info->copy_ms = ms;
  if (cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) info->sort_ms = ms;
  if (cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) info->pack_ms = ms;
  cleanup();
#undef FB_TRY
  return cudaSuccess;
}

}  // namespace cars
