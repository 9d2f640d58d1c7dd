// schedule_gpu.cuh -- K7f: the level-sorted record stream of the flagged wavefront, built on the device.
//
// Needed per rating n (reference order): ku / kj = its position in its user's / item's chain, and
//   level(n) = 1 + max(level(prev rating of u), level(prev rating of j))
// -- the longest path to n in the conflict DAG (schedule.cuh).  Everything runs on the GPU; the host only
// moves the caller's four arrays across PCIe (StagedCopier):
//   1. validate_ids_kernel      range check of u / j / ctx (first bad index via atomicMin)
//   2. CUB stable radix sort of (u, n) and of (j, n): neighbours in sorted order are chain neighbours
//      -> chain_heads_kernel / chain_link_kernel give ku, kj and each rating's successor in both chains
//   3. kahn_roots_kernel + kahn_levels_kernel: Kahn's algorithm by frontiers on a persistent cooperative
//      grid -- frontier L holds exactly the ratings of level L (in-degree <= 2, counted down with atomics);
//      one grid barrier per level, ~4 us each (3 268 levels at config 3)
//   4. CUB stable radix sort of (level, n), gather_recs_kernel packs the 32-byte RatingRec stream.
// The stable sort makes the stream independent of the order in which Kahn's frontiers were filled.
// The previous host pass (one thread, 8 ns per rating: 0.82 s at 100 M ratings) is kept behind the handle's
// tuning string ("levels=host") for A/B measurements and as a cross-check in the tests (build_flagged_host_levels).
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <vector>

#include "dev_mem.cuh"
#include "sgd_kernels.cuh"
#include "staged_copy.cuh"

namespace cars {

struct RatingSoA {  // device arrays in reference order
  int32_t *u = nullptr, *j = nullptr, *ctx = nullptr, *level = nullptr, *ku = nullptr, *kj = nullptr;
  double* r = nullptr;
};

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* v, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = (uint32_t)i;
}

// Records are packed in reference order first (coalesced reads of the six arrays), then moved to level order as
// whole 32-byte records: one DRAM sector per rating instead of six scattered 4 / 8-byte reads (measured
// 710 B of DRAM reads per rating before, profiles/r1/launches_r1b_default.txt).
// succ (may be null): the next rating of the same user / item, or -1 -- packed into `pad` as "last of its user" (bit 0)
// and "last of its item" (bit 1), which the tagged kernel uses to hand the row back with tag 0 for the next epoch
__global__ void __launch_bounds__(256) pack_recs_kernel(RatingSoA s, int64_t n, RatingRec* __restrict__ out,
                                                        const int32_t* __restrict__ succ = nullptr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    RatingRec x;
    x.u = s.u[k]; x.j = s.j[k]; x.ctx = s.ctx ? s.ctx[k] : 0; x.ku = s.ku ? s.ku[k] : 0;
    x.kj = s.kj ? s.kj[k] : 0; x.r = s.r[k];
    x.pad = succ ? ((succ[2 * k] < 0 ? 1 : 0) | (succ[2 * k + 1] < 0 ? 2 : 0)) : 0;
    out[k] = x;
  }
}
__global__ void __launch_bounds__(256) gather_recs_kernel(const RatingRec* __restrict__ in, const uint32_t* __restrict__ order,
                                                          int64_t n, RatingRec* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int4* src = reinterpret_cast<const int4*>(in + order[i]);
    const int4 a = __ldg(src), b = __ldg(src + 1);
    int4* dst = reinterpret_cast<int4*>(out + i);
    dst[0] = a;
    dst[1] = b;
  }
}

// first rating with an id out of range (unsigned compare also rejects negatives); *bad starts at ~0
__global__ void __launch_bounds__(256) validate_ids_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ j,
                                                           const int32_t* __restrict__ ctx, int64_t n, uint32_t num_users,
                                                           uint32_t num_items, uint32_t num_contexts,
                                                           unsigned long long* bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if ((uint32_t)u[i] >= num_users || (uint32_t)j[i] >= num_items || (ctx && (uint32_t)ctx[i] >= num_contexts))
      atomicMin(bad, (unsigned long long)i);
}

// skey / ord: the ratings stably sorted by one id (user or item).  start[key] = first sorted position.
__global__ void __launch_bounds__(256) chain_heads_kernel(const uint32_t* __restrict__ skey, int64_t n, uint32_t* __restrict__ start) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    if (p == 0 || skey[p - 1] != skey[p]) start[skey[p]] = (uint32_t)p;
}
// kpos[n] = earlier ratings of the same id; succ[2 n + which] = the next rating of the same id, or -1
__global__ void __launch_bounds__(256) chain_link_kernel(const uint32_t* __restrict__ skey, const uint32_t* __restrict__ ord,
                                                         int64_t n, const uint32_t* __restrict__ start,
                                                         int32_t* __restrict__ kpos, int32_t* __restrict__ succ, int which) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const uint32_t key = skey[p];
    const uint32_t me = ord[p];
    kpos[me] = (int32_t)((uint32_t)p - start[key]);
    succ[2 * (int64_t)me + which] = (p + 1 < n && skey[p + 1] == key) ? (int32_t)ord[p + 1] : -1;
  }
}

struct KahnCtl {  // zero-initialised
  unsigned cnt[3];  // frontier sizes, rotating: level L reads cnt[L % 3], fills cnt[(L + 1) % 3]
  unsigned barrier;
  unsigned num_levels;
  unsigned max_level;
  unsigned long long processed;
};

// in-degree of every rating in the conflict DAG (0, 1 or 2); the roots form frontier 1
__global__ void __launch_bounds__(256) kahn_roots_kernel(const int32_t* __restrict__ ku, const int32_t* __restrict__ kj, int64_t n,
                                                         int* __restrict__ indeg, uint32_t* __restrict__ frontier, KahnCtl* ctl) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
    const int64_t i = i0 + threadIdx.x;
    int d = -1;
    if (i < n) {
      d = (ku[i] > 0) + (kj[i] > 0);
      indeg[i] = d;
    }
    const unsigned m = __ballot_sync(0xffffffffu, d == 0);
    if (m) {
      const int lane = threadIdx.x & 31;
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&ctl->cnt[1], (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (d == 0) frontier[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    }
  }
}

__device__ __forceinline__ void kahn_grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add_u32(counter, 1u);
    while ((int)(ld_acquire_u32(counter) - target) < 0) {  // wrap-safe
    }
  }
  __syncthreads();
}

// Persistent cooperative grid.  Level L: every rating of frontier L gets level L and counts its (at most two)
// successors' in-degree down; a successor that reaches 0 joins frontier L + 1.  fr0 / fr1 alternate.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) kahn_levels_kernel(const int2* __restrict__ succ, int* __restrict__ indeg,
                                                                 int32_t* __restrict__ level, uint32_t* fr0, uint32_t* fr1,
                                                                 KahnCtl* ctl) {
  __shared__ unsigned s_cnt, s_base;
  __shared__ unsigned s_warp[THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned maxc = 0;
  unsigned long long total = 0;
  unsigned lvl = 1;
  for (;; lvl++) {
    const unsigned count = ld_relaxed_u32(&ctl->cnt[lvl % 3]);
    if (count == 0) break;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st_relaxed_u32(&ctl->cnt[(lvl + 2) % 3], 0u);  // read one level ago, filled one level from now
      if (count > maxc) maxc = count;
      total += count;
    }
    const uint32_t* cur = (lvl & 1) ? fr0 : fr1;
    uint32_t* nxt = (lvl & 1) ? fr1 : fr0;
    unsigned* ncnt = &ctl->cnt[(lvl + 1) % 3];
    for (unsigned i0 = blockIdx.x * THREADS; i0 < count; i0 += gridDim.x * THREADS) {
      const unsigned i = i0 + threadIdx.x;
      int a = -1, b = -1;
      if (i < count) {
        const uint32_t n = __ldcg(cur + i);  // written by other SMs one level ago: L2, not L1
        level[n] = (int32_t)lvl;
        const int2 s = __ldg(succ + n);
        if (s.x >= 0 && atomicSub(indeg + s.x, 1) == 1) a = s.x;
        if (s.y >= 0 && atomicSub(indeg + s.y, 1) == 1) b = s.y;
      }
      // CTA-aggregated append: warp scan -> shared -> one global atomic per CTA
      const unsigned k = (unsigned)(a >= 0) + (unsigned)(b >= 0);
      unsigned incl = k;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int w = 0; w < THREADS / 32; w++) {
          const unsigned t = s_warp[w];
          s_warp[w] = run;
          run += t;
        }
        s_cnt = run;
        if (run) s_base = atomicAdd(ncnt, run);
      }
      __syncthreads();
      if (k) {
        unsigned off = s_base + s_warp[warp] + incl - k;
        if (a >= 0) nxt[off++] = (uint32_t)a;
        if (b >= 0) nxt[off] = (uint32_t)b;
      }
      __syncthreads();  // s_warp / s_base are rewritten by the next pass
    }
    kahn_grid_barrier(&ctl->barrier, lvl * gridDim.x);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->num_levels = lvl - 1;
    ctl->max_level = maxc;
    ctl->processed = total;
  }
}

struct FlaggedBuild {
  int64_t num_levels = 0, max_level_size = 0, bad_index = -1;
  int64_t h2d_bytes = 0;
  int64_t kernel_launches = 0;
  double copy_ms = 0.0, levels_ms = 0.0, pack_ms = 0.0;
};

// Level pass on the host (the pre-Kahn builder): fills level / ku / kj for ratings in reference order.
inline bool build_flagged_host_levels(int32_t num_users, int32_t num_items, int64_t nnz, const int32_t* u, const int32_t* j,
                                      int32_t* level, int32_t* ku, int32_t* kj, int64_t* num_levels, int64_t* max_level) {
  struct Chain {
    int32_t level;
    uint32_t count;
  };
  try {
    std::vector<Chain> tu((size_t)num_users, Chain{0, 0u}), tj((size_t)num_items, Chain{0, 0u});
    std::vector<int64_t> level_count(1024, 0);
    int32_t nl = 0;
    for (int64_t n = 0; n < nnz; n++) {
      Chain& a = tu[(size_t)u[n]];
      Chain& b = tj[(size_t)j[n]];
      const int32_t l = 1 + (a.level > b.level ? a.level : b.level);
      a.level = l;
      b.level = l;
      if (l > nl) {
        nl = l;
        if ((size_t)l >= level_count.size()) level_count.resize((size_t)l * 2, 0);
      }
      level_count[l]++;
      level[n] = l;
      ku[n] = (int32_t)a.count++;
      kj[n] = (int32_t)b.count++;
    }
    *num_levels = nl;
    *max_level = 0;
    for (int32_t l = 1; l <= nl; l++)
      if (level_count[l] > *max_level) *max_level = level_count[l];
    return true;
  } catch (...) {
    return false;
  }
}

// Returns cudaSuccess, or the failing CUDA error; info->bad_index >= 0 when an id is out of range.
inline cudaError_t build_flagged_on_device(int32_t num_users, int32_t num_items, int32_t num_contexts, int64_t nnz,
                                           const int32_t* u, const int32_t* j, const int32_t* ctx, const double* r,
                                           cudaStream_t stream, int sm_count, StagedCopier& copier, RatingRec* d_rec,
                                           FlaggedBuild* info, const DevMem& mem, bool host_levels = false, bool trace = false) {
  if (nnz == 0) return cudaSuccess;
  // trace: host wall time of every phase, to stderr (tuning "sched_trace=1")
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[cars schedule] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  RatingSoA d;
  uint32_t *d_idx = nullptr, *d_ord = nullptr, *d_skey = nullptr, *d_start = nullptr, *d_fr0 = nullptr, *d_fr1 = nullptr;
  int32_t* d_succ = nullptr;
  RatingRec* d_tmp_rec = nullptr;
  int* d_indeg = nullptr;
  KahnCtl* d_ctl = nullptr;
  unsigned long long* d_bad = nullptr;
  void* d_temp = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaError_t e = cudaSuccess;
  char* arena = nullptr;  // every temporary below is carved from ONE allocation (cudaMalloc / cudaFree are not free)
  auto cleanup = [&]() {
    mem.free(arena);
    for (auto& x : ev)
      if (x) cudaEventDestroy(x);
  };
#define SG_TRY(x)              \
  do {                         \
    e = (x);                   \
    if (e != cudaSuccess) {    \
      cleanup();               \
      return e;                \
    }                          \
  } while (0)

  const size_t N = (size_t)nnz;
  const size_t ids = (size_t)(num_users > num_items ? num_users : num_items);
  const size_t fcap = (size_t)(num_users < num_items ? num_users : num_items);  // a level has distinct users and items
  for (auto& x : ev) SG_TRY(cudaEventCreate(&x));
  auto bits_for = [](int64_t n_values) {
    int b = 1;
    while ((1ll << b) < n_values) b++;
    return b;
  };
  size_t temp_bytes = 0;  // CUB scratch: the 32-bit-key sort bounds the three sorts below
  SG_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                         (uint32_t*)nullptr, nnz, 0, 32, stream));
  {
    size_t total = 0;
    auto reserve = [&](size_t bytes) {
      const size_t off = total;
      total += (bytes + 255) & ~(size_t)255;
      return off;
    };
    const size_t o_u = reserve(N * 4), o_j = reserve(N * 4), o_ctx = reserve(ctx ? N * 4 : 0), o_level = reserve(N * 4),
                 o_ku = reserve(N * 4), o_kj = reserve(N * 4), o_r = reserve(N * 8), o_idx = reserve(N * 4),
                 o_ord = reserve(N * 4), o_skey = reserve(N * 4), o_bad = reserve(8), o_temp = reserve(temp_bytes),
                 o_tmp = reserve(N * sizeof(RatingRec));
    size_t o_start = 0, o_succ = 0, o_indeg = 0, o_fr0 = 0, o_fr1 = 0, o_ctl = 0;
    if (!host_levels) {
      o_start = reserve(ids * 4); o_succ = reserve(N * 8); o_indeg = reserve(N * 4);
      o_fr0 = reserve((fcap + 1) * 4); o_fr1 = reserve((fcap + 1) * 4); o_ctl = reserve(sizeof(KahnCtl));
    }
    lap("size queries");
    SG_TRY(mem.alloc((void**)&arena, total));
    SG_TRY(cudaStreamSynchronize(stream));  // the copier fills the arena from its own streams
    lap("alloc(arena)");
    d.u = (int32_t*)(arena + o_u); d.j = (int32_t*)(arena + o_j); d.ctx = ctx ? (int32_t*)(arena + o_ctx) : nullptr;
    d.level = (int32_t*)(arena + o_level); d.ku = (int32_t*)(arena + o_ku); d.kj = (int32_t*)(arena + o_kj);
    d.r = (double*)(arena + o_r); d_idx = (uint32_t*)(arena + o_idx); d_ord = (uint32_t*)(arena + o_ord);
    d_skey = (uint32_t*)(arena + o_skey); d_bad = (unsigned long long*)(arena + o_bad); d_temp = arena + o_temp;
    d_tmp_rec = (RatingRec*)(arena + o_tmp);
    if (!host_levels) {
      d_start = (uint32_t*)(arena + o_start); d_succ = (int32_t*)(arena + o_succ); d_indeg = (int*)(arena + o_indeg);
      d_fr0 = (uint32_t*)(arena + o_fr0); d_fr1 = (uint32_t*)(arena + o_fr1); d_ctl = (KahnCtl*)(arena + o_ctl);
    }
  }

  // ---- the caller's arrays cross PCIe ------------------------------------------------------------------
  SG_TRY(cudaEventRecord(ev[0], stream));
  {
    CopySeg segs[4] = {{d.u, (void*)u, N * 4}, {d.j, (void*)j, N * 4}, {d.ctx, (void*)ctx, ctx ? N * 4 : 0}, {d.r, (void*)r, N * 8}};
    SG_TRY(copier.run(segs, 4, true));
  }
  info->h2d_bytes += nnz * (ctx ? 20 : 16);
  lap("H2D of u, j, ctx, r");
  SG_TRY(cudaEventRecord(ev[1], stream));
  const int blocks = sm_count * 8;
  SG_TRY(cudaMemsetAsync(d_bad, 0xff, 8, stream));
  validate_ids_kernel<<<blocks, 256, 0, stream>>>(d.u, d.j, d.ctx, nnz, (uint32_t)num_users, (uint32_t)num_items,
                                                  (uint32_t)num_contexts, d_bad);
  SG_TRY(cudaGetLastError());
  unsigned long long bad = 0;
  SG_TRY(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, stream));
  SG_TRY(cudaStreamSynchronize(stream));
  info->kernel_launches += 1;
  lap("validate + sync");
  if (bad != ~0ull) {
    info->bad_index = (int64_t)bad;
    cleanup();
    return cudaSuccess;
  }

  iota_kernel<<<blocks, 256, 0, stream>>>(d_idx, nnz);
  SG_TRY(cudaGetLastError());
  info->kernel_launches += 1;
  if (!host_levels) {
    SG_TRY(cudaMemsetAsync(d_ctl, 0, sizeof(KahnCtl), stream));
    // chains: stable sort by user, then by item
    const int bu = bits_for(num_users), bj = bits_for(num_items);
    for (int which = 0; which < 2; which++) {
      const uint32_t* keys = (const uint32_t*)(which == 0 ? d.u : d.j);
      size_t tb = temp_bytes;
      SG_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tb, keys, d_skey, d_idx, d_ord, nnz, 0, which == 0 ? bu : bj, stream));
      chain_heads_kernel<<<blocks, 256, 0, stream>>>(d_skey, nnz, d_start);
      SG_TRY(cudaGetLastError());
      chain_link_kernel<<<blocks, 256, 0, stream>>>(d_skey, d_ord, nnz, d_start, which == 0 ? d.ku : d.kj, d_succ, which);
      SG_TRY(cudaGetLastError());
      info->kernel_launches += 3;
    }
    kahn_roots_kernel<<<blocks, 256, 0, stream>>>(d.ku, d.kj, nnz, d_indeg, d_fr0, d_ctl);
    SG_TRY(cudaGetLastError());
    {
      constexpr int KT = 512;
      const int2* succ2 = reinterpret_cast<const int2*>(d_succ);
      void* args[] = {(void*)&succ2, (void*)&d_indeg, (void*)&d.level, (void*)&d_fr0, (void*)&d_fr1, (void*)&d_ctl};
      SG_TRY(cudaLaunchCooperativeKernel((const void*)kahn_levels_kernel<KT>, dim3(sm_count), dim3(KT), args, 0, stream));
    }
    info->kernel_launches += 2;
    KahnCtl ctl;
    SG_TRY(cudaMemcpyAsync(&ctl, d_ctl, sizeof ctl, cudaMemcpyDeviceToHost, stream));
    SG_TRY(cudaEventRecord(ev[2], stream));
    SG_TRY(cudaStreamSynchronize(stream));
    lap("chains + Kahn levels + sync");
    if ((int64_t)ctl.processed != nnz) {  // cannot happen for chains built above; refuse rather than train garbage
      cleanup();
      return cudaErrorUnknown;
    }
    info->num_levels = ctl.num_levels;
    info->max_level_size = ctl.max_level;
  } else {
    std::vector<int32_t> hl, hku, hkj;
    try {
      hl.resize(N); hku.resize(N); hkj.resize(N);
    } catch (...) {
      cleanup();
      return cudaErrorMemoryAllocation;
    }
    if (!build_flagged_host_levels(num_users, num_items, nnz, u, j, hl.data(), hku.data(), hkj.data(), &info->num_levels,
                                   &info->max_level_size)) {
      cleanup();
      return cudaErrorMemoryAllocation;
    }
    CopySeg segs[3] = {{d.level, hl.data(), N * 4}, {d.ku, hku.data(), N * 4}, {d.kj, hkj.data(), N * 4}};
    SG_TRY(copier.run(segs, 3, true));
    info->h2d_bytes += nnz * 12;
    SG_TRY(cudaEventRecord(ev[2], stream));
  }

  // ---- stable sort of (level, n), then pack the records in that order --------------------------------------
  {
    size_t tb = temp_bytes;
    SG_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tb, (const uint32_t*)d.level, d_skey, d_idx, d_ord, nnz, 0,
                                           bits_for(info->num_levels + 1), stream));
  }
  pack_recs_kernel<<<blocks, 256, 0, stream>>>(d, nnz, d_tmp_rec, host_levels ? nullptr : d_succ);
  SG_TRY(cudaGetLastError());
  gather_recs_kernel<<<blocks, 256, 0, stream>>>(d_tmp_rec, d_ord, nnz, d_rec);
  SG_TRY(cudaGetLastError());
  info->kernel_launches += 3;
  SG_TRY(cudaEventRecord(ev[3], stream));
  SG_TRY(cudaStreamSynchronize(stream));
  float ms = 0.f;
  lap("level sort + pack + sync");
  if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) info->copy_ms = ms;
  if (cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) info->levels_ms = ms;
  if (cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) info->pack_ms = ms;
  cleanup();
  lap("cudaFree(arena)");
#undef SG_TRY
  return cudaSuccess;
}

}  // namespace cars
