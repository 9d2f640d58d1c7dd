// fm_setup.cuh -- set-up of the FM sweep on the device: the engine's internal row order and, per field, the rows
// sorted by coordinate, cut into pieces (what fm_piece_reduce_kernel / fm_coord_kernel walk).
// The first version did this with counting sorts on one host thread (1.3 s at 25 M rows, more than six ALS
// iterations); here it is CUB radix sorts and a few streaming kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace carsfm {

__global__ void __launch_bounds__(256) fms_iota_kernel(uint32_t* v, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = (uint32_t)i;
}

// first row with an id out of range; *bad starts at ~0  (u < U, j < I, ctx >= 0 -- FM.java:81 tolerates ctx >= C)
__global__ void __launch_bounds__(256) fms_validate_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ j,
                                                           const int32_t* __restrict__ c, int64_t n, uint32_t U, uint32_t I,
                                                           unsigned long long* bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if ((uint32_t)u[i] >= U || (uint32_t)j[i] >= I || c[i] < 0) atomicMin(bad, (unsigned long long)i);
}

// storage-order key: (item block, context slot); ctx_slots == 1 leaves the context out
__global__ void __launch_bounds__(256) fms_row_key_kernel(const int32_t* __restrict__ j, const int32_t* __restrict__ c, int64_t n,
                                                          int32_t items_per_blk, int32_t ctx_slots, int32_t C,
                                                          uint32_t* __restrict__ key) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t slot = ctx_slots == 1 ? 0 : (c[i] < C ? c[i] : C);
    key[i] = (uint32_t)(j[i] / items_per_blk) * (uint32_t)ctx_slots + (uint32_t)slot;
  }
}

__global__ void __launch_bounds__(256) fms_permute_rows_kernel(const uint32_t* __restrict__ order, int64_t n,
                                                               const int32_t* __restrict__ u, const int32_t* __restrict__ j,
                                                               const int32_t* __restrict__ c, const double* __restrict__ r,
                                                               int32_t* __restrict__ pu, int32_t* __restrict__ pj,
                                                               int32_t* __restrict__ pc, double* __restrict__ pr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t k = order[i];
    pu[i] = u[k]; pj[i] = j[k]; pc[i] = c[k]; pr[i] = r[k];
  }
}

// second pass of the (item block, user) row order: the rows are already stably sorted by user (order1); key = item block
__global__ void __launch_bounds__(256) fms_block_key_kernel(const uint32_t* __restrict__ order1, const int32_t* __restrict__ j,
                                                            int64_t n, int32_t items_per_blk, uint32_t* __restrict__ key) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    key[i] = (uint32_t)(j[order1[i]] / items_per_blk);
}

// run table of fm_run_reduce_kernel: off[g * nblk + b] = first row (storage order, sorted by (item block, user)) whose
// (block, user) is not below (b, g * G); g = 0 .. ngroups inclusive, so that run (g, b) = [off[g][b], off[g + 1][b])
__global__ void __launch_bounds__(256) fms_run_offsets_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ j,
                                                              int64_t n, int32_t items_per_blk, int32_t nblk, int32_t G,
                                                              int32_t ngroups, uint32_t* __restrict__ off) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)(ngroups + 1) * nblk) return;
  const int32_t g = (int32_t)(t / nblk), b = (int32_t)(t % nblk);
  const int64_t ub = (int64_t)g * G;
  int64_t lo = 0, hi = n;  // first row with (block, user) >= (b, ub)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t mb = j[mid] / items_per_blk;
    const bool below = mb < b || (mb == b && (int64_t)u[mid] < ub);
    if (below) lo = mid + 1; else hi = mid;
  }
  off[t] = (uint32_t)lo;
}

// context coordinate of a row: its context id while the feature index stays below p (FM.java:81), else absent
__global__ void __launch_bounds__(256) fms_ctx_coord_kernel(const int32_t* __restrict__ c, int64_t n, int32_t C,
                                                            int32_t* __restrict__ coord) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) coord[i] = c[i] < C ? c[i] : -1;
}

// sort key of a row inside a field (rows without the feature last)
__global__ void __launch_bounds__(256) fms_field_key_kernel(const int32_t* __restrict__ coord, int64_t n, int32_t ncoord,
                                                            uint32_t* __restrict__ key) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t l = coord[i];
    key[i] = l >= 0 ? (uint32_t)l : (uint32_t)ncoord;
  }
}

// rows per coordinate from the SORTED keys: a run's first / last position (no atomics: a histogram with atomicAdd
// took 13.6 ms per field at 125 M rows, all of it on the 32 context counters)
__global__ void __launch_bounds__(256) fms_run_bounds_kernel(const uint32_t* __restrict__ skey, int64_t n, int32_t ncoord,
                                                             unsigned long long* __restrict__ run_beg,
                                                             unsigned long long* __restrict__ run_end) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const uint32_t k = skey[p];
    if (k >= (uint32_t)ncoord) continue;
    if (p == 0 || skey[p - 1] != k) run_beg[k] = (unsigned long long)p;
    if (p + 1 == n || skey[p + 1] != k) run_end[k] = (unsigned long long)p + 1;
  }
}

// rows and pieces per coordinate (slot ncoord = 0 so that the exclusive scans end with the totals)
__global__ void __launch_bounds__(256) fms_piece_count_kernel(const unsigned long long* __restrict__ run_beg,
                                                              const unsigned long long* __restrict__ run_end, int32_t ncoord,
                                                              int64_t piece, long long* __restrict__ rows_i64,
                                                              long long* __restrict__ npieces) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > ncoord) return;
  const long long r = l < ncoord ? (long long)(run_end[l] - run_beg[l]) : 0;
  rows_i64[l] = r;
  npieces[l] = (r + piece - 1) / piece;
}

// piece q belongs to the coordinate l with coord_piece[l] <= q < coord_piece[l + 1]
__global__ void __launch_bounds__(256) fms_piece_fill_kernel(const long long* __restrict__ coord_piece, const long long* __restrict__ start,
                                                             int32_t ncoord, int64_t num_pieces, int64_t piece, int64_t total,
                                                             long long* __restrict__ piece_beg, int32_t* __restrict__ piece_coord) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) piece_beg[num_pieces] = total;
  if (q >= num_pieces) return;
  int lo = 0, hi = ncoord;  // largest l with coord_piece[l] <= q
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (coord_piece[mid] <= q) lo = mid; else hi = mid;
  }
  piece_coord[q] = lo;
  piece_beg[q] = start[lo] + (q - coord_piece[lo]) * piece;
}

}  // namespace carsfm
