// fm_engine.cu -- C ABI of the FM recommender (cars_fm_* in include/carskit_b200.h) on the K4 kernels.
// No CPU implementation exists here: without a CUDA sm_100 device cars_fm_create() fails.
#include "../../include/carskit_b200.h"
#include "fm_kernels.cuh"
#include "fm_setup.cuh"
#include "dev_mem.cuh"
#include "staged_copy.cuh"
#include "tuning.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

using namespace carsfm;

static thread_local std::string g_fm_create_error = "";

struct FieldStore {
  bool owns_coord = true;  // users / items: coord_of_row IS the engine's u / j array
  int32_t *d_coord_of_row = nullptr, *d_perm = nullptr, *d_piece_coord = nullptr;
  int64_t *d_piece_beg = nullptr, *d_coord_piece = nullptr, *d_coord_rows = nullptr;
  double* d_delta = nullptr;
  uint32_t* d_run_off = nullptr;  // run mode (users): fm_run_reduce_kernel's table
  bool wide = false;  // few coordinates with many pieces each: one warp sums a coordinate's pieces
  bool short_pieces = false;  // on average < 64 rows per piece: 8 lanes per piece instead of a warp
  FmField f{};
};

struct cars_fm_handle {
  std::string err;
  cars::Tuning tune;  // developer knobs from cars_desc.tuning (no environment variable is read)
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int32_t U = 0, I = 0, C = 0, p = 0, k = 0, D = 1;
  int64_t N = 0, Nglobal = 0, Nq = 0;  // Nq: row stride of Qc (N rounded up to even: 16-byte aligned factor rows)
  double reg_lw = 0, reg_lf = 0, w0_denom = 1, xc = 1;
  int32_t *d_u = nullptr, *d_j = nullptr, *d_c = nullptr;
  double *d_r = nullptr, *d_e = nullptr, *d_Qc = nullptr;
  double *d_w0 = nullptr, *d_w = nullptr, *d_V = nullptr;
  double *d_part = nullptr, *d_scal = nullptr;
  double* h_scal = nullptr;
  int64_t max_pieces = 0;
  bool run_fuse_update = true;  // fm_run_reduce_kernel<MODE, 2>: the users' row update inside the reduce kernel
  int32_t row_items_per_blk = 0, row_blocks = 0;  // storage order = (item block, user) when row_blocks > 0
  int red_blocks = 0;
  FieldStore fld[3];
  bool uploaded = false, prepared = false;
  int ppg_short = 2, ppg_long = 1;  // pieces per lane group in fm_piece_reduce_kernel (CARS_FM_PPG_SHORT / _LONG)
  int64_t launches = 0, h2d = 0, d2h = 0;
  cudaEvent_t ev_beg = nullptr, ev_end = nullptr;
  double last_iter_ms = 0;
  cars::StagedCopier copier;
};

static int fm_fail(cars_fm_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_fm_create_error = buf;
  return code;
}

#define FM_TRY(h, expr)                                                                             \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fm_fail(h, _e == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA, "%s failed: %s", \
                     #expr, cudaGetErrorString(_e));                                                \
  } while (0)

template <typename T>
static cudaError_t fm_alloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), (n ? n : 1) * sizeof(T)); }

// One field on the device: rows sorted by coordinate (stable radix sort; rows without the feature last), cut into
// pieces of <= kPiece rows.  `coord` is the device array of the rows' coordinates (-1 = absent); `scratch_*` are
// caller-provided buffers of N uint32 each, `d_temp` CUB scratch.
static const int64_t kPiece = 256;
struct FmScratch {
  uint32_t *key = nullptr, *key_out = nullptr, *idx = nullptr;  // [N] each; idx = 0..N-1
  void* temp = nullptr;
  size_t temp_bytes = 0;
};
static int build_field(cars_fm_handle* h, int which, int32_t* coord, bool owns_coord, int32_t ncoord, int32_t offset, double x,
                       bool dense, const FmScratch& sc) {
  FieldStore& fs = h->fld[which];
  const int64_t N = h->N;
  const int blocks = h->sm_count * 8;
  fs.d_coord_of_row = coord;
  fs.owns_coord = owns_coord;
  unsigned long long *d_rows = nullptr, *d_rend = nullptr;  // first / one-past-last sorted position of every coordinate
  long long *d_np = nullptr, *d_start = nullptr;
  auto drop = [&]() { cudaFree(d_rows); cudaFree(d_rend); cudaFree(d_np); cudaFree(d_start); };
#define BF_TRY(expr)                                                                                    \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      drop();                                                                                           \
      return fm_fail(h, _e == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA, "%s failed: %s", #expr, \
                     cudaGetErrorString(_e));                                                           \
    }                                                                                                   \
  } while (0)
  const size_t nc1 = (size_t)ncoord + 1;
  BF_TRY(fm_alloc(&d_rows, nc1));
  BF_TRY(fm_alloc(&d_rend, nc1));
  BF_TRY(cudaMemsetAsync(d_rend, 0, nc1 * 8, h->stream));
  BF_TRY(fm_alloc(&d_np, nc1));
  BF_TRY(fm_alloc(&d_start, nc1));
  BF_TRY(fm_alloc(&fs.d_coord_rows, nc1));
  BF_TRY(fm_alloc(&fs.d_coord_piece, nc1));
  BF_TRY(fm_alloc(&fs.d_perm, (size_t)N));
  BF_TRY(fm_alloc(&fs.d_delta, (size_t)ncoord));
  BF_TRY(cudaMemsetAsync(d_rows, 0, nc1 * 8, h->stream));
  long long tot[2] = {0, 0};  // rows that carry the feature, pieces
  if (N > 0) {
    fms_field_key_kernel<<<blocks, 256, 0, h->stream>>>(coord, N, ncoord, sc.key);
    BF_TRY(cudaGetLastError());
    int bits = 1;
    while ((1ll << bits) < (long long)ncoord + 1) bits++;
    size_t tb = sc.temp_bytes;
    BF_TRY(cub::DeviceRadixSort::SortPairs(sc.temp, tb, sc.key, sc.key_out, sc.idx, reinterpret_cast<uint32_t*>(fs.d_perm), N, 0,
                                           bits, h->stream));
    fms_run_bounds_kernel<<<blocks, 256, 0, h->stream>>>(sc.key_out, N, ncoord, d_rows, d_rend);
    BF_TRY(cudaGetLastError());
    h->launches += 3;
  }
  fms_piece_count_kernel<<<(unsigned)((nc1 + 255) / 256), 256, 0, h->stream>>>(d_rows, d_rend, ncoord, kPiece,
                                                                              reinterpret_cast<long long*>(fs.d_coord_rows), d_np);
  BF_TRY(cudaGetLastError());
  {
    size_t tb = sc.temp_bytes;
    BF_TRY(cub::DeviceScan::ExclusiveSum(sc.temp, tb, reinterpret_cast<long long*>(fs.d_coord_rows), d_start, (int)nc1, h->stream));
    tb = sc.temp_bytes;
    BF_TRY(cub::DeviceScan::ExclusiveSum(sc.temp, tb, d_np, reinterpret_cast<long long*>(fs.d_coord_piece), (int)nc1, h->stream));
  }
  BF_TRY(cudaMemcpyAsync(&tot[0], d_start + ncoord, 8, cudaMemcpyDeviceToHost, h->stream));
  BF_TRY(cudaMemcpyAsync(&tot[1], fs.d_coord_piece + ncoord, 8, cudaMemcpyDeviceToHost, h->stream));
  BF_TRY(cudaStreamSynchronize(h->stream));
  h->launches += 3;
  const int64_t total = tot[0], num_pieces = dense ? 0 : tot[1];
  BF_TRY(fm_alloc(&fs.d_piece_beg, (size_t)num_pieces + 1));
  BF_TRY(fm_alloc(&fs.d_piece_coord, (size_t)num_pieces));
  fms_piece_fill_kernel<<<(unsigned)((num_pieces + 1 + 255) / 256), 256, 0, h->stream>>>(
      reinterpret_cast<long long*>(fs.d_coord_piece), d_start, ncoord, num_pieces, kPiece, total,
      reinterpret_cast<long long*>(fs.d_piece_beg), fs.d_piece_coord);
  BF_TRY(cudaGetLastError());
  BF_TRY(cudaStreamSynchronize(h->stream));
  h->launches++;
  drop();
#undef BF_TRY
  fs.f.coord_of_row = fs.d_coord_of_row; fs.f.perm = fs.d_perm; fs.f.piece_beg = fs.d_piece_beg;
  fs.f.piece_coord = fs.d_piece_coord; fs.f.coord_piece = fs.d_coord_piece; fs.f.coord_rows = fs.d_coord_rows;
  fs.f.num_pieces = num_pieces; fs.f.ncoord = ncoord; fs.f.offset = offset; fs.f.x = x;
  fs.f.dense_blocks = dense ? h->sm_count * 3 : 0;
  fs.wide = dense || (ncoord > 0 && fs.f.num_pieces / ncoord >= 8);
  fs.short_pieces = num_pieces > 0 && total / num_pieces < 64;
  if (const char* e = h->tune.get("fm_lanes_per_piece")) fs.short_pieces = atoi(e) == 8;
  if (fs.f.num_pieces > h->max_pieces) h->max_pieces = fs.f.num_pieces;
  return CARS_OK;
}

extern "C" int cars_fm_create(const cars_desc* d, cars_fm_handle** out) {
  if (!out) return fm_fail(nullptr, CARS_E_INVALID, "out is NULL");
  *out = nullptr;
  if (!d) return fm_fail(nullptr, CARS_E_INVALID, "desc is NULL");
  if (d->abi_version != CARS_ABI_VERSION) return fm_fail(nullptr, CARS_E_INVALID, "abi_version mismatch");
  if (d->model != CARS_FM) return fm_fail(nullptr, CARS_E_INVALID, "cars_fm_create needs model = CARS_FM");
  if (d->num_users <= 0 || d->num_items <= 0 || d->num_conditions < 0 || d->num_factors <= 0 || d->nnz < 0)
    return fm_fail(nullptr, CARS_E_INVALID, "bad sizes");
  if (d->num_context_dims <= 0) return fm_fail(nullptr, CARS_E_INVALID, "num_context_dims must be > 0 (rateDao.numContextDims())");
  if (d->nnz > 0 && (!d->u || !d->j || !d->ctx || !d->r)) return fm_fail(nullptr, CARS_E_INVALID, "u/j/ctx/r must not be NULL");
  if (d->nnz >= (1ll << 31)) return fm_fail(nullptr, CARS_E_UNSUPPORTED, "nnz >= 2^31 rows per handle");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) return fm_fail(nullptr, CARS_E_NO_DEVICE, "no CUDA device; this engine has no CPU path");
  if (d->device < 0 || d->device >= ndev) return fm_fail(nullptr, CARS_E_INVALID, "device %d out of range", d->device);
  cars::DeviceFacts prop;
  if (cars::device_facts(d->device, &prop) != cudaSuccess || prop.major != 10)
    return fm_fail(nullptr, CARS_E_NO_DEVICE, "device %d is not sm_100", d->device);

  cars_fm_handle* h = new (std::nothrow) cars_fm_handle();
  if (!h) return fm_fail(nullptr, CARS_E_OOM, "host allocation failed");
  auto bail = [&](int code) { g_fm_create_error = h->err; cars_fm_destroy(h); return code; };
  h->tune = cars::Tuning(d->tuning);
  h->device = d->device; h->sm_count = prop.sm_count;
  h->U = d->num_users; h->I = d->num_items; h->C = d->num_conditions; h->p = h->U + h->I + h->C;
  h->k = d->num_factors; h->D = d->num_context_dims; h->N = d->nnz;
  h->Nglobal = d->global_nnz > 0 ? d->global_nnz : d->nnz;
  h->reg_lw = d->reg_lw; h->reg_lf = d->reg_lf;
  h->xc = 1.0 / (double)h->D;                                   // FM.java:86
  h->w0_denom = (double)((float)h->Nglobal + (float)d->reg_lw);  // FM.java:159: int + float is a FLOAT addition
  int rc = CARS_OK;
#define FM_TRY_H(expr)                                      \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) {                                \
      fm_fail(h, CARS_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
      return bail(_e == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA);  \
    }                                                       \
  } while (0)
  FM_TRY_H(cudaSetDevice(h->device));
  if (d->stream) h->stream = (cudaStream_t)d->stream;
  else { FM_TRY_H(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  FM_TRY_H(cudaEventCreate(&h->ev_beg));
  FM_TRY_H(cudaEventCreate(&h->ev_end));
  const int64_t N = h->N;
  FM_TRY_H(fm_alloc(&h->d_u, (size_t)N)); FM_TRY_H(fm_alloc(&h->d_j, (size_t)N)); FM_TRY_H(fm_alloc(&h->d_c, (size_t)N));
  FM_TRY_H(fm_alloc(&h->d_r, (size_t)N)); FM_TRY_H(fm_alloc(&h->d_e, (size_t)N));
  h->Nq = (N + 1) & ~(int64_t)1;
  FM_TRY_H(fm_alloc(&h->d_Qc, (size_t)h->Nq * h->k));
  FM_TRY_H(fm_alloc(&h->d_w0, 1)); FM_TRY_H(fm_alloc(&h->d_w, (size_t)h->p)); FM_TRY_H(fm_alloc(&h->d_V, (size_t)h->p * h->k));
  FM_TRY_H(h->copier.init(h->device, (int)h->tune.get_ll("copy_threads", 0)));
  if (const char* e = h->tune.get("fm_ppg_short")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) h->ppg_short = v; }
  if (const char* e = h->tune.get("fm_ppg_long")) { const int v = atoi(e); if (v == 1 || v == 2) h->ppg_long = v; }
  {
    // Internal row order.  The rows (errors[n], Q[n][f]) are private to the engine, and every sum of the sweep
    // runs over the rows of ONE coordinate, so the rows may be stored in any order.  They are sorted by item
    // block (and by context when the context field is reduced by pieces), the caller's order inside: the item
    // field then gathers e / Qc[f] inside one block of ~block_rows rows (L2-resident) and the user field reads
    // short contiguous runs in which neighbouring users are neighbours -- instead of 8-byte reads scattered over
    // all N rows, one DRAM sector each (profiles/r1: 0.31 of the HBM roofline before).
    int64_t block_rows = 1536 * 1024;
    block_rows = h->tune.get_ll("fm_block_rows", block_rows);
    // few contexts: that field is reduced by streaming the rows (fm_dense_reduce_kernel), so the context does not
    // enter the row order and a user's rows stay contiguous inside every item block
    int64_t dense_min_rows = 65536;
    dense_min_rows = h->tune.get_ll("fm_dense_min_rows", dense_min_rows);
    const bool ctx_dense = h->C > 0 && h->C <= kDenseMaxCoord && N >= dense_min_rows;
    const int blocks = h->sm_count * 8;
    const size_t Na = (size_t)(N ? N : 1);
    int32_t *r_u = nullptr, *r_j = nullptr, *r_c = nullptr, *d_cc = nullptr;
    double* r_r = nullptr;
    unsigned long long* d_bad = nullptr;
    FmScratch sc;
    auto drop = [&]() {
      cudaFree(r_u); cudaFree(r_j); cudaFree(r_c); cudaFree(r_r); cudaFree(d_bad);
      cudaFree(sc.key); cudaFree(sc.key_out); cudaFree(sc.idx); cudaFree(sc.temp);
    };
#define FS_TRY(expr)                                                           \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      drop();                                                                  \
      fm_fail(h, CARS_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
      return bail(_e == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA); \
    }                                                                          \
  } while (0)
    FS_TRY(fm_alloc(&sc.key, Na)); FS_TRY(fm_alloc(&sc.key_out, Na)); FS_TRY(fm_alloc(&sc.idx, Na));
    FS_TRY(fm_alloc(&d_bad, 1));
    {
      size_t a = 0, b = 0;
      FS_TRY(cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                             (uint32_t*)nullptr, N, 0, 32, h->stream));
      int64_t m = h->U > h->I ? h->U : h->I;
      if (h->C > m) m = h->C;
      FS_TRY(cub::DeviceScan::ExclusiveSum(nullptr, b, (long long*)nullptr, (long long*)nullptr, (int)(m + 1), h->stream));
      sc.temp_bytes = a > b ? a : b;
      FS_TRY(cudaMalloc(&sc.temp, sc.temp_bytes ? sc.temp_bytes : 1));
    }
    if (N) {
      // the caller's arrays go straight into the engine's arrays (no row permutation) or into staging copies
      const bool permute = block_rows > 0 && N > block_rows;
      int32_t *t_u = h->d_u, *t_j = h->d_j, *t_c = h->d_c;
      double* t_r = h->d_r;
      if (permute) {
        FS_TRY(fm_alloc(&r_u, Na)); FS_TRY(fm_alloc(&r_j, Na)); FS_TRY(fm_alloc(&r_c, Na)); FS_TRY(fm_alloc(&r_r, Na));
        t_u = r_u; t_j = r_j; t_c = r_c; t_r = r_r;
      }
      cars::CopySeg segs[4] = {{t_u, (void*)d->u, (size_t)N * 4}, {t_j, (void*)d->j, (size_t)N * 4},
                               {t_c, (void*)d->ctx, (size_t)N * 4}, {t_r, (void*)d->r, (size_t)N * 8}};
      FS_TRY(h->copier.run(segs, 4, true));
      h->h2d += N * 20;
      FS_TRY(cudaMemsetAsync(d_bad, 0xff, 8, h->stream));
      fms_validate_kernel<<<blocks, 256, 0, h->stream>>>(t_u, t_j, t_c, N, (uint32_t)h->U, (uint32_t)h->I, d_bad);
      FS_TRY(cudaGetLastError());
      unsigned long long bad = 0;
      FS_TRY(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, h->stream));
      FS_TRY(cudaStreamSynchronize(h->stream));
      if (bad != ~0ull) {
        drop();
        fm_fail(h, CARS_E_INVALID, "rating %lld has an id out of range", (long long)bad);
        return bail(CARS_E_INVALID);
      }
      fms_iota_kernel<<<blocks, 256, 0, h->stream>>>(sc.idx, N);
      FS_TRY(cudaGetLastError());
      h->launches += 2;
      if (permute) {
        const int64_t nblk = (N + block_rows - 1) / block_rows;
        const int32_t items_per_blk = (int32_t)((h->I + nblk - 1) / nblk);
        const int32_t ctx_slots = ctx_dense ? 1 : h->C + 1;  // contexts without a feature (index >= p) share the last slot
        const int64_t nkeys = ((int64_t)(h->I - 1) / items_per_blk + 1) * ctx_slots;
        int bits = 1;
        while ((1ll << bits) < nkeys) bits++;
        if (bits > 32) {
          drop();
          fm_fail(h, CARS_E_UNSUPPORTED, "row-order key does not fit 32 bits");
          return bail(CARS_E_UNSUPPORTED);
        }
        // contexts streamed (ctx_slots == 1): the rows of a block are put in USER order (stable: the caller's order
        // inside a user), whatever the caller's order was -- fm_run_reduce_kernel needs the rows of consecutive users
        // inside a block to be one contiguous run.  Two stable passes: by user, then by item block.
        const bool by_user = ctx_slots == 1 && h->tune.get_ll("fm_runs", 1) != 0;
        uint32_t *order = nullptr, *order1 = nullptr;
        FS_TRY(fm_alloc(&order, Na));
        cudaError_t se = cudaSuccess;
        size_t tb = sc.temp_bytes;
        if (by_user) {
          se = fm_alloc(&order1, Na);
          int ubits = 1;
          while ((1ll << ubits) < (long long)h->U) ubits++;
          if (se == cudaSuccess)
            se = cub::DeviceRadixSort::SortPairs(sc.temp, tb, reinterpret_cast<const uint32_t*>(r_u), sc.key_out, sc.idx, order1, N, 0,
                                                 ubits, h->stream);
          if (se == cudaSuccess) {
            fms_block_key_kernel<<<blocks, 256, 0, h->stream>>>(order1, r_j, N, items_per_blk, sc.key);
            se = cudaGetLastError();
          }
          tb = sc.temp_bytes;
          if (se == cudaSuccess)
            se = cub::DeviceRadixSort::SortPairs(sc.temp, tb, sc.key, sc.key_out, order1, order, N, 0, bits, h->stream);
          h->launches += 2;
          h->row_items_per_blk = items_per_blk;
          h->row_blocks = (int32_t)((h->I - 1) / items_per_blk + 1);
        } else {
          fms_row_key_kernel<<<blocks, 256, 0, h->stream>>>(r_j, r_c, N, items_per_blk, ctx_slots, h->C, sc.key);
          se = cudaGetLastError();
          if (se == cudaSuccess)
            se = cub::DeviceRadixSort::SortPairs(sc.temp, tb, sc.key, sc.key_out, sc.idx, order, N, 0, bits, h->stream);
        }
        if (se == cudaSuccess) {
          fms_permute_rows_kernel<<<blocks, 256, 0, h->stream>>>(order, N, r_u, r_j, r_c, r_r, h->d_u, h->d_j, h->d_c, h->d_r);
          se = cudaGetLastError();
        }
        if (se == cudaSuccess) se = cudaStreamSynchronize(h->stream);
        cudaFree(order);
        cudaFree(order1);
        FS_TRY(se);
        h->launches += 3;
      }
    }
    // context coordinate: the context id while its feature index is < p (FM.java:81)
    FS_TRY(fm_alloc(&d_cc, Na));
    if (N) {
      fms_ctx_coord_kernel<<<blocks, 256, 0, h->stream>>>(h->d_c, N, h->C, d_cc);
      FS_TRY(cudaGetLastError());
      h->launches++;
    }
    rc = build_field(h, 0, h->d_u, false, h->U, 0, 1.0, false, sc);
    if (!rc) rc = build_field(h, 1, h->d_j, false, h->I, h->U, 1.0, false, sc);
    if (!rc) rc = build_field(h, 2, d_cc, true, h->C, h->U + h->I, h->xc, ctx_dense, sc);  // takes d_cc over
    else cudaFree(d_cc);
    drop();
    if (rc) return bail(rc);
#undef FS_TRY
    // run mode for the user field (fm_run_reduce_kernel): G consecutive users per CTA, G chosen so that a run -- the
    // group's rows inside one item block -- is 24..48 rows (measured at 125 M rows, 25 rows per user: G = 128 / 256 / 512 ->
    // 291.8 / 299.5 / 296.6 ms per iteration; shorter-lived CTAs leave more of their rows in L2 for the fused row update)
    if (h->row_blocks >= (int32_t)h->tune.get_ll("fm_run_min_blocks", 8) && N > 0) {
      FieldStore& fs = h->fld[0];
      const double rows_per_user_block = (double)N / ((double)h->U * (double)h->row_blocks);
      int G = 32;
      while (G < 512 && (double)(2 * G) * rows_per_user_block <= 48.0) G *= 2;
      const long long tg = h->tune.get_ll("fm_run_users", 0);
      if (tg == 32 || tg == 64 || tg == 128 || tg == 256 || tg == 512) G = (int)tg;
      const int32_t ngroups = (h->U + G - 1) / G;
      const int64_t entries = (int64_t)(ngroups + 1) * h->row_blocks;
      FM_TRY_H(fm_alloc(&fs.d_run_off, (size_t)entries));
      fms_run_offsets_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, h->stream>>>(h->d_u, h->d_j, N, h->row_items_per_blk,
                                                                                       h->row_blocks, G, ngroups, fs.d_run_off);
      FM_TRY_H(cudaGetLastError());
      h->launches++;
      fs.f.run_off = fs.d_run_off; fs.f.run_group = G; fs.f.run_groups = ngroups; fs.f.run_blocks = h->row_blocks;
      const int smem = (int)fm_run_smem_bytes(G);
      FM_TRY_H(cudaFuncSetAttribute(fm_run_reduce_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      FM_TRY_H(cudaFuncSetAttribute(fm_run_reduce_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      FM_TRY_H(cudaFuncSetAttribute(fm_run_reduce_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      FM_TRY_H(cudaFuncSetAttribute(fm_run_reduce_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      FM_TRY_H(cudaFuncSetAttribute(fm_run_reduce_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      FM_TRY_H(cudaFuncSetAttribute(fm_run_reduce_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      h->run_fuse_update = h->tune.get_ll("fm_run_fuse_update", 1) != 0;
      if ((int64_t)h->U > h->max_pieces) h->max_pieces = h->U;  // d_part holds one pair per user
    }
    if (ctx_dense) {
      const int smem = (int)((size_t)h->C * kDenseThreads * sizeof(double2));
      FM_TRY_H(cudaFuncSetAttribute(fm_dense_reduce_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      FM_TRY_H(cudaFuncSetAttribute(fm_dense_reduce_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
  }
  h->red_blocks = h->sm_count * 4;
  size_t part = (size_t)(2 * (h->max_pieces > h->red_blocks ? h->max_pieces : h->red_blocks));
  for (const FieldStore& fs : h->fld) {
    const size_t need = (size_t)2 * fs.f.dense_blocks * (fs.f.ncoord > 0 ? fs.f.ncoord : 0);
    if (need > part) part = need;
  }
  FM_TRY_H(fm_alloc(&h->d_part, part));
  FM_TRY_H(fm_alloc(&h->d_scal, 8));
  FM_TRY_H(cudaMallocHost((void**)&h->h_scal, 8 * sizeof(double)));
  FM_TRY_H(cudaStreamSynchronize(h->stream));
  *out = h;
  return CARS_OK;
#undef FM_TRY_H
}

static int fm_transfer(cars_fm_handle* h, const cars_fm_arrays* a, bool up) {
  if (!h) return CARS_E_INVALID;
  if (!a || !a->w0 || !a->w || !a->V) return fm_fail(h, CARS_E_INVALID, "w0, w and V are required");
  FM_TRY(h, cudaSetDevice(h->device));
  const cudaMemcpyKind kd = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  auto cp = [&](double* dev, double* host, size_t n) {
    return up ? cudaMemcpyAsync(dev, host, n * 8, kd, h->stream) : cudaMemcpyAsync(host, dev, n * 8, kd, h->stream);
  };
  FM_TRY(h, cp(h->d_w0, a->w0, 1));
  FM_TRY(h, cp(h->d_w, a->w, (size_t)h->p));
  // the caller's V is [p x k] row-major, the engine's is factor-major: transpose through a staging buffer
  const size_t vn = (size_t)h->p * h->k;
  double* d_stage = nullptr;
  FM_TRY(h, fm_alloc(&d_stage, vn));
  cudaError_t te = cudaSuccess;
  if (up) {
    te = h->copier.h2d(d_stage, a->V, vn * 8);
    if (te == cudaSuccess) {
      const dim3 grid((unsigned)((h->p + 31) / 32), (unsigned)((h->k + 31) / 32));
      fm_transpose_kernel<<<grid, 256, 0, h->stream>>>(d_stage, h->p, h->k, h->d_V);
      te = cudaGetLastError();
    }
    if (te == cudaSuccess) te = cudaStreamSynchronize(h->stream);
  } else {
    // d_V as a [k x p] row-major matrix -> [p x k]; the grid's x dimension runs over the long side (p)
    te = cudaStreamSynchronize(h->stream);
    if (te == cudaSuccess) {
      const dim3 grid((unsigned)((h->p + 31) / 32), (unsigned)((h->k + 31) / 32));
      fm_untranspose_kernel<<<grid, 256, 0, h->stream>>>(h->d_V, h->p, h->k, d_stage);
      te = cudaGetLastError();
    }
    if (te == cudaSuccess) te = cudaStreamSynchronize(h->stream);
    if (te == cudaSuccess) te = h->copier.d2h(a->V, d_stage, vn * 8);
  }
  cudaFree(d_stage);
  h->launches++;
  FM_TRY(h, te);
  FM_TRY(h, cudaStreamSynchronize(h->stream));
  (up ? h->h2d : h->d2h) += (int64_t)(1 + h->p + (int64_t)h->p * h->k) * 8;
  return CARS_OK;
}

extern "C" int cars_fm_upload(cars_fm_handle* h, const cars_fm_arrays* a) {
  int rc = fm_transfer(h, a, true);
  if (rc == CARS_OK) { h->uploaded = true; h->prepared = false; }
  return rc;
}
extern "C" int cars_fm_download(cars_fm_handle* h, const cars_fm_arrays* a) {
  if (h && !h->uploaded) return fm_fail(h, CARS_E_STATE, "cars_fm_download before cars_fm_upload");
  return fm_transfer(h, a, false);
}

extern "C" int cars_fm_prepare(cars_fm_handle* h) {
  if (!h) return CARS_E_INVALID;
  if (!h->uploaded) return fm_fail(h, CARS_E_STATE, "cars_fm_prepare before cars_fm_upload");
  FM_TRY(h, cudaSetDevice(h->device));
  const bool tiled = h->tune.get_ll("fm_prepare_tiled", h->N >= 65536 ? 1 : 0) != 0;
  if (h->N && tiled) {
    // coordinate-major scratch copy of V (a row's coefficient vectors contiguous), then fm_prepare_tiled_kernel
    double* d_vt = nullptr;
    FM_TRY(h, fm_alloc(&d_vt, (size_t)h->p * h->k));
    const dim3 tg((unsigned)((h->p + 31) / 32), (unsigned)((h->k + 31) / 32));
    fm_untranspose_kernel<<<tg, 256, 0, h->stream>>>(h->d_V, h->p, h->k, d_vt);
    const int smem = kPrepWarps * 3 * kPrepTile * (int)sizeof(double);
    cudaError_t pe = cudaFuncSetAttribute(fm_prepare_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (pe == cudaSuccess) {
      const int64_t want = (h->N + kPrepWarps * 32 - 1) / (kPrepWarps * 32);
      const int64_t cap = (int64_t)h->sm_count * 4 * 8;
      fm_prepare_tiled_kernel<<<(unsigned)(want < cap ? want : cap), kPrepWarps * 32, smem, h->stream>>>(
          h->d_u, h->d_j, h->d_c, h->d_r, h->d_w, d_vt, h->d_w0, h->U, h->I, h->p, h->k, h->xc, h->N, h->Nq, h->d_e, h->d_Qc);
      pe = cudaGetLastError();
    }
    if (pe == cudaSuccess) pe = cudaStreamSynchronize(h->stream);
    cudaFree(d_vt);
    FM_TRY(h, pe);
    h->launches += 2;
  } else if (h->N) {
    const unsigned blocks = (unsigned)((h->N + 255) / 256);
    fm_prepare_kernel<<<blocks, 256, 0, h->stream>>>(h->d_u, h->d_j, h->d_c, h->d_r, h->d_w, h->d_V, h->d_w0, h->U, h->I,
                                                    h->p, h->k, h->xc, h->N, h->Nq, h->d_e, h->d_Qc);
    FM_TRY(h, cudaGetLastError());
    h->launches++;
  }
  FM_TRY(h, cudaStreamSynchronize(h->stream));
  h->prepared = true;
  return CARS_OK;
}

// One sweep over the three fields (users, items, contexts) for one coefficient column: per field
// (a) reduce -> (b) new coordinates + deltas [all-reduce of the per-coordinate sums in between when sharded] ->
// (c) row update.  (Fusing (c) into the next field's streaming reduce was measured slower -- 416 us against
// 189 + 158 us, profiles/r1/launches_r1b_fm_fused.txt -- and is not kept.)
struct FmExchange {  // row-sharded run: caller's device buffer and all-reduce (cars_fm_iteration_sharded)
  double* dev_buf = nullptr;
  cars_allreduce_fn ar = nullptr;
  void* user = nullptr;
};

template <int MODE>
static int field_sweep(cars_fm_handle* h, double* coef, int stride, int col, double* Qf, double size_reg, const FmExchange* ex) {
  for (int which = 0; which < 3; which++) {
    FieldStore& fs = h->fld[which];
    const FmField& f = fs.f;
    if (f.ncoord == 0) continue;
    // (a)
    bool coord_done = false, rows_done = false;
    if (f.run_off) {
      const size_t smem = fm_run_smem_bytes(f.run_group);
      if (!ex && h->run_fuse_update) {
        fm_run_reduce_kernel<MODE, 2><<<f.run_groups, 256, smem, h->stream>>>(f, h->d_e, Qf, coef, stride, col, size_reg, h->d_part,
                                                                             fs.d_delta, h->d_e, Qf);
        coord_done = rows_done = true;  // steps (b) and (c) ran inside the kernel
      } else if (!ex) {
        fm_run_reduce_kernel<MODE, 1><<<f.run_groups, 256, smem, h->stream>>>(f, h->d_e, Qf, coef, stride, col, size_reg, h->d_part,
                                                                             fs.d_delta, nullptr, nullptr);
        coord_done = true;  // step (b) ran in the kernel's epilogue
      } else {
        fm_run_reduce_kernel<MODE, 0><<<f.run_groups, 256, smem, h->stream>>>(f, h->d_e, Qf, coef, stride, col, size_reg, h->d_part,
                                                                             fs.d_delta, nullptr, nullptr);
      }
      FM_TRY(h, cudaGetLastError());
      h->launches++;
    } else if (f.dense_blocks > 0) {
      const size_t smem = (size_t)f.ncoord * kDenseThreads * sizeof(double2);
      fm_dense_reduce_kernel<MODE><<<f.dense_blocks, kDenseThreads, smem, h->stream>>>(f, h->d_e, Qf, coef, stride, col, h->N,
                                                                                      h->d_part);
      FM_TRY(h, cudaGetLastError());
      h->launches++;
    } else if (f.num_pieces > 0) {
      const int ppg = fs.short_pieces ? h->ppg_short : h->ppg_long;
      const int lpp = fs.short_pieces ? 8 : 32;
      const int64_t groups = (f.num_pieces + ppg - 1) / ppg;
      const unsigned blocks = (unsigned)((groups * lpp + 255) / 256);
#define PIECE_LAUNCH(L, P) fm_piece_reduce_kernel<MODE, L, P><<<blocks, 256, 0, h->stream>>>(f, h->d_e, Qf, coef, stride, col, h->d_part)
      if (fs.short_pieces) {
        if (ppg == 4) PIECE_LAUNCH(8, 4); else if (ppg == 2) PIECE_LAUNCH(8, 2); else PIECE_LAUNCH(8, 1);
      } else {
        if (ppg == 2) PIECE_LAUNCH(32, 2); else PIECE_LAUNCH(32, 1);
      }
#undef PIECE_LAUNCH
      FM_TRY(h, cudaGetLastError());
      h->launches++;
    }
    // (b)
    const unsigned cb = fs.wide ? (unsigned)(((int64_t)f.ncoord * 32 + 255) / 256) : (unsigned)((f.ncoord + 255) / 256);
    if (coord_done) {
    } else if (!ex) {
      if (fs.wide)
        fm_coord_kernel<MODE, true><<<cb, 256, 0, h->stream>>>(f, h->d_part, size_reg, coef, stride, col, fs.d_delta);
      else
        fm_coord_kernel<MODE, false><<<cb, 256, 0, h->stream>>>(f, h->d_part, size_reg, coef, stride, col, fs.d_delta);
      FM_TRY(h, cudaGetLastError());
      h->launches++;
    } else {
      if (fs.wide)
        fm_coord_partial_kernel<MODE, true><<<cb, 256, 0, h->stream>>>(f, h->d_part, ex->dev_buf);
      else
        fm_coord_partial_kernel<MODE, false><<<cb, 256, 0, h->stream>>>(f, h->d_part, ex->dev_buf);
      FM_TRY(h, cudaGetLastError());
      if (ex->ar(ex->user, ex->dev_buf, 2 * (int64_t)f.ncoord) != 0) return fm_fail(h, CARS_E_STATE, "the all-reduce callback failed");
      fm_coord_finish_kernel<<<(unsigned)((f.ncoord + 255) / 256), 256, 0, h->stream>>>(f, ex->dev_buf, size_reg, coef, stride, col,
                                                                                       fs.d_delta);
      FM_TRY(h, cudaGetLastError());
      h->launches += 2;
    }
    // (c)
    if (h->N == 0 || rows_done) continue;
    fm_row_update_kernel<MODE><<<(unsigned)(((h->N + 1) / 2 + 255) / 256), 256, 0, h->stream>>>(f, fs.d_delta, h->N, h->d_e, Qf);
    FM_TRY(h, cudaGetLastError());
    h->launches++;
  }
  return CARS_OK;
}

extern "C" int cars_fm_iteration(cars_fm_handle* h, double* loss_out) {
  if (!h) return CARS_E_INVALID;
  if (!h->prepared) return fm_fail(h, CARS_E_STATE, "cars_fm_iteration before cars_fm_prepare");
  FM_TRY(h, cudaSetDevice(h->device));
  FM_TRY(h, cudaEventRecord(h->ev_beg, h->stream));
  int rc;
  // w0 (FM.java:152-169)
  fm_w0_reduce_kernel<<<h->red_blocks, 256, 0, h->stream>>>(h->d_e, h->d_w0, h->N, h->d_part);
  fm_w0_finish_kernel<<<1, 32, 0, h->stream>>>(h->d_part, h->red_blocks, h->w0_denom, h->reg_lw, h->d_w0, h->d_scal);
  if (h->N) fm_w0_apply_kernel<<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(h->d_e, h->d_scal, h->N);
  fm_wreg_kernel<<<1, 256, 0, h->stream>>>(h->d_w, h->p, h->reg_lw, h->d_scal + 3);  // uses the OLD w (:187)
  FM_TRY(h, cudaGetLastError());
  h->launches += 4;
  // w_l, l = 0..p-1 (:172-191): users, items, contexts
  const double w_reg = (double)h->Nglobal * h->reg_lw;
  if ((rc = field_sweep<0>(h, h->d_w, 1, 0, nullptr, w_reg, nullptr))) return rc;
  // V_lf, f = 0..k-1 { l = 0..p-1 } (:194-217)
  const double v_reg = (double)h->Nglobal * h->reg_lf;
  for (int f = 0; f < h->k; f++)
    if ((rc = field_sweep<1>(h, h->d_V + (int64_t)f * h->p, 1, 0, h->d_Qc + (int64_t)f * h->Nq, v_reg, nullptr))) return rc;
  FM_TRY(h, cudaEventRecord(h->ev_end, h->stream));
  FM_TRY(h, cudaMemcpyAsync(h->h_scal, h->d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FM_TRY(h, cudaStreamSynchronize(h->stream));
  h->d2h += 32;
  float ms = 0.f;
  FM_TRY(h, cudaEventElapsedTime(&ms, h->ev_beg, h->ev_end));
  h->last_iter_ms = ms;
  if (loss_out) *loss_out = (h->h_scal[2] + h->h_scal[3]) * 0.05;  // loss *= 0.05 (:218)
  return CARS_OK;
}

// ---- row-sharded iteration: the same sweep with an all-reduce of the per-coordinate sums ------------------
extern "C" int cars_fm_exchange_doubles(const cars_fm_handle* h, int64_t* out) {
  if (!h || !out) return CARS_E_INVALID;
  int64_t m = h->U > h->I ? h->U : h->I;
  if (h->C > m) m = h->C;
  *out = 2 * m + 2;
  return CARS_OK;
}

extern "C" int cars_fm_iteration_sharded(cars_fm_handle* h, double* dev_buf, cars_allreduce_fn ar, void* user,
                                         double* loss_out) {
  if (!h) return CARS_E_INVALID;
  if (!dev_buf || !ar) return fm_fail(h, CARS_E_INVALID, "dev_buf and allreduce are required");
  if (!h->prepared) return fm_fail(h, CARS_E_STATE, "cars_fm_iteration before cars_fm_prepare");
  FM_TRY(h, cudaSetDevice(h->device));
  FM_TRY(h, cudaEventRecord(h->ev_beg, h->stream));
  int rc;
  fm_w0_reduce_kernel<<<h->red_blocks, 256, 0, h->stream>>>(h->d_e, h->d_w0, h->N, h->d_part);
  fm_w0_partial_kernel<<<1, 32, 0, h->stream>>>(h->d_part, h->red_blocks, dev_buf);
  FM_TRY(h, cudaGetLastError());
  // sum(e - w0) is a sum over the LOCAL rows of (e_n - w0): the global sum is the sum of the parts
  if (ar(user, dev_buf, 2) != 0) return fm_fail(h, CARS_E_STATE, "the all-reduce callback failed");
  fm_w0_finish_global_kernel<<<1, 32, 0, h->stream>>>(dev_buf, h->w0_denom, h->reg_lw, h->d_w0, h->d_scal);
  if (h->N) fm_w0_apply_kernel<<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(h->d_e, h->d_scal, h->N);
  fm_wreg_kernel<<<1, 256, 0, h->stream>>>(h->d_w, h->p, h->reg_lw, h->d_scal + 3);
  FM_TRY(h, cudaGetLastError());
  h->launches += 5;
  const double w_reg = (double)h->Nglobal * h->reg_lw;
  FmExchange ex;
  ex.dev_buf = dev_buf; ex.ar = ar; ex.user = user;
  if ((rc = field_sweep<0>(h, h->d_w, 1, 0, nullptr, w_reg, &ex))) return rc;
  const double v_reg = (double)h->Nglobal * h->reg_lf;
  for (int f = 0; f < h->k; f++)
    if ((rc = field_sweep<1>(h, h->d_V + (int64_t)f * h->p, 1, 0, h->d_Qc + (int64_t)f * h->Nq, v_reg, &ex))) return rc;
  FM_TRY(h, cudaEventRecord(h->ev_end, h->stream));
  FM_TRY(h, cudaMemcpyAsync(h->h_scal, h->d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FM_TRY(h, cudaStreamSynchronize(h->stream));
  h->d2h += 32;
  float ms = 0.f;
  FM_TRY(h, cudaEventElapsedTime(&ms, h->ev_beg, h->ev_end));
  h->last_iter_ms = ms;
  if (loss_out) *loss_out = (h->h_scal[2] + h->h_scal[3]) * 0.05;  // scal[2] already holds the GLOBAL sum e^2
  return CARS_OK;
}

extern "C" int cars_fm_predict(cars_fm_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx,
                               int32_t bound, double min_rate, double max_rate, double* out) {
  if (!h) return CARS_E_INVALID;
  if (!h->uploaded) return fm_fail(h, CARS_E_STATE, "cars_fm_predict before cars_fm_upload");
  if (n < 0 || (n > 0 && (!u || !j || !ctx || !out))) return fm_fail(h, CARS_E_INVALID, "bad predict arguments");
  if (n == 0) return CARS_OK;
  FM_TRY(h, cudaSetDevice(h->device));
  // queries through the staged copier (pageable arrays), range check on the device before the kernel indexes with them
  int32_t *du = nullptr, *dj = nullptr, *dc = nullptr;
  double* dout = nullptr;
  unsigned long long* d_bad = nullptr;
  unsigned long long bad = ~0ull;
  cudaError_t e = fm_alloc(&du, (size_t)n);
  if (e == cudaSuccess) e = fm_alloc(&dj, (size_t)n);
  if (e == cudaSuccess) e = fm_alloc(&dc, (size_t)n);
  if (e == cudaSuccess) e = fm_alloc(&dout, (size_t)n);
  if (e == cudaSuccess) e = fm_alloc(&d_bad, 1);
  if (e == cudaSuccess) {
    cars::CopySeg segs[3] = {{du, (void*)u, (size_t)n * 4}, {dj, (void*)j, (size_t)n * 4}, {dc, (void*)ctx, (size_t)n * 4}};
    e = h->copier.run(segs, 3, true);
  }
  if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0xff, 8, h->stream);
  if (e == cudaSuccess) {
    fms_validate_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(du, dj, dc, n, (uint32_t)h->U, (uint32_t)h->I, d_bad);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && bad == ~0ull) {
    fm_predict_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(du, dj, dc, h->d_w, h->d_V, h->d_w0, h->U, h->I, h->p,
                                                                         h->k, h->xc, n, bound, min_rate, max_rate, dout);
    e = cudaGetLastError();
    h->launches += 2;
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = h->copier.d2h(out, dout, (size_t)n * 8);
  }
  cudaFree(du); cudaFree(dj); cudaFree(dc); cudaFree(dout); cudaFree(d_bad);
  if (e != cudaSuccess) return fm_fail(h, CARS_E_CUDA, "predict failed: %s", cudaGetErrorString(e));
  if (bad != ~0ull) return fm_fail(h, CARS_E_INVALID, "query %lld has an id out of range", (long long)bad);
  h->h2d += n * 12; h->d2h += n * 8;
  return CARS_OK;
}

extern "C" int cars_fm_get_stats(const cars_fm_handle* h, cars_fm_stats* out) {
  if (!h || !out) return CARS_E_INVALID;
  out->nnz = h->N; out->p = h->p; out->kernel_launches = h->launches; out->h2d_bytes = h->h2d; out->d2h_bytes = h->d2h;
  out->last_iteration_ms = h->last_iter_ms;
  out->pieces = h->fld[0].f.num_pieces + h->fld[1].f.num_pieces + h->fld[2].f.num_pieces;
  return CARS_OK;
}

extern "C" void* cars_fm_get_stream(const cars_fm_handle* h) { return h ? (void*)h->stream : nullptr; }

extern "C" const char* cars_fm_last_error(const cars_fm_handle* h) { return h ? h->err.c_str() : g_fm_create_error.c_str(); }

extern "C" void cars_fm_destroy(cars_fm_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_u); cudaFree(h->d_j); cudaFree(h->d_c); cudaFree(h->d_r); cudaFree(h->d_e); cudaFree(h->d_Qc);
  cudaFree(h->d_w0); cudaFree(h->d_w); cudaFree(h->d_V); cudaFree(h->d_part); cudaFree(h->d_scal);
  for (auto& f : h->fld) {
    if (f.owns_coord) cudaFree(f.d_coord_of_row);
    cudaFree(f.d_perm); cudaFree(f.d_piece_coord); cudaFree(f.d_piece_beg); cudaFree(f.d_run_off);
    cudaFree(f.d_coord_piece); cudaFree(f.d_coord_rows); cudaFree(f.d_delta);
  }
  if (h->h_scal) cudaFreeHost(h->h_scal);
  h->copier.destroy();
  if (h->ev_beg) cudaEventDestroy(h->ev_beg);
  if (h->ev_end) cudaEventDestroy(h->ev_end);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}
