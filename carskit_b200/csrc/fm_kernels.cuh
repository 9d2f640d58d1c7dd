// fm_kernels.cuh -- K4: sm_100a kernels of the FM recommender's ALS sweep
// (reference: src/carskit/alg/cars/adaptation/dependent/FM.java:93-113 predict, :115-220 buildModel).
//
// The reference loops `for l < p` over ALL coordinates and `for i < size` over ALL rows through a dense
// feature table (O(k*p*size)).  Every row has exactly three non-zero features -- x_u = 1, x_{U+j} = 1 and
// x_{U+I+ctx} = 1/numContextDims (only when that index is < p, FM.java:81) -- so the coordinates of one
// FIELD (all users / all items / all contexts) touch disjoint rows: solving them concurrently is
// identical to the reference's sequential order.  Fields and factors stay sequential.
//
// One coordinate step = (a) reduce numerator/denominator over rows(l), (b) new value + delta per
// coordinate, (c) e_n += delta*x (and Qc[n][f] += delta*x) for the rows of that coordinate.
// (a) is a deterministic two-level reduction: rows sorted by coordinate are cut into pieces of <= 256
// rows, one warp sums a piece in a fixed order, one thread (or warp) sums a coordinate's pieces in order.
// Fields with few coordinates (contexts) are reduced by streaming the rows instead (fm_dense_reduce_kernel).
// Arithmetic: fp64, no FMA contraction in the update formulas (Java semantics); sums are tree-ordered, so
// results equal the sparse oracle's up to summation order (tolerance stated in tests/test_fm_gpu.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace carsfm {

struct FmField {
  const int32_t* coord_of_row;  // [N] coordinate (within the field) of every row, -1 = feature absent
  const int32_t* perm;          // [rows_in_field] row ids sorted by coordinate (stable)
  const int64_t* piece_beg;     // [num_pieces + 1] offsets into perm
  const int32_t* piece_coord;   // [num_pieces]
  const int64_t* coord_piece;   // [ncoord + 1] first piece of every coordinate
  const int64_t* coord_rows;    // [ncoord] number of rows
  int64_t num_pieces;
  int32_t ncoord;
  int32_t offset;  // index of the field's first coordinate in w / V
  double x;        // feature value of the field
  int32_t dense_blocks;  // > 0: the field is reduced by fm_dense_reduce_kernel on this many CTAs (no pieces)
  // run mode (fm_run_reduce_kernel): the rows are stored sorted by (row block, coordinate), so the rows of
  // `run_group` consecutive coordinates inside one row block are ONE contiguous run of the storage order
  const uint32_t* run_off;  // [(run_groups + 1) x run_blocks] first row of (group, block); nullptr = pieces
  int32_t run_group, run_groups, run_blocks;
};

constexpr int kDenseThreads = 128;
constexpr int kDenseMaxCoord = 64;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __shfl_down_sync(0xffffffffu, __double2loint(v), o);
    int hi = __shfl_down_sync(0xffffffffu, __double2hiint(v), o);
    v = __dadd_rn(v, __hiloint2double(hi, lo));
  }
  return v;
}

// cp.async (LDGSTS): global -> shared copies that occupy no register while in flight
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// V is kept FACTOR-MAJOR on the device (V[f * p + i]; the caller's array is [p x k] row-major and is transposed on
// upload / download): every coordinate step works on one factor of all coordinates of a field, a unit-stride
// slice this way (a 512-byte stride -- one DRAM sector per coordinate, read and written -- the other way).
// [rows x cols] -> [cols x rows], 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) fm_transpose_kernel(const double* __restrict__ in, int64_t rows, int cols,
                                                           double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[r * cols + c] : 0.0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t r = r0 + tx;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = tile[tx][i];
  }
}

// the way back: in = [cols x rows] (factor-major), out = [rows x cols]; same grid as fm_transpose_kernel
__global__ void __launch_bounds__(256) fm_untranspose_kernel(const double* __restrict__ in, int64_t rows, int cols,
                                                             double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t r = r0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[(int64_t)c * rows + r] : 0.0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + tx;
    if (r < rows && c < cols) out[r * cols + c] = tile[tx][i];
  }
}

// FM.predict in sparse form, ascending feature index (FM.java:93-113)
__device__ __forceinline__ double fm_predict_one(const double* __restrict__ w, const double* __restrict__ V, double w0,
                                                 int U, int I, int p, int k, double xc, int u, int j, int c) {
  const int iu = u, ij = U + j, ic = U + I + c;
  const bool has_c = ic < p;
  double pred = w0;
  pred = __dadd_rn(pred, __dmul_rn(w[iu], 1.0));
  pred = __dadd_rn(pred, __dmul_rn(w[ij], 1.0));
  if (has_c) pred = __dadd_rn(pred, __dmul_rn(w[ic], xc));
  double sum = 0.0;
  for (int f = 0; f < k; ++f) {
    double sum1 = 0.0, sum2 = 0.0;
    double d = V[(int64_t)f * p + iu];
    sum1 = __dadd_rn(sum1, d); sum2 = __dadd_rn(sum2, __dmul_rn(d, d));
    d = V[(int64_t)f * p + ij];
    sum1 = __dadd_rn(sum1, d); sum2 = __dadd_rn(sum2, __dmul_rn(d, d));
    if (has_c) {
      d = __dmul_rn(V[(int64_t)f * p + ic], xc);
      sum1 = __dadd_rn(sum1, d); sum2 = __dadd_rn(sum2, __dmul_rn(d, d));
    }
    sum = __dadd_rn(sum, __dsub_rn(__dmul_rn(sum1, sum1), sum2));
  }
  return __dadd_rn(pred, __dmul_rn(0.5, sum));
}

// Pre-pass (FM.java:118-146): e_n = r_n - predict, Qc[f][n] = sum_i V[i][f] x_n[i].  Qc is factor-major.
__global__ void __launch_bounds__(256) fm_prepare_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ j,
                                                         const int32_t* __restrict__ c, const double* __restrict__ r,
                                                         const double* __restrict__ w, const double* __restrict__ V,
                                                         const double* __restrict__ w0p, int U, int I, int p, int k,
                                                         double xc, int64_t N, int64_t Nq /*Qc row stride*/,
                                                         double* __restrict__ e, double* __restrict__ Qc) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int uu = u[n], jj = j[n], cc = c[n];
  e[n] = __dsub_rn(r[n], fm_predict_one(w, V, *w0p, U, I, p, k, xc, uu, jj, cc));
  const int ic = U + I + cc;
  for (int f = 0; f < k; ++f) {
    double v = 0.0;
    v = __dadd_rn(v, V[(int64_t)f * p + uu]);
    v = __dadd_rn(v, V[(int64_t)f * p + U + jj]);
    if (ic < p) v = __dadd_rn(v, __dmul_rn(V[(int64_t)f * p + ic], xc));
    Qc[(int64_t)f * Nq + n] = v;
  }
}

// The pre-pass for large inputs.  fm_prepare_kernel above walks V factor-major (V[f * p + i]): in the engine's row order a
// warp's 32 rows belong to 32 different users, so every 8-byte read of a user's coefficient pulls its own 32-byte sector
// -- 412 GB of DRAM reads for 125 M rows x 64 factors, 87 ms (profiles/r2/ncu_summary_fm_prepare_125M.txt).  Here V comes
// coordinate-major (Vt[i * k + f], a scratch transpose made by cars_fm_prepare): a row's three coefficient vectors are
// contiguous, a warp stages 16 factors of its 32 rows at a time in shared memory with coalesced 128-byte reads
// (16 lanes per row), and every lane then walks ITS row's factors in order -- the same operations in the same order as
// fm_predict_one / fm_prepare_kernel, so e and Qc are bit-identical to theirs.
constexpr int kPrepFC = 16;         // factors staged per step
constexpr int kPrepWarps = 4;       // per CTA
constexpr int kPrepTile = 32 * (kPrepFC + 1);  // doubles per staged array (row stride 17: conflict-free both ways)
__global__ void __launch_bounds__(kPrepWarps * 32) fm_prepare_tiled_kernel(
    const int32_t* __restrict__ u, const int32_t* __restrict__ j, const int32_t* __restrict__ c, const double* __restrict__ r,
    const double* __restrict__ w, const double* __restrict__ Vt, const double* __restrict__ w0p, int U, int I, int p, int k,
    double xc, int64_t N, int64_t Nq, double* __restrict__ e, double* __restrict__ Qc) {
  extern __shared__ __align__(16) unsigned char prep_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tu = reinterpret_cast<double*>(prep_smem) + (size_t)warp * 3 * kPrepTile;
  double* tj = tu + kPrepTile;
  double* tc = tj + kPrepTile;
  const double w0 = *w0p;
  const int srow = lane >> 4, sf = lane & 15;  // staging: 16 lanes per row, two rows per load instruction
  for (int64_t row0 = ((int64_t)blockIdx.x * kPrepWarps + warp) * 32; row0 < N; row0 += (int64_t)gridDim.x * kPrepWarps * 32) {
    const int64_t n = row0 + lane;
    const bool ok = n < N;
    const int iu = ok ? u[n] : 0, ij = U + (ok ? j[n] : 0), ic = U + I + (ok ? c[n] : 0);
    const bool has_c = ic < p;
    double pred = w0;
    pred = __dadd_rn(pred, __dmul_rn(w[iu], 1.0));
    pred = __dadd_rn(pred, __dmul_rn(w[ij], 1.0));
    if (has_c) pred = __dadd_rn(pred, __dmul_rn(w[ic], xc));
    double sum = 0.0;
    for (int f0 = 0; f0 < k; f0 += kPrepFC) {
      const int fc = k - f0 < kPrepFC ? k - f0 : kPrepFC;
#pragma unroll 4
      for (int rr = 0; rr < 32; rr += 2) {
        const int row = rr + srow;
        const int cu = __shfl_sync(0xffffffffu, iu, row), cj = __shfl_sync(0xffffffffu, ij, row);
        const int cc = __shfl_sync(0xffffffffu, has_c ? ic : -1, row);
        if (sf < fc) {  // cp.async: all 48 copies of the step are in flight together, no register holds them
          cp_async_8(tu + row * (kPrepFC + 1) + sf, Vt + (int64_t)cu * k + f0 + sf);
          cp_async_8(tj + row * (kPrepFC + 1) + sf, Vt + (int64_t)cj * k + f0 + sf);
          if (cc >= 0) cp_async_8(tc + row * (kPrepFC + 1) + sf, Vt + (int64_t)cc * k + f0 + sf);
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      for (int f = 0; f < fc; f++) {
        const double vu = tu[lane * (kPrepFC + 1) + f], vj = tj[lane * (kPrepFC + 1) + f];
        double sum1 = 0.0, sum2 = 0.0, v = 0.0;
        sum1 = __dadd_rn(sum1, vu); sum2 = __dadd_rn(sum2, __dmul_rn(vu, vu));
        sum1 = __dadd_rn(sum1, vj); sum2 = __dadd_rn(sum2, __dmul_rn(vj, vj));
        v = __dadd_rn(v, vu);
        v = __dadd_rn(v, vj);
        if (has_c) {
          const double d = __dmul_rn(tc[lane * (kPrepFC + 1) + f], xc);
          sum1 = __dadd_rn(sum1, d); sum2 = __dadd_rn(sum2, __dmul_rn(d, d));
          v = __dadd_rn(v, d);
        }
        sum = __dadd_rn(sum, __dsub_rn(__dmul_rn(sum1, sum1), sum2));
        if (ok) Qc[(int64_t)(f0 + f) * Nq + n] = v;
      }
      __syncwarp();
    }
    if (ok) e[n] = __dsub_rn(r[n], __dadd_rn(pred, __dmul_rn(0.5, sum)));
  }
}

__global__ void __launch_bounds__(256) fm_predict_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ j,
                                                         const int32_t* __restrict__ c, const double* __restrict__ w,
                                                         const double* __restrict__ V, const double* __restrict__ w0p,
                                                         int U, int I, int p, int k, double xc, int64_t n, int bound,
                                                         double lo, double hi, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pr = fm_predict_one(w, V, *w0p, U, I, p, k, xc, u[i], j[i], c[i]);
  if (bound) {
    if (pr > hi) pr = hi;
    if (pr < lo) pr = lo;
  }
  out[i] = pr;
}

// w0 step, pass 1 (FM.java:152-158): block partials of sum(e - w0) and sum(e^2), fixed slices.
__global__ void __launch_bounds__(256) fm_w0_reduce_kernel(const double* __restrict__ e, const double* __restrict__ w0p,
                                                           int64_t N, double* __restrict__ part /*[2 x grid]*/) {
  __shared__ double sa[8], sb[8];
  const double w0 = *w0p;
  const int64_t per = (N + gridDim.x - 1) / gridDim.x;
  const int64_t beg = (int64_t)blockIdx.x * per, end = beg + per < N ? beg + per : N;
  double a = 0.0, b = 0.0;
  for (int64_t i = beg + threadIdx.x; i < end; i += 256) {
    const double x = e[i];
    a = __dadd_rn(a, __dsub_rn(x, w0));
    b = __dadd_rn(b, __dmul_rn(x, x));
  }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < 8; w++) { ta += sa[w]; tb += sb[w]; }
    part[blockIdx.x] = ta;
    part[gridDim.x + blockIdx.x] = tb;
  }
}

// w0 step, pass 2 (:159-169): new w0, loss terms; scal = {w0, delta (new - old as two separate values), loss}
__global__ void fm_w0_finish_kernel(const double* __restrict__ part, int nblocks, double denom, double reg_lw,
                                    double* __restrict__ w0p, double* __restrict__ scal /*[0]=new [1]=old [2]=loss*/) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ta = 0.0, tb = 0.0;
  for (int i = 0; i < nblocks; i++) { ta += part[i]; tb += part[nblocks + i]; }
  const double w0 = *w0p;
  double up = ta / denom;
  up = 0.0 - up;
  scal[0] = up;
  scal[1] = w0;
  scal[2] = tb + (reg_lw * w0) * w0;
  *w0p = up;
}
__global__ void __launch_bounds__(256) fm_w0_apply_kernel(double* __restrict__ e, const double* __restrict__ scal, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) e[i] = __dsub_rn(__dadd_rn(e[i], scal[0]), scal[1]);  // errors + update_w0 - w0 (:165)
}

// (a) one group of LPP lanes per piece (LPP = 32: a warp; LPP = 8 when the pieces are short -- a user with a few
// dozen rows -- so that a warp works on four pieces at a time), PPG consecutive pieces per group with the loads of
// all of them in flight together (a short piece is a chain of four dependent loads: piece -> coordinate ->
// coefficient, piece -> rows -> e / Qc; two chains per lane hide each other).
//                         MODE 0: w step (:175-179)  num += (e - w_l x) x.
//                         MODE 1: V step (:199-204)  h = x Qc - x^2 V_lf; num += (e - V_lf h) h; den += h^2.
template <int MODE, int LPP, int PPG>
__global__ void __launch_bounds__(256) fm_piece_reduce_kernel(FmField fld, const double* __restrict__ e,
                                                              const double* __restrict__ Qf /*Qc[f]*/,
                                                              const double* __restrict__ coef /*w or V*/, int coef_stride,
                                                              int coef_col, double* __restrict__ part /*[2 x pieces]*/) {
  const int64_t piece0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPP) * PPG;
  const int gl = threadIdx.x % LPP;
  const double x = fld.x;
  bool valid[PPG];  // no early return: the group reductions below shuffle warp-wide
  int64_t beg[PPG], end[PPG];
  double cl[PPG], num[PPG], den[PPG];
#pragma unroll
  for (int k = 0; k < PPG; k++) {
    valid[k] = piece0 + k < fld.num_pieces;
    beg[k] = valid[k] ? fld.piece_beg[piece0 + k] : 0;
    end[k] = valid[k] ? fld.piece_beg[piece0 + k + 1] : 0;
    const int l = valid[k] ? fld.piece_coord[piece0 + k] : 0;
    cl[k] = valid[k] ? coef[(int64_t)(fld.offset + l) * coef_stride + coef_col] : 0.0;
    num[k] = 0.0;
    den[k] = 0.0;
  }
  auto add_row = [&](int k, double en, double qn) {  // one row's terms, in the row order of the piece
    if (MODE == 0) {
      num[k] = __dadd_rn(num[k], __dmul_rn(__dsub_rn(en, __dmul_rn(cl[k], x)), x));
    } else {
      const double h = __dsub_rn(__dmul_rn(x, qn), __dmul_rn(__dmul_rn(x, x), cl[k]));
      num[k] = __dadd_rn(num[k], __dmul_rn(__dsub_rn(en, __dmul_rn(cl[k], h)), h));
      den[k] = __dadd_rn(den[k], __dmul_rn(h, h));
    }
  };
  // first four rows of every lane in every piece: all gathers in flight together
  int32_t n0[PPG][4];
  double e0[PPG][4], q0[PPG][4];
#pragma unroll
  for (int k = 0; k < PPG; k++)
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int64_t i = beg[k] + gl + t * LPP;
      n0[k][t] = i < end[k] ? fld.perm[i] : -1;
    }
#pragma unroll
  for (int k = 0; k < PPG; k++)
#pragma unroll
    for (int t = 0; t < 4; t++) {
      e0[k][t] = n0[k][t] >= 0 ? e[n0[k][t]] : 0.0;
      q0[k][t] = (MODE == 1 && n0[k][t] >= 0) ? Qf[n0[k][t]] : 0.0;
    }
#pragma unroll
  for (int k = 0; k < PPG; k++) {
#pragma unroll
    for (int t = 0; t < 4; t++)
      if (n0[k][t] >= 0) add_row(k, e0[k][t], q0[k][t]);
    // the rest of a long piece, four gathers at a time (same order of additions as a plain loop)
    int64_t i = beg[k] + gl + 4 * LPP;
    for (; i + 3 * LPP < end[k]; i += 4 * LPP) {
      const int64_t a0 = fld.perm[i], a1 = fld.perm[i + LPP], a2 = fld.perm[i + 2 * LPP], a3 = fld.perm[i + 3 * LPP];
      const double x0 = e[a0], x1 = e[a1], x2 = e[a2], x3 = e[a3];
      double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
      if (MODE == 1) { y0 = Qf[a0]; y1 = Qf[a1]; y2 = Qf[a2]; y3 = Qf[a3]; }
      add_row(k, x0, y0); add_row(k, x1, y1); add_row(k, x2, y2); add_row(k, x3, y3);
    }
    for (; i < end[k]; i += LPP) {
      const int64_t n = fld.perm[i];
      add_row(k, e[n], MODE == 1 ? Qf[n] : 0.0);
    }
  }
#pragma unroll
  for (int k = 0; k < PPG; k++) {
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) {  // fixed-order tree over the group's lanes
      num[k] = __dadd_rn(num[k], __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(num[k]), o, LPP),
                                                  __shfl_down_sync(0xffffffffu, __double2loint(num[k]), o, LPP)));
      if (MODE == 1)
        den[k] = __dadd_rn(den[k], __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(den[k]), o, LPP),
                                                    __shfl_down_sync(0xffffffffu, __double2loint(den[k]), o, LPP)));
    }
    if (valid[k] && gl == 0) {
      part[piece0 + k] = num[k];
      part[fld.num_pieces + piece0 + k] = den[k];
    }
  }
}

// (a'') fields with MANY coordinates of FEW rows each whose rows are stored sorted by (row block, coordinate) -- the
// users: 25 rows per user and GPU at config 4, spread over 82 item blocks.  Gathering e[perm[i]] reads a 32-byte
// sector for every 8 bytes it uses (fm_piece_reduce_kernel<1, 8, 2>: 1.35 ms per launch at 125 M rows, L1TEX / L2
// bound with DRAM at 35 %, profiles/r2/launches_fm_125M.txt).  Here ONE CTA owns `run_group` consecutive coordinates:
// their rows inside a row block are one contiguous run of the storage order (~80 rows), which the CTA's warps read
// coalesced, lane = row.  Warp w takes the blocks b = w, w + 8, ..; the terms of consecutive lanes with the same
// coordinate are joined in row order by shuffles, and the head lane adds the sum to the WARP's accumulator of that
// coordinate in shared memory (one lane per coordinate and instruction: no conflict, program order = row order).
// The coordinate's total = its 8 warp accumulators in warp order.  Deterministic.
// FUSE >= 1: step (b) in the epilogue (new value, delta) -- one launch less per field and factor.
// FUSE == 2: step (c) too -- the CTA owns ALL rows of its coordinates, so it applies e_n += delta x (Qf likewise) to the
// runs it has just read (L2 hits: a CTA's rows are ~100 KB) with fm_row_update_kernel's arithmetic; the separate row
// update's 2.8 GB DRAM read per launch goes away.
// Latency: the rows of the next kRunSlots - 1 chunks are always in flight -- cp.async into a per-warp ring in shared
// memory (a lane reads back only what it copied itself: no barrier on the ring) -- and pass 2 of FUSE == 2 loads four
// chunks before it stores any.
constexpr int kRunSlots = 4;
constexpr int kRunSlotBytes = 32 * 4 + 32 * 8 + 32 * 8;  // coordinate, e, Qf of 32 rows
__host__ __device__ constexpr size_t fm_run_smem_bytes(int G) {
  return (size_t)G * sizeof(double) + (size_t)8 * G * sizeof(double2) + (size_t)8 * kRunSlots * kRunSlotBytes;
}
// iterator over the 32-row chunks of a warp's runs of one batch (lane t holds the bounds of the batch's t-th run)
struct RunChunks {
  uint32_t rs, re, pos, end;
  int t, nruns;
  bool more;
  __device__ __forceinline__ void start(uint32_t rs_, uint32_t re_, int nruns_) {
    rs = rs_; re = re_; nruns = nruns_; t = 0;
    pos = __shfl_sync(0xffffffffu, rs, 0);
    end = __shfl_sync(0xffffffffu, re, 0);
    more = settle();
  }
  __device__ __forceinline__ bool settle() {  // skip exhausted / empty runs; false = batch done (warp-uniform)
    while (pos >= end) {
      if (++t >= nruns) return false;
      pos = __shfl_sync(0xffffffffu, rs, t);
      end = __shfl_sync(0xffffffffu, re, t);
    }
    return true;
  }
  __device__ __forceinline__ void next() { pos += 32; more = settle(); }
};

template <int MODE, int FUSE>
__global__ void __launch_bounds__(256) fm_run_reduce_kernel(FmField fld, const double* __restrict__ e,
                                                            const double* __restrict__ Qf, double* coef, int coef_stride,
                                                            int coef_col, double size_reg, double* __restrict__ part,
                                                            double* __restrict__ delta, double* e_rw, double* Qf_rw) {
  extern __shared__ __align__(16) unsigned char run_smem[];
  const int G = fld.run_group;
  double* coef_s = reinterpret_cast<double*>(run_smem);                              // [G]
  double2* acc = reinterpret_cast<double2*>(run_smem + (size_t)G * sizeof(double));  // [8][G]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* ring = run_smem + (size_t)G * sizeof(double) + (size_t)8 * G * sizeof(double2) +
                        (size_t)warp * kRunSlots * kRunSlotBytes;  // [kRunSlots][coordinate 32 x 4 | e 32 x 8 | Qf 32 x 8]
  const int g = blockIdx.x;
  const int l0 = g * G;
  const int nl = fld.ncoord - l0 < G ? fld.ncoord - l0 : G;
  const double x = fld.x;
  for (int t = tid; t < G; t += 256) coef_s[t] = t < nl ? coef[(int64_t)(fld.offset + l0 + t) * coef_stride + coef_col] : 0.0;
  for (int t = tid; t < 8 * G; t += 256) acc[t] = make_double2(0.0, 0.0);
  __syncthreads();
  double2* wacc = acc + (size_t)warp * G;
  // FUSE == 2 rewrites the rows it read: no read-only (non-coherent) loads through the __restrict__ pointers then
  const double* ep = FUSE == 2 ? e_rw : e;
  const double* qp = FUSE == 2 ? Qf_rw : Qf;
  const int nb = fld.run_blocks;
  const uint32_t* off0 = fld.run_off + (int64_t)g * nb;
  const uint32_t* off1 = off0 + nb;

  // the warp's runs come in batches of 32 (lane t of a batch holds the bounds of block b = base + warp + 8 t)
  for (int base = 0; base + warp < nb; base += 256) {
    const int myb = base + warp + 8 * lane;
    const uint32_t rs = myb < nb ? off0[myb] : 0u, re = myb < nb ? off1[myb] : 0u;
    int nruns = (nb - base - warp + 7) / 8;
    if (nruns > 32) nruns = 32;
    RunChunks it;
    it.start(rs, re, nruns);
    int issued = 0;
    auto issue = [&](int slot) {  // request the iterator's chunk into ring slot `slot` (or nothing), one group either way
      if (it.more) {
        unsigned char* sl = ring + slot * kRunSlotBytes;
        const uint32_t n = it.pos + lane;
        if (n < it.end) {
          cp_async_4(sl + lane * 4, fld.coord_of_row + n);
          cp_async_8(sl + 128 + lane * 8, ep + n);
          if (MODE == 1) cp_async_8(sl + 384 + lane * 8, qp + n);
        } else {
          reinterpret_cast<int*>(sl)[lane] = -1;  // past the run's end
        }
        it.next();
        issued++;
      }
      cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < kRunSlots - 1; k++) issue(k);
    for (int c = 0; c < issued; c++) {
      issue((c + kRunSlots - 1) % kRunSlots);
      cp_async_wait<kRunSlots - 1>();  // chunk c has landed (every lane waits for its own copies)
      const unsigned char* sl = ring + (c % kRunSlots) * kRunSlotBytes;
      const int cu = reinterpret_cast<const int*>(sl)[lane];
      const bool ok = cu >= 0;
      const double ce = ok ? reinterpret_cast<const double*>(sl + 128)[lane] : 0.0;
      const double cq = (MODE == 1 && ok) ? reinterpret_cast<const double*>(sl + 384)[lane] : 0.0;
      const int ul = ok ? cu - l0 : 0;
      const double cl = ok ? coef_s[ul] : 0.0;
      double num, den = 0.0;
      if (MODE == 0) {
        num = __dmul_rn(__dsub_rn(ce, __dmul_rn(cl, x)), x);
      } else {
        const double hh = __dsub_rn(__dmul_rn(x, cq), __dmul_rn(__dmul_rn(x, x), cl));
        num = __dmul_rn(__dsub_rn(ce, __dmul_rn(cl, hh)), hh);
        den = __dmul_rn(hh, hh);
      }
      const int key = ok ? ul : -1 - lane;  // rows past the run's end: singletons nobody adds
      const int pk = __shfl_up_sync(0xffffffffu, key, 1);
      const bool head = lane == 0 || pk != key;
      const unsigned hm = __ballot_sync(0xffffffffu, head);
      const int hpos = 31 - __clz(hm & (0xffffffffu >> (31 - lane)));
      const int dist = lane - hpos;
      for (int d = 1; __ballot_sync(0xffffffffu, dist >= d) != 0u; d++) {  // usually one round: 14 % of a coordinate's runs have a second row
        const double tn = __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(num), d),
                                           __shfl_down_sync(0xffffffffu, __double2loint(num), d));
        const int dd = __shfl_down_sync(0xffffffffu, dist, d);
        const bool take = lane + d < 32 && dd == d;  // only a head has a lane at distance d from it
        if (take) num = __dadd_rn(num, tn);
        if (MODE == 1) {
          const double td = __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(den), d),
                                             __shfl_down_sync(0xffffffffu, __double2loint(den), d));
          if (take) den = __dadd_rn(den, td);
        }
      }
      if (head && ok) {
        double2 a = wacc[ul];
        a.x = __dadd_rn(a.x, num);
        if (MODE == 1) a.y = __dadd_rn(a.y, den);
        wacc[ul] = a;
      }
      __syncwarp();
    }
    cp_async_wait<0>();
  }
  __syncthreads();
  for (int t = tid; t < nl; t += 256) {
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const double2 a = acc[(size_t)w * G + t];
      num = __dadd_rn(num, a.x);
      if (MODE == 1) den = __dadd_rn(den, a.y);
    }
    const int l = l0 + t;
    if (FUSE >= 1) {
      if (MODE == 0) den = __dmul_rn((double)fld.coord_rows[l], __dmul_rn(x, x));
      den = __dadd_rn(den, size_reg);
      const double old = coef_s[t];
      const double nv = __dsub_rn(0.0, num / den);
      coef[(int64_t)(fld.offset + l) * coef_stride + coef_col] = nv;
      const double dl = __dsub_rn(nv, old);
      delta[l] = dl;
      if (FUSE == 2) coef_s[t] = __dmul_rn(dl, x);  // only thread t reads or writes slot t here
    } else {
      part[l] = num;
      if (MODE == 1) part[(int64_t)fld.ncoord + l] = den;
    }
  }
  if (FUSE == 2) {
    __syncthreads();
    for (int base = 0; base + warp < nb; base += 256) {
      const int myb = base + warp + 8 * lane;
      const uint32_t rs = myb < nb ? off0[myb] : 0u, re = myb < nb ? off1[myb] : 0u;
      int nruns = (nb - base - warp + 7) / 8;
      if (nruns > 32) nruns = 32;
      RunChunks it;
      it.start(rs, re, nruns);
      while (it.more) {  // four chunks' loads (L2 hits mostly) before the first store
        uint32_t n[4];
        int uu[4];
        double ee[4], qq[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          n[k] = 0xffffffffu;
          if (it.more) {
            if (it.pos + lane < it.end) n[k] = it.pos + lane;
            it.next();
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const bool ok = n[k] != 0xffffffffu;
          uu[k] = ok ? fld.coord_of_row[n[k]] : l0;
          ee[k] = ok ? e_rw[n[k]] : 0.0;
          qq[k] = (MODE == 1 && ok) ? Qf_rw[n[k]] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (n[k] == 0xffffffffu) continue;
          const double d = coef_s[uu[k] - l0];
          e_rw[n[k]] = __dadd_rn(ee[k], d);
          if (MODE == 1) Qf_rw[n[k]] = __dadd_rn(qq[k], d);
        }
      }
    }
  }
}

// (a') fields with FEW coordinates (contexts: <= kDenseMaxCoord): the rows are streamed in storage order, fully
// coalesced, and every thread adds a row's terms into ITS OWN accumulator of the row's coordinate in shared
// memory (bins[coordinate][thread]: no conflicts, fixed order).  A CTA owns a fixed set of row chunks; its per-
// coordinate sums (fixed-order tree over the threads) go to part[(block * ncoord + c) * 2 + {0, 1}].  Deterministic.
template <int MODE>
__global__ void __launch_bounds__(kDenseThreads) fm_dense_reduce_kernel(FmField fld, const double* __restrict__ e,
                                                                        const double* __restrict__ Qf,
                                                                        const double* __restrict__ coef, int coef_stride,
                                                                        int coef_col, int64_t N, double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char dense_smem[];
  double2* bins = reinterpret_cast<double2*>(dense_smem);  // [ncoord][kDenseThreads]
  __shared__ double cls[kDenseMaxCoord];
  constexpr int T = kDenseThreads;
  constexpr int UNR = 8;  // rows in flight per thread
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nc = fld.ncoord;
  const double x = fld.x;
  for (int c = tid; c < nc; c += T) cls[c] = coef[(int64_t)(fld.offset + c) * coef_stride + coef_col];
  for (int i = tid; i < nc * T; i += T) bins[i] = make_double2(0.0, 0.0);
  __syncthreads();
  auto add_row = [&](int c, double en, double qn) {
    if (c < 0) return;
    const double cl = cls[c];
    double2 b = bins[c * T + tid];
    if (MODE == 0) {
      b.x = __dadd_rn(b.x, __dmul_rn(__dsub_rn(en, __dmul_rn(cl, x)), x));
    } else {
      const double h = __dsub_rn(__dmul_rn(x, qn), __dmul_rn(__dmul_rn(x, x), cl));
      b.x = __dadd_rn(b.x, __dmul_rn(__dsub_rn(en, __dmul_rn(cl, h)), h));
      b.y = __dadd_rn(b.y, __dmul_rn(h, h));
    }
    bins[c * T + tid] = b;
  };
  // The rows are dealt to the CTAs in chunks of UNR * T rows, round-robin (chunk q -> CTA q mod grid): at any time the
  // grid reads ONE window of grid * 20 KB of every array.  (Giving each CTA one contiguous slice -- 444 streams 2.25 MB
  // apart per array -- ran at 0.49 of the DRAM peak whatever the number of rows in flight:
  // profiles/r2/ncu_summary_fm_dense_reduce_125M.txt.)  A thread's rows and their order depend on the grid size only.
  // Two register buffers: the next chunk is requested before this one's shared-memory updates.
  constexpr int64_t CH = (int64_t)UNR * T;
  const int64_t nchunks = N / CH, G = gridDim.x;
  int ca[UNR], cb[UNR];
  double ea[UNR], qa[UNR], eb[UNR], qb[UNR];
  auto fetch = [&](int64_t q, int* c, double* en, double* qn) {
    const int64_t at = q * CH + tid;
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      c[k] = fld.coord_of_row[at + k * T];
      en[k] = e[at + k * T];
      qn[k] = MODE == 1 ? Qf[at + k * T] : 0.0;
    }
  };
  int64_t q = blockIdx.x;
  if (q < nchunks) {
    fetch(q, ca, ea, qa);
    for (;;) {
      const bool more_b = q + G < nchunks;
      if (more_b) fetch(q + G, cb, eb, qb);
#pragma unroll
      for (int k = 0; k < UNR; k++) add_row(ca[k], ea[k], qa[k]);
      q += G;
      if (!more_b) break;
      const bool more_a = q + G < nchunks;
      if (more_a) fetch(q + G, ca, ea, qa);
#pragma unroll
      for (int k = 0; k < UNR; k++) add_row(cb[k], eb[k], qb[k]);
      q += G;
      if (!more_a) break;
    }
  }
  if (blockIdx.x == 0)  // the rows after the last full chunk
    for (int64_t n = nchunks * CH + tid; n < N; n += T) add_row(fld.coord_of_row[n], e[n], MODE == 1 ? Qf[n] : 0.0);
  __syncthreads();
  for (int c = warp; c < nc; c += T / 32) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < T / 32; k++) {
      const double2 v = bins[c * T + lane + 32 * k];
      a = __dadd_rn(a, v.x);
      b = __dadd_rn(b, v.y);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      part[((int64_t)blockIdx.x * nc + c) * 2] = a;
      part[((int64_t)blockIdx.x * nc + c) * 2 + 1] = b;
    }
  }
}

// Sum of a coordinate's piece partials.  WIDE = false: one thread walks the pieces in order (fields with many
// coordinates of a few pieces each: users, items).  WIDE = true: one warp per coordinate, lanes stride the pieces
// and a fixed-order tree joins them (fields with few coordinates of many pieces: contexts) -- both deterministic.
template <int MODE, bool WIDE>
__device__ __forceinline__ void coord_piece_sums(const FmField& fld, const double* __restrict__ part, int l, int lane,
                                                 double& num, double& den) {
  num = 0.0;
  den = 0.0;
  if (WIDE && fld.dense_blocks > 0) {  // partials of fm_dense_reduce_kernel: one pair per CTA and coordinate
    for (int b = lane; b < fld.dense_blocks; b += 32) {
      num = __dadd_rn(num, part[((int64_t)b * fld.ncoord + l) * 2]);
      if (MODE == 1) den = __dadd_rn(den, part[((int64_t)b * fld.ncoord + l) * 2 + 1]);
    }
    num = warp_sum(num);
    if (MODE == 1) den = warp_sum(den);
    if (MODE == 0) den = __dmul_rn((double)fld.coord_rows[l], __dmul_rn(fld.x, fld.x));
    return;
  }
  if (fld.run_off) {  // fm_run_reduce_kernel left ONE pair per coordinate
    num = part[l];
    den = MODE == 1 ? part[(int64_t)fld.ncoord + l] : __dmul_rn((double)fld.coord_rows[l], __dmul_rn(fld.x, fld.x));
    return;
  }
  const int64_t q0 = fld.coord_piece[l], q1 = fld.coord_piece[l + 1];
  if (WIDE) {
    for (int64_t q = q0 + lane; q < q1; q += 32) {
      num = __dadd_rn(num, part[q]);
      if (MODE == 1) den = __dadd_rn(den, part[fld.num_pieces + q]);
    }
    num = warp_sum(num);
    if (MODE == 1) den = warp_sum(den);
  } else {
    for (int64_t q = q0; q < q1; q++) {
      num = __dadd_rn(num, part[q]);
      if (MODE == 1) den = __dadd_rn(den, part[fld.num_pieces + q]);
    }
  }
  if (MODE == 0) den = __dmul_rn((double)fld.coord_rows[l], __dmul_rn(fld.x, fld.x));
}

// (b) per coordinate: sum its pieces, new = 0 - num/den, delta = new - old.
template <int MODE, bool WIDE>
__global__ void __launch_bounds__(256) fm_coord_kernel(FmField fld, const double* __restrict__ part, double size_reg,
                                                       double* __restrict__ coef, int coef_stride, int coef_col,
                                                       double* __restrict__ delta /*[ncoord]*/) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = WIDE ? (t >> 5) : t;
  const int lane = threadIdx.x & 31;
  if (l >= fld.ncoord) return;
  double num, den;
  coord_piece_sums<MODE, WIDE>(fld, part, l, lane, num, den);
  if (WIDE && lane != 0) return;
  den = __dadd_rn(den, size_reg);
  double* cp = coef + (int64_t)(fld.offset + l) * coef_stride + coef_col;
  const double old = *cp;
  const double nv = __dsub_rn(0.0, num / den);
  *cp = nv;
  delta[l] = __dsub_rn(nv, old);
}

// Sharded variant of (b), split around the caller's all-reduce (rows of a coordinate live on several GPUs):
// (b1) per-coordinate LOCAL sums -> buf[l] = num, buf[ncoord + l] = den part (MODE 0: local row count * x^2)
template <int MODE, bool WIDE>
__global__ void __launch_bounds__(256) fm_coord_partial_kernel(FmField fld, const double* __restrict__ part,
                                                               double* __restrict__ buf /*[2 x ncoord]*/) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = WIDE ? (t >> 5) : t;
  const int lane = threadIdx.x & 31;
  if (l >= fld.ncoord) return;
  double num, den;
  coord_piece_sums<MODE, WIDE>(fld, part, l, lane, num, den);
  if (WIDE && lane != 0) return;
  buf[l] = num;
  buf[fld.ncoord + l] = den;
}
// (b2) after the all-reduce: every rank computes the same new value from the same global sums
__global__ void __launch_bounds__(256) fm_coord_finish_kernel(FmField fld, const double* __restrict__ buf, double size_reg,
                                                              double* __restrict__ coef, int coef_stride, int coef_col,
                                                              double* __restrict__ delta) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= fld.ncoord) return;
  const double den = __dadd_rn(buf[fld.ncoord + l], size_reg);
  double* cp = coef + (int64_t)(fld.offset + l) * coef_stride + coef_col;
  const double old = *cp;
  const double nv = __dsub_rn(0.0, buf[l] / den);
  *cp = nv;
  delta[l] = __dsub_rn(nv, old);
}
// w0 step, sharded: the two local sums -> buf[0..1]; finish from the global sums
__global__ void fm_w0_partial_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ buf) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ta = 0.0, tb = 0.0;
  for (int i = 0; i < nblocks; i++) { ta += part[i]; tb += part[nblocks + i]; }
  buf[0] = ta;
  buf[1] = tb;
}
__global__ void fm_w0_finish_global_kernel(const double* __restrict__ buf, double denom, double reg_lw,
                                           double* __restrict__ w0p, double* __restrict__ scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double w0 = *w0p;
  double up = buf[0] / denom;
  up = 0.0 - up;
  scal[0] = up;
  scal[1] = w0;
  scal[2] = buf[1] + (reg_lw * w0) * w0;
  *w0p = up;
}

// (c) row-parallel: e_n += (new - old) x ; V step also Qc[n][f] += (new - old) x  (:184-185, :208-211).
// Two consecutive rows per thread with 16-byte accesses (e and every Qc[f] start 16-byte aligned: the Qc row
// stride is padded to even).
template <int MODE>
__global__ void __launch_bounds__(256) fm_row_update_kernel(FmField fld, const double* __restrict__ delta, int64_t N,
                                                            double* __restrict__ e, double* __restrict__ Qf) {
  const int64_t n = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (n >= N) return;
  if (n + 1 < N) {
    const int2 l = *reinterpret_cast<const int2*>(fld.coord_of_row + n);
    if (l.x < 0 && l.y < 0) return;
    const double d0 = l.x >= 0 ? __dmul_rn(delta[l.x], fld.x) : 0.0;
    const double d1 = l.y >= 0 ? __dmul_rn(delta[l.y], fld.x) : 0.0;
    double2 ev = *reinterpret_cast<double2*>(e + n);
    if (l.x >= 0) ev.x = __dadd_rn(ev.x, d0);
    if (l.y >= 0) ev.y = __dadd_rn(ev.y, d1);
    *reinterpret_cast<double2*>(e + n) = ev;
    if (MODE == 1) {
      double2 qv = *reinterpret_cast<double2*>(Qf + n);
      if (l.x >= 0) qv.x = __dadd_rn(qv.x, d0);
      if (l.y >= 0) qv.y = __dadd_rn(qv.y, d1);
      *reinterpret_cast<double2*>(Qf + n) = qv;
    }
  } else {
    const int l = fld.coord_of_row[n];
    if (l < 0) return;
    const double d = __dmul_rn(delta[l], fld.x);
    e[n] = __dadd_rn(e[n], d);
    if (MODE == 1) Qf[n] = __dadd_rn(Qf[n], d);
  }
}

// sum_l (regLw * w_l) * w_l over all p coordinates, deterministic (one block)
__global__ void __launch_bounds__(256) fm_wreg_kernel(const double* __restrict__ w, int p, double reg_lw,
                                                      double* __restrict__ out) {
  __shared__ double sh[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < p; i += 256) t = __dadd_rn(t, __dmul_rn(__dmul_rn(reg_lw, w[i]), w[i]));
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

}  // namespace carsfm
