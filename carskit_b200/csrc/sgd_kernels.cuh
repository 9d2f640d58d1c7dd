// sgd_kernels.cuh -- sm_100a device code of the CARSKit SGD hot path.
//
// One "rating update" is the body of the per-rating loop of buildModel()
// (reference: src/carskit/alg/cars/adaptation/dependent/dev/CAMF_CI.java:80-121 and the sibling
// classes CAMF_C.java:80-128, CAMF_CU.java:77-118, baseline/cf/BiasedMF.java:63-99, PMF.java:52-71).
//
// Arithmetic contract: fp64, every * + - separately rounded (Java has no FMA
// contraction) -> all arithmetic below goes through __dmul_rn/__dadd_rn/__dsub_rn, which nvcc never
// fuses; the dot product P[u].Q[j] is summed in f = 0..F-1 order like librec's DenseMatrix.rowMult.
//
// Work decomposition: a *group* of LPR lanes (LPR in {8,16,32}) owns one rating at a time; lane l of
// the group holds the 16-byte chunks l, l+LPR, ... of the two factor rows, so every warp-wide
// ld.global.cg.v2.f64 covers whole 128-byte lines of a row (512 B per row at F = 64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cars {

enum : int { M_PMF = 0, M_BIASEDMF = 1, M_CAMF_C = 2, M_CAMF_CI = 3, M_CAMF_CU = 4 };

// Device-resident state shared by all kernels.  Plain pointers into the handle's allocations.
struct DeviceModel {
  double* P;          // [num_users x Fp]   row stride Fp = F rounded up to even (16-byte chunks)
  double* Q;          // [num_items x Fp]
  double* user_bias;  // [num_users]
  double* item_bias;  // [num_items]
  double* cond_bias;  // [C]
  double* ic_bias;    // [num_items x C]
  double* uc_bias;    // [num_users x C]
  const int32_t* ctx_tab;  // [num_contexts x Dmax] condition ids, -1 padded
  int32_t F, Fp, C, Dmax;
  double global_mean;
  double reg_u, reg_i, reg_b, reg_c;
};

// Ratings in schedule order (sorted by wavefront level; see schedule.cuh).
struct RatingStream {
  const int32_t* u;
  const int32_t* j;
  const int32_t* ctx;  // nullptr for the 2-D models
  const double* r;
  const int64_t* level_start;  // [num_levels + 1]
  int32_t num_levels;
};

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld_cg_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_cg_f64x2(double* p, double2 v) {
  asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_cg_f64(double* p, double v) {
  asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double shfl_f64(unsigned mask, double v, int src, int width) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(mask, lo, src, width);
  hi = __shfl_sync(mask, hi, src, width);
  return __hiloint2double(hi, lo);
}

// Butterfly-free tree sum over a warp (lane 0 holds the total); fixed order -> deterministic.
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __shfl_down_sync(0xffffffffu, __double2loint(v), o);
    int hi = __shfl_down_sync(0xffffffffu, __double2hiint(v), o);
    v += __hiloint2double(hi, lo);
  }
  return v;
}

// Grid-wide barrier for a co-resident (cooperative) persistent grid.  `target` is the value the
// monotonically increasing counter reaches once every CTA has arrived.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add_u32(counter, 1u);  // release: orders this CTA's prior stores (cumulative via bar.sync)
    while (ld_acquire_u32(counter) < target) {
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// One rating update by one group of LPR lanes.  V = 16-byte chunks per lane (LPR * V * 2 >= Fp).
// `prod` = this group's shared-memory scratch of Fp doubles (for the in-order dot product).
// Returns this lane's contribution to the epoch loss (un-halved).
// ------------------------------------------------------------------------------------------------
template <int MODEL, int LPR, int V>
__device__ __forceinline__ double rating_update(const DeviceModel& m, int u, int j, int ctx, double r,
                                                double lr, double* prod, int gl /*lane in group*/,
                                                unsigned gmask /*lanes of this group*/) {
  const int Fp = m.Fp;
  double* prow = m.P + (int64_t)u * Fp;
  double* qrow = m.Q + (int64_t)j * Fp;

  // ---- gather -------------------------------------------------------------------------------
  double2 p[V], q[V];
#pragma unroll
  for (int v = 0; v < V; v++) {
    const int c = gl + v * LPR;  // chunk index
    if (2 * c < Fp) {
      p[v] = ld_cg_f64x2(prow + 2 * c);
      q[v] = ld_cg_f64x2(qrow + 2 * c);
    } else {
      p[v] = make_double2(0.0, 0.0);
      q[v] = make_double2(0.0, 0.0);
    }
  }
  double bu = 0.0, bj = 0.0;
  if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CI) bu = ld_cg_f64(m.user_bias + u);
  if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CU) bj = ld_cg_f64(m.item_bias + j);

  // context-condition biases: lane d of the group owns condition d (d < D <= Dmax)
  constexpr bool kHasCond = (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU);
  double* cb_ptr = nullptr;  // this lane's condition-bias cell (first pass only when Dmax <= LPR)
  double cb = 0.0;
  const int Dmax = m.Dmax;
  if (kHasCond && gl < Dmax) {
    const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + gl);
    if (cond >= 0) {
      if (MODEL == M_CAMF_C) cb_ptr = m.cond_bias + cond;
      if (MODEL == M_CAMF_CI) cb_ptr = m.ic_bias + (int64_t)j * m.C + cond;
      if (MODEL == M_CAMF_CU) cb_ptr = m.uc_bias + (int64_t)u * m.C + cond;
      cb = ld_cg_f64(cb_ptr);
    }
  }

  // ---- dot product in f order (DenseMatrix.rowMult) --------------------------------------------
#pragma unroll
  for (int v = 0; v < V; v++) {
    const int c = gl + v * LPR;
    if (2 * c < Fp) {
      double2 t = make_double2(__dmul_rn(p[v].x, q[v].x), __dmul_rn(p[v].y, q[v].y));
      *reinterpret_cast<double2*>(prod + 2 * c) = t;
    }
  }
  __syncwarp(gmask);
  double dot = 0.0;
  {
    const int F = m.F;
    const int F2 = F & ~1;
    for (int f = 0; f < F2; f += 2) {
      double2 t = *reinterpret_cast<const double2*>(prod + f);
      dot = __dadd_rn(dot, t.x);
      dot = __dadd_rn(dot, t.y);
    }
    if (F & 1) dot = __dadd_rn(dot, prod[F - 1]);
  }
  __syncwarp(gmask);  // scratch is reused by the next rating

  // ---- predict --------------------------------------------------------------------------------
  double pred;
  if (MODEL == M_PMF) pred = dot;
  if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C)
    pred = __dadd_rn(__dadd_rn(__dadd_rn(m.global_mean, bu), bj), dot);
  if (MODEL == M_CAMF_CI) pred = __dadd_rn(__dadd_rn(m.global_mean, bu), dot);
  if (MODEL == M_CAMF_CU) pred = __dadd_rn(__dadd_rn(m.global_mean, bj), dot);

  double lane_loss = 0.0;
  if (kHasCond) {
    // conditions beyond the first LPR (Dmax > LPR) are handled by the slow path below
    const int D1 = Dmax < LPR ? Dmax : LPR;
    for (int d = 0; d < D1; d++) {
      const double b = shfl_f64(gmask, cb, d, LPR);
      pred = __dadd_rn(pred, b);  // adding a padded slot adds +0.0: exact
    }
    for (int d = LPR; d < Dmax; d++) {  // rare: more context dimensions than lanes in a group
      const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
      if (cond >= 0) {
        const double* bp = MODEL == M_CAMF_C    ? m.cond_bias + cond
                           : MODEL == M_CAMF_CI ? m.ic_bias + (int64_t)j * m.C + cond
                                                : m.uc_bias + (int64_t)u * m.C + cond;
        pred = __dadd_rn(pred, ld_cg_f64(bp));
      }
    }
  }
  const double e = __dsub_rn(r, pred);

  // ---- bias steps -----------------------------------------------------------------------------
  if (gl == 0) {
    lane_loss = __dmul_rn(e, e);
    if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CI) {
      const double sgd = __dsub_rn(e, __dmul_rn(m.reg_b, bu));
      st_cg_f64(m.user_bias + u, __dadd_rn(bu, __dmul_rn(lr, sgd)));
      lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bu), bu));
    }
    if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CU) {
      const double sgd = __dsub_rn(e, __dmul_rn(m.reg_b, bj));
      st_cg_f64(m.item_bias + j, __dadd_rn(bj, __dmul_rn(lr, sgd)));
      lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bj), bj));
    }
  }
  if (kHasCond) {
    if (cb_ptr != nullptr) {
      const double sgd = __dsub_rn(e, __dmul_rn(m.reg_c, cb));
      const double step = __dmul_rn(lr, sgd);
      st_cg_f64(cb_ptr, __dadd_rn(cb, step));
      // CAMF_C.java:115 adds regB * sum(bc) (not squared); CI/CU add regC * sum(b^2) (:108 / :105)
      if (MODEL == M_CAMF_C)
        lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_b, cb));
      else
        lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_c, __dmul_rn(cb, cb)));
    }
    if (gl == 0) {
      for (int d = LPR; d < Dmax; d++) {
        const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
        if (cond < 0) continue;
        double* bp = MODEL == M_CAMF_C    ? m.cond_bias + cond
                     : MODEL == M_CAMF_CI ? m.ic_bias + (int64_t)j * m.C + cond
                                          : m.uc_bias + (int64_t)u * m.C + cond;
        const double b = ld_cg_f64(bp);
        const double step = __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_c, b)));
        st_cg_f64(bp, __dadd_rn(b, step));
        lane_loss = __dadd_rn(lane_loss, MODEL == M_CAMF_C ? __dmul_rn(m.reg_b, b)
                                                             : __dmul_rn(m.reg_c, __dmul_rn(b, b)));
      }
    }
  }

  // ---- factor steps (both from the OLD p, q) and scatter -------------------------------------------
#pragma unroll
  for (int v = 0; v < V; v++) {
    const int c = gl + v * LPR;
    if (2 * c < Fp) {
      const double2 po = p[v], qo = q[v];
      double2 pn, qn;
      pn.x = __dadd_rn(po.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.x), __dmul_rn(m.reg_u, po.x))));
      qn.x = __dadd_rn(qo.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, po.x), __dmul_rn(m.reg_i, qo.x))));
      pn.y = __dadd_rn(po.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.y), __dmul_rn(m.reg_u, po.y))));
      qn.y = __dadd_rn(qo.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, po.y), __dmul_rn(m.reg_i, qo.y))));
      st_cg_f64x2(prow + 2 * c, pn);
      st_cg_f64x2(qrow + 2 * c, qn);
      lane_loss = __dadd_rn(lane_loss, __dadd_rn(__dmul_rn(__dmul_rn(m.reg_u, po.x), po.x),
                                                 __dmul_rn(__dmul_rn(m.reg_i, qo.x), qo.x)));
      lane_loss = __dadd_rn(lane_loss, __dadd_rn(__dmul_rn(__dmul_rn(m.reg_u, po.y), po.y),
                                                 __dmul_rn(__dmul_rn(m.reg_i, qo.y), qo.y)));
    }
  }
  return lane_loss;
}

// ------------------------------------------------------------------------------------------------
// K1: persistent wavefront SGD.  The ratings are sorted by dependency level; ratings of one level
// share neither a user nor an item, so they run concurrently with plain loads/stores; a grid-wide
// barrier separates levels.  Launched cooperatively with one CTA per SM slot.
// block_partial[blockIdx.x] receives the CTA's loss partial (reduced in fixed order by K3).
// ------------------------------------------------------------------------------------------------
template <int MODEL, int LPR, int V, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    sgd_wavefront_kernel(DeviceModel m, RatingStream s, double lr, unsigned* barrier_counter,
                         double* block_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int G = 32 / LPR;  // groups per warp
  constexpr int WARPS = THREADS / 32;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane % LPR;
  const int gw = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (gw * LPR));
  // per-group scratch, padded by 16 B so the G groups of a warp read distinct banks
  const int prod_stride = m.Fp + 2;
  double* prod = reinterpret_cast<double*>(smem_raw) + (size_t)(warp * G + gw) * prod_stride;

  const int64_t total_groups = (int64_t)gridDim.x * WARPS * G;
  const int64_t my_group = ((int64_t)blockIdx.x * WARPS + warp) * G + gw;

  double acc = 0.0;
  for (int L = 0; L < s.num_levels; L++) {
    const int64_t beg = s.level_start[L];
    const int64_t end = s.level_start[L + 1];
    for (int64_t n = beg + my_group; n < end; n += total_groups) {
      const int u = __ldg(s.u + n);
      const int j = __ldg(s.j + n);
      const int ctx = s.ctx ? __ldg(s.ctx + n) : 0;
      const double r = __ldg(s.r + n);
      acc = __dadd_rn(acc, rating_update<MODEL, LPR, V>(m, u, j, ctx, r, lr, prod, gl, gmask));
    }
    if (L + 1 < s.num_levels) grid_barrier(barrier_counter, (unsigned)(L + 1) * gridDim.x);
  }

  // CTA loss partial, fixed order: lanes -> warps -> block
  acc = warp_sum_f64(acc);
  __shared__ double warp_sum[WARPS];
  if (lane == 0) warp_sum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < WARPS; w++) t += warp_sum[w];
    block_partial[blockIdx.x] = t;
  }
}

// K1s: serial kernel (one warp walks the ratings in the reference's order).  Used for CAMF_C in EXACT
// mode, where the shared condBias vector makes every pair of ratings conflict (CAMF_C.java:107-113).
template <int MODEL, int V>
__global__ void __launch_bounds__(32, 1)
    sgd_serial_kernel(DeviceModel m, RatingStream s, int64_t nnz, double lr, double* block_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* prod = reinterpret_cast<double*>(smem_raw);
  const int lane = threadIdx.x;
  double acc = 0.0;
  for (int64_t n = 0; n < nnz; n++) {
    const int u = __ldg(s.u + n);
    const int j = __ldg(s.j + n);
    const int ctx = s.ctx ? __ldg(s.ctx + n) : 0;
    const double r = __ldg(s.r + n);
    acc = __dadd_rn(acc, rating_update<MODEL, 32, V>(m, u, j, ctx, r, lr, prod, lane, 0xffffffffu));
    __syncwarp();
    __threadfence_block();
  }
  acc = warp_sum_f64(acc);
  if (lane == 0) block_partial[0] = acc;
}

// K3: final loss reduction in fixed order, then `loss *= 0.5` (CAMF_CI.java:124).
__global__ void loss_finalize_kernel(const double* block_partial, int n, double* loss_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; i++) t += block_partial[i];
    *loss_out = t * 0.5;
  }
}

// ------------------------------------------------------------------------------------------------
// K5: batched predict(u, j, c, bound) -- one thread per test rating, dot product in f order so the
// value is bit-identical to Java's (Recommender.java:306-317 + the model's predict()).
// ------------------------------------------------------------------------------------------------
template <int MODEL>
__device__ __forceinline__ double predict_one(const DeviceModel& m, int u, int j, int ctx) {
  const double* p = m.P + (int64_t)u * m.Fp;
  const double* q = m.Q + (int64_t)j * m.Fp;
  double dot = 0.0;
  for (int f = 0; f < m.F; f++) dot = __dadd_rn(dot, __dmul_rn(p[f], q[f]));
  double pred;
  if (MODEL == M_PMF) pred = dot;
  if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C)
    pred = __dadd_rn(__dadd_rn(__dadd_rn(m.global_mean, m.user_bias[u]), m.item_bias[j]), dot);
  if (MODEL == M_CAMF_CI) pred = __dadd_rn(__dadd_rn(m.global_mean, m.user_bias[u]), dot);
  if (MODEL == M_CAMF_CU) pred = __dadd_rn(__dadd_rn(m.global_mean, m.item_bias[j]), dot);
  if (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU) {
    for (int d = 0; d < m.Dmax; d++) {
      const int cond = m.ctx_tab[(int64_t)ctx * m.Dmax + d];
      if (cond < 0) continue;
      const double b = MODEL == M_CAMF_C    ? m.cond_bias[cond]
                       : MODEL == M_CAMF_CI ? m.ic_bias[(int64_t)j * m.C + cond]
                                            : m.uc_bias[(int64_t)u * m.C + cond];
      pred = __dadd_rn(pred, b);
    }
  }
  return pred;
}

template <int MODEL>
__global__ void predict_kernel(DeviceModel m, int64_t n, const int32_t* u, const int32_t* j,
                               const int32_t* ctx, int bound, double min_rate, double max_rate,
                               double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pred = predict_one<MODEL>(m, u[i], j[i], ctx ? ctx[i] : 0);
  if (bound) {
    if (pred > max_rate) pred = max_rate;
    if (pred < min_rate) pred = min_rate;
  }
  out[i] = pred;
}

}  // namespace cars
