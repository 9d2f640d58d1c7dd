// sgd_kernels.cuh -- sm_100a device code of the CARSKit SGD hot path.
//
// One "rating update" is the body of the per-rating loop of buildModel()
// (reference: src/carskit/alg/cars/adaptation/dependent/dev/CAMF_CI.java:80-121 and the sibling
// classes CAMF_C.java:80-128, CAMF_CU.java:77-118, CAMF_CUCI.java:81-127, baseline/cf/BiasedMF.java:63-99,
// PMF.java:52-71).
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

enum : int { M_PMF = 0, M_BIASEDMF = 1, M_CAMF_C = 2, M_CAMF_CI = 3, M_CAMF_CU = 4, M_CAMF_CUCI = 5, M_CAMF_ICS = 6, M_CAMF_LCS = 7, M_CAMF_MCS = 8, M_SVDPP = 9 };

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
  double* cc_sim;             // CAMF_ICS: [C x C], the (max, min) cell of every unordered pair is the live one (SymmMatrix)
  const int32_t* empty_cond;  // CAMF_ICS / LCS / MCS: [Dmax] the "na" condition of every dimension (EmptyContextConditions)
  double* cf_lcs;             // CAMF_LCS: [C x numF] latent condition vectors (cfMatrix_LCS)
  double* c_mcs;              // CAMF_MCS: [C] condition positions (cVector_MCS)
  double mcs_upbound, mcs_lowbound;  // CAMF_MCS.java:44-45: 1 / sqrt(numContextDims), 1 / 10^100
  int32_t numF;               // CAMF_LCS: `-f`
  double* Y;                  // SVD++: [num_items x Fp] implicit-feedback item factors
  const int32_t* ui_ptr;      // SVD++: [num_users + 1] userItemsCache in CSR form: the items user u rated, ascending
  const int32_t* ui_items;    //        [nnz]
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

// One training rating in the reference's iteration order, plus its position in the two dependency
// chains it belongs to: ku / kj = number of EARLIER ratings (reference order) of the same user / item.
struct __align__(16) RatingRec {
  int32_t u, j, ctx, ku;
  int32_t kj, pad;
  double r;
};
static_assert(sizeof(RatingRec) == 32, "RatingRec must be 32 bytes");

// Dataflow schedule (K1d): ratings stay in reference order, cut into chunks of consecutive ratings.
struct DataflowStream {
  const RatingRec* rec;        // [nnz]
  const int64_t* chunk_start;  // [num_chunks + 1]
  double* chunk_loss;          // [num_chunks]  loss partial of each chunk (deterministic final reduce)
  unsigned* done_u;            // [num_users]   ratings of user u completed this epoch
  unsigned* done_j;            // [num_items]   ratings of item j completed this epoch
  unsigned* counter;           // next chunk to hand out
  uint32_t num_chunks;
};

// User-side state a group keeps in registers while consecutive ratings share the user.
template <int V>
struct UserRegs {
  double2 p[V];
  double bu;
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
// 256-bit L2 accesses (LDG/STG.E.ENL2.256 on sm_100): two adjacent 16-byte chunks in one instruction
__device__ __forceinline__ void ld_cg_f64x4(const double* p, double2& a, double2& b) {
  asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ void st_cg_f64x4(double* p, double2 a, double2 b) {
  asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
// Which 16-byte chunk of a row lane `gl` holds in slot v.  Narrow: chunks gl, gl + LPR, ... (a group covers 16 * LPR
// contiguous bytes per instruction).  WIDE: the lane owns 32-byte pieces gl, gl + LPR, ... -- slots 2w and 2w + 1 are
// the two halves of piece w -- so that one 256-bit access moves both (rows must be 32-byte aligned: Fp % 4 == 0).
template <int LPR, bool WIDE>
__device__ __forceinline__ int chunk_of(int gl, int v) {
  return WIDE ? 2 * (gl + (v >> 1) * LPR) + (v & 1) : gl + v * LPR;
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_cg_f64(double* p, double v) {
  asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// Model accesses by cache policy.  L1 = false: ld/st .cg (L2 only) -- required whenever another SM may have
// written the row (every multi-CTA schedule).  L1 = true: default caching, for the single-warp serial kernel
// where nobody else touches the model and an L1 hit saves an L2 round trip on every dependent load.
template <bool L1>
__device__ __forceinline__ double2 ldm_f64x2(const double* p) {
  if (L1) return *reinterpret_cast<const double2*>(p);
  return ld_cg_f64x2(p);
}
template <bool L1>
__device__ __forceinline__ double ldm_f64(const double* p) {
  if (L1) return *p;
  return ld_cg_f64(p);
}
template <bool L1>
__device__ __forceinline__ void stm_f64x2(double* p, double2 v) {
  if (L1) *reinterpret_cast<double2*>(p) = v;
  else st_cg_f64x2(p, v);
}
template <bool L1>
__device__ __forceinline__ void stm_f64(double* p, double v) {
  if (L1) *p = v;
  else st_cg_f64(p, v);
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Producer side: a RELEASE fence is all the release pattern needs.  fence.acq_rel.gpu lowers to MEMBAR.ALL.GPU + ERRBAR +
// CCTL.IVALL; fence.release.gpu to the same without the CCTL.IVALL (an L1 invalidation per rating that orders nothing the
// producer needs -- its own later loads of other rows go through the consumer pattern of THEIR ratings).
__device__ __forceinline__ void fence_release_gpu() { asm volatile("fence.release.gpu;" ::: "memory"); }
// Consumer side of the completion counters.  The producer publishes with the release pattern (fence.release.gpu, then
// relaxed stores).  Once the relaxed poll has succeeded, the lanes that polled re-read their counter with
// ld.acquire.gpu: that load observes the released value (or a later one), so it synchronizes-with the producer's fence
// -- the PTX-formal acquire pattern -- and the __syncwarp that follows extends the ordering to the group's other lanes
// before any of them gathers the rows.  It lowers to LDG.STRONG.GPU + CCTL.IVALL, once per RATING (not per poll: the
// spinning polls stay relaxed).  Measured cost at config 3: 40.38 ms/epoch against 40.28 ms without it
// (profiles/r2/bench_r2_exact100M_relaxed_vs_acquire.txt), i.e. nothing, so it is unconditional.
// -DCARS_RELAXED_POLL builds the variant without it (scripts/build_relaxed.sh), kept only to reproduce that A/B.
__device__ __forceinline__ void acquire_after_poll(const unsigned* p, bool mine) {
#ifndef CARS_RELAXED_POLL
  if (mine) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  }
#else
  (void)p;
  (void)mine;
#endif
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
// `us` holds the user's factor row (and userBias) in registers: it is read from memory when `load_user`
// and written back when `store_user`, so a run of consecutive ratings of one user touches P[u] once.
// Returns this lane's contribution to the epoch loss (un-halved).
// ------------------------------------------------------------------------------------------------
// Item-side (and per-rating) operands of one update, gathered before the arithmetic starts.
template <int V>
struct Operands {
  double2 q[V];
  double bj;       // itemBias[j]
  double cb;       // this lane's condition-bias cell (lane d owns context dimension d)
  double* cb_ptr;  // its address; nullptr when the lane owns none
  double cb2;      // CAMF_CUCI only: cb = icBias[j][cond], cb2 = ucBias[u][cond]
  double* cb2_ptr;
};

// Scalars of one update: itemBias[j] and the lane's condition-bias cell.
constexpr int kCondUnknown = -2;  // "look the condition id up" (a pre-fetched id is >= -1)

template <int MODEL, int LPR, int V, bool L1 = false>
__device__ __forceinline__ void gather_scalars(const DeviceModel& m, int u, int j, int ctx, int gl, Operands<V>& o,
                                               int cond_prefetched = kCondUnknown) {
  constexpr bool kItemBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CU);
  constexpr bool kHasCond = (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI);
  o.bj = 0.0;
  if (kItemBias) o.bj = ldm_f64<L1>(m.item_bias + j);
  o.cb_ptr = nullptr;
  o.cb = 0.0;
  o.cb2_ptr = nullptr;
  o.cb2 = 0.0;
  if (kHasCond && gl < m.Dmax) {
    const int cond = cond_prefetched != kCondUnknown ? cond_prefetched : __ldg(m.ctx_tab + (int64_t)ctx * m.Dmax + gl);
    if (cond >= 0) {
      if (MODEL == M_CAMF_C) o.cb_ptr = m.cond_bias + cond;
      if (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) o.cb_ptr = m.ic_bias + (int64_t)j * m.C + cond;
      if (MODEL == M_CAMF_CU) o.cb_ptr = m.uc_bias + (int64_t)u * m.C + cond;
      o.cb = ldm_f64<L1>(o.cb_ptr);
      if (MODEL == M_CAMF_CUCI) {
        o.cb2_ptr = m.uc_bias + (int64_t)u * m.C + cond;
        o.cb2 = ldm_f64<L1>(o.cb2_ptr);
      }
    }
  }
}

// Factor rows straight from global memory (L2) into registers.
// FIXF > 0: the number of factors is this compile-time constant (F = Fp = FIXF): the bounds checks fold away and the
// in-order dot product unrolls into straight-line code.
template <int MODEL, int LPR, int V, bool L1 = false, bool WIDE = false, int FIXF = 0>
__device__ __forceinline__ void gather_rows(const DeviceModel& m, int u, int j, int gl, UserRegs<V>& us,
                                            Operands<V>& o, bool load_user) {
  constexpr bool kUserBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CI);
  const int Fp = FIXF > 0 ? FIXF : m.Fp;
  const double* prow = m.P + (int64_t)u * Fp;
  const double* qrow = m.Q + (int64_t)j * Fp;
  if (WIDE) {
#pragma unroll
    for (int w = 0; w < V / 2; w++) {
      const int c = chunk_of<LPR, true>(gl, 2 * w);  // first 16-byte chunk of the lane's 32-byte piece
      if (2 * c < Fp) {
        ld_cg_f64x4(qrow + 2 * c, o.q[2 * w], o.q[2 * w + 1]);
        if (load_user) ld_cg_f64x4(prow + 2 * c, us.p[2 * w], us.p[2 * w + 1]);
      } else {
        o.q[2 * w] = o.q[2 * w + 1] = make_double2(0.0, 0.0);
        us.p[2 * w] = us.p[2 * w + 1] = make_double2(0.0, 0.0);
      }
    }
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = gl + v * LPR;  // chunk index
      if (2 * c < Fp) {
        o.q[v] = ldm_f64x2<L1>(qrow + 2 * c);
        if (load_user) us.p[v] = ldm_f64x2<L1>(prow + 2 * c);
      } else {
        o.q[v] = make_double2(0.0, 0.0);
        us.p[v] = make_double2(0.0, 0.0);
      }
    }
  }
  if (kUserBias && load_user) us.bu = ldm_f64<L1>(m.user_bias + u);
}

// ------------------------------------------------------------------------------------------------
// The arithmetic of one rating update and its scatter, by one group of LPR lanes holding the operands
// in registers.  V = 16-byte chunks per lane (LPR * V * 2 >= Fp).  `prod` = the group's shared-memory
// scratch of Fp doubles (for the in-order dot product).  The user's row (and userBias) live in `us`
// and are written back only when `store_user`.  Returns this lane's contribution to the epoch loss.
// ------------------------------------------------------------------------------------------------
template <int MODEL, int LPR, int V, bool L1 = false, bool WIDE = false, int FIXF = 0>
__device__ __forceinline__ double compute_scatter(const DeviceModel& m, int u, int j, int ctx, double r, double lr,
                                                  double* prod, int gl, unsigned gmask, UserRegs<V>& us,
                                                  const Operands<V>& o, bool store_user) {
  constexpr bool kUserBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CI);
  constexpr bool kItemBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CU);
  constexpr bool kHasCond = (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI);
  const int Fp = FIXF > 0 ? FIXF : m.Fp;
  const int Dmax = m.Dmax;
  double* prow = m.P + (int64_t)u * Fp;
  double* qrow = m.Q + (int64_t)j * Fp;
  const double2* q = o.q;
  const double bj = o.bj;
  const double cb = o.cb;
  double* cb_ptr = o.cb_ptr;
  const double bu = kUserBias ? us.bu : 0.0;

  // ---- dot product in f order (DenseMatrix.rowMult) --------------------------------------------
#pragma unroll
  for (int v = 0; v < V; v++) {
    const int c = chunk_of<LPR, WIDE>(gl, v);
    if (2 * c < Fp) {
      double2 t = make_double2(__dmul_rn(us.p[v].x, q[v].x), __dmul_rn(us.p[v].y, q[v].y));
      *reinterpret_cast<double2*>(prod + 2 * c) = t;
    }
  }
  __syncwarp(gmask);
  double dot = 0.0;
  {
    const int F = FIXF > 0 ? FIXF : m.F;
    const int F2 = F & ~1;
#pragma unroll
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
  if (MODEL == M_CAMF_CUCI) pred = __dadd_rn(m.global_mean, dot);

  double lane_loss = 0.0;
  if (kHasCond) {
    // conditions beyond the first LPR (Dmax > LPR) are handled by the slow path below
    const int D1 = Dmax < LPR ? Dmax : LPR;
    // CAMF_CUCI.java:71: pred += icBias(j,cond) + ucBias(u,cond) -- the two cells are added first
    const double cbs = MODEL == M_CAMF_CUCI ? __dadd_rn(cb, o.cb2) : cb;
    for (int d = 0; d < D1; d++) {
      const double b = shfl_f64(gmask, cbs, d, LPR);
      pred = __dadd_rn(pred, b);  // adding a padded slot adds +0.0: exact
    }
    for (int d = LPR; d < Dmax; d++) {  // rare: more context dimensions than lanes in a group
      const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
      if (cond >= 0) {
        const double* bp = MODEL == M_CAMF_C    ? m.cond_bias + cond
                           : (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) ? m.ic_bias + (int64_t)j * m.C + cond
                                                : m.uc_bias + (int64_t)u * m.C + cond;
        double b = ldm_f64<L1>(bp);
        if (MODEL == M_CAMF_CUCI) b = __dadd_rn(b, ldm_f64<L1>(m.uc_bias + (int64_t)u * m.C + cond));
        pred = __dadd_rn(pred, b);
      }
    }
  }
  const double e = __dsub_rn(r, pred);

  // ---- bias steps -----------------------------------------------------------------------------
  if (kUserBias) {  // every lane keeps the same copy of bu; lane 0 owns the loss term and the write-back
    const double sgd = __dsub_rn(e, __dmul_rn(m.reg_b, bu));
    us.bu = __dadd_rn(bu, __dmul_rn(lr, sgd));
    if (store_user && gl == 0) stm_f64<L1>(m.user_bias + u, us.bu);
  }
  if (gl == 0) {
    lane_loss = __dmul_rn(e, e);
    if (kUserBias) lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bu), bu));
    if (kItemBias) {
      const double sgd = __dsub_rn(e, __dmul_rn(m.reg_b, bj));
      stm_f64<L1>(m.item_bias + j, __dadd_rn(bj, __dmul_rn(lr, sgd)));
      lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bj), bj));
    }
  }
  if (kHasCond) {
    if (cb_ptr != nullptr) {
      const double sgd = __dsub_rn(e, __dmul_rn(m.reg_c, cb));
      const double step = __dmul_rn(lr, sgd);
      stm_f64<L1>(cb_ptr, __dadd_rn(cb, step));
      // CAMF_C.java:115 adds regB * sum(bc) (not squared); CI/CU add regC * sum(b^2) (:108 / :105)
      if (MODEL == M_CAMF_C)
        lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_b, cb));
      else
        lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_c, __dmul_rn(cb, cb)));
      if (MODEL == M_CAMF_CUCI) {  // CAMF_CUCI.java:104-112: the user-context cell, same step with Buc
        const double cb2 = o.cb2;
        const double step2 = __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_c, cb2)));
        stm_f64<L1>(o.cb2_ptr, __dadd_rn(cb2, step2));
        lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_c, __dmul_rn(cb2, cb2)));
      }
    }
    // dims beyond the group's lanes: every lane READ those cells for `pred` above; lane 0 rewrites them below
    if (Dmax > LPR) __syncwarp(gmask);
    if (gl == 0) {
      for (int d = LPR; d < Dmax; d++) {
        const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
        if (cond < 0) continue;
        double* bp = MODEL == M_CAMF_C    ? m.cond_bias + cond
                     : (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) ? m.ic_bias + (int64_t)j * m.C + cond
                                          : m.uc_bias + (int64_t)u * m.C + cond;
        const double b = ldm_f64<L1>(bp);
        const double step = __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_c, b)));
        stm_f64<L1>(bp, __dadd_rn(b, step));
        lane_loss = __dadd_rn(lane_loss, MODEL == M_CAMF_C ? __dmul_rn(m.reg_b, b)
                                                             : __dmul_rn(m.reg_c, __dmul_rn(b, b)));
        if (MODEL == M_CAMF_CUCI) {
          double* bp2 = m.uc_bias + (int64_t)u * m.C + cond;
          const double b2 = ldm_f64<L1>(bp2);
          const double step2 = __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_c, b2)));
          stm_f64<L1>(bp2, __dadd_rn(b2, step2));
          lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_c, __dmul_rn(b2, b2)));
        }
      }
    }
  }

  // ---- factor steps (both from the OLD p, q) and scatter -------------------------------------------
  // The regularisation terms of the loss, sum_f regU*p^2 + regI*q^2, are accumulated as regU*sum(p^2) +
  // regI*sum(q^2) with FMAs: `loss` is a report value (1e-11 relative), the model is not.
  double sp = 0.0, sq = 0.0;
  double2 q_even = make_double2(0.0, 0.0), p_even = make_double2(0.0, 0.0);  // WIDE: first half of a 32-byte piece
#pragma unroll
  for (int v = 0; v < V; v++) {
    const int c = chunk_of<LPR, WIDE>(gl, v);
    if (2 * c < Fp) {
      const double2 po = us.p[v], qo = q[v];
      double2 pn, qn;
      pn.x = __dadd_rn(po.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.x), __dmul_rn(m.reg_u, po.x))));
      qn.x = __dadd_rn(qo.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, po.x), __dmul_rn(m.reg_i, qo.x))));
      pn.y = __dadd_rn(po.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.y), __dmul_rn(m.reg_u, po.y))));
      qn.y = __dadd_rn(qo.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, po.y), __dmul_rn(m.reg_i, qo.y))));
      us.p[v] = pn;
      if (WIDE) {
        if ((v & 1) == 0) {
          q_even = qn;
          p_even = pn;
        } else {  // c - 1 is the piece's first chunk: one 256-bit store per row
          st_cg_f64x4(qrow + 2 * (c - 1), q_even, qn);
          if (store_user) st_cg_f64x4(prow + 2 * (c - 1), p_even, pn);
        }
      } else {
        stm_f64x2<L1>(qrow + 2 * c, qn);
        if (store_user) stm_f64x2<L1>(prow + 2 * c, pn);
      }
      sp = fma(po.x, po.x, sp);
      sq = fma(qo.x, qo.x, sq);
      sp = fma(po.y, po.y, sp);
      sq = fma(qo.y, qo.y, sq);
    }
  }
  lane_loss = __dadd_rn(lane_loss, fma(m.reg_u, sp, __dmul_rn(m.reg_i, sq)));
  return lane_loss;
}

// gather + compute + scatter with plain loads (wavefront, serial and dataflow kernels).
template <int MODEL, int LPR, int V, bool L1 = false>
__device__ __forceinline__ double rating_update(const DeviceModel& m, int u, int j, int ctx, double r,
                                                double lr, double* prod, int gl /*lane in group*/,
                                                unsigned gmask /*lanes of this group*/, UserRegs<V>& us,
                                                bool load_user, bool store_user) {
  Operands<V> o;
  gather_rows<MODEL, LPR, V, L1>(m, u, j, gl, us, o, load_user);
  gather_scalars<MODEL, LPR, V, L1>(m, u, j, ctx, gl, o);
  return compute_scatter<MODEL, LPR, V, L1>(m, u, j, ctx, r, lr, prod, gl, gmask, us, o, store_user);
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
      UserRegs<V> us;
      acc = __dadd_rn(acc, rating_update<MODEL, LPR, V>(m, u, j, ctx, r, lr, prod, gl, gmask, us, true, true));
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
    // one warp owns the whole model, so its accesses go through L1 (rating_update<.., L1 = true>)
    UserRegs<V> us;
    acc = __dadd_rn(acc, rating_update<MODEL, 32, V, true>(m, u, j, ctx, r, lr, prod, lane, 0xffffffffu, us, true, true));
    __syncwarp();
    __threadfence_block();
  }
  acc = warp_sum_f64(acc);
  if (lane == 0) block_partial[0] = acc;
}

// K1s-ICS: CAMF_ICS.buildModel (sim/CAMF_ICS.java:62-124), one warp in reference order.  Every rating reads and rewrites
// cells of the ONE condition-similarity matrix all ratings share (like CAMF_C's condBias), so EXACT mode is a single chain:
//   dot = P[u].Q[j];  pred = dot;  simc = 1
//   for i: sim = (cond_i != empty_i) ? cc(cond_i, empty_i) : 1;  [simc *= sim];  loss += regC*sim*sim;  pred = pred*sim
//   e = r - pred
//   for the cells read:  cc = sim + lr * (e*dot*simc/sim - regC*sim)
//   P[u][f] += lr * (e*q*simc - regU*p);  Q[j][f] += lr * (e*p*simc - regI*q)
template <int V>
__global__ void __launch_bounds__(32, 1)
    sgd_serial_ics_kernel(DeviceModel m, RatingStream s, int64_t nnz, double lr, double* block_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* prod = reinterpret_cast<double*>(smem_raw);
  const int lane = threadIdx.x;
  const int Fp = m.Fp, Dmax = m.Dmax;
  double acc = 0.0;
  for (int64_t n = 0; n < nnz; n++) {
    const int u = __ldg(s.u + n), j = __ldg(s.j + n), ctx = __ldg(s.ctx + n);
    const double r = __ldg(s.r + n);
    double* prow = m.P + (int64_t)u * Fp;
    double* qrow = m.Q + (int64_t)j * Fp;
    double2 p[V], q[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = lane + v * 32;
      if (2 * c < Fp) {
        p[v] = *reinterpret_cast<const double2*>(prow + 2 * c);
        q[v] = *reinterpret_cast<const double2*>(qrow + 2 * c);
        *reinterpret_cast<double2*>(prod + 2 * c) = make_double2(__dmul_rn(p[v].x, q[v].x), __dmul_rn(p[v].y, q[v].y));
      }
    }
    __syncwarp();
    double dot = 0.0;
    for (int f = 0; f < m.F; f++) dot = __dadd_rn(dot, prod[f]);
    __syncwarp();
    double pred = dot, simc = 1.0, lane_loss = 0.0;
    for (int d = 0; d < Dmax; d++) {  // warp-uniform: every lane reads the same cells
      const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
      if (cond < 0) continue;
      const int e2 = __ldg(m.empty_cond + d);
      double sim = 1.0;
      if (cond != e2) {
        const int hi = cond >= e2 ? cond : e2, lo = cond >= e2 ? e2 : cond;
        sim = m.cc_sim[(int64_t)hi * m.C + lo];
        simc = __dmul_rn(simc, sim);
      }
      if (lane == 0) lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_c, sim), sim));
      pred = __dmul_rn(pred, sim);
    }
    const double e = __dsub_rn(r, pred);
    if (lane == 0) lane_loss = __dadd_rn(lane_loss, __dmul_rn(e, e));
    __syncwarp();  // every lane has read the similarity cells before lane 0 rewrites them
    if (lane == 0) {
      for (int d = 0; d < Dmax; d++) {
        const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
        if (cond < 0) continue;
        const int e2 = __ldg(m.empty_cond + d);
        if (cond == e2) continue;
        const int hi = cond >= e2 ? cond : e2, lo = cond >= e2 ? e2 : cond;
        double* cell = m.cc_sim + (int64_t)hi * m.C + lo;
        const double sim = *cell;  // (the cells of one rating are distinct: still the value read above)
        const double t = __dsub_rn(__ddiv_rn(__dmul_rn(__dmul_rn(e, dot), simc), sim), __dmul_rn(m.reg_c, sim));
        *cell = __dadd_rn(sim, __dmul_rn(lr, t));
      }
    }
    double sp = 0.0, sq = 0.0;
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = lane + v * 32;
      if (2 * c < Fp) {
        const double2 po = p[v], qo = q[v];
        double2 pn, qn;
        pn.x = __dadd_rn(po.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, qo.x), simc), __dmul_rn(m.reg_u, po.x))));
        qn.x = __dadd_rn(qo.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, po.x), simc), __dmul_rn(m.reg_i, qo.x))));
        pn.y = __dadd_rn(po.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, qo.y), simc), __dmul_rn(m.reg_u, po.y))));
        qn.y = __dadd_rn(qo.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, po.y), simc), __dmul_rn(m.reg_i, qo.y))));
        *reinterpret_cast<double2*>(prow + 2 * c) = pn;
        *reinterpret_cast<double2*>(qrow + 2 * c) = qn;
        sp = fma(po.x, po.x, sp); sq = fma(qo.x, qo.x, sq);
        sp = fma(po.y, po.y, sp); sq = fma(qo.y, qo.y, sq);
      }
    }
    acc = __dadd_rn(acc, __dadd_rn(lane_loss, fma(m.reg_u, sp, __dmul_rn(m.reg_i, sq))));
    __syncwarp();
    __threadfence_block();
  }
  acc = warp_sum_f64(acc);
  if (lane == 0) block_partial[0] = acc;
}

// K1s-SVD++: SVDPlusPlus.buildModel (baseline/cf/SVDPlusPlus.java:55-124), one warp in reference order.  Every rating of
// user u reads and rewrites Y[k] for ALL items k the user rated, so two ratings conflict whenever their users share any
// item: the dependency DAG is ~1 rating wide (profiles/r2/svdpp_dag_width.txt) and EXACT mode is a single chain.  Inside a
// rating the work is parallel: lanes own factors (16-byte chunks) for the row updates and the sum over Y, and own ITEMS for
// the |items(u)| dot products Y[k].Q[j] of the prediction, which are then added in item order.
template <int V>
__global__ void __launch_bounds__(32, 1)
    sgd_serial_svdpp_kernel(DeviceModel m, RatingStream s, int64_t nnz, double lr, double* block_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* prod = reinterpret_cast<double*>(smem_raw);  // [Fp + 2] products for the in-order dot
  double* qs = prod + m.Fp + 2;                        // [Fp] Q[j] for the lanes' Y[k].Q[j] chains
  const int lane = threadIdx.x;
  const int Fp = m.Fp, F = m.F;
  double acc = 0.0;
  for (int64_t n = 0; n < nnz; n++) {
    const int u = __ldg(s.u + n), j = __ldg(s.j + n);
    const double r = __ldg(s.r + n);
    const int ia = __ldg(m.ui_ptr + u), ib = __ldg(m.ui_ptr + u + 1);
    const double w = __dsqrt_rn((double)(ib - ia));
    double* prow = m.P + (int64_t)u * Fp;
    double* qrow = m.Q + (int64_t)j * Fp;
    double2 p[V], q[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = lane + v * 32;
      if (2 * c < Fp) {
        p[v] = *reinterpret_cast<const double2*>(prow + 2 * c);
        q[v] = *reinterpret_cast<const double2*>(qrow + 2 * c);
        *reinterpret_cast<double2*>(prod + 2 * c) = make_double2(__dmul_rn(p[v].x, q[v].x), __dmul_rn(p[v].y, q[v].y));
        *reinterpret_cast<double2*>(qs + 2 * c) = q[v];
      }
    }
    __syncwarp();
    double dot = 0.0;
    for (int f = 0; f < F; f++) dot = __dadd_rn(dot, prod[f]);
    const double bu = m.user_bias[u], bj = m.item_bias[j];
    double pred = __dadd_rn(__dadd_rn(__dadd_rn(m.global_mean, bu), bj), dot);
    for (int base = ia; base < ib; base += 32) {  // 32 items at a time: lane = item, then added in item order
      double t = 0.0;
      if (base + lane < ib) {
        const double* yrow = m.Y + (int64_t)__ldg(m.ui_items + base + lane) * Fp;
        double yq = 0.0;
        for (int f = 0; f < F; f++) yq = __dadd_rn(yq, __dmul_rn(yrow[f], qs[f]));
        t = __ddiv_rn(yq, w);
      }
      const int cnt = ib - base < 32 ? ib - base : 32;
      for (int i = 0; i < cnt; i++)
        pred = __dadd_rn(pred, __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(t), i), __shfl_sync(0xffffffffu, __double2loint(t), i)));
    }
    const double e = __dsub_rn(r, pred);
    double lane_loss = 0.0;
    if (lane == 0) {
      lane_loss = __dmul_rn(e, e);
      m.user_bias[u] = __dadd_rn(bu, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_b, bu))));
      lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bu), bu));
      m.item_bias[j] = __dadd_rn(bj, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_b, bj))));
      lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bj), bj));
    }
    __syncwarp();  // every lane has read Y for the prediction before anyone rewrites it
    double sp = 0.0, sq = 0.0, sy = 0.0;
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = lane + v * 32;
      if (2 * c < Fp) {
        double2 sum = make_double2(0.0, 0.0);  // sum_ys[f], items in order (SVDPlusPlus.java:85-92)
        for (int i = ia; i < ib; i++) {
          const double2 y = *reinterpret_cast<const double2*>(m.Y + (int64_t)__ldg(m.ui_items + i) * Fp + 2 * c);
          sum.x = __dadd_rn(sum.x, y.x);
          sum.y = __dadd_rn(sum.y, y.y);
        }
        if (w > 0.0) { sum.x = __ddiv_rn(sum.x, w); sum.y = __ddiv_rn(sum.y, w); }
        const double2 po = p[v], qo = q[v];
        double2 pn, qn;
        pn.x = __dadd_rn(po.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.x), __dmul_rn(m.reg_u, po.x))));
        pn.y = __dadd_rn(po.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.y), __dmul_rn(m.reg_u, po.y))));
        qn.x = __dadd_rn(qo.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, __dadd_rn(po.x, sum.x)), __dmul_rn(m.reg_i, qo.x))));
        qn.y = __dadd_rn(qo.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, __dadd_rn(po.y, sum.y)), __dmul_rn(m.reg_i, qo.y))));
        *reinterpret_cast<double2*>(prow + 2 * c) = pn;
        *reinterpret_cast<double2*>(qrow + 2 * c) = qn;
        sp = fma(po.x, po.x, sp); sq = fma(qo.x, qo.x, sq);
        sp = fma(po.y, po.y, sp); sq = fma(qo.y, qo.y, sq);
        const double gx = __ddiv_rn(__dmul_rn(e, qo.x), w), gy = __ddiv_rn(__dmul_rn(e, qo.y), w);  // euj * qjf / w, the OLD qjf
        for (int i = ia; i < ib; i++) {
          double2* yp = reinterpret_cast<double2*>(m.Y + (int64_t)__ldg(m.ui_items + i) * Fp + 2 * c);
          const double2 y = *yp;
          double2 yn;
          yn.x = __dadd_rn(y.x, __dmul_rn(lr, __dsub_rn(gx, __dmul_rn(m.reg_u, y.x))));
          yn.y = __dadd_rn(y.y, __dmul_rn(lr, __dsub_rn(gy, __dmul_rn(m.reg_u, y.y))));
          *yp = yn;
          sy = fma(y.x, y.x, sy);
          sy = fma(y.y, y.y, sy);
        }
      }
    }
    acc = __dadd_rn(acc, __dadd_rn(lane_loss, fma(m.reg_u, __dadd_rn(sp, sy), __dmul_rn(m.reg_i, sq))));
    __syncwarp();
    __threadfence_block();
  }
  acc = warp_sum_f64(acc);
  if (lane == 0) block_partial[0] = acc;
}

// K1s-LCS / K1s-MCS: CAMF_LCS.buildModel (sim/CAMF_LCS.java:66-146) and CAMF_MCS.buildModel (sim/CAMF_MCS.java:71-167), one
// warp in reference order -- like CAMF_ICS every rating reads and rewrites cells all ratings share (the condition vectors /
// positions of the dimensions' "na" conditions), so EXACT mode is a single chain.  The similarity part is warp-uniform
// (every lane computes the same values from the same cells); lane 0 owns the loss terms and the rewrites.
//   LCS: sim_d = cf[cond_d] . cf[na_d] (f ascending);  pred = dot * prod sim_d;  cf rows += lr * (e*dot*simc*other/sim - regC*own)
//   MCS: dist = sqrt(sum diff_d^2), diff_d = pos[cond_d] - pos[na_d];  pred = dot * (1 - dist);  positions moved along diff,
//        clamped to (lowbound, upbound); a zero dist becomes lowbound in the update loop AND stays so for the factor steps
// DenseMatrix.rowMult(cfMatrix_LCS, a, cfMatrix_LCS, b) by one warp: the products are formed lane-parallel (lane = factor), the
// sum runs f ascending through shuffles -- the same additions in the same order, without a chain of dependent loads
__device__ __forceinline__ double lcs_row_mult(const double* c1, const double* c2, int numF, int lane) {
  double sim = 0.0;
  for (int f0 = 0; f0 < numF; f0 += 32) {
    const int f = f0 + lane;
    const double pr = f < numF ? __dmul_rn(c1[f], c2[f]) : 0.0;
    const int cnt = numF - f0 < 32 ? numF - f0 : 32;
    for (int i = 0; i < cnt; i++)
      sim = __dadd_rn(sim, __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(pr), i), __shfl_sync(0xffffffffu, __double2loint(pr), i)));
  }
  return sim;
}

template <int V, int KIND /*1 = LCS, 2 = MCS*/>
__global__ void __launch_bounds__(32, 1)
    sgd_serial_sim_kernel(DeviceModel m, RatingStream s, int64_t nnz, double lr, double* block_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* prod = reinterpret_cast<double*>(smem_raw);
  const int lane = threadIdx.x;
  const int Fp = m.Fp, Dmax = m.Dmax, numF = m.numF;
  double acc = 0.0;
  for (int64_t n = 0; n < nnz; n++) {
    const int u = __ldg(s.u + n), j = __ldg(s.j + n), ctx = __ldg(s.ctx + n);
    const double r = __ldg(s.r + n);
    double* prow = m.P + (int64_t)u * Fp;
    double* qrow = m.Q + (int64_t)j * Fp;
    double2 p[V], q[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = lane + v * 32;
      if (2 * c < Fp) {
        p[v] = *reinterpret_cast<const double2*>(prow + 2 * c);
        q[v] = *reinterpret_cast<const double2*>(qrow + 2 * c);
        *reinterpret_cast<double2*>(prod + 2 * c) = make_double2(__dmul_rn(p[v].x, q[v].x), __dmul_rn(p[v].y, q[v].y));
      }
    }
    __syncwarp();
    double dot = 0.0;
    for (int f = 0; f < m.F; f++) dot = __dadd_rn(dot, prod[f]);
    __syncwarp();
    double pred = dot, simc = 1.0, dist = 0.0, lane_loss = 0.0;
    for (int d = 0; d < Dmax; d++) {  // warp-uniform
      const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
      if (cond < 0) continue;
      const int e2 = __ldg(m.empty_cond + d);
      if (KIND == 1) {
        double sim = 1.0;
        if (cond != e2) {
          sim = lcs_row_mult(m.cf_lcs + (int64_t)cond * numF, m.cf_lcs + (int64_t)e2 * numF, numF, lane);
          simc = __dmul_rn(simc, sim);
        }
        pred = __dmul_rn(pred, sim);
      } else {
        const double pos1 = m.c_mcs[cond], pos2 = m.c_mcs[e2];
        const double diff = __dsub_rn(pos1, pos2);
        dist = __dadd_rn(dist, __dmul_rn(diff, diff));
        if (lane == 0)
          lane_loss = __dadd_rn(lane_loss, __dadd_rn(__dmul_rn(__dmul_rn(m.reg_c, pos1), pos1), __dmul_rn(__dmul_rn(m.reg_c, pos2), pos2)));
      }
    }
    if (KIND == 2) {
      dist = __dsqrt_rn(dist);
      pred = __dmul_rn(pred, __dsub_rn(1.0, dist));
    }
    const double e = __dsub_rn(r, pred);
    if (lane == 0) lane_loss = __dadd_rn(lane_loss, __dmul_rn(e, e));
    __syncwarp();  // every lane has read the shared cells before they are rewritten
    if (KIND == 1) {  // LCS: lane = factor of the two condition vectors (the factors' updates are independent of each other)
      for (int d = 0; d < Dmax; d++) {
        const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
        if (cond < 0) continue;
        const int e2 = __ldg(m.empty_cond + d);
        if (cond == e2) continue;
        double* c1 = m.cf_lcs + (int64_t)cond * numF;
        double* c2 = m.cf_lcs + (int64_t)e2 * numF;
        // the rows of one rating are distinct: still the value the prediction used
        const double sim = lcs_row_mult(c1, c2, numF, lane);
        const double g = __dmul_rn(__dmul_rn(e, dot), simc);
        for (int f = lane; f < numF; f += 32) {
          const double c1f = c1[f], c2f = c2[f];
          const double d1 = __dsub_rn(__ddiv_rn(__dmul_rn(g, c2f), sim), __dmul_rn(m.reg_c, c1f));
          const double d2 = __dsub_rn(__ddiv_rn(__dmul_rn(g, c1f), sim), __dmul_rn(m.reg_c, c2f));
          c1[f] = __dadd_rn(c1f, __dmul_rn(lr, d1));
          c2[f] = __dadd_rn(c2f, __dmul_rn(lr, d2));
          lane_loss = __dadd_rn(lane_loss, __dadd_rn(__dmul_rn(__dmul_rn(m.reg_c, c1f), c1f), __dmul_rn(__dmul_rn(m.reg_c, c2f), c2f)));
        }
        __syncwarp();
      }
    }
    if (KIND == 2 && lane == 0) {
      for (int d = 0; d < Dmax; d++) {
        const int cond = __ldg(m.ctx_tab + (int64_t)ctx * Dmax + d);
        if (cond < 0) continue;
        const int e2 = __ldg(m.empty_cond + d);
        if (cond == e2) continue;
        {
          const double pos1 = m.c_mcs[cond], pos2 = m.c_mcs[e2];
          const double diff = __dsub_rn(pos1, pos2);  // the cells of one rating are distinct: the first loop's value
          if (dist == 0.0) dist = m.mcs_lowbound;
          const double t = __ddiv_rn(__dmul_rn(__dmul_rn(e, dot), diff), dist);
          double n1 = __dadd_rn(pos1, __dmul_rn(lr, __dsub_rn(t, __dmul_rn(m.reg_c, pos1))));
          double n2 = __dsub_rn(pos2, __dmul_rn(lr, __dadd_rn(t, __dmul_rn(m.reg_c, pos2))));
          n1 = n1 < 0.0 ? m.mcs_lowbound : n1;
          n1 = n1 > m.mcs_upbound ? __dsub_rn(m.mcs_upbound, m.mcs_lowbound) : n1;
          n2 = n2 < 0.0 ? m.mcs_lowbound : n2;
          n2 = n2 > m.mcs_upbound ? __dsub_rn(m.mcs_upbound, m.mcs_lowbound) : n2;
          m.c_mcs[cond] = n1;
          m.c_mcs[e2] = n2;
        }
      }
    }
    if (KIND == 2) {  // lane 0 may have replaced a zero distance
      dist = __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(dist), 0), __shfl_sync(0xffffffffu, __double2loint(dist), 0));
      simc = __dsub_rn(1.0, dist);
    }
    double sp = 0.0, sq = 0.0;
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = lane + v * 32;
      if (2 * c < Fp) {
        const double2 po = p[v], qo = q[v];
        double2 pn, qn;
        pn.x = __dadd_rn(po.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, qo.x), simc), __dmul_rn(m.reg_u, po.x))));
        qn.x = __dadd_rn(qo.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, po.x), simc), __dmul_rn(m.reg_i, qo.x))));
        pn.y = __dadd_rn(po.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, qo.y), simc), __dmul_rn(m.reg_u, po.y))));
        qn.y = __dadd_rn(qo.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(__dmul_rn(e, po.y), simc), __dmul_rn(m.reg_i, qo.y))));
        *reinterpret_cast<double2*>(prow + 2 * c) = pn;
        *reinterpret_cast<double2*>(qrow + 2 * c) = qn;
        sp = fma(po.x, po.x, sp); sq = fma(qo.x, qo.x, sq);
        sp = fma(po.y, po.y, sp); sq = fma(qo.y, qo.y, sq);
      }
    }
    acc = __dadd_rn(acc, __dadd_rn(lane_loss, fma(m.reg_u, sp, __dmul_rn(m.reg_i, sq))));
    __syncwarp();
    __threadfence_block();
  }
  acc = warp_sum_f64(acc);
  // CAMF_MCS.java:158: loss *= 0.05 (the host-side finalisation halves): a report value, 1e-11 relative
  if (lane == 0) block_partial[0] = KIND == 2 ? acc * 0.1 : acc;
}

// ------------------------------------------------------------------------------------------------
// K1d: dataflow SGD (the default schedule).  Ratings stay in the reference's iteration order and are
// cut into chunks of consecutive ratings; groups take chunks IN ORDER from a global counter and walk
// each chunk sequentially.  A rating may run once the previous rating of its user and of its item (in
// reference order) has completed: done_u[u] == ku and done_j[j] == kj.  Completion is published with a
// release store after the group's stores; waiting is a non-blocking acquire poll (a group that is not
// ready skips the turn, so the other groups of its warp keep going -- no intra-warp deadlock).
// Because chunks are handed out in order and every dependency of a rating lies earlier in reference
// order, the earliest unfinished chunk can always progress: no grid-wide barrier, no co-residency needed.
// While consecutive ratings share the user (a ratings file sorted by user) P[u] and userBias[u] stay in
// registers; Q, itemBias and icBias (tens of MB) stay L2-resident.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ RatingRec ld_rec(const RatingRec* p) {
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  RatingRec x;
  x.u = a.x; x.j = a.y; x.ctx = a.z; x.ku = a.w;
  x.kj = b.x; x.pad = 0;
  x.r = __hiloint2double(b.w, b.z);
  return x;
}

// as ld_rec, keeping the `pad` word (flags of the tagged kernel: last rating of its user / item in the epoch)
__device__ __forceinline__ RatingRec ld_rec_pad(const RatingRec* p) {
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  RatingRec x;
  x.u = a.x; x.j = a.y; x.ctx = a.z; x.ku = a.w;
  x.kj = b.x; x.pad = b.y;
  x.r = __hiloint2double(b.w, b.z);
  return x;
}

template <int MODEL, int LPR, int V, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    sgd_dataflow_kernel(DeviceModel m, DataflowStream s, double lr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int G = 32 / LPR;  // groups per warp
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane % LPR;
  const int gw = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (gw * LPR));
  const int prod_stride = m.Fp + 2;
  double* prod = reinterpret_cast<double*>(smem_raw) + (size_t)(warp * G + gw) * prod_stride;

  int64_t n = 0, end = 0;
  int64_t chunk = -1;
  bool done = false;
  int prev_u = -1;
  double acc = 0.0;
  UserRegs<V> us;
#pragma unroll
  for (int v = 0; v < V; v++) us.p[v] = make_double2(0.0, 0.0);
  us.bu = 0.0;

  for (;;) {
    if (!done && n == end) {
      if (chunk >= 0) {  // publish the finished chunk's loss partial (fixed-order tree over the group)
        double t = acc;
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) t = __dadd_rn(t, shfl_f64(gmask, t, gl + o, LPR));
        if (gl == 0) s.chunk_loss[chunk] = t;
        acc = 0.0;
      }
      unsigned c = 0;
      if (gl == 0) c = atomicAdd(s.counter, 1u);
      c = __shfl_sync(gmask, c, 0, LPR);
      if (c >= s.num_chunks) {
        done = true;
      } else {
        chunk = c;
        n = __ldg(s.chunk_start + c);
        end = __ldg(s.chunk_start + c + 1);
        prev_u = -1;
      }
    }
    __syncwarp();
    if (__all_sync(0xffffffffu, done)) break;
    if (!done) {
      const RatingRec rec = ld_rec(s.rec + n);
      const bool first = (rec.u != prev_u);
      bool ok = true;
      if (gl == 0) ok = (ld_relaxed_u32(s.done_j + rec.j) == (unsigned)rec.kj);
      if (gl == 1 && first) ok = (ld_relaxed_u32(s.done_u + rec.u) == (unsigned)rec.ku);
      const unsigned b = __ballot_sync(gmask, ok);
      if ((b & gmask) == gmask) {
        acquire_after_poll(gl == 0 ? s.done_j + rec.j : s.done_u + rec.u, gl == 0 || (gl == 1 && first));
        __syncwarp(gmask);  // the leaders' polls happen-before every lane's loads below
        bool last = (n + 1 == end);
        if (!last) last = (__ldg(&s.rec[n + 1].u) != rec.u);
        acc = __dadd_rn(acc, rating_update<MODEL, LPR, V>(m, rec.u, rec.j, rec.ctx, rec.r, lr, prod, gl, gmask,
                                                          us, first, last));
        __syncwarp(gmask);  // the group's stores happen-before lane 0's fence
        if (gl == 0) {      // release pattern for BOTH counters: one fence, then two relaxed (strong) stores
          fence_release_gpu();
          st_relaxed_u32(s.done_j + rec.j, (unsigned)rec.kj + 1u);
          if (last) st_relaxed_u32(s.done_u + rec.u, (unsigned)rec.ku + 1u);
        }
        prev_u = rec.u;
        n++;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1f: flagged wavefront (the default schedule).
// Ratings are sorted by dependency level as for K1 and statically dealt round-robin to the co-resident
// groups (rating n -> group n mod T), but instead of a grid-wide barrier per level every rating waits
// only for ITS two predecessors: done_u[u] == ku and done_j[j] == kj (completion counters).
// The sorted order is a topological order of the conflict DAG and every group walks its ratings in
// that order, so the earliest unfinished rating is always runnable: deadlock-free given co-residency
// (cooperative launch).  A group that is not ready skips its turn (no blocking wait on another group
// inside a warp).  Levels overlap: the tail of level L runs beside the head of level L+1.
//
// Memory ordering: the producer stores its rows (st.global.cg), __syncwarp, then lane 0 issues ONE
// fence.release.gpu (MEMBAR.ALL.GPU, cumulative over the group's stores) followed by the two relaxed counter
// stores -- the release pattern for both counters.  The consumer polls with ld.relaxed.gpu and, after the branch
// on the polled values, one ld.acquire.gpu of each counter (acquire_after_poll) and __syncwarp, reads the rows with
// ld.global.cg: release pattern on one side, acquire pattern on the other.
// ------------------------------------------------------------------------------------------------
template <int MODEL, int LPR, int V, int THREADS, int MINB, bool WIDE = false, int FIXF = 0>
__global__ void __launch_bounds__(THREADS, MINB)
    sgd_flagged_kernel(DeviceModel m, const RatingRec* __restrict__ recs, int64_t nnz, unsigned* flags,
                       unsigned off_u, unsigned off_j, double lr, double* block_partial
#ifdef CARS_TRACE
                       , unsigned long long* trace, int64_t trace_lo, int64_t trace_n
#endif
    ) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int G = 32 / LPR;
  constexpr int WARPS = THREADS / 32;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane % LPR;
  const int gw = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (gw * LPR));
  const int prod_stride = (FIXF > 0 ? FIXF : m.Fp) + 2;
  double* prod = reinterpret_cast<double*>(smem_raw) + (size_t)(warp * G + gw) * prod_stride;
  unsigned* done_u = flags + off_u;
  unsigned* done_j = flags + off_j;
  constexpr bool kHasCond = (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI);

  // interleave CTAs so that consecutive ratings land on different SMs
  const int64_t T = (int64_t)gridDim.x * WARPS * G;
  int64_t n = ((int64_t)warp * gridDim.x + blockIdx.x) * G + gw;

  double acc = 0.0;
  RatingRec rec, next;
  rec.u = rec.j = rec.ctx = rec.ku = rec.kj = 0; rec.r = 0.0;
  if (n < nnz) rec = ld_rec(recs + n);
  next = rec;
#ifdef CARS_TRACE
  unsigned long long tries = 0;
#endif
  for (;;) {
    const bool active = n < nnz;
    if (!__any_sync(0xffffffffu, active)) break;
    if (active) {
#ifdef CARS_TRACE
      const long long tc0 = clock64();
      tries++;
#endif
      // the lane's condition id (static table) is fetched beside the poll, off the gather's critical path
      int cond = -1;
      if (kHasCond && gl < m.Dmax) cond = __ldg(m.ctx_tab + (int64_t)rec.ctx * m.Dmax + gl);
      bool ok = true;
      if (gl == 0) ok = (ld_relaxed_u32(done_j + rec.j) == (unsigned)rec.kj);
      if (gl == 1) ok = (ld_relaxed_u32(done_u + rec.u) == (unsigned)rec.ku);
      const unsigned b = __ballot_sync(gmask, ok);
      if ((b & gmask) == gmask) {
        acquire_after_poll(gl == 0 ? done_j + rec.j : done_u + rec.u, gl < 2);
        __syncwarp(gmask);
        const int64_t nn = n + T;
        if (nn < nnz) next = ld_rec(recs + nn);  // flies during this rating's gather and arithmetic
#ifdef CARS_TRACE
        const long long tc1 = clock64();
        unsigned long long tg1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg1));
#endif
        UserRegs<V> us;
        Operands<V> o;
        gather_rows<MODEL, LPR, V, false, WIDE, FIXF>(m, rec.u, rec.j, gl, us, o, true);
        gather_scalars<MODEL, LPR, V>(m, rec.u, rec.j, rec.ctx, gl, o, cond);
#ifdef CARS_TRACE
        long long tc2;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(tc2) : "d"(us.p[V - 1].y), "d"(o.q[V - 1].y), "d"(o.cb), "d"(us.p[0].x), "d"(o.q[0].x));
#endif
        acc = __dadd_rn(acc, compute_scatter<MODEL, LPR, V, false, WIDE, FIXF>(m, rec.u, rec.j, rec.ctx, rec.r, lr, prod, gl,
                                                                         gmask, us, o, true));
        __syncwarp(gmask);  // the group's stores happen-before lane 0's release
#ifdef CARS_TRACE
        const long long tc3 = clock64();
#endif
        if (gl == 0) {  // release pattern for BOTH counters: one fence (MEMBAR.ALL.GPU), then two relaxed stores
          fence_release_gpu();
          st_relaxed_u32(done_j + rec.j, (unsigned)rec.kj + 1u);
          st_relaxed_u32(done_u + rec.u, (unsigned)rec.ku + 1u);
        }
#ifdef CARS_TRACE
        if (gl == 0 && trace && n >= trace_lo && n < trace_lo + trace_n) {
          const long long tc4 = clock64();
          unsigned long long tg4;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg4));
          unsigned long long* t = trace + (n - trace_lo) * 8;
          t[0] = (unsigned long long)(tc1 - tc0); t[1] = (unsigned long long)(tc2 - tc1); t[2] = (unsigned long long)(tc3 - tc2);
          t[3] = (unsigned long long)(tc4 - tc3); t[4] = tg1; t[5] = tg4; t[6] = tries; t[7] = (unsigned long long)tc0;
        }
        tries = 0;
#endif
        rec = next;
        n = nn;
      }
    }
  }

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

// K3a: deterministic reduction of the per-chunk loss partials: block b sums a fixed contiguous slice.
__global__ void __launch_bounds__(256) chunk_loss_reduce_kernel(const double* chunk_loss, int64_t n, double* block_partial) {
  __shared__ double sh[256];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t beg = (int64_t)blockIdx.x * per;
  const int64_t end = beg + per < n ? beg + per : n;
  double t = 0.0;
  for (int64_t i = beg + threadIdx.x; i < end; i += 256) t += chunk_loss[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_partial[blockIdx.x] = sh[0];
}

// C1 helpers: the per-epoch exchange of the item block between user-range shards (one GPU each).
//   delta = cur - old   (fused with nothing else: one streaming pass, HBM-bound)
//   cur   = old + scale * sum   (after the caller's all-reduce of delta)
__global__ void __launch_bounds__(256) item_delta_kernel(const double* __restrict__ cur, const double* __restrict__ old,
                                                         double* __restrict__ delta, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    delta[i] = __dsub_rn(cur[i], old[i]);
}
__global__ void __launch_bounds__(256) item_apply_kernel(double* __restrict__ cur, const double* __restrict__ old,
                                                         const double* __restrict__ sum, double scale, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    cur[i] = __dadd_rn(old[i], __dmul_rn(scale, sum[i]));
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
__device__ __forceinline__ double predict_from_dot(const DeviceModel& m, int u, int j, int ctx, double dot);

template <int MODEL>
__device__ __forceinline__ double predict_one(const DeviceModel& m, int u, int j, int ctx) {
  const double* p = m.P + (int64_t)u * m.Fp;
  const double* q = m.Q + (int64_t)j * m.Fp;
  double dot = 0.0;
  for (int f = 0; f < m.F; f++) dot = __dadd_rn(dot, __dmul_rn(p[f], q[f]));
  return predict_from_dot<MODEL>(m, u, j, ctx, dot);
}

// everything of the model's predict(u, j, c) after DenseMatrix.rowMult, in the reference's order of additions
template <int MODEL>
__device__ __forceinline__ double predict_from_dot(const DeviceModel& m, int u, int j, int ctx, double dot) {
  double pred;
  if (MODEL == M_PMF) pred = dot;
  if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C)
    pred = __dadd_rn(__dadd_rn(__dadd_rn(m.global_mean, m.user_bias[u]), m.item_bias[j]), dot);
  if (MODEL == M_CAMF_CI) pred = __dadd_rn(__dadd_rn(m.global_mean, m.user_bias[u]), dot);
  if (MODEL == M_CAMF_CU) pred = __dadd_rn(__dadd_rn(m.global_mean, m.item_bias[j]), dot);
  if (MODEL == M_CAMF_CUCI) pred = __dadd_rn(m.global_mean, dot);  // CAMF_CUCI.java:69
  if (MODEL == M_SVDPP) {  // SVDPlusPlus.java:139-147: BiasedMF's prediction + sum_k rowMult(Y, k, Q, j) / sqrt(|items(u)|)
    pred = __dadd_rn(__dadd_rn(__dadd_rn(m.global_mean, m.user_bias[u]), m.item_bias[j]), dot);
    const int a = m.ui_ptr[u], b = m.ui_ptr[u + 1];
    const double w = __dsqrt_rn((double)(b - a));
    const double* qrow = m.Q + (int64_t)j * m.Fp;
    for (int i = a; i < b; i++) {
      const double* yrow = m.Y + (int64_t)m.ui_items[i] * m.Fp;
      double yq = 0.0;
      for (int f = 0; f < m.F; f++) yq = __dadd_rn(yq, __dmul_rn(yrow[f], qrow[f]));
      pred = __dadd_rn(pred, __ddiv_rn(yq, w));
    }
  }
  if (MODEL == M_CAMF_LCS) {  // CAMF_LCS.java:44-62: pred = pred * rowMult(cfMatrix_LCS, cond, cfMatrix_LCS, empty)
    pred = dot;
    for (int d = 0; d < m.Dmax; d++) {
      const int cond = m.ctx_tab[(int64_t)ctx * m.Dmax + d];
      if (cond < 0) continue;
      const double* c1 = m.cf_lcs + (int64_t)cond * m.numF;
      const double* c2 = m.cf_lcs + (int64_t)m.empty_cond[d] * m.numF;
      double sim = 0.0;
      for (int f = 0; f < m.numF; f++) sim = __dadd_rn(sim, __dmul_rn(c1[f], c2[f]));
      pred = __dmul_rn(pred, sim);
    }
  }
  if (MODEL == M_CAMF_MCS) {  // CAMF_MCS.java:52-69: pred = pred * (1 - sqrt(sum_d (pos(cond_d) - pos(empty_d))^2))
    double dist = 0.0;
    for (int d = 0; d < m.Dmax; d++) {
      const int cond = m.ctx_tab[(int64_t)ctx * m.Dmax + d];
      if (cond < 0) continue;
      const double diff = __dsub_rn(m.c_mcs[cond], m.c_mcs[m.empty_cond[d]]);
      dist = __dadd_rn(dist, __dmul_rn(diff, diff));
    }
    pred = __dmul_rn(dot, __dsub_rn(1.0, __dsqrt_rn(dist)));
  }
  if (MODEL == M_CAMF_ICS) {  // CAMF_ICS.java:52-58: pred = pred * ccMatrix_ICS.get(conditions.get(i), EmptyContextConditions.get(i))
    pred = dot;
    for (int d = 0; d < m.Dmax; d++) {
      const int cond = m.ctx_tab[(int64_t)ctx * m.Dmax + d];
      if (cond < 0) continue;
      const int e2 = m.empty_cond[d];
      const int hi = cond >= e2 ? cond : e2, lo = cond >= e2 ? e2 : cond;
      pred = __dmul_rn(pred, m.cc_sim[(int64_t)hi * m.C + lo]);
    }
  }
  if (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI) {
    for (int d = 0; d < m.Dmax; d++) {
      const int cond = m.ctx_tab[(int64_t)ctx * m.Dmax + d];
      if (cond < 0) continue;
      double b = MODEL == M_CAMF_C    ? m.cond_bias[cond]
                 : (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) ? m.ic_bias[(int64_t)j * m.C + cond]
                                      : m.uc_bias[(int64_t)u * m.C + cond];
      if (MODEL == M_CAMF_CUCI) b = __dadd_rn(b, m.uc_bias[(int64_t)u * m.C + cond]);  // :71 (ic + uc) first
      pred = __dadd_rn(pred, b);
    }
  }
  return pred;
}

// K5 (the kernel cars_predict launches): a group of 8 lanes per query.  The two factor rows are read with coalesced
// 16-byte loads (a group covers 128 contiguous bytes per instruction) instead of one thread striding a whole row; the
// products go through the group's shared-memory scratch and are summed in f = 0..F-1 order, so the value stays
// bit-identical to DenseMatrix.rowMult.  Algorithmic bytes per query: 2 * F * 8 (rows) + 12 (ids) + 8 (out) + biases.
template <int MODEL>
__global__ void __launch_bounds__(256) predict_group_kernel(DeviceModel m, int64_t n, const int32_t* __restrict__ u,
                                                            const int32_t* __restrict__ j, const int32_t* __restrict__ ctx, int bound,
                                                            double min_rate, double max_rate, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LPR = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, gw = lane / LPR;
  const unsigned gmask = ((1u << LPR) - 1u) << (gw * LPR);
  double* prod = reinterpret_cast<double*>(smem_raw) + (size_t)(warp * 4 + gw) * (m.Fp + 2);
  const int64_t groups = (int64_t)gridDim.x * 32;
  const int nchunks = m.Fp / 2;
  for (int64_t i0 = (int64_t)blockIdx.x * 32 + warp * 4; i0 < n; i0 += groups) {  // the 4 groups of a warp stay together
    const int64_t i = i0 + gw;
    const bool active = i < n;
    int uu = 0, jj = 0, cc = 0;
    if (active) {
      uu = __ldg(u + i);
      jj = __ldg(j + i);
      cc = ctx ? __ldg(ctx + i) : 0;
      const double* prow = m.P + (int64_t)uu * m.Fp;
      const double* qrow = m.Q + (int64_t)jj * m.Fp;
      for (int c = gl; c < nchunks; c += LPR) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(prow) + c);
        const double2 b = __ldg(reinterpret_cast<const double2*>(qrow) + c);
        *reinterpret_cast<double2*>(prod + 2 * c) = make_double2(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
      }
    }
    __syncwarp();
    if (active && gl == 0) {
      double dot = 0.0;
      for (int f = 0; f < m.F; f++) dot = __dadd_rn(dot, prod[f]);  // (the pad of an odd F is never added)
      double pred = predict_from_dot<MODEL>(m, uu, jj, cc, dot);
      if (bound) {
        if (pred > max_rate) pred = max_rate;
        if (pred < min_rate) pred = min_rate;
      }
      out[i] = pred;
    }
    __syncwarp();
    (void)gmask;
  }
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
