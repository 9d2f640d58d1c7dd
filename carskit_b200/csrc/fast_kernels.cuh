// fast_kernels.cuh -- K2: the FAST (hogwild) schedule of the CARSKit SGD hot path on sm_100a.
//
// Same per-rating arithmetic as sgd_kernels.cuh (CAMF_CI.java:80-121 and siblings), NOT the same order of
// execution: EXACT mode keeps, for every user and every item, that row's ratings in reference order, so its
// parallelism is nnz / max item degree -- on skewed data (every real data set under context-aware_data_sets/)
// the dependency DAG is a few ratings wide and a GPU has nothing to do.  FAST drops the item-side ordering:
//
//   * the ratings are stably sorted by user (fast_schedule.cuh) and cut into chunks at user boundaries; a group
//     of LPR lanes takes chunks from a global counter and walks each one in order, so ONE group runs all the
//     ratings of a user: P[u], userBias[u] stay in registers across the user's run and ucBias[u][.] is read and
//     written by one thread per cell -- the user side has no race at all;
//   * every item-side cell (Q[j], itemBias[j], icBias[j][.], condBias[.]) is read with a plain L2 load -- the
//     operands of rating n + 1 are requested before the arithmetic of rating n starts (software pipeline in
//     registers; nothing to poll, so the load is always useful) -- and updated with red.global.add.f64: the
//     step  lr * (e * p - reg * q)  computed from the possibly stale read is ADDED to whatever the cell holds;
//   * a cell that many in-flight ratings share would receive the sum of many stale gradients (an effective
//     learning rate of lr * concurrency: diverges for a Zipf-head item or for CAMF_C's condBias).  Such rows get
//     their step scaled by  min(1, max_conc / expected concurrency)  (item_scale / cond_scale, computed from the
//     row's degree at create time; nullptr when no row needs it).
//
// With no two in-flight ratings sharing an item the result equals the serial loop up to the order of the dot
// product (a tree here, sequential in Java) -- tests/test_fast_gpu.py checks exactly that.
#pragma once
#include "sgd_kernels.cuh"

namespace cars {

struct FastStream {
  const RatingRec* rec;        // [nnz] stably sorted by user
  const int64_t* chunk_start;  // [num_chunks + 1], every boundary is a user boundary
  unsigned* counter;           // next chunk to hand out (zeroed per epoch)
  uint32_t num_chunks;
  const double* item_scale;    // [num_items] step damping of the item-side cells, or nullptr (all 1)
  const double* cond_scale;    // [C] CAMF_C: step damping of condBias, or nullptr
  // Hot rows (a Zipf head): same-address reductions serialise in the L2 atomic unit (~35 ns per 64-factor row,
  // profiles/r2), so an item that owns 8 % of the ratings would bound the epoch on its own.  The steps of the
  // num_hot most popular items are summed per CTA in shared memory and flushed to Q[j] / itemBias[j] once every
  // hot_flush updates of the CTA (and at the end of the kernel): hot_flush times fewer global reductions on those rows.
  const signed char* hot_slot;  // [num_items] slot of a hot item, -1 otherwise; nullptr = no hot rows
  const int32_t* hot_items;     // [num_hot] item id of every slot
  int num_hot, hot_flush;
  int bulk_offset;              // BULK kernels: byte offset of the per-group delta-row slots in dynamic shared memory
};

// TMA add-reduce of one staged row: global[dst .. dst + bytes) += shared[src ..) as fp64, asynchronously (UBLKRED.G.S.ADD.F64)
__device__ __forceinline__ void bulk_reduce_add_f64(double* dst, const double* src_shared, int bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(src_shared);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(sa), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One hot row in shared memory: [Fp factors | itemBias | pad | C icBias cells (CAMF_CI / CAMF_CUCI only)]
template <int MODEL>
__host__ __device__ __forceinline__ int hot_row_stride(int Fp, int C) {
  return Fp + 2 + ((MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) ? C : 0);
}

__device__ __forceinline__ double atomic_exch_shared_f64(double* p, double v) {
  return __longlong_as_double((long long)atomicExch(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__double_as_longlong(v)));
}

__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Item-side operands of one rating, requested one rating ahead.
template <int V>
struct FastOps {
  double2 q[V];
  double bj;       // itemBias[j]
  double cb;       // this lane's item-side condition cell: icBias[j][cond] (CI, CUCI) or condBias[cond] (C)
  double* cb_ptr;  // nullptr when the lane owns none
  double scale;    // item_scale[j]
  double cscale;   // cond_scale[cond] (CAMF_C)
  int cond;        // this lane's condition id, -1 = none
  int slot;        // hot-row slot of item j, -1 = none
};

template <int MODEL, int LPR, int V, bool WIDE, int FIXF>
__device__ __forceinline__ void fast_load(const DeviceModel& m, const FastStream& s, const RatingRec& rec, int gl,
                                          FastOps<V>& o) {
  constexpr bool kItemBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CU);
  constexpr bool kHasCond = (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI);
  const int Fp = FIXF > 0 ? FIXF : m.Fp;
  const double* qrow = m.Q + (int64_t)rec.j * Fp;
  if (WIDE) {
#pragma unroll
    for (int w = 0; w < V / 2; w++) {
      const int c = chunk_of<LPR, true>(gl, 2 * w);
      if (2 * c < Fp) ld_cg_f64x4(qrow + 2 * c, o.q[2 * w], o.q[2 * w + 1]);
      else o.q[2 * w] = o.q[2 * w + 1] = make_double2(0.0, 0.0);
    }
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = gl + v * LPR;
      if (2 * c < Fp) o.q[v] = ld_cg_f64x2(qrow + 2 * c);
      else o.q[v] = make_double2(0.0, 0.0);
    }
  }
  o.bj = 0.0;
  if (kItemBias) o.bj = ld_cg_f64(m.item_bias + rec.j);
  o.scale = 1.0;
  if (s.item_scale) o.scale = __ldg(s.item_scale + rec.j);
  o.slot = -1;
  if (s.hot_slot) o.slot = __ldg(s.hot_slot + rec.j);
  o.cb = 0.0;
  o.cb_ptr = nullptr;
  o.cscale = 1.0;
  o.cond = -1;
  if (kHasCond && gl < m.Dmax) {
    const int cond = __ldg(m.ctx_tab + (int64_t)rec.ctx * m.Dmax + gl);
    o.cond = cond;
    if (cond >= 0) {
      if (MODEL == M_CAMF_C) {
        o.cb_ptr = m.cond_bias + cond;
        if (s.cond_scale) o.cscale = __ldg(s.cond_scale + cond);
      }
      if (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) o.cb_ptr = m.ic_bias + (int64_t)rec.j * m.C + cond;
      if (o.cb_ptr) o.cb = ld_cg_f64(o.cb_ptr);
    }
  }
}

// One rating: arithmetic + user-side stores + item-side reductions.  Returns this lane's loss contribution.
// BULK: the Q-row step is staged in the group's shared-memory slot `stage` and added to Q[j] by ONE TMA reduce
// (cp.reduce.async.bulk ... add.f64) instead of F scalar red.global.add.f64 -- the 64 REDG per rating are what bounds the
// scalar kernel (LSU: 1.3 cycles per lane, profiles/r2/ncu_summary_fast_uniform_100M.txt).
template <int MODEL, int LPR, int V, bool WIDE, int FIXF, bool BULK = false>
__device__ __forceinline__ double fast_update(const DeviceModel& m, const FastStream& s, const RatingRec& rec, double lr,
                                              int gl, unsigned gmask, UserRegs<V>& us, const FastOps<V>& o, double* sh_hot,
                                              double* stage = nullptr) {
  constexpr bool kUserBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CI);
  constexpr bool kItemBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CU);
  constexpr bool kHasCond = (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI);
  constexpr bool kUserCond = (MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI);  // ucBias[u][cond]: owned by this group
  const int Fp = FIXF > 0 ? FIXF : m.Fp;
  const int Dmax = m.Dmax;
  const int u = rec.u, j = rec.j;
  const int hstride = hot_row_stride<MODEL>(Fp, m.C);
  double* srow = sh_hot + (o.slot >= 0 ? o.slot : 0) * hstride;  // (shared address space: ATOMS, not generic ATOM)

  // the user-side condition cell is read NOW (after the previous rating of this user stored it; same thread)
  double ucb = 0.0;
  double* ucb_ptr = nullptr;
  if (kUserCond && o.cond >= 0) {
    ucb_ptr = m.uc_bias + (int64_t)u * m.C + o.cond;
    ucb = ld_cg_f64(ucb_ptr);
  }

  // ---- dot product: per-lane partial, then a butterfly over the group's lanes ---------------------------------
  double d0 = 0.0, d1 = 0.0;
#pragma unroll
  for (int v = 0; v < V; v++) {
    d0 = fma(us.p[v].x, o.q[v].x, d0);  // padded slots hold zeros
    d1 = fma(us.p[v].y, o.q[v].y, d1);
  }
  double dot = d0 + d1;
#pragma unroll
  for (int w = LPR / 2; w > 0; w >>= 1) {
    const int lo = __shfl_xor_sync(gmask, __double2loint(dot), w, LPR);
    const int hi = __shfl_xor_sync(gmask, __double2hiint(dot), w, LPR);
    dot += __hiloint2double(hi, lo);
  }

  // ---- predict (the model's own order of additions) ------------------------------------------------------------
  const double bu = kUserBias ? us.bu : 0.0;
  const double bj = o.bj;
  double pred;
  if (MODEL == M_PMF) pred = dot;
  if (MODEL == M_BIASEDMF || MODEL == M_CAMF_C) pred = __dadd_rn(__dadd_rn(__dadd_rn(m.global_mean, bu), bj), dot);
  if (MODEL == M_CAMF_CI) pred = __dadd_rn(__dadd_rn(m.global_mean, bu), dot);
  if (MODEL == M_CAMF_CU) pred = __dadd_rn(__dadd_rn(m.global_mean, bj), dot);
  if (MODEL == M_CAMF_CUCI) pred = __dadd_rn(m.global_mean, dot);
  double lane_loss = 0.0;
  if (kHasCond) {
    const int D1 = Dmax < LPR ? Dmax : LPR;
    const double cbs = MODEL == M_CAMF_CU ? ucb : MODEL == M_CAMF_CUCI ? __dadd_rn(o.cb, ucb) : o.cb;
    for (int d = 0; d < D1; d++) pred = __dadd_rn(pred, shfl_f64(gmask, cbs, d, LPR));
    for (int d = LPR; d < Dmax; d++) {  // more context dimensions than lanes in a group: every lane reads them
      const int cond = __ldg(m.ctx_tab + (int64_t)rec.ctx * Dmax + d);
      if (cond < 0) continue;
      double b = 0.0;
      if (MODEL == M_CAMF_C) b = ld_cg_f64(m.cond_bias + cond);
      if (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) b = ld_cg_f64(m.ic_bias + (int64_t)j * m.C + cond);
      if (MODEL == M_CAMF_CU) b = ld_cg_f64(m.uc_bias + (int64_t)u * m.C + cond);
      if (MODEL == M_CAMF_CUCI) b = __dadd_rn(b, ld_cg_f64(m.uc_bias + (int64_t)u * m.C + cond));
      pred = __dadd_rn(pred, b);
    }
  }
  const double e = __dsub_rn(rec.r, pred);
  const double lrj = __dmul_rn(lr, o.scale);  // item-side step (damped for rows many in-flight ratings share)

  // ---- bias steps ----------------------------------------------------------------------------------------------
  if (kUserBias) us.bu = __dadd_rn(bu, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_b, bu))));
  if (gl == 0) {
    lane_loss = __dmul_rn(e, e);
    if (kUserBias) lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bu), bu));
    if (kItemBias) {
      const double step = __dmul_rn(lrj, __dsub_rn(e, __dmul_rn(m.reg_b, bj)));
      if (o.slot >= 0) atomicAdd(srow + Fp, step);
      else red_add_f64(m.item_bias + j, step);
      lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bj), bj));
    }
  }
  if (kHasCond) {
    if (o.cb_ptr != nullptr) {  // item-side cell (or condBias): reduction
      const double lrc = MODEL == M_CAMF_C ? __dmul_rn(lr, o.cscale) : lrj;
      const double step = __dmul_rn(lrc, __dsub_rn(e, __dmul_rn(m.reg_c, o.cb)));
      if (MODEL != M_CAMF_C && o.slot >= 0) atomicAdd(srow + Fp + 2 + o.cond, step);  // a hot item's icBias cell
      else red_add_f64(o.cb_ptr, step);
      if (MODEL == M_CAMF_C) lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_b, o.cb));  // CAMF_C.java:115
      else lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_c, __dmul_rn(o.cb, o.cb)));
    }
    if (kUserCond && ucb_ptr != nullptr) {  // user-side cell: this thread is its only writer
      st_cg_f64(ucb_ptr, __dadd_rn(ucb, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_c, ucb)))));
      lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_c, __dmul_rn(ucb, ucb)));
    }
    if (gl == 0) {
      for (int d = LPR; d < Dmax; d++) {
        const int cond = __ldg(m.ctx_tab + (int64_t)rec.ctx * Dmax + d);
        if (cond < 0) continue;
        if (MODEL == M_CAMF_C || MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) {
          double* bp = MODEL == M_CAMF_C ? m.cond_bias + cond : m.ic_bias + (int64_t)j * m.C + cond;
          const double b = ld_cg_f64(bp);
          const double lrc = (MODEL == M_CAMF_C) ? __dmul_rn(lr, s.cond_scale ? __ldg(s.cond_scale + cond) : 1.0) : lrj;
          const double step = __dmul_rn(lrc, __dsub_rn(e, __dmul_rn(m.reg_c, b)));
          if (MODEL != M_CAMF_C && o.slot >= 0) atomicAdd(srow + Fp + 2 + cond, step);
          else red_add_f64(bp, step);
          lane_loss = __dadd_rn(lane_loss, MODEL == M_CAMF_C ? __dmul_rn(m.reg_b, b) : __dmul_rn(m.reg_c, __dmul_rn(b, b)));
        }
        if (kUserCond) {
          double* bp = m.uc_bias + (int64_t)u * m.C + cond;
          const double b = ld_cg_f64(bp);
          st_cg_f64(bp, __dadd_rn(b, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_c, b)))));
          lane_loss = __dadd_rn(lane_loss, __dmul_rn(m.reg_c, __dmul_rn(b, b)));
        }
      }
    }
    // lane 0 wrote user-side cells that every lane of the group reads for the user's next rating
    if (kUserCond && Dmax > LPR) __syncwarp(gmask);
  }

  // ---- factor steps (both from the values read) -----------------------------------------------------------------
  // a hot item's steps go to the CTA's shared-memory accumulator row, everything else straight to L2
  double* qrow = m.Q + (int64_t)j * Fp;
  double sp = 0.0, sq = 0.0;
  if (BULK && o.slot < 0) {  // the slot was last read by the bulk reduce issued two ratings ago: make sure it is done
    if (gl == 0) bulk_wait_read_1();
    __syncwarp(gmask);
  }
#pragma unroll
  for (int v = 0; v < V; v++) {
    const int c = chunk_of<LPR, WIDE>(gl, v);
    if (2 * c < Fp) {
      const double2 po = us.p[v], qo = o.q[v];
      us.p[v].x = __dadd_rn(po.x, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.x), __dmul_rn(m.reg_u, po.x))));
      us.p[v].y = __dadd_rn(po.y, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo.y), __dmul_rn(m.reg_u, po.y))));
      const double dx = __dmul_rn(lrj, __dsub_rn(__dmul_rn(e, po.x), __dmul_rn(m.reg_i, qo.x)));
      const double dy = __dmul_rn(lrj, __dsub_rn(__dmul_rn(e, po.y), __dmul_rn(m.reg_i, qo.y)));
      if (o.slot >= 0) {
        atomicAdd(srow + 2 * c, dx);
        atomicAdd(srow + 2 * c + 1, dy);
      } else if (BULK) {
        *reinterpret_cast<double2*>(stage + 2 * c) = make_double2(dx, dy);
      } else {
        red_add_f64(qrow + 2 * c, dx);
        red_add_f64(qrow + 2 * c + 1, dy);
      }
      sp = fma(po.x, po.x, sp);
      sq = fma(qo.x, qo.x, sq);
      sp = fma(po.y, po.y, sp);
      sq = fma(qo.y, qo.y, sq);
    }
  }
  if (BULK && o.slot < 0) {
    fence_proxy_async_shared();  // this lane's shared-memory writes -> visible to the async proxy
    __syncwarp(gmask);
    if (gl == 0) bulk_reduce_add_f64(qrow, stage, Fp * 8);
  }
  if (s.hot_slot) {  // every hot_flush-th update of a slot by this CTA moves the accumulated row to L2
    bool flush = false;
    if (o.slot >= 0) {
      __syncwarp(gmask);
      unsigned* cnt = reinterpret_cast<unsigned*>(sh_hot + s.num_hot * hstride);
      unsigned c = 0;
      if (gl == 0) c = atomicAdd(cnt + o.slot, 1u) + 1u;
      c = __shfl_sync(gmask, c, 0, LPR);
      flush = (c % (unsigned)s.hot_flush) == 0u;
    }
    if (flush) {
      for (int f = gl; f < Fp; f += LPR) {
        const double v = atomic_exch_shared_f64(srow + f, 0.0);
        if (v != 0.0) red_add_f64(qrow + f, v);
      }
      if (kItemBias && gl == 0) {
        const double v = atomic_exch_shared_f64(srow + Fp, 0.0);
        if (v != 0.0) red_add_f64(m.item_bias + j, v);
      }
      if (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) {
        for (int c = gl; c < m.C; c += LPR) {
          const double v = atomic_exch_shared_f64(srow + Fp + 2 + c, 0.0);
          if (v != 0.0) red_add_f64(m.ic_bias + (int64_t)j * m.C + c, v);
        }
      }
    }
  }
  return __dadd_rn(lane_loss, fma(m.reg_u, sp, __dmul_rn(m.reg_i, sq)));
}

// The user's row: read when a user's run starts, written back when it ends (plain L2 accesses: no other group
// touches P[u] / userBias[u] during the epoch).
template <int MODEL, int LPR, int V, bool WIDE, int FIXF>
__device__ __forceinline__ void fast_load_user(const DeviceModel& m, int u, int gl, UserRegs<V>& us) {
  constexpr bool kUserBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CI);
  const int Fp = FIXF > 0 ? FIXF : m.Fp;
  const double* prow = m.P + (int64_t)u * Fp;
  if (WIDE) {
#pragma unroll
    for (int w = 0; w < V / 2; w++) {
      const int c = chunk_of<LPR, true>(gl, 2 * w);
      if (2 * c < Fp) ld_cg_f64x4(prow + 2 * c, us.p[2 * w], us.p[2 * w + 1]);
      else us.p[2 * w] = us.p[2 * w + 1] = make_double2(0.0, 0.0);
    }
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = gl + v * LPR;
      if (2 * c < Fp) us.p[v] = ld_cg_f64x2(prow + 2 * c);
      else us.p[v] = make_double2(0.0, 0.0);
    }
  }
  us.bu = 0.0;
  if (kUserBias) us.bu = ld_cg_f64(m.user_bias + u);
}

template <int MODEL, int LPR, int V, bool WIDE, int FIXF>
__device__ __forceinline__ void fast_store_user(const DeviceModel& m, int u, int gl, const UserRegs<V>& us) {
  constexpr bool kUserBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CI);
  const int Fp = FIXF > 0 ? FIXF : m.Fp;
  double* prow = m.P + (int64_t)u * Fp;
  if (WIDE) {
#pragma unroll
    for (int w = 0; w < V / 2; w++) {
      const int c = chunk_of<LPR, true>(gl, 2 * w);
      if (2 * c < Fp) st_cg_f64x4(prow + 2 * c, us.p[2 * w], us.p[2 * w + 1]);
    }
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = gl + v * LPR;
      if (2 * c < Fp) st_cg_f64x2(prow + 2 * c, us.p[v]);
    }
  }
  if (kUserBias && gl == 0) st_cg_f64(m.user_bias + u, us.bu);
}

// K2: persistent grid (any size: chunks are only held by running groups, so no co-residency is needed).
template <int MODEL, int LPR, int V, int THREADS, int MINB, bool WIDE = false, int FIXF = 0, bool BULK = false>
__global__ void __launch_bounds__(THREADS, MINB) sgd_fast_kernel(DeviceModel m, FastStream s, double lr, double* block_partial) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sh_hot = reinterpret_cast<double*>(smem_raw);  // [num_hot x (Fp + 2)] accumulators, then num_hot counters
  constexpr int WARPS = THREADS / 32;
  const int hstride = hot_row_stride<MODEL>(FIXF > 0 ? FIXF : m.Fp, m.C);
  const int hot_words = s.hot_slot ? s.num_hot * hstride : 0;
  if (s.hot_slot) {
    for (int i = threadIdx.x; i < hot_words + (s.num_hot + 1) / 2; i += THREADS) sh_hot[i] = 0.0;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane % LPR;
  const int gw = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (gw * LPR));

  int64_t n = 0, end = 0;
  bool done = false;
  int prev_u = -1;
  double acc = 0.0;
  UserRegs<V> us;
#pragma unroll
  for (int v = 0; v < V; v++) us.p[v] = make_double2(0.0, 0.0);
  us.bu = 0.0;
  int stage_sel = 0;           // BULK: which of the group's two staging slots the next bulk reduce reads
  RatingRec rec, recn, recn2;  // current, next (its operands are requested this turn), the one after (record in flight)
  rec.u = rec.j = rec.ctx = rec.ku = rec.kj = rec.pad = 0; rec.r = 0.0;
  recn = rec;
  recn2 = rec;
  FastOps<V> cur, nxt;

  for (;;) {
    if (!done && n == end) {  // take the next non-empty chunk
      for (;;) {
        unsigned c = 0;
        if (gl == 0) c = atomicAdd(s.counter, 1u);
        c = __shfl_sync(gmask, c, 0, LPR);
        if (c >= s.num_chunks) {
          done = true;
          break;
        }
        n = __ldg(s.chunk_start + c);
        end = __ldg(s.chunk_start + c + 1);
        if (n < end) break;
      }
      if (!done) {
        rec = ld_rec(s.rec + n);
        if (n + 1 < end) recn = ld_rec(s.rec + n + 1);
        fast_load<MODEL, LPR, V, WIDE, FIXF>(m, s, rec, gl, cur);
        prev_u = -1;
      }
    }
    if (__all_sync(0xffffffffu, done)) break;
    if (!done) {
      const bool has_next = n + 1 < end;
      // the next rating's item-side operands fly during this rating's arithmetic; ITS record arrived a turn ago
      // (records are requested two turns ahead, so that the row gather never waits for the record that addresses it)
      if (has_next) fast_load<MODEL, LPR, V, WIDE, FIXF>(m, s, recn, gl, nxt);
      if (n + 2 < end) recn2 = ld_rec(s.rec + n + 2);
      if (rec.u != prev_u) fast_load_user<MODEL, LPR, V, WIDE, FIXF>(m, rec.u, gl, us);
      double* stage = nullptr;
      if (BULK) {  // two slots per group, used alternately: the TMA may still be reading the previous rating's row
        const int Fp_ = FIXF > 0 ? FIXF : m.Fp;
        stage = reinterpret_cast<double*>(smem_raw + s.bulk_offset) + ((size_t)((warp * (32 / LPR) + gw) * 2 + stage_sel)) * Fp_;
        if (cur.slot < 0) stage_sel ^= 1;  // this rating issues a bulk reduce from `stage`; the next one uses the other slot
      }
      acc = __dadd_rn(acc, fast_update<MODEL, LPR, V, WIDE, FIXF, BULK>(m, s, rec, lr, gl, gmask, us, cur, sh_hot, stage));
      if (!has_next || recn.u != rec.u) fast_store_user<MODEL, LPR, V, WIDE, FIXF>(m, rec.u, gl, us);
      prev_u = rec.u;
      if (has_next) {
        rec = recn;
        cur = nxt;
        recn = recn2;
      }
      n++;
    }
  }

  if (BULK && gl == 0) bulk_wait_all();  // the staged rows must have been consumed before the CTA's shared memory goes away
  acc = warp_sum_f64(acc);
  __shared__ double warp_sum[WARPS];
  if (lane == 0) warp_sum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < WARPS; w++) t += warp_sum[w];
    block_partial[blockIdx.x] = t;
  }
  if (s.hot_slot) {  // what the CTA still holds for the hot rows (every warp is past its last update: barrier above)
    constexpr bool kItemBias = (MODEL == M_BIASEDMF || MODEL == M_CAMF_C || MODEL == M_CAMF_CU);
    const int Fp = FIXF > 0 ? FIXF : m.Fp;
    for (int i = threadIdx.x; i < hot_words; i += THREADS) {
      const double v = sh_hot[i];
      if (v == 0.0) continue;
      const int slot = i / hstride, f = i % hstride;
      const int j = __ldg(s.hot_items + slot);
      if (f < Fp) red_add_f64(m.Q + (int64_t)j * Fp + f, v);
      else if (kItemBias && f == Fp) red_add_f64(m.item_bias + j, v);
      else if ((MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI) && f >= Fp + 2) red_add_f64(m.ic_bias + (int64_t)j * m.C + (f - Fp - 2), v);
    }
  }
}

}  // namespace cars
