// tagged_kernels.cuh -- K1t: the flagged wavefront with the flag IN the data ("LL128" rows).
//
// K1f (sgd_kernels.cuh) hands a row from one rating to the next through a completion counter: the producer stores the
// row, FENCES (MEMBAR.ALL.GPU: wait until L2 has acknowledged every store), then stores the counter; the consumer polls
// the counter (one L2 round trip), then gathers the row (another one).  A developer trace put fence + poll at 3.6 of
// the 8.2 us of a turn, and stall_membar at 22.6 % of the kernel's samples (profiles/r1).
//
// Here every 128-byte line of a row carries its own version: 15 doubles of payload + one 64-bit tag, written by ONE
// warp-wide store instruction (8 lanes x 16 bytes -- the store shape of NCCL's LL128 protocol).  The consumer simply
// loads the row and checks that every line carries the tag it expects (ku for P[u]: the number of ratings of that user
// already applied; kj for Q[j]); if one does not, the row is not ready (or only partly visible) and the group skips its
// turn.  No fence, no counter, no separate poll: the load IS the poll, and a row is handed over in one L2 round trip.
// The last rating of a row in the epoch writes tag 0 again, so the next epoch starts from "nothing applied".
//
// What this rests on: a 128-byte line stored by one warp instruction is never observed half old / half new by a 128-byte
// load.  The PTX memory model does not promise it; scripts/litmus/line_atomicity.cu hammers exactly this pattern on the
// B200 (5.6e8 overlapping loads, L2-resident and DRAM-streaming, 8 x 16 B and 4 x 32 B): no torn line
// (profiles/r2/litmus_line_atomicity.txt).  tuning "tagged=0" selects K1f (counters + fence + acquire) instead.
//
// Row layout (TaggedLayout): payload = [F factors | user / item bias if the model has one | the row's C condition-bias
// cells if the model has them (ucBias in P rows: CAMF_CU, CAMF_CUCI; icBias in Q rows: CAMF_CI, CAMF_CUCI)]; payload
// element x lives in line x / 15, slot x % 15.  Everything a rating reads or writes of a user (item) is in that user's
// (item's) row, so one tag per line orders all of it.  Costs bytes: config 3 moves 5 + 7 lines = 1 536 B per rating and
// direction instead of 1 160 B; the kernel it replaces left DRAM 36 % busy.
#pragma once
#include "sgd_kernels.cuh"

namespace cars {

constexpr int kTagPayload = 15;  // doubles of payload per 128-byte line (slot 15 = the tag)

struct TaggedLayout {
  int F = 0, C = 0;
  bool user_bias = false, item_bias = false, uc = false, ic = false;
  __host__ __device__ int p_payload() const { return F + (user_bias ? 1 : 0) + (uc ? C : 0); }
  __host__ __device__ int q_payload() const { return F + (item_bias ? 1 : 0) + (ic ? C : 0); }
  __host__ __device__ int p_lines() const { return (p_payload() + kTagPayload - 1) / kTagPayload; }
  __host__ __device__ int q_lines() const { return (q_payload() + kTagPayload - 1) / kTagPayload; }
};

inline TaggedLayout tagged_layout(int model, int F, int C) {
  TaggedLayout t;
  t.F = F; t.C = C;
  t.user_bias = (model == M_BIASEDMF || model == M_CAMF_CI);
  t.item_bias = (model == M_BIASEDMF || model == M_CAMF_CU);
  t.uc = (model == M_CAMF_CU || model == M_CAMF_CUCI);
  t.ic = (model == M_CAMF_CI || model == M_CAMF_CUCI);
  return t;
}

struct TaggedModel {
  double* Pt;  // [num_users x lp x 16]
  double* Qt;  // [num_items x lq x 16]
  int lp, lq;
};

// standard layout -> tagged rows (tags 0) and back.  which = 0: user side (P, userBias, ucBias), 1: item side.
__global__ void __launch_bounds__(256) tagged_pack_kernel(DeviceModel m, TaggedLayout t, TaggedModel tm, int which, int64_t rows) {
  const int lines = which == 0 ? tm.lp : tm.lq;
  const int64_t total = rows * lines * 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double* dst = which == 0 ? tm.Pt : tm.Qt;
  const double* fac = which == 0 ? m.P : m.Q;
  const double* bias = which == 0 ? m.user_bias : m.item_bias;
  const double* cells = which == 0 ? m.uc_bias : m.ic_bias;
  const bool hb = which == 0 ? t.user_bias : t.item_bias;
  const bool hc = which == 0 ? t.uc : t.ic;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / (lines * 16);
    const int w = (int)(i % (lines * 16)), line = w / 16, slot = w % 16;
    double v = 0.0;  // tag slot: bit pattern 0 == "nothing applied yet"
    if (slot < kTagPayload) {
      const int x = line * kTagPayload + slot;
      if (x < t.F) v = fac[row * m.Fp + x];
      else if (hb && x == t.F) v = bias[row];
      else if (hc && x - t.F - (hb ? 1 : 0) < t.C && x - t.F - (hb ? 1 : 0) >= 0) v = cells[row * t.C + (x - t.F - (hb ? 1 : 0))];
    }
    dst[i] = v;
  }
}

__global__ void __launch_bounds__(256) tagged_unpack_kernel(DeviceModel m, TaggedLayout t, TaggedModel tm, int which, int64_t rows) {
  const int lines = which == 0 ? tm.lp : tm.lq;
  const int64_t total = rows * lines * 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double* src = which == 0 ? tm.Pt : tm.Qt;
  double* fac = which == 0 ? m.P : m.Q;
  double* bias = which == 0 ? m.user_bias : m.item_bias;
  double* cells = which == 0 ? m.uc_bias : m.ic_bias;
  const bool hb = which == 0 ? t.user_bias : t.item_bias;
  const bool hc = which == 0 ? t.uc : t.ic;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / (lines * 16);
    const int w = (int)(i % (lines * 16)), line = w / 16, slot = w % 16;
    if (slot >= kTagPayload) continue;
    const int x = line * kTagPayload + slot;
    const double v = src[i];
    if (x < t.F) fac[row * m.Fp + x] = v;
    else if (hb && x == t.F) bias[row] = v;
    else if (hc && x - t.F - (hb ? 1 : 0) < t.C && x - t.F - (hb ? 1 : 0) >= 0) cells[row * t.C + (x - t.F - (hb ? 1 : 0))] = v;
  }
}

__device__ __forceinline__ unsigned long long f64_bits(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ double bits_f64(unsigned long long b) { return __longlong_as_double((long long)b); }

// RatingRec.pad carries two flags for this kernel: bit 0 = last rating of its user in the epoch, bit 1 = last of its item
constexpr int kRecLastOfUser = 1, kRecLastOfItem = 2;

// LP / LQ: compile-time upper bounds of the row lengths in lines (register arrays); tm.lp / tm.lq are the actual ones.
template <int MODEL, int LP, int LQ, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    sgd_tagged_kernel(DeviceModel m, TaggedLayout t, TaggedModel tm, const RatingRec* __restrict__ recs, int64_t nnz, double lr,
                      double* block_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LPR = 8, G = 4;
  constexpr int WARPS = THREADS / 32;
  constexpr bool kUB = (MODEL == M_BIASEDMF || MODEL == M_CAMF_CI);
  constexpr bool kIB = (MODEL == M_BIASEDMF || MODEL == M_CAMF_CU);
  constexpr bool kUC = (MODEL == M_CAMF_CU || MODEL == M_CAMF_CUCI);
  constexpr bool kIC = (MODEL == M_CAMF_CI || MODEL == M_CAMF_CUCI);
  constexpr bool kHasCond = kUC || kIC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, gw = lane / LPR;
  const unsigned gmask = ((1u << LPR) - 1u) << (gw * LPR);
  const int F = t.F, C = t.C, Dmax = m.Dmax;
  const int lp = tm.lp, lq = tm.lq;
  const int ep = t.p_payload() - F, eq = t.q_payload() - F;  // extras per row
  // per-group scratch: products [F], user-side extras [ep], item-side extras [eq]
  const int scratch = ((F + ep + eq + 1) & ~1) + 2;
  double* prod = reinterpret_cast<double*>(smem_raw) + (size_t)(warp * G + gw) * scratch;
  double* extp = prod + F;
  double* extq = extp + ep;

  const int64_t T = (int64_t)gridDim.x * WARPS * G;
  int64_t n = ((int64_t)warp * gridDim.x + blockIdx.x) * G + gw;  // consecutive ratings land on different SMs
  double acc = 0.0;
  RatingRec rec, next;
  rec.u = rec.j = rec.ctx = rec.ku = rec.kj = rec.pad = 0; rec.r = 0.0;
  if (n < nnz) rec = ld_rec_pad(recs + n);
  next = rec;
  // A rating's FIRST attempt loads both rows outright (the load is the poll; it succeeds for most ratings).  After a
  // failed attempt the group only watches the tag of each row's last line (one 16-byte load per row) until both say
  // "ready", so that a row that is late costs 32 bytes per turn and not the whole 1.5 KB.
  bool waiting = false;
  for (;;) {
    const bool active = n < nnz;
    if (!__any_sync(0xffffffffu, active)) break;
    if (active) {
      double* prow = tm.Pt + ((int64_t)rec.u * lp) * 16 + 2 * gl;
      double* qrow = tm.Qt + ((int64_t)rec.j * lq) * 16 + 2 * gl;
      if (waiting) {
        bool ready = true;
        if (gl == LPR - 1)
          ready = f64_bits(ld_cg_f64x2(prow + 16 * (lp - 1)).y) == (unsigned long long)(unsigned)rec.ku &&
                  f64_bits(ld_cg_f64x2(qrow + 16 * (lq - 1)).y) == (unsigned long long)(unsigned)rec.kj;
        const unsigned rb = __ballot_sync(gmask, ready);
        if ((rb & gmask) != gmask) continue;
      }
      double2 pl[LP], ql[LQ];
#pragma unroll
      for (int k = 0; k < LP; k++)
        if (k < lp) pl[k] = ld_cg_f64x2(prow + 16 * k);
#pragma unroll
      for (int k = 0; k < LQ; k++)
        if (k < lq) ql[k] = ld_cg_f64x2(qrow + 16 * k);
      // ---- the load is the poll: every line must carry the expected version ----------------------------------------
      bool ok = true;
      if (gl == LPR - 1) {
#pragma unroll
        for (int k = 0; k < LP; k++)
          if (k < lp) ok = ok && (f64_bits(pl[k].y) == (unsigned long long)(unsigned)rec.ku);
#pragma unroll
        for (int k = 0; k < LQ; k++)
          if (k < lq) ok = ok && (f64_bits(ql[k].y) == (unsigned long long)(unsigned)rec.kj);
      }
      const unsigned b = __ballot_sync(gmask, ok);
      waiting = (b & gmask) != gmask;
      if (!waiting) {
        const int64_t nn = n + T;
        if (nn < nnz) next = ld_rec_pad(recs + nn);  // flies during this rating's arithmetic
        // ---- products and extras to the group's scratch ----------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < (LP > LQ ? LP : LQ); k++) {
          const int x0 = k * kTagPayload + 2 * gl;
          if (x0 < F) prod[x0] = __dmul_rn(pl[k < LP ? k : 0].x, ql[k < LQ ? k : 0].x);
          else {
            if (k < LP && k < lp && x0 - F < ep) extp[x0 - F] = pl[k < LP ? k : 0].x;
            if (k < LQ && k < lq && x0 - F < eq) extq[x0 - F] = ql[k < LQ ? k : 0].x;
          }
          if (gl < LPR - 1) {
            const int x1 = x0 + 1;
            if (x1 < F) prod[x1] = __dmul_rn(pl[k < LP ? k : 0].y, ql[k < LQ ? k : 0].y);
            else {
              if (k < LP && k < lp && x1 - F < ep) extp[x1 - F] = pl[k < LP ? k : 0].y;
              if (k < LQ && k < lq && x1 - F < eq) extq[x1 - F] = ql[k < LQ ? k : 0].y;
            }
          }
        }
        __syncwarp(gmask);
        double dot = 0.0;
        for (int f = 0; f < F; f++) dot = __dadd_rn(dot, prod[f]);  // DenseMatrix.rowMult order
        const double bu = kUB ? extp[0] : 0.0;
        const double bj = kIB ? extq[0] : 0.0;
        double pred;
        if (MODEL == M_PMF) pred = dot;
        if (MODEL == M_BIASEDMF) pred = __dadd_rn(__dadd_rn(__dadd_rn(m.global_mean, bu), bj), dot);
        if (MODEL == M_CAMF_CI) pred = __dadd_rn(__dadd_rn(m.global_mean, bu), dot);
        if (MODEL == M_CAMF_CU) pred = __dadd_rn(__dadd_rn(m.global_mean, bj), dot);
        if (MODEL == M_CAMF_CUCI) pred = __dadd_rn(m.global_mean, dot);
        double cell_loss = 0.0;
        if (kHasCond) {
          const int32_t* conds = m.ctx_tab + (int64_t)rec.ctx * Dmax;
          for (int d = 0; d < Dmax; d++) {
            const int cond = __ldg(conds + d);
            if (cond < 0) continue;
            const double bic = kIC ? extq[(kIB ? 1 : 0) + cond] : 0.0;
            const double buc = kUC ? extp[(kUB ? 1 : 0) + cond] : 0.0;
            const double bsum = MODEL == M_CAMF_CUCI ? __dadd_rn(bic, buc) : (kIC ? bic : buc);
            pred = __dadd_rn(pred, bsum);
            if (kIC) cell_loss = fma(m.reg_c, __dmul_rn(bic, bic), cell_loss);
            if (kUC) cell_loss = fma(m.reg_c, __dmul_rn(buc, buc), cell_loss);
          }
        }
        const double e = __dsub_rn(rec.r, pred);
        __syncwarp(gmask);  // every lane has read the scratch; it is rewritten next turn
        // ---- steps, in the registers that hold the lines -----------------------------------------------------------------
        double sp = 0.0, sq = 0.0;
        auto step_cell = [&](double bval, int cidx) -> double {  // the row's condition cell cidx: stepped iff this context has it
          if (cidx < 0 || cidx >= C) return bval;
          const int32_t* conds = m.ctx_tab + (int64_t)rec.ctx * Dmax;
          bool hit = false;
          for (int d = 0; d < Dmax; d++) hit = hit || (__ldg(conds + d) == cidx);
          return hit ? __dadd_rn(bval, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_c, bval)))) : bval;
        };
        auto step_slot = [&](double& pv, double& qv, int x, bool has_p, bool has_q) {
          if (x < F) {
            const double po = pv, qo = qv;
            pv = __dadd_rn(po, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, qo), __dmul_rn(m.reg_u, po))));
            qv = __dadd_rn(qo, __dmul_rn(lr, __dsub_rn(__dmul_rn(e, po), __dmul_rn(m.reg_i, qo))));
            sp = fma(po, po, sp);
            sq = fma(qo, qo, sq);
          } else {
            const int ex = x - F;
            if (has_p) {
              if (kUB && ex == 0) pv = __dadd_rn(pv, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_b, pv))));
              else if (kUC) pv = step_cell(pv, ex - (kUB ? 1 : 0));
            }
            if (has_q) {
              if (kIB && ex == 0) qv = __dadd_rn(qv, __dmul_rn(lr, __dsub_rn(e, __dmul_rn(m.reg_b, qv))));
              else if (kIC) qv = step_cell(qv, ex - (kIB ? 1 : 0));
            }
          }
        };
        double dummy = 0.0;
#pragma unroll
        for (int k = 0; k < (LP > LQ ? LP : LQ); k++) {
          const int x0 = k * kTagPayload + 2 * gl;
          const bool hp = k < LP && k < lp, hq = k < LQ && k < lq;
          step_slot(k < LP ? pl[k < LP ? k : 0].x : dummy, k < LQ ? ql[k < LQ ? k : 0].x : dummy, x0, hp, hq);
          if (gl < LPR - 1) step_slot(k < LP ? pl[k < LP ? k : 0].y : dummy, k < LQ ? ql[k < LQ ? k : 0].y : dummy, x0 + 1, hp, hq);
        }
        // ---- publish: the new tag rides in the same store as the line's payload -------------------------------------------
        if (gl == LPR - 1) {
          const double tp = bits_f64((rec.pad & kRecLastOfUser) ? 0ull : (unsigned long long)(unsigned)rec.ku + 1ull);
          const double tq = bits_f64((rec.pad & kRecLastOfItem) ? 0ull : (unsigned long long)(unsigned)rec.kj + 1ull);
#pragma unroll
          for (int k = 0; k < LP; k++) pl[k].y = tp;
#pragma unroll
          for (int k = 0; k < LQ; k++) ql[k].y = tq;
        }
#pragma unroll
        for (int k = 0; k < LQ; k++)
          if (k < lq) st_cg_f64x2(qrow + 16 * k, ql[k]);
#pragma unroll
        for (int k = 0; k < LP; k++)
          if (k < lp) st_cg_f64x2(prow + 16 * k, pl[k]);
        // ---- loss (a report value: 1e-11 relative) --------------------------------------------------------------
        double lane_loss = fma(m.reg_u, sp, __dmul_rn(m.reg_i, sq));
        if (gl == 0) {
          lane_loss = __dadd_rn(lane_loss, __dmul_rn(e, e));
          if (kUB) lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bu), bu));
          if (kIB) lane_loss = __dadd_rn(lane_loss, __dmul_rn(__dmul_rn(m.reg_b, bj), bj));
          lane_loss = __dadd_rn(lane_loss, cell_loss);
        }
        acc = __dadd_rn(acc, lane_loss);
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
    double s = 0.0;
    for (int w = 0; w < WARPS; w++) s += warp_sum[w];
    block_partial[blockIdx.x] = s;
  }
}

}  // namespace cars
