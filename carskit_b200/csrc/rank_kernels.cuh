// rank_kernels.cuh -- K6: the scoring loop and top-N selection of Recommender.evalRankings()
// (src/carskit/generic/Recommender.java:797-824): for one (user, context) query, every candidate item that the
// user has not rated in that context is scored with ranking(u, j, c) = predict(u, j, c), kept if
// score > binThold, sorted by descending score (Lists.sortList: a STABLE sort, ties keep the candidates'
// iteration order) and cut to numRecs.
//
// Scores are bit-identical to Java's: the dot product is accumulated in f = 0..F-1 order with separately
// rounded multiply and add, exactly like predict_kernel -- which is why this is an fp64 SIMT contraction and
// not a tensor-core GEMM: DMMA accumulates with fused multiply-adds in a different order (and tcgen05 has no
// fp64 kind), and the north star asks for bit-exact top-N indices.  The P x Q^T structure is still exploited:
// a CTA stages a tile of 64 query rows and 64 candidate rows in shared memory and every thread keeps a 4 x 4
// block of independent accumulation chains in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sgd_kernels.cuh"

namespace cars {

constexpr int kRankTQ = 64;  // queries per tile
constexpr int kRankTJ = 64;  // candidates per tile
constexpr int kRankFC = 64;  // factors staged in shared memory per pass

// Double.compareTo order as an unsigned key (larger key = larger double; -0.0 < +0.0; NaN never gets here)
__device__ __forceinline__ unsigned long long sortable_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
// "not a candidate for this query": below every real key
constexpr unsigned long long kRankDropped = 0ull;

__device__ __forceinline__ void rank_cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// scores[q][c] = sortable key of predict(u_q, cand_c, ctx_q) if it is a number > bin_thold, else kRankDropped
template <int MODEL>
__global__ void __launch_bounds__(256) rank_score_kernel(DeviceModel m, int64_t q0, int64_t nq, const int32_t* __restrict__ qu,
                                                         const int32_t* __restrict__ qc, int32_t num_cand,
                                                         const int32_t* __restrict__ cand, double bin_thold,
                                                         unsigned long long* __restrict__ keys /*[nq x num_cand]*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = m.F;
  const int FC = F < kRankFC ? F : kRankFC;  // factors staged per pass; the chains simply continue across passes
  const int ld = FC + 1;                     // padded row stride: 16 rows read at the same f hit 16 different banks
  double* Ps = reinterpret_cast<double*>(smem_raw);
  double* Qs = Ps + kRankTQ * ld;
  const int64_t qt = (int64_t)blockIdx.y * kRankTQ;
  const int ct = blockIdx.x * kRankTJ;
  // 4 x 4 register tile: thread (ty, tx) scores queries 4 ty .. 4 ty + 3 against candidates tx, tx + 16, tx + 32, tx + 48 --
  // 8 shared-memory loads feed 16 multiply-add pairs per factor (the 1 x 4 tile of round 1 needed 5 loads for 4 pairs
  // and sat at 89 % of the L1TEX pipe with the fp64 pipe at 19 %, profiles/r2/ncu_summary_rank_score.txt)
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
  for (int f0 = 0; f0 < F; f0 += FC) {
    const int fl = F - f0 < FC ? F - f0 : FC;
    __syncthreads();
    // tile fill by cp.async: all of a thread's 32 copies are in flight together and hold no register (with plain loads
    // the STS behind each LDG carried 54 % of the kernel's stall samples, profiles/r2/ncu_summary_rank_score_4x4_tile.txt)
    for (int i = threadIdx.x; i < kRankTQ * fl; i += 256) {
      const int r = i / fl, f = i % fl;
      const int64_t q = qt + r;
      if (q < nq) rank_cp_async_8(Ps + r * ld + f, m.P + (int64_t)qu[q0 + q] * m.Fp + f0 + f);
      else Ps[r * ld + f] = 0.0;
    }
    for (int i = threadIdx.x; i < kRankTJ * fl; i += 256) {
      const int r = i / fl, f = i % fl;
      const int c = ct + r;
      if (c < num_cand) rank_cp_async_8(Qs + r * ld + f, m.Q + (int64_t)cand[c] * m.Fp + f0 + f);
      else Qs[r * ld + f] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const double* p0 = Ps + (4 * ty) * ld;
    const double* qq = Qs + tx * ld;
    for (int f = 0; f < fl; f++) {  // DenseMatrix.rowMult order: every pair's chain runs f ascending, mul and add rounded separately
      double pv[4], qv[4];
#pragma unroll
      for (int a = 0; a < 4; a++) pv[a] = p0[a * ld + f];
#pragma unroll
      for (int c = 0; c < 4; c++) qv[c] = qq[(16 * c) * ld + f];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c] = __dadd_rn(acc[a][c], __dmul_rn(pv[a], qv[c]));
    }
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int64_t q = qt + 4 * ty + a;
    if (q >= nq) continue;
    const int u = qu[q0 + q], ctx = qc ? qc[q0 + q] : 0;
#pragma unroll
    for (int c4 = 0; c4 < 4; c4++) {
      const int c = ct + tx + 16 * c4;
      if (c >= num_cand) continue;
      const double sc = predict_from_dot<MODEL>(m, u, cand[c], ctx, acc[a][c4]);
      keys[q * num_cand + c] = (sc == sc && sc > bin_thold) ? sortable_key(sc) : kRankDropped;  // !isNaN && rank > binThold
    }
  }
}

// items the user rated in this context in the TRAINING set leave the candidate list (Recommender.java:792,800)
__global__ void __launch_bounds__(256) rank_drop_rated_kernel(int64_t q0, int64_t nq, const int64_t* __restrict__ rated_ptr,
                                                              const int32_t* __restrict__ rated_items,
                                                              const int32_t* __restrict__ cand_index_of_item,
                                                              int32_t num_cand, unsigned long long* __restrict__ keys) {
  const int64_t q = blockIdx.x;
  if (q >= nq) return;
  for (int64_t i = rated_ptr[q0 + q] + threadIdx.x; i < rated_ptr[q0 + q + 1]; i += 256) {
    const int c = cand_index_of_item[rated_items[i]];
    if (c >= 0) keys[q * num_cand + c] = kRankDropped;
  }
}

// one CTA per query: num_recs rounds of "largest key, lowest candidate index on ties" == the first num_recs
// entries of the reference's stable descending sort
__global__ void __launch_bounds__(256) rank_select_kernel(int64_t q0, int64_t nq, int32_t num_cand, const int32_t* __restrict__ cand,
                                                          int32_t num_recs, unsigned long long* __restrict__ keys,
                                                          int32_t* __restrict__ out_items, double* __restrict__ out_scores,
                                                          int32_t* __restrict__ out_count, int32_t* __restrict__ out_kept) {
  __shared__ unsigned long long sk[8];
  __shared__ int si[8];
  __shared__ int skept[8];
  const int64_t q = blockIdx.x;
  if (q >= nq) return;
  unsigned long long* row = keys + q * num_cand;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int count = 0;
  for (int round = 0; round < num_recs; round++) {
    unsigned long long best = kRankDropped;
    int bi = 0x7fffffff, kept = 0;
    for (int c = threadIdx.x; c < num_cand; c += 256) {
      const unsigned long long k = row[c];
      if (k != kRankDropped) kept++;
      if (k > best) { best = k; bi = c; }  // strided scan visits a thread's candidates in ascending order
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ok > best || (ok == best && oi < bi)) { best = ok; bi = oi; }
      kept += __shfl_down_sync(0xffffffffu, kept, o);
    }
    if (lane == 0) { sk[warp] = best; si[warp] = bi; skept[warp] = kept; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int kt = 0;
      for (int w = 0; w < 8; w++) {
        kt += skept[w];
        if (sk[w] > best || (sk[w] == best && si[w] < bi)) { best = sk[w]; bi = si[w]; }
      }
      sk[0] = best; si[0] = bi; skept[0] = kt;
    }
    __syncthreads();
    best = sk[0]; bi = si[0];
    if (round == 0 && threadIdx.x == 0) out_kept[q0 + q] = skept[0];  // itemScores.size()
    if (best == kRankDropped) break;                                 // fewer than num_recs survivors
    if (threadIdx.x == 0) {
      out_items[(q0 + q) * num_recs + round] = cand[bi];
      const unsigned long long b = (best & 0x8000000000000000ull) ? (best & 0x7fffffffffffffffull) : ~best;
      out_scores[(q0 + q) * num_recs + round] = __longlong_as_double((long long)b);
      row[bi] = kRankDropped;
    }
    count++;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_count[q0 + q] = count;
}

// Single-pass selection for num_recs <= kRankRegK (the reference's default is 10): every thread keeps the num_recs best
// (key, candidate) pairs of ITS strided share of the row in registers -- so the row of keys is read ONCE instead of
// num_recs times -- then the CTA's 256 x num_recs survivors are merged in shared memory by num_recs rounds of
// "largest key, lowest candidate index".  Same result as rank_select_kernel: the reference's stable descending sort.
constexpr int kRankRegK = 16;

template <int K>  // K >= num_recs, a compile-time bound so that the thread's list stays in registers (K = 4, 8, 10, 16)
__global__ void __launch_bounds__(256) rank_select_topk_kernel(int64_t q0, int64_t nq, int32_t num_cand, const int32_t* __restrict__ cand,
                                                               int32_t num_recs, const unsigned long long* __restrict__ keys,
                                                               int32_t* __restrict__ out_items, double* __restrict__ out_scores,
                                                               int32_t* __restrict__ out_count, int32_t* __restrict__ out_kept) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* mk = reinterpret_cast<unsigned long long*>(smem_raw);  // [256 * num_recs] merged keys
  int* mi = reinterpret_cast<int*>(mk + 256 * num_recs);                     // [256 * num_recs] candidate indices
  __shared__ unsigned long long sk[8];
  __shared__ int si[8], sp[8], skept[8];
  const int64_t q = blockIdx.x;
  if (q >= nq) return;
  const unsigned long long* row = keys + q * num_cand;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long tk[K];
  int ti[K];
#pragma unroll
  for (int k = 0; k < K; k++) { tk[k] = kRankDropped; ti[k] = 0x7fffffff; }
  int kept = 0;
  for (int c = threadIdx.x; c < num_cand; c += 256) {
    const unsigned long long key = __ldg(row + c);
    if (key == kRankDropped) continue;
    kept++;
    if (key <= tk[K - 1]) continue;  // not better than the thread's K-th best (K >= num_recs keeps a superset)
    // insertion after equal keys: this thread sees its candidates in ascending index order (stable)
    unsigned long long ck = key;
    int ci = c;
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (ck > tk[k]) {
        const unsigned long long t = tk[k]; tk[k] = ck; ck = t;
        const int u = ti[k]; ti[k] = ci; ci = u;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; k++)
    if (k < num_recs) { mk[threadIdx.x * num_recs + k] = tk[k]; mi[threadIdx.x * num_recs + k] = ti[k]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_down_sync(0xffffffffu, kept, o);
  if (lane == 0) skept[warp] = kept;
  __syncthreads();
  if (threadIdx.x == 0) {
    int kt = 0;
    for (int w = 0; w < 8; w++) kt += skept[w];
    out_kept[q0 + q] = kt;  // itemScores.size()
  }
  const int total = 256 * num_recs;
  int count = 0;
  for (int round = 0; round < num_recs; round++) {
    unsigned long long best = kRankDropped;
    int bi = 0x7fffffff, bp = -1;
    for (int p = threadIdx.x; p < total; p += 256) {
      const unsigned long long k = mk[p];
      const int i = mi[p];
      if (k > best || (k == best && k != kRankDropped && i < bi)) { best = k; bi = i; bp = p; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      const int op = __shfl_down_sync(0xffffffffu, bp, o);
      if (ok > best || (ok == best && oi < bi)) { best = ok; bi = oi; bp = op; }
    }
    if (lane == 0) { sk[warp] = best; si[warp] = bi; sp[warp] = bp; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++)
        if (sk[w] > sk[0] || (sk[w] == sk[0] && si[w] < si[0])) { sk[0] = sk[w]; si[0] = si[w]; sp[0] = sp[w]; }
      if (sk[0] != kRankDropped) {
        out_items[(q0 + q) * num_recs + round] = cand[si[0]];
        const unsigned long long b = (sk[0] & 0x8000000000000000ull) ? (sk[0] & 0x7fffffffffffffffull) : ~sk[0];
        out_scores[(q0 + q) * num_recs + round] = __longlong_as_double((long long)b);
        mk[sp[0]] = kRankDropped;
      }
    }
    __syncthreads();
    if (sk[0] == kRankDropped) break;
    count++;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_count[q0 + q] = count;
}

}  // namespace cars
