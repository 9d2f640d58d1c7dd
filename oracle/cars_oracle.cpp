// cars_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Single-threaded CPU restatement of the CARSKit SGD hot path, operation for operation, in the
// reference's iteration order.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library; the product (carskit_b200/) never does.
//
// PARITY UNPINNED: the reference ships no tests, no golden vectors and cannot run here (no JVM in the
// image).  The only external pins are the java.util.Random known answers in
// tests/test_oracle.py and the layout of the reference's own sampleData/train_binary.csv (tests/test_data.py,
// against carskit_b200/data.py).  Everything else is a line-by-line reading of the Java sources cited at each
// function, cross-checked by an independent Python restatement (tests/test_oracle.py).
//
// Arithmetic rules: every value is fp64; Java never contracts a*b+c, so this
// file must be compiled with -ffp-contract=off (the Makefile does; a static_assert-style runtime
// check is in oracle_selftest_no_fma()).  The float-typed hyper-parameters regU/regI/regB/regC
// (IterativeRecommender.java:40) arrive already widened to double in cars_desc.
#include "../include/carskit_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {

// ---------------------------------------------------------------------------------------------
// java.util.Random.  librec.util.Randoms / happy.coding.math.Randoms wrap one
// static instance; DenseMatrix.init(mean, sigma) = mean + sigma * nextGaussian() row-major,
// DenseMatrix.init() / DenseVector.init() = nextDouble().
// ---------------------------------------------------------------------------------------------
struct oracle_jrandom {
  uint64_t seed;
  double next_next_gaussian;
  int have_next_next_gaussian;
};

static const uint64_t JR_MULT = 0x5DEECE66DULL;
static const uint64_t JR_MASK = (1ULL << 48) - 1;

void oracle_jr_seed(oracle_jrandom* g, int64_t seed) {
  g->seed = ((uint64_t)seed ^ JR_MULT) & JR_MASK;
  g->have_next_next_gaussian = 0;
  g->next_next_gaussian = 0.0;
}

static inline int32_t jr_next(oracle_jrandom* g, int bits) {
  g->seed = (g->seed * JR_MULT + 0xBULL) & JR_MASK;
  return (int32_t)((int64_t)g->seed >> (48 - bits));
}

int32_t oracle_jr_next_int(oracle_jrandom* g) { return jr_next(g, 32); }

// Random.nextInt(bound), bound > 0
int32_t oracle_jr_next_int_bound(oracle_jrandom* g, int32_t bound) {
  int32_t r = jr_next(g, 31);
  int32_t m = bound - 1;
  if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
  for (int32_t u = r; u - (r = u % bound) + m < 0; u = jr_next(g, 31)) {
  }
  return r;
}

double oracle_jr_next_double(oracle_jrandom* g) {
  int64_t hi = (int64_t)jr_next(g, 26);
  int64_t lo = (int64_t)jr_next(g, 27);
  return (double)((hi << 27) + lo) * 0x1.0p-53;
}


// fdlibm e_log.c (the algorithm java.lang.StrictMath.log is specified to use), for finite x > 0.
static double fdlibm_log(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                      two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                      Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                      Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                      Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  int32_t hx = (int32_t)(bits >> 32);
  int32_t k = 0;
  if (hx < 0x00100000) {  // subnormal: scale up
    k -= 54;
    x *= two54;
    std::memcpy(&bits, &x, 8);
    hx = (int32_t)(bits >> 32);
  }
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int32_t i = (hx + 0x95f64) & 0x100000;
  bits = (bits & 0xffffffffULL) | ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32);  // normalize x or x/2
  std::memcpy(&x, &bits, 8);
  k += (i >> 20);
  double f = x - 1.0, dk;
  if ((0x000fffff & (2 + hx)) < 3) {  // |f| < 2**-20
    if (f == 0.0) {
      if (k == 0) return 0.0;
      dk = (double)k;
      return dk * ln2_hi + dk * ln2_lo;
    }
    double R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    dk = (double)k;
    return dk * ln2_hi - ((R - dk * ln2_lo) - f);
  }
  double s = f / (2.0 + f);
  dk = (double)k;
  double z = s * s;
  i = hx - 0x6147a;
  double w = z * z;
  int32_t j = 0x6b851 - hx;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  i |= j;
  double R = t2 + t1;
  if (i > 0) {
    double hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

double oracle_jr_next_gaussian(oracle_jrandom* g) {
  if (g->have_next_next_gaussian) {
    g->have_next_next_gaussian = 0;
    return g->next_next_gaussian;
  }
  double v1, v2, s;
  do {
    v1 = 2 * oracle_jr_next_double(g) - 1;
    v2 = 2 * oracle_jr_next_double(g) - 1;
    s = v1 * v1 + v2 * v2;
  } while (s >= 1 || s == 0);
  // StrictMath.log is fdlibm's __ieee754_log (restated below; glibc's log differs from it in the last
  // bit on some inputs); StrictMath.sqrt is correctly rounded, like std::sqrt.
  double multiplier = std::sqrt(-2 * fdlibm_log(s) / s);
  g->next_next_gaussian = v2 * multiplier;
  g->have_next_next_gaussian = 1;
  return v1 * multiplier;
}

// librec DenseMatrix.init(mean, sigma) / DenseVector.init(mean, sigma): row-major gaussian fill.
void oracle_init_gaussian(oracle_jrandom* g, double* a, int64_t n, double mean, double sigma) {
  for (int64_t i = 0; i < n; i++) a[i] = mean + sigma * oracle_jr_next_gaussian(g);
}
// librec DenseMatrix.init() -> init(1.0) -> Randoms.uniform(0, 1) = 0 + (1 - 0) * nextDouble().
void oracle_init_uniform(oracle_jrandom* g, double* a, int64_t n) {
  for (int64_t i = 0; i < n; i++) a[i] = 0.0 + (1.0 - 0.0) * oracle_jr_next_double(g);
}

// ---------------------------------------------------------------------------------------------
// predict(u, j, c): the per-model overrides.
// ---------------------------------------------------------------------------------------------
// librec DenseMatrix.rowMult(m, mrow, n, nrow): res = 0; for j: res += m[mrow][j] * n[nrow][j]
static inline double row_mult(const double* P, int u, const double* Q, int j, int F) {
  const double* p = P + (int64_t)u * F;
  const double* q = Q + (int64_t)j * F;
  double res = 0;
  for (int f = 0; f < F; f++) res += p[f] * q[f];
  return res;
}

static inline double predict_one(const cars_desc* d, const cars_model_arrays* m, int u, int j, int c) {
  const int F = d->num_factors;
  const int C = d->num_conditions;
  double pred;
  switch (d->model) {
    case CARS_PMF:  // PMF.java:93-96
      return row_mult(m->P, u, m->Q, j, F);
    case CARS_BIASEDMF:  // BiasedMF.java:112-114
      return d->global_mean + m->user_bias[u] + m->item_bias[j] + row_mult(m->P, u, m->Q, j, F);
    case CARS_CAMF_C:  // CAMF_C.java:65-72
      pred = d->global_mean + m->user_bias[u] + m->item_bias[j] + row_mult(m->P, u, m->Q, j, F);
      for (int k = d->ctx_ptr[c]; k < d->ctx_ptr[c + 1]; k++) pred += m->cond_bias[d->ctx_cond[k]];
      return pred;
    case CARS_CAMF_CI:  // CAMF_CI.java:65-72
      pred = d->global_mean + m->user_bias[u] + row_mult(m->P, u, m->Q, j, F);
      for (int k = d->ctx_ptr[c]; k < d->ctx_ptr[c + 1]; k++)
        pred += m->ic_bias[(int64_t)j * C + d->ctx_cond[k]];
      return pred;
    case CARS_CAMF_CU:  // CAMF_CU.java:62-69
      pred = d->global_mean + m->item_bias[j] + row_mult(m->P, u, m->Q, j, F);
      for (int k = d->ctx_ptr[c]; k < d->ctx_ptr[c + 1]; k++)
        pred += m->uc_bias[(int64_t)u * C + d->ctx_cond[k]];
      return pred;
    case CARS_CAMF_CUCI:  // CAMF_CUCI.java:68-74 (`pred += ic + uc`: the two cells are added first)
      pred = d->global_mean + row_mult(m->P, u, m->Q, j, F);
      for (int k = d->ctx_ptr[c]; k < d->ctx_ptr[c + 1]; k++) {
        int cond = d->ctx_cond[k];
        pred += m->ic_bias[(int64_t)j * C + cond] + m->uc_bias[(int64_t)u * C + cond];
      }
      return pred;
    default:
      return NAN;
  }
}

// Recommender.predict(u, j, c, bound)  (Recommender.java:306-317)
static inline double predict_bound(const cars_desc* d, const cars_model_arrays* m, int u, int j, int c,
                                   int bound, double min_rate, double max_rate) {
  double pred = predict_one(d, m, u, j, c);
  if (bound) {
    if (pred > max_rate) pred = max_rate;
    if (pred < min_rate) pred = min_rate;
  }
  return pred;
}

int oracle_predict(const cars_desc* d, const cars_model_arrays* m, int64_t n, const int32_t* u,
                   const int32_t* j, const int32_t* ctx, int32_t bound, double min_rate,
                   double max_rate, double* out) {
  for (int64_t i = 0; i < n; i++)
    out[i] = predict_bound(d, m, u[i], j[i], ctx ? ctx[i] : -1, bound, min_rate, max_rate);
  return 0;
}

// evalRatings() accumulation (Recommender.java:518-545): skips NaN predictions, sums |err| and err^2.
int oracle_eval_ratings(const cars_desc* d, const cars_model_arrays* m, int64_t n, const int32_t* u,
                        const int32_t* j, const int32_t* ctx, const double* r, double min_rate,
                        double max_rate, double* sum_abs, double* sum_sq, int64_t* count) {
  double sum_maes = 0, sum_mses = 0;
  int64_t num = 0;
  for (int64_t i = 0; i < n; i++) {
    double pred = predict_bound(d, m, u[i], j[i], ctx ? ctx[i] : -1, 1, min_rate, max_rate);
    if (std::isnan(pred)) continue;
    double err = std::fabs(r[i] - pred);
    sum_maes += err;
    sum_mses += err * err;
    num++;
  }
  *sum_abs = sum_maes;
  *sum_sq = sum_mses;
  if (count) *count = num;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One epoch == one iteration of the `for (int iter ...)` body of buildModel(), through `loss *= 0.5`.
// Returns the loss.  The model arrays are updated in place, exactly like the Java containers.
// ---------------------------------------------------------------------------------------------
// The shared tail of every model (e.g. CAMF_CI.java:110-121).
static inline void factor_step(double* p, double* q, int F, double euj, double lRate, double regU,
                               double regI, double* loss) {
  for (int f = 0; f < F; f++) {
    double puf = p[f];
    double qjf = q[f];
    double delta_u = euj * qjf - regU * puf;
    double delta_j = euj * puf - regI * qjf;
    p[f] += lRate * delta_u;
    q[f] += lRate * delta_j;
    *loss += regU * puf * puf + regI * qjf * qjf;
  }
}

double oracle_epoch(const cars_desc* d, const cars_model_arrays* m, double lRate) {
  const int F = d->num_factors;
  const int C = d->num_conditions;
  const double regU = d->reg_u, regI = d->reg_i, regB = d->reg_b, regC = d->reg_c;
  double loss = 0;
  for (int64_t n = 0; n < d->nnz; n++) {
    const int u = d->u[n];
    const int j = d->j[n];
    const int ctx = d->ctx ? d->ctx[n] : -1;
    const double rujc = d->r[n];
    double* p = m->P + (int64_t)u * F;
    double* q = m->Q + (int64_t)j * F;

    double pred = predict_one(d, m, u, j, ctx);
    double euj = rujc - pred;
    loss += euj * euj;

    switch (d->model) {
      case CARS_PMF:  // PMF.java:52-71
        break;
      case CARS_BIASEDMF: {  // BiasedMF.java:75-85
        double bu = m->user_bias[u];
        double sgd = euj - regB * bu;
        m->user_bias[u] += lRate * sgd;
        loss += regB * bu * bu;
        double bj = m->item_bias[j];
        sgd = euj - regB * bj;
        m->item_bias[j] += lRate * sgd;
        loss += regB * bj * bj;
        break;
      }
      case CARS_CAMF_C: {  // CAMF_C.java:94-115
        double bu = m->user_bias[u];
        double sgd = euj - regB * bu;
        m->user_bias[u] += lRate * sgd;
        loss += regB * bu * bu;
        double bj = m->item_bias[j];
        sgd = euj - regB * bj;
        m->item_bias[j] += lRate * sgd;
        loss += regB * bj * bj;
        double bc_sum = 0;
        for (int k = d->ctx_ptr[ctx]; k < d->ctx_ptr[ctx + 1]; k++) {
          int cond = d->ctx_cond[k];
          double bc = m->cond_bias[cond];
          bc_sum += bc;
          sgd = euj - regC * bc;
          m->cond_bias[cond] += lRate * sgd;
        }
        loss += regB * bc_sum;  // reference quirk (:115): regB, and the sum is not squared
        break;
      }
      case CARS_CAMF_CI: {  // CAMF_CI.java:94-108
        double bu = m->user_bias[u];
        double sgd = euj - regB * bu;
        m->user_bias[u] += lRate * sgd;
        loss += regB * bu * bu;
        double Bic_sum = 0;
        for (int k = d->ctx_ptr[ctx]; k < d->ctx_ptr[ctx + 1]; k++) {
          int cond = d->ctx_cond[k];
          double Bic = m->ic_bias[(int64_t)j * C + cond];
          Bic_sum += Bic * Bic;  // Math.pow(Bic, 2): HotSpot folds pow(x, 2) to x * x
          sgd = euj - regC * Bic;
          m->ic_bias[(int64_t)j * C + cond] = Bic + lRate * sgd;
        }
        loss += regC * Bic_sum;
        break;
      }
      case CARS_CAMF_CU: {  // CAMF_CU.java:91-105
        double bj = m->item_bias[j];
        double sgd = euj - regB * bj;
        m->item_bias[j] += lRate * sgd;
        loss += regB * bj * bj;
        double Buc_sum = 0;
        for (int k = d->ctx_ptr[ctx]; k < d->ctx_ptr[ctx + 1]; k++) {
          int cond = d->ctx_cond[k];
          double Buc = m->uc_bias[(int64_t)u * C + cond];
          Buc_sum += Buc * Buc;
          sgd = euj - regC * Buc;
          m->uc_bias[(int64_t)u * C + cond] = Buc + lRate * sgd;
        }
        loss += regC * Buc_sum;
        break;
      }
      case CARS_CAMF_CUCI: {  // CAMF_CUCI.java:98-114
        double Buc_sum = 0;
        double Bic_sum = 0;
        for (int k = d->ctx_ptr[ctx]; k < d->ctx_ptr[ctx + 1]; k++) {
          int cond = d->ctx_cond[k];
          double Buc = m->uc_bias[(int64_t)u * C + cond];
          double Bic = m->ic_bias[(int64_t)j * C + cond];
          Buc_sum += Buc * Buc;
          Bic_sum += Bic * Bic;
          double sgdu = euj - regC * Buc;
          double sgdj = euj - regC * Bic;
          m->uc_bias[(int64_t)u * C + cond] = Buc + lRate * sgdu;
          m->ic_bias[(int64_t)j * C + cond] = Bic + lRate * sgdj;
        }
        loss += regC * Bic_sum + regC * Buc_sum;
        break;
      }
      default:
        return NAN;
    }
    factor_step(p, q, F, euj, lRate, regU, regI, &loss);
  }
  loss *= 0.5;
  return loss;
}

// ---------------------------------------------------------------------------------------------
// Epoch control: IterativeRecommender.isConverged (:145-199) and updateLRate (:216-229).
// earlyStopMeasure == null or Loss only (MAE/RMSE early stop needs evalRatings(); the caller sets
// `measure` before the call in that case and passes use_measure = 1).
// ---------------------------------------------------------------------------------------------
struct oracle_epoch_state {
  double lRate;         // instance field, initialised from the float initLRate (:106)
  double loss, last_loss;
  double measure, last_measure;
  float maxLRate;       // -max, default -1
  float decay;          // -decay, default -1
  int32_t isBoldDriver;
  int32_t early_stop;   // 0 none, 1 Loss, 2 external measure already stored in `measure`
};

void oracle_update_lrate(oracle_epoch_state* s, int iter) {
  if (s->lRate <= 0) return;
  if (s->isBoldDriver && iter > 1)
    s->lRate = std::fabs(s->last_loss) > std::fabs(s->loss) ? s->lRate * 1.05 : s->lRate * 0.5;
  else if (s->decay > 0 && s->decay < 1)
    s->lRate *= s->decay;
  if (s->maxLRate > 0 && s->lRate > s->maxLRate) s->lRate = s->maxLRate;
}

// returns 1 converged, 0 not, -1 NaN/Inf loss (Java calls System.exit(-1) there)
int oracle_is_converged(oracle_epoch_state* s, int iter) {
  if (s->early_stop == 1) {
    s->measure = s->loss;
    s->last_measure = s->last_loss;
  }
  float delta_measure = (float)(s->last_measure - s->measure);
  if (std::isnan(s->loss) || std::isinf(s->loss)) return -1;
  bool cond1 = std::fabs(s->loss) < 1e-5;
  bool cond2 = (delta_measure > 0) && (delta_measure < 1e-5);
  bool converged = cond1 || cond2;
  if (!converged) oracle_update_lrate(s, iter);
  s->last_loss = s->loss;
  s->last_measure = s->measure;
  return converged ? 1 : 0;
}

// Full buildModel(): epochs until numIters or convergence.  Returns the number of epochs run, or -1 on
// NaN/Inf loss.  losses (optional) receives the per-epoch loss.
int oracle_build_model(const cars_desc* d, const cars_model_arrays* m, oracle_epoch_state* s, int numIters,
                       double* losses) {
  int iter;
  for (iter = 1; iter <= numIters; iter++) {
    s->loss = oracle_epoch(d, m, s->lRate);
    if (losses) losses[iter - 1] = s->loss;
    int c = oracle_is_converged(s, iter);
    if (c < 0) return -1;
    if (c) return iter;
  }
  return numIters;
}

// ---------------------------------------------------------------------------------------------
// evalRankings scoring loop + top-N cut (Recommender.java:797-824), one query at a time exactly as written:
//   for (Integer j : candItems) if (!ratedItems.contains(j)) { rank = ranking(u, j, c);
//       if (!Double.isNaN(rank)) if (rank > binThold) itemScores.add((j, rank)); }
//   Lists.sortList(itemScores, true);   // stable Collections.sort, comparator -(a.value.compareTo(b.value))
//   recomd = itemScores.subList(0, numRecs)
// `cand` is candItems in the reference's HashSet iteration order.
// ---------------------------------------------------------------------------------------------
static inline int java_double_compare(double a, double b) {  // Double.compare for non-NaN values
  if (a < b) return -1;
  if (a > b) return 1;
  int64_t x, y;
  std::memcpy(&x, &a, 8);
  std::memcpy(&y, &b, 8);
  return x == y ? 0 : (x < y ? -1 : 1);  // -0.0 < +0.0
}

int oracle_rank_topn(const cars_desc* d, const cars_model_arrays* m, int64_t num_queries, const int32_t* qu,
                     const int32_t* qc, int32_t num_cand, const int32_t* cand, const int64_t* rated_ptr,
                     const int32_t* rated_items, double bin_thold, int32_t num_recs, int32_t* out_items,
                     double* out_scores, int32_t* out_count, int32_t* out_kept) {
  std::vector<char> rated((size_t)d->num_items, 0);
  std::vector<std::pair<int32_t, double>> scores;
  for (int64_t q = 0; q < num_queries; q++) {
    const int u = qu[q], c = qc ? qc[q] : 0;
    if (rated_ptr)
      for (int64_t i = rated_ptr[q]; i < rated_ptr[q + 1]; i++) rated[(size_t)rated_items[i]] = 1;
    scores.clear();
    for (int32_t k = 0; k < num_cand; k++) {
      const int j = cand[k];
      if (rated[(size_t)j]) continue;
      const double rank = predict_one(d, m, u, j, c);
      if (!std::isnan(rank))
        if (rank > bin_thold) scores.emplace_back(j, rank);
    }
    if (rated_ptr)
      for (int64_t i = rated_ptr[q]; i < rated_ptr[q + 1]; i++) rated[(size_t)rated_items[i]] = 0;
    std::stable_sort(scores.begin(), scores.end(), [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b) {
      return java_double_compare(a.second, b.second) > 0;
    });
    out_kept[q] = (int32_t)scores.size();
    const int32_t n = (int32_t)scores.size() < num_recs ? (int32_t)scores.size() : num_recs;
    out_count[q] = n;
    for (int32_t i = 0; i < num_recs; i++) {
      out_items[q * num_recs + i] = i < n ? scores[(size_t)i].first : -1;
      out_scores[q * num_recs + i] = i < n ? scores[(size_t)i].second : 0.0;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Data-side helpers the tests need to restate the reference's inputs.
// ---------------------------------------------------------------------------------------------
// SparseMatrix.getGlobalAvg (src/carskit/data/structure/SparseMatrix.java:49-56): sum()/size() where
// librec's sum() is a sequential sum of rowData and size() counts non-zero entries.
double oracle_global_mean(const double* r, int64_t nnz) {
  double sum = 0;
  int64_t cnt = 0;
  for (int64_t i = 0; i < nnz; i++) {
    sum += r[i];
    if (r[i] != 0) cnt++;
  }
  return sum / (double)cnt;
}

// DataDAO.toTraditionalSparseMatrix (DataDAO.java:1241-1257): the 2-D `train` matrix BiasedMF/PMF
// iterate: one entry per (user, item) = mean of that pair's ratings over contexts, CRS order (user
// ascending, item ascending).  Input: the {ui x ctx} arrays in CRS order plus ui -> (u, j).
// Output arrays must hold num_ui entries; returns the number of (u, j) entries written.
int64_t oracle_to_traditional(int64_t nnz, const int32_t* ui, const double* r, int32_t num_ui,
                              const int32_t* ui_user, const int32_t* ui_item, int32_t* out_u,
                              int32_t* out_j, double* out_r) {
  // mean per ui in entry order (Stats.mean = sequential sum / n), then sort by (u, j).
  std::vector<double> sum(num_ui, 0.0);
  std::vector<int64_t> cnt(num_ui, 0);
  for (int64_t n = 0; n < nnz; n++) {
    sum[ui[n]] += r[n];
    cnt[ui[n]]++;
  }
  std::vector<int32_t> order;
  for (int32_t k = 0; k < num_ui; k++)
    if (cnt[k] > 0) order.push_back(k);
  // insertion into a (u, j)-keyed table then CRS construction == sort by (u, j)
  std::vector<int64_t> key(order.size());
  for (size_t k = 0; k < order.size(); k++)
    key[k] = ((int64_t)ui_user[order[k]] << 32) | (uint32_t)ui_item[order[k]];
  std::vector<size_t> idx(order.size());
  for (size_t k = 0; k < idx.size(); k++) idx[k] = k;
  // simple stable sort
  std::vector<size_t> tmp(idx.size());
  // (std::stable_sort without <algorithm> lambdas kept trivial)
  struct Cmp {
    const std::vector<int64_t>* k;
    bool operator()(size_t a, size_t b) const { return (*k)[a] < (*k)[b]; }
  };
  // merge sort
  for (size_t width = 1; width < idx.size(); width *= 2) {
    for (size_t lo = 0; lo < idx.size(); lo += 2 * width) {
      size_t mid = lo + width < idx.size() ? lo + width : idx.size();
      size_t hi = lo + 2 * width < idx.size() ? lo + 2 * width : idx.size();
      size_t a = lo, b = mid, o = lo;
      while (a < mid && b < hi) tmp[o++] = key[idx[b]] < key[idx[a]] ? idx[b++] : idx[a++];
      while (a < mid) tmp[o++] = idx[a++];
      while (b < hi) tmp[o++] = idx[b++];
    }
    idx.swap(tmp);
  }
  int64_t w = 0;
  for (size_t k = 0; k < idx.size(); k++) {
    int32_t id = order[idx[k]];
    double mean = sum[id] / (double)cnt[id];
    if (mean == 0) continue;  // zero entries vanish from a librec SparseMatrix
    out_u[w] = ui_user[id];
    out_j[w] = ui_item[id];
    out_r[w] = mean;
    w++;
  }
  return w;
}

// 1 when this object was compiled without FMA contraction (required), 0 otherwise.
int oracle_selftest_no_fma(void) {
  volatile double a = 1.0 + 0x1.0p-30, b = 1.0 - 0x1.0p-30, c = -1.0;
  double t = a * b + c;  // exact product 1 - 2^-60 rounds to 1.0 -> 0.0 ; an FMA gives -2^-60
  return t == 0.0 ? 1 : 0;
}

const char* oracle_version(void) { return "cars_oracle 1 (CPU restatement of irecsys/CARSKit v0.4.0 SGD loops)"; }

}  // extern "C"
