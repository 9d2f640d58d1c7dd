// fm_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of carskit.alg.cars.adaptation.dependent.FM (src/carskit/alg/cars/adaptation/dependent/
// FM.java), the ALS / coordinate-descent factorization machine of SURVEY.md row A7.  PARITY UNPINNED (no
// reference tests, no JVM; see cars_oracle.cpp).  Two statements of the same algorithm:
//
//   oracle_fm_dense_*   the LITERAL algorithm: dense feature vectors of length p, `fvalues` as a dense
//                       size x p table, every loop over all p coordinates and all `size` rows exactly as
//                       FM.java:93-113 (predict) and :115-220 (buildModel) write them.  O(k*p*size):
//                       tiny inputs only.  This is the ground truth of the reference's arithmetic.
//   oracle_fm_sparse_*  the same coordinate order, using that every row has three non-zero features
//                       (x_u = 1, x_{U+j} = 1, x_{U+I+ctx} = 1/numContextDims when that index is < p).
//                       Numerators are summed over rows(l) in ascending row order (bit-identical to the
//                       dense loop, whose other terms are exact zeros).  Denominators: the dense loop adds
//                       (x^2 + reg) over ALL rows sequentially; `closed_den` = 0 replays that sequential
//                       sum (O(p*size) adds), `closed_den` = 1 uses cnt*x^2 + size*reg resp.
//                       sum_{rows(l)} h^2 + size*reg, which is what the CUDA engine computes and differs
//                       from the literal sum by rounding only (~1e-13 relative; tests state the bound).
//
// Java arithmetic notes kept here: `size + regLw` at FM.java:159 is int + float = FLOAT addition;
// regLw / regLf are floats widened to double at every other use; Math.pow(x, 2) == x * x (exact square,
// correctly rounded); `0 - x` (not -x) so an empty coordinate yields +0.0.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

struct oracle_fm_problem {
  int32_t num_users, num_items, num_conditions, num_context_dims, k;
  int32_t pad;
  int64_t size;  // number of training ratings, in trainMatrix iteration order
  const int32_t* u;
  const int32_t* j;
  const int32_t* ctx;
  const double* r;
  float reg_lw, reg_lf;
};

struct oracle_fm_model {
  double* w0;  // [1]
  double* w;   // [p]
  double* V;   // [p x k]
};

static inline int fm_p(const oracle_fm_problem* pr) { return pr->num_users + pr->num_items + pr->num_conditions; }

// FM.getFeatureVector (FM.java:76-91): dense x of length p
static void feature_vector(const oracle_fm_problem* pr, int u, int j, int c, double* fs) {
  const int p = fm_p(pr);
  const int indexu = u, indexj = pr->num_users + j, indexc = pr->num_users + pr->num_items + c;
  for (int i = 0; i < p; ++i) {
    if (i == indexu || i == indexj)
      fs[i] = 1;
    else if (i == indexc)
      fs[i] = 1.0 / pr->num_context_dims;
    else
      fs[i] = 0.0;
  }
}

// FM.predict (FM.java:93-113), literal
double oracle_fm_dense_predict(const oracle_fm_problem* pr, const oracle_fm_model* m, int u, int j, int c) {
  const int p = fm_p(pr), k = pr->k;
  std::vector<double> fs((size_t)p);
  feature_vector(pr, u, j, c, fs.data());
  double pred = *m->w0;
  for (int i = 0; i < p; ++i) pred += m->w[i] * fs[i];
  double sum = 0.0;
  for (int f = 0; f < k; ++f) {
    double sum1 = 0.0, sum2 = 0.0;
    for (int i = 0; i < p; ++i) {
      double dot = m->V[(size_t)i * k + f] * fs[i];
      sum1 += m->V[(size_t)i * k + f] * fs[i];
      sum2 += dot * dot;  // Math.pow(dot, 2)
    }
    sum += sum1 * sum1 - sum2;
  }
  pred += 0.5 * sum;
  return pred;
}

// FM.buildModel (FM.java:115-220), literal.  errors [size], Q [size x k] and fvalues [size x p] are
// work arrays owned by the caller (so tests can inspect them).  Returns the final `loss`.
double oracle_fm_dense_build(const oracle_fm_problem* pr, const oracle_fm_model* m, int num_iters, double* errors,
                             double* Q, double* fvalues) {
  const int p = fm_p(pr), k = pr->k;
  const int64_t size = pr->size;
  const float regLw = pr->reg_lw, regLf = pr->reg_lf;
  double w0 = *m->w0;
  std::vector<double> fs((size_t)p);
  for (int64_t n = 0; n < size; ++n) {  // :118-146
    feature_vector(pr, pr->u[n], pr->j[n], pr->ctx[n], fs.data());
    *m->w0 = w0;
    double pred = oracle_fm_dense_predict(pr, m, pr->u[n], pr->j[n], pr->ctx[n]);
    errors[n] = pr->r[n] - pred;
    for (int f = 0; f < k; ++f) {
      double value = 0;
      for (int i = 0; i < p; ++i) {
        value += m->V[(size_t)i * k + f] * fs[i];
        fvalues[(size_t)n * p + i] = fs[i];
      }
      Q[(size_t)n * k + f] = value;
    }
  }
  double loss = 0;
  for (int iter = 1; iter <= num_iters; iter++) {  // :148
    loss = 0;
    double update_w0 = 0;  // :152-169
    for (int64_t i = 0; i < size; ++i) {
      double err = errors[i];
      update_w0 += err - w0;
      loss += err * err;
    }
    update_w0 = update_w0 / ((float)size + regLw);  // int + float -> float
    update_w0 = 0 - update_w0;
    for (int64_t i = 0; i < size; ++i) errors[i] = errors[i] + update_w0 - w0;
    loss += regLw * w0 * w0;
    w0 = update_w0;
    for (int l = 0; l < p; ++l) {  // :172-191
      double update_wl = 0, sum = 0;
      for (int64_t i = 0; i < size; ++i) {
        double fl = fvalues[(size_t)i * p + l];
        update_wl += (errors[i] - m->w[l] * fl) * fl;
        sum += fl * fl + regLw;
      }
      update_wl = 0 - update_wl / sum;
      for (int64_t i = 0; i < size; ++i) errors[i] = errors[i] + (update_wl - m->w[l]) * fvalues[(size_t)i * p + l];
      loss += regLw * m->w[l] * m->w[l];
      m->w[l] = update_wl;
    }
    for (int f = 0; f < k; ++f)  // :194-217
      for (int l = 0; l < p; ++l) {
        double update_Vlf = 0, sum = 0;
        const double Vlf = m->V[(size_t)l * k + f];
        for (int64_t i = 0; i < size; ++i) {
          double fl = fvalues[(size_t)i * p + l];
          double hlf = fl * Q[(size_t)i * k + f] - fl * fl * Vlf;
          update_Vlf += (errors[i] - Vlf * hlf) * hlf;
          sum += hlf * hlf + regLf;
          loss += regLf * (Q[(size_t)i * k + f] * Q[(size_t)i * k + f]);
        }
        update_Vlf = 0 - update_Vlf / sum;
        for (int64_t i = 0; i < size; ++i) {
          errors[i] = errors[i] + (update_Vlf - Vlf) * fvalues[(size_t)i * p + l];
          Q[(size_t)i * k + f] = Q[(size_t)i * k + f] + (update_Vlf - Vlf) * fvalues[(size_t)i * p + l];
        }
        m->V[(size_t)l * k + f] = update_Vlf;
      }
    loss *= 0.05;
  }
  *m->w0 = w0;
  return loss;
}

// ---------------------------------------------------------------------------------------------------
// sparse form
// ---------------------------------------------------------------------------------------------------
// FM.predict with the three non-zero features, same summation order (ascending feature index; the zero
// features of the dense loop add exact zeros).
double oracle_fm_predict(const oracle_fm_problem* pr, const oracle_fm_model* m, int u, int j, int c) {
  const int k = pr->k;
  const int iu = u, ij = pr->num_users + j, ic = pr->num_users + pr->num_items + c;
  const bool has_c = ic < fm_p(pr);
  const double xc = 1.0 / pr->num_context_dims;
  double pred = *m->w0;
  pred += m->w[iu] * 1.0;
  pred += m->w[ij] * 1.0;
  if (has_c) pred += m->w[ic] * xc;
  double sum = 0.0;
  for (int f = 0; f < k; ++f) {
    double sum1 = 0.0, sum2 = 0.0;
    double d = m->V[(size_t)iu * k + f] * 1.0;
    sum1 += d; sum2 += d * d;
    d = m->V[(size_t)ij * k + f] * 1.0;
    sum1 += d; sum2 += d * d;
    if (has_c) {
      d = m->V[(size_t)ic * k + f] * xc;
      sum1 += d; sum2 += d * d;
    }
    sum += sum1 * sum1 - sum2;
  }
  pred += 0.5 * sum;
  return pred;
}

int oracle_fm_predict_batch(const oracle_fm_problem* pr, const oracle_fm_model* m, int64_t n, const int32_t* u,
                            const int32_t* j, const int32_t* c, int32_t bound, double lo, double hi, double* out) {
  for (int64_t i = 0; i < n; i++) {
    double p = oracle_fm_predict(pr, m, u[i], j[i], c[i]);
    if (bound) {
      if (p > hi) p = hi;
      if (p < lo) p = lo;
    }
    out[i] = p;
  }
  return 0;
}

// Pre-pass FM.java:118-146: errors[n] = r - predict, Q[n][f] = sum_i V[i][f] x_n[i]  (Q is [size x k]).
void oracle_fm_prepare(const oracle_fm_problem* pr, const oracle_fm_model* m, double* errors, double* Q) {
  const int k = pr->k, p = fm_p(pr);
  const double xc = 1.0 / pr->num_context_dims;
  for (int64_t n = 0; n < pr->size; ++n) {
    const int iu = pr->u[n], ij = pr->num_users + pr->j[n], ic = pr->num_users + pr->num_items + pr->ctx[n];
    errors[n] = pr->r[n] - oracle_fm_predict(pr, m, pr->u[n], pr->j[n], pr->ctx[n]);
    for (int f = 0; f < k; ++f) {
      double value = 0;
      value += m->V[(size_t)iu * k + f] * 1.0;
      value += m->V[(size_t)ij * k + f] * 1.0;
      if (ic < p) value += m->V[(size_t)ic * k + f] * xc;
      Q[(size_t)n * k + f] = value;
    }
  }
}

// One iteration of the `for (int iter ...)` loop (FM.java:148-219) in sparse form.
// Returns 0.05 * (sum e^2 + regLw*w0^2 + sum_l regLw*w_l^2): the O(size) part of the reference's `loss`
// (the reporting term regLf * sum_{f,l,n} Q[n][f]^2 of FM.java:206 is O(k*p*size) and never read by the
// reference -- FM.buildModel() does not call isConverged(); it is left out here and in the engine).
double oracle_fm_iteration(const oracle_fm_problem* pr, const oracle_fm_model* m, double* errors, double* Q,
                           int closed_den) {
  const int U = pr->num_users, I = pr->num_items, p = fm_p(pr), k = pr->k;
  const int64_t size = pr->size;
  const float regLw = pr->reg_lw, regLf = pr->reg_lf;
  const double xc = 1.0 / pr->num_context_dims;
  // rows of every coordinate, ascending row order
  std::vector<std::vector<int64_t>> rows((size_t)p);
  for (int64_t n = 0; n < size; ++n) {
    rows[(size_t)pr->u[n]].push_back(n);
    rows[(size_t)U + pr->j[n]].push_back(n);
    const int ic = U + I + pr->ctx[n];
    if (ic < p) rows[(size_t)ic].push_back(n);
  }
  auto xval = [&](int l) { return l < U + I ? 1.0 : xc; };

  double loss = 0;
  double w0 = *m->w0;
  double update_w0 = 0;
  for (int64_t i = 0; i < size; ++i) {
    double err = errors[i];
    update_w0 += err - w0;
    loss += err * err;
  }
  update_w0 = update_w0 / ((float)size + regLw);
  update_w0 = 0 - update_w0;
  for (int64_t i = 0; i < size; ++i) errors[i] = errors[i] + update_w0 - w0;
  loss += regLw * w0 * w0;
  w0 = update_w0;
  *m->w0 = w0;

  for (int l = 0; l < p; ++l) {
    const double fl = xval(l), wl = m->w[l];
    double num = 0, den = 0;
    if (closed_den) {
      for (int64_t n : rows[(size_t)l]) num += (errors[n] - wl * fl) * fl;
      den = (double)rows[(size_t)l].size() * (fl * fl) + (double)size * (double)regLw;
    } else {
      size_t pos = 0;
      const std::vector<int64_t>& rl = rows[(size_t)l];
      for (int64_t n = 0; n < size; ++n) {
        if (pos < rl.size() && rl[pos] == n) {
          num += (errors[n] - wl * fl) * fl;
          den += fl * fl + regLw;
          pos++;
        } else {
          den += 0.0 * 0.0 + regLw;
        }
      }
    }
    const double nw = 0 - num / den;
    for (int64_t n : rows[(size_t)l]) errors[n] = errors[n] + (nw - wl) * fl;
    loss += regLw * wl * wl;
    m->w[l] = nw;
  }

  for (int f = 0; f < k; ++f)
    for (int l = 0; l < p; ++l) {
      const double fl = xval(l), Vlf = m->V[(size_t)l * k + f];
      double num = 0, den = 0;
      const std::vector<int64_t>& rl = rows[(size_t)l];
      if (closed_den) {
        for (int64_t n : rl) {
          const double hlf = fl * Q[(size_t)n * k + f] - fl * fl * Vlf;
          num += (errors[n] - Vlf * hlf) * hlf;
          den += hlf * hlf;
        }
        den += (double)size * (double)regLf;
      } else {
        size_t pos = 0;
        for (int64_t n = 0; n < size; ++n) {
          if (pos < rl.size() && rl[pos] == n) {
            const double hlf = fl * Q[(size_t)n * k + f] - fl * fl * Vlf;
            num += (errors[n] - Vlf * hlf) * hlf;
            den += hlf * hlf + regLf;
            pos++;
          } else {
            den += 0.0 + regLf;  // hlf = 0*Q - 0*V = 0
          }
        }
      }
      const double nv = 0 - num / den;
      for (int64_t n : rl) {
        errors[n] = errors[n] + (nv - Vlf) * fl;
        Q[(size_t)n * k + f] = Q[(size_t)n * k + f] + (nv - Vlf) * fl;
      }
      m->V[(size_t)l * k + f] = nv;
    }
  return loss * 0.05;
}

}  // extern "C"
