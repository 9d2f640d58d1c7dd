// schedule.cuh -- K7: the dependency schedule that makes a parallel epoch serial-equivalent.
//
// The reference walks the ratings strictly in CRS order of trainMatrix and updates P[u], Q[j] and the
// biases in place (CAMF_CI.java:80-121), so rating n+1 sees what rating n wrote.  Two ratings commute
// iff they share neither the user nor the item (every cell a rating touches is keyed by its u or its j;
// SURVEY.md Appendix A "conflict sets").  Hence any execution that keeps, for every user and for
// every item, that user's / item's ratings in reference order produces bit-identical P/Q/biases.
//
// level(n) = 1 + max(level(previous rating of u), level(previous rating of j))  -- the longest path
// to n in the conflict DAG.  Ratings of one level are pairwise independent; levels run in order.
#pragma once
#include <cstdint>
#include <vector>

#include "sgd_kernels.cuh"

namespace cars {

struct HostSchedule {
  std::vector<int32_t> u, j, ctx;  // ratings sorted by (level, reference position)
  std::vector<double> r;
  std::vector<int64_t> level_start;  // [num_levels + 1]
  int64_t max_level_size = 0;
};

// Returns false on allocation failure.  ctx may be nullptr.
inline bool build_wavefront_schedule(int32_t num_users, int32_t num_items, int64_t nnz, const int32_t* u,
                                     const int32_t* j, const int32_t* ctx, const double* r, HostSchedule* out) {
  try {
    std::vector<int32_t> last_u((size_t)num_users, 0), last_j((size_t)num_items, 0);
    std::vector<int32_t> level((size_t)nnz);
    int32_t num_levels = 0;
    for (int64_t n = 0; n < nnz; n++) {
      const int32_t a = last_u[u[n]], b = last_j[j[n]];
      const int32_t l = 1 + (a > b ? a : b);
      last_u[u[n]] = l;
      last_j[j[n]] = l;
      level[n] = l;
      if (l > num_levels) num_levels = l;
    }
    // stable counting sort by level
    out->level_start.assign((size_t)num_levels + 1, 0);
    for (int64_t n = 0; n < nnz; n++) out->level_start[level[n]]++;  // level l counted at index l (1-based)
    // level_start[l] currently = size of level l (l >= 1); convert to exclusive starts indexed by l-1
    int64_t run = 0, max_sz = 0;
    for (int32_t l = 1; l <= num_levels; l++) {
      const int64_t sz = out->level_start[l];
      if (sz > max_sz) max_sz = sz;
      out->level_start[l - 1] = run;
      run += sz;
    }
    out->level_start[num_levels] = run;
    out->max_level_size = max_sz;
    std::vector<int64_t> cursor(out->level_start.begin(), out->level_start.end() - 1);
    out->u.resize((size_t)nnz);
    out->j.resize((size_t)nnz);
    out->r.resize((size_t)nnz);
    if (ctx) out->ctx.resize((size_t)nnz);
    for (int64_t n = 0; n < nnz; n++) {
      const int64_t pos = cursor[level[n] - 1]++;
      out->u[pos] = u[n];
      out->j[pos] = j[n];
      out->r[pos] = r[n];
      if (ctx) out->ctx[pos] = ctx[n];
    }
    return true;
  } catch (...) {
    return false;
  }
}

// K7f: level-sorted records for the flagged wavefront (K1f): same order as build_wavefront_schedule,
// each record carrying its position in its user's and its item's chain (counted in reference order).
inline bool build_flagged_schedule(int32_t num_users, int32_t num_items, int64_t nnz, const int32_t* u,
                                   const int32_t* j, const int32_t* ctx, const double* r,
                                   std::vector<RatingRec>* out, int64_t* num_levels_out, int64_t* max_level_out) {
  try {
    std::vector<int32_t> last_u((size_t)num_users, 0), last_j((size_t)num_items, 0);
    std::vector<uint32_t> cu((size_t)num_users, 0u), cj((size_t)num_items, 0u);
    std::vector<int32_t> level((size_t)nnz), ku((size_t)nnz), kj((size_t)nnz);
    int32_t num_levels = 0;
    for (int64_t n = 0; n < nnz; n++) {
      const int32_t a = last_u[u[n]], b = last_j[j[n]];
      const int32_t l = 1 + (a > b ? a : b);
      last_u[u[n]] = l;
      last_j[j[n]] = l;
      level[n] = l;
      ku[n] = (int32_t)cu[u[n]]++;
      kj[n] = (int32_t)cj[j[n]]++;
      if (l > num_levels) num_levels = l;
    }
    std::vector<int64_t> start((size_t)num_levels + 2, 0);
    for (int64_t n = 0; n < nnz; n++) start[level[n] + 1]++;
    int64_t max_sz = 0;
    for (int32_t l = 1; l <= num_levels + 1; l++) {
      if (start[l] > max_sz) max_sz = start[l];
      start[l] += start[l - 1];
    }
    out->resize((size_t)nnz);
    for (int64_t n = 0; n < nnz; n++) {
      RatingRec& x = (*out)[(size_t)start[level[n]]++];
      x.u = u[n]; x.j = j[n]; x.ctx = ctx ? ctx[n] : 0; x.ku = ku[n]; x.kj = kj[n]; x.pad = 0; x.r = r[n];
    }
    *num_levels_out = num_levels;
    *max_level_out = max_sz;
    return true;
  } catch (...) {
    return false;
  }
}

}  // namespace cars
