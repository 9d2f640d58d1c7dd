// multi_gpu.cuh -- ONE handle, N GPUs, one process (cars_desc.num_gpus > 1).
//
// The reference runs in one JVM process (src/carskit/main/CARSKit.java:392-412), so the multi-GPU path a Java caller
// can reach has to live behind the C ABI: cars_create() shards the users by contiguous range over gpu_ids, builds one
// ordinary single-GPU handle per shard (each on its own host thread), opens the NCCL communicators with
// ncclCommInitAll, and cars_epoch() then runs
//     every shard:  snapshot item block, epoch on the shard's ratings, delta = new - old      (its own stream)
//     ncclGroupStart .. ncclAllReduce(delta, sum) on every shard's stream .. ncclGroupEnd      (NVLink / NVSwitch)
//     every shard:  item block <- old + scale_j * sum of deltas
// P, userBias and ucBias rows never leave their GPU; upload / download address the caller's arrays by row offset.
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, a process that has
// torch's NCCL loaded shares it, and a single-GPU caller never needs it.
//
// Included by engine.cu (it needs the handle's internals); everything here is `static`.
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // declarations only

#include <mutex>
#include <thread>

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string err;
};

static const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
      return;
    }
#define CARS_NCCL_SYM(field, sym)                                           \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, #sym)); \
  if (!api.field) api.err = "libnccl lacks " #sym;
    CARS_NCCL_SYM(CommInitAll, ncclCommInitAll)
    CARS_NCCL_SYM(CommDestroy, ncclCommDestroy)
    CARS_NCCL_SYM(AllReduce, ncclAllReduce)
    CARS_NCCL_SYM(GroupStart, ncclGroupStart)
    CARS_NCCL_SYM(GroupEnd, ncclGroupEnd)
    CARS_NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef CARS_NCCL_SYM
  });
  return api;
}

struct MultiGpu {
  std::vector<cars_handle*> shards;
  std::vector<int> ids;
  std::vector<int32_t> lo;  // [N + 1] user range of every shard
  std::vector<ncclComm_t> comms;
  std::vector<double*> d_delta;      // per shard: the item block's delta (all-reduced in place)
  std::vector<double*> d_row_scale;  // per shard: COMBINE_TOUCHED scale per item, else empty
  int64_t item_doubles = 0;
  int combine = CARS_COMBINE_MEAN;
  double last_exchange_ms = 0.0;
  cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;  // around shard 0's all-reduce
};

// contiguous, balanced ranges: the first (U % N) shards get one extra user (same rule as carskit_b200/sharding.py)
static void user_ranges(int32_t U, int N, std::vector<int32_t>* lo) {
  lo->assign((size_t)N + 1, 0);
  const int32_t base = U / N, extra = U % N;
  for (int g = 0; g < N; g++) (*lo)[g + 1] = (*lo)[g] + base + (g < extra ? 1 : 0);
}
static inline int shard_of(int32_t u, int32_t U, int N) {
  const int32_t base = U / N, extra = U % N;
  const int32_t b = extra * (base + 1);
  if (u < b) return u / (base + 1);
  return base > 0 ? extra + (u - b) / base : N - 1;
}

// block <- old + scale[row] * sum : the per-row variant of item_apply_kernel (COMBINE_TOUCHED)
__global__ void __launch_bounds__(256) item_apply_rows_kernel(double* __restrict__ cur, const double* __restrict__ old,
                                                              const double* __restrict__ sum, const double* __restrict__ row_scale,
                                                              int64_t row_len, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    cur[i] = __dadd_rn(old[i], __dmul_rn(row_scale[i / row_len], sum[i]));
}

static void multi_destroy(cars_handle* h);

static int multi_create(const cars_desc* desc, cars_handle** out) {
  const int N = desc->num_gpus;
  const NcclApi& nccl = nccl_api();
  if (!nccl.err.empty()) return fail(nullptr, CARS_E_UNSUPPORTED, "num_gpus = %d needs NCCL: %s", N, nccl.err.c_str());
  if (desc->model == CARS_CAMF_C && desc->mode != CARS_FAST)
    return fail(nullptr, CARS_E_UNSUPPORTED, "CAMF_C in EXACT mode is one chain through condBias; use CARS_FAST with num_gpus > 1");
  if (desc->num_users < N) return fail(nullptr, CARS_E_INVALID, "num_gpus %d exceeds num_users %d", N, desc->num_users);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, CARS_E_NO_DEVICE, "no CUDA device; this engine has no CPU path");
  cars_handle* h = new (std::nothrow) cars_handle();
  if (!h) return fail(nullptr, CARS_E_OOM, "host allocation failed");
  MultiGpu* mg = new (std::nothrow) MultiGpu();
  if (!mg) { delete h; return fail(nullptr, CARS_E_OOM, "host allocation failed"); }
  h->multi = mg;
  h->d = *desc;
  h->nnz = desc->nnz;
  mg->combine = desc->combine;
  auto bail = [&](int code) {
    g_create_error = h->err;
    multi_destroy(h);
    return code;
  };
  for (int g = 0; g < N; g++) {
    const int id = desc->gpu_ids ? desc->gpu_ids[g] : g;
    if (id < 0 || id >= ndev) { fail(h, CARS_E_INVALID, "gpu_ids[%d] = %d out of range (have %d devices)", g, id, ndev); return bail(CARS_E_INVALID); }
    for (int k = 0; k < g; k++)
      if (mg->ids[k] == id) { fail(h, CARS_E_INVALID, "gpu_ids lists device %d twice", id); return bail(CARS_E_INVALID); }
    mg->ids.push_back(id);
  }
  if (desc->combine < CARS_COMBINE_MEAN || desc->combine > CARS_COMBINE_TOUCHED) { fail(h, CARS_E_INVALID, "unknown combine %d", desc->combine); return bail(CARS_E_INVALID); }
  const int32_t U = desc->num_users, I = desc->num_items;
  const int64_t nnz = desc->nnz;
  const bool has_ctx = model_has_ctx(desc->model);
  user_ranges(U, N, &mg->lo);

  // ---- partition the ratings by user range, reference order kept inside every shard -------------------------------
  std::vector<int64_t> count((size_t)N, 0), first((size_t)N, -1), last((size_t)N, -1);
  std::vector<std::vector<uint8_t>> touched;
  try {
    if (desc->combine == CARS_COMBINE_TOUCHED) touched.assign((size_t)N, std::vector<uint8_t>((size_t)I, 0));
  } catch (...) { fail(h, CARS_E_OOM, "host allocation failed"); return bail(CARS_E_OOM); }
  for (int64_t n = 0; n < nnz; n++) {
    const int32_t uu = desc->u[n];
    if ((uint32_t)uu >= (uint32_t)U || (uint32_t)desc->j[n] >= (uint32_t)I) {
      fail(h, CARS_E_INVALID, "rating %lld has an id out of range (u=%d j=%d)", (long long)n, uu, desc->j[n]);
      return bail(CARS_E_INVALID);
    }
    const int g = shard_of(uu, U, N);
    if (first[g] < 0) first[g] = n;
    last[g] = n;
    count[g]++;
    if (!touched.empty()) touched[g][(size_t)desc->j[n]] = 1;
  }
  bool contiguous = true;  // a ratings file sorted by user: every shard is one slice of the caller's arrays
  for (int g = 0; g < N; g++)
    if (count[g] > 0 && last[g] - first[g] + 1 != count[g]) contiguous = false;
  struct Part {
    std::vector<int32_t> u, j, ctx;
    std::vector<double> r;
  };
  std::vector<Part> parts((size_t)N);
  try {
    for (int g = 0; g < N; g++) {
      parts[g].u.resize((size_t)count[g]);
      if (!contiguous) {
        parts[g].j.resize((size_t)count[g]);
        parts[g].r.resize((size_t)count[g]);
        if (has_ctx) parts[g].ctx.resize((size_t)count[g]);
      }
    }
  } catch (...) { fail(h, CARS_E_OOM, "host allocation failed while sharding the ratings"); return bail(CARS_E_OOM); }
  {
    std::vector<int64_t> fill((size_t)N, 0);
    for (int64_t n = 0; n < nnz; n++) {
      const int g = shard_of(desc->u[n], U, N);
      const int64_t k = fill[g]++;
      parts[g].u[(size_t)k] = desc->u[n] - mg->lo[g];
      if (!contiguous) {
        parts[g].j[(size_t)k] = desc->j[n];
        parts[g].r[(size_t)k] = desc->r[n];
        if (has_ctx) parts[g].ctx[(size_t)k] = desc->ctx[n];
      }
    }
  }

  // ---- one ordinary handle per shard, built concurrently -------------------------------------------------------------
  mg->shards.assign((size_t)N, nullptr);
  std::vector<int> rcs((size_t)N, CARS_OK);
  std::vector<std::string> errs((size_t)N);
  std::vector<std::thread> th;
  for (int g = 0; g < N; g++) {
    th.emplace_back([&, g]() {
      cars_desc sd = *desc;
      sd.num_gpus = 0; sd.gpu_ids = nullptr; sd.device = mg->ids[g]; sd.stream = nullptr;
      sd.num_users = mg->lo[g + 1] - mg->lo[g];
      sd.nnz = count[g];
      const int64_t off = contiguous && count[g] > 0 ? first[g] : 0;
      sd.u = parts[g].u.data();
      sd.j = contiguous ? desc->j + off : parts[g].j.data();
      sd.r = contiguous ? desc->r + off : parts[g].r.data();
      sd.ctx = !has_ctx ? nullptr : (contiguous ? desc->ctx + off : parts[g].ctx.data());
      rcs[g] = cars_create(&sd, &mg->shards[g]);
      if (rcs[g] != CARS_OK) errs[g] = cars_last_error(nullptr);  // thread-local message of this thread's failed create
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < N; g++)
    if (rcs[g] != CARS_OK) { fail(h, rcs[g], "shard %d (device %d): %s", g, mg->ids[g], errs[g].c_str()); return bail(rcs[g]); }

  // ---- communicators, delta buffers, per-row scale ----------------------------------------------------------------------
  mg->comms.assign((size_t)N, nullptr);
  ncclResult_t nr = nccl.CommInitAll(mg->comms.data(), N, mg->ids.data());
  if (nr != ncclSuccess) { fail(h, CARS_E_CUDA, "ncclCommInitAll failed: %s", nccl.GetErrorString(nr)); mg->comms.clear(); return bail(CARS_E_CUDA); }
  cars_item_block_doubles(mg->shards[0], &mg->item_doubles);
  mg->d_delta.assign((size_t)N, nullptr);
  mg->d_row_scale.assign((size_t)N, nullptr);
  std::vector<double> row_scale;
  if (!touched.empty()) {
    row_scale.assign((size_t)I, 1.0);
    for (int32_t j = 0; j < I; j++) {
      int c = 0;
      for (int g = 0; g < N; g++) c += touched[g][(size_t)j];
      if (c > 1) row_scale[(size_t)j] = 1.0 / c;
    }
  }
  for (int g = 0; g < N; g++) {
    cars_handle* s = mg->shards[g];
    cudaError_t e = cudaSetDevice(s->device);
    if (e == cudaSuccess) e = s->mem.alloc((void**)&mg->d_delta[g], (size_t)mg->item_doubles * 8);
    if (e == cudaSuccess && !row_scale.empty()) {
      e = s->mem.alloc((void**)&mg->d_row_scale[g], (size_t)I * 8);
      if (e == cudaSuccess) e = cudaMemcpyAsync(mg->d_row_scale[g], row_scale.data(), (size_t)I * 8, cudaMemcpyHostToDevice, s->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) { fail(h, CARS_E_CUDA, "shard %d: %s", g, cudaGetErrorString(e)); return bail(CARS_E_CUDA); }
  }
  {
    cars_handle* s0 = mg->shards[0];
    cudaSetDevice(s0->device);
    cudaEventCreate(&mg->ev_x0);
    cudaEventCreate(&mg->ev_x1);
  }
  h->d.u = h->d.j = h->d.ctx = nullptr; h->d.r = nullptr; h->d.ctx_ptr = h->d.ctx_cond = nullptr; h->d.stream = nullptr;
  h->d.gpu_ids = nullptr; h->d.tuning = nullptr;
  *out = h;
  return CARS_OK;
}

static void multi_destroy(cars_handle* h) {
  MultiGpu* mg = h->multi;
  if (mg) {
    const NcclApi& nccl = nccl_api();
    for (size_t g = 0; g < mg->shards.size(); g++) {
      cars_handle* s = mg->shards[g];
      if (!s) continue;
      cudaSetDevice(s->device);
      cudaStreamSynchronize(s->stream);
      if (g < mg->d_delta.size()) s->mem.free(mg->d_delta[g]);
      if (g < mg->d_row_scale.size()) s->mem.free(mg->d_row_scale[g]);
    }
    if (mg->ev_x0) cudaEventDestroy(mg->ev_x0);
    if (mg->ev_x1) cudaEventDestroy(mg->ev_x1);
    for (ncclComm_t c : mg->comms)
      if (c && nccl.CommDestroy) nccl.CommDestroy(c);
    for (cars_handle* s : mg->shards)
      if (s) cars_destroy(s);
    delete mg;
  }
  delete h;
}

// The caller's arrays, offset to shard g's rows.  Item-side members are shared (full arrays).
static cars_model_arrays shard_arrays(const cars_handle* h, const cars_model_arrays* a, int g) {
  const MultiGpu* mg = h->multi;
  const int64_t lo = mg->lo[g];
  const int64_t F = h->d.num_factors, C = h->d.num_conditions;
  cars_model_arrays s = *a;
  if (a->P) s.P = a->P + lo * F;
  if (a->user_bias) s.user_bias = a->user_bias + lo;
  if (a->uc_bias) s.uc_bias = a->uc_bias + lo * C;
  return s;
}

static int multi_transfer(cars_handle* h, const cars_model_arrays* a, bool to_device) {
  MultiGpu* mg = h->multi;
  if (!a) return fail(h, CARS_E_INVALID, "arrays is NULL");
  const int N = (int)mg->shards.size();
  std::vector<int> rcs((size_t)N, CARS_OK);
  std::vector<std::thread> th;
  for (int g = 0; g < N; g++)
    th.emplace_back([&, g]() {
      cars_model_arrays s = shard_arrays(h, a, g);
      // every shard holds the same item block after an epoch: only shard 0 writes it back
      rcs[g] = transfer(mg->shards[g], &s, to_device, /*skip_item_side=*/!to_device && g > 0);
      if (rcs[g] == CARS_OK && to_device) mg->shards[g]->uploaded = true;
    });
  for (auto& t : th) t.join();
  for (int g = 0; g < N; g++)
    if (rcs[g] != CARS_OK) return fail(h, rcs[g], "shard %d: %s", g, cars_last_error(mg->shards[g]));
  if (to_device) h->uploaded = true;
  return CARS_OK;
}

static int sharded_finish_impl(cars_handle* h, const double* dev_delta, double scale, const double* row_scale, double* loss_out);

static int multi_epoch(cars_handle* h, double lrate, double* loss_out) {
  MultiGpu* mg = h->multi;
  const NcclApi& nccl = nccl_api();
  const int N = (int)mg->shards.size();
  if (!h->uploaded) return fail(h, CARS_E_STATE, "cars_epoch before cars_upload");
  for (int g = 0; g < N; g++) {  // asynchronous: every GPU starts its epoch
    int rc = cars_epoch_sharded_begin(mg->shards[g], lrate, mg->d_delta[g]);
    if (rc != CARS_OK) return fail(h, rc, "shard %d: %s", g, cars_last_error(mg->shards[g]));
  }
  cudaSetDevice(mg->shards[0]->device);
  cudaEventRecord(mg->ev_x0, mg->shards[0]->stream);
  ncclResult_t nr = nccl.GroupStart();
  for (int g = 0; g < N && nr == ncclSuccess; g++)
    nr = nccl.AllReduce(mg->d_delta[g], mg->d_delta[g], (size_t)mg->item_doubles, ncclDouble, ncclSum, mg->comms[g], mg->shards[g]->stream);
  if (nr == ncclSuccess) nr = nccl.GroupEnd();
  else nccl.GroupEnd();
  if (nr != ncclSuccess) return fail(h, CARS_E_CUDA, "ncclAllReduce failed: %s", nccl.GetErrorString(nr));
  cudaSetDevice(mg->shards[0]->device);
  cudaEventRecord(mg->ev_x1, mg->shards[0]->stream);
  const double scale = mg->combine == CARS_COMBINE_SUM ? 1.0 : 1.0 / N;
  double total = 0.0, max_ms = 0.0;
  int rc_all = CARS_OK;
  for (int g = 0; g < N; g++) {
    double loss = 0.0;
    int rc = sharded_finish_impl(mg->shards[g], mg->d_delta[g], scale, mg->combine == CARS_COMBINE_TOUCHED ? mg->d_row_scale[g] : nullptr, &loss);
    if (rc != CARS_OK && rc_all == CARS_OK) rc_all = fail(h, rc, "shard %d: %s", g, cars_last_error(mg->shards[g]));
    total += loss;  // the reference's loss is a sum over ratings (CAMF_CI.java:91-124)
    if (mg->shards[g]->st.last_epoch_ms > max_ms) max_ms = mg->shards[g]->st.last_epoch_ms;
  }
  if (rc_all != CARS_OK) return rc_all;
  float xms = 0.f;
  if (cudaEventElapsedTime(&xms, mg->ev_x0, mg->ev_x1) == cudaSuccess) mg->last_exchange_ms = xms;
  h->st.last_epoch_ms = max_ms;
  if (loss_out) *loss_out = total;
  return CARS_OK;
}

// Queries (u, ...) routed to the shard that owns the user; `run` is called per shard with the positions it got.
template <typename Fn>
static int multi_by_user(cars_handle* h, int64_t n, const int32_t* u, Fn run) {
  MultiGpu* mg = h->multi;
  const int N = (int)mg->shards.size();
  const int32_t U = h->d.num_users;
  std::vector<std::vector<int64_t>> pos((size_t)N);
  for (int64_t i = 0; i < n; i++) {
    if ((uint32_t)u[i] >= (uint32_t)U) return fail(h, CARS_E_INVALID, "query %lld has a user id out of range", (long long)i);
    pos[(size_t)shard_of(u[i], U, N)].push_back(i);
  }
  for (int g = 0; g < N; g++) {
    if (pos[g].empty()) continue;
    int rc = run(g, pos[g]);
    if (rc != CARS_OK) return fail(h, rc, "shard %d: %s", g, cars_last_error(mg->shards[g]));
  }
  return CARS_OK;
}

static int multi_predict(cars_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx, int32_t bound,
                         double min_rate, double max_rate, double* out) {
  MultiGpu* mg = h->multi;
  return multi_by_user(h, n, u, [&](int g, const std::vector<int64_t>& pos) {
    const size_t m = pos.size();
    std::vector<int32_t> su(m), sj(m), sc(ctx ? m : 0);
    std::vector<double> so(m);
    for (size_t k = 0; k < m; k++) {
      su[k] = u[pos[k]] - mg->lo[g];
      sj[k] = j[pos[k]];
      if (ctx) sc[k] = ctx[pos[k]];
    }
    int rc = cars_predict(mg->shards[g], (int64_t)m, su.data(), sj.data(), ctx ? sc.data() : nullptr, bound, min_rate, max_rate, so.data());
    for (size_t k = 0; k < m && rc == CARS_OK; k++) out[pos[k]] = so[k];
    return rc;
  });
}

static int multi_rank_topn(cars_handle* h, int64_t nq, const int32_t* qu, const int32_t* qc, int32_t num_cand, const int32_t* cand,
                           const int64_t* rated_ptr, const int32_t* rated_items, double bin_thold, int32_t num_recs,
                           int32_t* out_items, double* out_scores, int32_t* out_count, int32_t* out_kept) {
  MultiGpu* mg = h->multi;
  return multi_by_user(h, nq, qu, [&](int g, const std::vector<int64_t>& pos) {
    const size_t m = pos.size();
    std::vector<int32_t> su(m), sc(qc ? m : 0), items(m * (size_t)num_recs), count(m), kept(m), ritems;
    std::vector<double> scores(m * (size_t)num_recs);
    std::vector<int64_t> rptr(rated_ptr ? m + 1 : 0, 0);
    for (size_t k = 0; k < m; k++) {
      su[k] = qu[pos[k]] - mg->lo[g];
      if (qc) sc[k] = qc[pos[k]];
      if (rated_ptr) {
        for (int64_t t = rated_ptr[pos[k]]; t < rated_ptr[pos[k] + 1]; t++) ritems.push_back(rated_items[t]);
        rptr[k + 1] = (int64_t)ritems.size();
      }
    }
    int rc = cars_rank_topn(mg->shards[g], (int64_t)m, su.data(), qc ? sc.data() : nullptr, num_cand, cand,
                            rated_ptr ? rptr.data() : nullptr, ritems.empty() ? nullptr : ritems.data(), bin_thold, num_recs,
                            items.data(), scores.data(), count.data(), kept.data());
    for (size_t k = 0; k < m && rc == CARS_OK; k++) {
      memcpy(out_items + pos[k] * num_recs, items.data() + k * num_recs, (size_t)num_recs * 4);
      memcpy(out_scores + pos[k] * num_recs, scores.data() + k * num_recs, (size_t)num_recs * 8);
      out_count[pos[k]] = count[k];
      out_kept[pos[k]] = kept[k];
    }
    return rc;
  });
}

static void multi_stats(const cars_handle* h, cars_stats* out) {
  const MultiGpu* mg = h->multi;
  cars_stats t{};
  for (size_t g = 0; g < mg->shards.size(); g++) {
    const cars_stats& s = mg->shards[g]->st;
    t.nnz += s.nnz;
    t.kernel_launches += s.kernel_launches;
    t.h2d_bytes += s.h2d_bytes;
    t.d2h_bytes += s.d2h_bytes;
    if (s.num_levels > t.num_levels) t.num_levels = s.num_levels;
    if (s.max_level_size > t.max_level_size) t.max_level_size = s.max_level_size;
    if (s.schedule_ms > t.schedule_ms) t.schedule_ms = s.schedule_ms;
    if (s.last_epoch_ms > t.last_epoch_ms) t.last_epoch_ms = s.last_epoch_ms;
    if (s.max_item_degree > t.max_item_degree) t.max_item_degree = s.max_item_degree;
    if (g == 0) { t.grid_ctas = s.grid_ctas; t.block_threads = s.block_threads; t.sm_count = s.sm_count;
                  t.fast_min_item_scale = s.fast_min_item_scale; t.fast_min_cond_scale = s.fast_min_cond_scale; t.fast_hot_rows = s.fast_hot_rows; }
  }
  t.num_gpus = (int32_t)mg->shards.size();
  t.exchange_ms = mg->last_exchange_ms;
  *out = t;
}
