// engine.cu -- the C ABI of include/carskit_b200.h on top of the sm_100a kernels.
//
// What lives here: handle lifetime, host<->device marshalling of the flattened Java containers,
// the dependency schedule (wavefront levels) and kernel dispatch.  No CPU implementation of the
// update exists in this library: without a CUDA sm_100 device cars_create() fails.
#include "../../include/carskit_b200.h"
#include "schedule.cuh"
#include "schedule_gpu.cuh"
#include "sgd_kernels.cuh"
#include "fast_kernels.cuh"
#include "tagged_kernels.cuh"
#include "fast_schedule.cuh"
#include "staged_copy.cuh"
#include "rank_kernels.cuh"
#include "tuning.h"
#include "dev_mem.cuh"

#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace cars;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_create_error = "";

struct LaunchPlan {
  const void* fn = nullptr;
  int lpr = 0, v = 0, threads = 0;
  bool bulk = false;  // FAST: Q-row steps through TMA add-reduce of a staged row (fast_kernels.cuh, BULK)
};

struct MultiGpu;  // multi_gpu.cuh: one handle driving N GPUs (cars_desc.num_gpus > 1)

struct cars_handle {
  MultiGpu* multi = nullptr;  // non-null: this handle is a front for N per-GPU handles; the members below are unused
  cars_desc d;  // scalars only; pointer members are nulled after create
  LaunchPlan plan;  // the SGD kernel chosen at create (shape, shared memory and grid belong together)
  cars::Tuning tune;
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev_beg = nullptr, ev_end = nullptr;
  StagedCopier copier;  // host <-> device transfers of the caller's (pageable or pinned) arrays
  DevMem mem;           // device allocations: the device's stream-ordered pool (dev_mem.cuh)

  // model
  DeviceModel m{};
  int Dmax = 0;
  int32_t* d_ctx_tab = nullptr;
  int32_t* d_empty_cond = nullptr;  // CAMF_ICS
  int32_t *d_ui_ptr = nullptr, *d_ui_items = nullptr;  // SVD++: userItemsCache in CSR form
  bool uploaded = false;

  // ratings in schedule order
  int64_t nnz = 0;
  int32_t *d_u = nullptr, *d_j = nullptr, *d_ctx = nullptr;
  double* d_r = nullptr;
  int64_t* d_level_start = nullptr;
  int64_t num_levels = 0, max_level_size = 0;
  bool serial = false;    // CAMF_C exact: reference order, one warp
  bool dataflow = false;  // default schedule: reference order + per-user / per-item completion counters
  bool flagged = false;   // level order + completion counters (no barriers)
  bool fast = false;      // FAST mode: user-sorted chunks, item side by reductions (fast_kernels.cuh)
  // K1t: the flagged schedule on tagged rows (tagged_kernels.cuh).  The model then lives in TWO layouts; each knows
  // whether it is current, and whoever needs the other one converts (a streaming pass over the model: 0.2 ms at config 3)
  bool tagged = false, std_valid = true, tagged_valid = false;
  TaggedLayout tl;
  TaggedModel tm{};
  double* d_item_scale = nullptr;  // FAST: per-item step damping [num_items]
  double* d_cond_scale = nullptr;  // FAST, CAMF_C: per-condition step damping [C]
  bool damp_items = false, damp_conds = false;
  signed char* d_hot_slot = nullptr;  // FAST: hot-row slot of every item (-1 = none) [num_items]
  int32_t* d_hot_items = nullptr;     // FAST: item of every hot slot
  int num_hot = 0, hot_flush = 16, hot_stride = 0, bulk_offset = 0;
  RatingRec* d_rec = nullptr;
  int64_t* d_chunk_start = nullptr;
  double* d_chunk_loss = nullptr;
  unsigned* d_flags = nullptr;  // [counter (pad to 64 u32) | done_j | done_u]
  size_t flags_words = 0;
  int64_t num_chunks = 0;
  int loss_blocks = 0;
  double* d_item_old = nullptr;  // snapshot of the item block (multi-GPU exchange), allocated on first use
  bool sharded_pending = false;

  // kernel plumbing
  unsigned* d_barrier = nullptr;
  double* d_partial = nullptr;
  double* d_loss = nullptr;
  double* h_loss = nullptr;  // pinned
  int grid = 0, block = 0, sm_count = 0;
  size_t smem = 0;
  bool epoch_pending = false;

  cars_stats st{};
};

static int fail(cars_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(h, _e == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA, "%s failed: %s",  \
                  #expr, cudaGetErrorString(_e));                                                  \
  } while (0)

template <typename T>
static cudaError_t dev_alloc_on(const DevMem& mem, T** p, size_t n) {
  return mem.alloc(reinterpret_cast<void**>(p), (n ? n : 1) * sizeof(T));
}
#define dev_alloc(ptr, n) dev_alloc_on(h->mem, ptr, n)  /* every call site has the handle `h` in scope */

constexpr long long kTaggedDefault = 0;  // K1t is opt-in (tuning "tagged=1") until it has earned the default
static bool model_has_ctx(int model);
static int multi_create(const cars_desc* desc, cars_handle** out);
static void multi_destroy(cars_handle* h);
static int multi_transfer(cars_handle* h, const cars_model_arrays* a, bool to_device);
static int multi_epoch(cars_handle* h, double lrate, double* loss_out);
static int multi_predict(cars_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx, int32_t bound,
                         double min_rate, double max_rate, double* out);
static int multi_rank_topn(cars_handle* h, int64_t nq, const int32_t* qu, const int32_t* qc, int32_t num_cand, const int32_t* cand,
                           const int64_t* rated_ptr, const int32_t* rated_items, double bin_thold, int32_t num_recs,
                           int32_t* out_items, double* out_scores, int32_t* out_count, int32_t* out_kept);
static void multi_stats(const cars_handle* h, cars_stats* out);

// ------------------------------------------------------------------------------------------------
// kernel dispatch: model x (lanes per rating, chunks per lane) chosen from num_factors
// ------------------------------------------------------------------------------------------------
// One shape for the barrier (K1) and dataflow (K1d) kernels: 256 threads x 3 CTAs per SM, measured best on B200
// (profiles/r1/bench_wavefront_v2_256x3.json).
#define CARS_SHAPE_TABLE(KERNEL, MODEL, THREADS, MINB)                                                          \
  LaunchPlan p;                                                                                                 \
  p.threads = THREADS;                                                                                          \
  if (Fp <= 16) { p.fn = (const void*)KERNEL<MODEL, 8, 1, THREADS, MINB>; p.lpr = 8; p.v = 1; }                 \
  else if (Fp <= 32) { p.fn = (const void*)KERNEL<MODEL, 8, 2, THREADS, MINB>; p.lpr = 8; p.v = 2; }            \
  else if (Fp <= 64) { p.fn = (const void*)KERNEL<MODEL, 8, 4, THREADS, MINB>; p.lpr = 8; p.v = 4; }            \
  else if (Fp <= 128) { p.fn = (const void*)KERNEL<MODEL, 16, 4, THREADS, MINB>; p.lpr = 16; p.v = 4; }         \
  else if (Fp <= 256) { p.fn = (const void*)KERNEL<MODEL, 32, 4, THREADS, MINB>; p.lpr = 32; p.v = 4; }         \
  else if (Fp <= 512) { p.fn = (const void*)KERNEL<MODEL, 32, 8, THREADS, MINB>; p.lpr = 32; p.v = 8; }         \
  return p;

template <int MODEL>
static LaunchPlan pick_wavefront(int Fp) { CARS_SHAPE_TABLE(sgd_wavefront_kernel, MODEL, 256, 3) }
template <int MODEL>
static LaunchPlan pick_dataflow(int Fp) { CARS_SHAPE_TABLE(sgd_dataflow_kernel, MODEL, 256, 3) }

static LaunchPlan pick_plan(int model, int Fp) {
  switch (model) {
    case CARS_PMF: return pick_wavefront<M_PMF>(Fp);
    case CARS_BIASEDMF: return pick_wavefront<M_BIASEDMF>(Fp);
    case CARS_CAMF_C: return LaunchPlan{};  // every rating touches condBias: serial kernel only
    case CARS_CAMF_CI: return pick_wavefront<M_CAMF_CI>(Fp);
    case CARS_CAMF_CU: return pick_wavefront<M_CAMF_CU>(Fp);
    case CARS_CAMF_CUCI: return pick_wavefront<M_CAMF_CUCI>(Fp);
  }
  return LaunchPlan{};
}

// K1t instances: <row lengths in lines> that cover F <= 64 with up to 40 condition cells per row (config 3: 5 + 7 lines)
// and F <= 128 with a scalar bias; anything longer keeps the counter kernel.
template <int MODEL>
static LaunchPlan pick_tagged(int lp, int lq, int ctas) {
  LaunchPlan p;
  p.threads = 256; p.lpr = 8; p.v = 0;
  if (lp <= 5 && lq <= 7) p.fn = ctas == 2 ? (const void*)sgd_tagged_kernel<MODEL, 5, 7, 256, 2> : (const void*)sgd_tagged_kernel<MODEL, 5, 7, 256, 3>;
  else if (lp <= 9 && lq <= 9) p.fn = (const void*)sgd_tagged_kernel<MODEL, 9, 9, 256, 2>;
  return p;
}
static LaunchPlan pick_tagged_plan(int model, int lp, int lq, int ctas) {
  switch (model) {
    case CARS_PMF: return pick_tagged<M_PMF>(lp, lq, ctas);
    case CARS_BIASEDMF: return pick_tagged<M_BIASEDMF>(lp, lq, ctas);
    case CARS_CAMF_CI: return pick_tagged<M_CAMF_CI>(lp, lq, ctas);
    case CARS_CAMF_CU: return pick_tagged<M_CAMF_CU>(lp, lq, ctas);
    case CARS_CAMF_CUCI: return pick_tagged<M_CAMF_CUCI>(lp, lq, ctas);
  }
  return LaunchPlan{};
}

static LaunchPlan pick_dataflow_plan(int model, int Fp) {
  switch (model) {
    case CARS_PMF: return pick_dataflow<M_PMF>(Fp);
    case CARS_BIASEDMF: return pick_dataflow<M_BIASEDMF>(Fp);
    case CARS_CAMF_CI: return pick_dataflow<M_CAMF_CI>(Fp);
    case CARS_CAMF_CU: return pick_dataflow<M_CAMF_CU>(Fp);
    case CARS_CAMF_CUCI: return pick_dataflow<M_CAMF_CUCI>(Fp);
  }
  return LaunchPlan{};
}

// FAST (hogwild) kernel.  Default (shape 0): the Q-row step of a rating is staged in shared memory and added to Q[j] by
// ONE TMA add-reduce (BULK, UBLKRED.G.S.ADD.F64) for Fp >= 64 -- 27.5 ms per 100 M ratings against 34.4 ms with 64 scalar
// REDG per rating on the same box (profiles/r2/fast_bulk_reduce.txt); rows shorter than 64 factors keep the scalar
// reductions.  F = 64: 8 lanes per rating, 256-bit row accesses, F compiled in, 256 threads x 2 CTAs per SM; F = 128: 16
// lanes.  shape 5 = the scalar-REDG kernels (A/B), shape 1 / 2 / 4 = other launch shapes measured and rejected.
template <int MODEL, bool BULK>
static LaunchPlan pick_fast_generic(int Fp) {
  LaunchPlan p;
  p.threads = 256;
  p.bulk = BULK;
  if (Fp <= 16) { p.fn = (const void*)sgd_fast_kernel<MODEL, 8, 1, 256, 2, false, 0, BULK>; p.lpr = 8; p.v = 1; }
  else if (Fp <= 32) { p.fn = (const void*)sgd_fast_kernel<MODEL, 8, 2, 256, 2, false, 0, BULK>; p.lpr = 8; p.v = 2; }
  else if (Fp <= 64) { p.fn = (const void*)sgd_fast_kernel<MODEL, 8, 4, 256, 2, false, 0, BULK>; p.lpr = 8; p.v = 4; }
  else if (Fp <= 128) { p.fn = (const void*)sgd_fast_kernel<MODEL, 16, 4, 256, 2, false, 0, BULK>; p.lpr = 16; p.v = 4; }
  else if (Fp <= 256) { p.fn = (const void*)sgd_fast_kernel<MODEL, 32, 4, 256, 2, false, 0, BULK>; p.lpr = 32; p.v = 4; }
  else if (Fp <= 512) { p.fn = (const void*)sgd_fast_kernel<MODEL, 32, 8, 256, 2, false, 0, BULK>; p.lpr = 32; p.v = 8; }
  return p;
}
template <int MODEL>
static LaunchPlan pick_fast(int Fp, int F, int shape) {
  LaunchPlan p;
  p.threads = 256;
  const bool bulk = shape != 5 && shape != 1 && shape != 2;
  if (Fp > 16 && Fp <= 32 && shape == 7) {  // 4 lanes per rating up to 32 factors: 15.0 against 18.2 ms per 100 M ratings (F = 32)
    p.fn = (const void*)sgd_fast_kernel<MODEL, 4, 4, 256, 2, false, 0, false>; p.lpr = 4; p.v = 4; return p;
  }
  if (Fp <= 16 && shape == 6) {  // short rows, 4 lanes per rating (8 ratings per warp instruction instead of 4)
    p.fn = (const void*)sgd_fast_kernel<MODEL, 4, 2, 256, 2, false, 0, false>; p.lpr = 4; p.v = 2; return p;
  }
  if (F == 64 && (shape == 0 || shape == 3)) { p.fn = (const void*)sgd_fast_kernel<MODEL, 8, 4, 256, 2, true, 64, true>; p.lpr = 8; p.v = 4; p.bulk = true; return p; }
  if (F == 64 && shape == 4) { p.fn = (const void*)sgd_fast_kernel<MODEL, 8, 4, 256, 3, true, 64, true>; p.lpr = 8; p.v = 4; p.bulk = true; return p; }
  if (F == 64 && shape == 5) { p.fn = (const void*)sgd_fast_kernel<MODEL, 8, 4, 256, 2, true, 64>; p.lpr = 8; p.v = 4; return p; }
  if (F == 64 && shape == 1) { p.fn = (const void*)sgd_fast_kernel<MODEL, 8, 4, 256, 3, true, 64>; p.lpr = 8; p.v = 4; return p; }
  if (F == 64 && shape == 2) { p.fn = (const void*)sgd_fast_kernel<MODEL, 4, 8, 256, 2, true, 64>; p.lpr = 4; p.v = 8; return p; }
  if (F == 128 && bulk) { p.fn = (const void*)sgd_fast_kernel<MODEL, 16, 4, 256, 2, true, 128, true>; p.lpr = 16; p.v = 4; p.bulk = true; return p; }
  if (F == 128) { p.fn = (const void*)sgd_fast_kernel<MODEL, 16, 4, 256, 2, true, 128>; p.lpr = 16; p.v = 4; return p; }
  if (bulk && Fp >= 64) return pick_fast_generic<MODEL, true>(Fp);  // F = 32: one 256-byte TMA reduce per rating LOSES to 32 REDG (24.4 vs 18.2 ms)
  return pick_fast_generic<MODEL, false>(Fp);
}
static LaunchPlan pick_fast_plan(int model, int Fp, int F, int shape) {
  switch (model) {
    case CARS_PMF: return pick_fast<M_PMF>(Fp, F, shape);
    case CARS_BIASEDMF: return pick_fast<M_BIASEDMF>(Fp, F, shape);
    case CARS_CAMF_C: return pick_fast<M_CAMF_C>(Fp, F, shape);
    case CARS_CAMF_CI: return pick_fast<M_CAMF_CI>(Fp, F, shape);
    case CARS_CAMF_CU: return pick_fast<M_CAMF_CU>(Fp, F, shape);
    case CARS_CAMF_CUCI: return pick_fast<M_CAMF_CUCI>(Fp, F, shape);
  }
  return LaunchPlan{};
}

template <int MODEL>
static LaunchPlan pick_flagged_generic(int Fp) { CARS_SHAPE_TABLE(sgd_flagged_kernel, MODEL, 256, 3) }

template <int MODEL>
static LaunchPlan pick_flagged(int Fp, int variant, int F = 0) {
  switch (variant) {
    case 3: {  // 4 lanes per rating (16 factors per lane): more ratings in flight per SM at F <= 64
      LaunchPlan p;
      if (Fp > 32 && Fp <= 64) {
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 4, 8, 256, 2>; p.lpr = 4; p.v = 8;
        return p;
      }
      if (Fp <= 16) {  // the reference's default F = 10: 4 lanes per rating, 27.5 against 30.6 ms per 100 M ratings with 8
                       // (CAMF_CI; BiasedMF 26.7 / 27.4; profiles/r2/config5_shapes.txt)
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 4, 2, 256, 3>; p.lpr = 4; p.v = 2;
        return p;
      }
      if (Fp <= 32) {  // F = 32: 31.3 against 34.7 ms per 100 M ratings with 8 lanes
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 4, 4, 256, 3>; p.lpr = 4; p.v = 4;
        return p;
      }
      return pick_flagged_generic<MODEL>(Fp);
    }
    case 11:  // A/B of the default for short rows: the generic table (8 lanes per rating)
      return pick_flagged_generic<MODEL>(Fp);
    case 9: {  // A/B of the default at F = 128: 16 lanes per rating x 3 CTAs per SM (48 ratings in flight per SM)
      LaunchPlan p;
      if (F == 128) {
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 16, 4, 256, 3, true, 128>; p.lpr = 16; p.v = 4;
        return p;
      }
      return pick_flagged<MODEL>(Fp, 8, F);
    }
    case 8: {  // as 7 with the factor count a compile-time constant (F = 64 / F = 128 exactly)
      LaunchPlan p;
      if (F == 64) {
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 4, 8, 256, 2, true, 64>; p.lpr = 4; p.v = 8;
        return p;
      }
      if (F == 128) {  // 8 lanes per rating (16 factors per lane), 2 CTAs per SM: 64 ratings in flight per SM -- 118.4 ms
                       // against 131.1 ms per epoch of config 5's shard for 16 lanes x 3 CTAs (profiles/r2/config5_shapes.txt)
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 8, 8, 256, 2, true, 128>; p.lpr = 8; p.v = 8;
        return p;
      }
      return pick_flagged<MODEL>(Fp, 7, F);
    }
    case 7: {  // as 3, rows moved with 256-bit loads / stores (rows 32-byte aligned: Fp % 4 == 0)
      LaunchPlan p;
      if (Fp > 32 && Fp <= 64 && Fp % 4 == 0) {
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 4, 8, 256, 2, true>; p.lpr = 4; p.v = 8;
        return p;
      }
      if (Fp > 64 && Fp <= 128 && Fp % 4 == 0) {
        p.threads = 256; p.fn = (const void*)sgd_flagged_kernel<MODEL, 16, 4, 256, 3, true>; p.lpr = 16; p.v = 4;
        return p;
      }
      return pick_flagged<MODEL>(Fp, 3);
    }
    default: return pick_flagged_generic<MODEL>(Fp);
  }
}

// v = 8 (default): F = 64 / 128 compiled in + 256-bit row accesses; else 7 (256-bit where Fp % 4 == 0); else 3
static LaunchPlan pick_flagged_plan(int model, int Fp, int F, int v) {
  switch (model) {
    case CARS_PMF: return pick_flagged<M_PMF>(Fp, v, F);
    case CARS_BIASEDMF: return pick_flagged<M_BIASEDMF>(Fp, v, F);
    case CARS_CAMF_CI: return pick_flagged<M_CAMF_CI>(Fp, v, F);
    case CARS_CAMF_CU: return pick_flagged<M_CAMF_CU>(Fp, v, F);
    case CARS_CAMF_CUCI: return pick_flagged<M_CAMF_CUCI>(Fp, v, F);
  }
  return LaunchPlan{};
}

template <int MODEL>
static const void* pick_serial_m(int Fp) {
  if (Fp <= 64) return (const void*)sgd_serial_kernel<MODEL, 1>;
  if (Fp <= 128) return (const void*)sgd_serial_kernel<MODEL, 2>;
  if (Fp <= 256) return (const void*)sgd_serial_kernel<MODEL, 4>;
  if (Fp <= 512) return (const void*)sgd_serial_kernel<MODEL, 8>;
  return nullptr;
}
static const void* pick_serial(int model, int Fp) {
  if (model == CARS_SVDPP) {
    if (Fp <= 64) return (const void*)sgd_serial_svdpp_kernel<1>;
    if (Fp <= 128) return (const void*)sgd_serial_svdpp_kernel<2>;
    if (Fp <= 256) return (const void*)sgd_serial_svdpp_kernel<4>;
    if (Fp <= 512) return (const void*)sgd_serial_svdpp_kernel<8>;
    return nullptr;
  }
  if (model == CARS_CAMF_LCS || model == CARS_CAMF_MCS) {
#define CARS_SIM_PICK(K)                                               \
  if (Fp <= 64) return (const void*)sgd_serial_sim_kernel<1, K>;       \
  if (Fp <= 128) return (const void*)sgd_serial_sim_kernel<2, K>;      \
  if (Fp <= 256) return (const void*)sgd_serial_sim_kernel<4, K>;      \
  if (Fp <= 512) return (const void*)sgd_serial_sim_kernel<8, K>;      \
  return nullptr;
    if (model == CARS_CAMF_LCS) { CARS_SIM_PICK(1) }
    CARS_SIM_PICK(2)
#undef CARS_SIM_PICK
  }
  if (model == CARS_CAMF_ICS) {
    if (Fp <= 64) return (const void*)sgd_serial_ics_kernel<1>;
    if (Fp <= 128) return (const void*)sgd_serial_ics_kernel<2>;
    if (Fp <= 256) return (const void*)sgd_serial_ics_kernel<4>;
    if (Fp <= 512) return (const void*)sgd_serial_ics_kernel<8>;
    return nullptr;
  }
  switch (model) {
    case CARS_PMF: return pick_serial_m<M_PMF>(Fp);
    case CARS_BIASEDMF: return pick_serial_m<M_BIASEDMF>(Fp);
    case CARS_CAMF_C: return pick_serial_m<M_CAMF_C>(Fp);
    case CARS_CAMF_CI: return pick_serial_m<M_CAMF_CI>(Fp);
    case CARS_CAMF_CU: return pick_serial_m<M_CAMF_CU>(Fp);
    case CARS_CAMF_CUCI: return pick_serial_m<M_CAMF_CUCI>(Fp);
  }
  return nullptr;
}

static bool model_has_ctx(int model) {
  return model == CARS_CAMF_C || model == CARS_CAMF_CI || model == CARS_CAMF_CU || model == CARS_CAMF_CUCI || model == CARS_CAMF_ICS ||
         model == CARS_CAMF_LCS || model == CARS_CAMF_MCS;
}

// ------------------------------------------------------------------------------------------------
// cars_create
// ------------------------------------------------------------------------------------------------
static int validate(const cars_desc* d) {
  if (!d) return fail(nullptr, CARS_E_INVALID, "desc is NULL");
  if (d->abi_version != CARS_ABI_VERSION)
    return fail(nullptr, CARS_E_INVALID, "abi_version %d != %d", d->abi_version, CARS_ABI_VERSION);
  if (d->model < CARS_PMF || d->model > CARS_SVDPP) return fail(nullptr, CARS_E_INVALID, "unknown model %d", d->model);
  if (d->model == CARS_SVDPP && (d->mode == CARS_FAST || d->num_gpus > 1))
    return fail(nullptr, CARS_E_UNSUPPORTED, "SVD++ is built for EXACT mode on one GPU (every rating rewrites Y of all the user's items: one chain)");
  if (d->model == CARS_CAMF_LCS && (d->num_context_factors <= 0 || d->num_context_factors > 4096))
    return fail(nullptr, CARS_E_INVALID, "CAMF_LCS needs num_context_factors in 1..4096 (the `-f` option, CAMF_LCS.java:38)");
  if (d->model == CARS_CAMF_MCS && d->num_context_dims <= 0)
    return fail(nullptr, CARS_E_INVALID, "CAMF_MCS needs num_context_dims > 0 (rateDao.numContextDims(), CAMF_MCS.java:44)");
  if (d->model == CARS_CAMF_ICS || d->model == CARS_CAMF_LCS || d->model == CARS_CAMF_MCS) {
    if (d->mode == CARS_FAST || d->num_gpus > 1)
      return fail(nullptr, CARS_E_UNSUPPORTED, "CAMF_ICS / LCS / MCS are built for EXACT mode on one GPU (one chain through the shared similarity cells)");
    if (!d->empty_conditions || d->num_empty_conditions <= 0)
      return fail(nullptr, CARS_E_INVALID, "CAMF_ICS / LCS / MCS need empty_conditions (rateDao.getEmptyContextConditions())");
  }
  if (d->model == CARS_FM)
    return fail(nullptr, CARS_E_INVALID, "FM is an ALS model with its own entry points: use cars_fm_create");
  if (d->mode != CARS_EXACT && d->mode != CARS_FAST) return fail(nullptr, CARS_E_INVALID, "unknown mode %d", d->mode);
  if (d->nnz > 0x7fffffffll)  // rating indices are 32-bit in the record stream, the chains and the frontiers
    return fail(nullptr, CARS_E_UNSUPPORTED, "nnz %lld exceeds 2^31 - 1 ratings per handle; shard the users over more handles",
                (long long)d->nnz);
  if (d->num_users <= 0 || d->num_items <= 0) return fail(nullptr, CARS_E_INVALID, "num_users/num_items must be > 0");
  if (d->num_factors <= 0 || d->num_factors > 512)
    return fail(nullptr, CARS_E_UNSUPPORTED, "num_factors %d outside 1..512", d->num_factors);
  if (d->nnz < 0) return fail(nullptr, CARS_E_INVALID, "nnz < 0");
  if (d->nnz > 0 && (!d->u || !d->j || !d->r)) return fail(nullptr, CARS_E_INVALID, "u/j/r must not be NULL");
  if (model_has_ctx(d->model)) {
    if (d->num_conditions <= 0 || d->num_contexts <= 0)
      return fail(nullptr, CARS_E_INVALID, "context model needs num_conditions/num_contexts > 0");
    if (!d->ctx_ptr || !d->ctx_cond || (d->nnz > 0 && !d->ctx))
      return fail(nullptr, CARS_E_INVALID, "context model needs ctx, ctx_ptr, ctx_cond");
  }
  return CARS_OK;
}

extern "C" int cars_create(const cars_desc* desc, cars_handle** out) {
  const auto t_create = std::chrono::steady_clock::now();
  if (!out) return fail(nullptr, CARS_E_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate(desc);
  if (rc) return rc;
  if (desc->num_gpus > 1) return multi_create(desc, out);

  const bool create_trace = desc->tuning && strstr(desc->tuning, "sched_trace=1") != nullptr;
  auto clap = [&](const char* what) {
    if (create_trace)
      fprintf(stderr, "[cars create] +%8.2f ms  %s\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_create).count(), what);
  };
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, CARS_E_NO_DEVICE, "no CUDA device (%s); this engine has no CPU path",
                ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (desc->device < 0 || desc->device >= ndev)
    return fail(nullptr, CARS_E_INVALID, "device %d out of range (have %d)", desc->device, ndev);
  DeviceFacts prop;
  CUDA_TRY(nullptr, device_facts(desc->device, &prop));
  clap("cudaGetDeviceCount + device attributes");
  if (prop.major != 10)
    return fail(nullptr, CARS_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                desc->device, prop.major, prop.minor);

  cars_handle* h = new (std::nothrow) cars_handle();
  if (!h) return fail(nullptr, CARS_E_OOM, "host allocation failed");
  // from here on failures must free h
  auto bail = [&](int code) {
    g_create_error = h->err;
    cars_destroy(h);
    return code;
  };
#define CUDA_TRY_H(expr)                                                                                \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      fail(h, CARS_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));                             \
      return bail(_e == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA);                          \
    }                                                                                                   \
  } while (0)

  h->d = *desc;
  h->tune = Tuning(desc->tuning);
  h->device = desc->device;
  h->sm_count = prop.sm_count;
  CUDA_TRY_H(cudaSetDevice(h->device));
  if (desc->stream) {
    h->stream = (cudaStream_t)desc->stream;
  } else {
    CUDA_TRY_H(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  CUDA_TRY_H(cudaEventCreate(&h->ev_beg));
  CUDA_TRY_H(cudaEventCreate(&h->ev_end));
  h->mem.stream = h->stream;
  h->mem.pooled = h->tune.get_ll("pool", 1) != 0;
  clap("stream + events");
  if (h->mem.pooled) CUDA_TRY_H(pool_setup(h->device));
  CUDA_TRY_H(h->copier.init(h->device, (int)h->tune.get_ll("copy_threads", 0)));
  clap("pool_setup + copier.init");

  const int F = desc->num_factors;
  const int Fp = (F + 1) & ~1;
  const bool has_ctx = model_has_ctx(desc->model);

  // ---- context -> condition table (dense, -1 padded) ------------------------------------------------
  int Dmax = 0;
  std::vector<int32_t> ctx_tab;
  if (has_ctx) {
    for (int c = 0; c < desc->num_contexts; c++) {
      int len = desc->ctx_ptr[c + 1] - desc->ctx_ptr[c];
      if (len < 0) { fail(h, CARS_E_INVALID, "ctx_ptr not monotone at %d", c); return bail(CARS_E_INVALID); }
      if (len > Dmax) Dmax = len;
    }
    if (Dmax == 0) Dmax = 1;
    ctx_tab.assign((size_t)desc->num_contexts * Dmax, -1);
    for (int c = 0; c < desc->num_contexts; c++) {
      int k0 = desc->ctx_ptr[c], k1 = desc->ctx_ptr[c + 1];
      for (int k = k0; k < k1; k++) {
        int cond = desc->ctx_cond[k];
        if (cond < 0 || cond >= desc->num_conditions) {
          fail(h, CARS_E_INVALID, "condition id %d of context %d out of range", cond, c);
          return bail(CARS_E_INVALID);
        }
        for (int k2 = k0; k2 < k; k2++)
          if (desc->ctx_cond[k2] == cond) {
            fail(h, CARS_E_UNSUPPORTED, "context %d lists condition %d twice", c, cond);
            return bail(CARS_E_UNSUPPORTED);
          }
        ctx_tab[(size_t)c * Dmax + (k - k0)] = cond;
      }
    }
    CUDA_TRY_H(dev_alloc(&h->d_ctx_tab, ctx_tab.size()));
    CUDA_TRY_H(cudaMemcpyAsync(h->d_ctx_tab, ctx_tab.data(), ctx_tab.size() * 4, cudaMemcpyHostToDevice, h->stream));
    h->st.h2d_bytes += (int64_t)ctx_tab.size() * 4;
  }
  h->Dmax = Dmax;
  clap("context table");

  // ---- range checks on the rating arrays -------------------------------------------------------------
  const int64_t nnz = desc->nnz;
  const int sched_req = desc->schedule;
  const bool fast = desc->mode == CARS_FAST;
  // the flagged and the FAST schedule validate the ids inside their own device pass over the ratings
  const bool fused_check = fast || (sched_req == CARS_SCHED_FLAGGED && desc->model != CARS_CAMF_C && desc->model != CARS_CAMF_ICS &&
                                   desc->model != CARS_CAMF_LCS && desc->model != CARS_CAMF_MCS && desc->model != CARS_SVDPP);
  for (int64_t n = 0; n < nnz && !fused_check; n++) {
    if ((unsigned)desc->u[n] >= (unsigned)desc->num_users || (unsigned)desc->j[n] >= (unsigned)desc->num_items ||
        (has_ctx && (unsigned)desc->ctx[n] >= (unsigned)desc->num_contexts)) {
      fail(h, CARS_E_INVALID, "rating %lld has an id out of range (u=%d j=%d ctx=%d)", (long long)n, desc->u[n],
           desc->j[n], has_ctx ? desc->ctx[n] : -1);
      return bail(CARS_E_INVALID);
    }
  }

  // ---- launch geometry (the dataflow schedule sizes its chunks from the number of resident groups) ------
  const int model = desc->model;
  h->fast = fast;
  h->serial = !fast && (model == CARS_CAMF_C || model == CARS_CAMF_ICS || model == CARS_CAMF_LCS || model == CARS_CAMF_MCS ||
                        model == CARS_SVDPP);
  const int sched = sched_req;
  if (sched != CARS_SCHED_DATAFLOW && sched != CARS_SCHED_WAVEFRONT && sched != CARS_SCHED_FLAGGED) {
    fail(h, CARS_E_INVALID, "unknown schedule %d", sched);
    return bail(CARS_E_INVALID);
  }
  h->dataflow = !fast && !h->serial && sched == CARS_SCHED_DATAFLOW;
  h->flagged = !fast && !h->serial && sched == CARS_SCHED_FLAGGED;
  int groups_per_cta = 1;
  if (h->serial) {
    h->grid = 1; h->block = 32;
    h->smem = (size_t)((model == CARS_SVDPP ? 2 : 1) * Fp + 2) * 8;  // SVD++: + Q[j] for the lanes' Y[k].Q[j] chains
  } else {
    // EXACT default 8; rows of at most 32 factors with more than 4 context dimensions keep 8 lanes per rating (shape 11): the lanes of a
    // group fetch one condition cell each, dimensions beyond the group's lanes take the slow path
    // FAST default 0; rows of at most 16 / 32 factors with at most 4 context dimensions run 4 lanes per rating (shape 6 / 7): 8 ratings per warp
    // instruction -- BiasedMF F = 10: 7.8 against 11.4 ms per 100 M ratings, CAMF_CI F = 10: 11.6 against 16.0
    const bool few_dims = Dmax <= 4;
    const int shape = (int)h->tune.get_ll("shape", fast ? (few_dims && Fp <= 16 ? 6 : few_dims && Fp <= 32 ? 7 : 0)
                                                        : (Fp <= 32 && !few_dims ? 11 : 8));
    LaunchPlan plan = fast ? pick_fast_plan(model, Fp, F, shape)
                      : h->dataflow ? pick_dataflow_plan(model, Fp)
                      : h->flagged ? pick_flagged_plan(model, Fp, F, shape) : pick_plan(model, Fp);
    if (h->flagged && h->tune.get_ll("tagged", kTaggedDefault) != 0 && !h->tune.is("levels", "host")) {
      static const int kModelOf[] = {M_PMF, M_BIASEDMF, M_CAMF_C, M_CAMF_CI, M_CAMF_CU, -1, M_CAMF_CUCI, M_CAMF_ICS, M_CAMF_LCS, M_CAMF_MCS, M_SVDPP};
      h->tl = tagged_layout(kModelOf[model], F, has_ctx ? desc->num_conditions : 0);
      LaunchPlan tp = pick_tagged_plan(model, h->tl.p_lines(), h->tl.q_lines(), (int)h->tune.get_ll("tagged_ctas", 3));
      if (tp.fn) {
        plan = tp;
        h->tagged = true;
        h->tm.lp = h->tl.p_lines();
        h->tm.lq = h->tl.q_lines();
      }
    }
    h->plan = plan;
    if (!plan.fn) { fail(h, CARS_E_UNSUPPORTED, "no kernel for model %d mode %d F %d", model, desc->mode, F); return bail(CARS_E_UNSUPPORTED); }
    const int G = 32 / plan.lpr;
    groups_per_cta = (plan.threads / 32) * G;
    h->block = plan.threads;
    // FAST: room for the hot rows' accumulators (at most 32 rows / 32 KB) enters the occupancy up front
    h->hot_stride = Fp + 2 + ((model == CARS_CAMF_CI || model == CARS_CAMF_CUCI) ? desc->num_conditions : 0);
    const int hot_max = fast ? (4096 / h->hot_stride < 32 ? 4096 / h->hot_stride : 32) : 0;
    h->smem = fast ? (size_t)hot_max * h->hot_stride * 8 + (size_t)hot_max * 4 + 16 : (size_t)groups_per_cta * (Fp + 2) * 8;
    if (fast && plan.bulk) {  // two staged delta rows per group behind the hot-row area
      h->bulk_offset = (int)((h->smem + 127) & ~(size_t)127);
      h->smem = (size_t)h->bulk_offset + (size_t)groups_per_cta * 2 * Fp * 8;
    }
    if (h->tagged) {
      const int extras = (h->tl.p_payload() - F) + (h->tl.q_payload() - F);
      h->smem = (size_t)groups_per_cta * ((((F + extras + 1) & ~1) + 2) * 8);
    }
    CUDA_TRY_H(cudaFuncSetAttribute(plan.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    int per_sm = 0;
    CUDA_TRY_H(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan.fn, plan.threads, h->smem));
    if (per_sm < 1) { fail(h, CARS_E_CUDA, "SGD kernel does not fit on an SM (smem %zu)", h->smem); return bail(CARS_E_CUDA); }
    h->grid = h->sm_count * per_sm;
    if (fast) {
      // A small training set on a full grid would have most of its ratings in flight at once -- every update computed
      // from values that miss a large share of the epoch's other updates.  Keep at least 256 ratings per group.
      const int64_t want = (nnz + 256ll * groups_per_cta - 1) / (256ll * groups_per_cta);
      if (want < h->grid) h->grid = (int)(want < 1 ? 1 : want);
    }
  }

  clap("launch plan (cudaFuncSetAttribute + occupancy)");
  // ---- schedule ----------------------------------------------------------------------------------------
  auto t0 = std::chrono::steady_clock::now();
  h->nnz = nnz;
  const size_t U = desc->num_users, I = desc->num_items, C = desc->num_conditions;
  const bool sched_trace = h->tune.get_ll("sched_trace", 0) != 0;
  if (fast) {
    // K7h: user-sorted record stream, chunk table and damping factors, all on the device (fast_schedule.cuh)
    const int64_t total_groups = (int64_t)h->grid * groups_per_cta;
    int64_t chunk_len = h->tune.get_ll("fast_chunk", 0);
    if (chunk_len <= 0) {
      chunk_len = nnz / (8 * total_groups);
      chunk_len = chunk_len < 32 ? 32 : (chunk_len > 512 ? 512 : chunk_len);
    }
    h->num_chunks = nnz ? (nnz + chunk_len - 1) / chunk_len : 0;
    const double in_flight = (double)(h->num_chunks < total_groups ? h->num_chunks : total_groups);
    // Default cap on the expected number of concurrent updates of one shared row: 4.  c concurrent ratings add c stale
    // steps at once, i.e. the row sees a learning rate of c * lr against a curvature of about |p|^2 + D + 1 (the row's
    // factors, its D condition cells and its bias all move the same prediction); with the bold driver taking lr to ~0.05
    // and curvature ~6, c * lr * curvature stays below the stability bound 2 for c <= 4 (8 was measured to diverge on
    // the Frappe-shaped CAMF_C input after the bold driver had tripled lr).
    const double max_conc = desc->fast_max_conc == 0.0 ? 4.0 : desc->fast_max_conc;
    CUDA_TRY_H(dev_alloc(&h->d_rec, (size_t)nnz));
    CUDA_TRY_H(dev_alloc(&h->d_chunk_start, (size_t)h->num_chunks + 1));
    CUDA_TRY_H(dev_alloc(&h->d_item_scale, I));
    if (model == CARS_CAMF_C) CUDA_TRY_H(dev_alloc(&h->d_cond_scale, C));
    const int hot_max = (int)h->tune.get_ll("fast_hot_rows", 4096 / h->hot_stride < 32 ? 4096 / h->hot_stride : 32);
    h->hot_flush = (int)h->tune.get_ll("fast_hot_flush", 16);
    if (h->hot_flush < 1) h->hot_flush = 1;
    CUDA_TRY_H(dev_alloc(&h->d_hot_slot, I));
    CUDA_TRY_H(dev_alloc(&h->d_hot_items, 64));
    h->flags_words = 64;
    CUDA_TRY_H(dev_alloc(&h->d_flags, h->flags_words));
    CUDA_TRY_H(cudaStreamSynchronize(h->stream));  // the context table must have landed
    FastBuild fb;
    cudaError_t be = build_fast_on_device(desc->num_users, desc->num_items, desc->num_contexts, nnz, desc->u, desc->j,
                                          has_ctx ? desc->ctx : nullptr, desc->r, h->stream, h->sm_count, h->copier, chunk_len,
                                          in_flight, max_conc, h->d_ctx_tab, Dmax, desc->num_conditions, h->d_rec,
                                          h->d_chunk_start, h->d_item_scale, h->d_cond_scale,
                                          hot_max > 4096 / h->hot_stride ? 4096 / h->hot_stride : (hot_max > 32 ? 32 : hot_max), h->hot_flush, h->grid,
                                          kHotMinFlushes * h->hot_flush * (((int64_t)h->grid * groups_per_cta + 31) / 32), h->d_hot_slot, h->d_hot_items, h->mem, &fb);
    if (fb.bad_index >= 0) {
      const int64_t n = fb.bad_index;
      fail(h, CARS_E_INVALID, "rating %lld has an id out of range (u=%d j=%d ctx=%d)", (long long)n, desc->u[n], desc->j[n],
           has_ctx ? desc->ctx[n] : -1);
      return bail(CARS_E_INVALID);
    }
    if (be != cudaSuccess) {
      fail(h, be == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA, "building the FAST schedule failed: %s", cudaGetErrorString(be));
      return bail(be == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA);
    }
    h->damp_items = fb.min_item_scale < 1.0;
    h->num_hot = fb.num_hot;
    h->st.fast_hot_rows = fb.num_hot;
    h->damp_conds = h->d_cond_scale && fb.min_cond_scale < 1.0;
    h->num_levels = h->num_chunks;
    h->max_level_size = fb.max_chunk;
    h->st.h2d_bytes += fb.h2d_bytes;
    h->st.kernel_launches += fb.kernel_launches;
    h->st.schedule_copy_ms = fb.copy_ms;
    h->st.schedule_levels_ms = fb.sort_ms;
    h->st.schedule_pack_ms = fb.pack_ms;
    h->st.fast_min_item_scale = fb.min_item_scale;
    h->st.fast_min_cond_scale = fb.min_cond_scale;
    h->st.max_item_degree = fb.max_item_degree;
    h->st.schedule_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  } else if (h->dataflow) {
    // K7d: reference order is kept; every rating learns its position in its user's and its item's chain,
    // and the stream is cut into chunks (at user changes) that the kernel hands out in order.
    std::vector<RatingRec> recs;
    std::vector<int64_t> chunk_start;
    int64_t max_chunk = 0;
    try {
      recs.resize((size_t)nnz);
      std::vector<uint32_t> cu(U, 0u), cj(I, 0u);
      const int64_t total_groups = (int64_t)h->grid * groups_per_cta;
      int64_t lmin = nnz / (4 * total_groups);
      lmin = lmin < 1 ? 1 : (lmin > 64 ? 64 : lmin);
      chunk_start.reserve((size_t)(nnz / lmin + 2));
      int64_t cur = 0;
      for (int64_t n = 0; n < nnz; n++) {
        const int32_t uu = desc->u[n], jj = desc->j[n];
        if (n == 0) {
          chunk_start.push_back(0);
        } else if (uu != desc->u[n - 1] && n - cur >= lmin) {
          if (n - cur > max_chunk) max_chunk = n - cur;
          chunk_start.push_back(n);
          cur = n;
        }
        RatingRec& x = recs[(size_t)n];
        x.u = uu; x.j = jj; x.ctx = has_ctx ? desc->ctx[n] : 0;
        x.ku = (int32_t)cu[uu]++; x.kj = (int32_t)cj[jj]++; x.pad = 0; x.r = desc->r[n];
      }
      if (nnz - cur > max_chunk) max_chunk = nnz - cur;
      chunk_start.push_back(nnz);
    } catch (...) {
      fail(h, CARS_E_OOM, "host allocation failed while building the schedule");
      return bail(CARS_E_OOM);
    }
    h->num_chunks = nnz ? (int64_t)chunk_start.size() - 1 : 0;
    h->num_levels = h->num_chunks;
    h->max_level_size = max_chunk;
    h->st.schedule_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CUDA_TRY_H(dev_alloc(&h->d_rec, (size_t)nnz));
    CUDA_TRY_H(dev_alloc(&h->d_chunk_start, chunk_start.size()));
    CUDA_TRY_H(dev_alloc(&h->d_chunk_loss, (size_t)h->num_chunks));
    h->flags_words = 64 + I + U;
    CUDA_TRY_H(dev_alloc(&h->d_flags, h->flags_words));
    if (nnz) {
      CUDA_TRY_H(cudaMemcpyAsync(h->d_rec, recs.data(), (size_t)nnz * sizeof(RatingRec), cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY_H(cudaMemcpyAsync(h->d_chunk_start, chunk_start.data(), chunk_start.size() * 8, cudaMemcpyHostToDevice, h->stream));
      h->st.h2d_bytes += nnz * (int64_t)sizeof(RatingRec) + (int64_t)chunk_start.size() * 8;
    }
    h->loss_blocks = (int)((h->num_chunks + 1023) / 1024);
    if (h->loss_blocks < 1) h->loss_blocks = 1;
    if (h->loss_blocks > 1024) h->loss_blocks = 1024;
    CUDA_TRY_H(cudaStreamSynchronize(h->stream));  // recs / chunk_start die at the end of this block
  } else if (h->flagged) {
    // chains, dependency levels (Kahn by frontiers) and the level sort all run on the device (schedule_gpu.cuh)
    CUDA_TRY_H(dev_alloc(&h->d_rec, (size_t)nnz));
    h->flags_words = 64 + I + U;
    CUDA_TRY_H(dev_alloc(&h->d_flags, h->flags_words));
    if (sched_trace)
      fprintf(stderr, "[cars schedule] %-28s %8.2f ms\n", "cudaMalloc(records, flags)",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    FlaggedBuild fb;
    cudaError_t be = build_flagged_on_device(desc->num_users, desc->num_items, desc->num_contexts, nnz, desc->u, desc->j,
                                             has_ctx ? desc->ctx : nullptr, desc->r, h->stream, h->sm_count, h->copier,
                                             h->d_rec, &fb, h->mem, h->tune.is("levels", "host"), sched_trace);
    if (fb.bad_index >= 0) {
      const int64_t n = fb.bad_index;
      fail(h, CARS_E_INVALID, "rating %lld has an id out of range (u=%d j=%d ctx=%d)", (long long)n, desc->u[n], desc->j[n],
           has_ctx ? desc->ctx[n] : -1);
      return bail(CARS_E_INVALID);
    }
    if (be != cudaSuccess) {
      fail(h, be == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA, "building the schedule failed: %s", cudaGetErrorString(be));
      return bail(be == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA);
    }
    h->num_levels = fb.num_levels;
    h->max_level_size = fb.max_level_size;
    h->st.h2d_bytes += fb.h2d_bytes;
    h->st.kernel_launches += fb.kernel_launches;
    h->st.schedule_copy_ms = fb.copy_ms;
    h->st.schedule_levels_ms = fb.levels_ms;
    h->st.schedule_pack_ms = fb.pack_ms;
    h->st.schedule_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  } else {
    std::vector<int64_t> level_start;
    HostSchedule sched_host;
    const int32_t* su = desc->u;
    const int32_t* sj = desc->j;
    const int32_t* sc = desc->ctx;
    const double* sr = desc->r;
    if (!h->serial) {
      if (!build_wavefront_schedule(desc->num_users, desc->num_items, nnz, desc->u, desc->j, has_ctx ? desc->ctx : nullptr,
                                    desc->r, &sched_host)) {
        fail(h, CARS_E_OOM, "host allocation failed while building the schedule");
        return bail(CARS_E_OOM);
      }
      su = sched_host.u.data(); sj = sched_host.j.data(); sc = has_ctx ? sched_host.ctx.data() : nullptr; sr = sched_host.r.data();
      h->num_levels = (int64_t)sched_host.level_start.size() - 1;
      h->max_level_size = sched_host.max_level_size;
    } else {
      h->num_levels = nnz;  // every rating is its own level
      h->max_level_size = nnz ? 1 : 0;
    }
    h->st.schedule_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    CUDA_TRY_H(dev_alloc(&h->d_u, nnz));
    CUDA_TRY_H(dev_alloc(&h->d_j, nnz));
    CUDA_TRY_H(dev_alloc(&h->d_r, nnz));
    if (has_ctx) CUDA_TRY_H(dev_alloc(&h->d_ctx, nnz));
    if (nnz) {
      CUDA_TRY_H(cudaMemcpyAsync(h->d_u, su, nnz * 4, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY_H(cudaMemcpyAsync(h->d_j, sj, nnz * 4, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY_H(cudaMemcpyAsync(h->d_r, sr, nnz * 8, cudaMemcpyHostToDevice, h->stream));
      if (has_ctx) CUDA_TRY_H(cudaMemcpyAsync(h->d_ctx, sc, nnz * 4, cudaMemcpyHostToDevice, h->stream));
      h->st.h2d_bytes += nnz * (has_ctx ? 20 : 16);
    }
    if (!h->serial) {
      CUDA_TRY_H(dev_alloc(&h->d_level_start, sched_host.level_start.size()));
      CUDA_TRY_H(cudaMemcpyAsync(h->d_level_start, sched_host.level_start.data(), sched_host.level_start.size() * 8,
                                 cudaMemcpyHostToDevice, h->stream));
      h->st.h2d_bytes += (int64_t)sched_host.level_start.size() * 8;
    }
    CUDA_TRY_H(cudaStreamSynchronize(h->stream));  // host staging vectors die at the end of this block
  }

  // ---- model storage ----------------------------------------------------------------------------------
  DeviceModel& m = h->m;
  m.F = F; m.Fp = Fp; m.C = desc->num_conditions; m.Dmax = Dmax;
  m.global_mean = desc->global_mean;
  m.reg_u = desc->reg_u; m.reg_i = desc->reg_i; m.reg_b = desc->reg_b; m.reg_c = desc->reg_c;
  m.ctx_tab = h->d_ctx_tab;
  CUDA_TRY_H(dev_alloc(&m.P, U * Fp));
  CUDA_TRY_H(dev_alloc(&m.Q, I * Fp));
  if (model == CARS_BIASEDMF || model == CARS_CAMF_C || model == CARS_CAMF_CI || model == CARS_SVDPP) CUDA_TRY_H(dev_alloc(&m.user_bias, U));
  if (model == CARS_BIASEDMF || model == CARS_CAMF_C || model == CARS_CAMF_CU || model == CARS_SVDPP) CUDA_TRY_H(dev_alloc(&m.item_bias, I));
  if (model == CARS_SVDPP) {
    // userItemsCache = train.rowColumnsCache() (SVDPlusPlus.java:51): every user's items ascending, whatever the order of the
    // rating stream -- built here on the host (the one-warp chain is for small data) and kept on the device in CSR form
    CUDA_TRY_H(dev_alloc(&m.Y, I * Fp));
    CUDA_TRY_H(cudaMemsetAsync(m.Y, 0, I * Fp * sizeof(double), h->stream));
    std::vector<int32_t> uptr(U + 1, 0), uitems((size_t)nnz);
    for (int64_t n = 0; n < nnz; n++) {
      if ((uint32_t)desc->u[n] >= (uint32_t)U || (uint32_t)desc->j[n] >= (uint32_t)I) {
        fail(h, CARS_E_INVALID, "rating %lld has an id out of range (u=%d j=%d ctx=-1)", (long long)n, desc->u[n], desc->j[n]);
        return bail(CARS_E_INVALID);
      }
      uptr[desc->u[n] + 1]++;
    }
    for (size_t k = 0; k < U; k++) uptr[k + 1] += uptr[k];
    {
      std::vector<int32_t> fill(uptr.begin(), uptr.end() - 1);
      for (int64_t n = 0; n < nnz; n++) uitems[fill[desc->u[n]]++] = desc->j[n];
      for (size_t k = 0; k < U; k++) std::sort(uitems.begin() + uptr[k], uitems.begin() + uptr[k + 1]);
    }
    CUDA_TRY_H(dev_alloc(&h->d_ui_ptr, U + 1));
    CUDA_TRY_H(dev_alloc(&h->d_ui_items, (size_t)nnz));
    CUDA_TRY_H(cudaMemcpyAsync(h->d_ui_ptr, uptr.data(), (U + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    if (nnz) CUDA_TRY_H(cudaMemcpyAsync(h->d_ui_items, uitems.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY_H(cudaStreamSynchronize(h->stream));  // the host vectors die here
    m.ui_ptr = h->d_ui_ptr; m.ui_items = h->d_ui_items;
  }
  if (model == CARS_CAMF_C) CUDA_TRY_H(dev_alloc(&m.cond_bias, C));
  if (model == CARS_CAMF_CI || model == CARS_CAMF_CUCI) CUDA_TRY_H(dev_alloc(&m.ic_bias, I * C));
  if (model == CARS_CAMF_CU || model == CARS_CAMF_CUCI) CUDA_TRY_H(dev_alloc(&m.uc_bias, U * C));
  if (model == CARS_CAMF_ICS || model == CARS_CAMF_LCS || model == CARS_CAMF_MCS) {
    if (desc->num_empty_conditions < Dmax) {
      fail(h, CARS_E_INVALID, "a context has %d conditions but only %d empty conditions were given", Dmax, desc->num_empty_conditions);
      return bail(CARS_E_INVALID);
    }
    for (int d = 0; d < desc->num_empty_conditions; d++)
      if ((unsigned)desc->empty_conditions[d] >= (unsigned)desc->num_conditions) {
        fail(h, CARS_E_INVALID, "empty_conditions[%d] = %d is not a condition id", d, desc->empty_conditions[d]);
        return bail(CARS_E_INVALID);
      }
    if (model == CARS_CAMF_ICS) CUDA_TRY_H(dev_alloc(&m.cc_sim, C * C));
    if (model == CARS_CAMF_LCS) {
      m.numF = desc->num_context_factors;
      CUDA_TRY_H(dev_alloc(&m.cf_lcs, C * (size_t)m.numF));
    }
    if (model == CARS_CAMF_MCS) {
      m.mcs_upbound = 1.0 / std::sqrt((double)desc->num_context_dims);  // CAMF_MCS.java:44-45
      m.mcs_lowbound = 1.0 / 1e100;
      CUDA_TRY_H(dev_alloc(&m.c_mcs, C));
    }
    CUDA_TRY_H(dev_alloc(&h->d_empty_cond, (size_t)desc->num_empty_conditions));
    CUDA_TRY_H(cudaMemcpyAsync(h->d_empty_cond, desc->empty_conditions, (size_t)desc->num_empty_conditions * 4, cudaMemcpyHostToDevice, h->stream));
    m.empty_cond = h->d_empty_cond;
  }

  if (h->tagged) {
    CUDA_TRY_H(dev_alloc(&h->tm.Pt, U * (size_t)h->tm.lp * 16));
    CUDA_TRY_H(dev_alloc(&h->tm.Qt, I * (size_t)h->tm.lq * 16));
  }
  CUDA_TRY_H(dev_alloc(&h->d_barrier, 1));
  CUDA_TRY_H(dev_alloc(&h->d_partial, (size_t)(h->grid > 1024 ? h->grid : 1024)));
  CUDA_TRY_H(dev_alloc(&h->d_loss, 1));
  CUDA_TRY_H(cudaMallocHost((void**)&h->h_loss, sizeof(double)));
  CUDA_TRY_H(cudaStreamSynchronize(h->stream));  // host staging vectors die at return

  h->st.nnz = nnz;
  h->st.num_levels = h->num_levels;
  h->st.max_level_size = h->max_level_size;
  h->st.grid_ctas = h->grid;
  h->st.block_threads = h->block;
  h->st.sm_count = h->sm_count;
  // the handle must not keep caller pointers
  h->d.u = h->d.j = h->d.ctx = nullptr; h->d.r = nullptr; h->d.ctx_ptr = h->d.ctx_cond = nullptr; h->d.stream = nullptr;
  h->d.gpu_ids = nullptr; h->d.tuning = nullptr; h->d.empty_conditions = nullptr;
  if (h->tune.get_ll("sched_trace", 0) != 0)
    fprintf(stderr, "[cars create] total %8.2f ms (before the schedule %8.2f ms, schedule %8.2f ms)\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_create).count(),
            std::chrono::duration<double, std::milli>(t0 - t_create).count(), h->st.schedule_ms);
  *out = h;
  return CARS_OK;
#undef CUDA_TRY_H
}

// ------------------------------------------------------------------------------------------------
// upload / download
// ------------------------------------------------------------------------------------------------
static int copy_rows(cars_handle* h, bool to_device, double* dev, double* host, size_t rows, int F, int Fp,
                     std::vector<CopySeg>* segs) {
  if (rows == 0) return CARS_OK;
  if (F == Fp) {  // contiguous on both sides: goes with the other arrays through the staged copier
    segs->push_back(CopySeg{dev, host, rows * (size_t)F * 8});
    if (to_device) h->st.h2d_bytes += (int64_t)rows * F * 8;
    else h->st.d2h_bytes += (int64_t)rows * F * 8;
    return CARS_OK;
  }
  if (to_device) {
    if (F != Fp) CUDA_TRY(h, cudaMemsetAsync(dev, 0, rows * Fp * 8, h->stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(dev, (size_t)Fp * 8, host, (size_t)F * 8, (size_t)F * 8, rows, cudaMemcpyHostToDevice, h->stream));
    h->st.h2d_bytes += (int64_t)rows * F * 8;
  } else {
    CUDA_TRY(h, cudaMemcpy2DAsync(host, (size_t)F * 8, dev, (size_t)Fp * 8, (size_t)F * 8, rows, cudaMemcpyDeviceToHost, h->stream));
    h->st.d2h_bytes += (int64_t)rows * F * 8;
  }
  return CARS_OK;
}

static int copy_vec(cars_handle* h, bool to_device, double* dev, double* host, size_t n, std::vector<CopySeg>* segs) {
  if (n == 0) return CARS_OK;
  segs->push_back(CopySeg{dev, host, n * 8});
  if (to_device) h->st.h2d_bytes += (int64_t)n * 8;
  else h->st.d2h_bytes += (int64_t)n * 8;
  return CARS_OK;
}

// K1t keeps the model in tagged rows; every other consumer reads the standard layout.  Convert on demand.
static int ensure_std(cars_handle* h) {
  if (!h->tagged || h->std_valid) return CARS_OK;
  const int blocks = h->sm_count * 8;
  tagged_unpack_kernel<<<blocks, 256, 0, h->stream>>>(h->m, h->tl, h->tm, 0, (int64_t)h->d.num_users);
  tagged_unpack_kernel<<<blocks, 256, 0, h->stream>>>(h->m, h->tl, h->tm, 1, (int64_t)h->d.num_items);
  CUDA_TRY(h, cudaGetLastError());
  h->st.kernel_launches += 2;
  h->std_valid = true;
  return CARS_OK;
}
static int ensure_tagged(cars_handle* h) {
  if (!h->tagged || h->tagged_valid) return CARS_OK;
  const int blocks = h->sm_count * 8;
  tagged_pack_kernel<<<blocks, 256, 0, h->stream>>>(h->m, h->tl, h->tm, 0, (int64_t)h->d.num_users);
  tagged_pack_kernel<<<blocks, 256, 0, h->stream>>>(h->m, h->tl, h->tm, 1, (int64_t)h->d.num_items);
  CUDA_TRY(h, cudaGetLastError());
  h->st.kernel_launches += 2;
  h->tagged_valid = true;
  return CARS_OK;
}

static int transfer(cars_handle* h, const cars_model_arrays* a, bool to_device, bool skip_item_side = false) {
  if (!h) return CARS_E_INVALID;
  if (!a) return fail(h, CARS_E_INVALID, "arrays is NULL");
  if (h->epoch_pending) return fail(h, CARS_E_STATE, "an epoch is pending; call cars_epoch_wait first");
  CUDA_TRY(h, cudaSetDevice(h->device));
  const DeviceModel& m = h->m;
  const size_t U = h->d.num_users, I = h->d.num_items, C = h->d.num_conditions;
  struct Item { double* dev; double* host; const char* name; };
  const Item need[] = {{m.P, a->P, "P"}, {m.Q, a->Q, "Q"}, {m.user_bias, a->user_bias, "user_bias"},
                       {m.item_bias, a->item_bias, "item_bias"}, {m.cond_bias, a->cond_bias, "cond_bias"},
                       {m.ic_bias, a->ic_bias, "ic_bias"}, {m.uc_bias, a->uc_bias, "uc_bias"}, {m.cc_sim, a->cc_sim, "cc_sim"},
                       {m.cf_lcs, a->cf_lcs, "cf_lcs"}, {m.c_mcs, a->c_mcs, "c_mcs"}, {m.Y, a->Y, "Y"}};
  for (const Item& it : need) {
    const bool item_side = it.dev == m.Q || it.dev == m.item_bias || it.dev == m.cond_bias || it.dev == m.ic_bias;
    if (it.dev && !it.host && !(skip_item_side && item_side))
      return fail(h, CARS_E_INVALID, "model array %s is required for this model but NULL", it.name);
  }
  int rc;
  std::vector<CopySeg> segs;
  if (!to_device && (rc = ensure_std(h))) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // the staged copies run on the copier's own streams
  if ((rc = copy_rows(h, to_device, m.P, a->P, U, m.F, m.Fp, &segs))) return rc;
  if (!skip_item_side && (rc = copy_rows(h, to_device, m.Q, a->Q, I, m.F, m.Fp, &segs))) return rc;
  if (m.Y && (rc = copy_rows(h, to_device, m.Y, a->Y, I, m.F, m.Fp, &segs))) return rc;
  if (m.user_bias && (rc = copy_vec(h, to_device, m.user_bias, a->user_bias, U, &segs))) return rc;
  if (!skip_item_side && m.item_bias && (rc = copy_vec(h, to_device, m.item_bias, a->item_bias, I, &segs))) return rc;
  if (!skip_item_side && m.cond_bias && (rc = copy_vec(h, to_device, m.cond_bias, a->cond_bias, C, &segs))) return rc;
  if (!skip_item_side && m.ic_bias && (rc = copy_vec(h, to_device, m.ic_bias, a->ic_bias, I * C, &segs))) return rc;
  if (m.uc_bias && (rc = copy_vec(h, to_device, m.uc_bias, a->uc_bias, U * C, &segs))) return rc;
  if (m.cc_sim && (rc = copy_vec(h, to_device, m.cc_sim, a->cc_sim, C * C, &segs))) return rc;
  if (m.cf_lcs && (rc = copy_vec(h, to_device, m.cf_lcs, a->cf_lcs, C * (size_t)m.numF, &segs))) return rc;
  if (m.c_mcs && (rc = copy_vec(h, to_device, m.c_mcs, a->c_mcs, C, &segs))) return rc;
  CUDA_TRY(h, h->copier.run(segs.data(), (int)segs.size(), to_device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // the strided (odd F) row copies
  if (to_device) { h->std_valid = true; h->tagged_valid = false; }
  if (!to_device && m.cc_sim)  // SymmMatrix: the (max, min) cell is the live one; show it on both sides of the caller's array
    for (size_t i = 0; i < C; i++)
      for (size_t k = 0; k < i; k++) a->cc_sim[k * C + i] = a->cc_sim[i * C + k];
  return CARS_OK;
}

extern "C" int cars_upload(cars_handle* h, const cars_model_arrays* host) {
  if (h && h->multi) return multi_transfer(h, host, true);
  const auto t0 = std::chrono::steady_clock::now();
  int rc = transfer(h, host, true);
  if (rc == CARS_OK) h->uploaded = true;
  if (h && h->tune.get_ll("sched_trace", 0) != 0)
    fprintf(stderr, "[cars upload] %8.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return rc;
}

extern "C" int cars_download(cars_handle* h, const cars_model_arrays* host) {
  if (h && !h->uploaded) return fail(h, CARS_E_STATE, "cars_download before cars_upload");
  if (h && h->multi) return multi_transfer(h, host, false);
  return transfer(h, host, false);
}

// ------------------------------------------------------------------------------------------------
// epoch
// ------------------------------------------------------------------------------------------------
extern "C" int cars_epoch_begin(cars_handle* h, double lrate) {
  if (!h) return CARS_E_INVALID;
  if (h->multi) return fail(h, CARS_E_UNSUPPORTED, "a multi-GPU handle runs whole epochs: use cars_epoch");
  if (!h->uploaded) return fail(h, CARS_E_STATE, "cars_epoch before cars_upload");
  if (h->epoch_pending) return fail(h, CARS_E_STATE, "previous epoch not collected; call cars_epoch_wait");
  CUDA_TRY(h, cudaSetDevice(h->device));
  RatingStream s;
  s.u = h->d_u; s.j = h->d_j; s.ctx = h->d_ctx; s.r = h->d_r;
  s.level_start = h->d_level_start;
  s.num_levels = (int32_t)h->num_levels;
  DeviceModel m = h->m;
  CUDA_TRY(h, cudaEventRecord(h->ev_beg, h->stream));
  int partials = h->grid;
  if (h->serial) {
    const void* fn = pick_serial(h->d.model, m.Fp);
    int64_t nnz = h->nnz;
    void* args[] = {&m, &s, &nnz, &lrate, &h->d_partial};
    CUDA_TRY(h, cudaLaunchKernel(fn, dim3(1), dim3(32), args, h->smem, h->stream));
    h->st.kernel_launches += 1;
  } else if (h->fast) {
    FastStream fs;
    fs.rec = h->d_rec; fs.chunk_start = h->d_chunk_start; fs.counter = h->d_flags; fs.num_chunks = (uint32_t)h->num_chunks;
    fs.item_scale = h->damp_items ? h->d_item_scale : nullptr;
    fs.cond_scale = h->damp_conds ? h->d_cond_scale : nullptr;
    fs.hot_slot = h->num_hot > 0 ? h->d_hot_slot : nullptr;
    fs.hot_items = h->d_hot_items; fs.num_hot = h->num_hot; fs.hot_flush = h->hot_flush;
    fs.bulk_offset = h->bulk_offset;
    CUDA_TRY(h, cudaMemsetAsync(h->d_flags, 0, h->flags_words * sizeof(unsigned), h->stream));
    void* args[] = {&m, &fs, &lrate, &h->d_partial};
    CUDA_TRY(h, cudaLaunchKernel(h->plan.fn, dim3(h->grid), dim3(h->block), args, h->smem, h->stream));
    h->st.kernel_launches += 1;
  } else if (h->dataflow) {
    const LaunchPlan& plan = h->plan;
    DataflowStream ds;
    ds.rec = h->d_rec; ds.chunk_start = h->d_chunk_start; ds.chunk_loss = h->d_chunk_loss;
    ds.counter = h->d_flags; ds.done_j = h->d_flags + 64; ds.done_u = h->d_flags + 64 + h->d.num_items;
    ds.num_chunks = (uint32_t)h->num_chunks;
    CUDA_TRY(h, cudaMemsetAsync(h->d_flags, 0, h->flags_words * sizeof(unsigned), h->stream));
    if (h->num_chunks > 0) {
      void* args[] = {&m, &ds, &lrate};
      CUDA_TRY(h, cudaLaunchKernel(plan.fn, dim3(h->grid), dim3(h->block), args, h->smem, h->stream));
      h->st.kernel_launches += 1;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_end, h->stream));
    chunk_loss_reduce_kernel<<<h->loss_blocks, 256, 0, h->stream>>>(h->d_chunk_loss, h->num_chunks, h->d_partial);
    CUDA_TRY(h, cudaGetLastError());
    h->st.kernel_launches += 1;
    partials = h->loss_blocks;
  } else if (h->flagged && h->tagged) {
    int rc = ensure_tagged(h);  // (the conversion is part of the epoch it serves: inside the timed region)
    if (rc) return rc;
    const RatingRec* recs = h->d_rec;
    int64_t nnz = h->nnz;
    void* args[] = {&m, &h->tl, &h->tm, &recs, &nnz, &lrate, &h->d_partial};
    CUDA_TRY(h, cudaLaunchCooperativeKernel(h->plan.fn, dim3(h->grid), dim3(h->block), args, h->smem, h->stream));
    h->st.kernel_launches += 1;
    h->std_valid = false;
  } else if (h->flagged) {
    const LaunchPlan& plan = h->plan;
    CUDA_TRY(h, cudaMemsetAsync(h->d_flags, 0, h->flags_words * sizeof(unsigned), h->stream));
    const RatingRec* recs = h->d_rec;
    int64_t nnz = h->nnz;
    unsigned off_j = 64u;
    unsigned off_u = 64u + (unsigned)h->d.num_items;
#ifdef CARS_TRACE
    // developer build only (scripts/trace_flagged.py): per-rating stage timestamps of a window of ratings
    static unsigned long long* d_trace = nullptr;
    int64_t trace_lo = 0, trace_n = 0;
    if (const char* e = getenv("CARS_TRACE_WINDOW")) sscanf(e, "%lld:%lld", (long long*)&trace_lo, (long long*)&trace_n);
    if (trace_n > 0 && !d_trace) cudaMalloc((void**)&d_trace, (size_t)trace_n * 64);
    void* args[] = {&m, &recs, &nnz, &h->d_flags, &off_u, &off_j, &lrate, &h->d_partial, &d_trace, &trace_lo, &trace_n};
#else
    void* args[] = {&m, &recs, &nnz, &h->d_flags, &off_u, &off_j, &lrate, &h->d_partial};
#endif
    CUDA_TRY(h, cudaLaunchCooperativeKernel(plan.fn, dim3(h->grid), dim3(h->block), args, h->smem, h->stream));
#ifdef CARS_TRACE
    if (trace_n > 0) {
      if (const char* out = getenv("CARS_TRACE_OUT")) {
        std::vector<unsigned long long> hb((size_t)trace_n * 8);
        cudaStreamSynchronize(h->stream);
        cudaMemcpy(hb.data(), d_trace, hb.size() * 8, cudaMemcpyDeviceToHost);
        std::vector<RatingRec> hr((size_t)trace_n);
        cudaMemcpy(hr.data(), h->d_rec + trace_lo, hr.size() * sizeof(RatingRec), cudaMemcpyDeviceToHost);
        FILE* f = fopen(out, "wb");
        if (f) { fwrite(hb.data(), 8, hb.size(), f); fwrite(hr.data(), sizeof(RatingRec), hr.size(), f); fclose(f); }
      }
    }
#endif
    h->st.kernel_launches += 1;
  } else {
    const LaunchPlan& plan = h->plan;
    CUDA_TRY(h, cudaMemsetAsync(h->d_barrier, 0, sizeof(unsigned), h->stream));
    void* args[] = {&m, &s, &lrate, &h->d_barrier, &h->d_partial};
    CUDA_TRY(h, cudaLaunchCooperativeKernel(plan.fn, dim3(h->grid), dim3(h->block), args, h->smem, h->stream));
    h->st.kernel_launches += 1;
  }
  if (!h->dataflow) CUDA_TRY(h, cudaEventRecord(h->ev_end, h->stream));
  loss_finalize_kernel<<<1, 32, 0, h->stream>>>(h->d_partial, partials, h->d_loss);
  CUDA_TRY(h, cudaGetLastError());
  h->st.kernel_launches += 1;
  CUDA_TRY(h, cudaMemcpyAsync(h->h_loss, h->d_loss, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->st.d2h_bytes += 8;
  h->epoch_pending = true;
  return CARS_OK;
}

extern "C" int cars_epoch_wait(cars_handle* h, double* loss_out) {
  if (!h) return CARS_E_INVALID;
  if (h->multi) return fail(h, CARS_E_UNSUPPORTED, "a multi-GPU handle runs whole epochs: use cars_epoch");
  if (!h->epoch_pending) return fail(h, CARS_E_STATE, "no epoch pending");
  h->epoch_pending = false;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev_beg, h->ev_end));
  h->st.last_epoch_ms = ms;
  if (loss_out) *loss_out = *h->h_loss;
  return CARS_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU exchange of the item block (C1)
// ------------------------------------------------------------------------------------------------
struct ItemPart {
  double* ptr;
  int64_t n;
};
static int item_parts(const cars_handle* h, ItemPart out[4]) {
  const DeviceModel& m = h->m;
  const int64_t I = h->d.num_items;
  int k = 0;
  out[k++] = {m.Q, I * m.Fp};
  if (m.item_bias) out[k++] = {m.item_bias, I};
  if (m.ic_bias) out[k++] = {m.ic_bias, I * (int64_t)m.C};
  if (m.cond_bias) out[k++] = {m.cond_bias, (int64_t)m.C};  // CAMF_C (FAST only): shared by every shard
  return k;
}

extern "C" int cars_item_block_doubles(const cars_handle* h, int64_t* out) {
  if (!h || !out || h->multi) return CARS_E_INVALID;
  ItemPart parts[4];
  const int k = item_parts(h, parts);
  int64_t n = 0;
  for (int i = 0; i < k; i++) n += parts[i].n;
  *out = n;
  return CARS_OK;
}

extern "C" int cars_epoch_sharded_begin(cars_handle* h, double lrate, double* dev_delta) {
  if (!h) return CARS_E_INVALID;
  if (h->multi) return fail(h, CARS_E_UNSUPPORTED, "a multi-GPU handle exchanges its item block itself: use cars_epoch");
  if (!dev_delta) return fail(h, CARS_E_INVALID, "dev_delta is NULL");
  if (h->d.model == CARS_CAMF_C && !h->fast)
    return fail(h, CARS_E_UNSUPPORTED, "CAMF_C in EXACT mode is one chain through condBias; shard it in FAST mode");
  if (h->sharded_pending) return fail(h, CARS_E_STATE, "previous sharded epoch not finished");
  if (!h->uploaded) return fail(h, CARS_E_STATE, "cars_epoch before cars_upload");
  CUDA_TRY(h, cudaSetDevice(h->device));
  int rcs = ensure_std(h);
  if (rcs) return rcs;
  ItemPart parts[4];
  const int k = item_parts(h, parts);
  int64_t total = 0;
  for (int i = 0; i < k; i++) total += parts[i].n;
  if (!h->d_item_old) CUDA_TRY(h, dev_alloc(&h->d_item_old, (size_t)total));
  int64_t off = 0;
  for (int i = 0; i < k; i++) {
    CUDA_TRY(h, cudaMemcpyAsync(h->d_item_old + off, parts[i].ptr, (size_t)parts[i].n * 8, cudaMemcpyDeviceToDevice, h->stream));
    off += parts[i].n;
  }
  int rc = cars_epoch_begin(h, lrate);
  if (rc) return rc;
  if ((rc = ensure_std(h))) return rc;  // K1t trained the tagged rows: the delta is taken on the standard layout
  off = 0;
  const int blocks = h->sm_count * 8;
  for (int i = 0; i < k; i++) {
    item_delta_kernel<<<blocks, 256, 0, h->stream>>>(parts[i].ptr, h->d_item_old + off, dev_delta + off, parts[i].n);
    CUDA_TRY(h, cudaGetLastError());
    h->st.kernel_launches += 1;
    off += parts[i].n;
  }
  h->sharded_pending = true;
  return CARS_OK;
}

__global__ void item_apply_rows_kernel(double* __restrict__ cur, const double* __restrict__ old, const double* __restrict__ sum,
                                       const double* __restrict__ row_scale, int64_t row_len, int64_t n);  // multi_gpu.cuh

// row_scale != nullptr: block <- old + row_scale[item] * sum (COMBINE_TOUCHED), else old + scale * sum
static int sharded_finish_impl(cars_handle* h, const double* dev_delta, double scale, const double* row_scale, double* loss_out) {
  if (!h) return CARS_E_INVALID;
  if (h->multi) return fail(h, CARS_E_UNSUPPORTED, "a multi-GPU handle exchanges its item block itself: use cars_epoch");
  if (!h->sharded_pending) return fail(h, CARS_E_STATE, "no sharded epoch pending");
  if (!dev_delta) return fail(h, CARS_E_INVALID, "dev_delta is NULL");
  CUDA_TRY(h, cudaSetDevice(h->device));
  ItemPart parts[4];
  const int k = item_parts(h, parts);
  int64_t off = 0;
  const int blocks = h->sm_count * 8;
  const int64_t I = h->d.num_items;
  for (int i = 0; i < k; i++) {
    const int64_t row_len = parts[i].n / (I > 0 ? I : 1);
    if (row_scale && parts[i].ptr != h->m.cond_bias && row_len > 0)
      item_apply_rows_kernel<<<blocks, 256, 0, h->stream>>>(parts[i].ptr, h->d_item_old + off, dev_delta + off, row_scale, row_len, parts[i].n);
    else
      item_apply_kernel<<<blocks, 256, 0, h->stream>>>(parts[i].ptr, h->d_item_old + off, dev_delta + off, scale, parts[i].n);
    CUDA_TRY(h, cudaGetLastError());
    h->st.kernel_launches += 1;
    off += parts[i].n;
  }
  h->sharded_pending = false;
  h->tagged_valid = false;  // the item block changed in the standard layout
  return cars_epoch_wait(h, loss_out);
}

extern "C" int cars_epoch_sharded_finish(cars_handle* h, const double* dev_delta, double scale, double* loss_out) {
  return sharded_finish_impl(h, dev_delta, scale, nullptr, loss_out);
}

extern "C" int cars_epoch(cars_handle* h, double lrate, double* loss_out) {
  if (h && h->multi) return multi_epoch(h, lrate, loss_out);
  int rc = cars_epoch_begin(h, lrate);
  if (rc) return rc;
  return cars_epoch_wait(h, loss_out);
}

// ------------------------------------------------------------------------------------------------
// predict / evalRatings
// ------------------------------------------------------------------------------------------------
static int predict_device(cars_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx,
                          int bound, double lo, double hi, double* d_out, int32_t** du_, int32_t** dj_, int32_t** dc_) {
  const bool has_ctx = model_has_ctx(h->d.model);
  if (has_ctx && !ctx) return fail(h, CARS_E_INVALID, "ctx is required for this model");
  int rce = ensure_std(h);
  if (rce) return rce;
  CUDA_TRY(h, dev_alloc(du_, (size_t)n));
  CUDA_TRY(h, dev_alloc(dj_, (size_t)n));
  if (has_ctx) CUDA_TRY(h, dev_alloc(dc_, (size_t)n));
  {
    // the queries cross PCIe through the staged copier (pageable arrays: several host threads with pinned double
    // buffers instead of the driver's one bounce buffer) and are range-checked on the device, before the kernel that
    // would index with them (a 10 M-query host loop + plain copies were 30 of cars_predict's 44 ms)
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // the allocations above are stream-ordered
    CopySeg segs[3] = {{*du_, (void*)u, (size_t)n * 4}, {*dj_, (void*)j, (size_t)n * 4}, {has_ctx ? *dc_ : nullptr, (void*)ctx, (size_t)n * 4}};
    CUDA_TRY(h, h->copier.run(segs, has_ctx ? 3 : 2, true));
    h->st.h2d_bytes += n * (has_ctx ? 12 : 8);
    unsigned long long* d_bad = nullptr;
    CUDA_TRY(h, dev_alloc(&d_bad, 1));
    cudaError_t ve = cudaMemsetAsync(d_bad, 0xff, 8, h->stream);
    unsigned long long bad = ~0ull;
    if (ve == cudaSuccess) {
      validate_ids_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(*du_, *dj_, has_ctx ? *dc_ : nullptr, n, (uint32_t)h->d.num_users,
                                                                  (uint32_t)h->d.num_items, (uint32_t)h->d.num_contexts, d_bad);
      ve = cudaGetLastError();
    }
    if (ve == cudaSuccess) ve = cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, h->stream);
    if (ve == cudaSuccess) ve = cudaStreamSynchronize(h->stream);
    h->mem.free(d_bad);
    CUDA_TRY(h, ve);
    h->st.kernel_launches += 1;
    if (bad != ~0ull) return fail(h, CARS_E_INVALID, "query %lld has an id out of range", (long long)bad);
  }
  // K5: a group of 8 lanes per query, coalesced 16-byte row loads, in-order dot through shared memory (sgd_kernels.cuh)
  const int threads = 256;
  const int64_t groups = (int64_t)threads / 8;
  int64_t want = (n + groups - 1) / groups;
  const int64_t cap = (int64_t)h->sm_count * 16;
  const unsigned blocks = (unsigned)(want < cap ? want : cap);
  DeviceModel m = h->m;
  const size_t smem = (size_t)groups * (m.Fp + 2) * 8;
#define CARS_PREDICT(M)                                                                                                \
  CUDA_TRY(h, cudaFuncSetAttribute(predict_group_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  predict_group_kernel<M><<<blocks, threads, smem, h->stream>>>(m, n, *du_, *dj_, *dc_, bound, lo, hi, d_out)
  switch (h->d.model) {
    case CARS_PMF: CARS_PREDICT(M_PMF); break;
    case CARS_BIASEDMF: CARS_PREDICT(M_BIASEDMF); break;
    case CARS_CAMF_C: CARS_PREDICT(M_CAMF_C); break;
    case CARS_CAMF_CI: CARS_PREDICT(M_CAMF_CI); break;
    case CARS_CAMF_CU: CARS_PREDICT(M_CAMF_CU); break;
    case CARS_CAMF_CUCI: CARS_PREDICT(M_CAMF_CUCI); break;
    case CARS_CAMF_ICS: CARS_PREDICT(M_CAMF_ICS); break;
    case CARS_CAMF_LCS: CARS_PREDICT(M_CAMF_LCS); break;
    case CARS_SVDPP: CARS_PREDICT(M_SVDPP); break;
    case CARS_CAMF_MCS: CARS_PREDICT(M_CAMF_MCS); break;
    default: return fail(h, CARS_E_UNSUPPORTED, "predict: model %d", h->d.model);
  }
#undef CARS_PREDICT
  CUDA_TRY(h, cudaGetLastError());
  h->st.kernel_launches += 1;
  return CARS_OK;
}

extern "C" int cars_predict(cars_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx,
                            int32_t bound, double min_rate, double max_rate, double* out) {
  if (!h) return CARS_E_INVALID;
  if (!h->uploaded) return fail(h, CARS_E_STATE, "cars_predict before cars_upload");
  if (h->epoch_pending) return fail(h, CARS_E_STATE, "an epoch is pending; call cars_epoch_wait first");
  if (n < 0 || (n > 0 && (!u || !j || !out))) return fail(h, CARS_E_INVALID, "bad predict arguments");
  if (n == 0) return CARS_OK;
  if (h->multi) return multi_predict(h, n, u, j, ctx, bound, min_rate, max_rate, out);
  CUDA_TRY(h, cudaSetDevice(h->device));
  int32_t *du = nullptr, *dj = nullptr, *dc = nullptr;
  double* d_out = nullptr;
  int rc = CARS_OK;
  cudaError_t e = dev_alloc(&d_out, (size_t)n);
  if (e != cudaSuccess) rc = fail(h, CARS_E_OOM, "cudaMalloc failed: %s", cudaGetErrorString(e));
  if (!rc) rc = predict_device(h, n, u, j, ctx, bound, min_rate, max_rate, d_out, &du, &dj, &dc);
  if (!rc) {
    e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = h->copier.d2h(out, d_out, (size_t)n * 8);
    if (e != cudaSuccess) rc = fail(h, CARS_E_CUDA, "predict copy-back failed: %s", cudaGetErrorString(e));
    h->st.d2h_bytes += n * 8;
  } else {
    cudaStreamSynchronize(h->stream);
  }
  h->mem.free(du); h->mem.free(dj); h->mem.free(dc); h->mem.free(d_out);
  return rc;
}

extern "C" int cars_eval_ratings(cars_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx,
                                 const double* r, double min_rate, double max_rate, double* sum_abs_err,
                                 double* sum_sq_err) {
  if (!h) return CARS_E_INVALID;
  if (n < 0 || (n > 0 && !r) || !sum_abs_err || !sum_sq_err) return fail(h, CARS_E_INVALID, "bad eval arguments");
  std::vector<double> pred;
  try { pred.resize((size_t)n); } catch (...) { return fail(h, CARS_E_OOM, "host allocation failed"); }
  int rc = cars_predict(h, n, u, j, ctx, 1, min_rate, max_rate, pred.data());
  if (rc) return rc;
  // Java accumulates sequentially in test-matrix order (Recommender.java:518-545); keep that order so
  // MAE/RMSE are bit-identical given identical predictions.
  double sa = 0.0, ss = 0.0;
  for (int64_t i = 0; i < n; i++) {
    if (std::isnan(pred[i])) continue;
    double err = std::fabs(r[i] - pred[i]);
    sa += err;
    ss += err * err;
  }
  *sum_abs_err = sa;
  *sum_sq_err = ss;
  return CARS_OK;
}

// ------------------------------------------------------------------------------------------------
// evalRankings scoring + top-N (K6)
// ------------------------------------------------------------------------------------------------
template <int MODEL>
static cudaError_t launch_rank_score(cars_handle* h, int64_t q0, int64_t nq, const int32_t* qu, const int32_t* qc,
                                     int32_t num_cand, const int32_t* cand, double thold, unsigned long long* keys) {
  const int FC = h->m.F < kRankFC ? h->m.F : kRankFC;
  const size_t smem = (size_t)(kRankTQ + kRankTJ) * (FC + 1) * 8;
  cudaError_t e = cudaFuncSetAttribute(rank_score_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((num_cand + kRankTJ - 1) / kRankTJ), (unsigned)((nq + kRankTQ - 1) / kRankTQ));
  rank_score_kernel<MODEL><<<grid, 256, smem, h->stream>>>(h->m, q0, nq, qu, qc, num_cand, cand, thold, keys);
  return cudaGetLastError();
}

extern "C" int cars_rank_topn(cars_handle* h, int64_t num_queries, const int32_t* qu, const int32_t* qc, int32_t num_cand,
                              const int32_t* cand, const int64_t* rated_ptr, const int32_t* rated_items, double bin_thold,
                              int32_t num_recs, int32_t* out_items, double* out_scores, int32_t* out_count,
                              int32_t* out_kept) {
  if (!h) return CARS_E_INVALID;
  if (!h->uploaded) return fail(h, CARS_E_STATE, "cars_rank_topn before cars_upload");
  if (h->epoch_pending) return fail(h, CARS_E_STATE, "an epoch is pending; call cars_epoch_wait first");
  const bool has_ctx = model_has_ctx(h->d.model);
  if (num_queries < 0 || num_cand < 0 || num_recs <= 0 || num_recs > 4096)
    return fail(h, CARS_E_INVALID, "bad sizes (num_recs must be in 1..4096)");
  if (num_queries > 0 && (!qu || (has_ctx && !qc) || !out_items || !out_scores || !out_count || !out_kept))
    return fail(h, CARS_E_INVALID, "NULL argument");
  if (num_cand > 0 && !cand) return fail(h, CARS_E_INVALID, "cand is NULL");
  if (num_queries == 0) return CARS_OK;
  if (h->multi)
    return multi_rank_topn(h, num_queries, qu, qc, num_cand, cand, rated_ptr, rated_items, bin_thold, num_recs, out_items,
                           out_scores, out_count, out_kept);
  const int32_t I = h->d.num_items;
  for (int64_t q = 0; q < num_queries; q++)
    if ((unsigned)qu[q] >= (unsigned)h->d.num_users || (has_ctx && (unsigned)qc[q] >= (unsigned)h->d.num_contexts))
      return fail(h, CARS_E_INVALID, "query %lld has an id out of range", (long long)q);
  std::vector<int32_t> cidx;
  try { cidx.assign((size_t)I, -1); } catch (...) { return fail(h, CARS_E_OOM, "host allocation failed"); }
  for (int32_t c = 0; c < num_cand; c++) {
    if ((unsigned)cand[c] >= (unsigned)I) return fail(h, CARS_E_INVALID, "candidate %d is not an item id", c);
    if (cidx[cand[c]] >= 0) return fail(h, CARS_E_INVALID, "candidate item %d listed twice", cand[c]);
    cidx[cand[c]] = c;
  }
  const int64_t nrated = rated_ptr ? rated_ptr[num_queries] : 0;
  if (rated_ptr) {
    if (nrated > 0 && !rated_items) return fail(h, CARS_E_INVALID, "rated_items is NULL");
    for (int64_t i = 0; i < nrated; i++)
      if ((unsigned)rated_items[i] >= (unsigned)I) return fail(h, CARS_E_INVALID, "rated item %lld out of range", (long long)i);
  }
  for (int64_t q = 0; q < num_queries; q++) { out_count[q] = 0; out_kept[q] = 0; }
  if (num_cand == 0) return CARS_OK;
  CUDA_TRY(h, cudaSetDevice(h->device));
  {
    int rce = ensure_std(h);
    if (rce) return rce;
  }

  int32_t *d_qu = nullptr, *d_qc = nullptr, *d_cand = nullptr, *d_cidx = nullptr, *d_rated = nullptr, *d_items = nullptr,
          *d_count = nullptr, *d_kept = nullptr;
  int64_t* d_rptr = nullptr;
  double* d_scores = nullptr;
  unsigned long long* d_keys = nullptr;
  // queries per pass: the key matrix [chunk x num_cand] is kept under ~1 GiB
  int64_t chunk = (int64_t)((1ull << 30) / ((uint64_t)num_cand * 8));
  if (chunk < kRankTQ) chunk = kRankTQ;
  if (chunk > num_queries) chunk = num_queries;
  if (chunk > 65535ll * kRankTQ) chunk = 65535ll * kRankTQ;
  int rc = CARS_OK;
  cudaError_t e = cudaSuccess;
#define RK(x) do { if (e == cudaSuccess) e = (x); } while (0)
  RK(dev_alloc(&d_qu, (size_t)num_queries));
  if (has_ctx) RK(dev_alloc(&d_qc, (size_t)num_queries));
  RK(dev_alloc(&d_cand, (size_t)num_cand));
  RK(dev_alloc(&d_cidx, (size_t)I));
  RK(dev_alloc(&d_items, (size_t)num_queries * num_recs));
  RK(dev_alloc(&d_scores, (size_t)num_queries * num_recs));
  RK(dev_alloc(&d_count, (size_t)num_queries));
  RK(dev_alloc(&d_kept, (size_t)num_queries));
  RK(dev_alloc(&d_keys, (size_t)chunk * num_cand));
  if (rated_ptr) {
    RK(dev_alloc(&d_rptr, (size_t)num_queries + 1));
    RK(dev_alloc(&d_rated, (size_t)nrated));
  }
  RK(cudaMemcpyAsync(d_qu, qu, num_queries * 4, cudaMemcpyHostToDevice, h->stream));
  if (has_ctx) RK(cudaMemcpyAsync(d_qc, qc, num_queries * 4, cudaMemcpyHostToDevice, h->stream));
  RK(cudaMemcpyAsync(d_cand, cand, (size_t)num_cand * 4, cudaMemcpyHostToDevice, h->stream));
  RK(cudaMemcpyAsync(d_cidx, cidx.data(), (size_t)I * 4, cudaMemcpyHostToDevice, h->stream));
  RK(cudaMemsetAsync(d_count, 0, (size_t)num_queries * 4, h->stream));
  RK(cudaMemsetAsync(d_kept, 0, (size_t)num_queries * 4, h->stream));
  RK(cudaMemsetAsync(d_items, 0xff, (size_t)num_queries * num_recs * 4, h->stream));
  RK(cudaMemsetAsync(d_scores, 0, (size_t)num_queries * num_recs * 8, h->stream));
  if (rated_ptr) {
    RK(cudaMemcpyAsync(d_rptr, rated_ptr, ((size_t)num_queries + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    if (nrated) RK(cudaMemcpyAsync(d_rated, rated_items, (size_t)nrated * 4, cudaMemcpyHostToDevice, h->stream));
  }
  h->st.h2d_bytes += num_queries * (has_ctx ? 8 : 4) + (int64_t)num_cand * 4 + (int64_t)I * 4 + nrated * 4;
  for (int64_t q0 = 0; q0 < num_queries && e == cudaSuccess; q0 += chunk) {
    const int64_t nq = num_queries - q0 < chunk ? num_queries - q0 : chunk;
    switch (h->d.model) {
      case CARS_PMF: RK(launch_rank_score<M_PMF>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_BIASEDMF: RK(launch_rank_score<M_BIASEDMF>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_CAMF_C: RK(launch_rank_score<M_CAMF_C>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_CAMF_CI: RK(launch_rank_score<M_CAMF_CI>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_CAMF_CU: RK(launch_rank_score<M_CAMF_CU>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_CAMF_CUCI: RK(launch_rank_score<M_CAMF_CUCI>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_CAMF_ICS: RK(launch_rank_score<M_CAMF_ICS>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_CAMF_LCS: RK(launch_rank_score<M_CAMF_LCS>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_SVDPP: RK(launch_rank_score<M_SVDPP>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      case CARS_CAMF_MCS: RK(launch_rank_score<M_CAMF_MCS>(h, q0, nq, d_qu, d_qc, num_cand, d_cand, bin_thold, d_keys)); break;
      default: rc = fail(h, CARS_E_UNSUPPORTED, "rank: model %d", h->d.model); break;
    }
    if (rc) break;
    if (rated_ptr && e == cudaSuccess) {
      rank_drop_rated_kernel<<<(unsigned)nq, 256, 0, h->stream>>>(q0, nq, d_rptr, d_rated, d_cidx, num_cand, d_keys);
      RK(cudaGetLastError());
      h->st.kernel_launches += 1;
    }
    if (e == cudaSuccess) {
      if (num_recs <= kRankRegK) {  // one pass over the keys: per-thread top-K in registers, merged in shared memory
        const size_t sm = (size_t)256 * num_recs * 12;
#define CARS_SELECT(KK)                                                                                                        \
  RK(cudaFuncSetAttribute(rank_select_topk_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * kRankRegK * 12));   \
  rank_select_topk_kernel<KK><<<(unsigned)nq, 256, sm, h->stream>>>(q0, nq, num_cand, d_cand, num_recs, d_keys, d_items, d_scores, \
                                                                    d_count, d_kept)
        if (num_recs <= 4) { CARS_SELECT(4); }
        else if (num_recs <= 8) { CARS_SELECT(8); }
        else if (num_recs <= 10) { CARS_SELECT(10); }
        else { CARS_SELECT(16); }
#undef CARS_SELECT
      } else
        rank_select_kernel<<<(unsigned)nq, 256, 0, h->stream>>>(q0, nq, num_cand, d_cand, num_recs, d_keys, d_items, d_scores,
                                                               d_count, d_kept);
      RK(cudaGetLastError());
      h->st.kernel_launches += 2;
    }
  }
  RK(cudaMemcpyAsync(out_items, d_items, (size_t)num_queries * num_recs * 4, cudaMemcpyDeviceToHost, h->stream));
  RK(cudaMemcpyAsync(out_scores, d_scores, (size_t)num_queries * num_recs * 8, cudaMemcpyDeviceToHost, h->stream));
  RK(cudaMemcpyAsync(out_count, d_count, (size_t)num_queries * 4, cudaMemcpyDeviceToHost, h->stream));
  RK(cudaMemcpyAsync(out_kept, d_kept, (size_t)num_queries * 4, cudaMemcpyDeviceToHost, h->stream));
  cudaError_t es = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = es;
#undef RK
  h->st.d2h_bytes += num_queries * ((int64_t)num_recs * 12 + 8);
  h->mem.free(d_qu); h->mem.free(d_qc); h->mem.free(d_cand); h->mem.free(d_cidx); h->mem.free(d_rated); h->mem.free(d_items);
  h->mem.free(d_count); h->mem.free(d_kept); h->mem.free(d_rptr); h->mem.free(d_scores); h->mem.free(d_keys);
  if (rc) return rc;
  if (e != cudaSuccess)
    return fail(h, e == cudaErrorMemoryAllocation ? CARS_E_OOM : CARS_E_CUDA, "cars_rank_topn failed: %s", cudaGetErrorString(e));
  return CARS_OK;
}

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
#include "multi_gpu.cuh"

extern "C" void cars_destroy(cars_handle* h) {
  if (!h) return;
  if (h->multi) { multi_destroy(h); return; }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->mem.free(h->d_ctx_tab);
  h->mem.free(h->d_u); h->mem.free(h->d_j); h->mem.free(h->d_ctx); h->mem.free(h->d_r); h->mem.free(h->d_level_start);
  h->mem.free(h->m.P); h->mem.free(h->m.Q); h->mem.free(h->m.user_bias); h->mem.free(h->m.item_bias);
  h->mem.free(h->m.cond_bias); h->mem.free(h->m.ic_bias); h->mem.free(h->m.uc_bias);
  h->mem.free(h->d_rec); h->mem.free(h->d_chunk_start); h->mem.free(h->d_chunk_loss); h->mem.free(h->d_flags);
  h->mem.free(h->d_item_old); h->mem.free(h->d_item_scale); h->mem.free(h->d_cond_scale);
  h->mem.free(h->d_hot_slot); h->mem.free(h->d_hot_items);
  h->mem.free(h->tm.Pt); h->mem.free(h->tm.Qt);
  h->mem.free(h->m.cc_sim); h->mem.free(h->m.cf_lcs); h->mem.free(h->m.c_mcs); h->mem.free(h->m.Y); h->mem.free(h->d_ui_ptr); h->mem.free(h->d_ui_items); h->mem.free(h->d_empty_cond);
  h->mem.free(h->d_barrier); h->mem.free(h->d_partial); h->mem.free(h->d_loss);
  if (h->stream) cudaStreamSynchronize(h->stream);  // the pool's frees are stream-ordered
  if (h->h_loss) cudaFreeHost(h->h_loss);
  h->copier.destroy();
  if (h->ev_beg) cudaEventDestroy(h->ev_beg);
  if (h->ev_end) cudaEventDestroy(h->ev_end);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" const char* cars_last_error(const cars_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int cars_get_stats(const cars_handle* h, cars_stats* out) {
  if (!h || !out) return CARS_E_INVALID;
  if (h->multi) { multi_stats(h, out); return CARS_OK; }
  *out = h->st;
  out->num_gpus = 1;
  return CARS_OK;
}

extern "C" void* cars_get_stream(const cars_handle* h) { return h ? (void*)h->stream : nullptr; }

extern "C" int cars_device_count(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < ndev; d++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
  }
  return ok;
}

extern "C" const char* cars_version(void) { return "carskit_b200 abi 3, sm_100a, fp64 SGD (EXACT serial-equivalent / FAST hogwild), FM ALS, top-N"; }
