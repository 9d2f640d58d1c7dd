/*
 * carskit_b200.h -- C ABI of the B200-native engine for the CARSKit SGD hot path.
 *
 * The reference (irecsys/CARSKit, 100 % Java) has no FFI; the plug-in point is the
 * template-method hook `protected void buildModel()` (src/carskit/generic/Recommender.java:1088)
 * that Recommender.execute() calls between initModel() and evalRatings()/evalRankings()
 * (Recommender.java:319-346).  A subclass such as CAMF_CI_B200 overrides only buildModel(), flattens
 * the Java containers once and calls the entry points below through a 1:1 JNI wrapper
 * (see INTEGRATION.md).  Every entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are caller-owned HOST memory, valid for the duration
 *     of the call; the handle owns every device allocation.
 *   - return 0 on success, a negative CARS_E_* code otherwise; message via cars_last_error().
 *     The library never calls exit(); a NaN/Inf loss is *returned* so the Java-side check
 *     (IterativeRecommender.java:181-184) still fires.
 *   - re-entrant per handle (CARSKit runs K cross-validation folds on K threads,
 *     src/carskit/main/CARSKit.java:395-412); no global mutable state.
 *   - there is NO CPU fallback: without a usable CUDA device cars_create() fails with
 *     CARS_E_NO_DEVICE.
 */
#ifndef CARSKIT_B200_H
#define CARSKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CARS_ABI_VERSION 3 /* 3: FAST mode built; num_gpus / gpu_ids / combine (single-process multi-GPU);
                              fast_max_conc; tuning string (replaces every environment knob) */

/* Recommender classes on the hot path. */
enum cars_model {
  CARS_PMF      = 0, /* src/carskit/alg/baseline/cf/PMF.java:47-82            (2-D `train`, ctx = NULL) */
  CARS_BIASEDMF = 1, /* src/carskit/alg/baseline/cf/BiasedMF.java:58-108      (2-D `train`, ctx = NULL) */
  CARS_CAMF_C   = 2, /* .../cars/adaptation/dependent/dev/CAMF_C.java:74-138  */
  CARS_CAMF_CI  = 3, /* .../cars/adaptation/dependent/dev/CAMF_CI.java:74-131 */
  CARS_CAMF_CU  = 4, /* .../cars/adaptation/dependent/dev/CAMF_CU.java:71-128 */
  CARS_FM       = 5, /* .../cars/adaptation/dependent/FM.java:115-220 (ALS; cars_fm_* entry points below) */
  CARS_CAMF_CUCI = 6, /* .../cars/adaptation/dependent/dev/CAMF_CUCI.java:78-134: ic_bias AND uc_bias, no user /
                        item bias; the reference's Guava tables icBias / ucBias are passed as dense
                        [num_items x C] / [num_users x C] arrays (every cell is initialised, :58-64) */
  CARS_CAMF_ICS = 7, /* .../cars/adaptation/dependent/sim/CAMF_ICS.java:60-129: similarity-based,
                        pred = P[u].Q[j] * prod_d sim(cond_d, empty_d) with ONE condition-similarity matrix shared by all
                        ratings (cc_sim; needs cars_desc.empty_conditions).  Like CAMF_C, EXACT runs on one warp. */
  CARS_CAMF_LCS = 8, /* .../sim/CAMF_LCS.java:66-146: latent context similarity -- sim(cond, empty) = the dot product of
                        two rows of cf_lcs [num_conditions x num_context_factors] (`-f`, CAMF_LCS.java:38) */
  CARS_CAMF_MCS = 9, /* .../sim/CAMF_MCS.java:71-167: multidimensional context similarity -- every condition is a point
                        c_mcs[cond] on its dimension's axis, sim = 1 - Euclidean distance between the context and the
                        all-"na" context; positions clamped to (1e-100, 1/sqrt(num_context_dims)); loss *= 0.05 */
  CARS_SVDPP = 10    /* src/carskit/alg/baseline/cf/SVDPlusPlus.java:55-147: BiasedMF plus the implicit-
                        feedback factors Y [num_items x F] of every item the user rated (2-D `train`, ctx = NULL).  Every
                        rating of user u rewrites Y[k] for ALL items k the user rated: measured DAG width 1.08-1.10
                        (profiles/r2/svdpp_dag_width.txt), so EXACT runs on one warp; userItemsCache = the user's items in
                        ascending order (train.getColumns(u)), derived from u / j by cars_create */
};

/* Update mode.
 * EXACT: serial-equivalent.  Ratings that share a user or an item run in the reference's iteration
 *        order (CRS order of trainMatrix, CAMF_CI.java:80); independent ratings run concurrently.
 *        P, Q and every bias end bit-identical to the Java loop; only `loss` (a 67*nnz-term
 *        sequential fp64 sum in Java) differs, by summation order (~1e-13 relative).
 *        CAMF_C in EXACT mode runs on one warp (every rating touches the shared condBias vector,
 *        CAMF_C.java:107-113), so it is meant for small data.
 * FAST:  hogwild (kernel K2).  NOT serial-equivalent.  The ratings are re-ordered by
 *        user; a user's ratings run in order on one group of lanes with P[u] / userBias[u] / ucBias[u][.]
 *        held privately (no race on the user side), while every item-side cell (Q[j], itemBias[j],
 *        icBias[j][.], condBias[.]) is read without synchronisation and updated with red.global.add.f64.
 *        Throughput does not depend on the skew of the data (EXACT's parallelism is nnz / max item degree).
 *        Rows that many in-flight ratings share (a Zipf-head item, every condBias cell of CAMF_C) would see
 *        the SUM of many stale gradients, i.e. an effective learning rate multiplied by the concurrency;
 *        their step is therefore damped to fast_max_conc expected concurrent updates (DESIGN.md "FAST").
 *        The result differs from the Java loop; tests and bench report the RMSE difference. */
enum cars_mode { CARS_EXACT = 0, CARS_FAST = 1 };

/* How EXACT mode orders independent ratings (all three are serial-equivalent and bit-identical):
 * FLAGGED   (default) ratings are sorted into dependency levels and dealt round-robin to the resident
 *           groups; every rating waits only for the previous rating of its user and of its item through
 *           per-row completion counters (no grid-wide barrier; levels overlap).
 * WAVEFRONT same order, a grid-wide barrier separates levels.
 * DATAFLOW  ratings stay in reference order; same counters; a user's factor row stays in registers across
 *           consecutive ratings of that user (cheapest set-up, fastest on small inputs sorted by user). */
enum cars_schedule { CARS_SCHED_FLAGGED = 0, CARS_SCHED_WAVEFRONT = 1, CARS_SCHED_DATAFLOW = 2 };

enum cars_error {
  CARS_OK            =  0,
  CARS_E_INVALID     = -1, /* bad argument / inconsistent descriptor            */
  CARS_E_NO_DEVICE   = -2, /* no CUDA device, or device is not sm_100           */
  CARS_E_CUDA        = -3, /* a CUDA runtime call failed                        */
  CARS_E_OOM         = -4, /* host or device allocation failed                  */
  CARS_E_UNSUPPORTED = -5, /* model/mode combination not built                  */
  CARS_E_STATE       = -6  /* call order violated (e.g. epoch before upload)    */
};

/* Training-set descriptor: what buildModel() can reach without reflection.
 *   u/j/ctx/r   one entry per training rating, in the reference's iteration order
 *               (`for (MatrixEntry me : trainMatrix)`: row = user-item pair id ascending, column =
 *               context id ascending; u = rateDao.getUserIdFromUI(ui), j = getItemIdFromUI(ui),
 *               DataDAO.java:1038-1046).  For PMF/BiasedMF iterate the 2-D `train` matrix and pass
 *               ctx = NULL.
 *   ctx_ptr/ctx_cond   CSR form of rateDao.getContextConditionsList() (DataDAO.java:1035): the
 *               condition ids of context c are ctx_cond[ctx_ptr[c] .. ctx_ptr[c+1]), in the order
 *               ContextRecommender.getConditions() yields them (ContextRecommender.java:53-61).
 *   reg_*       the static floats regU/regI/regB/regC (IterativeRecommender.java:92-99) and FM's
 *               regLw/regLf (FM.java:53-54) ALREADY widened float->double by the caller
 *               (1e-4f -> 9.999999747378752e-05); never re-parse the decimal string as a double.
 */
typedef struct cars_desc {
  int32_t abi_version;     /* = CARS_ABI_VERSION */
  int32_t model;           /* enum cars_model */
  int32_t mode;            /* enum cars_mode  */
  int32_t device;          /* CUDA device ordinal */
  int32_t num_users;
  int32_t num_items;
  int32_t num_conditions;  /* rateDao.numConditions(), ContextRecommender.java:43 */
  int32_t num_contexts;
  int32_t num_factors;     /* IterativeRecommender.numFactors (:101) */
  int32_t schedule;        /* enum cars_schedule (EXACT mode only) */
  int64_t nnz;
  const int32_t* u;
  const int32_t* j;
  const int32_t* ctx;      /* NULL for PMF / BiasedMF */
  const double*  r;
  const int32_t* ctx_ptr;  /* [num_contexts + 1], NULL when ctx == NULL */
  const int32_t* ctx_cond; /* [ctx_ptr[num_contexts]] */
  double global_mean;      /* Recommender.globalMean (Recommender.java:265) */
  double reg_u, reg_i, reg_b, reg_c, reg_lw, reg_lf;
  int32_t num_context_dims; /* rateDao.numContextDims() (FM.java:86); FM only */
  int32_t num_gpus;         /* 0 or 1: one GPU (`device`).  N > 1: ONE handle drives N GPUs from this process --
                               users are sharded by contiguous range over gpu_ids, the item block is combined with
                               ncclAllReduce inside cars_epoch; `device` and `stream` are ignored */
  int64_t global_nnz;       /* FM row sharding: rows over all ranks; 0 = nnz (single GPU) */
  void*   stream;           /* cudaStream_t to launch on; NULL = the handle creates its own */
  const int32_t* gpu_ids;   /* [num_gpus] CUDA ordinals; NULL = 0 .. num_gpus-1 */
  double  fast_max_conc;    /* FAST: cap on the expected number of concurrent updates of one shared row before
                               its step is damped; 0 = default (4); < 0 = never damp */
  const char* tuning;       /* developer knobs "key=value;key=value" (tests / profiling); NULL in normal use.
                               The library reads NO environment variable. */
  int32_t combine;          /* num_gpus > 1: enum cars_combine */
  int32_t num_empty_conditions; /* CAMF_ICS: length of empty_conditions (= number of context dimensions) */
  const int32_t* empty_conditions; /* CAMF_ICS: rateDao.getEmptyContextConditions() (DataDAO.java:214-215): the "dim:na"
                               condition of every dimension, in dimension order; the i-th condition of a context is
                               compared with empty_conditions[i] (CAMF_ICS.java:56, 88).  CAMF_LCS / CAMF_MCS too */
  int32_t num_context_factors; /* CAMF_LCS: numF, the `-f` option (CAMF_LCS.java:38; default 10) */
} cars_desc;

/* How the item block is combined between user-range shards once per epoch (DESIGN.md "Multi-GPU"):
 *   block <- old + scale_j * sum_over_shards(new_shard - old)
 * MEAN  scale = 1/shards for every row (averages the shards' rows; stable for any shard size)
 * SUM   scale = 1 (adds the shards' steps; only safe while an epoch moves a row a little)
 * TOUCHED  scale_j = 1 / (number of shards that have at least one rating of item j): a row only one shard trained
 *          keeps that shard's full step (sparse item sets), a row every shard trained is averaged */
enum cars_combine { CARS_COMBINE_MEAN = 0, CARS_COMBINE_SUM = 1, CARS_COMBINE_TOUCHED = 2 };

typedef struct cars_handle cars_handle;

/* Model arrays, row-major contiguous, NULL where the model has no such member:
 *   P [num_users x F], Q [num_items x F]              IterativeRecommender.java:55-58
 *   user_bias [num_users], item_bias [num_items]      IterativeRecommender.java:61-63
 *   cond_bias [num_conditions]                        CAMF.java (condBias), CAMF_C.java:60
 *   ic_bias [num_items x num_conditions]              CAMF_CI.java:58
 *   uc_bias [num_users x num_conditions]              CAMF_CU.java:55
 *   cc_sim  [num_conditions x num_conditions]         CAMF_ICS.java:45-48
 *   cf_lcs  [num_conditions x num_context_factors]    CAMF_LCS.java:38-40;   c_mcs [num_conditions]   CAMF_MCS.java:47-48
 *   (CAMF_CUCI: both ic_bias and uc_bias, CAMF_CUCI.java:42-43, 58-64) */
typedef struct cars_model_arrays {
  double* P;
  double* Q;
  double* user_bias;
  double* item_bias;
  double* cond_bias;
  double* ic_bias;
  double* uc_bias;
  double* cc_sim; /* CAMF_ICS: ccMatrix_ICS as a dense [C x C] array (CAMF.java:44, CAMF_ICS.java:45-48).  librec's
                     SymmMatrix keeps ONE cell per unordered pair, (max(i,j), min(i,j)): upload reads that cell of the
                     caller's array, download writes the trained value to BOTH (i,j) and (j,i) */
  double* cf_lcs; /* CAMF_LCS: cfMatrix_LCS [num_conditions x num_context_factors] (CAMF.java:46, CAMF_LCS.java:38-40) */
  double* c_mcs;  /* CAMF_MCS: cVector_MCS [num_conditions] (CAMF.java:47, CAMF_MCS.java:47-48) */
  double* Y;      /* SVD++: implicit-feedback item factors [num_items x F] (SVDPlusPlus.java:37, 48-49) */
} cars_model_arrays;

/* Replaces the set-up a Java buildModel() does implicitly by holding trainMatrix/rateDao: copies the
 * rating arrays to the device and builds the dependency schedule (wavefront levels). */
int cars_create(const cars_desc* desc, cars_handle** out);

/* Hands over the arrays Java's initModel() produced (IterativeRecommender.java:231-247 and
 * CAMF_*.initModel).  May be called again to reset the model. */
int cars_upload(cars_handle* h, const cars_model_arrays* host);

/* One pass over all training ratings == one iteration of the `for (int iter = 1; ...)` loop body of
 * buildModel() up to and including `loss *= 0.5` (CAMF_CI.java:79-124 and siblings).  `lrate` is the
 * instance field lRate.  The caller keeps isConverged(iter)/updateLRate (IterativeRecommender.java:
 * 145-229) in Java and passes the new lRate next time. */
int cars_epoch(cars_handle* h, double lrate, double* loss_out);

/* Asynchronous variant: enqueue the epoch on the handle's stream and return; cars_epoch_wait() blocks
 * and fetches the loss.  cars_epoch == cars_epoch_begin + cars_epoch_wait. */
int cars_epoch_begin(cars_handle* h, double lrate);
int cars_epoch_wait(cars_handle* h, double* loss_out);

/* Copies the trained arrays back so the inherited Java predict()/evalRatings()/evalRankings()
 * (Recommender.java:306-317, 504-594, 672-960) see them. */
int cars_download(cars_handle* h, const cars_model_arrays* host);

/* Batched Recommender.predict(u, j, c, bound) (Recommender.java:306-317) with the model's
 * predict(u,j,c) (CAMF_CI.java:65-72 etc.); sequential-order dot product, so bit-identical to Java.
 * ctx may be NULL for PMF/BiasedMF. */
int cars_predict(cars_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx,
                 int32_t bound, double min_rate, double max_rate, double* out);

/* evalRatings() core (Recommender.java:518-545): sum |err| and sum err^2 over a test set, computed on
 * the device from the resident model; the caller derives MAE/RMSE (:567-570). */
int cars_eval_ratings(cars_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx,
                      const double* r, double min_rate, double max_rate, double* sum_abs_err,
                      double* sum_sq_err);

/* The scoring loop and top-N cut of evalRankings() (Recommender.java:797-824) for a batch of (user, context)
 * queries: every candidate item not in the query's `rated` list (items the user rated in that context in the
 * TRAINING set, :792/:800) is scored with ranking(u, j, c) = predict(u, j, c) (:807, :1016), dropped if NaN or
 * not > bin_thold (:808-811), and the survivors are ordered like Lists.sortList(itemScores, true) (:822): by
 * descending score, ties in candidate order -- `cand` must therefore list candItems in the iteration order of
 * the reference's HashSet (:704).  Scores are bit-identical to Java's, so are the ranked ids.
 *   out_items / out_scores  [num_queries x num_recs] the first min(num_recs, kept) entries are valid
 *   out_count               ranked entries per query (rankedItems.size())
 *   out_kept                itemScores.size() before the cut (numDropped = numCands - rankedItems.size(), :841)
 * rated_ptr may be NULL (nothing excluded); qc may be NULL for PMF / BiasedMF.  The measures themselves
 * (Measures.PrecAt .. RRAt, :852-858) stay with the caller. */
int cars_rank_topn(cars_handle* h, int64_t num_queries, const int32_t* qu, const int32_t* qc, int32_t num_cand,
                   const int32_t* cand, const int64_t* rated_ptr, const int32_t* rated_items, double bin_thold,
                   int32_t num_recs, int32_t* out_items, double* out_scores, int32_t* out_count, int32_t* out_kept);

/* ---- multi-GPU -------------------------------------------------------------------------------------------------
 * Two ways in.  (1) cars_desc.num_gpus > 1: ONE handle drives N GPUs from the calling process -- what a JVM needs
 * (CARSKit is one process, CARSKit.java:392-412).  cars_create shards the users by contiguous range over gpu_ids,
 * cars_upload / cars_download address the caller's FULL arrays (user-side rows by offset), cars_epoch runs the shards'
 * epochs concurrently and combines the item block with ncclAllReduce (NCCL is loaded at run time), cars_predict /
 * cars_eval_ratings / cars_rank_topn route every query to the GPU that owns its user.  The entry points below this
 * comment are NOT used then.  (2) One process and one single-GPU handle per GPU (torchrun, MPI): the caller owns the
 * collective and uses the three entry points below.
 * ---- (2) one process (and one handle) per GPU, users sharded by contiguous range ------------------
 * The reference is single-process; SURVEY.md 8e defines the sharded semantics: every rank trains the
 * ratings of ITS users against its own copy of the item-side arrays (Q, itemBias, icBias -- "the item
 * block"), then the ranks combine  item_block <- old + sum_over_ranks(new_rank - old)  once per epoch.
 * Within a rank the epoch is EXACT (serial-equivalent on the rank's ratings); across ranks it is a
 * block-Jacobi step, not serial-equivalent (DESIGN.md "Multi-GPU").  The library does not link NCCL:
 * the caller owns the DEVICE buffer `dev_delta` (cars_item_block_doubles() doubles, e.g. a torch tensor),
 * all-reduces it (sum) on the handle's stream between the two calls below.
 *   cars_epoch_sharded_begin : snapshot the item block, run the epoch, write (new - old) to dev_delta
 *   cars_epoch_sharded_finish: item block <- old + scale * dev_delta (dev_delta now holds the sum over ranks;
 *                              scale = 1/world averages the ranks' item blocks, scale = 1 sums their
 *                              steps); returns the local loss
 * Layout of the item block: [Q (num_items x Fp, row stride Fp = F rounded up to even) | item_bias | ic_bias],
 * members the model lacks are absent. */
int cars_item_block_doubles(const cars_handle* h, int64_t* out);
int cars_epoch_sharded_begin(cars_handle* h, double lrate, double* dev_delta);
int cars_epoch_sharded_finish(cars_handle* h, const double* dev_delta, double scale, double* loss_out);

void cars_destroy(cars_handle* h);

/* Message of the last failure on this handle (or of the last failed cars_create when h == NULL;
 * thread-local).  Never NULL. */
const char* cars_last_error(const cars_handle* h);

/* Introspection used by the benchmark and the tests (no reference counterpart). */
typedef struct cars_stats {
  int64_t nnz;              /* ratings trained by this handle (after sharding)         */
  int64_t num_levels;       /* WAVEFRONT: levels of the dependency DAG; DATAFLOW: chunks */
  int64_t max_level_size;   /* WAVEFRONT: largest level; DATAFLOW: longest chunk       */
  int64_t kernel_launches;  /* engine kernels launched so far on this handle           */
  int64_t h2d_bytes;        /* bytes copied host->device so far                        */
  int64_t d2h_bytes;        /* bytes copied device->host so far                        */
  double  schedule_ms;      /* wall time spent building the schedule in cars_create (H2D of the ratings included) */
  double  last_epoch_ms;    /* device time of the last epoch's SGD kernel (CUDA events on the
                               launching stream)                                       */
  int32_t grid_ctas;        /* persistent grid of the SGD kernel                       */
  int32_t block_threads;
  int32_t sm_count;
  int32_t reserved;
  /* FLAGGED schedule build, split: the caller's arrays crossing PCIe / chains + dependency levels on the
     device / level sort + record packing (device times, CUDA events) */
  double  schedule_copy_ms;
  double  schedule_levels_ms;
  double  schedule_pack_ms;
  /* FAST mode: the smallest step-damping factor of any item row / condBias cell (1 = nothing damped) and the largest
     item degree (EXACT's parallelism is nnz / max_item_degree) */
  double  fast_min_item_scale;
  double  fast_min_cond_scale;
  int64_t max_item_degree;
  int32_t fast_hot_rows;    /* FAST: item rows whose steps are summed per CTA in shared memory before they reach L2 */
  int32_t num_gpus;         /* devices this handle drives (cars_desc.num_gpus) */
  double  exchange_ms;      /* num_gpus > 1: device time of the last epoch's ncclAllReduce of the item block (shard 0) */
} cars_stats;
int cars_get_stats(const cars_handle* h, cars_stats* out);
void* cars_get_stream(const cars_handle* h); /* cudaStream_t the kernels are launched on */

/* ---- FM (src/carskit/alg/cars/adaptation/dependent/FM.java): ALS factorization machine ------------------
 * Not SGD: coordinate descent with cached residuals over the one-hot features x_u = 1, x_{U+j} = 1,
 * x_{U+I+ctx} = 1/numContextDims (FM.java:76-91; the context feature only exists while its index is < p =
 * numUsers + numItems + numConditions).  Uses cars_desc with model = CARS_FM, u/j/ctx/r in trainMatrix
 * iteration order, num_factors = k, reg_lw / reg_lf = the floats of `FM=-lw .. -lf ..` (FM.java:53-54)
 * widened to double, num_context_dims.  learn.rate and isConverged() are not used by FM.java.
 *   cars_fm_upload     w0, w [p], V [p x k] as FM.initModel() made them (FM.java:57-74)
 *   cars_fm_prepare    the pre-pass of buildModel(): errors[n] = r - predict, Q[n][f] = sum_i V[i][f] x_n[i] (:118-146)
 *   cars_fm_iteration  one iteration of `for (iter ...)`: w0 step, w_l steps, V_lf steps (:148-219).  The
 *                      coordinates of one field touch disjoint rows and are solved concurrently (identical
 *                      to the sequential order); sums are tree-ordered, so results match the reference up to
 *                      summation order.  loss_out = 0.05 * (sum e^2 + regLw w0^2 + sum_l regLw w_l^2): the
 *                      O(size) part of FM.java's `loss`; its O(k p size) reporting term (:206) is never read
 *                      by the reference and is omitted.
 *   cars_fm_predict    FM.predict(u, j, c) (:93-113) + Recommender.predict(.., bound) (Recommender.java:306-317) */
typedef struct cars_fm_handle cars_fm_handle;
typedef struct cars_fm_arrays {
  double* w0; /* [1] */
  double* w;  /* [p] */
  double* V;  /* [p x k] row-major */
} cars_fm_arrays;
typedef struct cars_fm_stats {
  int64_t nnz, p, pieces, kernel_launches, h2d_bytes, d2h_bytes;
  double last_iteration_ms; /* device time of the last cars_fm_iteration (CUDA events) */
} cars_fm_stats;
int cars_fm_create(const cars_desc* desc, cars_fm_handle** out);
int cars_fm_upload(cars_fm_handle* h, const cars_fm_arrays* host);
int cars_fm_prepare(cars_fm_handle* h);
int cars_fm_iteration(cars_fm_handle* h, double* loss_out);
int cars_fm_download(cars_fm_handle* h, const cars_fm_arrays* host);
int cars_fm_predict(cars_fm_handle* h, int64_t n, const int32_t* u, const int32_t* j, const int32_t* ctx,
                    int32_t bound, double min_rate, double max_rate, double* out);
/* Multi-GPU FM: rows sharded by contiguous range (one handle per GPU over its rows; w0, w, V replicated).
 * desc->global_nnz = number of rows over ALL ranks (the `size` of FM.java's denominators); 0 = nnz.
 * Every coordinate step needs the sums over all rows: the engine writes its local per-coordinate sums into
 * `dev_buf` (caller-owned DEVICE memory, >= cars_fm_exchange_doubles() doubles) and calls `allreduce`, which
 * must sum dev_buf[0 .. count) over the ranks in place, ordered on the handle's stream (e.g.
 * torch.distributed.all_reduce under that stream).  3 * (1 + k) + 1 calls per iteration.  All ranks then
 * compute identical new coordinates.  loss_out is the GLOBAL loss. */
typedef int (*cars_allreduce_fn)(void* user, double* dev_ptr, int64_t count);
int cars_fm_exchange_doubles(const cars_fm_handle* h, int64_t* out);
int cars_fm_iteration_sharded(cars_fm_handle* h, double* dev_buf, cars_allreduce_fn allreduce, void* user,
                              double* loss_out);
int cars_fm_get_stats(const cars_fm_handle* h, cars_fm_stats* out);
void* cars_fm_get_stream(const cars_fm_handle* h);
const char* cars_fm_last_error(const cars_fm_handle* h);
void cars_fm_destroy(cars_fm_handle* h);

/* ---- rating data: native ingest, k-fold split, columnar file (host code; csrc/ingest.cpp) ------------------------
 * The arrays cars_desc takes, produced without the JVM holding the ratings in boxed Guava tables (DataDAO.java:174, 342):
 *   cars_dataset_read_binary_csv  DataDAO.readData (DataDAO.java:199-354) over a binary-format ratings file: inner ids by
 *                                 first appearance (users, items, user-item pairs, contexts), condition id = header column,
 *                                 duplicates overwrite, zero ratings dropped, CRS order (pair id, context id)
 *   cars_dataset_from_arrays      wraps arrays that are already in CRS order (e.g. synthetic data)
 *   cars_dataset_save / _load     columnar file: one contiguous column per array, loads with one read per column
 *   cars_dataset_kfold            DataSplitter (DataSplitter.java:68-133): fold labels from java.util.Random(seed),
 *                                 fold k's label = its TEST set, the rest = training set, order kept
 *   cars_dataset_get_view         the pointers / sizes to copy into a cars_desc (owned by the data set) */
typedef struct cars_dataset cars_dataset;
typedef struct cars_dataset_view {
  int32_t num_users, num_items, num_pairs, num_contexts, num_conditions, num_context_dims;
  int64_t nnz;
  const int32_t *u, *j, *ctx, *pair; /* [nnz]; pair = CRS row (user-item pair id); ctx NULL for 2-D data */
  const double* r;
  const int32_t *ctx_ptr, *ctx_cond;
  double global_mean, min_rate, max_rate; /* SparseMatrix.getGlobalAvg; first / last of rateDao.getRatingScale() */
  int32_t num_empty_conditions;
  const int32_t* empty_conditions;        /* rateDao.getEmptyContextConditions() (the "dim:na" columns) */
} cars_dataset_view;
int cars_dataset_read_binary_csv(const char* path, cars_dataset** out);
int cars_dataset_from_arrays(int32_t num_users, int32_t num_items, int32_t num_conditions, int32_t num_contexts,
                             int32_t num_context_dims, int64_t nnz, const int32_t* u, const int32_t* j, const int32_t* ctx,
                             const double* r, const int32_t* ctx_ptr, const int32_t* ctx_cond, cars_dataset** out);
int cars_dataset_save(const cars_dataset* d, const char* path);
int cars_dataset_load(const char* path, cars_dataset** out);
int cars_dataset_kfold(const cars_dataset* d, int32_t kfold, int64_t seed, int32_t fold, cars_dataset** train, cars_dataset** test);
int cars_dataset_get_view(const cars_dataset* d, cars_dataset_view* out);
void cars_dataset_free(cars_dataset* d);
const char* cars_dataset_last_error(void);

/* Library self-description: "carskit_b200 <abi> sm_100a ..." */
const char* cars_version(void);

/* Number of CUDA devices this library can train on (sm_100 only); 0 when there is none (then cars_create fails with
 * CARS_E_NO_DEVICE).  Lets a caller spread cross-validation folds over the GPUs (java/carskit/b200/B200.devicesFor). */
int cars_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CARSKIT_B200_H */
