/* carskit_b200_jni.c -- the C half of java/carskit/b200/Native.java: one JNI function per native method, each a thin
 * wrapper of one entry point of include/carskit_b200.h.
 *
 *   - arrays are reached with GetPrimitiveArrayCritical / ReleasePrimitiveArrayCritical around the single C call that
 *     reads or writes them (the library copies to / from the device inside the call and keeps no host pointer);
 *     read-only arrays are released with JNI_ABORT (nothing to copy back);
 *   - a failing cars_* call becomes java.lang.RuntimeException(cars_last_error(...)), which Recommender.run() logs like
 *     any other algorithm failure (src/carskit/generic/Recommender.java:1162-1171); a NaN / Inf loss is RETURNED, so the
 *     Java check in isConverged() (IterativeRecommender.java:181-184) still fires;
 *   - no global state: CARSKit runs the K folds of `cv -p on` on K threads (CARSKit.java:395-412), one handle each.
 *
 * Build (on a box with a JDK):  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *                                   jni/carskit_b200_jni.c -Lcarskit_b200 -lcarskit_b200 -o libcarskit_b200_jni.so
 * Here (no JDK): compiled against jni/stub/jni.h and driven by a fake JNIEnv, tests/test_jni.py. */
#include <jni.h>
#include <string.h>

#include "carskit_b200.h"

static void throw_runtime(JNIEnv* env, const char* msg) {
  jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (cls) (*env)->ThrowNew(env, cls, msg ? msg : "carskit_b200: unknown error");
}

/* A Java primitive array pinned for the duration of one C call; NULL arrays stay NULL. */
typedef struct {
  jarray arr;
  void* p;
} pin_t;

static pin_t pin(JNIEnv* env, jarray a) {
  pin_t x;
  x.arr = a;
  x.p = a ? (*env)->GetPrimitiveArrayCritical(env, a, NULL) : NULL;
  return x;
}
static void unpin(JNIEnv* env, pin_t x, jint mode) {
  if (x.arr && x.p) (*env)->ReleasePrimitiveArrayCritical(env, x.arr, x.p, mode);
}
static jsize len_of(JNIEnv* env, jarray a) { return a ? (*env)->GetArrayLength(env, a) : 0; }

JNIEXPORT jint JNICALL Java_carskit_b200_Native_deviceCount(JNIEnv* env, jclass cls) {
  (void)env; (void)cls;
  return cars_device_count();
}

JNIEXPORT jstring JNICALL Java_carskit_b200_Native_version(JNIEnv* env, jclass cls) {
  (void)cls;
  return (*env)->NewStringUTF(env, cars_version());
}

JNIEXPORT jlong JNICALL Java_carskit_b200_Native_create(JNIEnv* env, jclass cls, jint model, jint mode, jint numUsers,
                                                        jint numItems, jint numConditions, jint numContexts, jint numFactors,
                                                        jintArray u, jintArray j, jintArray ctx, jdoubleArray r,
                                                        jintArray ctxPtr, jintArray ctxCond, jdouble globalMean, jdouble regU,
                                                        jdouble regI, jdouble regB, jdouble regC, jintArray gpuIds,
                                                        jint combine, jdouble fastMaxConc, jintArray emptyConditions,
                                                        jint numContextFactors) {
  (void)cls;
  cars_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = CARS_ABI_VERSION;
  d.model = model;
  d.mode = mode;
  d.schedule = CARS_SCHED_FLAGGED;
  d.num_users = numUsers; d.num_items = numItems; d.num_conditions = numConditions; d.num_contexts = numContexts;
  d.num_factors = numFactors;
  d.nnz = len_of(env, r);
  d.global_mean = globalMean;
  d.reg_u = regU; d.reg_i = regI; d.reg_b = regB; d.reg_c = regC;
  d.combine = combine;
  d.fast_max_conc = fastMaxConc;
  if (len_of(env, u) != d.nnz || len_of(env, j) != d.nnz || (ctx && len_of(env, ctx) != d.nnz)) {
    throw_runtime(env, "carskit_b200: u / j / ctx / r must have the same length");
    return 0;
  }
  const jsize nempty = len_of(env, emptyConditions);
  const jsize ngpu = len_of(env, gpuIds);
  pin_t pu = pin(env, u), pj = pin(env, j), pc = pin(env, ctx), pr = pin(env, r), pp = pin(env, ctxPtr), pq = pin(env, ctxCond),
        pg = pin(env, gpuIds), pe = pin(env, emptyConditions);
  d.empty_conditions = (const int32_t*)pe.p;
  d.num_empty_conditions = nempty;
  d.num_context_factors = numContextFactors; /* CAMF_LCS */
  d.num_context_dims = nempty;               /* CAMF_MCS: one "na" condition per context dimension */
  d.u = (const int32_t*)pu.p; d.j = (const int32_t*)pj.p; d.ctx = (const int32_t*)pc.p; d.r = (const double*)pr.p;
  d.ctx_ptr = (const int32_t*)pp.p; d.ctx_cond = (const int32_t*)pq.p;
  if (ngpu > 1) {
    d.num_gpus = ngpu;
    d.gpu_ids = (const int32_t*)pg.p;
  } else if (ngpu == 1) {
    d.device = ((const int32_t*)pg.p)[0];
  }
  cars_handle* h = NULL;
  const int rc = cars_create(&d, &h);
  char msg[512];
  if (rc != CARS_OK) { strncpy(msg, cars_last_error(NULL), sizeof msg - 1); msg[sizeof msg - 1] = 0; }
  unpin(env, pe, JNI_ABORT); unpin(env, pg, JNI_ABORT); unpin(env, pq, JNI_ABORT); unpin(env, pp, JNI_ABORT); unpin(env, pr, JNI_ABORT);
  unpin(env, pc, JNI_ABORT); unpin(env, pj, JNI_ABORT); unpin(env, pu, JNI_ABORT);
  if (rc != CARS_OK) {  /* thrown only after every critical region is closed: no JNI call is allowed inside one */
    throw_runtime(env, msg);
    return 0;
  }
  return (jlong)(intptr_t)h;
}

static void transfer(JNIEnv* env, jlong handle, int to_device, jdoubleArray P, jdoubleArray Q, jdoubleArray userBias,
                     jdoubleArray itemBias, jdoubleArray condBias, jdoubleArray icBias, jdoubleArray ucBias, jdoubleArray ccSim) {
  cars_handle* h = (cars_handle*)(intptr_t)handle;
  pin_t p[8] = {pin(env, P), pin(env, Q), pin(env, userBias), pin(env, itemBias), pin(env, condBias), pin(env, icBias),
                pin(env, ucBias), pin(env, ccSim)};
  cars_model_arrays a;
  a.P = (double*)p[0].p; a.Q = (double*)p[1].p; a.user_bias = (double*)p[2].p; a.item_bias = (double*)p[3].p;
  a.cond_bias = (double*)p[4].p; a.ic_bias = (double*)p[5].p; a.uc_bias = (double*)p[6].p;
  /* the similarity model's own array: the engine reads the member its model has (ccMatrix_ICS / cfMatrix_LCS / cVector_MCS / SVD++'s Y) */
  a.cc_sim = a.cf_lcs = a.c_mcs = a.Y = (double*)p[7].p;
  const int rc = to_device ? cars_upload(h, &a) : cars_download(h, &a);
  for (int k = 7; k >= 0; k--) unpin(env, p[k], to_device ? JNI_ABORT : 0);  /* download: copy back / commit */
  if (rc != CARS_OK) throw_runtime(env, cars_last_error(h));
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_upload(JNIEnv* env, jclass cls, jlong h, jdoubleArray P, jdoubleArray Q,
                                                       jdoubleArray userBias, jdoubleArray itemBias, jdoubleArray condBias,
                                                       jdoubleArray icBias, jdoubleArray ucBias, jdoubleArray ccSim) {
  (void)cls;
  transfer(env, h, 1, P, Q, userBias, itemBias, condBias, icBias, ucBias, ccSim);
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_download(JNIEnv* env, jclass cls, jlong h, jdoubleArray P, jdoubleArray Q,
                                                         jdoubleArray userBias, jdoubleArray itemBias, jdoubleArray condBias,
                                                         jdoubleArray icBias, jdoubleArray ucBias, jdoubleArray ccSim) {
  (void)cls;
  transfer(env, h, 0, P, Q, userBias, itemBias, condBias, icBias, ucBias, ccSim);
}

JNIEXPORT jdouble JNICALL Java_carskit_b200_Native_epoch(JNIEnv* env, jclass cls, jlong handle, jdouble lRate) {
  (void)cls;
  cars_handle* h = (cars_handle*)(intptr_t)handle;
  double loss = 0.0;
  if (cars_epoch(h, lRate, &loss) != CARS_OK) {
    throw_runtime(env, cars_last_error(h));
    return 0.0;
  }
  return loss; /* NaN / Inf are returned: IterativeRecommender.java:181-184 decides */
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_predict(JNIEnv* env, jclass cls, jlong handle, jintArray u, jintArray j,
                                                        jintArray ctx, jboolean bound, jdouble minRate, jdouble maxRate,
                                                        jdoubleArray out) {
  (void)cls;
  cars_handle* h = (cars_handle*)(intptr_t)handle;
  const jsize n = len_of(env, u);
  if (len_of(env, j) != n || len_of(env, out) != n || (ctx && len_of(env, ctx) != n)) {
    throw_runtime(env, "carskit_b200: predict arrays must have the same length");
    return;
  }
  pin_t pu = pin(env, u), pj = pin(env, j), pc = pin(env, ctx), po = pin(env, out);
  const int rc = cars_predict(h, n, (const int32_t*)pu.p, (const int32_t*)pj.p, (const int32_t*)pc.p, bound ? 1 : 0, minRate,
                              maxRate, (double*)po.p);
  unpin(env, po, 0); unpin(env, pc, JNI_ABORT); unpin(env, pj, JNI_ABORT); unpin(env, pu, JNI_ABORT);
  if (rc != CARS_OK) throw_runtime(env, cars_last_error(h));
}

JNIEXPORT jdoubleArray JNICALL Java_carskit_b200_Native_evalRatings(JNIEnv* env, jclass cls, jlong handle, jintArray u,
                                                                    jintArray j, jintArray ctx, jdoubleArray r, jdouble minRate,
                                                                    jdouble maxRate) {
  (void)cls;
  cars_handle* h = (cars_handle*)(intptr_t)handle;
  const jsize n = len_of(env, u);
  double sums[2] = {0.0, 0.0};
  pin_t pu = pin(env, u), pj = pin(env, j), pc = pin(env, ctx), pr = pin(env, r);
  const int rc = cars_eval_ratings(h, n, (const int32_t*)pu.p, (const int32_t*)pj.p, (const int32_t*)pc.p, (const double*)pr.p,
                                   minRate, maxRate, &sums[0], &sums[1]);
  unpin(env, pr, JNI_ABORT); unpin(env, pc, JNI_ABORT); unpin(env, pj, JNI_ABORT); unpin(env, pu, JNI_ABORT);
  if (rc != CARS_OK) {
    throw_runtime(env, cars_last_error(h));
    return NULL;
  }
  jdoubleArray out = (*env)->NewDoubleArray(env, 2);
  if (out) (*env)->SetDoubleArrayRegion(env, out, 0, 2, sums);
  return out;
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_rankTopN(JNIEnv* env, jclass cls, jlong handle, jintArray qu, jintArray qc,
                                                         jintArray cand, jlongArray ratedPtr, jintArray ratedItems,
                                                         jdouble binThold, jint numRecs, jintArray outItems,
                                                         jdoubleArray outScores, jintArray outCount, jintArray outKept) {
  (void)cls;
  cars_handle* h = (cars_handle*)(intptr_t)handle;
  const jsize nq = len_of(env, qu);
  if (len_of(env, outCount) != nq || len_of(env, outKept) != nq || (jlong)len_of(env, outItems) != (jlong)nq * numRecs ||
      (jlong)len_of(env, outScores) != (jlong)nq * numRecs || (ratedPtr && len_of(env, ratedPtr) != nq + 1)) {
    throw_runtime(env, "carskit_b200: rankTopN output arrays have the wrong length");
    return;
  }
  pin_t a = pin(env, qu), b = pin(env, qc), c = pin(env, cand), d = pin(env, ratedPtr), e = pin(env, ratedItems),
        f = pin(env, outItems), g = pin(env, outScores), k = pin(env, outCount), l = pin(env, outKept);
  const int rc = cars_rank_topn(h, nq, (const int32_t*)a.p, (const int32_t*)b.p, len_of(env, cand), (const int32_t*)c.p,
                                (const int64_t*)d.p, (const int32_t*)e.p, binThold, numRecs, (int32_t*)f.p, (double*)g.p,
                                (int32_t*)k.p, (int32_t*)l.p);
  unpin(env, l, 0); unpin(env, k, 0); unpin(env, g, 0); unpin(env, f, 0);
  unpin(env, e, JNI_ABORT); unpin(env, d, JNI_ABORT); unpin(env, c, JNI_ABORT); unpin(env, b, JNI_ABORT); unpin(env, a, JNI_ABORT);
  if (rc != CARS_OK) throw_runtime(env, cars_last_error(h));
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_destroy(JNIEnv* env, jclass cls, jlong handle) {
  (void)env; (void)cls;
  cars_destroy((cars_handle*)(intptr_t)handle);
}

/* ---- FM ------------------------------------------------------------------------------------------------------------- */
JNIEXPORT jlong JNICALL Java_carskit_b200_Native_fmCreate(JNIEnv* env, jclass cls, jint numUsers, jint numItems,
                                                          jint numConditions, jint numContexts, jint numFactors,
                                                          jint numContextDims, jintArray u, jintArray j, jintArray ctx,
                                                          jdoubleArray r, jdouble regLw, jdouble regLf, jint device) {
  (void)cls;
  cars_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = CARS_ABI_VERSION;
  d.model = CARS_FM;
  d.device = device;
  d.num_users = numUsers; d.num_items = numItems; d.num_conditions = numConditions; d.num_contexts = numContexts;
  d.num_factors = numFactors; d.num_context_dims = numContextDims;
  d.nnz = len_of(env, r);
  d.reg_lw = regLw; d.reg_lf = regLf;
  pin_t pu = pin(env, u), pj = pin(env, j), pc = pin(env, ctx), pr = pin(env, r);
  d.u = (const int32_t*)pu.p; d.j = (const int32_t*)pj.p; d.ctx = (const int32_t*)pc.p; d.r = (const double*)pr.p;
  cars_fm_handle* h = NULL;
  const int rc = cars_fm_create(&d, &h);
  char msg[512];
  if (rc != CARS_OK) { strncpy(msg, cars_fm_last_error(NULL), sizeof msg - 1); msg[sizeof msg - 1] = 0; }
  unpin(env, pr, JNI_ABORT); unpin(env, pc, JNI_ABORT); unpin(env, pj, JNI_ABORT); unpin(env, pu, JNI_ABORT);
  if (rc != CARS_OK) {
    throw_runtime(env, msg);
    return 0;
  }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_fmUploadAndPrepare(JNIEnv* env, jclass cls, jlong handle, jdouble w0,
                                                                   jdoubleArray w, jdoubleArray V) {
  (void)cls;
  cars_fm_handle* h = (cars_fm_handle*)(intptr_t)handle;
  pin_t pw = pin(env, w), pv = pin(env, V);
  cars_fm_arrays a;
  double w0v = w0;
  a.w0 = &w0v; a.w = (double*)pw.p; a.V = (double*)pv.p;
  int rc = cars_fm_upload(h, &a);
  unpin(env, pv, JNI_ABORT); unpin(env, pw, JNI_ABORT);
  if (rc == CARS_OK) rc = cars_fm_prepare(h);
  if (rc != CARS_OK) throw_runtime(env, cars_fm_last_error(h));
}

JNIEXPORT jdouble JNICALL Java_carskit_b200_Native_fmIteration(JNIEnv* env, jclass cls, jlong handle) {
  (void)cls;
  cars_fm_handle* h = (cars_fm_handle*)(intptr_t)handle;
  double loss = 0.0;
  if (cars_fm_iteration(h, &loss) != CARS_OK) {
    throw_runtime(env, cars_fm_last_error(h));
    return 0.0;
  }
  return loss;
}

JNIEXPORT jdouble JNICALL Java_carskit_b200_Native_fmDownload(JNIEnv* env, jclass cls, jlong handle, jdoubleArray w,
                                                              jdoubleArray V) {
  (void)cls;
  cars_fm_handle* h = (cars_fm_handle*)(intptr_t)handle;
  pin_t pw = pin(env, w), pv = pin(env, V);
  cars_fm_arrays a;
  double w0 = 0.0;
  a.w0 = &w0; a.w = (double*)pw.p; a.V = (double*)pv.p;
  const int rc = cars_fm_download(h, &a);
  unpin(env, pv, 0); unpin(env, pw, 0);
  if (rc != CARS_OK) throw_runtime(env, cars_fm_last_error(h));
  return w0;
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_fmPredict(JNIEnv* env, jclass cls, jlong handle, jintArray u, jintArray j,
                                                          jintArray ctx, jboolean bound, jdouble minRate, jdouble maxRate,
                                                          jdoubleArray out) {
  (void)cls;
  cars_fm_handle* h = (cars_fm_handle*)(intptr_t)handle;
  const jsize n = len_of(env, u);
  pin_t pu = pin(env, u), pj = pin(env, j), pc = pin(env, ctx), po = pin(env, out);
  const int rc = cars_fm_predict(h, n, (const int32_t*)pu.p, (const int32_t*)pj.p, (const int32_t*)pc.p, bound ? 1 : 0, minRate,
                                 maxRate, (double*)po.p);
  unpin(env, po, 0); unpin(env, pc, JNI_ABORT); unpin(env, pj, JNI_ABORT); unpin(env, pu, JNI_ABORT);
  if (rc != CARS_OK) throw_runtime(env, cars_fm_last_error(h));
}

JNIEXPORT void JNICALL Java_carskit_b200_Native_fmDestroy(JNIEnv* env, jclass cls, jlong handle) {
  (void)env; (void)cls;
  cars_fm_destroy((cars_fm_handle*)(intptr_t)handle);
}
