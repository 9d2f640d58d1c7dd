/* jni_fake_env.c -- TEST INFRASTRUCTURE: a fake JNIEnv (function table over plain C structs) that drives the JNI glue
 * jni/carskit_b200_jni.c exactly as a JVM would call it from java/carskit/b200/B200.train(): create, upload, three
 * epochs, evalRatings, predict, download, destroy, on the same 6-rating toy set as examples/c_client.c -- whose output
 * the GPU test compares with this program's, line for line.  Without an sm_100 device create() must throw
 * RuntimeException("... no CPU path ...") and the program exits 3. */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct _jobject {
  jsize len;
  void* data;
  int critical; /* open GetPrimitiveArrayCritical regions */
};

static char g_exception[600];
static int g_pending = 0, g_open_critical = 0, g_jni_call_inside_critical = 0;
static struct _jobject g_runtime_exception_class;

static void check_not_critical(void) {
  if (g_open_critical) g_jni_call_inside_critical++; /* the JNI spec forbids other JNI calls inside a critical region */
}
static jclass fe_FindClass(JNIEnv* env, const char* name) {
  (void)env;
  check_not_critical();
  return strcmp(name, "java/lang/RuntimeException") == 0 ? &g_runtime_exception_class : NULL;
}
static jint fe_ThrowNew(JNIEnv* env, jclass c, const char* msg) {
  (void)env; (void)c;
  check_not_critical();
  strncpy(g_exception, msg, sizeof g_exception - 1);
  g_pending = 1;
  return 0;
}
static jboolean fe_ExceptionCheck(JNIEnv* env) { (void)env; return (jboolean)g_pending; }
static jsize fe_GetArrayLength(JNIEnv* env, jarray a) { (void)env; return a->len; }
static void* fe_GetCritical(JNIEnv* env, jarray a, jboolean* is_copy) {
  (void)env;
  if (is_copy) *is_copy = JNI_FALSE;
  a->critical++;
  g_open_critical++;
  return a->data;
}
static void fe_ReleaseCritical(JNIEnv* env, jarray a, void* p, jint mode) {
  (void)env; (void)mode;
  if (p != a->data || a->critical <= 0) { fprintf(stderr, "fake env: bad ReleasePrimitiveArrayCritical\n"); exit(2); }
  a->critical--;
  g_open_critical--;
}
static jdoubleArray fe_NewDoubleArray(JNIEnv* env, jsize n) {
  (void)env;
  check_not_critical();
  struct _jobject* o = (struct _jobject*)calloc(1, sizeof *o);
  o->len = n;
  o->data = calloc((size_t)n, sizeof(double));
  return o;
}
static void fe_SetDoubleArrayRegion(JNIEnv* env, jdoubleArray a, jsize start, jsize len, const jdouble* buf) {
  (void)env;
  memcpy((double*)a->data + start, buf, (size_t)len * sizeof(double));
}
static jstring fe_NewStringUTF(JNIEnv* env, const char* s) {
  (void)env;
  struct _jobject* o = (struct _jobject*)calloc(1, sizeof *o);
  o->len = (jsize)strlen(s);
  o->data = malloc((size_t)o->len + 1);
  memcpy(o->data, s, (size_t)o->len + 1);
  return o;
}

static const struct JNINativeInterface_ g_table = {fe_FindClass,       fe_ThrowNew,        fe_ExceptionCheck,
                                                   fe_GetArrayLength,  fe_GetCritical,     fe_ReleaseCritical,
                                                   fe_NewDoubleArray,  fe_SetDoubleArrayRegion, fe_NewStringUTF};

static struct _jobject arr(void* data, jsize n) {
  struct _jobject o;
  o.len = n; o.data = data; o.critical = 0;
  return o;
}

/* the glue's entry points (javah signatures) */
jint Java_carskit_b200_Native_deviceCount(JNIEnv*, jclass);
jstring Java_carskit_b200_Native_version(JNIEnv*, jclass);
jlong Java_carskit_b200_Native_create(JNIEnv*, jclass, jint, jint, jint, jint, jint, jint, jint, jintArray, jintArray, jintArray,
                                      jdoubleArray, jintArray, jintArray, jdouble, jdouble, jdouble, jdouble, jdouble, jintArray,
                                      jint, jdouble, jintArray, jint);
void Java_carskit_b200_Native_upload(JNIEnv*, jclass, jlong, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray,
                                     jdoubleArray, jdoubleArray, jdoubleArray);
void Java_carskit_b200_Native_download(JNIEnv*, jclass, jlong, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray,
                                       jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray);
jdouble Java_carskit_b200_Native_epoch(JNIEnv*, jclass, jlong, jdouble);
void Java_carskit_b200_Native_predict(JNIEnv*, jclass, jlong, jintArray, jintArray, jintArray, jboolean, jdouble, jdouble,
                                      jdoubleArray);
jdoubleArray Java_carskit_b200_Native_evalRatings(JNIEnv*, jclass, jlong, jintArray, jintArray, jintArray, jdoubleArray, jdouble,
                                                  jdouble);
void Java_carskit_b200_Native_destroy(JNIEnv*, jclass, jlong);

int main(void) {
  JNIEnv envp = &g_table;
  JNIEnv* env = &envp;
  int32_t u[] = {0, 0, 1, 1, 2, 2}, j[] = {0, 1, 0, 1, 0, 1}, ctx[] = {0, 1, 1, 0, 0, 1};
  double r[] = {4, 5, 3, 4, 2, 5};
  int32_t ctx_ptr[] = {0, 1, 2}, ctx_cond[] = {0, 1};
  enum { U = 3, I = 2, C = 2, F = 4 };
  double P[U * F], Q[I * F], user_bias[U] = {0.01, -0.02, 0.03}, ic_bias[I * C] = {0.5, 0.25, 0.75, 0.125};
  for (int k = 0; k < U * F; k++) P[k] = 0.1 * ((k % 5) - 2);
  for (int k = 0; k < I * F; k++) Q[k] = 0.05 * ((k % 7) - 3);
  struct _jobject au = arr(u, 6), aj = arr(j, 6), ac = arr(ctx, 6), ar = arr(r, 6), ap = arr(ctx_ptr, 3), aq = arr(ctx_cond, 2);
  struct _jobject aP = arr(P, U * F), aQ = arr(Q, I * F), aub = arr(user_bias, U), aic = arr(ic_bias, I * C);

  printf("deviceCount = %d\n", (int)Java_carskit_b200_Native_deviceCount(env, NULL));
  jlong h = Java_carskit_b200_Native_create(env, NULL, 3 /* Native.CAMF_CI */, 0, U, I, C, 2, F, &au, &aj, &ac, &ar, &ap, &aq, 23.0 / 6.0,
                                            (double)1e-4f, (double)1e-4f, (double)1e-4f, (double)1e-3f, NULL, 0, 0.0, NULL, 0);
  if (g_pending) {
    printf("RuntimeException: %s\n", g_exception);
    return (g_open_critical == 0 && g_jni_call_inside_critical == 0 && h == 0) ? 3 : 2;
  }
  Java_carskit_b200_Native_upload(env, NULL, h, &aP, &aQ, &aub, NULL, NULL, &aic, NULL, NULL);
  for (int iter = 1; iter <= 3 && !g_pending; iter++)
    printf("iter %d: loss = %.17g\n", iter, Java_carskit_b200_Native_epoch(env, NULL, h, (double)0.02f));
  double out[6];
  struct _jobject ao = arr(out, 6);
  Java_carskit_b200_Native_predict(env, NULL, h, &au, &aj, &ac, JNI_TRUE, 1.0, 5.0, &ao);
  jdoubleArray sums = Java_carskit_b200_Native_evalRatings(env, NULL, h, &au, &aj, &ac, &ar, 1.0, 5.0);
  Java_carskit_b200_Native_download(env, NULL, h, &aP, &aQ, &aub, NULL, NULL, &aic, NULL, NULL);
  Java_carskit_b200_Native_destroy(env, NULL, h);
  if (g_pending) { printf("RuntimeException: %s\n", g_exception); return 1; }
  printf("P[0][0] = %.17g, pred[0] = %.17g, sumAbs = %.17g, sumSq = %.17g\n", P[0], out[0], ((double*)sums->data)[0],
         ((double*)sums->data)[1]);
  return (g_open_critical == 0 && g_jni_call_inside_critical == 0) ? 0 : 2;
}
