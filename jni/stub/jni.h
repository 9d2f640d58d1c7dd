/* jni/stub/jni.h -- NOT the JDK's jni.h.  A stand-in with the types and the JNIEnv members jni/carskit_b200_jni.c uses,
 * so that the glue can be compiled, linked and driven by a fake JNIEnv (tests/jni_fake_env.c) in an image without a JDK.
 * Names, signatures and constants follow the JNI specification (chapter 4, "JNI Functions"); the ORDER of the members of
 * JNINativeInterface_ does not (the real table has ~230 slots), so an object built against this header must never be
 * loaded into a real JVM: build the shipping glue with -I$JAVA_HOME/include -I$JAVA_HOME/include/linux instead. */
#ifndef CARSKIT_STUB_JNI_H
#define CARSKIT_STUB_JNI_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_FALSE 0
#define JNI_TRUE 1

typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef unsigned char jboolean;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv* env, const char* name);
  jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
  jboolean (*ExceptionCheck)(JNIEnv* env);
  jsize (*GetArrayLength)(JNIEnv* env, jarray array);
  void* (*GetPrimitiveArrayCritical)(JNIEnv* env, jarray array, jboolean* isCopy);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv* env, jarray array, void* carray, jint mode);
  jdoubleArray (*NewDoubleArray)(JNIEnv* env, jsize len);
  void (*SetDoubleArrayRegion)(JNIEnv* env, jdoubleArray array, jsize start, jsize len, const jdouble* buf);
  jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
};
#endif
