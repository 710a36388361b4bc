// jni_stub.h — a minimal stand-in for <jni.h> so that jni/se_jni.cpp can be SYNTAX- and TYPE-checked in an image
// without a JDK (g++ -fsyntax-only -DSE_JNI_STUB; tests/test_host_cpu.py).  Declarations follow the JNI specification
// (types, JNIEnv member functions used by the shim); nothing here is linked or executed.
#pragma once
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;
class _jobject {};
typedef _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jobject jbyteArray;
typedef jobject jintArray;
typedef jobject jfloatArray;
typedef jobject jdoubleArray;
typedef jobject jthrowable;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

struct JNIEnv {
  jclass FindClass(const char*);
  jint ThrowNew(jclass, const char*);
  jboolean ExceptionCheck();
  jsize GetArrayLength(jarray);
  jbyteArray NewByteArray(jsize);
  jdoubleArray NewDoubleArray(jsize);
  void GetByteArrayRegion(jbyteArray, jsize, jsize, jbyte*);
  void SetByteArrayRegion(jbyteArray, jsize, jsize, const jbyte*);
  void GetIntArrayRegion(jintArray, jsize, jsize, jint*);
  void GetFloatArrayRegion(jfloatArray, jsize, jsize, jfloat*);
  void SetFloatArrayRegion(jfloatArray, jsize, jsize, const jfloat*);
  void GetDoubleArrayRegion(jdoubleArray, jsize, jsize, jdouble*);
  void SetDoubleArrayRegion(jdoubleArray, jsize, jsize, const jdouble*);
  const char* GetStringUTFChars(jstring, jboolean*);
  void ReleaseStringUTFChars(jstring, const char*);
  jobject NewDirectByteBuffer(void*, jlong);
  void* GetDirectBufferAddress(jobject);
  jlong GetDirectBufferCapacity(jobject);
};
