/* TEST INFRASTRUCTURE: a minimal stand-in for the JDK's <jni.h> — only the types and JNIEnv members that
 * bindings/jni/wtg_jni.c uses — so that the shim can be syntax- and type-checked on a machine without a JDK.
 * Never used to build a loadable library. */
#ifndef WTG_TEST_JNI_STUB_H
#define WTG_TEST_JNI_STUB_H
#include <stdint.h>
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef float jfloat;
typedef jint jsize;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jbyteArray;
typedef void* jfieldID;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*GetObjectClass)(JNIEnv*, jobject);
  jfieldID (*GetFieldID)(JNIEnv*, jclass, const char*, const char*);
  jlong (*GetLongField)(JNIEnv*, jobject, jfieldID);
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
  void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
  jbyte* (*GetByteArrayElements)(JNIEnv*, jbyteArray, jboolean*);
  void (*ReleaseByteArrayElements)(JNIEnv*, jbyteArray, jbyte*, jint);
  jlongArray (*NewLongArray)(JNIEnv*, jsize);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
  jintArray (*NewIntArray)(JNIEnv*, jsize);
  void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
  jbyteArray (*NewByteArray)(JNIEnv*, jsize);
  void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
};
#endif
