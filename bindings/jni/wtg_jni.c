/* JNI shim: Java_net_consensys_wittgenstein_core_NativeNetwork_* -> the C ABI of include/wtg.h, one wrapper per native
 * method of NativeNetwork.java.  Contract violations (negative status) are re-thrown as the unchecked exception the
 * reference would have thrown (IllegalStateException with wtg_last_error() as message).
 *
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude bindings/jni/wtg_jni.c \
 *       -Lwittgenstein_b200 -lwtg_b200 -o libwtg_jni.so
 *
 * Compile-gated: __graft_entry__.build() builds it only where a JDK's jni.h exists (not in this image);
 * tests/test_abi.py checks that it compiles against a minimal stub of jni.h. */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>

#include "wtg.h"

#define NN(name) Java_net_consensys_wittgenstein_core_NativeNetwork_##name

static wtg_net* H(JNIEnv* e, jobject o) {
  jclass c = (*e)->GetObjectClass(e, o);
  jfieldID f = (*e)->GetFieldID(e, c, "handle", "J");
  return (wtg_net*)(intptr_t)(*e)->GetLongField(e, o, f);
}
static int check(JNIEnv* e, int rc) {
  if (rc < 0) (*e)->ThrowNew(e, (*e)->FindClass(e, "java/lang/IllegalStateException"), wtg_last_error());
  return rc;
}
static void withName(JNIEnv* e, jobject o, jstring s, int (*fn)(wtg_net*, const char*)) {
  const char* c = s ? (*e)->GetStringUTFChars(e, s, 0) : 0;
  check(e, fn(H(e, o), c));
  if (c) (*e)->ReleaseStringUTFChars(e, s, c);
}
static void withInts(JNIEnv* e, jobject o, jintArray a, int (*fn)(wtg_net*, const int*)) {
  jint* v = (*e)->GetIntArrayElements(e, a, 0);
  check(e, fn(H(e, o), (const int*)v));
  (*e)->ReleaseIntArrayElements(e, a, v, JNI_ABORT);
}

JNIEXPORT jlong JNICALL NN(create)(JNIEnv* e, jclass c) {
  (void)c;
  wtg_net* n = wtg_create();
  if (!n) (*e)->ThrowNew(e, (*e)->FindClass(e, "java/lang/IllegalStateException"), wtg_last_error());
  return (jlong)(intptr_t)n;
}
JNIEXPORT jlong JNICALL NN(shardCreate)(JNIEnv* e, jclass c, jint rank, jint world, jint device) {
  (void)c;
  wtg_net* n = wtg_shard_create(rank, world, device);
  if (!n) (*e)->ThrowNew(e, (*e)->FindClass(e, "java/lang/IllegalStateException"), wtg_last_error());
  return (jlong)(intptr_t)n;
}
JNIEXPORT void JNICALL NN(destroy)(JNIEnv* e, jclass c, jlong h) {
  (void)e;
  (void)c;
  wtg_destroy((wtg_net*)(intptr_t)h);
}
JNIEXPORT void JNICALL NN(setSeed)(JNIEnv* e, jobject o, jlong s) { check(e, wtg_set_seed(H(e, o), (long long)s)); }
JNIEXPORT void JNICALL NN(setNodeBuilder)(JNIEnv* e, jobject o, jstring s) { withName(e, o, s, wtg_set_node_builder); }
JNIEXPORT void JNICALL NN(setNetworkLatency)(JNIEnv* e, jobject o, jstring s) { withName(e, o, s, wtg_set_network_latency); }
JNIEXPORT void JNICALL NN(setNetworkLatencyMeasured)(JNIEnv* e, jobject o, jintArray p, jintArray v) {
  jsize n = (*e)->GetArrayLength(e, p);
  jint* pp = (*e)->GetIntArrayElements(e, p, 0);
  jint* vv = (*e)->GetIntArrayElements(e, v, 0);
  check(e, wtg_set_network_latency_measured(H(e, o), (const int*)pp, (const int*)vv, (int)n));
  (*e)->ReleaseIntArrayElements(e, p, pp, JNI_ABORT);
  (*e)->ReleaseIntArrayElements(e, v, vv, JNI_ABORT);
}
JNIEXPORT void JNICALL NN(setMsgDiscardTime)(JNIEnv* e, jobject o, jint ms) { check(e, wtg_set_msg_discard_time(H(e, o), ms)); }
JNIEXPORT void JNICALL NN(pingPongInit)(JNIEnv* e, jobject o, jint n) { check(e, wtg_pingpong_init(H(e, o), n)); }
JNIEXPORT void JNICALL NN(gsfInit)(JNIEnv* e, jobject o, jintArray p) { withInts(e, o, p, wtg_gsf_init); }
JNIEXPORT void JNICALL NN(handelInit)(JNIEnv* e, jobject o, jintArray p) { withInts(e, o, p, wtg_handel_init); }
JNIEXPORT void JNICALL NN(sanFerminConstruct)(JNIEnv* e, jobject o, jintArray p) { withInts(e, o, p, wtg_sanfermin_construct); }
JNIEXPORT void JNICALL NN(sanFerminInit)(JNIEnv* e, jobject o) { check(e, wtg_sanfermin_init(H(e, o))); }
JNIEXPORT void JNICALL NN(capposInit)(JNIEnv* e, jobject o, jintArray p) { withInts(e, o, p, wtg_cappos_init); }
JNIEXPORT void JNICALL NN(casperConstruct)(JNIEnv* e, jobject o, jintArray p) { withInts(e, o, p, wtg_casper_construct); }
JNIEXPORT void JNICALL NN(casperInit)(JNIEnv* e, jobject o, jint delay) { check(e, wtg_casper_init(H(e, o), delay)); }
JNIEXPORT jboolean JNICALL NN(runMs)(JNIEnv* e, jobject o, jint ms) { return check(e, wtg_run_ms(H(e, o), ms)) == 1; }
JNIEXPORT jint JNICALL NN(time)(JNIEnv* e, jobject o) { return wtg_time(H(e, o)); }
JNIEXPORT jint JNICALL NN(msgsSize)(JNIEnv* e, jobject o) { return check(e, wtg_msgs_size(H(e, o))); }
JNIEXPORT jint JNICALL NN(msgsSizeAt)(JNIEnv* e, jobject o, jint t) { return check(e, wtg_msgs_size_at(H(e, o), t)); }
JNIEXPORT jintArray JNICALL NN(peekMessages)(JNIEnv* e, jobject o, jint maxRows) {
  wtg_net* n = H(e, o);
  if (maxRows < 0) maxRows = 0;
  int* col = (int*)malloc(sizeof(int) * 6 * (size_t)(maxRows > 0 ? maxRows : 1));
  jintArray out = 0;
  int total = check(e, wtg_peek_messages(n, col, col + maxRows, col + 2 * (size_t)maxRows, col + 3 * (size_t)maxRows,
                                         col + 4 * (size_t)maxRows, col + 5 * (size_t)maxRows, maxRows));
  if (total >= 0) {
    int rows = total < maxRows ? total : maxRows;
    int* flat = (int*)malloc(sizeof(int) * 6 * (size_t)(rows > 0 ? rows : 1));
    for (int i = 0; i < rows; ++i)
      for (int k = 0; k < 6; ++k) flat[6 * i + k] = col[(size_t)k * (size_t)maxRows + (size_t)i];
    out = (*e)->NewIntArray(e, 6 * rows);
    (*e)->SetIntArrayRegion(e, out, 0, 6 * rows, (const jint*)flat);
    free(flat);
  }
  free(col);
  return out;
}
JNIEXPORT void JNICALL NN(stopNode)(JNIEnv* e, jobject o, jint id) { check(e, wtg_stop_node(H(e, o), id)); }
JNIEXPORT void JNICALL NN(startNode)(JNIEnv* e, jobject o, jint id) { check(e, wtg_start_node(H(e, o), id)); }
JNIEXPORT void JNICALL NN(partition)(JNIEnv* e, jobject o, jfloat part) { check(e, wtg_partition(H(e, o), part)); }
JNIEXPORT void JNICALL NN(endPartition)(JNIEnv* e, jobject o) { check(e, wtg_end_partition(H(e, o))); }

static int localCount(wtg_net* n) {
  int first = 0, count = wtg_node_count(n);
  wtg_shard_range(n, &first, &count);
  return count;
}
JNIEXPORT jlongArray JNICALL NN(nodeCounters)(JNIEnv* e, jobject o) {
  wtg_net* n = H(e, o);
  int cnt = 5 * localCount(n);
  long long* buf = (long long*)malloc(sizeof(long long) * (size_t)cnt);
  jlongArray out = 0;
  if (check(e, wtg_node_counters(n, buf)) >= 0) {
    out = (*e)->NewLongArray(e, cnt);
    (*e)->SetLongArrayRegion(e, out, 0, cnt, (const jlong*)buf);
  }
  free(buf);
  return out;
}
JNIEXPORT jlongArray JNICALL NN(gsfVerified)(JNIEnv* e, jobject o) {
  wtg_net* n = H(e, o);
  int words = wtg_node_count(n) / 64;
  if (words < 1) words = 1;
  size_t cnt = (size_t)localCount(n) * (size_t)words;
  unsigned long long* buf = (unsigned long long*)malloc(sizeof(unsigned long long) * cnt);
  jlongArray out = 0;
  if (check(e, wtg_gsf_verified(n, buf)) >= 0) {
    out = (*e)->NewLongArray(e, (jsize)cnt);
    (*e)->SetLongArrayRegion(e, out, 0, (jsize)cnt, (const jlong*)buf);
  }
  free(buf);
  return out;
}
JNIEXPORT jintArray JNICALL NN(pingPongPongs)(JNIEnv* e, jobject o) {
  wtg_net* n = H(e, o);
  int cnt = wtg_node_count(n);
  int* buf = (int*)malloc(sizeof(int) * (size_t)cnt);
  jintArray out = 0;
  if (check(e, wtg_pingpong_pongs(n, buf)) >= 0) {
    out = (*e)->NewIntArray(e, cnt);
    (*e)->SetIntArrayRegion(e, out, 0, cnt, (const jint*)buf);
  }
  free(buf);
  return out;
}
JNIEXPORT jbyteArray JNICALL NN(shardExport)(JNIEnv* e, jobject o) {
  unsigned char h[128];
  if (check(e, wtg_shard_export(H(e, o), h)) < 0) return 0;
  jbyteArray out = (*e)->NewByteArray(e, 128);
  (*e)->SetByteArrayRegion(e, out, 0, 128, (const jbyte*)h);
  return out;
}
JNIEXPORT void JNICALL NN(shardLink)(JNIEnv* e, jobject o, jbyteArray all) {
  jbyte* v = (*e)->GetByteArrayElements(e, all, 0);
  check(e, wtg_shard_link(H(e, o), (const unsigned char*)v));
  (*e)->ReleaseByteArrayElements(e, all, v, JNI_ABORT);
}
JNIEXPORT jintArray JNICALL NN(shardRange)(JNIEnv* e, jobject o) {
  int r[2] = {0, 0};
  if (check(e, wtg_shard_range(H(e, o), &r[0], &r[1])) < 0) return 0;
  jintArray out = (*e)->NewIntArray(e, 2);
  (*e)->SetIntArrayRegion(e, out, 0, 2, (const jint*)r);
  return out;
}
