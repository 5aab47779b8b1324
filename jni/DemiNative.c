/*
 * JNI shim: one native method per C-ABI entry point of include/demi_b200.h.
 * Not compiled in this image (no JDK / jni.h); build on a JVM host with
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *       DemiNative.c -L../demi_b200 -ldemi_b200 -o libdemijni.so
 * All buffers are direct java.nio.ByteBuffers (little-endian), so no copies
 * are made on the Java side; handles travel as jlong.
 * Scala side: jni/DemiNative.scala.
 */
#include <jni.h>
#include <stdint.h>
#include "demi_b200.h"

#define H(h) ((demi_handle*)(intptr_t)(h))
#define BUF(env, b) ((*(env))->GetDirectBufferAddress((env), (b)))

JNIEXPORT jlong JNICALL Java_akka_dispatch_verification_DemiNative_create(JNIEnv* env, jclass c, jobject cfg) {
  demi_handle* h = 0;
  int32_t rc = demi_create((const demi_config*)BUF(env, cfg), &h);
  return rc == DEMI_OK ? (jlong)(intptr_t)h : (jlong)rc;          /* negative = error code */
}
JNIEXPORT void JNICALL Java_akka_dispatch_verification_DemiNative_destroy(JNIEnv* env, jclass c, jlong h) {
  demi_destroy(H(h));
}
JNIEXPORT jstring JNICALL Java_akka_dispatch_verification_DemiNative_lastError(JNIEnv* env, jclass c, jlong h) {
  return (*env)->NewStringUTF(env, demi_last_error(H(h)));
}
JNIEXPORT jint JNICALL Java_akka_dispatch_verification_DemiNative_setExternals(JNIEnv* env, jclass c, jlong h, jobject ev, jint n) {
  return demi_set_externals(H(h), (const demi_ext_event*)BUF(env, ev), (uint32_t)n);
}
JNIEXPORT jint JNICALL Java_akka_dispatch_verification_DemiNative_fuzzBatch(JNIEnv* env, jclass c, jlong h, jobject params, jobject out) {
  return demi_fuzz_batch(H(h), (const demi_fuzz_params*)BUF(env, params), (demi_fuzz_result*)BUF(env, out));
}
JNIEXPORT jint JNICALL Java_akka_dispatch_verification_DemiNative_fuzzTrace(JNIEnv* env, jclass c, jlong h, jobject params, jlong seed,
    jobject events, jint capEvents, jobject depParent, jint capNodes, jobject counts /* 2 x u32 */, jobject result) {
  uint32_t* cnt = (uint32_t*)BUF(env, counts);
  return demi_fuzz_trace(H(h), (const demi_fuzz_params*)BUF(env, params), (int64_t)seed,
                         (demi_event*)BUF(env, events), (uint32_t)capEvents, &cnt[0],
                         (uint16_t*)BUF(env, depParent), (uint32_t)capNodes, &cnt[1],
                         (demi_fuzz_result*)BUF(env, result));
}
JNIEXPORT jint JNICALL Java_akka_dispatch_verification_DemiNative_setTrace(JNIEnv* env, jclass c, jlong h, jobject events, jint nEvents,
    jobject externals, jint nExternals) {
  return demi_set_trace(H(h), (const demi_event*)BUF(env, events), (uint32_t)nEvents,
                        (const demi_ext_event*)BUF(env, externals), (uint32_t)nExternals);
}
JNIEXPORT jint JNICALL Java_akka_dispatch_verification_DemiNative_replayBatch(JNIEnv* env, jclass c, jlong h, jobject masks, jint nMasks,
    jint maskWords, jint lookingFor, jint flags, jobject out) {
  return demi_replay_batch(H(h), (const uint64_t*)BUF(env, masks), (uint32_t)nMasks, (uint32_t)maskWords,
                           (uint32_t)lookingFor, (uint32_t)flags, (demi_replay_result*)BUF(env, out));
}
JNIEXPORT jint JNICALL Java_akka_dispatch_verification_DemiNative_ddmin(JNIEnv* env, jclass c, jlong h, jint lookingFor, jint flags,
    jint checkUnmodified, jobject mcsMask, jint maskWords, jobject iterationSizes, jint capIterations, jobject out) {
  return demi_ddmin(H(h), (uint32_t)lookingFor, (uint32_t)flags, checkUnmodified, (uint64_t*)BUF(env, mcsMask),
                    (uint32_t)maskWords, (uint32_t*)BUF(env, iterationSizes), (uint32_t)capIterations,
                    (demi_ddmin_out*)BUF(env, out));
}
JNIEXPORT jint JNICALL Java_akka_dispatch_verification_DemiNative_stats(JNIEnv* env, jclass c, jlong h, jobject out) {
  return demi_stats(H(h), (demi_perf*)BUF(env, out));
}
