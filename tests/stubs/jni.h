/* test stub (tests/test_callers_compile.py): the part of the Java Native Interface the reference's glue uses, for a syntax-only
   compile of /root/reference/src/jni/org_janelia_simview_lfm_LFMJNI.cpp against this repository's include/ */
#ifndef STUB_JNI_H
#define STUB_JNI_H
#include <stdint.h>
#define JNIEXPORT
#define JNICALL
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_OK 0
#define JNI_COMMIT 1
#define JNI_ABORT 2
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef uint16_t jchar;
typedef int16_t jshort; typedef float jfloat; typedef double jdouble; typedef jint jsize;
class _jobject {}; class _jclass : public _jobject {}; class _jstring : public _jobject {}; class _jarray : public _jobject {};
class _jbooleanArray : public _jarray {}; class _jbyteArray : public _jarray {}; class _jcharArray : public _jarray {};
class _jshortArray : public _jarray {}; class _jintArray : public _jarray {}; class _jlongArray : public _jarray {};
class _jfloatArray : public _jarray {}; class _jdoubleArray : public _jarray {}; class _jobjectArray : public _jarray {};
typedef _jobject* jobject; typedef _jclass* jclass; typedef _jstring* jstring; typedef _jarray* jarray;
typedef _jbooleanArray* jbooleanArray; typedef _jbyteArray* jbyteArray; typedef _jcharArray* jcharArray; typedef _jshortArray* jshortArray;
typedef _jintArray* jintArray; typedef _jlongArray* jlongArray; typedef _jfloatArray* jfloatArray; typedef _jdoubleArray* jdoubleArray;
typedef _jobjectArray* jobjectArray;
struct JNIEnv_ {
	const char* GetStringUTFChars(jstring, jboolean*);
	void ReleaseStringUTFChars(jstring, const char*);
	jsize GetArrayLength(jarray);
	void* GetDirectBufferAddress(jobject);
	jlong GetDirectBufferCapacity(jobject);
#define STUB_JNI_ARRAY(T, N) T* Get##N##ArrayElements(T##Array, jboolean*); void Release##N##ArrayElements(T##Array, T*, jint); \
	void Get##N##ArrayRegion(T##Array, jsize, jsize, T*); void Set##N##ArrayRegion(T##Array, jsize, jsize, const T*); T##Array New##N##Array(jsize);
	STUB_JNI_ARRAY(jbyte, Byte) STUB_JNI_ARRAY(jshort, Short) STUB_JNI_ARRAY(jint, Int) STUB_JNI_ARRAY(jlong, Long)
	STUB_JNI_ARRAY(jfloat, Float) STUB_JNI_ARRAY(jdouble, Double) STUB_JNI_ARRAY(jboolean, Boolean) STUB_JNI_ARRAY(jchar, Char)
#undef STUB_JNI_ARRAY
	jstring NewStringUTF(const char*);
	jclass FindClass(const char*);
	jint ThrowNew(jclass, const char*);
};
typedef JNIEnv_ JNIEnv;
#endif
