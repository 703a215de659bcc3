/*
 * klb_Cwrapper.h -- the C ABI of the compress/decompress path. Drop-in for the reference's src/klb_Cwrapper.h:40-64:
 * the six entry points below have the same names, argument order, argument meaning, ownership rules and integer
 * return codes, so the JNI glue (src/jni/org_janelia_simview_lfm_LFMJNI.cpp:19,45,191,375) and the MEX wrappers bind
 * to this library without change. Extensions (header knobs the old ABI cannot express, memory-to-memory and
 * device-resident entry points) are in lfm_b200.h.
 *
 * Return codes (src/klb_Cwrapper.cpp:33-46, src/klb_imageIO.cpp:220-221,531-532,2267-2268,2619-2631):
 *   0 ok; 2 block codec failure / no blocks; 3 cannot open for read / API misuse; 5 cannot create output or unknown codec;
 *   new: 6 CUDA failure (the reference ignores CUDA errors); 7 unsupported (non-16-bit data, ZLIB codec, video stacks with
 *   predictor way angle/space, KLB blocks that need more than 32 bzip2 blocks per stream).
 */
#ifndef __KLB_IMAGE_C_WRAPPER_H__
#define __KLB_IMAGE_C_WRAPPER_H__

#ifdef __cplusplus
extern "C" {
#endif

#include <stdint.h>
#include "common.h"

#define DECLSPECIFIER

/* replaces src/klb_Cwrapper.cpp:19-50. pixelSize, blockSize, metadata may be NULL (defaults 1.0 / {96,96,8,1,1} / zeros).
   Always auto-selects the predictor, image mode, Nnum = 13 (as the reference's wrapper does). numThreads is accepted
   for compatibility; the work runs on the GPU(s). */
DECLSPECIFIER int writeKLBstack(const void* im, const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE dataType, int numThreads, float32_t pixelSize[KLB_DATA_DIMS], uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE compressionType, char metadata[KLB_METADATA_SIZE]);

/* replaces src/klb_Cwrapper.cpp:53-87: one pointer per XY slice; xyzct[3] = xyzct[4] = 1 required (else 3) */
DECLSPECIFIER int writeKLBstackSlices(const void** im, const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE dataType, int numThreads, float32_t pixelSize[KLB_DATA_DIMS], uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE compressionType, char metadata[KLB_METADATA_SIZE]);

/* replaces src/klb_Cwrapper.cpp:90-110 */
DECLSPECIFIER int readKLBheader(const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE *dataType, float32_t pixelSize[KLB_DATA_DIMS], uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE *compressionType, char metadata[KLB_METADATA_SIZE]);

/* replaces src/klb_Cwrapper.cpp:112-152: returns malloc()ed memory the caller free()s, NULL on error;
   every argument after numThreads may be NULL */
DECLSPECIFIER void* readKLBstack(const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE *dataType, int numThreads, float32_t pixelSize[KLB_DATA_DIMS], uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE *compressionType, char metadata[KLB_METADATA_SIZE]);

/* replaces src/klb_Cwrapper.cpp:154-174 */
DECLSPECIFIER int readKLBstackInPlace(const char* filename, void* im, enum KLB_DATA_TYPE *dataType, int numThreads);

/* replaces src/klb_Cwrapper.cpp:176-189; bounds are inclusive; im holds the ROI only (x fastest) */
DECLSPECIFIER int readKLBroiInPlace(const char* filename, void* im, uint32_t xyzctLB[KLB_DATA_DIMS], uint32_t xyzctUB[KLB_DATA_DIMS], int numThreads);

#ifdef __cplusplus
}
#endif
#endif
