/*
 * common.h -- shared enums and constants of the KLB / LFM container, API-compatible with the reference's
 * src/common.h:9-64 (same names, same values) so existing callers compile unchanged.
 * One deliberate difference: LFM_PREDICTOR_WAY is no longer a compile-time macro that changes the library's
 * behaviour (src/common.h:19); the "way" is a run-time setting, see lfm_b200.h (lfmSetPredictorWay).
 */
#ifndef __KLB_IMAGE_COMMON_H__
#define __KLB_IMAGE_COMMON_H__

typedef float  float32_t;
typedef double float64_t;

#define KLB_DATA_DIMS (5)                /* x, y, z, c, t */
#define KLB_METADATA_SIZE (256)
#define KLB_DEFAULT_HEADER_VERSION (0)   /* 0..7: auto-select the predictor; 8+k: force predictor k; bit 7: video */
#define NUM_PREDICTORS (8)
#define LFM_PREDICTOR_WAY_DEFAULT (0)

enum KLB_DATA_TYPE
{
	UINT8_TYPE = 0, UINT16_TYPE = 1, UINT32_TYPE = 2, UINT64_TYPE = 3,
	INT8_TYPE = 4, INT16_TYPE = 5, INT32_TYPE = 6, INT64_TYPE = 7,
	FLOAT32_TYPE = 8, FLOAT64_TYPE = 9
};

enum KLB_COMPRESSION_TYPE
{
	NONE = 0,
	BZIP2 = 1,
	ZLIB = 2
};

enum LFM_PREDICTORS
{
	ANGLE_AND_SPACE = 0,
	ANGLE = 1,
	SPACE = 2
};

enum LFM_PREDICTORS_TYPE
{
	NO_PREIDICTORS = 0,
	PREIDCTORS_A = 1,
	PREIDCTORS_B = 2,
	PREIDCTORS_C = 3,
	PREIDCTORS_APB_DC = 4,
	PREIDCTORS_A_BDC_Div2 = 5,
	PREIDCTORS_B_ADC_Div2 = 6,
	PREIDCTORS_APB_Div2 = 7,
	PREIDCTORS_APB_Div2_Exten = 8
};

#endif
