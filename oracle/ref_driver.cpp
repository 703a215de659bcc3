/*
 * TEST INFRASTRUCTURE ONLY (oracle/): thin extern "C" driver linked against the UNMODIFIED reference
 * objects (built from /root/reference/src by oracle/build_ref.py into oracle/_ref/). It exposes the
 * knobs the reference's own C ABI cannot express (headerVersion, Nnum; klb_Cwrapper.cpp:19-50 always
 * uses the defaults) and direct access to the per-frame predictor launchers, the host inverse
 * predictors and the 2-D entropy estimate, so tests can pin the oracle restatement against the real thing.
 *
 * The predictor "way" is compile-time in the reference (src/common.h:19); one library per way is built.
 */
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "klb_imageIO.h"
#include "lfm_Predictors.h"
#include "lfm_Predictors_space.h"
#include "lfm_Predictors_angle.h"

typedef void (*fwd_fn)(const uint16_t*, int16_t*, int, int, int, int, int);
typedef void (*inv_fn)(const int16_t*, uint16_t*, int, int, int, int, int);

static fwd_fn FWD[3][7] = {
	{ predictor1_tiles_GPU, predictor2_tiles_GPU, predictor3_tiles_GPU, predictor4_tiles_GPU, predictor5_tiles_GPU, predictor6_tiles_GPU, predictor7_tiles_GPU },
	{ predictor1_angle_GPU, predictor2_angle_GPU, predictor3_angle_GPU, predictor4_angle_GPU, predictor5_angle_GPU, predictor6_angle_GPU, predictor7_angle_GPU },
	{ predictor1_space_GPU, predictor2_space_GPU, predictor3_space_GPU, predictor4_space_GPU, predictor5_space_GPU, predictor6_space_GPU, predictor7_space_GPU } };
static inv_fn INV[3][7] = {
	{ unPredictor1_tiles, unPredictor2_tiles, unPredictor3_tiles, unPredictor4_tiles, unPredictor5_tiles, unPredictor6_tiles, unPredictor7_tiles },
	{ unPredictor1_angle, unPredictor2_angle, unPredictor3_angle, unPredictor4_angle, unPredictor5_angle, unPredictor6_angle, unPredictor7_angle },
	{ unPredictor1_space, unPredictor2_space, unPredictor3_space, unPredictor4_space, unPredictor5_space, unPredictor6_space, unPredictor7_space } };

extern "C" {

int ref_way() { return LFM_PREDICTOR_WAY; }

/* writeImage with every header knob exposed (mirrors test/mainTest_lfmIO.cxx:60-97) */
int ref_write(const void* img, const char* filename, const uint32_t xyzct[5], const uint32_t* blockSize,
              int headerVersion, int Nnum, int numThreads, int* storedHeaderVersion)
{
	klb_imageIO io{ std::string(filename) };
	io.header.setHeader(xyzct, KLB_DATA_TYPE::UINT16_TYPE, NULL, blockSize, KLB_COMPRESSION_TYPE::BZIP2, NULL,
	                    (uint8_t)headerVersion, (uint8_t)Nnum);
	int err = io.writeImage((const char*)img, numThreads);
	if (storedHeaderVersion) *storedHeaderVersion = io.header.headerVersion;
	return err;
}

int ref_read_full(const char* filename, void* out, int numThreads)
{
	klb_imageIO io{ std::string(filename) };
	int err = io.readHeader();
	if (err) return err;
	return io.readImageFull((char*)out, numThreads);
}

int ref_read_roi(const char* filename, void* out, const uint32_t lb[5], const uint32_t ub[5], int numThreads)
{
	klb_imageIO io{ std::string(filename) };
	int err = io.readHeader();
	if (err) return err;
	klb_ROI roi;
	for (int d = 0; d < 5; d++) { roi.xyzctLB[d] = lb[d]; roi.xyzctUB[d] = ub[d]; }
	return io.readImage((char*)out, &roi, numThreads);
}

/* one frame through the reference's forward kernel + symbolizeKernel.
   cur_prev: host uint16[2*W*H] (current frame followed by previous frame, as Predictor_both lays them out,
   klb_imageIO.cpp:1250-1257); sym: host uint16[W*H]. way 0 tiles / 1 angle / 2 space; pred 1..7. */
int ref_predict_frame(const uint16_t* cur_prev, uint16_t* sym, int W, int H, int T, int way, int pred, int zflag)
{
	if (way < 0 || way > 2 || pred < 1 || pred > 7) return 1;
	size_t n = (size_t)W * H;
	uint16_t* dIn = nullptr; int16_t* dRes = nullptr; uint16_t* dSym = nullptr;
	cudaMalloc(&dIn, 2 * n * sizeof(uint16_t));
	cudaMalloc(&dRes, n * sizeof(int16_t));
	cudaMalloc(&dSym, n * sizeof(uint16_t));
	cudaMemcpy(dIn, cur_prev, 2 * n * sizeof(uint16_t), cudaMemcpyHostToDevice);
	cudaMemcpy(dRes, dIn, n * sizeof(uint16_t), cudaMemcpyDeviceToDevice);
	FWD[way][pred - 1](dIn, dRes, W * (int)sizeof(uint16_t), W, H, zflag, T);
	symbolize_GPU(dSym, dRes, W, H, 1, 0, 0);
	cudaMemcpy(sym, dSym, n * sizeof(uint16_t), cudaMemcpyDeviceToHost);
	cudaFree(dIn); cudaFree(dRes); cudaFree(dSym);
	return 0;
}

/* host inverse of the reference. res: int16[W*H] residuals of this frame; out: points at THIS frame inside a
   buffer whose preceding W*H entries hold the decoded previous frame (the reference reads out - W*H). */
int ref_unpredict_frame(const int16_t* res, uint16_t* out, int W, int H, int T, int way, int pred, int zflag)
{
	if (way < 0 || way > 2 || pred < 1 || pred > 7) return 1;
	INV[way][pred - 1](res, out, W * (int)sizeof(int16_t), W, H, zflag, T);
	return 0;
}

/* klb_imageIO::bwt_entropy_2D (klb_imageIO.cpp:2030-2093) on a symbol image of npx pixels */
float ref_entropy_2d(const uint16_t* sym, uint64_t npx, int id)
{
	klb_imageIO io;
	uint32_t xyzct[5] = { (uint32_t)npx, 1, 1, 1, 1 };
	io.header.setHeader(xyzct, KLB_DATA_TYPE::UINT16_TYPE);
	uint16_t* d = nullptr;
	cudaMalloc(&d, npx * sizeof(uint16_t));
	cudaMemcpy(d, sym, npx * sizeof(uint16_t), cudaMemcpyHostToDevice);
	float e = 0.f;
	io.bwt_entropy_2D(d, &e, id);
	cudaFree(d);
	return e;
}

}
