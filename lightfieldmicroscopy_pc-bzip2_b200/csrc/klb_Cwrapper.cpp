// klb_Cwrapper.cpp -- the six C entry points of the reference's C ABI (src/klb_Cwrapper.cpp:19-189), same contracts,
// implemented over this library's klb_imageIO (GPU engine). Messages to stdout as the reference does.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "klb_Cwrapper.h"
#include "klb_imageIO.h"

static void report_write_error(int error, bool slices)
{
	switch (error) {
	case 0: return;
	case 2: printf("Error during BZIP compression of one of the blocks"); break;
	case 3: if (slices) { printf("Error: number of channels or number of time points must be 1 for this API call\n"); break; }
	        printf("Error writing the image"); break;
	case 5: printf("Error generating the output file in the specified location"); break;
	case 6: printf("Error: CUDA failure while writing the image"); break;
	case 7: printf("Error: data type / compression type / block size not supported by the GPU engine"); break;
	default: printf("Error writing the image");
	}
}

extern "C" {

int writeKLBstack(const void* im, const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE dataType, int numThreads,
                  float32_t pixelSize[KLB_DATA_DIMS], uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE compressionType,
                  char metadata[KLB_METADATA_SIZE])
{
	klb_imageIO imgIO{ std::string(filename ? filename : "") };
	imgIO.header.setHeader(xyzct, dataType, pixelSize, blockSize, compressionType, metadata);   // headerVersion 0 (auto), Nnum 13
	int error = imgIO.writeImage((const char*)im, numThreads);
	report_write_error(error, false);
	return error;
}

int writeKLBstackSlices(const void** im, const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE dataType, int numThreads,
                        float32_t pixelSize[KLB_DATA_DIMS], uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE compressionType,
                        char metadata[KLB_METADATA_SIZE])
{
	klb_imageIO imgIO{ std::string(filename ? filename : "") };
	imgIO.header.setHeader(xyzct, dataType, pixelSize, blockSize, compressionType, metadata);
	int error = imgIO.writeImageStackSlices((const char**)im, numThreads);
	report_write_error(error, true);
	return error;
}

int readKLBheader(const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE* dataType, float32_t pixelSize[KLB_DATA_DIMS],
                  uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE* compressionType, char metadata[KLB_METADATA_SIZE])
{
	klb_image_header header;
	int error = header.readHeader(filename);
	if (error != 0) return error;
	memcpy(xyzct, header.xyzct, sizeof(uint32_t) * KLB_DATA_DIMS);
	*dataType = header.dataType;
	*compressionType = header.compressionType;
	memcpy(pixelSize, header.pixelSize, sizeof(float32_t) * KLB_DATA_DIMS);
	memcpy(metadata, header.metadata, KLB_METADATA_SIZE);
	memcpy(blockSize, header.blockSize, sizeof(uint32_t) * KLB_DATA_DIMS);
	return 0;
}

void* readKLBstack(const char* filename, uint32_t xyzct[KLB_DATA_DIMS], enum KLB_DATA_TYPE* dataType, int numThreads,
                   float32_t pixelSize[KLB_DATA_DIMS], uint32_t blockSize[KLB_DATA_DIMS], enum KLB_COMPRESSION_TYPE* compressionType,
                   char metadata[KLB_METADATA_SIZE])
{
	klb_imageIO imgFull{ std::string(filename ? filename : "") };
	if (imgFull.readHeader() > 0) return NULL;
	void* im = malloc((size_t)imgFull.header.getImageSizeBytes());
	if (!im) return NULL;
	if (imgFull.readImageFull((char*)im, numThreads) > 0) { free(im); return NULL; }
	memcpy(xyzct, imgFull.header.xyzct, sizeof(uint32_t) * KLB_DATA_DIMS);
	*dataType = imgFull.header.dataType;
	if (compressionType) *compressionType = imgFull.header.compressionType;
	if (pixelSize) memcpy(pixelSize, imgFull.header.pixelSize, sizeof(float32_t) * KLB_DATA_DIMS);
	if (metadata) memcpy(metadata, imgFull.header.metadata, KLB_METADATA_SIZE);
	if (blockSize) memcpy(blockSize, imgFull.header.blockSize, sizeof(uint32_t) * KLB_DATA_DIMS);
	return im;
}

int readKLBstackInPlace(const char* filename, void* im, enum KLB_DATA_TYPE* dataType, int numThreads)
{
	klb_imageIO imgFull{ std::string(filename ? filename : "") };
	int err = imgFull.readHeader();
	if (err > 0) return err;
	*dataType = imgFull.header.dataType;
	return imgFull.readImageFull((char*)im, numThreads);
}

int readKLBroiInPlace(const char* filename, void* im, uint32_t xyzctLB[KLB_DATA_DIMS], uint32_t xyzctUB[KLB_DATA_DIMS], int numThreads)
{
	klb_imageIO img{ std::string(filename ? filename : "") };
	klb_ROI roi;
	for (int d = 0; d < KLB_DATA_DIMS; d++) { roi.xyzctLB[d] = xyzctLB[d]; roi.xyzctUB[d] = xyzctUB[d]; }
	return img.readImage((char*)im, &roi, numThreads);
}

}
