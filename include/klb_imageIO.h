/*
 * klb_imageIO.h -- the I/O object of the compress/decompress path, source-compatible with the public part of the
 * reference's src/klb_imageIO.h:47-101: public `header` and `numThreads`, the same constructors and
 * writeImage / writeImageStackSlices / readImage / readImageFull / readHeader signatures and return codes.
 * Behind it the work is done by the B200 engine (one engine per GPU, KLB blocks sharded over the GPUs with no
 * collective on the data path; only a host prefix sum of the block sizes, cf. blockWriter src/klb_imageIO.cpp:1145-1225).
 */
#ifndef __KLB_IMAGE_IO_H__
#define __KLB_IMAGE_IO_H__

#include <string>
#include "klb_imageHeader.h"
#include "klb_ROI.h"

class klb_imageIO
{
public:
	klb_image_header header;
	int numThreads;      // kept for API compatibility (host staging threads); the codec runs on the GPU

	klb_imageIO();
	klb_imageIO(const std::string& filename_);

	std::string getFilename() const { return filename; }
	void setFilename(const std::string& filename_) { filename = filename_; }

	int readHeader() { return header.readHeader(filename.c_str()); }
	int readHeader(const std::string& filename_) { filename = filename_; return readHeader(); }

	// header must be set before; header.headerVersion 0..7 = auto-select, 8+k = force predictor k, |0x80 = video.
	// On return header.headerVersion holds the value stored in the file (src/klb_imageIO.cpp:2272-2398).
	int writeImage(const char* img, int numThreads);
	int writeImageStackSlices(const char** img, int numThreads);
	int readImage(char* img, const klb_ROI* ROI, int numThreads);
	int readImageFull(char* imgOut, int numThreads);

	// ---- extensions of this implementation
	// build / parse a complete file image in host memory (what writeImage writes / readImageFull reads)
	int writeImageToMemory(const char* img, std::string& fileBytes);
	int readImageFromMemory(const char* fileBytes, size_t fileSize, char* imgOut, const klb_ROI* ROI);

private:
	std::string filename;
};

#endif
