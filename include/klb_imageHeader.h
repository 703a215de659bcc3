/*
 * klb_imageHeader.h -- .lfm / KLB header object, source-compatible with the reference's src/klb_imageHeader.h:32-94
 * (same public fields and methods; callers such as matlabWrapper/writeLFMstack.cpp:59-63,424-443 mutate the fields
 * directly). On-disk layout, little endian, no padding (src/klb_imageHeader.cpp:164-176):
 *   0 headerVersion(1) | 1 Nnum(1) | 2 xyzct(5 x u32) | 22 pixelSize(5 x f32) | 42 dataType(1) | 43 compressionType(1)
 *   | 44 metadata(256) | 300 blockSize(5 x u32) | 320 blockOffset(Nb x u64, END offset of each block) | payload
 */
#ifndef __KLB_IMAGE_HEADER_H__
#define __KLB_IMAGE_HEADER_H__

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include "common.h"

class klb_image_header
{
public:
	std::uint8_t         headerVersion;   // bit 7: video stack; bits 0-6: predictor (stored) / request (on input to writeImage)
	std::uint8_t         Nnum;            // microlens pitch in pixels
	std::uint32_t        xyzct[KLB_DATA_DIMS];
	float32_t            pixelSize[KLB_DATA_DIMS];
	KLB_DATA_TYPE        dataType;
	KLB_COMPRESSION_TYPE compressionType;
	char                 metadata[KLB_METADATA_SIZE];
	std::uint32_t        blockSize[KLB_DATA_DIMS];
	std::uint64_t*       blockOffset;     // Nb entries, inclusive prefix sum of the compressed block sizes
	size_t               Nb;

	klb_image_header();
	klb_image_header(const klb_image_header& p);
	~klb_image_header();
	klb_image_header& operator=(const klb_image_header& p);

	void writeHeader(std::ostream& fid);      // legacy (pre-LFM) layout, kept for source compatibility
	void writeHeader(FILE* fid);
	void readHeader(std::istream& fid);
	int  readHeader(const char* filename);

	size_t getNumBlocks() const { return Nb; }
	int    getMetadataSizeInBytes() const { return KLB_METADATA_SIZE; }
	size_t calculateNumBlocks() const;
	size_t getSizeInBytes() const { return getSizeInBytesFixPortion() + Nb * sizeof(std::uint64_t); }
	size_t getSizeInBytesFixPortion() const { return KLB_DATA_DIMS * (2 * sizeof(std::uint32_t) + sizeof(float32_t)) + 3 * sizeof(std::uint8_t) + sizeof(char) * (KLB_METADATA_SIZE + 1); }
	size_t getBytesPerPixel() const;
	std::uint32_t getBlockSizeBytes() const;
	std::uint64_t getImageSizeBytes() const;
	std::uint64_t getImageSizePixels() const;
	size_t getBlockCompressedSizeBytes(size_t blockId) const;
	std::uint64_t getBlockOffset(size_t blockIdx) const;     // START offset of the block inside the payload
	std::uint64_t getCompressedFileSizeInBytes() const;
	void setDefaultBlockSize();
	void resizeBlockOffset(size_t Nb_);
	void setOptimalBlockSizeInBytes() { optimalBlockSizeInBytes[0] = 192; optimalBlockSizeInBytes[1] = 192; optimalBlockSizeInBytes[2] = 16; optimalBlockSizeInBytes[3] = 1; optimalBlockSizeInBytes[4] = 1; }

	char* getMetadataPtr() { return metadata; }
	char* cloneMetadata() const { char* p = new char[KLB_METADATA_SIZE]; memcpy(p, metadata, KLB_METADATA_SIZE); return p; }
	void  setMetadata(char meta[KLB_METADATA_SIZE]) { memcpy(metadata, meta, KLB_METADATA_SIZE); }

	void setHeader(const std::uint32_t xyzct_[KLB_DATA_DIMS], const KLB_DATA_TYPE dataType_, const float32_t pixelSize_[KLB_DATA_DIMS] = NULL,
	               const std::uint32_t blockSize_[KLB_DATA_DIMS] = NULL, const KLB_COMPRESSION_TYPE compressionType_ = KLB_COMPRESSION_TYPE::BZIP2,
	               const char metadata_[KLB_METADATA_SIZE] = NULL, const std::uint8_t headerVersion_ = KLB_DEFAULT_HEADER_VERSION,
	               const std::uint8_t Nnum_ = 13);

	// serialise / parse the 320-byte fixed part (helpers of this implementation)
	bool numBlocksBounded(size_t limit, size_t* nb) const;    // calculateNumBlocks() with an overflow / size guard
	void packFixed(std::uint8_t out[320]) const;
	void unpackFixed(const std::uint8_t in[320]);

private:
	std::uint32_t optimalBlockSizeInBytes[KLB_DATA_DIMS];
};

#endif
