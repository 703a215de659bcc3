// klb_imageHeader.cpp -- header object of the .lfm container (behaviour of the reference's src/klb_imageHeader.cpp,
// re-implemented; byte layout in include/klb_imageHeader.h).
#include <algorithm>
#include <cmath>
#include <fstream>
#include <limits>
#include "klb_imageHeader.h"

klb_image_header::klb_image_header() : blockOffset(NULL), Nb(0)
{
	const std::uint32_t zero[KLB_DATA_DIMS] = { 0, 0, 0, 0, 0 };
	setHeader(zero, KLB_DATA_TYPE::UINT16_TYPE);
}

klb_image_header::klb_image_header(const klb_image_header& p) : blockOffset(NULL), Nb(0)
{
	memcpy(optimalBlockSizeInBytes, p.optimalBlockSizeInBytes, sizeof(optimalBlockSizeInBytes));
	setHeader(p.xyzct, p.dataType, p.pixelSize, p.blockSize, p.compressionType, p.metadata, p.headerVersion, p.Nnum);
	resizeBlockOffset(p.Nb);
	if (Nb) memcpy(blockOffset, p.blockOffset, sizeof(std::uint64_t) * Nb);
}

klb_image_header::~klb_image_header()
{
	delete[] blockOffset;
	blockOffset = NULL; Nb = 0;
}

klb_image_header& klb_image_header::operator=(const klb_image_header& p)
{
	if (this == &p) return *this;
	memcpy(optimalBlockSizeInBytes, p.optimalBlockSizeInBytes, sizeof(optimalBlockSizeInBytes));
	// note: the reference forgets Nnum here (src/klb_imageHeader.cpp:30) and silently resets it to 13; we copy it
	setHeader(p.xyzct, p.dataType, p.pixelSize, p.blockSize, p.compressionType, p.metadata, p.headerVersion, p.Nnum);
	resizeBlockOffset(p.Nb);
	if (Nb) memcpy(blockOffset, p.blockOffset, sizeof(std::uint64_t) * Nb);
	return *this;
}

size_t klb_image_header::calculateNumBlocks() const
{
	size_t n = 1;
	for (int d = 0; d < KLB_DATA_DIMS; d++) n *= (size_t)std::ceil((float)xyzct[d] / (float)blockSize[d]);   // float ceil, as the reference
	return n;
}

size_t klb_image_header::getBytesPerPixel() const
{
	static const size_t bpp[10] = { 1, 2, 4, 8, 1, 2, 4, 8, 4, 8 };
	int t = (int)dataType;
	return (t >= 0 && t < 10) ? bpp[t] : 0;
}

std::uint32_t klb_image_header::getBlockSizeBytes() const
{
	std::uint32_t n = 1;
	for (int d = 0; d < KLB_DATA_DIMS; d++) n *= blockSize[d];
	return n * (std::uint32_t)getBytesPerPixel();
}

std::uint64_t klb_image_header::getImageSizePixels() const
{
	std::uint64_t n = 1;
	for (int d = 0; d < KLB_DATA_DIMS; d++) n *= xyzct[d];
	return n;
}
std::uint64_t klb_image_header::getImageSizeBytes() const { return getImageSizePixels() * getBytesPerPixel(); }

void klb_image_header::packFixed(std::uint8_t out[320]) const
{
	out[0] = headerVersion; out[1] = Nnum;
	memcpy(out + 2, xyzct, 20); memcpy(out + 22, pixelSize, 20);
	out[42] = (std::uint8_t)dataType; out[43] = (std::uint8_t)compressionType;
	memcpy(out + 44, metadata, KLB_METADATA_SIZE); memcpy(out + 300, blockSize, 20);
}
void klb_image_header::unpackFixed(const std::uint8_t in[320])
{
	headerVersion = in[0]; Nnum = in[1];
	memcpy(xyzct, in + 2, 20); memcpy(pixelSize, in + 22, 20);
	dataType = (KLB_DATA_TYPE)in[42]; compressionType = (KLB_COMPRESSION_TYPE)in[43];
	memcpy(metadata, in + 44, KLB_METADATA_SIZE); memcpy(blockSize, in + 300, 20);
}

void klb_image_header::writeHeader(FILE* fid)
{
	std::uint8_t fixed[320];
	packFixed(fixed);
	fwrite(fixed, 1, sizeof(fixed), fid);
	if (Nb) fwrite(blockOffset, sizeof(std::uint64_t), Nb, fid);
}

// legacy KLB stream layout (no headerVersion / Nnum / metadata), as the reference's ostream overload writes it
void klb_image_header::writeHeader(std::ostream& fid)
{
	fid.write((const char*)xyzct, sizeof(xyzct));
	fid.write((const char*)pixelSize, sizeof(pixelSize));
	std::uint8_t dt = (std::uint8_t)dataType, ct = (std::uint8_t)compressionType;
	fid.write((const char*)&dt, 1); fid.write((const char*)&ct, 1);
	fid.write((const char*)blockSize, sizeof(blockSize));
	fid.write((const char*)blockOffset, sizeof(std::uint64_t) * Nb);
}

// number of blocks of the current xyzct / blockSize if it does not exceed `limit` (the header of an untrusted file: the
// product of five 32-bit ratios must not overflow or drive an allocation); false otherwise
bool klb_image_header::numBlocksBounded(size_t limit, size_t* nb) const
{
	size_t n = 1;
	for (int d = 0; d < KLB_DATA_DIMS; d++) {
		if (blockSize[d] == 0) return false;
		const size_t r = (size_t)std::ceil((float)xyzct[d] / (float)blockSize[d]);
		if (r != 0 && n > limit / r) return false;
		n *= r;
	}
	if (nb) *nb = n;
	return n <= limit;
}

// The stream must hold the 320 fixed bytes and the whole blockOffset table; a short or inconsistent header leaves Nb == 0
// (callers report "no blocks", code 2) instead of parsing uninitialised bytes or allocating a table the file cannot hold.
void klb_image_header::readHeader(std::istream& fid)
{
	std::uint8_t fixed[320];
	resizeBlockOffset(0);
	const std::istream::pos_type here = fid.tellg();
	std::uint64_t avail = ~0ull;
	if (here != std::istream::pos_type(-1)) {
		fid.seekg(0, std::ios::end);
		const std::istream::pos_type endp = fid.tellg();
		fid.seekg(here);
		if (endp != std::istream::pos_type(-1) && endp >= here) avail = (std::uint64_t)(endp - here);
	}
	fid.read((char*)fixed, sizeof(fixed));
	if ((size_t)fid.gcount() != sizeof(fixed)) return;
	unpackFixed(fixed);
	size_t nb = 0;
	const std::uint64_t room = avail >= sizeof(fixed) ? (avail - sizeof(fixed)) / sizeof(std::uint64_t) : 0;
	if (!numBlocksBounded((size_t)std::min<std::uint64_t>(room, (std::uint64_t)1 << 40), &nb)) return;
	resizeBlockOffset(nb);
	if (Nb) {
		fid.read((char*)blockOffset, sizeof(std::uint64_t) * Nb);
		if ((size_t)fid.gcount() != sizeof(std::uint64_t) * Nb) resizeBlockOffset(0);
	}
}

int klb_image_header::readHeader(const char* filename)
{
	std::ifstream fid(filename, std::ios::binary | std::ios::in);
	if (!fid.is_open()) {
		std::cout << "ERROR: klb_image_header::readHeader : file " << filename << " could not be opened to read header" << std::endl;
		return 2;
	}
	try { readHeader(fid); } catch (const std::bad_alloc&) { resizeBlockOffset(0); }
	if (Nb == 0) {
		std::cout << "ERROR: klb_image_header::readHeader : file " << filename << " holds no complete header / blockOffset table" << std::endl;
		return 2;
	}
	return 0;
}

void klb_image_header::resizeBlockOffset(size_t Nb_)
{
	if (Nb == Nb_) return;
	delete[] blockOffset;
	Nb = Nb_;
	blockOffset = Nb ? new std::uint64_t[Nb]() : NULL;
}

size_t klb_image_header::getBlockCompressedSizeBytes(size_t i) const
{
	if (i >= Nb) return 0;
	return i ? (size_t)(blockOffset[i] - blockOffset[i - 1]) : (size_t)blockOffset[0];
}

std::uint64_t klb_image_header::getBlockOffset(size_t i) const
{
	if (i >= Nb) return std::numeric_limits<std::uint64_t>::max();
	return i ? blockOffset[i - 1] : 0;
}

std::uint64_t klb_image_header::getCompressedFileSizeInBytes() const { return getSizeInBytes() + (Nb ? blockOffset[Nb - 1] : 0); }

void klb_image_header::setDefaultBlockSize()
{
	setOptimalBlockSizeInBytes();
	std::uint32_t bpp = (std::uint32_t)getBytesPerPixel();
	for (int d = 0; d < KLB_DATA_DIMS; d++) blockSize[d] = std::max(bpp ? optimalBlockSizeInBytes[d] / bpp : 1u, 1u);
}

void klb_image_header::setHeader(const std::uint32_t xyzct_[KLB_DATA_DIMS], const KLB_DATA_TYPE dataType_, const float32_t pixelSize_[KLB_DATA_DIMS],
                                 const std::uint32_t blockSize_[KLB_DATA_DIMS], const KLB_COMPRESSION_TYPE compressionType_,
                                 const char metadata_[KLB_METADATA_SIZE], const std::uint8_t headerVersion_, const std::uint8_t Nnum_)
{
	memcpy(xyzct, xyzct_, sizeof(xyzct));
	dataType = dataType_; compressionType = compressionType_;
	headerVersion = headerVersion_; Nnum = Nnum_;
	if (pixelSize_) memcpy(pixelSize, pixelSize_, sizeof(pixelSize));
	else for (int d = 0; d < KLB_DATA_DIMS; d++) pixelSize[d] = 1.0f;
	if (metadata_) memcpy(metadata, metadata_, KLB_METADATA_SIZE); else memset(metadata, 0, KLB_METADATA_SIZE);
	if (blockSize_) memcpy(blockSize, blockSize_, sizeof(blockSize)); else setDefaultBlockSize();
}
