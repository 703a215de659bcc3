/*
 * klb_ROI.h -- inclusive region of interest, source-compatible with the reference's src/klb_ROI.h:24-42.
 */
#ifndef __KLB_ROI_H__
#define __KLB_ROI_H__

#include <cstdint>
#include "klb_imageHeader.h"

class klb_ROI
{
public:
	std::uint32_t xyzctLB[KLB_DATA_DIMS];   // lower bound, included
	std::uint32_t xyzctUB[KLB_DATA_DIMS];   // upper bound, included

	void defineSlice(int val, int dim, const std::uint32_t xyzct[KLB_DATA_DIMS]);
	void defineFullImage(const std::uint32_t xyzct[KLB_DATA_DIMS]);
	std::uint32_t getSizePixels(int dim) const { return xyzctUB[dim] - xyzctLB[dim] + 1; }
	std::uint64_t getSizePixels() const
	{
		std::uint64_t n = 1;
		for (int d = 0; d < KLB_DATA_DIMS; d++) n *= (std::uint64_t)(xyzctUB[d] - xyzctLB[d] + 1);
		return n;
	}
};

#endif
