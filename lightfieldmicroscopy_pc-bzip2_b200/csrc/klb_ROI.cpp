// klb_ROI.cpp -- see include/klb_ROI.h (behaviour of the reference's src/klb_ROI.cpp:21-40)
#include <cassert>
#include "klb_ROI.h"

void klb_ROI::defineFullImage(const std::uint32_t xyzct[KLB_DATA_DIMS])
{
	for (int d = 0; d < KLB_DATA_DIMS; d++) { xyzctLB[d] = 0; xyzctUB[d] = xyzct[d] - 1; }
}

void klb_ROI::defineSlice(int val, int dim, const std::uint32_t xyzct[KLB_DATA_DIMS])
{
	assert(dim >= 0 && dim < KLB_DATA_DIMS);
	defineFullImage(xyzct);
	xyzctLB[dim] = xyzctUB[dim] = (std::uint32_t)val;
}
