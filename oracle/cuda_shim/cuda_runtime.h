/*
 * TEST INFRASTRUCTURE ONLY (oracle/): a host-only stand-in for the handful of CUDA runtime
 * features the reference uses, so that the reference's own predictor / entropy kernels
 * (/root/reference/src/lfm_Predictors*.cu) and orchestrator (klb_imageIO.cpp) can be compiled
 * with g++ and executed on a CPU-only machine. "Device memory" is host memory, a kernel launch
 * is a nested loop over the grid. The reference kernels use no shared memory, no barriers and no
 * warp intrinsics, so sequential execution of the threads is a faithful execution.
 *
 * Nothing in the product links this. See oracle/build_ref.py.
 */
#ifndef LFM_ORACLE_CUDA_SHIM_H
#define LFM_ORACLE_CUDA_SHIM_H

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <math.h>
#include <mutex>
#include <unordered_set>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline

struct uint3 { unsigned int x, y, z; };
struct dim3 {
	unsigned int x, y, z;
	dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};

inline thread_local uint3 blockIdx, threadIdx;
inline thread_local dim3 blockDim, gridDim;

typedef int cudaError_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };

namespace lfmshim {
inline std::mutex& mtx() { static std::mutex m; return m; }
inline std::unordered_set<void*>& owned() { static std::unordered_set<void*> s; return s; }

template <class F>
inline void launch(dim3 grid, dim3 block, F&& body)
{
	gridDim = grid; blockDim = block;
	for (unsigned bz = 0; bz < grid.z; bz++)
	for (unsigned by = 0; by < grid.y; by++)
	for (unsigned bx = 0; bx < grid.x; bx++)
	{
		blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
		for (unsigned tz = 0; tz < block.z; tz++)
		for (unsigned ty = 0; ty < block.y; ty++)
		for (unsigned tx = 0; tx < block.x; tx++)
		{
			threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
			body();
		}
	}
}
}

#define LFMSHIM_LAUNCH(kernel, grid, block, ...) ::lfmshim::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })

template <class T>
inline cudaError_t cudaMalloc(T** p, size_t bytes)
{
	/* +64: the reference's pair histogram writes bin 65535 of a 65535-entry table
	   (lfm_Predictors.cu:2845 vs klb_imageIO.cpp:2043); keep that inside the allocation. */
	void* q = calloc(1, bytes + 64);
	*p = (T*)q;
	std::lock_guard<std::mutex> g(lfmshim::mtx());
	lfmshim::owned().insert(q);
	return q ? 0 : 2;
}
inline cudaError_t cudaFree(void* p)
{
	/* the reference also cudaFree()s a host stack array (klb_imageIO.cpp:2312): ignore foreign pointers */
	std::lock_guard<std::mutex> g(lfmshim::mtx());
	auto it = lfmshim::owned().find(p);
	if (it == lfmshim::owned().end()) return 1;
	lfmshim::owned().erase(it);
	free(p);
	return 0;
}
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { memmove(dst, src, n); return 0; }
inline cudaError_t cudaMemset(void* dst, int v, size_t n) { memset(dst, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* dst, int v, size_t n) { memset(dst, v, n); return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }

template <class T, class U>
inline T atomicAdd(T* addr, U v) { return __atomic_fetch_add(addr, (T)v, __ATOMIC_RELAXED); }

#endif
