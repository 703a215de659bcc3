// lfm_predict.cu -- forward predictor + symbolize, and the inverse (unsymbolize + wavefront un-predict).
//
// Forward replaces _predictorN_{tiles,angle,space} + symbolizeKernel + the per-frame cudaMemcpy loop of
// klb_imageIO::Predictor_{both,angle,space}[_GPU] (src/klb_imageIO.cpp:1227-1746): one fused launch over all frames.
// Inverse replaces unsymbolizeKernel + the single-threaded HOST loops unPredictorN_* (src/lfm_Predictors.cu:1470-2739,
// src/klb_imageIO.cpp:1748-1821) with a GPU wavefront, one CTA per frame, frames in parallel:
//   * schedule A, w = tx+ty+u+v (~ tilesX+tilesY+2T steps): valid whenever the near neighbours L/U/UL are only used
//     inside a tile -- every way/predictor except predictor 2 of the ways "tiles" and "angle";
//   * schedule B, row by row (H steps): predictor 2 of those two ways reads U across the tile border (first tile row
//     of interior tiles), which makes every image column one long chain. Rows then only depend on earlier rows, plus
//     short left-to-right chains inside the first tile column / the first image row, which one thread walks.
#include <algorithm>
#include <cooperative_groups.h>
#include "lfm_device.cuh"
#include "lfm_predict.cuh"

namespace cg = cooperative_groups;

namespace lfm {

constexpr int PF_NT = 256;

__global__ void __launch_bounds__(PF_NT)
k_predict_fwd(const uint16_t* __restrict__ img, uint16_t* __restrict__ sym, int W, int H, int T, int way, int k,
              int video, uint32_t z0, uint32_t nz)
{
	const uint64_t fpx = (uint64_t)W * H;
	const uint64_t idx = (uint64_t)blockIdx.x * PF_NT + threadIdx.x;
	if (idx >= fpx * nz) return;
	const uint32_t zi = (uint32_t)(idx / fpx);
	const uint32_t rem = (uint32_t)(idx - (uint64_t)zi * fpx);
	const int y = (int)(rem / (uint32_t)W), x = (int)(rem - (uint32_t)y * (uint32_t)W);
	const uint32_t z = z0 + zi;
	const uint16_t* cur = img + (uint64_t)z * fpx;
	const int tx = x / T, ty = y / T, u = x - tx * T, v = y - ty * T;
	auto px = [&](int dx, int dy) { return (int)__ldg(cur + (size_t)(y + dy) * W + (x + dx)); };
	int p = predict0(px, T, way, k, tx, ty, u, v);
	if (video & (int)z & 1) {                       // `i_or_v & z`: only odd frames look back (klb_imageIO.cpp:1243)
		int P = (int)__ldg(cur - fpx + (size_t)y * W + x);
		p = (x == 0 && y == 0) ? P : ((p + P) >> 1);
	}
	sym[(uint64_t)z * fpx + rem] = symbolize16(px(0, 0) - p);
}

// ---------------------------------------------------------------------------------------------------------
// Inverse predictor: ONE cooperative launch over all frames of the call. Every wavefront step handles the ready
// pixels of all frames, spread over the whole grid; steps are separated by a grid-wide barrier. Neighbours written
// by other SMs are read with ld.global.cg (L2), never through the non-coherent L1.
constexpr int UP_NT = 512;

__global__ void __launch_bounds__(UP_NT, 2)
k_unpredict(const uint16_t* __restrict__ sym, uint16_t* out, int W, int H, int T, int way, int k, int video,
            uint32_t z_start, uint32_t z_step, uint32_t nframes)
{
	cg::grid_group grid = cg::this_grid();
	__shared__ uint32_t pre[512];
	__shared__ uint32_t red[64];
	const uint32_t tid = threadIdx.x;
	const uint64_t gtid = (uint64_t)blockIdx.x * UP_NT + tid, gsize = (uint64_t)gridDim.x * UP_NT;
	const uint64_t fpx = (uint64_t)W * H;
	const int tilesX = (W + T - 1) / T, tilesY = (H + T - 1) / T;
	const int nr = 2 * T - 1;                         // pixel anti-diagonals inside a tile (<= 509)
	const int nsteps = tilesX + tilesY - 1 + nr - 1;

	auto decode_px = [&](uint32_t f, int x, int y, int tx, int ty, int u, int v) {
		const uint32_t z = z_start + f * z_step;
		const uint16_t* s = sym + (uint64_t)z * fpx;
		uint16_t* o = out + (uint64_t)z * fpx;
		auto px = [&](int dx, int dy) { return (int)__ldcg(o + (size_t)(y + dy) * W + (x + dx)); };
		int p = predict0(px, T, way, k, tx, ty, u, v);
		if (video & (int)z & 1) {
			int P = (int)__ldcg(o + (size_t)y * W + x - fpx);    // previous (even) frame, reconstructed by an earlier launch
			p = (x == 0 && y == 0) ? P : ((p + P) >> 1);
		}
		o[(size_t)y * W + x] = (uint16_t)(unsymbolize16(__ldg(s + (size_t)y * W + x)) + p);
	};

	if (k == 2 && way != 2) {
		// ---- schedule B: rows. Leading pixels with left-neighbour chains are walked by one thread per frame.
		for (int y = 0; y < H; y++) {
			const int ty = y / T, v = y - ty * T;
			const int seq = (y == 0) ? W : min(T, W);
			if (gtid < nframes) for (int x = 0; x < seq; x++) decode_px((uint32_t)gtid, x, y, x / T, ty, x % T, v);
			const uint64_t par = (uint64_t)(W - seq) * nframes;
			for (uint64_t i = gtid; i < par; i += gsize) {
				uint32_t f = (uint32_t)(i / (uint32_t)(W - seq)); int x = seq + (int)(i - (uint64_t)f * (W - seq));
				decode_px(f, x, y, x / T, ty, x % T, v);
			}
			grid.sync();
		}
		return;
	}

	// ---- schedule A: 4-D wavefront w = tx + ty + u + v
	for (int w = 0; w < nsteps; w++) {
		uint32_t c = 0;
		if ((int)tid < nr) {
			int r = (int)tid, sd = w - r;
			if (sd >= 0 && sd <= tilesX + tilesY - 2) {
				int lo = max(0, sd - (tilesY - 1)), hi = min(sd, tilesX - 1);
				int np = min(r, 2 * T - 2 - r) + 1;
				c = (uint32_t)((hi - lo + 1) * np);
			}
		}
		uint32_t total; uint32_t inc = block_scan_add<UP_NT>(c, red, &total);
		if ((int)tid < nr) pre[tid] = inc - c;
		__syncthreads();
		const uint64_t all = (uint64_t)total * nframes;
		for (uint64_t ii = gtid; ii < all; ii += gsize) {
			const uint32_t f = (uint32_t)(ii / total), i = (uint32_t)(ii - (uint64_t)f * total);
			int lo_r = 0, hi_r = nr - 1;                 // last r with pre[r] <= i
			while (lo_r < hi_r) { int mid = (lo_r + hi_r + 1) >> 1; if (pre[mid] <= i) lo_r = mid; else hi_r = mid - 1; }
			const int r = lo_r, sd = w - r;
			const uint32_t j = i - pre[r];
			const int np = min(r, 2 * T - 2 - r) + 1;
			const int ti = (int)(j / (uint32_t)np), pi = (int)(j - (uint32_t)ti * (uint32_t)np);
			const int tx = max(0, sd - (tilesY - 1)) + ti, ty = sd - tx;
			const int u = max(0, r - (T - 1)) + pi, v = r - u;
			const int x = tx * T + u, y = ty * T + v;
			if (x < W && y < H) decode_px(f, x, y, tx, ty, u, v);
		}
		grid.sync();
	}
}

void launch_predict_fwd(const uint16_t* img, uint16_t* sym, int W, int H, int T, int way, int k, int video,
                        uint32_t z0, uint32_t nz, cudaStream_t st)
{
	uint64_t total = (uint64_t)W * H * nz;
	uint64_t blocks = (total + PF_NT - 1) / PF_NT;
	k_predict_fwd<<<(unsigned)blocks, PF_NT, 0, st>>>(img, sym, W, H, T, way, k, video, z0, nz);
}

// frames z_start, z_start+z_step, ... (count of them); video stacks: even frames first, then odd frames.
// Returns 0, or 1 if the cooperative launch is not possible.
int launch_unpredict(const uint16_t* sym, uint16_t* out, int W, int H, int T, int way, int k, int video,
                     uint32_t z_start, uint32_t z_step, uint32_t count, int sm_count, cudaStream_t st)
{
	if (count == 0) return 0;
	int per_sm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_unpredict, UP_NT, 0);
	if (per_sm < 1) return 1;
	const int max_grid = per_sm * sm_count;
	const int tilesX = (W + T - 1) / T, tilesY = (H + T - 1) / T;
	const uint64_t steps = (k == 2 && way != 2) ? (uint64_t)H : (uint64_t)(tilesX + tilesY + 2 * T);
	uint64_t per_step = ((uint64_t)W * H * count + steps - 1) / steps;            // mean ready pixels per step
	uint64_t want = (per_step + UP_NT - 1) / UP_NT;                                // about one pixel per thread per step
	int grid = (int)std::min<uint64_t>((uint64_t)max_grid, std::max<uint64_t>(1, want));
	void* args[] = { (void*)&sym, (void*)&out, (void*)&W, (void*)&H, (void*)&T, (void*)&way, (void*)&k, (void*)&video,
	                 (void*)&z_start, (void*)&z_step, (void*)&count };
	return cudaLaunchCooperativeKernel((void*)k_unpredict, dim3(grid), dim3(UP_NT), args, 0, st) == cudaSuccess ? 0 : 1;
}

}  // namespace lfm
