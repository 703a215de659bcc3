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
// Inverse predictor: one thread-block CLUSTER per frame (up to 8 CTAs, hardware cluster barrier between wavefront
// steps -- an order of magnitude cheaper than a grid-wide barrier), frames in parallel across clusters.
// Neighbours written by the other CTAs of the cluster are read with ld.global.cg (L2), never through the
// non-coherent L1; barrier.cluster arrive.release / wait.acquire orders the global writes.
constexpr int UP_NT = 512;

__global__ void __launch_bounds__(UP_NT, 2)
k_unpredict(const uint16_t* __restrict__ sym, uint16_t* out, int W, int H, int T, int way, int k, int video,
            uint32_t z_start, uint32_t z_step, uint32_t nframes)
{
	cg::cluster_group cluster = cg::this_cluster();
	__shared__ uint32_t pre[512];
	__shared__ uint32_t red[64];
	const uint32_t tid = threadIdx.x;
	const uint32_t csz = cluster.num_blocks();
	const uint32_t frame = blockIdx.x / csz;
	const uint32_t gtid = cluster.block_rank() * UP_NT + tid, gsize = csz * UP_NT;
	if (frame >= nframes) return;                            // whole clusters only: grid = nframes * csz
	const uint64_t fpx = (uint64_t)W * H;
	const int tilesX = (W + T - 1) / T, tilesY = (H + T - 1) / T;
	const int nr = 2 * T - 1;                                // pixel anti-diagonals inside a tile (<= 509)
	const int nsteps = tilesX + tilesY - 1 + nr - 1;
	const uint32_t z = z_start + frame * z_step;
	const uint16_t* s = sym + (uint64_t)z * fpx;
	uint16_t* o = out + (uint64_t)z * fpx;
	const bool zflag = (video & (int)z & 1) != 0;

	auto decode_px = [&](int x, int y, int tx, int ty, int u, int v) {
		auto px = [&](int dx, int dy) { return (int)__ldcg(o + (size_t)(y + dy) * W + (x + dx)); };
		int p = predict0(px, T, way, k, tx, ty, u, v);
		if (zflag) {
			int P = (int)__ldcg(o + (size_t)y * W + x - fpx);    // previous (even) frame, reconstructed by an earlier launch
			p = (x == 0 && y == 0) ? P : ((p + P) >> 1);
		}
		o[(size_t)y * W + x] = (uint16_t)(unsymbolize16(__ldg(s + (size_t)y * W + x)) + p);
	};

	if (k == 2 && way != 2) {
		// ---- schedule B: rows. Leading pixels with left-neighbour chains are walked by one thread.
		for (int y = 0; y < H; y++) {
			const int ty = y / T, v = y - ty * T;
			const int seq = (y == 0) ? W : min(T, W);
			if (gtid == 0) for (int x = 0; x < seq; x++) decode_px(x, y, x / T, ty, x % T, v);
			for (int x = seq + (int)gtid - 1; x < W; x += (int)gsize - 1) if (gtid > 0) decode_px(x, y, x / T, ty, x % T, v);
			cluster.sync();
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
		for (uint32_t i = gtid; i < total; i += gsize) {
			int lo_r = 0, hi_r = nr - 1;                 // last r with pre[r] <= i
			while (lo_r < hi_r) { int mid = (lo_r + hi_r + 1) >> 1; if (pre[mid] <= i) lo_r = mid; else hi_r = mid - 1; }
			const int r = lo_r, sd = w - r;
			const uint32_t j = i - pre[r];
			const int np = min(r, 2 * T - 2 - r) + 1;
			const int ti = (int)(j / (uint32_t)np), pi = (int)(j - (uint32_t)ti * (uint32_t)np);
			const int tx = max(0, sd - (tilesY - 1)) + ti, ty = sd - tx;
			const int u = max(0, r - (T - 1)) + pi, v = r - u;
			const int x = tx * T + u, y = ty * T + v;
			if (x < W && y < H) decode_px(x, y, tx, ty, u, v);
		}
		cluster.sync();
	}
}

// ---------------------------------------------------------------------------------------------------------
// Barrier-free inverse for the two ways whose dependency graph factorises (no cross-CTA synchronisation at all):
//   way "space": after the first tile, pixel (tx,ty,u,v) only depends on the SAME (u,v) of tiles (tx-1,ty), (tx,ty-1),
//     (tx-1,ty-1): T*T independent 2-D recurrences over the tile grid -> one warp per (frame, u, v) walks the tile
//     anti-diagonals with __syncwarp() only.
//   way "angle" (predictor != 2): only the tile DC looks at other tiles (their DCs) -> one warp per frame solves
//     the DC recurrence over tile anti-diagonals, then every tile is an independent intra-tile DPCM: one warp per tile.
// k_unpredict_first_tile decodes tile (0,0) (plain intra-tile DPCM in every way) / the DC grid.
constexpr int UF_NT = 128;

// mode 0: tile (0,0) of every frame (way space);  mode 1: the DC pixel of every tile (way angle).  One warp per frame.
__global__ void __launch_bounds__(UF_NT)
k_unpredict_seed(const uint16_t* __restrict__ sym, uint16_t* out, int W, int H, int T, int way, int k,
                 uint32_t z_start, uint32_t z_step, uint32_t nframes, int mode)
{
	const uint32_t lane = lane_id();
	const uint32_t f = blockIdx.x * (UF_NT / 32) + warp_id();
	if (f >= nframes) return;
	const uint64_t fpx = (uint64_t)W * H;
	const uint32_t z = z_start + f * z_step;
	const uint16_t* s = sym + (uint64_t)z * fpx;
	uint16_t* o = out + (uint64_t)z * fpx;
	auto decode_px = [&](int x, int y, int tx, int ty, int u, int v) {
		auto px = [&](int dx, int dy) { return (int)__ldcg(o + (size_t)(y + dy) * W + (x + dx)); };
		o[(size_t)y * W + x] = (uint16_t)(unsymbolize16(__ldg(s + (size_t)y * W + x)) + predict0(px, T, way, k, tx, ty, u, v));
	};
	if (mode == 0) {
		const int tw = min(T, W), th = min(T, H);
		for (int d = 0; d <= tw + th - 2; d++) {             // pixel anti-diagonals of the first tile
			for (int u = (int)lane; u <= d; u += 32) { int v = d - u; if (u < tw && v < th) decode_px(u, v, 0, 0, u, v); }
			__syncwarp();
		}
	} else {
		const int tilesX = (W + T - 1) / T, tilesY = (H + T - 1) / T;
		for (int d = 0; d <= tilesX + tilesY - 2; d++) {     // tile anti-diagonals, DC pixel only
			const int lo = max(0, d - (tilesY - 1)), hi = min(d, tilesX - 1);
			for (int tx = lo + (int)lane; tx <= hi; tx += 32) { int ty = d - tx; decode_px(tx * T, ty * T, tx, ty, 0, 0); }
			__syncwarp();
		}
	}
}

// way space: one warp per (frame, u, v)
__global__ void __launch_bounds__(UF_NT)
k_unpredict_space(const uint16_t* __restrict__ sym, uint16_t* out, int W, int H, int T, int k,
                  uint32_t z_start, uint32_t z_step, uint32_t nframes)
{
	const uint32_t lane = lane_id();
	const uint32_t wg = blockIdx.x * (UF_NT / 32) + warp_id();
	const uint32_t per_frame = (uint32_t)T * T;
	const uint32_t f = wg / per_frame;
	if (f >= nframes) return;
	const uint32_t uv = wg - f * per_frame;
	const int v = (int)(uv / (uint32_t)T), u = (int)(uv - (uint32_t)v * T);
	const uint64_t fpx = (uint64_t)W * H;
	const uint32_t z = z_start + f * z_step;
	const uint16_t* s = sym + (uint64_t)z * fpx;
	uint16_t* o = out + (uint64_t)z * fpx;
	const int tilesX = (W + T - 1) / T, tilesY = (H + T - 1) / T;
	for (int d = 1; d <= tilesX + tilesY - 2; d++) {         // tile (0,0) was decoded by k_unpredict_seed
		const int lo = max(0, d - (tilesY - 1)), hi = min(d, tilesX - 1);
		for (int tx = lo + (int)lane; tx <= hi; tx += 32) {
			const int ty = d - tx, x = tx * T + u, y = ty * T + v;
			if (x < W && y < H) {
				auto px = [&](int dx, int dy) { return (int)__ldcg(o + (size_t)(y + dy) * W + (x + dx)); };
				o[(size_t)y * W + x] = (uint16_t)(unsymbolize16(__ldg(s + (size_t)y * W + x)) + predict0(px, T, 2, k, tx, ty, u, v));
			}
		}
		__syncwarp();
	}
}

// way angle (predictor != 2): one warp per (frame, tile); the DC is already in place
__global__ void __launch_bounds__(UF_NT)
k_unpredict_angle(const uint16_t* __restrict__ sym, uint16_t* out, int W, int H, int T, int k,
                  uint32_t z_start, uint32_t z_step, uint32_t nframes)
{
	const uint32_t lane = lane_id();
	const int tilesX = (W + T - 1) / T, tilesY = (H + T - 1) / T;
	const uint64_t wg = (uint64_t)blockIdx.x * (UF_NT / 32) + warp_id();
	const uint64_t per_frame = (uint64_t)tilesX * tilesY;
	const uint32_t f = (uint32_t)(wg / per_frame);
	if (f >= nframes) return;
	const uint32_t tile = (uint32_t)(wg - (uint64_t)f * per_frame);
	const int ty = (int)(tile / (uint32_t)tilesX), tx = (int)(tile - (uint32_t)ty * tilesX);
	const uint64_t fpx = (uint64_t)W * H;
	const uint32_t z = z_start + f * z_step;
	const uint16_t* s = sym + (uint64_t)z * fpx;
	uint16_t* o = out + (uint64_t)z * fpx;
	const int x0 = tx * T, y0 = ty * T, tw = min(T, W - x0), th = min(T, H - y0);
	for (int d = 1; d <= tw + th - 2; d++) {
		for (int u = (int)lane; u <= d; u += 32) {
			const int v = d - u;
			if (u < tw && v < th) {
				const int x = x0 + u, y = y0 + v;
				auto px = [&](int dx, int dy) { return (int)__ldcg(o + (size_t)(y + dy) * W + (x + dx)); };
				o[(size_t)y * W + x] = (uint16_t)(unsymbolize16(__ldg(s + (size_t)y * W + x)) + predict0(px, T, 1, k, tx, ty, u, v));
			}
		}
		__syncwarp();
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
// Returns 0, or 1 if the cluster launch fails.
int launch_unpredict(const uint16_t* sym, uint16_t* out, int W, int H, int T, int way, int k, int video,
                     uint32_t z_start, uint32_t z_step, uint32_t count, int sm_count, cudaStream_t st)
{
	(void)sm_count;
	if (count == 0) return 0;
	const int tilesX = (W + T - 1) / T, tilesY = (H + T - 1) / T;
	const unsigned wpb = UF_NT / 32;
	if (!video && way == 2) {                               // barrier-free: first tile, then T*T recurrences per frame
		k_unpredict_seed<<<(count + wpb - 1) / wpb, UF_NT, 0, st>>>(sym, out, W, H, T, way, k, z_start, z_step, count, 0);
		const uint64_t warps = (uint64_t)count * T * T;
		k_unpredict_space<<<(unsigned)((warps + wpb - 1) / wpb), UF_NT, 0, st>>>(sym, out, W, H, T, k, z_start, z_step, count);
		return cudaGetLastError() == cudaSuccess ? 0 : 1;
	}
	if (!video && way == 1 && k != 2) {                     // barrier-free: DC grid, then every tile on its own
		k_unpredict_seed<<<(count + wpb - 1) / wpb, UF_NT, 0, st>>>(sym, out, W, H, T, way, k, z_start, z_step, count, 1);
		const uint64_t warps = (uint64_t)count * tilesX * tilesY;
		k_unpredict_angle<<<(unsigned)((warps + wpb - 1) / wpb), UF_NT, 0, st>>>(sym, out, W, H, T, k, z_start, z_step, count);
		return cudaGetLastError() == cudaSuccess ? 0 : 1;
	}
	const uint64_t steps = (k == 2 && way != 2) ? (uint64_t)H : (uint64_t)(tilesX + tilesY + 2 * T);
	const uint64_t per_step = ((uint64_t)W * H + steps - 1) / steps;              // mean ready pixels per step and frame
	unsigned csz = 1;
	while (csz < 8 && (uint64_t)csz * UP_NT < per_step) csz <<= 1;                // about one pixel per thread per step
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(count * csz); cfg.blockDim = dim3(UP_NT); cfg.dynamicSmemBytes = 0; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, k_unpredict, sym, out, W, H, T, way, k, video, z_start, z_step, count) == cudaSuccess ? 0 : 1;
}

}  // namespace lfm
