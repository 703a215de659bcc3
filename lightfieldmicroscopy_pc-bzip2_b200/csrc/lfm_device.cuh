// Shared device-side definitions for the B200 LFM engine (sm_100a).
// Data layout in HBM (see DESIGN.md "Data layout"):
//   image / symbol image : uint16 [frames][H][W], x fastest (the caller's layout, klb_imageIO.cpp:133-183)
//   per KLB-block ("job") working arrays are strided by `cap` = max post-RLE1 length rounded to 16
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lfm {

constexpr int kMaxAlpha = 258;
constexpr int kGroups   = 6;
constexpr int kGSize    = 50;
constexpr int kMaxSel   = 18002;     // compress.c:456 (2 + 900000/50)

// geometry of the KLB block grid (klb_imageIO.cpp:97-183)
struct Geom {
	uint32_t xyzct[5];
	uint32_t bs[5];
	uint32_t nb[5];
	uint64_t stride[5];   // in pixels
};

// One KLB block = one bzip2 stream = one or more bzip2 blocks (bzlib.c:  a block is closed once it holds
// nblockMAX = 100000 * level - 19 run-length coded bytes).  A KLB block owns `nsub` consecutive job records ("sub-jobs",
// one per possible bzip2 block, nsub = 1 for every block shape that fits one bzip2 block): record job * nsub + k describes
// bzip2 block k of KLB block `job`; all per-block working arrays are strided by the sub-slot size.
constexpr uint32_t kSubUnused = 1u;    // EncJob/DecJob.flags: this record holds no bzip2 block
constexpr uint32_t kSubFirst  = 2u;    // first bzip2 block of its stream (the stream header precedes it)
constexpr uint32_t kSubLast   = 4u;    // last bzip2 block of its stream (the stream trailer follows it)
constexpr uint32_t kSubRand   = 8u;    // decoder: the block carries bzip2's "randomised" bit (streams of bzip2 <= 0.9.0)
constexpr int kMaxSub = 32;            // bzip2 blocks per KLB block this engine handles

struct EncJob {
	uint32_t raw_bytes;    // bytes gathered from the image (gcount, klb_imageIO.cpp:146-151) -- of the whole KLB block
	uint32_t n;            // post-RLE1 length (nblock)
	uint32_t crc;          // block CRC over the pre-RLE bytes
	uint32_t orig_ptr;
	uint32_t n_mtf;
	uint32_t n_in_use;
	uint32_t n_groups;
	uint32_t n_sel;
	uint32_t total_bits;   // bits this record contributes to the stream (before byte padding)
	uint32_t out_bytes;    // bytes of its slot that hold them
	uint32_t periodic;
	uint32_t status;       // 0 ok
	uint32_t in_use[8];    // 256-bit map
	uint32_t flags;        // kSub*
	uint32_t stream_crc;   // last record of a stream: the combined CRC of all its blocks
	uint32_t pad[2];
};

struct DecJob {
	uint32_t n;            // post-RLE1 length of the bzip2 block
	uint32_t n_mtf;        // Huffman-coded symbols incl. EOB
	uint32_t n_in_use;
	uint32_t orig_ptr;
	uint32_t stored_crc;
	uint32_t out_bytes;    // bytes produced by un-RLE1
	uint32_t status;       // 0 ok, 1 bad magic, 2 corrupt, 3 crc mismatch, 4 unsupported (more bzip2 blocks than the geometry allows)
	uint32_t level;
	uint32_t max_block;
	uint32_t flags;        // kSub*
	uint32_t pad[2];
	uint32_t in_use[8];    // 256-bit symbol map of the block
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// block coordinates of KLB block `id`: origin c0[5] and clamped extent ext[5]
__device__ __forceinline__ void block_box(const Geom& g, uint64_t id, uint32_t c0[5], uint32_t ext[5])
{
	#pragma unroll
	for (int i = 0; i < 5; i++) {
		uint32_t c = (uint32_t)(id % g.nb[i]); id /= g.nb[i];
		c0[i] = c * g.bs[i];
		uint32_t rem = g.xyzct[i] - c0[i];
		ext[i] = rem < g.bs[i] ? rem : g.bs[i];
	}
}

// ---- block-wide scans over one value per thread; `red` is a shared array of >= 33 uint32_t ----
// returns the inclusive scan; *total = sum over the block. Contains __syncthreads: call from all threads.
template <int NT>
__device__ __forceinline__ uint32_t block_scan_add(uint32_t v, uint32_t* red, uint32_t* total)
{
	uint32_t lane = lane_id(), w = warp_id();
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
	__syncthreads();
	if (lane == 31) red[w] = v;
	__syncthreads();
	if (w == 0) {
		uint32_t s = lane < NT / 32 ? red[lane] : 0;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
		red[lane] = s;
	}
	__syncthreads();
	uint32_t off = w ? red[w - 1] : 0;
	*total = red[NT / 32 - 1];
	return v + off;
}
template <int NT>
__device__ __forceinline__ uint32_t block_scan_max(uint32_t v, uint32_t* red)
{
	uint32_t lane = lane_id(), w = warp_id();
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = max(v, t); }
	__syncthreads();
	if (lane == 31) red[w] = v;
	__syncthreads();
	if (w == 0) {
		uint32_t s = lane < NT / 32 ? red[lane] : 0;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s = max(s, t); }
		red[lane] = s;
	}
	__syncthreads();
	uint32_t off = w ? red[w - 1] : 0;
	return max(v, off);
}

// exclusive scan of 256 counters held in shared memory (in place); NT >= 256. *total receives the sum.
template <int NT>
__device__ __forceinline__ void scan256_excl(uint32_t* arr, uint32_t* red)
{
	uint32_t v = threadIdx.x < 256 ? arr[threadIdx.x] : 0, tot;
	uint32_t inc = block_scan_add<NT>(v, red, &tot);
	if (threadIdx.x < 256) arr[threadIdx.x] = inc - v;
	__syncthreads();
}

// Sequential per-thread byte reader (global, shared or generic pointers): 16 bytes per load with one block of lookahead,
// the byte comes out of registers -- a thread that walks its chunk byte by byte otherwise pays one memory latency per
// byte.  The base must be 16-byte aligned; 16-byte blocks that START below `lim` are loaded whole (buffers are padded).
struct ByteReader {
	const uint4* base; uint32_t blk, lim; uint4 cur, nxt;
	__device__ __forceinline__ void init(const uint8_t* p, uint32_t i, uint32_t lim_) {
		base = reinterpret_cast<const uint4*>(p); lim = lim_; blk = i >> 4; cur = base[blk];
		nxt = ((blk + 1) << 4) < lim ? base[blk + 1] : make_uint4(0, 0, 0, 0);
	}
	__device__ __forceinline__ uint32_t get(uint32_t i) {
		const uint32_t b = i >> 4;
		if (b != blk) {
			cur = (b == blk + 1) ? nxt : base[b];
			blk = b;
			nxt = ((b + 1) << 4) < lim ? base[b + 1] : make_uint4(0, 0, 0, 0);
		}
		const uint32_t w = (i >> 2) & 3u;
		const uint32_t v = w == 0 ? cur.x : w == 1 ? cur.y : w == 2 ? cur.z : cur.w;
		return (v >> ((i & 3u) * 8u)) & 0xffu;
	}
};

// bzip2 CRC-32 table (poly 0x04C11DB7, MSB first; crctable.c) computed on the fly
__device__ __forceinline__ uint32_t crc_table_entry(uint32_t i)
{
	uint32_t c = i << 24;
	#pragma unroll
	for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
	return c;
}

}  // namespace lfm
