// lfm_radix.cuh -- CTA-wide stable 8-bit counting-sort pass and group-splitting helpers shared by the
// BWT (bz_bwt.cu), the inverse BWT (bz_decode.cu) and the order-1 context sort of the mode selection (lfm_select.cu).
#pragma once
#include "lfm_device.cuh"

namespace lfm {

constexpr int BWT_NT = 1024;
constexpr int BWT_R  = 8;                 // elements per thread per radix tile
constexpr int BWT_NW = BWT_NT / 32;
constexpr int BWT_WS = 257;               // row stride (words) of the per-warp digit counters: conflict-free rows AND columns
constexpr int SPLIT_R = 4;                // elements per thread per tile in split_groups
#define LFM_WC(w, d) wcnt[(w) * BWT_WS + (d)]

struct Trip { uint32_t g, r, v; };

// One stable 8-bit counting-sort pass over m elements, tile by tile (tile = BWT_NT * BWT_R elements).
// run[256] (shared) must hold the exclusive bucket starts on entry; it is advanced as tiles are placed.
// wcnt: BWT_NW x BWT_WS shared words.
//   1. all loads of the tile are issued first (independent global loads in flight together);
//   2. per warp, BWT_R rounds of match.any multisplit give each element its rank among equal digits of the warp;
//   3. the per-warp digit counts are scanned ACROSS warps by warp shuffles (warp w owns digits 8w..8w+7);
//   4. scatter.
template <class P, class LoadFn, class DigitFn, class StoreFn>
__device__ __forceinline__ void radix_scatter(uint32_t m, uint32_t* run, uint32_t* wcnt,
                                              LoadFn load, DigitFn digit, StoreFn store)
{
	const uint32_t lane = lane_id(), w = warp_id();
	constexpr uint32_t TILE = BWT_NT * BWT_R;
	for (uint32_t t0 = 0; t0 < m; t0 += TILE) {
		for (uint32_t i = threadIdx.x; i < BWT_NW * BWT_WS; i += BWT_NT) wcnt[i] = 0;
		P pay[BWT_R]; uint32_t dl[BWT_R];                 // digit << 16 | rank inside the warp
		#pragma unroll
		for (int r = 0; r < BWT_R; r++) {
			uint32_t e = t0 + w * (32 * BWT_R) + r * 32 + lane;
			if (e < m) pay[r] = load(e);
		}
		__syncthreads();
		#pragma unroll
		for (int r = 0; r < BWT_R; r++) {
			uint32_t e = t0 + w * (32 * BWT_R) + r * 32 + lane;
			bool act = e < m;
			uint32_t amask = __ballot_sync(0xffffffffu, act);
			dl[r] = 0;
			if (act) {
				uint32_t d = digit(pay[r]);
				uint32_t peers = __match_any_sync(amask, d);
				uint32_t leader = __ffs(peers) - 1;
				uint32_t old = 0;
				if (lane == leader) { old = LFM_WC(w, d); LFM_WC(w, d) = old + __popc(peers); }
				old = __shfl_sync(peers, old, leader);
				dl[r] = (d << 16) | (old + __popc(peers & ((1u << lane) - 1u)));
			}
			__syncwarp();
		}
		__syncthreads();
		#pragma unroll
		for (int q = 0; q < 256 / BWT_NW; q++) {          // exclusive scan over the warps, digit by digit
			const uint32_t d = w * (256 / BWT_NW) + q;
			const uint32_t v = LFM_WC(lane, d);
			uint32_t incl = v;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
			const uint32_t base = run[d];
			LFM_WC(lane, d) = base + incl - v;
			if (lane == 31) run[d] = base + incl;
		}
		__syncthreads();
		#pragma unroll
		for (int r = 0; r < BWT_R; r++) {
			uint32_t e = t0 + w * (32 * BWT_R) + r * 32 + lane;
			if (e < m) store(LFM_WC(w, dl[r] >> 16) + (dl[r] & 0xffffu), pay[r]);
		}
		__syncthreads();
	}
}

// per-warp private histograms (match.any aggregated, no atomics) of an 8-bit digit over m elements,
// reduced and exclusive-scanned into run[256]
template <class DigitOfIndex>
__device__ __forceinline__ void digit_starts(uint32_t m, uint32_t* run, uint32_t* wcnt, uint32_t* red, DigitOfIndex dig)
{
	const uint32_t lane = lane_id(), w = warp_id();
	for (uint32_t i = threadIdx.x; i < BWT_NW * BWT_WS; i += BWT_NT) wcnt[i] = 0;
	__syncthreads();
	for (uint32_t e0 = 0; e0 < m; e0 += BWT_NT) {
		uint32_t e = e0 + threadIdx.x;
		bool act = e < m;
		uint32_t amask = __ballot_sync(0xffffffffu, act);
		if (act) {
			uint32_t d = dig(e);
			uint32_t peers = __match_any_sync(amask, d);
			if (lane == (uint32_t)(__ffs(peers) - 1)) LFM_WC(w, d) += __popc(peers);
		}
		__syncwarp();
	}
	__syncthreads();
	if (threadIdx.x < 256) {
		uint32_t s = 0;
		#pragma unroll 8
		for (int ww = 0; ww < BWT_NW; ww++) s += LFM_WC(ww, threadIdx.x);
		run[threadIdx.x] = s;
	}
	__syncthreads();
	scan256_excl<BWT_NT>(run, red);
}

// Split a sorted list of `cnt` entries into groups of equal keys.
//   is_head(i)   : entry i starts a new group (must return true for i == 0 and i >= cnt)
//   pos_of(i)    : position of entry i in the suffix array (ascending in i)
//   emit(i, head_pos, unresolved, slot) : called once per entry; slot = index among the unresolved entries
// returns the number of unresolved entries (members of groups larger than 1). Uniform across the CTA.
template <class HeadFn, class PosFn, class EmitFn>
__device__ __forceinline__ uint32_t split_groups(uint32_t cnt, uint32_t* red, HeadFn is_head, PosFn pos_of, EmitFn emit)
{
	uint32_t carry_head = 0, carry_cnt = 0;
	for (uint32_t t0 = 0; t0 < cnt; t0 += BWT_NT * SPLIT_R) {
		uint32_t i0 = t0 + threadIdx.x * SPLIT_R;
		bool hd[SPLIT_R + 1]; uint32_t gh[SPLIT_R];
		#pragma unroll
		for (int r = 0; r <= SPLIT_R; r++) hd[r] = is_head(i0 + r);
		uint32_t local = 0;
		#pragma unroll
		for (int r = 0; r < SPLIT_R; r++) { if (i0 + r < cnt && hd[r]) local = pos_of(i0 + r); gh[r] = local; }
		uint32_t incl = block_scan_max<BWT_NT>(local, red);
		uint32_t tile_max = red[BWT_NW - 1];
		uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1);
		if (lane_id() == 0) before = warp_id() ? red[warp_id() - 1] : 0;
		before = max(before, carry_head);
		uint32_t c = 0; bool un[SPLIT_R];
		#pragma unroll
		for (int r = 0; r < SPLIT_R; r++) {
			gh[r] = max(gh[r], before);
			un[r] = (i0 + r < cnt) && !(hd[r] && hd[r + 1]);
			c += un[r];
		}
		uint32_t tot; uint32_t inc = block_scan_add<BWT_NT>(c, red, &tot);
		uint32_t o = carry_cnt + inc - c;
		#pragma unroll
		for (int r = 0; r < SPLIT_R; r++) if (i0 + r < cnt) { emit(i0 + r, gh[r], un[r], o); o += un[r]; }
		carry_cnt += tot;
		carry_head = max(carry_head, tile_max);
	}
	__syncthreads();
	return carry_cnt;
}

}  // namespace lfm
