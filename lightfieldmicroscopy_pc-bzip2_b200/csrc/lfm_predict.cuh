// lfm_predict.cuh -- the light-field DPCM prediction rules, one device function shared by the forward
// (lfm_predict.cu: k_predict_fwd) and the inverse (k_unpredict) kernels.
//
// Behavioural spec: the 21 reference kernels _predictorN_{tiles,angle,space}
//   src/lfm_Predictors.cu:35-1334, src/lfm_Predictors_angle.cu:17-1279, src/lfm_Predictors_space.cu:17-1213
// (z == 0 branches; the z != 0 branch of way "tiles" adds the previous frame, see predict_px()).
// T = Nnum; tile (tx,ty), in-tile (u,v); px(dx,dy) fetches the pixel at (x+dx, y+dy) of the current frame.
// Every operand sits at strictly smaller tx+ty+u+v, which is what the inverse wavefront relies on.
#pragma once
#include <cstdint>

namespace lfm {

template <class Fetch>
__device__ __forceinline__ int predict0(Fetch px, int T, int way, int k, int tx, int ty, int u, int v)
{
	const bool t00 = (tx == 0 && ty == 0), tcol0 = (tx == 0 && ty != 0), trow0 = (tx != 0 && ty == 0);
	if (t00) {                                    // first tile: plain intra-tile DPCM borders, all ways
		if (u == 0 && v == 0) return 0;
		if (u == 0) return px(0, -1);
		if (v == 0) return px(-1, 0);
	}
	if (way == 2 && !t00) {                       // space: the same sub-aperture pixel of neighbouring microlenses
		if (tcol0) return px(0, -T);
		if (trow0) return px(-T, 0);
		switch (k) {
		case 1: return px(-T, 0);
		case 2: return px(0, -T);
		case 3: return px(-T, -T);
		case 4: return px(0, -T) + px(-T, 0) - px(-T, -T);
		case 5: return px(0, -T) + ((px(-T, 0) - px(-T, -T)) >> 1);
		case 6: return px(-T, 0) + ((px(0, -T) - px(-T, -T)) >> 1);
		default: return (u > 0 && v > 0) ? ((px(0, -T) + px(-T, 0)) >> 1) : (px(0, -T) + px(-T, 0) - px(-T, -T));
		}
	}
	if (way == 1 || (way == 2 && t00)) {          // angle: neighbours under the same microlens; DC from the next tile
		if (u == 0 && v == 0) {
			if (tcol0) return px(0, -T);
			if (trow0) return px(-T, 0);
			switch (k) {
			case 1: return px(-T, 0);
			case 2: return px(0, -T);
			case 3: return px(-T, -T);
			case 4: case 7: return px(0, -T) + px(-T, 0) - px(-T, -T);
			case 5: return px(0, -T) + ((px(-T, 0) - px(-T, -T)) >> 1);
			default: return px(-T, 0) + ((px(0, -T) - px(-T, -T)) >> 1);
			}
		}
		if (u == 0) return px(0, -1);
		const bool tin = !(t00 || tcol0 || trow0);
		if (v == 0) return (k == 2 && tin) ? px(0, -1) : px(-1, 0);
		int kk = k;
		if (tin && (k == 5 || k == 6)) kk = 11 - k;           // 5 and 6 trade places in interior tiles
		switch (kk) {
		case 1: return px(-1, 0);
		case 2: return px(0, -1);
		case 3: return px(-1, -1);
		case 4: return px(-1, 0) + px(0, -1) - px(-1, -1);
		case 5: return px(-1, 0) + ((px(0, -1) - px(-1, -1)) >> 1);
		case 6: return px(0, -1) + ((px(-1, 0) - px(-1, -1)) >> 1);
		default: return (px(-1, 0) + px(0, -1)) >> 1;
		}
	}
	// way 0: "tiles" = blend of both
	const bool a = (u == 0 && v > 0), b = (u == 0 && v == 0), c = (u > 0 && v == 0);
	if (t00) {
		switch (k) {
		case 1: return px(-1, 0);
		case 2: return px(0, -1);
		case 3: return px(-1, -1);
		case 4: return px(-1, 0) + px(0, -1) - px(-1, -1);
		case 5: return (px(-1, 0) + (px(0, -1) - px(-1, -1))) >> 1;    // sic: the shift binds last (lfm_Predictors.cu:744)
		case 6: return px(0, -1) + ((px(-1, 0) - px(-1, -1)) >> 1);
		default: return (px(-1, 0) + px(0, -1)) >> 1;
		}
	}
	if (k == 1) {
		if (tcol0) return a ? px(0, -1) : b ? px(0, -T) : px(-1, 0);
		return (a || b) ? px(-T, 0) : ((px(-1, 0) + px(-T, 0)) >> 1);
	}
	if (k == 2) {
		if (tcol0) return (b || c) ? px(0, -T) : ((px(0, -1) + px(0, -T)) >> 1);
		if (trow0) return b ? px(-T, 0) : c ? px(-1, 0) : px(0, -1);
		return (a || b) ? px(0, -T) : ((px(0, -1) + px(0, -T)) >> 1);
	}
	if (k == 3) {
		const int far = tcol0 ? px(0, -T) : trow0 ? px(-T, 0) : px(-T, -T);
		if (b) return far;
		const int near = a ? px(0, -1) : c ? px(-1, 0) : px(-1, -1);
		return (near + far) >> 1;
	}
	if (tcol0 || trow0) {
		const int far = tcol0 ? px(0, -T) : px(-T, 0);
		if (b) return far;
		if (a) return (px(0, -1) + far) >> 1;
		if (c) return (px(-1, 0) + far) >> 1;
		switch (k) {
		case 4: return (px(-1, 0) + px(0, -1) - px(-1, -1) + far) >> 1;
		case 5: return (px(-1, 0) + ((px(0, -1) - px(-1, -1)) >> 1) + far) >> 1;
		case 6: return (px(0, -1) + ((px(-1, 0) - px(-1, -1)) >> 1) + far) >> 1;
		default:
			if (tcol0) return (px(-1, 0) + px(0, -1) + px(-1, -T) + px(0, -T - 1)) >> 2;
			return (px(-1, 0) + px(0, -1) + px(-T, -1) + px(-T - 1, 0)) >> 2;
		}
	}
	int g;
	switch (k) {
	case 4: case 7: g = px(0, -T) + px(-T, 0) - px(-T, -T); break;
	case 5: g = px(0, -T) + ((px(-T, 0) - px(-T, -T)) >> 1); break;
	default: g = px(-T, 0) + ((px(0, -T) - px(-T, -T)) >> 1); break;
	}
	if (b) return g;
	if (a) return (g + px(0, -1)) >> 1;
	if (c) return (g + px(-1, 0)) >> 1;
	switch (k) {
	case 4: return (g + px(0, -1) + px(-1, 0) - px(-1, -1)) >> 1;
	case 5: return (g + px(0, -1) + ((px(-1, 0) - px(-1, -1)) >> 1)) >> 1;
	case 6: return (g + px(-1, 0) + ((px(0, -1) - px(-1, -1)) >> 1)) >> 1;
	default: return (px(0, -T - 1) + px(-T - 1, 0) + px(0, -1) + px(-1, 0)) >> 2;
	}
}

// Straight-line form of predict0 for the bulk of the image: INTERIOR tiles (tx > 0 and ty > 0) and, for the ways that use
// near neighbours (tiles, angle), rows with v > 0.  WAY and K are compile-time, the only data-dependent choice left is
// u == 0 (a select).  near(dx,dy) is only called with the literal offsets (-1,0) (0,-1) (-1,-1), so a caller may serve
// them from registers; far(dx,dy) gets the offsets that involve T.  Must agree with predict0 on its domain -- the
// forward kernel uses predict0 everywhere else.
template <int WAY, int K, class Near, class Far>
__device__ __forceinline__ int predict_interior(Near near, Far far, int T, int u, int v)
{
	if (WAY == 2) {
		const int fu = far(0, -T), fl = far(-T, 0);
		if (K == 1) return fl;
		if (K == 2) return fu;
		const int ful = far(-T, -T);
		if (K == 3) return ful;
		if (K == 4) return fu + fl - ful;
		if (K == 5) return fu + ((fl - ful) >> 1);
		if (K == 6) return fl + ((fu - ful) >> 1);
		return (u > 0 && v > 0) ? ((fu + fl) >> 1) : (fu + fl - ful);
	}
	const int up = near(0, -1);
	if (WAY == 1) {                                   // v > 0: u == 0 -> up; else the near-neighbour rule (5 and 6 trade places)
		const int left = near(-1, 0);
		int gen;
		if (K == 1) gen = left;
		else if (K == 2) gen = up;
		else {
			const int ul = near(-1, -1);
			if (K == 3) gen = ul;
			else if (K == 4) gen = left + up - ul;
			else if (K == 5) gen = up + ((left - ul) >> 1);          // kk = 6
			else if (K == 6) gen = left + ((up - ul) >> 1);          // kk = 5
			else gen = (left + up) >> 1;
		}
		return u == 0 ? up : gen;
	}
	// WAY 0 (tiles), v > 0: a = (u == 0); b and c cannot occur
	const bool a = (u == 0);
	if (K == 1) { const int fl = far(-T, 0); return a ? fl : ((near(-1, 0) + fl) >> 1); }
	if (K == 2) { const int fu = far(0, -T); return a ? fu : ((up + fu) >> 1); }
	if (K == 3) { const int f3 = far(-T, -T); return ((a ? up : near(-1, -1)) + f3) >> 1; }
	const int left = near(-1, 0);
	if (K == 7) {
		const int g = far(0, -T) + far(-T, 0) - far(-T, -T);
		return a ? ((g + up) >> 1) : ((far(0, -T - 1) + far(-T - 1, 0) + up + left) >> 2);
	}
	const int fu = far(0, -T), fl = far(-T, 0), ful = far(-T, -T), ul = near(-1, -1);
	int g, full;
	if (K == 4) { g = fu + fl - ful; full = (g + up + left - ul) >> 1; }
	else if (K == 5) { g = fu + ((fl - ful) >> 1); full = (g + up + ((left - ul) >> 1)) >> 1; }
	else { g = fl + ((fu - ful) >> 1); full = (g + left + ((up - ul) >> 1)) >> 1; }
	return a ? ((g + up) >> 1) : full;
}

// The same for the FIRST row of an interior tile (tx > 0, ty > 0, v == 0) of the ways tiles and angle -- one row in Nnum,
// but through predict0 it cost as much as five interior rows.  near(-1,0) = left, near(0,-1) = the pixel above (last row of
// the tile above); must agree with predict0 (cases b: u == 0, c: u > 0).
template <int WAY, int K, class Near, class Far>
__device__ __forceinline__ int predict_interior_v0(Near near, Far far, int T, int u)
{
	int g;                                              // the far-neighbour term (DC rule of way angle, g of way tiles)
	if (K == 1) g = far(-T, 0);
	else if (K == 2) g = far(0, -T);
	else if (K == 3) g = far(-T, -T);
	else {
		const int fu = far(0, -T), fl = far(-T, 0), ful = far(-T, -T);
		if (K == 4 || K == 7) g = fu + fl - ful;
		else if (K == 5) g = fu + ((fl - ful) >> 1);
		else g = fl + ((fu - ful) >> 1);
	}
	if (WAY == 1) return u == 0 ? g : (K == 2 ? near(0, -1) : near(-1, 0));
	// WAY 0
	if (K == 2) return u == 0 ? g : ((near(0, -1) + g) >> 1);
	return u == 0 ? g : ((near(-1, 0) + g) >> 1);
}

// zig-zag residual <-> symbol map (lfm_Predictors.cu:16-33), on the int16-truncated residual
__device__ __forceinline__ uint16_t symbolize16(int residual)
{
	const int r = (int)(int16_t)residual;
	return (uint16_t)((r << 1) ^ (r >> 31));       // r >= 0: 2r;  r < 0: ~(2r) = 2|r| - 1   (== 2*abs(r) + (r >> 31))
}
__device__ __forceinline__ int unsymbolize16(uint16_t s)
{
	return ((int)s >> 1) ^ -((int)s & 1);         // even: s/2;  odd: ~(s >> 1) = -(s + 1)/2   (== (1 - 2 neg) * ((s + neg) / 2))
}

}  // namespace lfm
