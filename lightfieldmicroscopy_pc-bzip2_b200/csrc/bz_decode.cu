// bz_decode.cu -- bzip2 block decoder for sm_100a.
//
// Replaces BZ2_bzBuffToBuffDecompress(dst,&n,src,len,0,0) as called per KLB block by
// klb_imageIO::blockUncompressor{,ImageFull} (src/klb_imageIO.cpp:627, :1034) plus the scatter of the block into
// the image (:673-746, :1080-1125):
//   k_decode      stream/block header, coding tables, Huffman decode, inverse MTF, RUNA/RUNB expansion
//                 (decompress.c:196-487, huffman.c:170-205)            -> last column L + symbol counts
//   k_inv_bwt     inverse BWT: stable counting sort = LF mapping (decompress.c:494-573), then the n-step walk is
//                 cut at splitters and walked by 1024 threads in parallel (list ranking by sampling)
//   k_unrle       inverse of the initial run-length coding + CRC check (bzlib.c:561-728) fused with the scatter
//                 of the block into the symbol image
#include "lfm_radix.cuh"

namespace lfm {

// =====================================================================================================
// k_decode : one warp per KLB block stream; lane 0 decodes (sequential by nature), blocks in parallel
// =====================================================================================================
constexpr int DEC_NT = 128;
constexpr int DEC_NW = DEC_NT / 32;

struct BitReader {
	const uint8_t* p; uint64_t nbytes; uint64_t pos; uint64_t buf; uint32_t cnt; bool overrun;
	__device__ void init(const uint8_t* p_, uint64_t n_) { p = p_; nbytes = n_; pos = 0; buf = 0; cnt = 0; overrun = false; }
	__device__ __forceinline__ void refill() {
		while (cnt <= 56) {
			uint32_t b = 0;
			if (pos < nbytes) b = __ldg(p + pos); else if (pos >= nbytes + 8) overrun = true;
			pos++;
			buf |= (uint64_t)b << (56 - cnt);
			cnt += 8;
		}
	}
	__device__ __forceinline__ uint32_t get(uint32_t nb) {    // nb <= 32
		if (cnt < nb) refill();
		uint32_t v = (uint32_t)(buf >> (64 - nb));
		buf <<= nb; cnt -= nb;
		return v;
	}
	__device__ __forceinline__ uint32_t peek(uint32_t nb) { if (cnt < nb) refill(); return (uint32_t)(buf >> (64 - nb)); }
	__device__ __forceinline__ void skip(uint32_t nb) { buf <<= nb; cnt -= nb; }
	__device__ uint64_t bits_used() const { return pos * 8 - cnt; }
};

__global__ void __launch_bounds__(DEC_NT)
k_decode(const uint8_t* __restrict__ payload, const uint64_t* __restrict__ begin, const uint64_t* __restrict__ end,
         uint32_t njobs, DecJob* __restrict__ jobs, uint8_t* __restrict__ bwt_all, uint32_t cap,
         uint8_t* __restrict__ sel_all, uint32_t selcap)
{
	__shared__ int32_t  s_limit[DEC_NW][kGroups][24];
	__shared__ int32_t  s_base[DEC_NW][kGroups][24];
	__shared__ uint16_t s_perm[DEC_NW][kGroups][kMaxAlpha + 2];
	__shared__ uint8_t  s_len[DEC_NW][kMaxAlpha + 2];
	__shared__ uint8_t  s_unseq[DEC_NW][256];
	__shared__ uint8_t  s_mtf[DEC_NW][256];
	__shared__ uint32_t s_cnt[DEC_NW][256];
	__shared__ int32_t  s_minlen[DEC_NW][kGroups];

	const uint32_t w = warp_id();
	const uint32_t job = blockIdx.x * DEC_NW + w;
	if (job >= njobs || lane_id() != 0) return;
	DecJob& J = jobs[job];
	uint8_t* L = bwt_all + (size_t)job * cap;
	uint8_t* selector = sel_all + (size_t)job * selcap;
	BitReader br; br.init(payload + begin[job], end[job] - begin[job]);
	J.n = 0; J.out_bytes = 0; J.orig_ptr = 0; J.stored_crc = 0;
	#define FAIL(code) do { J.status = (code); return; } while (0)

	if (br.get(8) != 'B' || br.get(8) != 'Z' || br.get(8) != 'h') FAIL(1);
	int level = (int)br.get(8) - '0';
	if (level < 1 || level > 9) FAIL(1);
	J.level = (uint32_t)level;
	const uint32_t max_block = min((uint32_t)(100000 * level), cap);
	uint32_t m1 = br.get(24), m2 = br.get(24);
	if (m1 == 0x177245 && m2 == 0x385090) { J.status = 0; for (int i = 0; i <= 256; i++) J.cftab[i] = 0; return; }   // empty stream
	if (m1 != 0x314159 || m2 != 0x265359) FAIL(2);
	J.stored_crc = br.get(32);
	if (br.get(1)) FAIL(4);                                   // randomised blocks are never produced (compress.c:629)
	const uint32_t orig_ptr = br.get(24);
	uint32_t n_in_use = 0;
	{
		uint32_t used16 = br.get(16);
		for (int i = 0; i < 16; i++) if (used16 & (0x8000u >> i)) {
			uint32_t bits = br.get(16);
			for (int j = 0; j < 16; j++) if (bits & (0x8000u >> j)) s_unseq[w][n_in_use++] = (uint8_t)(i * 16 + j);
		}
	}
	if (n_in_use == 0) FAIL(2);
	const int alpha = (int)n_in_use + 2;
	const int n_groups = (int)br.get(3);
	const int n_sel = (int)br.get(15);
	if (n_groups < 2 || n_groups > 6 || n_sel < 1 || (uint32_t)n_sel > selcap) FAIL(2);
	{
		uint8_t pos[kGroups];
		for (int i = 0; i < n_groups; i++) pos[i] = (uint8_t)i;
		for (int i = 0; i < n_sel; i++) {
			int j = 0;
			while (br.get(1)) { j++; if (j >= n_groups) FAIL(2); }
			uint8_t t = pos[j];
			for (; j > 0; j--) pos[j] = pos[j - 1];
			pos[0] = t; selector[i] = t;
		}
	}
	for (int t = 0; t < n_groups; t++) {
		int curr = (int)br.get(5);
		for (int i = 0; i < alpha; i++) {
			for (;;) {
				if (curr < 1 || curr > 20) FAIL(2);
				if (!br.get(1)) break;
				if (br.get(1)) curr--; else curr++;
			}
			s_len[w][i] = (uint8_t)curr;
		}
		// decode tables (huffman.c:170-205); perm by counting sort on the length (stable in the symbol)
		int mn = 32, mx = 0;
		int32_t* base = s_base[w][t]; int32_t* limit = s_limit[w][t];
		for (int i = 0; i < 24; i++) { base[i] = 0; limit[i] = 0; }
		for (int i = 0; i < alpha; i++) { int l = s_len[w][i]; mx = l > mx ? l : mx; mn = l < mn ? l : mn; base[l + 1]++; }
		for (int i = 1; i < 23; i++) base[i] += base[i - 1];
		{
			int32_t nxt[24];
			for (int i = 0; i < 24; i++) nxt[i] = base[i];
			for (int i = 0; i < alpha; i++) { int l = s_len[w][i]; s_perm[w][t][nxt[l]++] = (uint16_t)i; }
		}
		int vec = 0;
		for (int i = mn; i <= mx; i++) { vec += base[i + 1] - base[i]; limit[i] = vec - 1; vec <<= 1; }
		for (int i = mn + 1; i <= mx; i++) base[i] = ((limit[i - 1] + 1) << 1) - base[i];
		s_minlen[w][t] = mn;
	}
	if (br.overrun) FAIL(2);

	// ---- MTF / run decoding (decompress.c:349-487)
	for (int i = 0; i < 256; i++) { s_mtf[w][i] = (uint8_t)i; s_cnt[w][i] = 0; }
	const int EOB = (int)n_in_use + 1;
	uint32_t nblock = 0, acc = 0;
	auto put = [&](uint32_t b) {
		acc |= b << ((nblock & 3) * 8);
		nblock++;
		if ((nblock & 3) == 0) { *reinterpret_cast<uint32_t*>(L + nblock - 4) = acc; acc = 0; }
	};
	int grp = -1, left = 0, t = 0;
	uint32_t run = 0, run_w = 1; bool in_run = false;
	const int32_t* limit = nullptr; const int32_t* base = nullptr; const uint16_t* perm = nullptr; int mn = 0;
	for (;;) {
		if (left == 0) {
			grp++; if (grp >= n_sel) FAIL(2);
			left = kGSize; t = selector[grp];
			limit = s_limit[w][t]; base = s_base[w][t]; perm = s_perm[w][t]; mn = s_minlen[w][t];
		}
		left--;
		uint32_t window = br.peek(20);
		int zn = mn; int32_t zvec = (int32_t)(window >> (20 - zn));
		for (;;) {
			if (zn > 20) FAIL(2);
			if (zvec <= limit[zn]) break;
			zn++;
			if (zn <= 20) zvec = (int32_t)(window >> (20 - zn));
		}
		br.skip((uint32_t)zn);
		int32_t idx = zvec - base[zn];
		if (idx < 0 || idx >= kMaxAlpha) FAIL(2);
		int sym = perm[idx];
		if (sym <= 1) {
			if (!in_run) { in_run = true; run = 0; run_w = 1; }
			run += (uint32_t)(sym + 1) * run_w; run_w <<= 1;
			if (run > max_block) FAIL(2);
			continue;
		}
		if (in_run) {
			uint32_t uc = s_unseq[w][s_mtf[w][0]];
			if (nblock + run > max_block) FAIL(2);
			s_cnt[w][uc] += run;
			for (uint32_t i = 0; i < run; i++) put(uc);
			in_run = false;
		}
		if (sym == EOB) break;
		if (nblock >= max_block) FAIL(2);
		{
			int p = sym - 1; uint8_t v = s_mtf[w][p];
			for (; p > 0; p--) s_mtf[w][p] = s_mtf[w][p - 1];
			s_mtf[w][0] = v;
			uint32_t uc = s_unseq[w][v];
			s_cnt[w][uc]++; put(uc);
		}
		if (br.overrun) FAIL(2);
	}
	const uint32_t n = nblock;
	while (nblock & 3) put(0);
	if (orig_ptr >= n) FAIL(2);
	// the stream must end here: end-of-stream magic + combined CRC (single block: == block CRC)
	uint32_t e1 = br.get(24), e2 = br.get(24);
	if (e1 == 0x314159 && e2 == 0x265359) FAIL(4);            // multi-block stream: not supported in this version
	if (e1 != 0x177245 || e2 != 0x385090) FAIL(2);
	uint32_t combined = br.get(32);
	if (combined != J.stored_crc) FAIL(3);
	uint32_t s = 0;
	for (int i = 0; i < 256; i++) { J.cftab[i] = s; s += s_cnt[w][i]; }
	J.cftab[256] = s;
	J.n = n; J.orig_ptr = orig_ptr; J.status = 0;
	#undef FAIL
}

// =====================================================================================================
// k_inv_bwt : one CTA per block
// =====================================================================================================
constexpr int IB_MAXS = 1500;     // max number of splitters
constexpr int IB_VIS  = 2 * (IB_MAXS + 2);   // max sublist visits (a periodic block laps its cycle)

__global__ void __launch_bounds__(BWT_NT, 1)
k_inv_bwt(const uint8_t* __restrict__ bwt_all, uint32_t cap, DecJob* __restrict__ jobs, uint32_t njobs,
          uint32_t* __restrict__ tt_all, uint8_t* __restrict__ txt_all)
{
	__shared__ uint32_t wcnt[BWT_NW][256];
	__shared__ uint32_t run[256];
	__shared__ uint32_t s_next[IB_MAXS + 2];
	__shared__ uint32_t s_len[IB_MAXS + 2];
	__shared__ uint32_t s_nvis;
	const uint32_t tid = threadIdx.x;
	uint32_t* tt = tt_all + (size_t)blockIdx.x * cap;          // per-CTA scratch
	uint32_t* vis = tt_all + (size_t)gridDim.x * cap + (size_t)blockIdx.x * 2 * IB_VIS;   // visit list (id, offset)

	for (uint32_t job = blockIdx.x; job < njobs; job += gridDim.x) {
		DecJob& J = jobs[job];
		const uint32_t n = J.n;
		if (J.status != 0 || n == 0) continue;
		const uint8_t* L = bwt_all + (size_t)job * cap;
		uint8_t* txt = txt_all + (size_t)job * cap;

		// ---- LF mapping: T[pos] = i for the i-th occurrence ... stable counting sort of positions by byte
		if (tid < 256) run[tid] = J.cftab[tid];
		__syncthreads();
		radix_scatter<uint32_t>(n, run, wcnt,
			[&](uint32_t e) { return (e << 8) | (uint32_t)L[e]; },
			[&](uint32_t p) { return p & 255u; },
			[&](uint32_t pos, uint32_t p) { tt[pos] = p >> 8; });
		__syncthreads();
		for (uint32_t j = tid; j < n; j += BWT_NT) tt[j] = (tt[j] << 8) | (uint32_t)L[j];    // bzip2's tt layout
		__syncthreads();

		// ---- splitters: every K-th position, plus the start of the walk
		uint32_t K = 64;
		while ((n + K - 1) / K > IB_MAXS) K <<= 1;
		const uint32_t S = (n + K - 1) / K;
		const uint32_t p0 = tt[J.orig_ptr] >> 8;
		const bool p0_regular = (p0 % K) == 0;
		const uint32_t nspl = S + (p0_regular ? 0 : 1);
		auto splitter_id = [&](uint32_t p, uint32_t& id) -> bool {
			if (p % K == 0) { id = p / K; return true; }
			if (p == p0) { id = S; return true; }
			return false;
		};
		for (uint32_t q = tid; q < nspl; q += BWT_NT) {
			uint32_t p = q < S ? q * K : p0, len = 0, id = 0;
			do { p = tt[p] >> 8; len++; } while (!splitter_id(p, id) && len < n);
			s_next[q] = id; s_len[q] = len;
		}
		__syncthreads();
		// ---- chain the sublists from the start for exactly n steps (a periodic block laps its cycle several times)
		if (tid == 0) {
			uint32_t q = p0_regular ? p0 / K : S, off = 0, nv = 0;
			while (off < n && nv < (uint32_t)IB_VIS) { vis[2 * nv] = q; vis[2 * nv + 1] = off; nv++; off += s_len[q]; q = s_next[q]; }
			if (off < n) J.status = 2;
			s_nvis = nv;
		}
		__syncthreads();
		const uint32_t nvis = s_nvis;
		for (uint32_t i = tid; i < nvis; i += BWT_NT) {
			uint32_t q = vis[2 * i], off = vis[2 * i + 1];
			uint32_t p = q < S ? q * K : p0, len = s_len[q];
			for (uint32_t k = 0; k < len && off + k < n; k++) { uint32_t e = tt[p]; txt[off + k] = (uint8_t)e; p = e >> 8; }
		}
		__syncthreads();
	}
}

// =====================================================================================================
// k_unrle : one warp per block, lane 0 (v1)
// =====================================================================================================
constexpr int UR_NT = 128;
__global__ void __launch_bounds__(UR_NT)
k_unrle(const uint8_t* __restrict__ txt_all, uint32_t cap, DecJob* __restrict__ jobs, uint32_t njobs,
        uint16_t* __restrict__ sym, Geom g, const uint64_t* __restrict__ block_ids)
{
	__shared__ uint32_t crc_tab[256];
	for (uint32_t i = threadIdx.x; i < 256; i += UR_NT) crc_tab[i] = crc_table_entry(i);
	__syncthreads();
	uint32_t job = blockIdx.x * (UR_NT / 32) + warp_id();
	if (job >= njobs || lane_id() != 0) return;
	DecJob& J = jobs[job];
	if (J.status != 0) return;
	const uint8_t* txt = txt_all + (size_t)job * cap;
	const uint32_t n = J.n;
	uint32_t c0[5], ext[5];
	block_box(g, block_ids[job], c0, ext);
	const uint32_t row_bytes = ext[0] * 2;
	const uint64_t total = (uint64_t)row_bytes * ext[1] * ext[2] * ext[3] * ext[4];
	uint32_t y = 0, z = 0, c = 0, t = 0, xb = 0;           // position of the next output byte
	uint16_t* row = sym + (c0[0] + (uint64_t)c0[1] * g.stride[1] + (uint64_t)c0[2] * g.stride[2] + (uint64_t)c0[3] * g.stride[3] + (uint64_t)c0[4] * g.stride[4]);
	uint64_t produced = 0; uint32_t lo = 0, crc = 0xFFFFFFFFu;
	bool overflow = false;
	auto emit = [&](uint32_t b) {
		if (produced >= total) { overflow = true; return; }
		crc = (crc << 8) ^ crc_tab[(crc >> 24) ^ b];
		if (xb & 1) row[xb >> 1] = (uint16_t)(lo | (b << 8)); else lo = b;
		xb++; produced++;
		if (xb == row_bytes) {
			xb = 0;
			if (++y == ext[1]) { y = 0; if (++z == ext[2]) { z = 0; if (++c == ext[3]) { c = 0; ++t; } } }
			row = sym + (c0[0] + (uint64_t)(c0[1] + y) * g.stride[1] + (uint64_t)(c0[2] + z) * g.stride[2]
			             + (uint64_t)(c0[3] + c) * g.stride[3] + (uint64_t)(c0[4] + t) * g.stride[4]);
		}
	};
	int prev = -1; uint32_t cnt = 0;
	for (uint32_t i = 0; i < n; i++) {
		uint32_t ch = txt[i];
		if (cnt == 4) { for (uint32_t k = 0; k < ch; k++) emit((uint32_t)prev); cnt = 0; prev = -1; continue; }
		if ((int)ch == prev) cnt++; else { prev = (int)ch; cnt = 1; }
		emit(ch);
	}
	J.out_bytes = (uint32_t)produced;
	if (overflow || produced != total) J.status = 2;
	else if (~crc != J.stored_crc) J.status = 3;
}

// ------------------------------------------------------------------------------------------------ launchers
void launch_decode(const uint8_t* payload, const uint64_t* begin, const uint64_t* end, uint32_t njobs, DecJob* jobs,
                   uint8_t* bwt, uint32_t cap, uint8_t* sel, uint32_t selcap, cudaStream_t st)
{
	k_decode<<<(njobs + DEC_NW - 1) / DEC_NW, DEC_NT, 0, st>>>(payload, begin, end, njobs, jobs, bwt, cap, sel, selcap);
}
size_t inv_bwt_scratch_elems(int grid, uint32_t cap) { return (size_t)grid * cap + (size_t)grid * 2 * IB_VIS; }
void launch_inv_bwt(const uint8_t* bwt, uint32_t cap, DecJob* jobs, uint32_t njobs, uint32_t* tt_scratch, uint8_t* txt,
                    int grid, cudaStream_t st)
{
	k_inv_bwt<<<grid, BWT_NT, 0, st>>>(bwt, cap, jobs, njobs, tt_scratch, txt);
}
void launch_unrle(const uint8_t* txt, uint32_t cap, DecJob* jobs, uint32_t njobs, uint16_t* sym, const Geom& g,
                  const uint64_t* block_ids, cudaStream_t st)
{
	uint32_t per = UR_NT / 32;
	k_unrle<<<(njobs + per - 1) / per, UR_NT, 0, st>>>(txt, cap, jobs, njobs, sym, g, block_ids);
}

}  // namespace lfm
