// bz_encode.cu -- bzip2 block encoder stages around the BWT, for sm_100a.
//
// Replaces, bit-exactly, what BZ2_bzBuffToBuffCompress(dst,&len,src,n,level,0,30) does for one KLB block
// (src/klb_imageIO.cpp:217):
//   k_rle1      gather of the block from the symbol image (klb_imageIO.cpp:133-183) + initial run-length coding +
//               block CRC + inUse map                      (bzlib.c:216-354 ADD_CHAR_TO_BLOCK / add_pair_to_block)
//   k_mtf       move-to-front + RUNA/RUNB zero-run coding   (compress.c:121-232 generateMTFValues)
//   k_huff_pack coding-table selection, 4 refinement passes, exact bzip2 Huffman code lengths, canonical codes,
//               MSB-first bit packing and stream framing     (compress.c:240-600 sendMTFValues, huffman.c:63-166,
//               compress.c:603-676 BZ2_compressBlock header/trailer)
#include "lfm_device.cuh"

namespace lfm {

// =====================================================================================================
// k_rle1 : one warp per KLB block; lane 0 walks the block (v1: latency bound, parallel over blocks)
// =====================================================================================================
constexpr int RLE_NT = 128;

__global__ void __launch_bounds__(RLE_NT)
k_rle1(const uint16_t* __restrict__ sym, Geom g, uint64_t first_block, uint32_t njobs,
       uint8_t* __restrict__ txt_all, uint32_t cap, EncJob* __restrict__ jobs)
{
	__shared__ uint32_t crc_tab[256];
	for (uint32_t i = threadIdx.x; i < 256; i += RLE_NT) crc_tab[i] = crc_table_entry(i);
	__syncthreads();
	uint32_t job = blockIdx.x * (RLE_NT / 32) + warp_id();
	if (job >= njobs || lane_id() != 0) return;

	uint32_t c0[5], ext[5];
	block_box(g, first_block + job, c0, ext);
	uint8_t* out = txt_all + (size_t)job * cap;
	uint32_t nout = 0, acc = 0;
	uint64_t first8 = 0;               // the first 8 output bytes, replayed after the block as wrap-around
	uint32_t crc = 0xFFFFFFFFu;
	uint32_t in_use[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	int run_ch = -1; uint32_t run_len = 0;

	auto put = [&](uint32_t b) {
		if (nout < 8) first8 |= (uint64_t)b << (nout * 8);
		acc |= b << ((nout & 3) * 8);
		nout++;
		if ((nout & 3) == 0) { *reinterpret_cast<uint32_t*>(out + nout - 4) = acc; acc = 0; }
	};
	auto mark = [&](uint32_t b) {
		#pragma unroll
		for (int k = 0; k < 8; k++) if ((b >> 5) == (uint32_t)k) in_use[k] |= 1u << (b & 31);
	};
	auto flush_run = [&]() {
		if (run_ch < 0) return;
		mark((uint32_t)run_ch);
		uint32_t k = run_len < 4 ? run_len : 4;
		for (uint32_t i = 0; i < k; i++) put((uint32_t)run_ch);
		if (run_len >= 4) { put(run_len - 4); mark(run_len - 4); }
	};
	auto feed = [&](uint32_t b) {
		crc = (crc << 8) ^ crc_tab[(crc >> 24) ^ b];
		if ((int)b != run_ch || run_len == 255) { flush_run(); run_ch = (int)b; run_len = 1; }
		else run_len++;
	};

	uint32_t raw = 0;
	for (uint32_t t = 0; t < ext[4]; t++) for (uint32_t c = 0; c < ext[3]; c++)
	for (uint32_t z = 0; z < ext[2]; z++) for (uint32_t y = 0; y < ext[1]; y++) {
		const uint16_t* row = sym + (c0[0] + (uint64_t)(c0[1] + y) * g.stride[1] + (uint64_t)(c0[2] + z) * g.stride[2]
		                             + (uint64_t)(c0[3] + c) * g.stride[3] + (uint64_t)(c0[4] + t) * g.stride[4]);
		for (uint32_t x = 0; x < ext[0]; x++) {
			uint32_t v = __ldg(row + x);
			feed(v & 255u); feed(v >> 8);
		}
		raw += ext[0] * 2;
	}
	flush_run();
	uint32_t n = nout;
	// 8 wrap-around bytes after the block (the sort reads text[i + 0..7])
	if (n > 0) for (uint32_t k = 0; k < 8; k++) put((uint32_t)(first8 >> ((k % n) * 8)) & 255u);
	while (nout & 3) put(0);

	EncJob& J = jobs[job];
	J.raw_bytes = raw; J.n = n; J.crc = ~crc; J.status = 0; J.periodic = 0; J.orig_ptr = 0;
	uint32_t niu = 0;
	#pragma unroll
	for (int k = 0; k < 8; k++) { J.in_use[k] = in_use[k]; niu += __popc(in_use[k]); }
	J.n_in_use = niu;
}

// =====================================================================================================
// k_mtf : one warp per block.  The first 32 entries of the move-to-front list live one per lane,
// the tail in shared memory; the scan over the BWT output is sequential, parallel over blocks.
// =====================================================================================================
constexpr int MTF_NT = 128;
constexpr int MTF_NW = MTF_NT / 32;

__global__ void __launch_bounds__(MTF_NT)
k_mtf(const uint8_t* __restrict__ bwt_all, uint32_t cap, EncJob* __restrict__ jobs, uint32_t njobs,
      uint16_t* __restrict__ mtfv_all, uint32_t mcap)
{
	__shared__ uint8_t s_seq[MTF_NW][256];
	__shared__ uint8_t s_list[MTF_NW][256];
	const uint32_t lane = lane_id(), w = warp_id();
	const uint32_t job = blockIdx.x * MTF_NW + w;
	if (job >= njobs) return;
	const uint32_t n = jobs[job].n;
	const uint8_t* bwt = bwt_all + (size_t)job * cap;
	uint16_t* mtfv = mtfv_all + (size_t)job * mcap;

	// unseqToSeq (compress.c:105-116): rank of each used byte value
	uint32_t iu[8];
	#pragma unroll
	for (int k = 0; k < 8; k++) iu[k] = jobs[job].in_use[k];
	uint32_t n_in_use = 0;
	#pragma unroll
	for (int k = 0; k < 8; k++) {
		uint32_t c = k * 32 + lane;
		uint32_t below = n_in_use + __popc(iu[k] & ((1u << lane) - 1u));
		s_seq[w][c] = (uint8_t)below;
		s_list[w][c] = (uint8_t)c;
		n_in_use += __popc(iu[k]);
	}
	__syncwarp();
	const uint32_t EOB = n_in_use + 1;

	uint32_t y0 = lane;            // list positions 0..31
	uint32_t front = 0;
	uint32_t zpend = 0, wr = 0, buf = 0;
	auto emit = [&](uint32_t v) {
		if (lane == (wr & 31)) buf = v;
		wr++;
		if ((wr & 31) == 0) mtfv[wr - 32 + lane] = (uint16_t)buf;
	};
	auto flush_zeros = [&]() {
		if (zpend == 0) return;
		uint32_t z = zpend - 1;
		for (;;) { emit(z & 1u); if (z < 2) break; z = (z - 2) >> 1; }
		zpend = 0;
	};

	for (uint32_t b0 = 0; b0 < n; b0 += 32) {
		uint32_t mine = (b0 + lane < n) ? s_seq[w][bwt[b0 + lane]] : 0;
		uint32_t cnt = min(32u, n - b0);
		for (uint32_t t = 0; t < cnt; t++) {
			uint32_t c = __shfl_sync(0xffffffffu, mine, t);
			if (c == front) { zpend++; continue; }
			flush_zeros();
			uint32_t pos;
			uint32_t bal = __ballot_sync(0xffffffffu, y0 == c);
			uint32_t up = __shfl_up_sync(0xffffffffu, y0, 1);
			if (bal) {
				pos = __ffs(bal) - 1;
				if (lane == 0) y0 = c; else if (lane <= pos) y0 = up;
			} else {
				uint32_t carry = __shfl_sync(0xffffffffu, y0, 31);
				if (lane == 0) y0 = c; else y0 = up;
				pos = 0;
				for (uint32_t ch = 1; ch < 8; ch++) {
					uint32_t v = s_list[w][ch * 32 + lane];
					uint32_t hit = __ballot_sync(0xffffffffu, v == c);
					uint32_t nextc = __shfl_sync(0xffffffffu, v, 31);
					uint32_t upv = __shfl_up_sync(0xffffffffu, v, 1);
					if (lane == 0) upv = carry;
					uint32_t limit = hit ? (uint32_t)(__ffs(hit) - 1) : 31u;
					if (lane <= limit) s_list[w][ch * 32 + lane] = (uint8_t)upv;
					carry = nextc;
					if (hit) { pos = ch * 32 + limit; break; }
				}
				__syncwarp();
			}
			front = c;
			emit(pos + 1);
		}
	}
	flush_zeros();
	emit(EOB);
	if (wr & 31) { if (lane < (wr & 31)) mtfv[(wr & ~31u) + lane] = (uint16_t)buf; }
	if (lane == 0) { jobs[job].n_mtf = wr; jobs[job].n_in_use = n_in_use; }
}

// =====================================================================================================
// k_huff_pack : one CTA per block
// =====================================================================================================
constexpr int HP_NT = 256;
constexpr int HP_SYMS = 8;     // symbols per thread per packing tile

// exact restatement of BZ2_hbMakeCodeLengths (huffman.c:63-148): same heap, same tie rules, same rescale loop
__device__ void make_code_lengths(uint8_t* len, const uint32_t* freq, int alpha, int max_len,
                                  int32_t* heap, int32_t* weight, int32_t* parent)
{
	for (int i = 0; i < alpha; i++) weight[i + 1] = (int32_t)((freq[i] == 0 ? 1u : freq[i]) << 8);
	for (;;) {
		int n_nodes = alpha, n_heap = 0;
		heap[0] = 0; weight[0] = 0; parent[0] = -2;
		for (int i = 1; i <= alpha; i++) {
			parent[i] = -1;
			int z = ++n_heap, t = i;
			int32_t wt = weight[t];
			while (wt < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
			heap[z] = t;
		}
		while (n_heap > 1) {
			int pick[2];
			#pragma unroll
			for (int q = 0; q < 2; q++) {
				pick[q] = heap[1]; heap[1] = heap[n_heap--];
				int z = 1, t = heap[1];
				int32_t wt = weight[t];
				for (;;) {
					int y = z << 1;
					if (y > n_heap) break;
					if (y < n_heap && weight[heap[y + 1]] < weight[heap[y]]) y++;
					if (wt < weight[heap[y]]) break;
					heap[z] = heap[y]; z = y;
				}
				heap[z] = t;
			}
			n_nodes++;
			parent[pick[0]] = parent[pick[1]] = n_nodes;
			uint32_t w1 = (uint32_t)weight[pick[0]], w2 = (uint32_t)weight[pick[1]];
			uint32_t d1 = w1 & 0xffu, d2 = w2 & 0xffu;
			int32_t nw = (int32_t)(((w1 & 0xffffff00u) + (w2 & 0xffffff00u)) | (1u + (d1 > d2 ? d1 : d2)));
			weight[n_nodes] = nw;
			parent[n_nodes] = -1;
			int z = ++n_heap;
			while (nw < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
			heap[z] = n_nodes;
		}
		bool too_long = false;
		for (int i = 1; i <= alpha; i++) {
			int j = 0, k = i;
			while (parent[k] >= 0) { k = parent[k]; j++; }
			len[i - 1] = (uint8_t)j;
			if (j > max_len) too_long = true;
		}
		if (!too_long) break;
		for (int i = 1; i <= alpha; i++) { int j = weight[i] >> 8; j = 1 + (j / 2); weight[i] = j << 8; }
	}
}

// sequential MSB-first bit writer used by thread 0 for the block header (whole 32-bit words, big-endian)
struct HdrWriter {
	uint32_t* out; uint64_t acc; uint32_t live; uint32_t words;
	__device__ void put(uint32_t nbits, uint32_t v) {
		acc = (acc << nbits) | (uint64_t)v; live += nbits;
		if (live >= 32) {
			uint32_t wv = (uint32_t)(acc >> (live - 32));
			out[words++] = __byte_perm(wv, 0, 0x0123);
			live -= 32;
		}
	}
	__device__ uint32_t bits() const { return words * 32 + live; }
	// the < 32 leftover bits are OR-ed into the (zeroed) next word
	__device__ void finish() { if (live) { uint32_t wv = (uint32_t)(acc << (32 - live)); atomicOr(&out[words], __byte_perm(wv, 0, 0x0123)); } }
};

// OR `nbits` (<= 32) bits of v into the big-endian bit stream at bit position bitpos (words pre-zeroed)
__device__ __forceinline__ void or_bits(uint32_t* out, uint32_t bitpos, uint32_t nbits, uint32_t v)
{
	uint32_t wi = bitpos >> 5, off = bitpos & 31u;
	uint64_t x = (uint64_t)v << (64 - off - nbits);
	uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
	if (hi) atomicOr(&out[wi], __byte_perm(hi, 0, 0x0123));
	if (lo) atomicOr(&out[wi + 1], __byte_perm(lo, 0, 0x0123));
}

__global__ void __launch_bounds__(HP_NT)
k_huff_pack(const uint16_t* __restrict__ mtfv_all, uint32_t mcap, EncJob* __restrict__ jobs, uint32_t njobs,
            uint8_t* __restrict__ sel_all, uint32_t selcap, uint8_t* __restrict__ out_all, uint32_t ocap, int level)
{
	__shared__ uint32_t freq[kMaxAlpha + 6];
	__shared__ uint32_t rfreq[kGroups][kMaxAlpha + 2];
	__shared__ uint8_t  len[kGroups][kMaxAlpha + 2];
	__shared__ uint32_t code[kGroups][kMaxAlpha + 2];
	__shared__ int32_t  hheap[kGroups][kMaxAlpha + 2];
	__shared__ int32_t  hweight[kGroups][kMaxAlpha * 2];
	__shared__ int32_t  hparent[kGroups][kMaxAlpha * 2];
	__shared__ uint32_t red[64];
	uint32_t* win = reinterpret_cast<uint32_t*>(&hweight[0][0]);   // packing window; reused once the tables are final
	static_assert(sizeof(hweight) >= (HP_NT * HP_SYMS * 20 / 32 + 4) * 4, "window must fit");
	__shared__ uint32_t s_hdr_bits, s_ngroups;

	const uint32_t tid = threadIdx.x;
	const uint32_t job = blockIdx.x;
	if (job >= njobs) return;
	EncJob& J = jobs[job];
	const uint32_t n_mtf = J.n_mtf, n_in_use = J.n_in_use;
	const int alpha = (int)n_in_use + 2;
	const uint16_t* mtfv = mtfv_all + (size_t)job * mcap;
	uint8_t* selector = sel_all + (size_t)job * selcap;
	uint32_t* out = reinterpret_cast<uint32_t*>(out_all + (size_t)job * ocap);

	if (J.n == 0) {     // empty input: stream header + trailer only (compress.c:603-676 with nblock == 0)
		if (tid == 0) {
			out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 0;
			HdrWriter hw{ out, 0, 0, 0 };
			hw.put(8, 'B'); hw.put(8, 'Z'); hw.put(8, 'h'); hw.put(8, (uint32_t)('0' + level));
			hw.put(24, 0x177245); hw.put(24, 0x385090); hw.put(32, 0);
			hw.finish();
			J.total_bits = hw.bits(); J.out_bytes = (hw.bits() + 7) / 8; J.n_groups = 0; J.n_sel = 0;
		}
		return;
	}

	// ---- symbol frequencies
	for (uint32_t i = tid; i < kMaxAlpha + 6; i += HP_NT) freq[i] = 0;
	__syncthreads();
	for (uint32_t i = tid; i < n_mtf; i += HP_NT) atomicAdd(&freq[mtfv[i]], 1u);
	__syncthreads();

	// ---- number of tables + initial partition (compress.c:268-316)
	if (tid == 0) {
		int ng = n_mtf < 200 ? 2 : n_mtf < 600 ? 3 : n_mtf < 1200 ? 4 : n_mtf < 2400 ? 5 : 6;
		s_ngroups = ng;
		int n_part = ng, rem = (int)n_mtf, gs = 0;
		while (n_part > 0) {
			int target = rem / n_part, ge = gs - 1, acc = 0;
			while (acc < target && ge < alpha - 1) { ge++; acc += (int)freq[ge]; }
			if (ge > gs && n_part != ng && n_part != 1 && ((ng - n_part) % 2 == 1)) { acc -= (int)freq[ge]; ge--; }
			for (int v = 0; v < alpha; v++) len[n_part - 1][v] = (v >= gs && v <= ge) ? 0 : 15;
			n_part--; gs = ge + 1; rem -= acc;
		}
	}
	__syncthreads();
	const int ng = (int)s_ngroups;
	const uint32_t n_sel = (n_mtf + kGSize - 1) / kGSize;

	// ---- 4 refinement passes (compress.c:321-453)
	for (int iter = 0; iter < 4; iter++) {
		for (uint32_t i = tid; i < kGroups * (kMaxAlpha + 2); i += HP_NT) (&rfreq[0][0])[i] = 0;
		__syncthreads();
		for (uint32_t gi = tid; gi < n_sel; gi += HP_NT) {
			uint32_t gs = gi * kGSize, ge = min(gs + kGSize, n_mtf);
			uint32_t cost[kGroups] = { 0, 0, 0, 0, 0, 0 };
			for (uint32_t i = gs; i < ge; i++) {
				uint32_t s = mtfv[i];
				#pragma unroll
				for (int t = 0; t < kGroups; t++) if (t < ng) cost[t] += len[t][s];
			}
			int bt = 0; uint32_t bc = cost[0];
			#pragma unroll
			for (int t = 1; t < kGroups; t++) if (t < ng && cost[t] < bc) { bc = cost[t]; bt = t; }
			selector[gi] = (uint8_t)bt;
			for (uint32_t i = gs; i < ge; i++) atomicAdd(&rfreq[bt][mtfv[i]], 1u);
		}
		__syncthreads();
		if ((tid & 31) == 0 && (int)(tid >> 5) < ng) {
			int t = tid >> 5;
			make_code_lengths(len[t], rfreq[t], alpha, 17, hheap[t], hweight[t], hparent[t]);
		}
		__syncthreads();
	}

	// ---- canonical codes (huffman.c:152-166), one table per warp leader
	if ((tid & 31) == 0 && (int)(tid >> 5) < ng) {
		int t = tid >> 5, mn = 32, mx = 0;
		for (int i = 0; i < alpha; i++) { int l = len[t][i]; mx = l > mx ? l : mx; mn = l < mn ? l : mn; }
		uint32_t vec = 0;
		for (int l = mn; l <= mx; l++) {
			for (int i = 0; i < alpha; i++) if (len[t][i] == l) code[t][i] = vec++;
			vec <<= 1;
		}
	}
	__syncthreads();

	// ---- total size: header bits (computed by thread 0 while writing) + sum of code lengths
	uint32_t sym_bits_local = 0;
	for (uint32_t i = tid; i < (uint32_t)ng * (uint32_t)alpha; i += HP_NT) {
		uint32_t t = i / (uint32_t)alpha, v = i % (uint32_t)alpha;
		sym_bits_local += rfreq[t][v] * len[t][v];      // rfreq of the last pass == usage with the final selectors
	}
	uint32_t sym_bits; block_scan_add<HP_NT>(sym_bits_local, red, &sym_bits);

	// upper bound of the header so the output words can be cleared before anybody writes
	// (32 stream + 48+32+1+24 + 16+256 + 3+15 + n_sel*6 + ng*(5+alpha*(1+2*20)))
	{
		uint32_t hdr_max = 32 + 105 + 272 + 18 + n_sel * 6 + (uint32_t)ng * (5 + (uint32_t)alpha * 41);
		uint32_t words = (hdr_max + sym_bits + 80 + 31) / 32 + 2;
		if ((size_t)words * 4 > ocap) { if (tid == 0) { J.status = 2; J.out_bytes = 0; } return; }
		for (uint32_t i = tid; i < words; i += HP_NT) out[i] = 0;
	}
	__syncthreads();

	if (tid == 0) {
		HdrWriter hw{ out, 0, 0, 0 };
		hw.put(8, 'B'); hw.put(8, 'Z'); hw.put(8, 'h'); hw.put(8, (uint32_t)('0' + level));
		hw.put(24, 0x314159); hw.put(24, 0x265359);
		hw.put(32, J.crc);
		hw.put(1, 0);
		hw.put(24, J.orig_ptr);
		uint32_t iu[8]; uint32_t used16 = 0;
		for (int k = 0; k < 8; k++) { iu[k] = J.in_use[k]; }
		for (int i = 0; i < 16; i++) { uint32_t chunk = (iu[i >> 1] >> ((i & 1) * 16)) & 0xffffu; if (chunk) used16 |= 1u << (15 - i); }
		hw.put(16, used16);
		for (int i = 0; i < 16; i++) {
			uint32_t chunk = (iu[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
			if (chunk) hw.put(16, __brev(chunk) >> 16);     // bit j of the chunk is sent j-th (MSB first)
		}
		hw.put(3, (uint32_t)ng);
		hw.put(15, n_sel);
		{   // selectors, move-to-front coded, unary (compress.c:462-479, :530-535)
			uint8_t pos[kGroups];
			for (int i = 0; i < ng; i++) pos[i] = (uint8_t)i;
			for (uint32_t i = 0; i < n_sel; i++) {
				uint8_t s = selector[i];
				int j = 0; uint8_t tmp = pos[0];
				while (tmp != s) { j++; uint8_t t2 = pos[j]; pos[j] = tmp; tmp = t2; }
				pos[0] = tmp;
				hw.put((uint32_t)j + 1, ((1u << j) - 1u) << 1);
			}
		}
		for (int t = 0; t < ng; t++) {   // delta coded lengths (compress.c:537-548)
			int curr = len[t][0];
			hw.put(5, (uint32_t)curr);
			for (int i = 0; i < alpha; i++) {
				int l = len[t][i];
				while (curr < l) { hw.put(2, 2); curr++; }
				while (curr > l) { hw.put(2, 3); curr--; }
				hw.put(1, 0);
			}
		}
		s_hdr_bits = hw.bits();
		hw.finish();
	}
	__syncthreads();
	const uint32_t hdr_bits = s_hdr_bits;

	// ---- symbols: tiles of HP_NT*HP_SYMS, bit offsets by block scan, assembled in a shared window
	uint32_t bitpos = hdr_bits;
	constexpr uint32_t TILE = HP_NT * HP_SYMS;
	constexpr uint32_t WIN_WORDS = HP_NT * HP_SYMS * 20 / 32 + 4;
	for (uint32_t t0 = 0; t0 < n_mtf; t0 += TILE) {
		for (uint32_t i = tid; i < WIN_WORDS; i += HP_NT) win[i] = 0;
		uint32_t i0 = t0 + tid * HP_SYMS;
		uint32_t l[HP_SYMS], c[HP_SYMS], sum = 0;
		#pragma unroll
		for (int k = 0; k < HP_SYMS; k++) {
			uint32_t i = i0 + k;
			l[k] = 0; c[k] = 0;
			if (i < n_mtf) { uint32_t t = selector[i / kGSize], s = mtfv[i]; l[k] = len[t][s]; c[k] = code[t][s]; }
			sum += l[k];
		}
		uint32_t tot; uint32_t inc = block_scan_add<HP_NT>(sum, red, &tot);   // includes the barrier after zeroing win
		uint32_t rel = (bitpos & 31u) + inc - sum;
		#pragma unroll
		for (int k = 0; k < HP_SYMS; k++) {
			if (l[k]) {
				uint32_t wi = rel >> 5, off = rel & 31u;
				uint64_t v = (uint64_t)c[k] << (64 - off - l[k]);
				atomicOr(&win[wi], (uint32_t)(v >> 32));
				uint32_t lo = (uint32_t)v;
				if (lo) atomicOr(&win[wi + 1], lo);
				rel += l[k];
			}
		}
		__syncthreads();
		uint32_t nwords = ((bitpos & 31u) + tot + 31) / 32;
		uint32_t w0 = bitpos >> 5;
		for (uint32_t i = tid; i < nwords; i += HP_NT) {
			uint32_t wv = win[i];
			if (wv) atomicOr(&out[w0 + i], __byte_perm(wv, 0, 0x0123));
		}
		bitpos += tot;
		__syncthreads();
	}

	// ---- trailer: end-of-stream magic + combined CRC (= block CRC for a single block), pad to a byte
	if (tid == 0) {
		or_bits(out, bitpos, 24, 0x177245u);
		or_bits(out, bitpos + 24, 24, 0x385090u);
		or_bits(out, bitpos + 48, 32, J.crc);
		uint32_t total = bitpos + 80;
		J.total_bits = total; J.out_bytes = (total + 7) / 8; J.n_groups = (uint32_t)ng; J.n_sel = n_sel;
	}
}

// ------------------------------------------------------------------------------------------------ launchers
void launch_rle1(const uint16_t* sym, const Geom& g, uint64_t first_block, uint32_t njobs, uint8_t* txt, uint32_t cap,
                 EncJob* jobs, cudaStream_t st)
{
	uint32_t per = RLE_NT / 32;
	k_rle1<<<(njobs + per - 1) / per, RLE_NT, 0, st>>>(sym, g, first_block, njobs, txt, cap, jobs);
}
void launch_mtf(const uint8_t* bwt, uint32_t cap, EncJob* jobs, uint32_t njobs, uint16_t* mtfv, uint32_t mcap, cudaStream_t st)
{
	k_mtf<<<(njobs + MTF_NW - 1) / MTF_NW, MTF_NT, 0, st>>>(bwt, cap, jobs, njobs, mtfv, mcap);
}
void launch_huff_pack(const uint16_t* mtfv, uint32_t mcap, EncJob* jobs, uint32_t njobs, uint8_t* sel, uint32_t selcap,
                      uint8_t* out, uint32_t ocap, int level, cudaStream_t st)
{
	k_huff_pack<<<njobs, HP_NT, 0, st>>>(mtfv, mcap, jobs, njobs, sel, selcap, out, ocap, level);
}

}  // namespace lfm
