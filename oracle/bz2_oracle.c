/*
 * oracle/bz2_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not linked, imported or executed by the product.
 *
 * Plain sequential C restatement of the bzip2 1.0.6 block codec as the reference drives it:
 *   BZ2_bzBuffToBuffCompress(dst,&len,src,n, level, 0, 30)   (src/klb_imageIO.cpp:217)
 *   BZ2_bzBuffToBuffDecompress(dst,&n,src,len, 0, 0)          (src/klb_imageIO.cpp:627, :1034)
 * The algorithm lives in the reference tree under src/external/bzip2-1.0.6/; every stage below cites the
 * file:line it follows.  The stages are exposed one by one (bz2o_trace) so the CUDA kernels can be checked
 * stage by stage, not only on the final stream.
 *
 * Parity pin: tests/test_oracle_bz2.py checks this file byte-for-byte against
 *   - the reference's own known-answer vectors sample{1,2,3}.ref <-> .bz2 (levels 1/2/3),
 *   - oracle/_ref/libbz2ref.so (the vendored bzip2 compiled as is) and Python's bz2 module on image-like,
 *     random, constant, periodic, empty and multi-block inputs.
 * Known deviation: for EXACTLY periodic blocks bzip2's origPtr depends on where
 * its quicksort happens to leave rotation 0 inside the group of identical rotations; this restatement (and the
 * CUDA engine) place rotation 0 LAST in that group, which is what bzip2 does for the periods that occur in
 * practice (period <= 3, e.g. constant uint16 pixels).  Everything else is byte-identical.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BZO_MAX_ALPHA 258
#define BZO_GROUPS    6
#define BZO_GSIZE     50
#define BZO_ITERS     4
#define BZO_MAX_SEL   (2 + (900000 / BZO_GSIZE))

/* ------------------------------------------------------------------ CRC (crctable.c; bzlib_private.h:157-171) */
static uint32_t crc_tab[256];
static int crc_ready = 0;
static void crc_init(void)
{
	if (crc_ready) return;
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t c = i << 24;
		for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
		crc_tab[i] = c;
	}
	crc_ready = 1;
}
uint32_t bz2o_crc32(const uint8_t* p, size_t n)
{
	crc_init();
	uint32_t c = 0xFFFFFFFFu;
	for (size_t i = 0; i < n; i++) c = (c << 8) ^ crc_tab[(c >> 24) ^ p[i]];
	return ~c;
}

/* ------------------------------------------------------------------ MSB-first bit writer (compress.c:74-99) */
typedef struct { uint8_t* out; size_t cap, pos; uint64_t acc; int live; int overflow; } bitw;
static void bw_put(bitw* w, int nbits, uint32_t v)
{
	w->acc = (w->acc << nbits) | (uint64_t)(v & ((nbits == 32) ? 0xFFFFFFFFu : ((1u << nbits) - 1u)));
	w->live += nbits;
	while (w->live >= 8) {
		if (w->pos < w->cap) w->out[w->pos] = (uint8_t)(w->acc >> (w->live - 8)); else w->overflow = 1;
		w->pos++; w->live -= 8;
	}
}
static void bw_flush(bitw* w) { if (w->live > 0) bw_put(w, 8 - w->live, 0); }

/* ------------------------------------------------------------------ stage trace (one bzip2 block) */
typedef struct {
	int32_t  nblock;        /* bytes after RLE1 */
	uint32_t block_crc;
	int32_t  orig_ptr;
	int32_t  n_in_use;
	int32_t  n_mtf;
	int32_t  n_groups;
	int32_t  n_selectors;
	int32_t  periodic;      /* 1 if the rotation sort ended with ties */
	uint8_t* rle1;          /* [nblock]  post-RLE1 bytes                         */
	uint8_t* bwt;           /* [nblock]  last column                              */
	uint16_t* mtfv;         /* [n_mtf]                                            */
	uint8_t* selector;      /* [n_selectors]                                      */
	uint8_t  len[BZO_GROUPS][BZO_MAX_ALPHA];
	int64_t  bit_start, bit_end; /* position of this block inside the stream     */
} bz2o_block_trace;

typedef struct { int n_blocks; int cap_blocks; bz2o_block_trace* blk; } bz2o_trace;

void bz2o_trace_free(bz2o_trace* t)
{
	if (!t) return;
	for (int i = 0; i < t->n_blocks; i++) { free(t->blk[i].rle1); free(t->blk[i].bwt); free(t->blk[i].mtfv); free(t->blk[i].selector); }
	free(t->blk); t->blk = NULL; t->n_blocks = t->cap_blocks = 0;
}

/* ------------------------------------------------------------------ rotation sort (blocksort.c:1031-1090 contract)
 * Contract of BZ2_blockSort: ptr[] = cyclic rotations of block[0..n) in ascending order, origPtr = rank of
 * rotation 0.  Implemented here as prefix doubling with a radix pass per round (O(n log n), immune to
 * repetitive data); how the order is produced does not matter, only the order. */
static void rot_sort(const uint8_t* b, int32_t n, int32_t* sa, int32_t* periodic)
{
	int32_t* rank = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	int32_t* tmp  = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	int32_t* nr   = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	int32_t* cnt  = (int32_t*)calloc((size_t)(n > 256 ? n : 256) + 1, sizeof(int32_t));
	*periodic = 0;
	/* h = 1: counting sort on the first byte; rank = first index of the bucket */
	for (int32_t i = 0; i < n; i++) cnt[b[i] + 1]++;
	for (int i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
	{
		int32_t start[257]; memcpy(start, cnt, sizeof(start));
		for (int32_t i = 0; i < n; i++) sa[cnt[b[i]]++] = i;
		for (int32_t i = 0; i < n; i++) rank[i] = start[b[i]];
	}
	for (int64_t h = 1; ; h <<= 1) {
		/* all ranks distinct? */
		int done = 1;
		for (int32_t j = 1; j < n; j++) if (rank[sa[j]] == rank[sa[j - 1]]) { done = 0; break; }
		if (done) break;
		if (h >= n) { *periodic = 1; break; }
		/* sort by (rank[i], rank[i+h]) : LSD = counting sort on second key, then stable on first */
		memset(cnt, 0, sizeof(int32_t) * ((size_t)n + 1));
		for (int32_t i = 0; i < n; i++) cnt[rank[(i + h) % n] + 1]++;
		for (int32_t i = 0; i < n; i++) cnt[i + 1] += cnt[i];
		for (int32_t i = 0; i < n; i++) tmp[cnt[rank[(i + h) % n]]++] = i;
		memset(cnt, 0, sizeof(int32_t) * ((size_t)n + 1));
		for (int32_t i = 0; i < n; i++) cnt[rank[i] + 1]++;
		for (int32_t i = 0; i < n; i++) cnt[i + 1] += cnt[i];
		for (int32_t j = 0; j < n; j++) { int32_t i = tmp[j]; sa[cnt[rank[i]]++] = i; }
		nr[sa[0]] = 0;
		for (int32_t j = 1; j < n; j++) {
			int32_t a = sa[j - 1], c = sa[j];
			int same = rank[a] == rank[c] && rank[(a + h) % n] == rank[(c + h) % n];
			nr[c] = same ? nr[a] : j;
		}
		memcpy(rank, nr, sizeof(int32_t) * (size_t)n);
	}
	if (*periodic) {
		/* identical rotations: order them by descending start index so that rotation 0 is LAST of its group */
		int32_t j = 0;
		while (j < n) {
			int32_t e = j + 1;
			while (e < n && rank[sa[e]] == rank[sa[j]]) e++;
			for (int32_t a = j; a < e; a++) for (int32_t c = a + 1; c < e; c++) if (sa[c] > sa[a]) { int32_t t = sa[a]; sa[a] = sa[c]; sa[c] = t; }
			j = e;
		}
	}
	free(rank); free(tmp); free(nr); free(cnt);
}

/* ------------------------------------------------------------------ Huffman code lengths (huffman.c:63-148) */
static void make_code_lengths(uint8_t* len, const int32_t* freq, int alpha, int max_len)
{
	int32_t heap[BZO_MAX_ALPHA + 2], weight[BZO_MAX_ALPHA * 2], parent[BZO_MAX_ALPHA * 2];
	for (int i = 0; i < alpha; i++) weight[i + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
	for (;;) {
		int n_nodes = alpha, n_heap = 0;
		heap[0] = 0; weight[0] = 0; parent[0] = -2;
		for (int i = 1; i <= alpha; i++) {
			parent[i] = -1;
			int z = ++n_heap, t = i;
			while (weight[t] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
			heap[z] = t;
		}
		while (n_heap > 1) {
			int pick[2];
			for (int q = 0; q < 2; q++) {
				pick[q] = heap[1]; heap[1] = heap[n_heap--];
				int z = 1, t = heap[1];
				for (;;) {
					int y = z << 1;
					if (y > n_heap) break;
					if (y < n_heap && weight[heap[y + 1]] < weight[heap[y]]) y++;
					if (weight[t] < weight[heap[y]]) break;
					heap[z] = heap[y]; z = y;
				}
				heap[z] = t;
			}
			n_nodes++;
			parent[pick[0]] = parent[pick[1]] = n_nodes;
			{
				uint32_t w1 = (uint32_t)weight[pick[0]], w2 = (uint32_t)weight[pick[1]];
				uint32_t d1 = w1 & 0xff, d2 = w2 & 0xff;
				weight[n_nodes] = (int32_t)(((w1 & 0xffffff00u) + (w2 & 0xffffff00u)) | (1 + (d1 > d2 ? d1 : d2)));
			}
			parent[n_nodes] = -1;
			int z = ++n_heap, t = n_nodes;
			while (weight[t] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
			heap[z] = t;
		}
		int too_long = 0;
		for (int i = 1; i <= alpha; i++) {
			int j = 0, k = i;
			while (parent[k] >= 0) { k = parent[k]; j++; }
			len[i - 1] = (uint8_t)j;
			if (j > max_len) too_long = 1;
		}
		if (!too_long) break;
		for (int i = 1; i <= alpha; i++) { int j = weight[i] >> 8; j = 1 + (j / 2); weight[i] = j << 8; }
	}
}

/* canonical codes by (length, symbol) (huffman.c:152-166) */
static void assign_codes(int32_t* code, const uint8_t* len, int min_len, int max_len, int alpha)
{
	int vec = 0;
	for (int n = min_len; n <= max_len; n++) {
		for (int i = 0; i < alpha; i++) if (len[i] == n) code[i] = vec++;
		vec <<= 1;
	}
}

/* ------------------------------------------------------------------ one block: BWT .. bits (compress.c:603-676) */
#include "bz2_randtable.h"
static const int32_t bz2o_rnums[512] = BZ_RNUMS_INIT;
/* decompress.c BZ_RAND_INIT_MASK / BZ_RAND_UPD_MASK / BZ_RAND_MASK applied to a whole block: byte j ^= 1 where the mask is set */
static void randomise_block(uint8_t* blk, int32_t nblock)
{
	int32_t n_to_go = 0, t_pos = 0;
	for (int32_t j = 0; j < nblock; j++) {
		if (n_to_go == 0) { n_to_go = bz2o_rnums[t_pos]; t_pos++; if (t_pos == 512) t_pos = 0; }
		n_to_go--;
		if (n_to_go == 1) blk[j] ^= 1;
	}
}

static int g_force_randomised = 0;      /* test hook: write blocks the way bzip2 <= 0.9.0 did after a failed sort (randomised bit set) */
void bz2o_set_randomised(int on) { g_force_randomised = on; }

static void compress_block(bitw* w, const uint8_t* blk_in, int32_t nblock, const uint8_t* in_use_in, uint32_t block_crc,
                           bz2o_block_trace* tr)
{
	const int randomised = g_force_randomised;
	uint8_t* blk_r = NULL; uint8_t in_use_r[256];
	const uint8_t* blk = blk_in; const uint8_t* in_use = in_use_in;
	if (randomised) {                   /* the block is XORed with the mask, then sorted and coded; the CRC stays that of the input */
		blk_r = (uint8_t*)malloc((size_t)nblock);
		memcpy(blk_r, blk_in, (size_t)nblock);
		randomise_block(blk_r, nblock);
		memset(in_use_r, 0, 256);
		for (int32_t i = 0; i < nblock; i++) in_use_r[blk_r[i]] = 1;
		blk = blk_r; in_use = in_use_r;
	}
	int32_t* sa = (int32_t*)malloc(sizeof(int32_t) * (size_t)nblock);
	int32_t periodic = 0;
	rot_sort(blk, nblock, sa, &periodic);
	int32_t orig_ptr = -1;
	for (int32_t j = 0; j < nblock; j++) if (sa[j] == 0) { orig_ptr = j; break; }

	/* symbol map (compress.c:105-116) */
	uint8_t to_seq[256]; int n_in_use = 0;
	for (int i = 0; i < 256; i++) if (in_use[i]) to_seq[i] = (uint8_t)n_in_use++;
	int eob = n_in_use + 1, alpha = n_in_use + 2;

	/* MTF + zero-run coding (compress.c:121-232) */
	uint16_t* mtfv = (uint16_t*)malloc(sizeof(uint16_t) * ((size_t)nblock + 2));
	int32_t mtf_freq[BZO_MAX_ALPHA]; memset(mtf_freq, 0, sizeof(mtf_freq));
	uint8_t order[256];
	for (int i = 0; i < n_in_use; i++) order[i] = (uint8_t)i;
	int32_t wr = 0, zpend = 0;
	uint8_t* bwt = tr ? (uint8_t*)malloc((size_t)nblock) : NULL;
	for (int32_t i = 0; i <= nblock; i++) {
		int flush = (i == nblock), sym = 0, pos = 0;
		if (!flush) {
			int32_t j = sa[i] - 1; if (j < 0) j += nblock;
			if (bwt) bwt[i] = blk[j];
			sym = to_seq[blk[j]];
			if (order[0] == sym) { zpend++; continue; }
		}
		if (zpend > 0) {          /* bijective base 2 with RUNA=0 / RUNB=1 */
			int32_t z = zpend - 1;
			for (;;) {
				int s = (z & 1) ? 1 : 0;
				mtfv[wr++] = (uint16_t)s; mtf_freq[s]++;
				if (z < 2) break;
				z = (z - 2) / 2;
			}
			zpend = 0;
		}
		if (flush) break;
		while (order[pos] != sym) pos++;
		memmove(order + 1, order, (size_t)pos);
		order[0] = (uint8_t)sym;
		mtfv[wr++] = (uint16_t)(pos + 1); mtf_freq[pos + 1]++;
	}
	mtfv[wr++] = (uint16_t)eob; mtf_freq[eob]++;
	int32_t n_mtf = wr;

	/* coding tables (compress.c:240-453) */
	uint8_t* sel_buf = (uint8_t*)malloc(BZO_MAX_SEL); uint8_t* sel_mtf = (uint8_t*)malloc(BZO_MAX_SEL);
	uint8_t len[BZO_GROUPS][BZO_MAX_ALPHA];
	int32_t rfreq[BZO_GROUPS][BZO_MAX_ALPHA], code[BZO_GROUPS][BZO_MAX_ALPHA];
	memset(len, 15, sizeof(len));
	int n_groups = n_mtf < 200 ? 2 : n_mtf < 600 ? 3 : n_mtf < 1200 ? 4 : n_mtf < 2400 ? 5 : 6;
	{
		int n_part = n_groups, rem = n_mtf, gs = 0;
		while (n_part > 0) {
			int target = rem / n_part, ge = gs - 1, acc = 0;
			while (acc < target && ge < alpha - 1) { ge++; acc += mtf_freq[ge]; }
			if (ge > gs && n_part != n_groups && n_part != 1 && ((n_groups - n_part) % 2 == 1)) { acc -= mtf_freq[ge]; ge--; }
			for (int v = 0; v < alpha; v++) len[n_part - 1][v] = (v >= gs && v <= ge) ? 0 : 15;
			n_part--; gs = ge + 1; rem -= acc;
		}
	}
	int n_sel = 0;
	for (int iter = 0; iter < BZO_ITERS; iter++) {
		memset(rfreq, 0, sizeof(rfreq));
		n_sel = 0;
		for (int32_t gs = 0; gs < n_mtf; gs += BZO_GSIZE) {
			int32_t ge = gs + BZO_GSIZE - 1; if (ge >= n_mtf) ge = n_mtf - 1;
			uint16_t cost[BZO_GROUPS] = { 0, 0, 0, 0, 0, 0 };
			for (int32_t i = gs; i <= ge; i++) for (int t = 0; t < n_groups; t++) cost[t] = (uint16_t)(cost[t] + len[t][mtfv[i]]);
			int bt = -1; int32_t bc = 999999999;
			for (int t = 0; t < n_groups; t++) if (cost[t] < bc) { bc = cost[t]; bt = t; }
			sel_buf[n_sel++] = (uint8_t)bt;
			for (int32_t i = gs; i <= ge; i++) rfreq[bt][mtfv[i]]++;
		}
		for (int t = 0; t < n_groups; t++) make_code_lengths(len[t], rfreq[t], alpha, 17);
	}
	/* MTF of the selectors (compress.c:462-479) */
	{
		uint8_t pos[BZO_GROUPS];
		for (int i = 0; i < n_groups; i++) pos[i] = (uint8_t)i;
		for (int i = 0; i < n_sel; i++) {
			int j = 0; while (pos[j] != sel_buf[i]) j++;
			memmove(pos + 1, pos, (size_t)j); pos[0] = sel_buf[i];
			sel_mtf[i] = (uint8_t)j;
		}
	}
	for (int t = 0; t < n_groups; t++) {
		int mn = 32, mx = 0;
		for (int i = 0; i < alpha; i++) { if (len[t][i] > mx) mx = len[t][i]; if (len[t][i] < mn) mn = len[t][i]; }
		assign_codes(code[t], len[t], mn, mx, alpha);
	}

	/* emission (compress.c:620-650 header, :482-600 tables + data) */
	int64_t bit_start = (int64_t)w->pos * 8 + w->live;
	bw_put(w, 8, 0x31); bw_put(w, 8, 0x41); bw_put(w, 8, 0x59); bw_put(w, 8, 0x26); bw_put(w, 8, 0x53); bw_put(w, 8, 0x59);
	bw_put(w, 32, block_crc);
	bw_put(w, 1, (uint32_t)randomised);
	bw_put(w, 24, (uint32_t)orig_ptr);
	{
		int in_use16[16];
		for (int i = 0; i < 16; i++) { in_use16[i] = 0; for (int j = 0; j < 16; j++) if (in_use[i * 16 + j]) in_use16[i] = 1; }
		for (int i = 0; i < 16; i++) bw_put(w, 1, (uint32_t)in_use16[i]);
		for (int i = 0; i < 16; i++) if (in_use16[i]) for (int j = 0; j < 16; j++) bw_put(w, 1, in_use[i * 16 + j] ? 1u : 0u);
	}
	bw_put(w, 3, (uint32_t)n_groups);
	bw_put(w, 15, (uint32_t)n_sel);
	for (int i = 0; i < n_sel; i++) { for (int j = 0; j < sel_mtf[i]; j++) bw_put(w, 1, 1); bw_put(w, 1, 0); }
	for (int t = 0; t < n_groups; t++) {
		int curr = len[t][0];
		bw_put(w, 5, (uint32_t)curr);
		for (int i = 0; i < alpha; i++) {
			while (curr < len[t][i]) { bw_put(w, 2, 2); curr++; }
			while (curr > len[t][i]) { bw_put(w, 2, 3); curr--; }
			bw_put(w, 1, 0);
		}
	}
	{
		int sel = 0;
		for (int32_t gs = 0; gs < n_mtf; gs += BZO_GSIZE, sel++) {
			int32_t ge = gs + BZO_GSIZE - 1; if (ge >= n_mtf) ge = n_mtf - 1;
			int t = sel_buf[sel];
			for (int32_t i = gs; i <= ge; i++) bw_put(w, len[t][mtfv[i]], (uint32_t)code[t][mtfv[i]]);
		}
	}
	if (tr) {
		tr->nblock = nblock; tr->block_crc = block_crc; tr->orig_ptr = orig_ptr; tr->n_in_use = n_in_use;
		tr->n_mtf = n_mtf; tr->n_groups = n_groups; tr->n_selectors = n_sel; tr->periodic = periodic;
		tr->rle1 = (uint8_t*)malloc((size_t)nblock); memcpy(tr->rle1, blk, (size_t)nblock);
		tr->bwt = bwt;
		tr->mtfv = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)n_mtf); memcpy(tr->mtfv, mtfv, sizeof(uint16_t) * (size_t)n_mtf);
		tr->selector = (uint8_t*)malloc((size_t)n_sel); memcpy(tr->selector, sel_buf, (size_t)n_sel);
		memcpy(tr->len, len, sizeof(len));
		tr->bit_start = bit_start; tr->bit_end = (int64_t)w->pos * 8 + w->live;
	}
	free(sa); free(mtfv); free(sel_buf); free(sel_mtf); free(blk_r);
}

/* ------------------------------------------------------------------ whole stream
 * RLE1 + block splitting: bzlib.c:216-354 (add_pair_to_block, ADD_CHAR_TO_BLOCK, copy_input_until_stop),
 * :386-418 (handle_compress: a full block is closed WITHOUT flushing the pending run), stream framing
 * compress.c:603-676.  Returns the stream size (even when it does not fit cap: check against cap). */
size_t bz2o_compress(const uint8_t* in, size_t n, int level, uint8_t* out, size_t cap, bz2o_trace* trace)
{
	crc_init();
	if (level < 1) level = 1;
	if (level > 9) level = 9;
	int32_t nblock_max = 100000 * level - 19;
	uint8_t* blk = (uint8_t*)malloc((size_t)100000 * level + 64);
	bitw w; memset(&w, 0, sizeof(w)); w.out = out; w.cap = cap;
	bw_put(&w, 8, 'B'); bw_put(&w, 8, 'Z'); bw_put(&w, 8, 'h'); bw_put(&w, 8, (uint32_t)('0' + level));
	uint32_t combined = 0;
	if (trace) { trace->n_blocks = 0; trace->cap_blocks = 0; trace->blk = NULL; }

	int run_ch = 256, run_len = 0;     /* pending run (state_in_ch / state_in_len) */
	size_t ip = 0;
	for (;;) {
		int32_t nblock = 0; uint32_t crc = 0xFFFFFFFFu; uint8_t in_use[256]; memset(in_use, 0, 256);
		int last = 0;
		for (;;) {
			if (ip >= n) { last = 1; break; }                 /* input exhausted -> flush + last block */
			if (nblock >= nblock_max) break;                  /* block full, pending run carries over  */
			int ch = in[ip++];
			if (ch != run_ch || run_len == 255) {
				if (run_ch < 256) {                           /* add_pair_to_block */
					for (int i = 0; i < run_len; i++) crc = (crc << 8) ^ crc_tab[(crc >> 24) ^ (uint32_t)run_ch];
					in_use[run_ch] = 1;
					int k = run_len < 4 ? run_len : 4;
					for (int i = 0; i < k; i++) blk[nblock++] = (uint8_t)run_ch;
					if (run_len >= 4) { in_use[run_len - 4] = 1; blk[nblock++] = (uint8_t)(run_len - 4); }
				}
				run_ch = ch; run_len = 1;
			} else run_len++;
		}
		if (last && run_ch < 256) {                           /* flush_RL (bzlib.c:304-310) */
			for (int i = 0; i < run_len; i++) crc = (crc << 8) ^ crc_tab[(crc >> 24) ^ (uint32_t)run_ch];
			in_use[run_ch] = 1;
			int k = run_len < 4 ? run_len : 4;
			for (int i = 0; i < k; i++) blk[nblock++] = (uint8_t)run_ch;
			if (run_len >= 4) { in_use[run_len - 4] = 1; blk[nblock++] = (uint8_t)(run_len - 4); }
			run_ch = 256; run_len = 0;
		}
		if (nblock > 0) {
			crc = ~crc;
			combined = (combined << 1) | (combined >> 31);
			combined ^= crc;
			bz2o_block_trace* tr = NULL;
			if (trace) {
				if (trace->n_blocks == trace->cap_blocks) {
					trace->cap_blocks = trace->cap_blocks ? trace->cap_blocks * 2 : 4;
					trace->blk = (bz2o_block_trace*)realloc(trace->blk, sizeof(bz2o_block_trace) * (size_t)trace->cap_blocks);
				}
				tr = &trace->blk[trace->n_blocks++]; memset(tr, 0, sizeof(*tr));
			}
			compress_block(&w, blk, nblock, in_use, crc, tr);
		}
		if (last) break;
	}
	bw_put(&w, 8, 0x17); bw_put(&w, 8, 0x72); bw_put(&w, 8, 0x45); bw_put(&w, 8, 0x38); bw_put(&w, 8, 0x50); bw_put(&w, 8, 0x90);
	bw_put(&w, 32, combined);
	bw_flush(&w);
	free(blk);
	return w.pos;
}

/* accessors so ctypes callers need not mirror the struct layout */
int bz2o_trace_nblocks(const bz2o_trace* t) { return t->n_blocks; }
const bz2o_block_trace* bz2o_trace_block(const bz2o_trace* t, int i) { return &t->blk[i]; }
int32_t bz2o_blk_i32(const bz2o_block_trace* b, int what)
{
	switch (what) { case 0: return b->nblock; case 1: return (int32_t)b->block_crc; case 2: return b->orig_ptr; case 3: return b->n_in_use;
	case 4: return b->n_mtf; case 5: return b->n_groups; case 6: return b->n_selectors; case 7: return b->periodic; default: return -1; }
}
const void* bz2o_blk_ptr(const bz2o_block_trace* b, int what)
{
	switch (what) { case 0: return b->rle1; case 1: return b->bwt; case 2: return b->mtfv; case 3: return b->selector; case 4: return b->len; default: return NULL; }
}
bz2o_trace* bz2o_trace_new(void) { return (bz2o_trace*)calloc(1, sizeof(bz2o_trace)); }
void bz2o_trace_delete(bz2o_trace* t) { bz2o_trace_free(t); free(t); }

/* ------------------------------------------------------------------ decoder (decompress.c:106-646, bzlib.c:561-728)
 * returns 0 ok, <0 error: -1 bad magic, -2 truncated/corrupt, -3 CRC mismatch, -4 output too small */
typedef struct { const uint8_t* p; size_t n; size_t bit; int err; } bitr;
static uint32_t br_get(bitr* r, int nbits)
{
	uint32_t v = 0;
	for (int i = 0; i < nbits; i++) {
		size_t by = r->bit >> 3;
		if (by >= r->n) { r->err = 1; return 0; }
		v = (v << 1) | ((r->p[by] >> (7 - (r->bit & 7))) & 1u);
		r->bit++;
	}
	return v;
}

int bz2o_decompress(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len)
{
	crc_init();
	bitr r = { in, n, 0, 0 };
	if (br_get(&r, 8) != 'B' || br_get(&r, 8) != 'Z' || br_get(&r, 8) != 'h') return -1;
	int level = (int)br_get(&r, 8) - '0';
	if (level < 1 || level > 9) return -1;
	int32_t max_block = 100000 * level;
	uint32_t* tt = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)max_block);
	uint8_t* selector = (uint8_t*)malloc(BZO_MAX_SEL);
	size_t op = 0; uint32_t combined = 0; int rc = 0;
	for (;;) {
		uint32_t m1 = br_get(&r, 24), m2 = br_get(&r, 24);
		if (r.err) { rc = -2; break; }
		if (m1 == 0x177245 && m2 == 0x385090) {
			uint32_t stored = br_get(&r, 32);
			if (r.err) rc = -2; else if (stored != combined) rc = -3;
			break;
		}
		if (m1 != 0x314159 || m2 != 0x265359) { rc = -2; break; }
		uint32_t stored_crc = br_get(&r, 32);
		int randomised = (int)br_get(&r, 1);
		int32_t orig_ptr = (int32_t)br_get(&r, 24);
		uint8_t seq_to_unseq[256]; int n_in_use = 0;
		{
			int used16[16];
			for (int i = 0; i < 16; i++) used16[i] = (int)br_get(&r, 1);
			for (int i = 0; i < 16; i++) if (used16[i]) for (int j = 0; j < 16; j++) if (br_get(&r, 1)) seq_to_unseq[n_in_use++] = (uint8_t)(i * 16 + j);
		}
		if (n_in_use == 0 || r.err) { rc = -2; break; }
		int alpha = n_in_use + 2;
		int n_groups = (int)br_get(&r, 3);
		int n_sel = (int)br_get(&r, 15);
		if (n_groups < 2 || n_groups > 6 || n_sel < 1 || n_sel > BZO_MAX_SEL) { rc = -2; break; }
		{
			uint8_t pos[BZO_GROUPS];
			for (int i = 0; i < n_groups; i++) pos[i] = (uint8_t)i;
			for (int i = 0; i < n_sel && !r.err; i++) {
				int j = 0;
				while (br_get(&r, 1)) { j++; if (j >= n_groups) { r.err = 1; break; } }
				if (r.err) break;
				uint8_t t = pos[j]; memmove(pos + 1, pos, (size_t)j); pos[0] = t; selector[i] = t;
			}
		}
		if (r.err) { rc = -2; break; }
		uint8_t len[BZO_GROUPS][BZO_MAX_ALPHA];
		int32_t limit[BZO_GROUPS][24], base[BZO_GROUPS][24], perm[BZO_GROUPS][BZO_MAX_ALPHA]; int min_len[BZO_GROUPS];
		for (int t = 0; t < n_groups && !r.err; t++) {
			int curr = (int)br_get(&r, 5);
			for (int i = 0; i < alpha; i++) {
				for (;;) {
					if (curr < 1 || curr > 20) { r.err = 1; break; }
					if (!br_get(&r, 1)) break;
					if (br_get(&r, 1)) curr--; else curr++;
				}
				if (r.err) break;
				len[t][i] = (uint8_t)curr;
			}
		}
		if (r.err) { rc = -2; break; }
		for (int t = 0; t < n_groups; t++) {           /* BZ2_hbCreateDecodeTables huffman.c:170-205 */
			int mn = 32, mx = 0;
			for (int i = 0; i < alpha; i++) { if (len[t][i] > mx) mx = len[t][i]; if (len[t][i] < mn) mn = len[t][i]; }
			int pp = 0;
			for (int i = mn; i <= mx; i++) for (int j = 0; j < alpha; j++) if (len[t][j] == i) perm[t][pp++] = j;
			for (int i = 0; i < 23; i++) base[t][i] = 0;
			for (int i = 0; i < alpha; i++) base[t][len[t][i] + 1]++;
			for (int i = 1; i < 23; i++) base[t][i] += base[t][i - 1];
			for (int i = 0; i < 23; i++) limit[t][i] = 0;
			int vec = 0;
			for (int i = mn; i <= mx; i++) { vec += (base[t][i + 1] - base[t][i]); limit[t][i] = vec - 1; vec <<= 1; }
			for (int i = mn + 1; i <= mx; i++) base[t][i] = ((limit[t][i - 1] + 1) << 1) - base[t][i];
			min_len[t] = mn;
		}
		/* MTF / run decoding (decompress.c:349-487) */
		int eob = n_in_use + 1;
		int32_t unzftab[256]; memset(unzftab, 0, sizeof(unzftab));
		uint8_t order[256]; for (int i = 0; i < 256; i++) order[i] = (uint8_t)i;
		int32_t nblock = 0; int grp = -1, left = 0, t = 0; int bad = 0;
		int32_t run = 0, run_w = 1; int in_run = 0;
		for (;;) {
			if (left == 0) { grp++; if (grp >= n_sel) { bad = 1; break; } left = BZO_GSIZE; t = selector[grp]; }
			left--;
			int zn = min_len[t]; int32_t zvec = (int32_t)br_get(&r, zn);
			for (;;) {
				if (zn > 20 || r.err) { bad = 1; break; }
				if (zvec <= limit[t][zn]) break;
				zn++; zvec = (zvec << 1) | (int32_t)br_get(&r, 1);
			}
			if (bad) break;
			int32_t idx = zvec - base[t][zn];
			if (idx < 0 || idx >= BZO_MAX_ALPHA) { bad = 1; break; }
			int sym = perm[t][idx];
			if (sym == 0 || sym == 1) {               /* RUNA / RUNB */
				if (!in_run) { in_run = 1; run = 0; run_w = 1; }
				run += (sym == 0 ? 1 : 2) * run_w; run_w <<= 1;
				if (run > max_block) { bad = 1; break; }
				continue;
			}
			if (in_run) {
				uint8_t uc = seq_to_unseq[order[0]];
				if (nblock + run > max_block) { bad = 1; break; }
				unzftab[uc] += run;
				for (int32_t i = 0; i < run; i++) tt[nblock++] = uc;
				in_run = 0;
			}
			if (sym == eob) break;
			if (nblock >= max_block) { bad = 1; break; }
			{
				int p = sym - 1; uint8_t v = order[p];
				memmove(order + 1, order, (size_t)p); order[0] = v;
				uint8_t uc = seq_to_unseq[v];
				unzftab[uc]++; tt[nblock++] = uc;
			}
		}
		if (bad || r.err || orig_ptr < 0 || orig_ptr >= nblock) { rc = -2; break; }
		/* inverse BWT (decompress.c:494-573) */
		{
			int32_t cf[257]; cf[0] = 0;
			for (int i = 0; i < 256; i++) cf[i + 1] = cf[i] + unzftab[i];
			for (int32_t i = 0; i < nblock; i++) { uint8_t uc = (uint8_t)(tt[i] & 0xff); tt[cf[uc]++] |= ((uint32_t)i << 8); }
		}
		/* un-RLE1 + CRC (bzlib.c:561-728), with the de-randomisation of bzip2 <= 0.9.0 blocks (decompress.c BZ_RAND_*, bzlib.c:577-640) */
		{
			uint32_t crc = 0xFFFFFFFFu, pos = tt[orig_ptr] >> 8;
			int prev = -1, cnt = 0;
			int32_t n_to_go = 0, t_pos = 0;
			for (int32_t i = 0; i < nblock; i++) {
				pos = tt[pos]; uint8_t ch = (uint8_t)(pos & 0xff); pos >>= 8;
				if (randomised) {
					if (n_to_go == 0) { n_to_go = bz2o_rnums[t_pos]; t_pos++; if (t_pos == 512) t_pos = 0; }
					n_to_go--;
					if (n_to_go == 1) ch ^= 1;
				}
				if (cnt == 4) {                          /* ch is a repeat count */
					for (int k = 0; k < ch; k++) { if (op >= cap) { rc = -4; break; } out[op++] = (uint8_t)prev; crc = (crc << 8) ^ crc_tab[(crc >> 24) ^ (uint32_t)prev]; }
					if (rc) break;
					cnt = 0; prev = -1;
					continue;
				}
				if (ch == prev) cnt++; else { prev = ch; cnt = 1; }
				if (op >= cap) { rc = -4; break; }
				out[op++] = ch; crc = (crc << 8) ^ crc_tab[(crc >> 24) ^ ch];
			}
			if (rc) break;
			crc = ~crc;
			if (crc != stored_crc) { rc = -3; break; }
			combined = ((combined << 1) | (combined >> 31)) ^ crc;
		}
	}
	free(tt); free(selector);
	if (out_len) *out_len = op;
	return rc;
}
