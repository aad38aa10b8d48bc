/* TEST INFRASTRUCTURE ONLY -- see fmd_oracle.h.
 *
 * CPU restatement of the lh3/fermi FMD-index hot path, written from the behaviour of the
 * reference (citations are file:line of the reference).  The index lives in ONE flat array of
 * 64-bit words (the reference chunks it into 2^23-word pieces, rld.h:9-11; the chunking only
 * shows up here as the "one word less in the last block of a chunk" rule, rld.h:66) and the
 * decoder works on an absolute bit cursor instead of a (pointer, bits-left) pair.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <pthread.h>
#include <sys/time.h>
#include <math.h>
#include "fmd_oracle.h"

#define FO_CHUNK_WORDS (1ull << 23)     /* rld.h:9-10 */
#define FO_ASIZE 6
#define FO_SBITS 3

struct fo_index_s {
	int ssize;                  /* words per block = 1<<sbits (8 => 64 B) */
	int ibits;
	int hdr16, hdr32;           /* header size in words: 2 and 4 (rld.c:76-77) */
	uint64_t n_bytes, n_frames;
	uint64_t n_words, cap_words;
	uint64_t *w;                /* bit stream, zero padded by >= 2 words */
	uint64_t *frame;            /* n_frames x 7 */
	uint64_t cnt[8], mcnt[8];   /* cnt[c] = #symbols < c ; mcnt[0]=total, mcnt[c+1]=#c (rld.c:233,282-284) */
};

static __thread uint64_t tl_n_locate, tl_n_extend;

static double now_s(void)
{
	struct timeval tv;
	gettimeofday(&tv, 0);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static int floor_log2_u32(uint32_t v) { return v ? 31 - __builtin_clz(v) : -1; } /* rld.c:40-45 */

void fo_free(void *p) { free(p); }

/********************
 * bit-level access *
 ********************/

static inline int hdr_is32(uint64_t head) { return (uint32_t)head >> 31; }   /* rld.h:68 */

/* 64-bit window starting at absolute bit position b (w is padded) */
static inline uint64_t window_at(const uint64_t *w, uint64_t b)
{
	uint64_t i = b >> 6;
	int s = b & 63;
	return s ? (w[i] << s) | (w[i + 1] >> (64 - s)) : w[i];
}

/* last word of block `blk` in which a code may start (rld.h:66) */
static inline uint64_t block_tail(const fo_index_t *e, uint64_t blk)
{
	uint64_t end = blk + e->ssize;
	return end - ((end & (FO_CHUNK_WORDS - 1)) == 0 ? 2 : 1);
}

/* Decode one run at bit cursor *b (A.2 / rld.c:395-415): Elias-delta length then a 3-bit symbol.
 * Returns the run length, 0 if the window holds no code (>= 6 leading zeros, rld.h:84). */
static inline int64_t decode_run(const uint64_t *w, uint64_t *b, int *sym)
{
	uint64_t x = window_at(w, *b);
	int z, g, y;
	int64_t len;
	if (x >> 63) { /* "1": length 1 */
		*sym = (x >> 60) & 7;
		*b += 4;
		return 1;
	}
	z = __builtin_clzll(x | 1);
	if (z > 5) return 0;
	g = 2 * z + 1;                         /* width of the gamma-coded bit count */
	y = (int)(x >> (64 - g)) - 1;          /* number of mantissa bits */
	len = (int64_t)((x << g) >> (64 - y)) | (int64_t)(1u << y);
	*sym = (x << (g + y)) >> 61;
	*b += g + y + 3;
	return len;
}

/************
 * encoding *
 ************/

typedef struct {
	uint64_t p;        /* current word */
	int r;             /* free bits in word p */
	uint64_t head, tail;
	int pc;            /* pending symbol (-1: none) */
	int64_t pl;        /* pending length */
	uint64_t run[8], at_head[8]; /* running totals now / at the start of the current block: [0]=all, [c+1]=#c */
} fo_enc_t;

static void ensure_words(fo_index_t *e, uint64_t need)
{
	if (need + 4 > e->cap_words) {
		uint64_t nc = e->cap_words ? e->cap_words : 1024;
		while (nc < need + 4) nc <<= 1;
		e->w = (uint64_t*)realloc(e->w, nc * 8);
		memset(e->w + e->cap_words, 0, (nc - e->cap_words) * 8);
		e->cap_words = nc;
	}
}

static fo_index_t *index_new(void)
{
	fo_index_t *e = (fo_index_t*)calloc(1, sizeof(fo_index_t));
	e->ssize = 1 << FO_SBITS;
	e->hdr16 = (7 * 16 + 63) / 64;
	e->hdr32 = (7 * 32 + 63) / 64;
	return e;
}

static void enc_begin(fo_index_t *e, fo_enc_t *t)
{
	memset(t, 0, sizeof(*t));
	ensure_words(e, e->ssize * 2);
	t->head = 0; t->p = e->hdr16; t->r = 64; t->pc = -1;   /* first block: zero 16-bit header (rld.c:97-107) */
	t->tail = block_tail(e, 0);
}

/* open the next block and write the symbol counts of the block just closed (rld.c:111-134) */
static void enc_open_block(fo_index_t *e, fo_enc_t *t)
{
	int i;
	uint64_t d[7];
	t->head += e->ssize;
	ensure_words(e, t->head + 2 * e->ssize);
	for (i = 0; i < 7; ++i) d[i] = t->run[i] - t->at_head[i];
	if (d[0] >= 0x8000) {
		uint32_t *h = (uint32_t*)(e->w + t->head);
		for (i = 0; i < 7; ++i) h[i] = (uint32_t)d[i];
		h[0] |= 1u << 31;
		t->p = t->head + e->hdr32;
	} else {
		uint16_t *h = (uint16_t*)(e->w + t->head);
		for (i = 0; i < 7; ++i) h[i] = (uint16_t)d[i];
		t->p = t->head + e->hdr16;
	}
	t->tail = block_tail(e, t->head);
	t->r = 64;
	for (i = 0; i < 7; ++i) t->at_head[i] = t->run[i];
}

/* emit one maximal run (rld.c:47-53,159-173) */
static void enc_emit(fo_index_t *e, fo_enc_t *t, int64_t l, int c)
{
	int y = floor_log2_u32((uint32_t)l), z = floor_log2_u32(y + 1);
	int w = 2 * z + 1 + y + 3;
	uint64_t code = ((((uint64_t)l ^ (1ull << y)) | (uint64_t)(y + 1) << y) << 3) | (uint64_t)c;
	if (w >= t->r && t->p == t->tail) enc_open_block(e, t);
	if (w > t->r) {
		w -= t->r;
		e->w[t->p++] |= code >> w;
		t->r = 64 - w;
		e->w[t->p] = code << t->r;
	} else {
		t->r -= w;
		e->w[t->p] |= code << t->r;
	}
	t->run[0] += l; t->run[c + 1] += l;
}

/* rld_enc, rld.c:176-184: merge adjacent runs of the same symbol */
static void enc_push(fo_index_t *e, fo_enc_t *t, int64_t l, int c)
{
	if (l == 0) return;
	if (t->pc != c) {
		if (t->pl) enc_emit(e, t, t->pl, t->pc);
		t->pl = l; t->pc = c;
	} else t->pl += l;
}

/* rld_rank_index, rld.c:186-224 */
static void build_frames(fo_index_t *e)
{
	uint64_t n_blks = e->n_bytes * 8 / 64 / e->ssize + 1;
	uint64_t last = (e->n_bytes >> 3 >> FO_SBITS) << FO_SBITS;      /* rld_last_blk, rld.h:62 */
	uint64_t acc[6] = {0, 0, 0, 0, 0, 0}, i, k, f;
	int j;
	e->ibits = floor_log2_u32((uint32_t)(e->mcnt[0] / n_blks)) + 4;
	e->n_frames = ((e->mcnt[0] + (1ull << e->ibits) - 1) >> e->ibits) + 1;
	e->frame = (uint64_t*)calloc(e->n_frames * 7, 8);
	for (i = e->ssize, k = 1; i <= last; i += e->ssize) {
		uint64_t sum = 0, head = e->w[i];
		if (hdr_is32(head)) {
			const uint32_t *h = (const uint32_t*)(e->w + i);
			for (j = 0; j < 6; ++j) acc[j] += h[j + 1];
		} else {
			const uint16_t *h = (const uint16_t*)(e->w + i);
			for (j = 0; j < 6; ++j) acc[j] += h[j + 1];
		}
		for (j = 0; j < 6; ++j) sum += acc[j];
		while (sum >= k << e->ibits) ++k;
		if (k < e->n_frames) {
			e->frame[k * 7] = i;
			for (j = 0; j < 6; ++j) e->frame[k * 7 + 1 + j] = acc[j];
		}
	}
	for (f = 1; f < e->n_frames; ++f)          /* rows no block start fell into inherit the previous row */
		if (e->frame[f * 7] == 0)
			memcpy(e->frame + f * 7, e->frame + (f - 1) * 7, 7 * 8);
}

/* rld_enc_finish, rld.c:226-236 */
static void enc_end(fo_index_t *e, fo_enc_t *t)
{
	int i;
	if (t->pl) enc_emit(e, t, t->pl, t->pc);
	enc_open_block(e, t);                  /* trailing header-only pseudo block */
	e->n_words = t->p;
	e->n_bytes = t->p * 8;
	for (i = 0; i < 7; ++i) e->mcnt[i] = t->run[i];
	e->cnt[0] = 0;
	for (i = 1; i <= 6; ++i) e->cnt[i] = e->cnt[i - 1] + e->mcnt[i];
	e->cnt[7] = e->cnt[6];
	build_frames(e);
}

fo_index_t *fo_from_bwt(int64_t n, const uint8_t *bwt)
{
	fo_index_t *e = index_new();
	fo_enc_t t;
	int64_t i, k = 1;
	int c = bwt[0];
	enc_begin(e, &t);
	for (i = 1; i < n; ++i) {
		if (bwt[i] != c) { enc_push(e, &t, k, c); c = bwt[i]; k = 1; }
		else ++k;
	}
	enc_push(e, &t, k, c);
	enc_end(e, &t);
	return e;
}

fo_index_t *fo_from_rle6(int64_t n, const uint8_t *rle)
{
	fo_index_t *e = index_new();
	fo_enc_t t;
	int64_t i;
	enc_begin(e, &t);
	for (i = 0; i < n; ++i)
		if (rle[i] >> 3) enc_push(e, &t, rle[i] >> 3, rle[i] & 7);
	enc_end(e, &t);
	return e;
}

/***************
 * file format *
 ***************/

int fo_dump(const fo_index_t *e, const char *fn)
{
	FILE *fp = fopen(fn, "wb");
	uint32_t a = FO_ASIZE << 16 | FO_SBITS;
	uint64_t zero = 0;
	if (fp == 0) return -1;
	fwrite("RLD\2", 1, 4, fp);
	fwrite(&a, 4, 1, fp);
	fwrite(&zero, 8, 1, fp);
	fwrite(&e->n_bytes, 8, 1, fp);
	fwrite(&e->n_frames, 8, 1, fp);
	fwrite(e->mcnt + 1, 8, 6, fp);
	fwrite(e->w, 8, e->n_bytes / 8, fp);
	fwrite(e->frame, 8 * 7, e->n_frames, fp);
	fclose(fp);
	return 0;
}

fo_index_t *fo_load(const char *fn)
{
	FILE *fp = fopen(fn, "rb");
	char magic[4];
	fo_index_t *e;
	if (fp == 0) return 0;
	if (fread(magic, 1, 4, fp) != 4) { fclose(fp); return 0; }
	if (memcmp(magic, "RLD\2", 4) == 0) {
		uint32_t a;
		uint64_t h[3], n_blks;
		int i;
		e = index_new();
		if (fread(&a, 4, 1, fp) != 1 || fread(h, 8, 3, fp) != 3 || (a >> 16) != FO_ASIZE || (a & 0xffff) != FO_SBITS) {
			fclose(fp); free(e); return 0;
		}
		e->n_bytes = h[1]; e->n_frames = h[2]; e->n_words = e->n_bytes / 8;
		if (fread(e->mcnt + 1, 8, 6, fp) != 6) { fclose(fp); free(e); return 0; }
		for (i = 1, e->cnt[0] = 0; i <= 6; ++i) e->cnt[i] = e->cnt[i - 1] + e->mcnt[i];
		e->cnt[7] = e->cnt[6];
		e->mcnt[0] = e->cnt[6];
		ensure_words(e, e->n_words);
		e->frame = (uint64_t*)malloc(e->n_frames * 7 * 8);
		if (fread(e->w, 8, e->n_words, fp) != e->n_words || fread(e->frame, 56, e->n_frames, fp) != e->n_frames) {
			fclose(fp); fo_destroy(e); return 0;
		}
		n_blks = e->n_bytes * 8 / 64 / e->ssize + 1;
		e->ibits = floor_log2_u32((uint32_t)(e->mcnt[0] / n_blks)) + 4;   /* rld.c:322-323 */
	} else {
		/* anything else is taken as a raw byte-RLE stream read from the CURRENT offset
		 * (rld.c:295-302): the 4 bytes already consumed ("RLE\6", ropebwt.c:133) are skipped. */
		uint8_t *buf = (uint8_t*)malloc(1 << 16);
		fo_enc_t t;
		size_t l, i;
		e = index_new();
		enc_begin(e, &t);
		while ((l = fread(buf, 1, 1 << 16, fp)) != 0)
			for (i = 0; i < l; ++i)
				if (buf[i] >> 3) enc_push(e, &t, buf[i] >> 3, buf[i] & 7);
		free(buf);
		enc_end(e, &t);
	}
	fclose(fp);
	return e;
}

void fo_destroy(fo_index_t *e)
{
	if (e == 0) return;
	free(e->w); free(e->frame); free(e);
}

void fo_info(const fo_index_t *e, uint64_t out[17])
{
	int i;
	for (i = 0; i < 7; ++i) out[i] = e->mcnt[i], out[7 + i] = e->cnt[i];
	out[14] = e->n_bytes; out[15] = e->n_frames; out[16] = e->ibits;
}

const uint64_t *fo_words(const fo_index_t *e) { return e->w; }
const uint64_t *fo_frame(const fo_index_t *e) { return e->frame; }

int64_t fo_decode_bwt(const fo_index_t *e, uint8_t *out)
{
	uint64_t last = (e->n_bytes >> 3 >> FO_SBITS) << FO_SBITS, blk;
	int64_t n = 0;
	for (blk = 0; blk < last; blk += e->ssize) {
		uint64_t b = (blk + (hdr_is32(e->w[blk]) ? e->hdr32 : e->hdr16)) * 64;
		uint64_t end = (block_tail(e, blk) + 1) * 64;
		while (b < end) {
			int c;
			uint64_t x = window_at(e->w, b);
			int64_t l;
			if (end - b < 64) x &= ~0ull << (64 - (end - b));     /* zero pad past the tail word (rld.h:82) */
			if (x == 0) break;
			{ /* decode from the masked window */
				uint64_t tmpw[2] = { x, 0 }, tb = 0;
				l = decode_run(tmpw, &tb, &c);
				if (l == 0 || c > FO_ASIZE) break;
				b += tb;
			}
			if (out) memset(out + n, c, l);
			n += l;
		}
	}
	return n;
}

/********
 * rank *
 ********/

/* rld_locate_blk, rld.c:352-392.  Returns the first coordinate after the block holding k;
 * cnt[] = symbol counts before that block, *z = coordinates before it, *bit = payload cursor. */
static inline uint64_t locate(const fo_index_t *e, uint64_t k, uint64_t cnt[6], uint64_t *z, uint64_t *bit)
{
	const uint64_t *row = e->frame + (k >> e->ibits) * 7;
	uint64_t blk = row[0], sum = 0, c;
	int j;
	++tl_n_locate;
	for (j = 0; j < 6; ++j) sum += (cnt[j] = row[1 + j]);
	for (;;) {
		uint64_t nxt = blk + e->ssize, head = e->w[nxt];
		if (hdr_is32(head)) {
			const uint32_t *h = (const uint32_t*)(e->w + nxt);
			c = h[0] & 0x7fffffff;
			if (sum + c > k) break;
			for (j = 0; j < 6; ++j) cnt[j] += h[1 + j];
		} else {
			const uint16_t *h = (const uint16_t*)(e->w + nxt);
			c = h[0];
			if (sum + c > k) break;
			for (j = 0; j < 6; ++j) cnt[j] += h[1 + j];
		}
		sum += c;
		blk = nxt;
	}
	*z = sum;
	*bit = (blk + (hdr_is32(e->w[blk]) ? e->hdr32 : e->hdr16)) * 64;
	return sum + c;
}

/* rld_rank1a, rld.c:424-446 */
int fo_rank1a(const fo_index_t *e, uint64_t k, uint64_t ok[6])
{
	uint64_t z, bit;
	int64_t l;
	int a = -1;
	if (k == (uint64_t)-1) { memset(ok, 0, 48); return -1; }
	locate(e, k, ok, &z, &bit);
	++k;
	for (;;) {
		l = decode_run(e->w, &bit, &a);
		if (z + l >= k) break;
		z += l; ok[a] += l;
	}
	ok[a] += k - z;
	return a;
}

/* rld_rank2a, rld.c:457-492 */
void fo_rank2a(const fo_index_t *e, uint64_t k, uint64_t l, uint64_t ok[6], uint64_t ol[6])
{
	uint64_t z, y, bit;
	int64_t len;
	int a = -1;
	if (k == (uint64_t)-1) { memset(ok, 0, 48); fo_rank1a(e, l, ol); return; }
	y = locate(e, k, ok, &z, &bit);
	++k;
	for (;;) {
		len = decode_run(e->w, &bit, &a);
		if (z + len >= k) break;
		z += len; ok[a] += len;
	}
	if (y > l) { /* l is in the same block: keep decoding */
		++l;
		memcpy(ol, ok, 48);
		ok[a] += k - z;
		if (z + len < l) {
			z += len; ol[a] += len;
			for (;;) {
				len = decode_run(e->w, &bit, &a);
				if (z + len >= l) break;
				z += len; ol[a] += len;
			}
		}
		ol[a] += l - z;
	} else {
		ok[a] += k - z;
		fo_rank1a(e, l, ol);
	}
}

void fo_rank1a_batch(const fo_index_t *e, int64_t n, const uint64_t *k, uint64_t *ok, int32_t *sym)
{
	int64_t i;
	for (i = 0; i < n; ++i) sym[i] = fo_rank1a(e, k[i], ok + 6 * i);
}

void fo_rank2a_batch(const fo_index_t *e, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol)
{
	int64_t i;
	for (i = 0; i < n; ++i) fo_rank2a(e, k[i], l[i], ok + 6 * i, ol + 6 * i);
}

/******************
 * FMD primitives *
 ******************/

static inline int comp6(int c) { return c >= 1 && c <= 4 ? 5 - c : c; }   /* fermi.h:52 */

static inline void set_intv(const fo_index_t *e, int c, fo_intv_t *ik)   /* fermi.h:53 */
{
	ik->x[0] = e->cnt[c]; ik->x[2] = e->cnt[c + 1] - e->cnt[c]; ik->x[1] = e->cnt[comp6(c)]; ik->info = 0;
}

/* fm6_extend, exact.c:72-88 (A.5) */
int fo_extend(const fo_index_t *e, const fo_intv_t *ik, fo_intv_t ok[6], int is_back)
{
	uint64_t tk[6], tl[6];
	int c, far = !is_back, near = !!is_back;
	static const int order[6] = { 0, 4, 3, 2, 1, 5 };
	uint64_t acc;
	++tl_n_extend;
	fo_rank2a(e, ik->x[far] - 1, ik->x[far] - 1 + ik->x[2], tk, tl);
	for (c = 0; c < 6; ++c) {
		ok[c].x[far] = e->cnt[c] + tk[c];
		ok[c].x[2] = tl[c] - tk[c];
	}
	for (c = 0, acc = ik->x[near]; c < 6; ++c) {
		ok[order[c]].x[near] = acc;
		acc += ok[order[c]].x[2];
	}
	return 0;
}

/* fm6_extend0, exact.c:90-98 */
static void extend0(const fo_index_t *e, const fo_intv_t *ik, fo_intv_t *ok0, int is_back)
{
	uint64_t tk[6], tl[6];
	++tl_n_extend;
	fo_rank2a(e, ik->x[!is_back] - 1, ik->x[!is_back] - 1 + ik->x[2], tk, tl);
	ok0->x[!is_back] = tk[0];
	ok0->x[!!is_back] = ik->x[!!is_back];
	ok0->x[2] = tl[0] - tk[0];
}

void fo_extend_batch(const fo_index_t *e, int64_t n, const fo_intv_t *ik, const uint8_t *is_back, fo_intv_t *ok6)
{
	int64_t i;
	int c;
	for (i = 0; i < n; ++i) {
		fo_extend(e, ik + i, ok6 + 6 * i, is_back[i]);
		for (c = 0; c < 6; ++c) ok6[6 * i + c].info = 0;
	}
}

/* fm_backward_search, exact.c:7-23 (two independent rank11, rld.c:418-422) */
uint64_t fo_backward_search(const fo_index_t *e, int len, const uint8_t *s, uint64_t *sa_beg, uint64_t *sa_end)
{
	uint64_t k, l, tk[6], tl[6];
	int i, c = s[len - 1];
	k = e->cnt[c]; l = e->cnt[c + 1] - 1;
	for (i = len - 2; i >= 0; --i) {
		c = s[i];
		fo_rank1a(e, k - 1, tk);
		fo_rank1a(e, l, tl);
		k = e->cnt[c] + tk[c];
		l = e->cnt[c] + tl[c] - 1;
		if (k > l) break;
	}
	if (k > l) return 0;
	*sa_beg = k; *sa_end = l;
	return l - k + 1;
}

void fo_backward_search_batch(const fo_index_t *e, int64_t n, const uint8_t *seq, const uint64_t *off,
							  uint64_t *sa_beg, uint64_t *sa_end, uint64_t *size)
{
	int64_t i;
	for (i = 0; i < n; ++i) {
		sa_beg[i] = sa_end[i] = 0;
		size[i] = fo_backward_search(e, (int)(off[i + 1] - off[i]), seq + off[i], &sa_beg[i], &sa_end[i]);
	}
}

/* fm_retrieve, exact.c:59-70: LF walk from sentinel rank x, emits the read reversed */
int64_t fo_retrieve(const fo_index_t *e, uint64_t x, uint8_t *out, int max_len, int *len)
{
	uint64_t k = x, ok[6];
	int n = 0;
	for (;;) {
		int c = fo_rank1a(e, k, ok);
		k = e->cnt[c] + ok[c] - 1;
		if (c == 0) break;
		if (n < max_len) out[n] = c;
		++n;
	}
	*len = n;
	return k;
}

/********
 * SMEM *
 ********/

typedef struct { size_t n, m; fo_intv_t *a; } ivec_t;

static inline void iv_push(ivec_t *v, const fo_intv_t *x)
{
	if (v->n == v->m) {
		v->m = v->m ? v->m << 1 : 16;
		v->a = (fo_intv_t*)realloc(v->a, v->m * sizeof(fo_intv_t));
	}
	v->a[v->n++] = *x;
}

static void iv_reverse(ivec_t *v)
{
	size_t i;
	for (i = 0; i < v->n >> 1; ++i) {
		fo_intv_t t = v->a[i]; v->a[i] = v->a[v->n - 1 - i]; v->a[v->n - 1 - i] = t;
	}
}

/* fm6_smem1_core, smem.c:13-80 (S1 / A.6): all SMEMs covering query position x; returns next x */
static int smem1(const fo_index_t *e, int len, const uint8_t *q, int x, ivec_t *mem, int self_match, ivec_t *va, ivec_t *vb)
{
	ivec_t *prev = va, *curr = vb, *t;
	fo_intv_t ik, ok[6];
	int i, c, ret;
	size_t j;

	prev->n = curr->n = 0;
	set_intv(e, q[x], &ik);
	ik.info = x + 1;
	for (i = x + 1; i < len; ++i) { /* forward sweep */
		c = comp6(q[i]);
		fo_extend(e, &ik, ok, 0);
		if (ok[c].x[2] != ik.x[2]) {
			if (ik.x[2] != ok[0].x[2]) iv_push(curr, &ik);
			if (!self_match && ok[0].x[2]) { ok[0].info = i; iv_push(curr, &ok[0]); }
		}
		if (self_match ? ok[c].x[2] < 2 : ok[c].x[2] == 0) break;
		ik = ok[c]; ik.info = i + 1;
	}
	if (i == len) {
		iv_push(curr, &ik);
		if (!self_match) {
			fo_extend(e, &ik, ok, 0);
			if (ok[0].x[2]) { ok[0].info = len; iv_push(curr, &ok[0]); }
		}
	}
	iv_reverse(curr);
	ret = (int)curr->a[0].info;
	t = curr; curr = prev; prev = t;

	mem->n = 0;
	for (i = x - 1; i >= -1; --i) { /* backward sweep */
		c = i < 0 ? 0 : q[i];
		for (j = 0, curr->n = 0; j < prev->n; ++j) {
			fo_intv_t *p = &prev->a[j];
			int cont, fl;
			fo_extend(e, p, ok, 1);
			fl = ok[0].x[2] && p->x[1] < e->mcnt[1];
			cont = self_match ? ok[c].x[2] > 1 : ok[c].x[2] != 0;
			if (!cont || fl || i == -1) {
				if (curr->n == 0 || fl) {
					if (fl || mem->n == 0 || (uint64_t)(i + 1) < (mem->a[mem->n - 1].info >> 32 & 0x3fffffff)) {
						ik = *p;
						ik.info |= (uint64_t)(ok[0].x[2] != 0) << 63 | (uint64_t)(i + 1) << 32;
						iv_push(mem, &ik);
					}
				}
			}
			if (cont && (p->x[1] < e->mcnt[1] || curr->n == 0 || ok[c].x[2] != curr->a[curr->n - 1].x[2])) {
				ok[c].info = p->info;
				iv_push(curr, &ok[c]);
			}
		}
		if (curr->n == 0) break;
		t = curr; curr = prev; prev = t;
	}
	iv_reverse(mem);
	return ret;
}

typedef struct {
	const fo_index_t *e;
	int64_t n, start, step;
	const uint8_t *seq;
	const uint64_t *off;
	int self_match;
	ivec_t out;
	uint32_t *cnt;
	uint64_t n_locate, n_extend;
} smem_job_t;

/* fm6_smem, smem.c:397-410: repeat smem1 from the returned x until the end of the read */
static void *smem_job(void *data)
{
	smem_job_t *w = (smem_job_t*)data;
	ivec_t mem = {0, 0, 0}, va = {0, 0, 0}, vb = {0, 0, 0};
	int64_t i;
	size_t j;
	tl_n_locate = tl_n_extend = 0;
	for (i = w->start; i < w->n; i += w->step) {
		int len = (int)(w->off[i + 1] - w->off[i]), x = 0;
		const uint8_t *q = w->seq + w->off[i];
		uint32_t c = 0;
		while (x < len) {
			x = smem1(w->e, len, q, x, &mem, w->self_match, &va, &vb);
			for (j = 0; j < mem.n; ++j) iv_push(&w->out, &mem.a[j]);
			c += mem.n;
		}
		w->cnt[i] = c;
	}
	w->n_locate = tl_n_locate; w->n_extend = tl_n_extend;
	free(mem.a); free(va.a); free(vb.a);
	return 0;
}

int fo_smem_batch(const fo_index_t *e, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match, int n_threads,
				  fo_intv_t **mem, uint64_t *mem_off, double *secs, uint64_t *n_locate, uint64_t *n_extend)
{
	smem_job_t *w;
	pthread_t *tid;
	uint32_t *cnt;
	size_t *cur;
	int64_t i;
	int t;
	double t0;
	if (n_threads < 1) n_threads = 1;
	w = (smem_job_t*)calloc(n_threads, sizeof(smem_job_t));
	tid = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
	cnt = (uint32_t*)calloc(n + 1, 4);
	for (t = 0; t < n_threads; ++t) {
		w[t].e = e; w[t].n = n; w[t].start = t; w[t].step = n_threads;
		w[t].seq = seq; w[t].off = off; w[t].self_match = self_match; w[t].cnt = cnt;
	}
	t0 = now_s();
	if (n_threads == 1) smem_job(&w[0]);
	else {
		for (t = 0; t < n_threads; ++t) pthread_create(&tid[t], 0, smem_job, &w[t]);
		for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
	}
	if (secs) *secs = now_s() - t0;
	for (i = 0, mem_off[0] = 0; i < n; ++i) mem_off[i + 1] = mem_off[i] + cnt[i];
	if (mem) {
		*mem = (fo_intv_t*)malloc((mem_off[n] ? mem_off[n] : 1) * sizeof(fo_intv_t));
		cur = (size_t*)calloc(n_threads, sizeof(size_t));
		for (i = 0; i < n; ++i) {
			t = i % n_threads;
			memcpy(*mem + mem_off[i], w[t].out.a + cur[t], cnt[i] * sizeof(fo_intv_t));
			cur[t] += cnt[i];
		}
		free(cur);
	}
	if (n_locate) for (t = 0, *n_locate = 0; t < n_threads; ++t) *n_locate += w[t].n_locate;
	if (n_extend) for (t = 0, *n_extend = 0; t < n_threads; ++t) *n_extend += w[t].n_extend;
	for (t = 0; t < n_threads; ++t) free(w[t].out.a);
	free(w); free(tid); free(cnt);
	return 0;
}

/***********
 * overlap *
 ***********/

typedef struct { size_t n, m; int *a; } cvec_t;
typedef struct { int l, m; uint8_t *s; } bstr_t;

static void bs_reserve(bstr_t *s, int m)
{
	if (m > s->m) { s->m = m + 64; s->s = (uint8_t*)realloc(s->s, s->m); }
}

static void bs_putc(bstr_t *s, int c) { bs_reserve(s, s->l + 2); s->s[s->l++] = c; s->s[s->l] = 0; }

/* overlap_intv, unitig.c:38-64 (U1) */
static fo_intv_t overlap_intv(const fo_index_t *e, int len, const uint8_t *seq, int min, int j, int at5, ivec_t *p, int inc_sentinel)
{
	int c, depth, dir = at5 ? 1 : -1, end = at5 ? len : -1;
	fo_intv_t ik, ok[6];
	p->n = 0;
	set_intv(e, seq[j], &ik);
	for (depth = 1, j += dir; j != end; j += dir, ++depth) {
		c = at5 ? comp6(seq[j]) : seq[j];
		fo_extend(e, &ik, ok, !at5);
		if (!ok[c].x[2]) break;
		if (depth >= min && ok[0].x[2]) {
			if (inc_sentinel) { ok[0].info = j - dir; iv_push(p, &ok[0]); }
			else { ik.info = j - dir; iv_push(p, &ik); }
		}
		ik = ok[c];
	}
	iv_reverse(p);
	return ik;
}

/* fm6_is_contained, unitig.c:77-91 (U2) */
static int is_contained(const fo_index_t *e, int min_match, const bstr_t *s, fo_intv_t *intv, ivec_t *ovlp)
{
	fo_intv_t ik, ok[6];
	int ret = 0;
	ovlp->n = 0;
	ik = overlap_intv(e, s->l, s->s, min_match, s->l - 1, 0, ovlp, 0);
	fo_extend(e, &ik, ok, 1);
	if (ik.x[2] != ok[0].x[2]) ret = -1;
	ik = ok[0];
	fo_extend(e, &ik, ok, 0);
	if (ik.x[2] != ok[0].x[2]) ret = -1;
	*intv = ok[0];
	return ret;
}

static int cmp_info(const void *a, const void *b)
{
	uint64_t x = ((const fo_intv_t*)a)->info, y = ((const fo_intv_t*)b)->info;
	return x < y ? -1 : x > y;
}

/* fm6_get_nei, unitig.c:93-179 (U3 / A.7), without the optional `used` bitmap */
static int get_nei(const fo_index_t *e, int min_match, int beg, bstr_t *s, ivec_t *nei, ivec_t *prev, ivec_t *curr, cvec_t *cat)
{
	int ori_l = s->l, c, rbeg, is_forked = 0;
	size_t i, j;
	ivec_t *t;
	fo_intv_t ok[6], ok0;

	curr->n = nei->n = 0;
	if (prev->n == 0) {
		overlap_intv(e, s->l - beg, s->s + beg, min_match, s->l - beg - 1, 0, prev, 0);
		if (prev->n == 0) return -1;
		for (j = 0; j < prev->n; ++j) prev->a[j].info += beg;
	}
#define CAT_RESERVE(need) do { if ((need) > cat->m) { cat->m = (need) + 16; cat->a = (int*)realloc(cat->a, cat->m * sizeof(int)); } } while (0)
	CAT_RESERVE(prev->n + 1);
	for (j = 0; j < prev->n; ++j) cat->a[j] = 0;
	while (prev->n) {
		for (j = 0, curr->n = 0; j < prev->n; ++j) {
			fo_intv_t *p = &prev->a[j];
			if (cat->a[j] < 0) continue;
			fo_extend(e, p, ok, 0);
			if (ok[0].x[2] && ori_l != s->l) {
				extend0(e, &ok[0], &ok0, 1);
				if (ok0.x[2]) {
					if (ok[0].x[2] == p->x[2] && p->x[2] == ok0.x[2]) {
						int cat0 = cat->a[j];
						ok0.info = ori_l - (p->info & 0xffffffffu);
						for (i = j; i < prev->n && cat->a[i] == cat0; ++i) cat->a[i] = -1;
						iv_push(nei, &ok0);
						continue;
					}
				}
			}
			if (cat->a[j] < 0) continue;
			for (c = 1; c < 5; ++c)
				if (ok[c].x[2]) {
					extend0(e, &ok[c], &ok0, 1);
					if (ok0.x[2]) {
						ok[c].info = (p->info & 0xfffffff0ffffffffull) | (uint64_t)c << 32;
						iv_push(curr, &ok[c]);
					}
				}
		}
		if (curr->n) {
			uint32_t last, cat0;
			CAT_RESERVE(curr->n + 1);
			c = curr->a[0].info >> 32 & 0xf;
			bs_putc(s, comp6(c));
			qsort(curr->a, curr->n, sizeof(fo_intv_t), cmp_info);   /* keys are unique (unitig.c:142) */
			last = curr->a[0].info >> 32;
			cat->a[0] = 0;
			curr->a[0].info &= 0xffffffff;
			for (j = 1, cat0 = 0; j < curr->n; ++j) {
				if (curr->a[j].info >> 32 != last) last = curr->a[j].info >> 32, cat0 = j;
				cat->a[j] = cat0;
				curr->a[j].info = (curr->a[j].info & 0xffffffff) | (uint64_t)cat0 << 36;
			}
			if (cat0 != 0) is_forked = 1;
		}
		t = curr; curr = prev; prev = t;
	}
	if (nei->n == 0) return -1;
	rbeg = ori_l - (uint32_t)nei->a[0].info;
	if (nei->n == 1 && is_forked) { /* contained reads forked the path: rebuild along the one neighbour */
		int k;
		set_intv(e, 0, &ok0);
		for (k = rbeg; k < ori_l; ++k) {
			fo_extend(e, &ok0, ok, 0);
			ok0 = ok[comp6(s->s[k])];
		}
		for (k = ori_l; k < s->l; ++k) {
			int c0 = -1, hits = 0;
			fo_extend(e, &ok0, ok, 0);
			for (c = 1; c < 5; ++c)
				if (ok[c].x[2] && ok[c].x[0] <= nei->a[0].x[0] && ok[c].x[0] + ok[c].x[2] >= nei->a[0].x[0] + nei->a[0].x[2])
					++hits, c0 = c;
			if (hits == 0 && ok[0].x[2]) break;
			assert(hits == 1);
			s->s[k] = comp6(c0);
			ok0 = ok[c0];
		}
		s->l = k; s->s[s->l] = 0;
	}
	if (nei->n > 1) s->l = ori_l, s->s[s->l] = 0;
	return rbeg;
}

int fo_overlap_batch(const fo_index_t *e, int min_match, int64_t n, const uint64_t *seeds, int64_t *rec,
					 fo_intv_t **nei_out, uint64_t *nei_off, uint64_t *n_locate)
{
	bstr_t s = {0, 0, 0};
	ivec_t a[2] = {{0, 0, 0}, {0, 0, 0}}, nei = {0, 0, 0}, all = {0, 0, 0};
	cvec_t cat = {0, 0, 0};
	int64_t i;
	size_t j;
	tl_n_locate = 0;
	nei_off[0] = 0;
	bs_reserve(&s, 1 << 16);
	for (i = 0; i < n; ++i) {
		int64_t *r = rec + 9 * i;
		fo_intv_t intv0;
		int ret, rbeg, len, k;
		memset(r, 0, 9 * sizeof(int64_t));
		r[0] = fo_retrieve(e, seeds[i], s.s, s.m - 1, &len);
		s.l = len; s.s[len] = 0;
		for (k = 0; k < len >> 1; ++k) { uint8_t t = s.s[k]; s.s[k] = s.s[len - 1 - k]; s.s[len - 1 - k] = t; }
		r[1] = len; r[6] = -1;
		a[0].n = a[1].n = nei.n = 0;
		if (len <= min_match) { r[2] = -9; nei_off[i + 1] = all.n; continue; }
		ret = is_contained(e, min_match, &s, &intv0, &a[0]);
		r[2] = ret; r[3] = intv0.x[0]; r[4] = intv0.x[1]; r[5] = intv0.x[2];
		if (ret >= 0 && a[0].n) {
			rbeg = get_nei(e, min_match, 0, &s, &nei, &a[0], &a[1], &cat);
			r[6] = rbeg; r[7] = nei.n; r[8] = s.l;
			for (j = 0; j < nei.n; ++j) iv_push(&all, &nei.a[j]);
		}
		nei_off[i + 1] = all.n;
	}
	*nei_out = all.a ? all.a : (fo_intv_t*)malloc(1);
	if (n_locate) *n_locate = tl_n_locate;
	free(s.s); free(a[0].a); free(a[1].a); free(nei.a); free(cat.a);
	return 0;
}

/* check_left_simple, unitig.c:186-204 (U4): 0 = the unique neighbour has no other left neighbour, -1 = potential backward
 * bifurcation.  s = the consensus grown by get_nei, rbeg = start of the neighbour in it. */
static int check_left_simple(const fo_index_t *e, int min_match, int beg, int rbeg, const bstr_t *s, ivec_t *prev, ivec_t *curr)
{
	fo_intv_t ok[6];
	ivec_t *swap;
	int i;
	size_t j;
	overlap_intv(e, s->l, s->s, min_match, rbeg, 1, prev, 1);
	for (i = rbeg - 1; i >= beg; --i) {
		for (j = 0, curr->n = 0; j < prev->n; ++j) {
			fo_intv_t *p = &prev->a[j];
			fo_extend(e, p, ok, 1);
			if (ok[0].x[2] + ok[s->s[i]].x[2] != p->x[2]) return -1;
			iv_push(curr, &ok[s->s[i]]);
		}
		swap = curr; curr = prev; prev = swap;
	}
	return 0;
}

/* Block lookups (calls of the rld_locate_blk restatement) of the work `fermi unitig` does for ONE sequence that its walk visits
 * (unitig.c:227-317): locates[0] fm_retrieve of the seed, [1] fm6_is_contained (its overlap_intv is the one fm6_get_nei would run
 * for a read reached through a neighbour, unitig.c:100-104), [2] fm6_get_nei, [3] check_left_simple when there is exactly one
 * neighbour; left[i] (optional) = result of check_left_simple (1 = not evaluated).  bench.py sums the four over a sample of the
 * seed rows of the benchmark index: SURVEY.md 8d's N_locate per input read. */
int fo_unitig_locates(const fo_index_t *e, int min_match, int64_t n, const uint64_t *seeds, uint64_t locates[4], int8_t *left)
{
	bstr_t s = {0, 0, 0};
	ivec_t a[2] = {{0, 0, 0}, {0, 0, 0}}, nei = {0, 0, 0};
	cvec_t cat = {0, 0, 0};
	int64_t i;
	locates[0] = locates[1] = locates[2] = locates[3] = 0;
	bs_reserve(&s, 1 << 16);
	for (i = 0; i < n; ++i) {
		fo_intv_t intv0;
		int ret, rbeg, len, k;
		uint64_t c0;
		if (left) left[i] = 1;
		c0 = tl_n_locate = 0;
		fo_retrieve(e, seeds[i], s.s, s.m - 1, &len);
		locates[0] += tl_n_locate - c0; c0 = tl_n_locate;
		s.l = len; s.s[len] = 0;
		for (k = 0; k < len >> 1; ++k) { uint8_t t = s.s[k]; s.s[k] = s.s[len - 1 - k]; s.s[len - 1 - k] = t; }
		a[0].n = a[1].n = nei.n = 0;
		if (len <= min_match) continue;
		ret = is_contained(e, min_match, &s, &intv0, &a[0]);
		locates[1] += tl_n_locate - c0; c0 = tl_n_locate;
		if (ret < 0 || a[0].n == 0) continue;
		rbeg = get_nei(e, min_match, 0, &s, &nei, &a[0], &a[1], &cat);
		locates[2] += tl_n_locate - c0; c0 = tl_n_locate;
		if (rbeg >= 0 && nei.n == 1) {
			a[0].n = a[1].n = 0;
			ret = check_left_simple(e, min_match, 0, rbeg, &s, &a[0], &a[1]);
			locates[3] += tl_n_locate - c0;
			if (left) left[i] = (int8_t)ret;
		}
	}
	free(s.s); free(a[0].a); free(a[1].a); free(nei.a); free(cat.a);
	return 0;
}

/*******************
 * k-mer collection *
 *******************/

static int cmp_u64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
	return x < y ? -1 : x > y;
}

/* One recursive walk of the backward-extension trie.  Above depth suf_len every non-empty child is followed
 * (fm6_traverse, exact.c:158-164); below it only children with >= min_occ occurrences (correct.c:77-82); at depth w
 * the node is summarised (correct.c:56-75).  `path` packs the prepended bases, 2 bits each, first base lowest. */
typedef struct { uint64_t *a; uint64_t n, m; int64_t cnt[2]; int w, suf_len; uint64_t min_occ; } eccol_t;

static void ec_walk(const fo_index_t *e, eccol_t *z, const fo_intv_t *ik, int depth, uint64_t path)
{
	fo_intv_t ok[6];
	int c;
	fo_extend(e, ik, ok, 1);
	if (depth == z->w) {
		uint64_t max = 0, rest, key, suffix;
		int max_c = 6;
		double r;
		for (c = 1; c <= 4; ++c) if (ok[c].x[2] > max) max = ok[c].x[2], max_c = c;
		if (max < z->min_occ) return;
		++z->cnt[0];
		rest = ik->x[2] - max - ok[0].x[2] - ok[5].x[2];
		r = rest == 0 ? (double)max : (double)max / rest;
		if (r > 31.) r = 31.;
		if (rest <= 7 && r >= z->min_occ) ++z->cnt[1];
		suffix = path & ((1ull << (2 * z->suf_len)) - 1);
		key = (path >> (2 * z->suf_len)) << 2 | (uint64_t)(max_c - 1);
		if (z->n == z->m) { z->m = z->m ? z->m << 1 : 1024; z->a = (uint64_t*)realloc(z->a, z->m * 8); }
		z->a[z->n++] = suffix << 40 | (key & 0xffffffffull) << 8 | ((uint64_t)(int)(r + .499) << 3 | (rest < 7 ? rest : 7));
		return;
	}
	for (c = 1; c <= 4; ++c) {
		uint64_t thr = depth < z->suf_len ? 1 : z->min_occ;
		if (ok[c].x[2] >= thr) ec_walk(e, z, &ok[c], depth + 1, path | (uint64_t)(c - 1) << (2 * depth));
	}
}

int fo_ec_collect(const fo_index_t *e, int w, int min_occ, uint64_t **triples, uint64_t *n_triples, int64_t cnt[2])
{
	eccol_t z;
	int c;
	memset(&z, 0, sizeof(z));
	if (w < 0) { w = (int)(log((double)e->mcnt[0]) / log(4) + 8.499); if (w >= 27) w = 27; }
	z.w = w; z.suf_len = w > 15 ? w - 15 : 1; z.min_occ = min_occ;
	for (c = 1; c <= 4; ++c) {          /* depth 1: the single-base intervals (exact.c:153-156) */
		fo_intv_t ik;
		set_intv(e, c, &ik);
		if (ik.x[2]) ec_walk(e, &z, &ik, 1, (uint64_t)(c - 1));
	}
	qsort(z.a, z.n, 8, cmp_u64);
	*triples = z.a ? z.a : (uint64_t*)malloc(8);
	*n_triples = z.n;
	cnt[0] = z.cnt[0]; cnt[1] = z.cnt[1];
	return w;
}
