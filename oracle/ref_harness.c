/* TEST INFRASTRUCTURE ONLY.
 *
 * Batch entry points over the UNMODIFIED reference (lh3/fermi), linked against the objects
 * compiled from /root/reference by oracle/Makefile into oracle/_ref/libfermi_ref.so.
 * Nothing here re-implements fermi; every function only loops over a reference function:
 *   rld_rank1a / rld_rank2a        rld.c:424,457
 *   fm6_extend                     exact.c:72
 *   fm_backward_search             exact.c:7
 *   fm6_smem                       smem.c:397
 *   fm_retrieve                    exact.c:59
 *   fm6_is_contained / fm6_get_nei unitig.c:77,93
 * It is used (a) to pin oracle/fmd_oracle.c and to generate tests/golden/, (b) as the
 * "reference" CPU baseline in bench.py.  The product never links it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <sys/time.h>
#include "rld.h"
#include "fermi.h"
#include "kvec.h"
#include "kstring.h"

int fm6_is_contained(const rld_t *e, int min_match, const kstring_t *s, fmintv_t *intv, fmintv_v *ovlp);
int fm6_get_nei(const rld_t *e, int min_match, int beg, kstring_t *s, fmintv_v *nei, fmintv_v *prev, fmintv_v *curr,
				fm32s_v *cat, uint64_t *used, const uint64_t *sorted);
void seq_reverse(int l, unsigned char *s);
void seq_revcomp6(int l, unsigned char *s);

static double now_s(void)
{
	struct timeval tv;
	gettimeofday(&tv, 0);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

void *refh_restore(const char *fn) { return rld_restore(fn); }
void refh_destroy(void *e) { rld_destroy((rld_t*)e); }
int refh_dump(const void *e, const char *fn) { return rld_dump((const rld_t*)e, fn); }
void refh_free(void *p) { free(p); }

/* out[0..6]=mcnt, out[7..13]=cnt, out[14]=n_bytes, out[15]=n_frames, out[16]=ibits */
void refh_info(const void *_e, uint64_t out[17])
{
	const rld_t *e = (const rld_t*)_e;
	int i;
	for (i = 0; i < 7; ++i) out[i] = e->mcnt[i], out[7 + i] = e->cnt[i];
	out[14] = e->n_bytes; out[15] = e->n_frames; out[16] = e->ibits;
}

/* s: nt6 text with 0 sentinels (fwd $ rev $ ...), overwritten by its BWT (build.c:33) */
void *refh_build_text(int64_t l, uint8_t *s) { return fm_build(0, 6, 3, l, s); }

void refh_rank1a_batch(const void *e, int64_t n, const uint64_t *k, uint64_t *ok, int32_t *sym)
{
	int64_t i;
	for (i = 0; i < n; ++i) sym[i] = rld_rank1a((const rld_t*)e, k[i], ok + 6 * i);
}

void refh_rank2a_batch(const void *e, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol)
{
	int64_t i;
	for (i = 0; i < n; ++i) rld_rank2a((const rld_t*)e, k[i], l[i], ok + 6 * i, ol + 6 * i);
}

void refh_extend_batch(const void *e, int64_t n, const fmintv_t *ik, const uint8_t *is_back, fmintv_t *ok6)
{
	int64_t i;
	int c;
	for (i = 0; i < n; ++i) {
		fmintv_t *o = ok6 + 6 * i;
		fm6_extend((const rld_t*)e, ik + i, o, is_back[i]);
		for (c = 0; c < 6; ++c) o[c].info = 0; /* fm6_extend leaves info undefined */
	}
}

void refh_backward_search_batch(const void *e, int64_t n, const uint8_t *seq, const uint64_t *off,
								uint64_t *sa_beg, uint64_t *sa_end, uint64_t *size)
{
	int64_t i;
	for (i = 0; i < n; ++i) {
		sa_beg[i] = sa_end[i] = 0;
		size[i] = fm_backward_search((const rld_t*)e, (int)(off[i+1] - off[i]), seq + off[i], &sa_beg[i], &sa_end[i]);
	}
}

typedef struct {
	const rld_t *e;
	int64_t n, start, step;
	const uint8_t *seq;
	const uint64_t *off;
	int self_match;
	fmintv_v out;      /* concatenated records of the reads of this worker, in read order */
	uint32_t *cnt;     /* shared: records per read */
} smem_worker_t;

static void *smem_worker(void *data)
{
	smem_worker_t *w = (smem_worker_t*)data;
	fmintv_v a;
	int64_t i;
	size_t j;
	kv_init(a);
	for (i = w->start; i < w->n; i += w->step) {
		fm6_smem(w->e, (int)(w->off[i+1] - w->off[i]), w->seq + w->off[i], &a, w->self_match);
		w->cnt[i] = a.n;
		for (j = 0; j < a.n; ++j) kv_push(fmintv_t, w->out, a.a[j]);
	}
	free(a.a);
	return 0;
}

/* fm6_smem over n reads with n_threads strided workers (the split of smem.c:346-381).
 * *mem is malloc'd (refh_free); mem_off has n+1 entries. Returns compute wall time in *secs. */
int refh_smem_batch(const void *e, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match, int n_threads,
					fmintv_t **mem, uint64_t *mem_off, double *secs)
{
	smem_worker_t *w;
	pthread_t *tid;
	uint32_t *cnt;
	int64_t i;
	int t;
	double t0;
	size_t *cur;
	if (n_threads < 1) n_threads = 1;
	w = (smem_worker_t*)calloc(n_threads, sizeof(smem_worker_t));
	tid = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
	cnt = (uint32_t*)calloc(n + 1, 4);
	for (t = 0; t < n_threads; ++t) {
		w[t].e = (const rld_t*)e; w[t].n = n; w[t].start = t; w[t].step = n_threads;
		w[t].seq = seq; w[t].off = off; w[t].self_match = self_match; w[t].cnt = cnt;
	}
	t0 = now_s();
	if (n_threads == 1) smem_worker(&w[0]);
	else {
		for (t = 0; t < n_threads; ++t) pthread_create(&tid[t], 0, smem_worker, &w[t]);
		for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
	}
	if (secs) *secs = now_s() - t0;
	for (i = 0, mem_off[0] = 0; i < n; ++i) mem_off[i+1] = mem_off[i] + cnt[i];
	if (mem) {
		*mem = (fmintv_t*)malloc((mem_off[n] ? mem_off[n] : 1) * sizeof(fmintv_t));
		cur = (size_t*)calloc(n_threads, sizeof(size_t));
		for (i = 0; i < n; ++i) {
			t = i % n_threads;
			memcpy(*mem + mem_off[i], w[t].out.a + cur[t], cnt[i] * sizeof(fmintv_t));
			cur[t] += cnt[i];
		}
		free(cur);
	}
	for (t = 0; t < n_threads; ++t) free(w[t].out.a);
	free(w); free(tid); free(cnt);
	return 0;
}

/* fm_retrieve for each sentinel rank x[i]: sequences come back reversed (exact.c:59-70);
 * seq must hold max_len bytes per entry; len[i] receives the length, ret[i] the returned k. */
void refh_retrieve_batch(const void *e, int64_t n, const uint64_t *x, int max_len, uint8_t *seq, int32_t *len, int64_t *ret)
{
	kstring_t s = {0, 0, 0};
	int64_t i;
	for (i = 0; i < n; ++i) {
		ret[i] = fm_retrieve((const rld_t*)e, x[i], &s);
		len[i] = s.l;
		memcpy(seq + (size_t)i * max_len, s.s, s.l < (uint32_t)max_len ? s.l : (uint32_t)max_len);
	}
	free(s.s);
}

/* Per-seed overlap record, exactly the first steps of unitig1 (unitig.c:284-305) without the
 * shared bitmaps: retrieve the read of sentinel rank `seed`, fm6_is_contained, then one
 * fm6_get_nei to the right.  Output per seed:
 *   rec[0]=k returned by fm_retrieve, rec[1]=read length, rec[2]=is_contained ret (0/-1, -9 if too short),
 *   rec[3..5]=intv0.x[0..2], rec[6]=rbeg (or -1), rec[7]=n_nei, rec[8]=extended length of s
 * neighbours (x[0],x[1],x[2],info) are appended to *nei, nei_off has n+1 entries. */
int refh_overlap_batch(const void *_e, int min_match, int64_t n, const uint64_t *seeds, int64_t *rec,
					   fmintv_t **nei_out, uint64_t *nei_off)
{
	const rld_t *e = (const rld_t*)_e;
	kstring_t s = {0, 0, 0};
	fmintv_v a[2], nei, all;
	fm32s_v cat;
	int64_t i;
	size_t j;
	kv_init(a[0]); kv_init(a[1]); kv_init(nei); kv_init(all); kv_init(cat);
	nei_off[0] = 0;
	for (i = 0; i < n; ++i) {
		int64_t *r = rec + 9 * i;
		fmintv_t intv0;
		int ret, rbeg;
		memset(r, 0, 9 * sizeof(int64_t));
		r[0] = fm_retrieve(e, seeds[i], &s);
		seq_reverse(s.l, (uint8_t*)s.s);
		r[1] = s.l; r[6] = -1;
		a[0].n = a[1].n = nei.n = 0;
		if ((int)s.l <= min_match) { r[2] = -9; nei_off[i+1] = all.n; continue; }
		ret = fm6_is_contained(e, min_match, &s, &intv0, &a[0]);
		r[2] = ret; r[3] = intv0.x[0]; r[4] = intv0.x[1]; r[5] = intv0.x[2];
		if (ret >= 0 && a[0].n) {
			rbeg = fm6_get_nei(e, min_match, 0, &s, &nei, &a[0], &a[1], &cat, 0, 0);
			r[6] = rbeg; r[7] = nei.n; r[8] = s.l;
			for (j = 0; j < nei.n; ++j) kv_push(fmintv_t, all, nei.a[j]);
		}
		nei_off[i+1] = all.n;
	}
	*nei_out = all.a ? all.a : (fmintv_t*)malloc(1);
	free(s.s); free(a[0].a); free(a[1].a); free(nei.a); free(cat.a);
	return 0;
}

/* fm6_contrast (cmp.c:94-126) of the reference on two loaded indexes: sub0 / sub1 receive (mcnt[1] + 63) / 64 words each */
void fm6_contrast(rld_t *const e[2], int k, int min_occ, int n_threads, uint64_t *sub[2]);
int refh_contrast(const void *_e0, const void *_e1, int k, int min_occ, int n_threads, uint64_t *sub0, uint64_t *sub1)
{
	rld_t *e[2];
	uint64_t *sub[2];
	e[0] = (rld_t*)_e0; e[1] = (rld_t*)_e1;
	fm6_contrast(e, k, min_occ, n_threads, sub);
	memcpy(sub0, sub[0], (e[0]->mcnt[1] + 63) / 64 * 8);
	memcpy(sub1, sub[1], (e[1]->mcnt[1] + 63) / 64 * 8);
	free(sub[0]); free(sub[1]);
	return 0;
}

/* fm_compute_gap_bits (merge.c:68-94) of the reference: bits receives (n0 + n1 + 63) / 64 words */
uint64_t *fm_compute_gap_bits(const rld_t *e0, const rld_t *e1, int n_threads);
int refh_gap_bits(const void *_e0, const void *_e1, int n_threads, uint64_t *bits)
{
	const rld_t *e0 = (const rld_t*)_e0, *e1 = (const rld_t*)_e1;
	uint64_t *b = fm_compute_gap_bits(e0, e1, n_threads);
	memcpy(bits, b, (e0->mcnt[0] + e1->mcnt[0] + 63) / 64 * 8);
	free(b);
	return 0;
}
