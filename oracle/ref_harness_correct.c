/* TEST INFRASTRUCTURE ONLY.
 *
 * ec_collect (correct.c:35-87) is static in the reference, so this translation unit compiles the reference's
 * correct.c IN PLACE (#include of the file where it lies under /root/reference, found through -I; nothing is
 * copied into this repo) and adds one entry point that runs the collect phase of fm6_ec_correct
 * (correct.c:305-360) and dumps the hash tables.  It replaces correct.o in libfermi_ref.so.
 */
#include "correct.c"

static int cmp_u64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
	return x < y ? -1 : x > y;
}

/* triples = suffix<<40 | key<<8 | val, sorted; cnt[0..1] as accumulated by ec_collect */
int refh_ec_collect(const void *_e, int w, int min_occ, uint64_t **triples, uint64_t *n_triples, int64_t cnt[2])
{
	const rld_t *e = (const rld_t*)_e;
	fmecopt_t opt;
	fmintv_t *top;
	uint64_t *out = 0, n = 0, m = 0;
	int i;
	memset(&opt, 0, sizeof(opt));
	opt.w = w; opt.min_occ = min_occ;
	if (opt.w < 0) {                         /* correct.c:313-318 */
		opt.w = (int)(log(e->mcnt[0]) / log(4) + 8.499);
		if (opt.w >= MAX_KMER) opt.w = MAX_KMER;
	}
	compute_SUF(opt.w > 15? opt.w - 15 : 1);  /* correct.c:319 */
	top = fm6_traverse(e, SUF_LEN);
	cnt[0] = cnt[1] = 0;
	for (i = 0; i < SUF_NUM; ++i) {
		shash_t *h = kh_init(solid);
		khint_t k;
		ec_collect(e, &opt, SUF_LEN, &top[i], h, cnt);
		for (k = kh_begin(h); k != kh_end(h); ++k) {
			if (!kh_exist(h, k)) continue;
			if (n == m) { m = m? m<<1 : 1024; out = (uint64_t*)realloc(out, m * 8); }
			out[n++] = (uint64_t)i<<40 | (uint64_t)kh_key(h, k)<<8 | kh_val(h, k);
		}
		kh_destroy(solid, h);
	}
	free(top);
	qsort(out, n, 8, cmp_u64);
	*triples = out ? out : (uint64_t*)malloc(8);
	*n_triples = n;
	return opt.w;
}

/* The fix phase of `fermi correct` (correct.c:361-452) from a GIVEN set of k-mer triples: the hash tables solid[suffix] are filled
 * from (suffix, key, val) exactly as ec_collect stores them (correct.c:71-75) and the reference's own worker2 / ec_fix run on the
 * reads of `fq`; the corrected reads are written to `out_path` in the format of correct.c:405-418 (unpaired, defaults of
 * main_correct, cmd.c:257).  The loop below is ours (one thread, batches of BATCH_SIZE like correct.c:389); every decision about a
 * base is the reference's.  tests/: triples of the reference's ec_collect -> byte-identical to `fermi correct -t1` (pins this
 * driver); triples of fmg_ec_collect (GPU) -> byte-identical too (the hand-off of SURVEY.md 8b "Correct"). */
int refh_ec_fix_from_triples(const void *_e, int w, int min_occ, uint64_t n_triples, const uint64_t *triples, const char *fq, const char *out_path)
{
	const rld_t *e = (const rld_t*)_e;
	fmecopt_t opt;
	shash_t **solid;
	worker2_t w2;
	gzFile fp;
	kseq_t *seq;
	FILE *out;
	uint64_t i, id = 0;
	int j, max_seqs, ret;
	opt.w = w; opt.min_occ = min_occ; opt.keep_bad = 0; opt.is_paired = 0; opt.max_corr = 0.3; opt.trim_l = 0; opt.step = 5;
	if (opt.w < 0) {
		opt.w = (int)(log(e->mcnt[0]) / log(4) + 8.499);
		if (opt.w >= MAX_KMER) opt.w = MAX_KMER;
	}
	compute_SUF(opt.w > 15? opt.w - 15 : 1);
	solid = calloc(SUF_NUM, sizeof(void*));
	for (j = 0; j < SUF_NUM; ++j) solid[j] = kh_init(solid);
	for (i = 0; i < n_triples; ++i) {
		khint_t k = kh_put(solid, solid[triples[i]>>40], (uint32_t)(triples[i]>>8), &ret);
		kh_val(solid[triples[i]>>40], k) = (uint8_t)triples[i];
	}
	out = fopen(out_path, "wb");
	fp = gzopen(fq, "r");
	if (out == 0 || fp == 0) return -1;
	seq = kseq_init(fp);
	max_seqs = BATCH_SIZE;
	memset(&w2, 0, sizeof(w2));
	w2.e = e; w2.solid = solid; w2.opt = &opt;
	w2.seq = calloc(max_seqs, sizeof(void*)); w2.qual = calloc(max_seqs, sizeof(void*)); w2.info = calloc(max_seqs, sizeof(int));
	for (;;) {
		ret = kseq_read(seq);
		if (ret < 0 || w2.n_seqs == max_seqs) {
			worker2(&w2);
			for (j = 0; j < w2.n_seqs; ++j, ++id) {
				if (!(w2.info[j]>>16&1) || opt.keep_bad)
					fprintf(out, "@%llu_%d_%d\n%s\n+\n%s\n", (unsigned long long)id, w2.info[j]&0xffff, w2.info[j]>>18, w2.seq[j], w2.qual[j]);
				free(w2.seq[j]); free(w2.qual[j]);
			}
			w2.n_seqs = 0;
		}
		if (ret < 0) break;
		w2.seq[w2.n_seqs] = strdup(seq->seq.s);
		if (seq->qual.l == 0) {
			w2.qual[w2.n_seqs] = malloc(seq->seq.l + 1);
			for (j = 0; j < (int)seq->seq.l; ++j) w2.qual[w2.n_seqs][j] = 33 + 15;
			w2.qual[w2.n_seqs][j] = 0;
		} else w2.qual[w2.n_seqs] = strdup(seq->qual.s);
		++w2.n_seqs;
	}
	free(w2.seq); free(w2.qual); free(w2.info);
	kseq_destroy(seq); gzclose(fp); fclose(out);
	for (j = 0; j < SUF_NUM; ++j) kh_destroy(solid, solid[j]);
	free(solid);
	return opt.w;
}
