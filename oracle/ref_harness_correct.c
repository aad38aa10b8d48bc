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
