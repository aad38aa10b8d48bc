/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the lh3/fermi FMD-index hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (fermi_b200/) never links, imports or executes it.
 *
 * Parity status: PINNED.  Every function below is checked against the unmodified reference
 * compiled into oracle/_ref/ (tests/test_oracle_vs_ref.py, run where /root/reference exists)
 * and against the golden vectors under tests/golden/ that were generated from that reference
 * by tests/golden/make_golden.py.
 *
 * All citations are file:line of the reference (lh3/fermi @ 1.1-r751-beta).
 */
#ifndef FMD_ORACLE_H
#define FMD_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x[3]; uint64_t info; } fo_intv_t;      /* fermi.h:13-16 */

typedef struct fo_index_s fo_index_t;

/* container (rld.h:20-39), flat in memory instead of 64 MB chunks */
fo_index_t *fo_load(const char *fn);                 /* rld_restore, rld.c:288-325 (RLD\2 and raw RLE\6) */
fo_index_t *fo_from_bwt(int64_t n, const uint8_t *bwt); /* fm_bwtenc, build.c:11-31 */
fo_index_t *fo_from_rle6(int64_t n, const uint8_t *rle); /* rld.c:295-309 */
int         fo_dump(const fo_index_t *e, const char *fn); /* rld_dump, rld.c:242-263 */
void        fo_destroy(fo_index_t *e);
/* out[0..6]=mcnt, out[7..13]=cnt, out[14]=n_bytes, out[15]=n_frames, out[16]=ibits */
void        fo_info(const fo_index_t *e, uint64_t out[17]);
const uint64_t *fo_words(const fo_index_t *e);
const uint64_t *fo_frame(const fo_index_t *e);
/* expand the whole BWT to one nt6 byte per symbol; returns number of symbols */
int64_t     fo_decode_bwt(const fo_index_t *e, uint8_t *out);

/* rank (rld.c:352-492) */
int  fo_rank1a(const fo_index_t *e, uint64_t k, uint64_t ok[6]);
void fo_rank2a(const fo_index_t *e, uint64_t k, uint64_t l, uint64_t ok[6], uint64_t ol[6]);
void fo_rank1a_batch(const fo_index_t *e, int64_t n, const uint64_t *k, uint64_t *ok, int32_t *sym);
void fo_rank2a_batch(const fo_index_t *e, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol);

/* FMD primitives (exact.c:7-88) */
int  fo_extend(const fo_index_t *e, const fo_intv_t *ik, fo_intv_t ok[6], int is_back);
void fo_extend_batch(const fo_index_t *e, int64_t n, const fo_intv_t *ik, const uint8_t *is_back, fo_intv_t *ok6);
uint64_t fo_backward_search(const fo_index_t *e, int len, const uint8_t *s, uint64_t *sa_beg, uint64_t *sa_end);
void fo_backward_search_batch(const fo_index_t *e, int64_t n, const uint8_t *seq, const uint64_t *off,
							  uint64_t *sa_beg, uint64_t *sa_end, uint64_t *size);
int64_t fo_retrieve(const fo_index_t *e, uint64_t x, uint8_t *out, int max_len, int *len); /* exact.c:59-70 */

/* SMEM (smem.c:13-112, 397-410).  fo_smem_batch: strided n_threads workers like smem.c:346-381.
 * *mem is malloc'd (fo_free), mem_off has n+1 entries; *n_locate (optional) receives the number
 * of block lookups (calls of the rld_locate_blk restatement), *n_extend the fm6_extend count. */
int fo_smem_batch(const fo_index_t *e, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match, int n_threads,
				  fo_intv_t **mem, uint64_t *mem_off, double *secs, uint64_t *n_locate, uint64_t *n_extend);
void fo_free(void *p);

/* overlap (unitig.c:38-179): per seed, the record described in oracle/ref_harness.c:refh_overlap_batch */
int fo_overlap_batch(const fo_index_t *e, int min_match, int64_t n, const uint64_t *seeds, int64_t *rec,
					 fo_intv_t **nei_out, uint64_t *nei_off, uint64_t *n_locate);

/* block lookups per stage of the unitig work on one visited sequence (retrieve, is_contained, get_nei, check_left_simple,
 * unitig.c:77-204): the N_locate of SURVEY.md 8d counted on the benchmark input; left (optional, n) = check_left_simple results */
int fo_unitig_locates(const fo_index_t *e, int min_match, int64_t n, const uint64_t *seeds, uint64_t locates[4], int8_t *left);

/* k-mer collection of `fermi correct`: fm6_traverse (exact.c:141-171) + ec_collect (correct.c:35-87) over all
 * suffixes; triples = suffix<<40 | key<<8 | val, sorted; returns the k-mer length used (w<0: correct.c:313-318) */
int fo_ec_collect(const fo_index_t *e, int w, int min_occ, uint64_t **triples, uint64_t *n_triples, int64_t cnt[2]);

#ifdef __cplusplus
}
#endif
#endif
