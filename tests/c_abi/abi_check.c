/* Plain C99 client of include/fermi_b200.h: proves that the header is valid C, that every declared entry point links, and
 * exercises the host-only part of the C-ABI (container build / info / decode / dump) -- what fermi's own C mains would do.
 * Built and run by tests/test_host_logic.py with gcc; no GPU needed.  With a second argument "gpu" it also drives the device
 * entry points the way a C main would (upload, fmg_rank1a_batch against counts taken from the BWT here, fmg_check_rank,
 * fmg_backward_search_batch) and fails if no device answers: tests/test_gpu_parity.py runs that on the GPU box. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fermi_b200.h"

/* taking the address of every symbol makes the link fail if the library does not export one of them */
typedef void (*anyfn_t)(void);
static const anyfn_t all_symbols[] = {
	(anyfn_t)fmg_fmd_restore, (anyfn_t)fmg_fmd_from_bwt, (anyfn_t)fmg_fmd_from_rle6, (anyfn_t)fmg_fmd_from_rld,
	(anyfn_t)fmg_fmd_dump, (anyfn_t)fmg_fmd_destroy, (anyfn_t)fmg_fmd_info, (anyfn_t)fmg_fmd_decode_bwt,
	(anyfn_t)fmg_index_upload, (anyfn_t)fmg_index_free, (anyfn_t)fmg_index_bytes, (anyfn_t)fmg_index_device, (anyfn_t)fmg_index_export,
	(anyfn_t)fmg_rank2a_batch, (anyfn_t)fmg_extend_batch, (anyfn_t)fmg_backward_search_batch, (anyfn_t)fmg_smem_batch,
	(anyfn_t)fmg_smem_batch_into, (anyfn_t)fmg_free, (anyfn_t)fmg_smem_session_create, (anyfn_t)fmg_smem_session_destroy,
	(anyfn_t)fmg_smem_session_run, (anyfn_t)fmg_smem_session_result, (anyfn_t)fmg_smem_session_set_timing, (anyfn_t)fmg_smem_session_kernel_ms,
	(anyfn_t)fmg_release_cache, (anyfn_t)fmg_launch_count, (anyfn_t)fmg_overlap_batch, (anyfn_t)fmg_unitig_assemble, (anyfn_t)fmg_unitig,
	(anyfn_t)fmg_overlap_shard, (anyfn_t)fmg_rldx_upload, (anyfn_t)fmg_rldx_free, (anyfn_t)fmg_rldx_bytes, (anyfn_t)fmg_rldx_rank2a_batch, (anyfn_t)fmg_rldx_extend_batch, (anyfn_t)fmg_contrast, (anyfn_t)fmg_gap_bits, (anyfn_t)fmg_merge, (anyfn_t)fmg_rank1a_batch, (anyfn_t)fmg_check_rank, (anyfn_t)fmg_smem_batch_into16, (anyfn_t)fmg_intv16_expand, (anyfn_t)fmg_overlap_merge, (anyfn_t)fmg_overlap_left_fix, (anyfn_t)fmg_overlap_left_fix_rows, (anyfn_t)fmg_overlap_left_flags, (anyfn_t)fmg_unitig_part, (anyfn_t)fmg_magpart_write, (anyfn_t)fmg_magpart_free, (anyfn_t)fmg_unitig_from_device, (anyfn_t)fmg_seqsort, (anyfn_t)fmg_overlap_stats,
	(anyfn_t)fmg_ec_collect, (anyfn_t)fmg_ec_collect_part, (anyfn_t)fmg_ec_kmer_length, (anyfn_t)fmg_build_bwt, (anyfn_t)fmg_build_fmd, (anyfn_t)fmg_fmd_from_bwt_device,
	(anyfn_t)fmg_bcr_init, (anyfn_t)fmg_bcr_append, (anyfn_t)fmg_bcr_append_batch, (anyfn_t)fmg_bcr_build, (anyfn_t)fmg_bcr_size,
	(anyfn_t)fmg_bcr_bwt, (anyfn_t)fmg_bcr_rle, (anyfn_t)fmg_bcr_fmd, (anyfn_t)fmg_bcr_destroy,
	(anyfn_t)fmg_synth_genome, (anyfn_t)fmg_synth_reads, (anyfn_t)fmg_fmd_text
};

int main(int argc, char *argv[])
{
	/* BWT of "ACGT$" + "ACGT$" style toy text is not needed: any nt6 string encodes and decodes */
	const uint8_t bwt[] = {4,4,0,0,1,1,1,2,2,3,3,3,3,3,5,1,0,2};
	uint8_t back[sizeof(bwt)];
	uint64_t info[17];
	fmg_fmd_t *e;
	size_t i, n_sym = sizeof(all_symbols) / sizeof(all_symbols[0]);
	for (i = 0; i < n_sym; ++i) if (all_symbols[i] == 0) return 2;
	fmg_verbose = 0;
	e = fmg_fmd_from_bwt((int64_t)sizeof(bwt), bwt);
	if (e == 0) return 3;
	fmg_fmd_info(e, info);
	if (info[0] != sizeof(bwt) || info[1] != 3 || info[2] != 4 || info[6] != 1) return 4;       /* mcnt: total, $, A, ..., N */
	if (fmg_fmd_decode_bwt(e, back) != (int64_t)sizeof(bwt) || memcmp(back, bwt, sizeof(bwt)) != 0) return 5;
	if (argc > 1 && fmg_fmd_dump(e, argv[1]) != 0) return 6;
	if (argc > 2 && strcmp(argv[2], "gpu") == 0) {
		fmg_index_t *idx = fmg_index_upload(e, 0);
		uint64_t k[sizeof(bwt)], ok[6 * sizeof(bwt)], cnt[6] = {0, 0, 0, 0, 0, 0}, n_bad = 1, first_bad = 0;
		int32_t sym[sizeof(bwt)];
		/* the read "3" (G) as nt6: its interval is the G bucket, C(G) = #$ + #A + #C, size = #G */
		const uint8_t q[1] = {3};
		const uint64_t off[2] = {0, 1};
		uint64_t sa_beg = 0, sa_end = 0, size = 0;
		int c;
		if (idx == 0) return 20;
		if (fmg_index_device(idx) != 0 || fmg_index_bytes(idx) == 0) return 21;
		for (i = 0; i < sizeof(bwt); ++i) k[i] = i;
		if (fmg_rank1a_batch(idx, (int64_t)sizeof(bwt), k, ok, sym) != 0) return 22;
		for (i = 0; i < sizeof(bwt); ++i) {             /* rld_rank1a: counts of BWT[0..k] and the symbol at k */
			++cnt[bwt[i]];
			if (sym[i] != bwt[i]) return 23;
			for (c = 0; c < 6; ++c) if (ok[6 * i + c] != cnt[c]) return 24;
		}
		if (fmg_check_rank(idx, &n_bad, &first_bad) != 0 || n_bad != 0) return 25;
		if (fmg_backward_search_batch(idx, 1, q, off, &sa_beg, &sa_end, &size) != 0) return 26;
		if (size != cnt[3] || sa_beg != cnt[0] + cnt[1] + cnt[2]) return 27;
		fmg_index_free(idx);
		printf("gpu ok: %d ranks, %ld kernel launches\n", (int)sizeof(bwt), (long)fmg_launch_count());
	}
	fmg_fmd_destroy(e);
	if (argc > 1) {                                  /* the file reads back to the same container */
		e = fmg_fmd_restore(argv[1]);
		if (e == 0) return 7;
		if (fmg_fmd_decode_bwt(e, back) != (int64_t)sizeof(bwt) || memcmp(back, bwt, sizeof(bwt)) != 0) return 8;
		fmg_fmd_destroy(e);
	}
	/* no CUDA device here (or one): the device entry points must fail loudly, never fall back to a CPU path */
	printf("ok %d symbols, fmg_ec_kmer_length(3e9)=%d, version %s\n", (int)n_sym, fmg_ec_kmer_length(3000000000ull), FMG_VERSION);
	return 0;
}
