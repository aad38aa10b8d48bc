"""fermi_b200 -- B200 (sm_100a) drop-in for the FMD-index hot path of lh3/fermi.

The product is libfermi_b200.so (C-ABI in include/fermi_b200.h, sources in fermi_b200/csrc/);
this package is the thin host-side mirror of the reference's interface used by tests and bench.
"""
from .api import (INTV, Bcr, Fmd, FmdIndex, SmemSession, overlap_stats, release_cache, fm6_seqsort, ec_kmer_length, fm6_ec_collect, fm6_extend, fm6_overlap, fm6_smem, fm6_unitig, fm6_unitig_assemble, fm6_smem_raw, fm6_smem_raw16, rld_rank1a, check_rank, fm_merge, fm_gap_bits, fm6_contrast, RldIndex, intv16_expand, fm_backward_search,
                  fm_build, fm_build_bwt, fm_ropebwt, fmd_text, launch_count, rld_rank2a, synth_genome, synth_reads)

__all__ = ["INTV", "Bcr", "Fmd", "fm_ropebwt", "FmdIndex", "SmemSession", "fm6_ec_collect", "fm6_extend", "fm6_overlap", "fm6_smem", "fm6_unitig", "fm6_unitig_assemble", "fm6_smem_raw", "fm_backward_search",
           "fm_build", "fm_build_bwt", "fmd_text", "launch_count", "rld_rank2a", "synth_genome", "synth_reads"]
