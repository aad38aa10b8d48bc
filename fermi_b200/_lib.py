"""ctypes binding of libfermi_b200.so (the C-ABI declared in include/fermi_b200.h).

The library is the product; this module only loads it and declares the prototypes.  There is no
Python/NumPy implementation of any query behind it: if the shared library is missing or cannot be
loaded, importing a query fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libfermi_b200.so")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
vpp = C.POINTER(C.c_void_p)

_lib = None

_PROTOS = {
    # host container
    "fmg_fmd_restore": (C.c_void_p, [C.c_char_p]),
    "fmg_fmd_from_bwt": (C.c_void_p, [C.c_int64, u8p]),
    "fmg_fmd_from_rle6": (C.c_void_p, [C.c_int64, u8p]),
    "fmg_fmd_dump": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fmg_fmd_destroy": (None, [C.c_void_p]),
    "fmg_fmd_info": (None, [C.c_void_p, u64p]),
    "fmg_fmd_decode_bwt": (C.c_int64, [C.c_void_p, u8p]),
    # device index
    "fmg_index_upload": (C.c_void_p, [C.c_void_p, C.c_int]),
    "fmg_index_free": (None, [C.c_void_p]),
    "fmg_index_bytes": (C.c_uint64, [C.c_void_p]),
    "fmg_index_device": (C.c_int, [C.c_void_p]),
    "fmg_index_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64p]),
    # rank on the .fmd stream itself
    "fmg_rldx_upload": (C.c_void_p, [C.c_void_p, C.c_int]),
    "fmg_rldx_free": (None, [C.c_void_p]),
    "fmg_rldx_bytes": (C.c_uint64, [C.c_void_p]),
    "fmg_rldx_rank2a_batch": (C.c_int, [C.c_void_p, C.c_int64, u64p, u64p, u64p, u64p]),
    "fmg_rldx_extend_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, u8p, C.c_void_p]),
    # batched queries
    "fmg_rank2a_batch": (C.c_int, [C.c_void_p, C.c_int64, u64p, u64p, u64p, u64p]),
    "fmg_rank1a_batch": (C.c_int, [C.c_void_p, C.c_int64, u64p, u64p, C.c_void_p]),
    "fmg_check_rank": (C.c_int, [C.c_void_p, u64p, u64p]),
    "fmg_extend_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, u8p, C.c_void_p]),
    "fmg_backward_search_batch": (C.c_int, [C.c_void_p, C.c_int64, u8p, u64p, u64p, u64p, u64p]),
    "fmg_smem_batch": (C.c_int, [C.c_void_p, C.c_int64, u8p, u64p, C.c_int, vpp, u64p]),
    "fmg_smem_batch_into": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64,
                                      C.c_void_p, u64p, C.c_int64]),
    "fmg_smem_batch_into16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64,
                                        C.c_void_p, u64p, C.c_int64]),
    "fmg_intv16_expand": (None, [C.c_uint64, C.c_void_p, C.c_void_p]),
    "fmg_free": (None, [C.c_void_p]),
    # device-resident session
    "fmg_smem_session_create": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int]),
    "fmg_smem_session_destroy": (None, [C.c_void_p]),
    "fmg_smem_session_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fmg_smem_session_result": (C.c_int, [C.c_void_p, u64p, vpp, vpp]),
    "fmg_smem_session_set_timing": (None, [C.c_void_p, C.c_int]),
    "fmg_smem_session_kernel_ms": (C.c_double, [C.c_void_p, C.POINTER(C.c_int)]),
    "fmg_release_cache": (None, []),
    "fmg_launch_count": (C.c_uint64, []),
    # overlap / unitig
    "fmg_overlap_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, u64p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, vpp, u64p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "fmg_unitig_assemble": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, u64p, C.c_void_p, C.c_void_p, C.c_char_p, u64p]),
    "fmg_unitig": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, u64p]),
    "fmg_overlap_stats": (None, [C.POINTER(C.c_double)]),
    "fmg_seqsort": (C.c_int, [C.c_void_p, u64p, C.POINTER(C.c_int64)]),
    "fmg_overlap_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                    C.c_void_p, C.c_uint64, u64p]),
    "fmg_overlap_merge": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fmg_overlap_left_fix": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "fmg_overlap_left_fix_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "fmg_overlap_left_flags": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]),
    "fmg_unitig_part": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fmg_magpart_write": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64]),
    "fmg_magpart_free": (None, [C.c_void_p]),
    "fmg_unitig_from_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                         C.c_char_p, u64p]),
    "fmg_contrast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, u64p, u64p]),
    "fmg_gap_bits": (C.c_int, [C.c_void_p, C.c_void_p, u64p]),
    "fmg_merge": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int]),
    "fmg_ec_collect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, vpp, u64p, C.POINTER(C.c_int64)]),
    "fmg_ec_collect_part": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, vpp, u64p, C.POINTER(C.c_int64)]),
    "fmg_ec_kmer_length": (C.c_int, [C.c_uint64]),
    # construction + synthetic data
    "fmg_build_bwt": (C.c_int, [C.c_int, C.c_int64, u8p, u8p]),
    "fmg_build_fmd": (C.c_void_p, [C.c_int, C.c_int64, u8p]),
    "fmg_fmd_from_bwt_device": (C.c_void_p, [C.c_int, C.c_int64, u8p]),
    "fmg_bcr_fmd": (C.c_void_p, [C.c_void_p]),
    "fmg_bcr_init": (C.c_void_p, [C.c_int]),
    "fmg_bcr_append": (C.c_int, [C.c_void_p, C.c_int, u8p]),
    "fmg_bcr_append_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, u8p]),
    "fmg_bcr_build": (C.c_int, [C.c_void_p]),
    "fmg_bcr_size": (C.c_int64, [C.c_void_p]),
    "fmg_bcr_bwt": (C.c_int, [C.c_void_p, u8p]),
    "fmg_bcr_rle": (C.c_int, [C.c_void_p, vpp, C.POINTER(C.c_int64)]),
    "fmg_bcr_destroy": (None, [C.c_void_p]),
    "fmg_synth_genome": (None, [C.c_uint64, C.c_int64, u8p]),
    "fmg_synth_reads": (None, [C.c_uint64, C.c_int64, u8p, C.c_int64, C.c_int, C.c_double, u8p]),
    "fmg_fmd_text": (C.c_int64, [C.c_int64, C.c_int, u8p, u8p]),
}


def lib():
    """The loaded shared library; raises if it has not been built (python -m fermi_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libfermi_b200.so is missing (%s): build it with `python -m fermi_b200.build`; "
                "fermi_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_PROTOS)
