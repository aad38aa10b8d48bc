"""Host-side mirror of the reference's interface for the FMD-index hot path, over libfermi_b200.so.

Names follow lh3/fermi (rld.h:45-58, fermi.h:61-104) so that the parity tests read like calls into
the reference; every function is batched (NumPy arrays in, NumPy arrays out) and executes on the GPU
through the C-ABI.  There is no CPU path here.
"""
import ctypes as C

import numpy as np

from ._lib import lib, u8p, u64p

INTV = np.dtype([("x0", "<u8"), ("x1", "<u8"), ("x2", "<u8"), ("info", "<u8")])   # fmintv_t, fermi.h:13-16


def _p(a, t):
    return a.ctypes.data_as(t)


class Fmd:
    """Host .fmd container: rld_t (rld.h:20-39)."""

    def __init__(self, handle):
        if not handle:
            raise IOError("fermi_b200: could not create the .fmd container")
        self.h = handle
        o = np.zeros(17, np.uint64)
        lib().fmg_fmd_info(self.h, _p(o, u64p))
        self.mcnt = o[0:7].copy()
        self.cnt = o[7:14].copy()
        self.n_bytes, self.n_frames, self.ibits = int(o[14]), int(o[15]), int(o[16])

    @classmethod
    def restore(cls, fn):
        """rld_restore (rld.c:288): accepts "RLD\\2" files and raw "RLE\\6" streams."""
        return cls(lib().fmg_fmd_restore(str(fn).encode()))

    @classmethod
    def from_bwt(cls, bwt):
        """fm_bwtenc (build.c:11)."""
        bwt = np.ascontiguousarray(bwt, np.uint8)
        return cls(lib().fmg_fmd_from_bwt(len(bwt), _p(bwt, u8p)))

    @classmethod
    def from_bwt_device(cls, bwt, device=0):
        """fm_bwtenc (build.c:11) with the RLD encoder on the GPU (rld_enc.cu); byte-identical to from_bwt."""
        bwt = np.ascontiguousarray(bwt, np.uint8)
        return cls(lib().fmg_fmd_from_bwt_device(device, len(bwt), _p(bwt, u8p)))

    @classmethod
    def from_rle6(cls, rle):
        rle = np.ascontiguousarray(rle, np.uint8)
        return cls(lib().fmg_fmd_from_rle6(len(rle), _p(rle, u8p)))

    def dump(self, fn):
        """rld_dump (rld.c:242)."""
        if lib().fmg_fmd_dump(self.h, str(fn).encode()) != 0:
            raise IOError("fermi_b200: cannot write " + str(fn))

    def decode_bwt(self):
        out = np.zeros(int(self.mcnt[0]), np.uint8)
        n = lib().fmg_fmd_decode_bwt(self.h, _p(out, u8p))
        assert n == len(out)
        return out

    def close(self):
        if self.h:
            lib().fmg_fmd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FmdIndex:
    """The index resident in the HBM of one GPU (replaces holding an rld_t for queries)."""

    def __init__(self, fmd, device=0):
        self.fmd = fmd
        self.h = lib().fmg_index_upload(fmd.h, device)
        if not self.h:
            raise RuntimeError("fermi_b200: index upload failed (no CUDA device? see stderr); there is no CPU fallback")
        self.device = device
        self.mcnt, self.cnt = fmd.mcnt, fmd.cnt

    @property
    def nbytes(self):
        return int(lib().fmg_index_bytes(self.h))

    def export(self):
        """(blocks u32[n_blocks,16], cs u64[n_super,8]): the occ-block layout as it sits in HBM."""
        nb, ns = C.c_uint64(), C.c_uint64()
        _check(lib().fmg_index_export(self.h, None, None, C.byref(nb), C.byref(ns)), "index export")
        blocks = np.zeros((nb.value, 16), np.uint32)
        cs = np.zeros((ns.value, 8), np.uint64)
        _check(lib().fmg_index_export(self.h, blocks.ctypes.data, cs.ctypes.data, None, None), "index export")
        return blocks, cs

    def close(self):
        if self.h:
            lib().fmg_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("fermi_b200: %s failed (rc=%d), see stderr" % (what, rc))


def rld_rank2a(idx, k, l):
    """n x rld_rank2a (rld.c:457): returns (ok[n,6], ol[n,6])."""
    k = np.ascontiguousarray(k, np.uint64)
    l = np.ascontiguousarray(l, np.uint64)
    ok = np.zeros((len(k), 6), np.uint64)
    ol = np.zeros((len(k), 6), np.uint64)
    _check(lib().fmg_rank2a_batch(idx.h, len(k), _p(k, u64p), _p(l, u64p), _p(ok, u64p), _p(ol, u64p)), "rld_rank2a")
    return ok, ol


def fm6_extend(idx, ik, is_back):
    """n x fm6_extend (exact.c:72): ik INTV[n] -> ok INTV[n,6] (info = 0)."""
    ik = np.ascontiguousarray(ik, INTV)
    is_back = np.ascontiguousarray(is_back, np.uint8)
    ok = np.zeros((len(ik), 6), INTV)
    _check(lib().fmg_extend_batch(idx.h, len(ik), ik.ctypes.data, _p(is_back, u8p), ok.ctypes.data), "fm6_extend")
    return ok


def fm_backward_search(idx, seq, off):
    """n x fm_backward_search (exact.c:7): returns (sa_beg, sa_end, size); size 0 = no match."""
    seq = np.ascontiguousarray(seq, np.uint8)
    off = np.ascontiguousarray(off, np.uint64)
    n = len(off) - 1
    b, e, s = (np.zeros(n, np.uint64) for _ in range(3))
    _check(lib().fmg_backward_search_batch(idx.h, n, _p(seq, u8p), _p(off, u64p), _p(b, u64p), _p(e, u64p), _p(s, u64p)),
           "fm_backward_search")
    return b, e, s


def fm6_smem(idx, seq, off, self_match=0):
    """n x fm6_smem (smem.c:397): returns (records INTV[], mem_off u64[n+1])."""
    seq = np.ascontiguousarray(seq, np.uint8)
    off = np.ascontiguousarray(off, np.uint64)
    n = len(off) - 1
    mo = np.zeros(n + 1, np.uint64)
    mem = C.c_void_p()
    _check(lib().fmg_smem_batch(idx.h, n, _p(seq, u8p), _p(off, u64p), int(self_match), C.byref(mem), _p(mo, u64p)), "fm6_smem")
    tot = int(mo[-1])
    rec = np.frombuffer(C.string_at(mem.value, tot * 32), dtype=INTV).copy() if tot else np.zeros(0, INTV)
    lib().fmg_free(mem)
    return rec, mo


def fm6_smem_raw(idx, n, seq_ptr, off_ptr, mem_ptr, mem_cap, mem_off_ptr, self_match=0, batch_reads=0):
    """fmg_smem_batch_into on raw host pointers (pinned torch tensors in bench.py). Returns n_records."""
    got = C.c_uint64()
    rc = lib().fmg_smem_batch_into(idx.h, n, seq_ptr, off_ptr, int(self_match), mem_ptr, mem_cap, mem_off_ptr,
                                   C.byref(got), batch_reads)
    if rc == 1:
        raise RuntimeError("fermi_b200: record buffer too small (%d needed)" % got.value)
    _check(rc, "fm6_smem")
    return got.value


class RldIndex:
    """The .fmd stream itself in HBM + a dense block directory (fmg_rldx_*): rank by one warp per query (rldx.cu)."""

    def __init__(self, fmd, device=0):
        self.fmd = fmd
        self.h = lib().fmg_rldx_upload(fmd.h, device)
        if not self.h:
            raise RuntimeError("fermi_b200: RLD index upload failed (no CUDA device? see stderr); there is no CPU fallback")

    @property
    def nbytes(self):
        return int(lib().fmg_rldx_bytes(self.h))

    def rank2a(self, k, l):
        k = np.ascontiguousarray(k, np.uint64)
        l = np.ascontiguousarray(l, np.uint64)
        ok = np.zeros((len(k), 6), np.uint64)
        ol = np.zeros((len(k), 6), np.uint64)
        _check(lib().fmg_rldx_rank2a_batch(self.h, len(k), _p(k, u64p), _p(l, u64p), _p(ok, u64p), _p(ol, u64p)), "rld_rank2a (RLD index)")
        return ok, ol

    def extend(self, ik, is_back):
        ik = np.ascontiguousarray(ik, INTV)
        is_back = np.ascontiguousarray(is_back, np.uint8)
        out = np.zeros((len(ik), 6), INTV)
        _check(lib().fmg_rldx_extend_batch(self.h, len(ik), ik.ctypes.data, _p(is_back, u8p), out.ctypes.data), "fm6_extend (RLD index)")
        return out

    def close(self):
        if self.h:
            lib().fmg_rldx_free(self.h)
            self.h = None


def fm6_contrast(idx0, idx1, k, min_occ):
    """fm6_contrast (cmp.c:94-126): (sub0, sub1) bitmaps over the sequence ranks of the two indexes (uint64 words)"""
    s0 = np.zeros((int(idx0.mcnt[1]) + 63) // 64, np.uint64)
    s1 = np.zeros((int(idx1.mcnt[1]) + 63) // 64, np.uint64)
    _check(lib().fmg_contrast(idx0.h, idx1.h, int(k), int(min_occ), _p(s0, u64p), _p(s1, u64p)), "fm6_contrast")
    return s0, s1


def fm_merge(fmd0, fmd1, device=0):
    """fm_merge (merge.c:100-137) / `fermi merge`: the gap vector, the interleaving and the RLD encoding on the GPU"""
    return Fmd(lib().fmg_merge(fmd0.h, fmd1.h, device))


def fm_gap_bits(idx0, idx1):
    """fm_compute_gap_bits (merge.c:68-94): uint64 words, bit q set when symbol q of the merged BWT comes from idx1"""
    n = int(idx0.mcnt[0]) + int(idx1.mcnt[0])
    bits = np.zeros((n + 63) // 64, np.uint64)
    _check(lib().fmg_gap_bits(idx0.h, idx1.h, _p(bits, u64p)), "fm_compute_gap_bits")
    return bits


def rld_rank1a(idx, k):
    """rld_rank1a (rld.c:424-446) for an array of positions: (ok[n,6], symbol[n]); k = 2^64-1 gives zeros and -1"""
    k = np.ascontiguousarray(k, np.uint64)
    ok = np.zeros((len(k), 6), np.uint64)
    sym = np.zeros(len(k), np.int32)
    _check(lib().fmg_rank1a_batch(idx.h, len(k), _p(k, u64p), _p(ok, u64p), sym.ctypes.data), "rld_rank1a")
    return ok, sym


def check_rank(idx):
    """`fermi chkbwt -r` on the device: (number of positions where the rank function disagrees with the BWT, first such position)"""
    bad, first = C.c_uint64(), C.c_uint64()
    _check(lib().fmg_check_rank(idx.h, C.byref(bad), C.byref(first)), "chkbwt -r")
    return int(bad.value), int(first.value)


def fm6_smem_raw16(idx, n, seq_ptr, off_ptr, mem_ptr, mem_cap, mem_off_ptr, self_match=0, batch_reads=0):
    """fmg_smem_batch_into16 on raw host pointers: packed 16-byte records (x0, x1, x2, end | start << 16 | closed << 31)."""
    got = C.c_uint64()
    rc = lib().fmg_smem_batch_into16(idx.h, n, seq_ptr, off_ptr, int(self_match), mem_ptr, mem_cap, mem_off_ptr,
                                     C.byref(got), batch_reads)
    if rc == 1:
        raise RuntimeError("fermi_b200: record buffer too small (%d needed)" % got.value)
    _check(rc, "fm6_smem (packed)")
    return got.value


def intv16_expand(packed):
    """uint32[n,4] packed records -> INTV[n] (fmg_intv16_expand)"""
    packed = np.ascontiguousarray(packed, np.uint32).reshape(-1, 4)
    out = np.zeros(len(packed), INTV)
    lib().fmg_intv16_expand(len(packed), packed.ctypes.data, out.ctypes.data)
    return out


class SmemSession:
    """Device-resident SMEM session (fmg_smem_session_*): reads and results stay in HBM."""

    def __init__(self, idx, max_reads, max_len):
        self.idx = idx
        self.h = lib().fmg_smem_session_create(idx.h, max_reads, max_len)
        if not self.h:
            raise RuntimeError("fermi_b200: cannot create the SMEM session")

    def run(self, n, d_seq_ptr, d_off_ptr, self_match=0, stream=0):
        _check(lib().fmg_smem_session_run(self.h, n, d_seq_ptr, d_off_ptr, int(self_match), stream), "smem session run")

    def result(self):
        """(n_records, device pointer of records, device pointer of offsets); synchronises the stream."""
        n = C.c_uint64()
        mem = C.c_void_p()
        off = C.c_void_p()
        _check(lib().fmg_smem_session_result(self.h, C.byref(n), C.byref(mem), C.byref(off)), "smem session result")
        return n.value, mem.value, off.value

    def set_timing(self, on=True):
        lib().fmg_smem_session_set_timing(self.h, 1 if on else 0)

    def kernel_ms(self):
        """(summed k_smem milliseconds, number of launches) since the last call (CUDA events)."""
        n = C.c_int()
        ms = lib().fmg_smem_session_kernel_ms(self.h, C.byref(n))
        return ms, n.value

    def close(self):
        if self.h:
            lib().fmg_smem_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fm6_overlap(idx, min_match, ids=None, first=0, step=1, n=None, max_len=128):
    """Per-sequence overlap records (fm_retrieve + fm6_is_contained + fm6_get_nei + check_left_simple,
    unitig.c:77-204).  Returns dict(rec[n,10], nei INTV[], nei_off[n+1], seq[n,max_len], len[n], ext[n,max_len])."""
    if ids is not None:
        ids = np.ascontiguousarray(ids, np.uint64)
        n = len(ids)
    rec = np.zeros((n, 10), np.int64)
    off = np.zeros(n + 1, np.uint64)
    seq = np.zeros((n, max_len), np.uint8)
    ext = np.zeros((n, max_len), np.uint8)
    ln = np.zeros(n, np.int32)
    nei = C.c_void_p()
    rc = lib().fmg_overlap_batch(idx.h, int(min_match), n, _p(ids, u64p) if ids is not None else None, first, step, max_len,
                                 rec.ctypes.data, C.byref(nei), _p(off, u64p), seq.ctypes.data, ln.ctypes.data, ext.ctypes.data)
    _check(rc, "fm6_overlap")
    tot = int(off[-1])
    out = np.frombuffer(C.string_at(nei.value, tot * 32), dtype=INTV).copy() if tot else np.zeros(0, INTV)
    lib().fmg_free(nei)
    return dict(rec=rec, nei=out, nei_off=off, seq=seq, len=ln, ext=ext)


def fm6_unitig_assemble(n_seq, min_match, o, out_path):
    """The unitig walk (unitig.c:227-362) over the overlap records `o` of all n_seq sequences; writes MAG text."""
    n = C.c_uint64()
    nei = np.ascontiguousarray(o["nei"], INTV)
    if len(nei) == 0:
        nei = np.zeros(1, INTV)
    _check(lib().fmg_unitig_assemble(n_seq, o["seq"].shape[1], int(min_match), o["rec"].ctypes.data, nei.ctypes.data,
                                     _p(np.ascontiguousarray(o["nei_off"], np.uint64), u64p), o["seq"].ctypes.data, o["ext"].ctypes.data,
                                     str(out_path).encode(), C.byref(n)), "fm6_unitig_assemble")
    return n.value


def fm6_unitig(idx, min_match, out_path, max_len=0):
    """fm6_unitig (unitig.c:378) / `fermi unitig -l`: MAG records to out_path; returns the number of unitigs."""
    n = C.c_uint64()
    _check(lib().fmg_unitig(idx.h, int(min_match), int(max_len), str(out_path).encode(), C.byref(n)), "fm6_unitig")
    return n.value


def fm6_seqsort(idx):
    """fm6_seqsort (seqsort.c:37) / `fermi seqrank`: (sorted[mcnt[1]], (zeros, contained, duplicates))."""
    n = int(idx.fmd.mcnt[1])
    out = np.zeros(max(n, 1), np.uint64)
    st = (C.c_int64 * 3)()
    _check(lib().fmg_seqsort(idx.h, _p(out, u64p), st), "fm6_seqsort")
    return out[:n], (int(st[0]), int(st[1]), int(st[2]))


def overlap_stats():
    """kernel milliseconds of the last overlap pass: dict(contained, neighbours, left_fix, left_rows, pack, batches)"""
    ms = (C.c_double * 8)()
    lib().fmg_overlap_stats(ms)
    return {"contained": ms[1], "neighbours": ms[2], "left_fix": ms[3], "left_rows": int(ms[4]), "pack": ms[5], "batches": int(ms[7])}


def fm_build_bwt(text, device=0):
    """BWT of the FMD text on the GPU (replaces fm_bwtgen/ksa_bwt, build.c:5, ksa.c:231)."""
    text = np.ascontiguousarray(text, np.uint8)
    bwt = np.zeros(len(text), np.uint8)
    _check(lib().fmg_build_bwt(device, len(text), _p(text, u8p), _p(bwt, u8p)), "fm_build_bwt")
    return bwt


def fm_build(text, device=0, host_encode=False):
    """fm_build (build.c:33): text -> Fmd.  Suffix sort, BWT and the RLD encoder all run on the GPU (fmg_build_fmd);
    host_encode=True takes the BWT to the host and encodes it there (fm_bwtenc as serial host code)."""
    if host_encode:
        return Fmd.from_bwt(fm_build_bwt(text, device))
    text = np.ascontiguousarray(text, np.uint8)
    return Fmd(lib().fmg_build_fmd(device, len(text), _p(text, u8p)))


def fm6_ec_collect(idx, w=-1, min_occ=3, part=0, n_parts=1):
    """ec_collect (correct.c:35-87) over the whole trie: returns (triples u64[] sorted = suffix<<40|key<<8|val, cnt[2]).
    n_parts > 1: only the subtrees of the suffixes s with s % n_parts == part (one GPU's share)."""
    p = C.c_void_p()
    n = C.c_uint64()
    cnt = (C.c_int64 * 2)()
    _check(lib().fmg_ec_collect_part(idx.h, int(w), int(min_occ), int(part), int(n_parts), C.byref(p), C.byref(n), cnt), "fm6_ec_collect")
    out = np.ctypeslib.as_array(C.cast(p, u64p), shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint64)     # one copy out of the malloc'd result
    lib().fmg_free(p)
    return out, (int(cnt[0]), int(cnt[1]))


def ec_kmer_length(n_symbols):
    """k-mer length fm6_ec_correct picks for an index of n_symbols (correct.c:313-318)."""
    return int(lib().fmg_ec_kmer_length(int(n_symbols)))


class Bcr:
    """bcr_t (bcr.h:43-49): append sequences, build the BWT on the GPU, read it back."""

    def __init__(self, device=0):
        self.h = lib().fmg_bcr_init(device)

    def append(self, seq):
        seq = np.ascontiguousarray(seq, np.uint8)
        _check(lib().fmg_bcr_append(self.h, len(seq), _p(seq, u8p)), "bcr_append")

    def append_batch(self, seqs):
        seqs = np.ascontiguousarray(seqs, np.uint8)
        _check(lib().fmg_bcr_append_batch(self.h, seqs.shape[0], seqs.shape[1], _p(seqs, u8p)), "bcr_append")

    def build(self):
        _check(lib().fmg_bcr_build(self.h), "bcr_build")

    def build_fmd(self):
        """`fermi ropebwt | fermi recode`: BCR and the RLD encoder on the GPU; returns the Fmd (the BWT itself is not copied out)."""
        _check(lib().fmg_bcr_build(self.h), "bcr_build")
        return Fmd(lib().fmg_bcr_fmd(self.h))

    def bwt(self):
        n = lib().fmg_bcr_size(self.h)
        out = np.zeros(n, np.uint8)
        _check(lib().fmg_bcr_bwt(self.h, _p(out, u8p)), "bcr_bwt")
        return out

    def rle(self):
        p = C.c_void_p()
        n = C.c_int64()
        _check(lib().fmg_bcr_rle(self.h, C.byref(p), C.byref(n)), "bcr_rle")
        out = np.frombuffer(C.string_at(p.value, n.value), np.uint8).copy()
        lib().fmg_free(p)
        return out

    def close(self):
        if self.h:
            lib().fmg_bcr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fm_ropebwt(reads, device=0):
    """`fermi ropebwt -a bcr` (ropebwt.c:47-158): every read, then its reverse complement (even-length
    reverse-complement palindromes lose their last base first, ropebwt.c:25-29), through the GPU BCR. Returns the BWT."""
    b = Bcr(device)
    reads = np.ascontiguousarray(reads, np.uint8)
    if reads.ndim == 2 and reads.shape[1] % 2 == 1:            # odd length: no palindromes, one batch of (read, revcomp) pairs
        rc = 5 - reads[:, ::-1]
        both = np.empty((2 * len(reads), reads.shape[1]), np.uint8)
        both[0::2], both[1::2] = reads, rc
        b.append_batch(both)
    else:
        for r in reads:
            r = np.asarray(r, np.uint8)
            l = len(r)
            if l % 2 == 0 and l > 0 and np.all(r[: l // 2] + r[::-1][: l // 2] == 5):
                r = r[:-1]
            b.append(r)
            b.append((5 - r[::-1]).astype(np.uint8))
    b.build()
    out = b.bwt()
    b.close()
    return out


def launch_count():
    return int(lib().fmg_launch_count())


def release_cache():
    """return the device scratch the library keeps between calls to the driver"""
    lib().fmg_release_cache()


# ------------------------------------------------------------------ synthetic data (host helpers)
def synth_genome(seed, n):
    g = np.zeros(n, np.uint8)
    lib().fmg_synth_genome(seed, n, _p(g, u8p))
    return g


def synth_reads(seed, genome, n_reads, length, err, out=None):
    genome = np.ascontiguousarray(genome, np.uint8)
    if out is None:
        out = np.zeros((n_reads, length), np.uint8)
    lib().fmg_synth_reads(seed, len(genome), _p(genome, u8p), n_reads, length, float(err), _p(out, u8p))
    return out


def fmd_text(seqs):
    seqs = np.ascontiguousarray(seqs, np.uint8)
    n, L = seqs.shape
    total = lib().fmg_fmd_text(n, L, _p(seqs, u8p), None)
    text = np.zeros(total, np.uint8)
    lib().fmg_fmd_text(n, L, _p(seqs, u8p), _p(text, u8p))
    return text
