"""Test-side helpers: ctypes bindings for the CPU oracle (oracle/liboracle.so), the compiled
reference (oracle/_ref/libfermi_ref.so, when present) and seeded synthetic data.

TEST INFRASTRUCTURE ONLY: nothing in fermi_b200/ imports this module or anything under oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

INTV = np.dtype([("x0", "<u8"), ("x1", "<u8"), ("x2", "<u8"), ("info", "<u8")])

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


# ----------------------------------------------------------------------------- synthetic data
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    """Vectorised splitmix64 finaliser; same constants as fermi_b200/csrc/synth.cpp."""
    z = (np.asarray(z, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _stream(seed, stream, idx):
    with np.errstate(over="ignore"):
        base = splitmix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1B54A32D192ED03)))
        return splitmix64(base + np.asarray(idx, dtype=np.uint64))


def synth_genome(n, seed):
    """i.i.d. uniform ACGT genome, nt6 codes 1..4 (one 64-bit hash per 32 bases)."""
    with np.errstate(over="ignore"):
        h = _stream(seed, 0, np.arange((n + 31) // 32, dtype=np.uint64))
        sh = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
        g = ((h[:, None] >> sh) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n] + 1
    return g


def revcomp(a):
    a = np.asarray(a)
    r = a[..., ::-1].copy()
    m = (r >= 1) & (r <= 4)
    r[m] = 5 - r[m]
    return r


def synth_reads(genome, n_reads, L, err, seed):
    """n_reads x L nt6 reads: uniform start, 50/50 strand, uniform substitutions at rate err."""
    G = len(genome)
    with np.errstate(over="ignore"):
        r = np.arange(n_reads, dtype=np.uint64)
        start = (_stream(seed, 1, r) % np.uint64(G - L + 1)).astype(np.int64)
        strand = (_stream(seed, 2, r) >> np.uint64(17)) & np.uint64(1)
        idx = start[:, None] + np.arange(L, dtype=np.int64)[None, :]
        reads = genome[idx].copy()
        e = _stream(seed, 3, (r[:, None] * np.uint64(L) + np.arange(L, dtype=np.uint64)[None, :]))
        thr = np.uint64(int(err * 4294967296.0))
        hit = (e & np.uint64(0xFFFFFFFF)) < thr
        delta = ((e >> np.uint64(32)) % np.uint64(3)).astype(np.uint8) + 1
        sub = ((reads - 1 + delta) & 3) + 1
        reads[hit] = sub[hit]
    rc = revcomp(reads)
    reads[strand == 1] = rc[strand == 1]
    return np.ascontiguousarray(reads)


def fmd_text(seqs):
    """The text fermi indexes (cmd.c:457-469): r0 $ rc(r0) $ r1 $ rc(r1) $ ..., with the
    "even-length reverse-complement palindromes lose their last base" rule (cmd.c:458-463)."""
    out = []
    for s in seqs:
        s = np.asarray(s, dtype=np.uint8)
        l = len(s)
        if l % 2 == 0 and l > 0 and np.all(s[: l // 2] + s[::-1][: l // 2] == 5):
            s = s[:-1]
        out += [s, np.zeros(1, np.uint8), revcomp(s), np.zeros(1, np.uint8)]
    return np.ascontiguousarray(np.concatenate(out))


def naive_bwt(text):
    """BWT of a multi-sentinel text, sentinels ordered by position (ksa.c:54, 231-242)."""
    n = len(text)
    t = text.astype(np.int64)
    sent = np.flatnonzero(t == 0)
    key = t.copy() + len(sent)          # bases above every sentinel
    key[sent] = np.arange(len(sent))    # sentinels distinct, ordered by position
    rank = key.copy()
    sa = np.argsort(rank, kind="stable")
    h = 1
    while True:
        nxt = np.zeros(n, np.int64)
        nxt[: n - h] = rank[h:] + 1 if h < n else 0
        order = np.lexsort((nxt, rank))
        rk = rank[order]
        nk = nxt[order]
        diff = np.ones(n, bool)
        diff[1:] = (rk[1:] != rk[:-1]) | (nk[1:] != nk[:-1])
        newrank = np.empty(n, np.int64)
        newrank[order] = np.cumsum(diff) - 1
        rank = newrank
        sa = order
        if diff.all():
            break
        h *= 2
    bwt = np.where(sa > 0, text[sa - 1], 0).astype(np.uint8)
    return bwt


def reads_to_flat(reads):
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n, L = reads.shape
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L)).astype(np.uint64)
    return reads.reshape(-1), off


# ----------------------------------------------------------------------------- libraries
def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


class _Lib:
    """Common surface of the oracle port ("fo_") and the reference harness ("refh_")."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.p = prefix
        L = self.lib
        g = lambda n: getattr(L, prefix + n)
        self.kind = "port" if prefix == "fo_" else "reference"
        self._load = g("load") if prefix == "fo_" else g("restore")
        self._load.restype = C.c_void_p
        self._load.argtypes = [C.c_char_p]
        self._destroy = g("destroy")
        self._destroy.argtypes = [C.c_void_p]
        self._info = g("info")
        self._info.argtypes = [C.c_void_p, u64p]
        self._dump = g("dump")
        self._dump.argtypes = [C.c_void_p, C.c_char_p]
        self._dump.restype = C.c_int
        self._free = g("free")
        self._free.argtypes = [C.c_void_p]
        self._rank1a = g("rank1a_batch")
        self._rank1a.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i32p]
        self._rank2a = g("rank2a_batch")
        self._rank2a.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, u64p, u64p]
        self._extend = g("extend_batch")
        self._extend.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, u8p, C.c_void_p]
        self._bsearch = g("backward_search_batch")
        self._bsearch.argtypes = [C.c_void_p, C.c_int64, u8p, u64p, u64p, u64p, u64p]
        self._smem = g("smem_batch")
        if prefix == "fo_":
            self._smem.argtypes = [C.c_void_p, C.c_int64, u8p, u64p, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p), u64p, C.POINTER(C.c_double), u64p, u64p]
            self._overlap = g("overlap_batch")
            self._overlap.argtypes = [C.c_void_p, C.c_int, C.c_int64, u64p, i64p, C.POINTER(C.c_void_p), u64p, u64p]
            L.fo_from_bwt.restype = C.c_void_p
            L.fo_from_bwt.argtypes = [C.c_int64, u8p]
            L.fo_from_rle6.restype = C.c_void_p
            L.fo_from_rle6.argtypes = [C.c_int64, u8p]
            L.fo_decode_bwt.restype = C.c_int64
            L.fo_decode_bwt.argtypes = [C.c_void_p, u8p]
        else:
            self._smem.argtypes = [C.c_void_p, C.c_int64, u8p, u64p, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p), u64p, C.POINTER(C.c_double)]
            self._overlap = g("overlap_batch")
            self._overlap.argtypes = [C.c_void_p, C.c_int, C.c_int64, u64p, i64p, C.POINTER(C.c_void_p), u64p]
            L.refh_build_text.restype = C.c_void_p
            L.refh_build_text.argtypes = [C.c_int64, u8p]

    def ec_collect(self, h, w, min_occ):
        """(triples u64[] sorted, (cnt0, cnt1), w)"""
        f = getattr(self.lib, self.p + "ec_collect")
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), u64p, i64p]
        f.restype = C.c_int
        p = C.c_void_p()
        n = C.c_uint64()
        cnt = np.zeros(2, np.int64)
        w = f(h, w, min_occ, C.byref(p), C.byref(n), _ptr(cnt, i64p))
        out = np.frombuffer(C.string_at(p.value, n.value * 8), np.uint64).copy() if n.value else np.zeros(0, np.uint64)
        self._free(p)
        return out, (int(cnt[0]), int(cnt[1])), w

    def ec_fix_from_triples(self, h, w, min_occ, triples, fq, out_path):
        """reference harness only: the fix phase of `fermi correct` (the reference's worker2 / ec_fix) on the reads of `fq` with the
        hash tables filled from `triples`; corrected FASTQ to out_path; returns the k-mer length used"""
        assert self.p == "refh_"
        f = self.lib.refh_ec_fix_from_triples
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, u64p, C.c_char_p, C.c_char_p]
        t = np.ascontiguousarray(triples, np.uint64)
        return f(h, w, min_occ, len(t), _ptr(t, u64p), fq.encode(), out_path.encode())

    def contrast(self, h0, h1, n_seq0, n_seq1, k, min_occ, n_threads=1):
        """reference harness only: fm6_contrast (cmp.c:94-126) -> (sub0, sub1) uint64 bitmaps over sequence ranks"""
        assert self.p == "refh_"
        f = self.lib.refh_contrast
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, u64p, u64p]
        s0, s1 = np.zeros((n_seq0 + 63) // 64, np.uint64), np.zeros((n_seq1 + 63) // 64, np.uint64)
        f(h0, h1, k, min_occ, n_threads, _ptr(s0, u64p), _ptr(s1, u64p))
        return s0, s1

    def gap_bits(self, h0, h1, n_total, n_threads=1):
        """reference harness only: fm_compute_gap_bits (merge.c:68-94)"""
        assert self.p == "refh_"
        f = self.lib.refh_gap_bits
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, u64p]
        bits = np.zeros((n_total + 63) // 64, np.uint64)
        f(h0, h1, n_threads, _ptr(bits, u64p))
        return bits

    # --- index lifecycle
    def load(self, fn):
        h = self._load(fn.encode())
        if not h:
            raise IOError("cannot load " + fn)
        return h

    def destroy(self, h):
        self._destroy(h)

    def dump(self, h, fn):
        return self._dump(h, fn.encode())

    def info(self, h):
        o = np.zeros(17, np.uint64)
        self._info(h, _ptr(o, u64p))
        return dict(mcnt=o[0:7].copy(), cnt=o[7:14].copy(), n_bytes=int(o[14]), n_frames=int(o[15]), ibits=int(o[16]))

    def from_bwt(self, bwt):
        assert self.p == "fo_"
        bwt = np.ascontiguousarray(bwt, np.uint8)
        return self.lib.fo_from_bwt(len(bwt), _ptr(bwt, u8p))

    def from_rle6(self, rle):
        assert self.p == "fo_"
        rle = np.ascontiguousarray(rle, np.uint8)
        return self.lib.fo_from_rle6(len(rle), _ptr(rle, u8p))

    def decode_bwt(self, h):
        assert self.p == "fo_"
        n = int(self.info(h)["mcnt"][0])
        out = np.zeros(n, np.uint8)
        m = self.lib.fo_decode_bwt(h, _ptr(out, u8p))
        assert m == n, (m, n)
        return out

    def build_text(self, text):
        """reference SA-IS build (fm_build, build.c:33) from the nt6 text; consumes a copy."""
        assert self.p == "refh_"
        t = np.ascontiguousarray(text, np.uint8).copy()
        return self.lib.refh_build_text(len(t), _ptr(t, u8p))

    # --- queries
    def rank1a(self, h, k):
        k = np.ascontiguousarray(k, np.uint64)
        ok = np.zeros((len(k), 6), np.uint64)
        sym = np.zeros(len(k), np.int32)
        self._rank1a(h, len(k), _ptr(k, u64p), _ptr(ok, u64p), _ptr(sym, i32p))
        return ok, sym

    def rank2a(self, h, k, l):
        k = np.ascontiguousarray(k, np.uint64)
        l = np.ascontiguousarray(l, np.uint64)
        ok = np.zeros((len(k), 6), np.uint64)
        ol = np.zeros((len(k), 6), np.uint64)
        self._rank2a(h, len(k), _ptr(k, u64p), _ptr(l, u64p), _ptr(ok, u64p), _ptr(ol, u64p))
        return ok, ol

    def extend(self, h, ik, is_back):
        ik = np.ascontiguousarray(ik, INTV)
        is_back = np.ascontiguousarray(is_back, np.uint8)
        ok = np.zeros((len(ik), 6), INTV)
        self._extend(h, len(ik), ik.ctypes.data, _ptr(is_back, u8p), ok.ctypes.data)
        return ok

    def backward_search(self, h, seq, off):
        n = len(off) - 1
        b = np.zeros(n, np.uint64)
        e = np.zeros(n, np.uint64)
        s = np.zeros(n, np.uint64)
        self._bsearch(h, n, _ptr(seq, u8p), _ptr(off, u64p), _ptr(b, u64p), _ptr(e, u64p), _ptr(s, u64p))
        return b, e, s

    def smem(self, h, seq, off, self_match=0, n_threads=1, want_records=True):
        """returns (records INTV[], mem_off u64[n+1], seconds, n_locate, n_extend)"""
        seq = np.ascontiguousarray(seq, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        n = len(off) - 1
        mo = np.zeros(n + 1, np.uint64)
        mem = C.c_void_p()
        secs = C.c_double()
        nl = C.c_uint64()
        ne = C.c_uint64()
        args = [h, n, _ptr(seq, u8p), _ptr(off, u64p), self_match, n_threads,
                C.byref(mem) if want_records else None, _ptr(mo, u64p), C.byref(secs)]
        if self.p == "fo_":
            args += [C.byref(nl), C.byref(ne)]
        self._smem(*args)
        rec = np.zeros(0, INTV)
        if want_records:
            tot = int(mo[-1])
            if tot:
                rec = np.frombuffer(C.string_at(mem.value, tot * 32), dtype=INTV).copy()
            self._free(mem)
        return rec, mo, secs.value, nl.value, ne.value

    def overlap(self, h, min_match, seeds):
        seeds = np.ascontiguousarray(seeds, np.uint64)
        n = len(seeds)
        rec = np.zeros((n, 9), np.int64)
        no = np.zeros(n + 1, np.uint64)
        nei = C.c_void_p()
        nl = C.c_uint64()
        args = [h, min_match, n, _ptr(seeds, u64p), _ptr(rec, i64p), C.byref(nei), _ptr(no, u64p)]
        if self.p == "fo_":
            args.append(C.byref(nl))
        self._overlap(*args)
        tot = int(no[-1])
        out = np.frombuffer(C.string_at(nei.value, tot * 32), dtype=INTV).copy() if tot else np.zeros(0, INTV)
        self._free(nei)
        return rec, out, no, nl.value


    def unitig_locates(self, h, min_match, seeds):
        """oracle port only: (locates[4] = block lookups of retrieve / is_contained / get_nei / check_left_simple over the seeds,
        left int8[n] = check_left_simple per seed, 1 = not evaluated)"""
        assert self.p == "fo_"
        seeds = np.ascontiguousarray(seeds, np.uint64)
        loc = (C.c_uint64 * 4)()
        left = np.ones(len(seeds), np.int8)
        f = self.lib.fo_unitig_locates
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_int64, u64p, C.POINTER(C.c_uint64), C.c_void_p]
        f(h, min_match, len(seeds), _ptr(seeds, u64p), loc, left.ctypes.data)
        return [int(x) for x in loc], left


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _oracle = _Lib(path, "fo_")
    return _oracle


def reference():
    """The compiled, unmodified reference (None when oracle/_ref/ has not been built)."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libfermi_ref.so")
        if not os.path.exists(path):
            return None
        _ref = _Lib(path, "refh_")
    return _ref


def ref_fermi_binary():
    p = os.path.join(ORACLE_DIR, "_ref", "fermi")
    return p if os.path.exists(p) else None


# ----------------------------------------------------------------------------- MAG records (mag.c:149-174)
def write_fastq(path, reads, qual_char="I"):
    """nt6 reads [n, L] -> FASTQ with constant qualities"""
    tab = np.array(list("$ACGTN"))
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@r%d\n%s\n+\n%s\n" % (i, "".join(tab[r]), qual_char * len(r)))


def parse_mag(text):
    """MAG text -> list of (k0, k1, nsr, nei0, nei1, seq, cov); nei = tuple of (id, ovlp)."""
    out = []
    lines = text.split("\n")
    i = 0
    while i + 3 < len(lines):
        if not lines[i].startswith("@"):
            i += 1
            continue
        f = lines[i][1:].split("\t")
        k0, k1 = (int(x) for x in f[0].split(":"))
        nei = []
        for col in f[2:4]:
            nei.append(tuple(tuple(int(v) for v in e.split(",")) for e in col.split(";") if e and e != "."))
        out.append((k0, k1, int(f[1]), nei[0], nei[1], lines[i + 1], lines[i + 3]))
        i += 4
    return out


def _rc_str(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def canonical_mag(records):
    """orientation-free, order-free form of a unitig set."""
    canon = []
    for k0, k1, nsr, n0, n1, seq, cov in records:
        a = (k0, k1, nsr, tuple(sorted(n0)), tuple(sorted(n1)), seq, cov)
        b = (k1, k0, nsr, tuple(sorted(n1)), tuple(sorted(n0)), _rc_str(seq), cov[::-1])
        canon.append(min(a, b) if k0 == k1 else (a if k0 < k1 else b))
    return sorted(canon)


def reference_unitig(fmd_path, min_match, threads=1):
    """MAG text of the compiled reference: fermi unitig -l <min_match> -t <threads> (cmd.c:184)."""
    binary = ref_fermi_binary()
    if binary is None:
        return None
    res = subprocess.run([binary, "unitig", "-l", str(min_match), "-t", str(threads), fmd_path], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, check=True)
    return res.stdout.decode()
