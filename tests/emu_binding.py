"""TEST INFRASTRUCTURE ONLY: builds and binds tests/emu/emu.cpp, the host compilation of the
device-side core in fermi_b200/csrc/fmd_device.cuh (see the header of emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

import helpers as H

ROOT = H.ROOT
OUT = os.path.join(ROOT, "build", "libemu.so")
SRC = [os.path.join(ROOT, "tests", "emu", "emu.cpp"),
       os.path.join(ROOT, "fermi_b200", "csrc", "fmd_host.cpp"),
       os.path.join(ROOT, "fermi_b200", "csrc", "occ_build_host.cpp")]
DEPS = SRC + [os.path.join(ROOT, "fermi_b200", "csrc", f) for f in ("fmd_device.cuh", "fmd_overlap.cuh", "fmd_host.hpp", "occ_layout.hpp", "bcr_tile.cuh")]


class Emu:
    def __init__(self, path):
        L = self.lib = C.CDLL(path)
        L.fmg_fmd_restore.restype = C.c_void_p
        L.fmg_fmd_restore.argtypes = [C.c_char_p]
        L.fmg_fmd_destroy.argtypes = [C.c_void_p]
        L.emu_index_build.restype = C.c_void_p
        L.emu_index_build.argtypes = [C.c_void_p]
        L.emu_index_free.argtypes = [C.c_void_p]
        L.emu_rank2a.argtypes = [C.c_void_p, C.c_int64, H.u64p, H.u64p, H.u64p, H.u64p]
        L.emu_extend.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, H.u8p, C.c_void_p]
        L.emu_smem.argtypes = [C.c_void_p, C.c_int64, H.u8p, H.u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), H.u64p]
        L.emu_smem.restype = C.c_int
        L.fmg_free.argtypes = [C.c_void_p]
        L.emu_bcr_tile.argtypes = [C.c_int, C.c_int, C.c_int, H.u8p, H.u8p, C.c_void_p, C.c_int, H.u8p, H.u64p, C.c_void_p]
        L.emu_bcr_tile.restype = C.c_int
        L.emu_overlap.argtypes = [C.c_void_p, C.c_int, C.c_int64, H.u64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  H.i64p, C.c_void_p, C.c_void_p, H.u8p, H.i32p, H.u8p]

    def index(self, fmd_path):
        f = self.lib.fmg_fmd_restore(fmd_path.encode())
        assert f
        x = self.lib.emu_index_build(f)
        self.lib.fmg_fmd_destroy(f)
        return x

    def rank2a(self, x, k, l):
        k = np.ascontiguousarray(k, np.uint64)
        l = np.ascontiguousarray(l, np.uint64)
        ok = np.zeros((len(k), 6), np.uint64)
        ol = np.zeros((len(k), 6), np.uint64)
        self.lib.emu_rank2a(x, len(k), H._ptr(k, H.u64p), H._ptr(l, H.u64p), H._ptr(ok, H.u64p), H._ptr(ol, H.u64p))
        return ok, ol

    def extend(self, x, ik, is_back):
        ik = np.ascontiguousarray(ik, H.INTV)
        is_back = np.ascontiguousarray(is_back, np.uint8)
        ok = np.zeros((len(ik), 6), H.INTV)
        self.lib.emu_extend(x, len(ik), ik.ctypes.data, H._ptr(is_back, H.u8p), ok.ctypes.data)
        return ok

    def smem(self, x, seq, off, self_match, n_lanes=7, out_cap=64, wide=0):
        seq = np.ascontiguousarray(seq, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        mo = np.zeros(len(off), np.uint64)
        mem = C.c_void_p()
        ov = self.lib.emu_smem(x, len(off) - 1, H._ptr(seq, H.u8p), H._ptr(off, H.u64p), self_match, n_lanes, out_cap, wide,
                               C.byref(mem), H._ptr(mo, H.u64p))
        tot = int(mo[-1])
        rec = np.frombuffer(C.string_at(mem.value, tot * 32), dtype=H.INTV).copy() if tot else np.zeros(0, H.INTV)
        self.lib.fmg_free(mem)
        return rec, mo, ov


    def overlap(self, x, min_match, ids, max_len, cap=512, nei_cap=16, wide=0):
        """returns (rec[n,10], nei INTV[], nei_off[n+1], seq[n,max_len], len[n], ext[n,max_len])"""
        ids = np.ascontiguousarray(ids, np.uint64)
        n = len(ids)
        rec = np.zeros((n, 10), np.int64)
        nei = np.zeros((n, nei_cap), H.INTV)
        cnt = np.zeros(n, np.uint32)
        seq = np.zeros((n, max_len), np.uint8)
        ln = np.zeros(n, np.int32)
        ext = np.zeros((n, max_len), np.uint8)
        self.lib.emu_overlap(x, min_match, n, H._ptr(ids, H.u64p), max_len, cap, nei_cap, wide, H._ptr(rec, H.i64p),
                             nei.ctypes.data, cnt.ctypes.data, H._ptr(seq, H.u8p), H._ptr(ln, H.i32p), H._ptr(ext, H.u8p))
        off = np.concatenate([[0], np.cumsum(np.minimum(cnt, nei_cap))]).astype(np.uint64)
        flat = np.concatenate([nei[i, :min(int(cnt[i]), nei_cap)] for i in range(n)]) if n else np.zeros(0, H.INTV)
        return rec, flat, off, seq, ln, ext


    def bcr_tile(self, k_per, n_threads, flags, syms, old, shift):
        """one tile of the BCR merge: flags / syms per output position, old = the old symbols in order -> (out, hist[4], ranks)"""
        flags = np.ascontiguousarray(flags, np.uint8)
        syms = np.ascontiguousarray(syms, np.uint8)
        staged = np.full(shift + len(old) + 8 + (-(shift + len(old)) % 4), 0xdd, np.uint8)
        staged[shift: shift + len(old)] = old
        out = np.zeros(len(flags), np.uint8)
        hist = np.zeros(4, np.uint64)
        ranks = np.zeros(max(1, int(flags.sum())), np.uint32)
        n = self.lib.emu_bcr_tile(k_per, n_threads, len(flags), H._ptr(flags, H.u8p), H._ptr(syms, H.u8p), staged.view(np.uint32).ctypes.data, shift,
                                  H._ptr(out, H.u8p), H._ptr(hist, H.u64p), ranks.ctypes.data)
        assert n == int(flags.sum())
        return out, hist, ranks[:n]


def load():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", OUT] + SRC, check=True)
    return Emu(OUT)
