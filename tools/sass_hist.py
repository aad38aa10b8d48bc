#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels of libfermi_b200 (cuobjdump -sass on the objects the library is linked from):
    python tools/sass_hist.py > profiles/r02_sass_histogram.txt
Evidence for the instructions the design relies on: 256-bit index loads (LDG.E.*.256), the paired gather's shuffles (SHFL.BFLY),
warp votes (VOTE), population counts (POPC), bulk asynchronous copies of the BCR merge (UBLKCP) and their mbarriers (SYNCS)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "obj")
KERNELS = [("overlap.cu.o", "k_ov_chainIjLi1"), ("overlap.cu.o", "k_ov_neiIjLi4"), ("overlap.cu.o", "k_ov_chainImLi1"), ("overlap.cu.o", "k_ov_neiImLi4"),
           ("fmg_cuda.cu.o", "k_smemIjLb0"), ("fmg_cuda.cu.o", "k_smemIjLb1"), ("fmg_cuda.cu.o", "k_smemImLb1"), ("bcr.cu.o", "k_bcr_merge"),
           ("ec.cu.o", "k_trie_expand"), ("unitig_gpu.cu.o", "k_mag_body")]
MARK = ["LDG.E.ENL2.256", "LDG.E.128", "SHFL.BFLY", "SHFL.IDX", "VOTE", "POPC", "UBLKCP", "SYNCS", "PRMT", "ATOMG", "LDS", "STS", "BAR.SYNC", "WARPSYNC", "LDL", "STL"]
for obj, pat in KERNELS:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode(errors="replace")
    cur, hist, name = None, None, None
    out = {}
    for ln in txt.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            cur = out.setdefault(name, collections.Counter()) if pat in name else None
            continue
        if cur is None:
            continue
        m = re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            cur[m.group(1)] += 1
    for name, h in out.items():
        tot = sum(h.values())
        print("== %s  (%s, %d instructions)" % (name, obj, tot))
        print("   top: " + ", ".join("%s %d" % kv for kv in h.most_common(14)))
        marks = []
        for mk in MARK:
            c = sum(v for k, v in h.items() if k.startswith(mk))
            if c:
                marks.append("%s* %d" % (mk, c))
        print("   marked: " + ", ".join(marks))
        print()
