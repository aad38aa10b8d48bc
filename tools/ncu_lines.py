#!/usr/bin/env python
"""Aggregate the warp-stall samples of one kernel of an ncu report by CUDA source line.
ncu's CSV source page is per SASS instruction without line numbers; this joins it (by instruction offset) with the
line table nvdisasm prints for the same kernel of the cubin inside libfermi_b200.so.
usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX MANGLED_SUBSTRING [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, kre, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "fermi_b200", "lib", "libfermi_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode(errors="replace")
    cur, loc, inl = None, None, None
    for ln in txt.split("\n"):
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            cur = m.group(1) if mangled in m.group(1) else None
            continue
        if cur is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and loc:
            lines[int(m.group(1), 16)] = (loc, m.group(2))
    if lines:
        break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(io.StringIO(out)))
# several kernels may match: take the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None and r:
        cur["rows"].append(r)
b = blocks[0]
hdr = b["rows"][0]; data = b["rows"][1:]
ia, isamp, iinst, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ilsb = hdr.index("stall_long_sb")
base = int(data[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = 0
for r in data:
    off = int(r[ia], 16) - base
    loc = lines.get(off, (("?", 0), ""))[0]
    a = agg[loc]
    a[0] += int(r[isamp] or 0); a[1] += int(r[iinst] or 0); a[2] += int(r[ithr] or 0); a[3] += int(r[ilsb] or 0)
    tot += int(r[isamp] or 0)
print(b["name"], "samples", tot, "instructions", len(data))
src = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        p = os.path.join(ROOT, "fermi_b200", "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print("%5.1f%%  %-18s %4d  winst=%-10d thr/inst=%4.1f long_sb=%-6d %s" % (100.0 * a[0] / max(tot, 1), f, l, a[1], a[2] / max(a[1], 1), a[3], text))
