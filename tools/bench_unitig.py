#!/usr/bin/env python
"""Secondary measurement: `unitig -l50` overlap path (BASELINE config 3, scaled by --reads) on one GPU,
next to `fermi unitig -l50 -t<nproc>` of the compiled reference on the same .fmd.
    python tools/bench_unitig.py --reads 1000000 --err 0.0
"""
import argparse, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fermi_b200 as fb
import ctypes
from fermi_b200._lib import lib
ctypes.c_int.in_dll(lib(), 'fmg_verbose').value = 4
import helpers as H

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1000000)
ap.add_argument("--len", type=int, default=100)
ap.add_argument("--cov", type=float, default=10.0)
ap.add_argument("--err", type=float, default=0.0)
ap.add_argument("--no-ref", action="store_true")
a = ap.parse_args()
glen = int(a.reads * a.len / a.cov)
genome = fb.synth_genome(41, glen)
reads = fb.synth_reads(42, genome, a.reads, a.len, a.err)
t = time.time(); text = fb.fmd_text(reads); fmd = fb.fm_build(text, 0); t_bwt = time.time() - t
t = time.time(); fn = os.path.join(tempfile.gettempdir(), "bench_unitig.fmd"); fmd.dump(fn); t_enc = time.time() - t
t = time.time(); idx = fb.FmdIndex(fmd, 0); t_up = time.time() - t
out = os.path.join(tempfile.gettempdir(), "bench_unitig.mag")
fb.fm6_unitig(idx, 50, out)                      # warm-up (allocations, first touch)
t = time.time(); n = fb.fm6_unitig(idx, 50, out); t_gpu = time.time() - t
res = {"reads": a.reads, "err": a.err, "index_symbols": int(fmd.mcnt[0]), "unitigs": n, "build_fmd_s": t_bwt, "write_s": t_enc, "upload_s": t_up,
       "ours_total_s": t_gpu, "ours_reads_per_s": a.reads / t_gpu}
if not a.no_ref and H.ref_fermi_binary():
    cores = os.cpu_count()
    t = time.time()
    ref = subprocess.run([H.ref_fermi_binary(), "unitig", "-l", "50", "-t", str(cores), fn], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    t_ref = time.time() - t
    same = H.canonical_mag(H.parse_mag(ref)) == H.canonical_mag(H.parse_mag(open(out).read()))
    res.update({"ref_cores": cores, "ref_total_s": t_ref, "ref_reads_per_s": a.reads / t_ref, "set_equal": same})
print(json.dumps(res))
