#!/usr/bin/env python
"""Secondary measurement: RLD encoding of a read-set BWT on the GPU (rld_enc.cu) next to the serial host encoder and, as the
reference, `fermi recode` of the raw BWT stream.   python tools/bench_encode.py --reads 10000000"""
import argparse, ctypes, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fermi_b200 as fb
from fermi_b200._lib import lib
ctypes.c_int.in_dll(lib(), 'fmg_verbose').value = 4
import helpers as H
ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10000000)
ap.add_argument("--len", type=int, default=100)
ap.add_argument("--no-host", action="store_true")
a = ap.parse_args()
genome = fb.synth_genome(41, a.reads * a.len // 10)
reads = fb.synth_reads(42, genome, a.reads, a.len, 0.0)
text = fb.fmd_text(reads)
fb.fm_build(text[: 202 * 1000], 0)                      # warm-up
t = time.time(); bwt = fb.fm_build_bwt(text, 0); t_bwt = time.time() - t
t = time.time(); dev = fb.Fmd.from_bwt_device(bwt, 0); t_dev = time.time() - t
t = time.time(); dev2 = fb.fm_build(text, 0); t_all = time.time() - t
res = {"bench": "rld_encode", "symbols": int(len(text)), "fmd_bytes": int(dev.n_bytes), "gpu_bwt_s": t_bwt, "gpu_encode_incl_h2d_s": t_dev,
       "gpu_build_total_s": t_all, "gpu_encode_symbols_per_s": len(text) / t_dev}
if not a.no_host:
    t = time.time(); host = fb.Fmd.from_bwt(bwt); t_host = time.time() - t
    d = tempfile.gettempdir()
    dev.dump(os.path.join(d, "enc_dev.fmd")); host.dump(os.path.join(d, "enc_host.fmd")); dev2.dump(os.path.join(d, "enc_dev2.fmd"))
    same = open(os.path.join(d, "enc_dev.fmd"), "rb").read() == open(os.path.join(d, "enc_host.fmd"), "rb").read() == open(os.path.join(d, "enc_dev2.fmd"), "rb").read()
    res.update({"host_encode_s": t_host, "identical": bool(same)})
print(json.dumps(res))
