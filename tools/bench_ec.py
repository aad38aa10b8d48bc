#!/usr/bin/env python
"""Secondary measurement: k-mer collection of `fermi correct` (BASELINE config 5, scaled by --reads): fmg_ec_collect on one GPU
next to the reference's ec_collect (single thread, through oracle/_ref) on the same .fmd.
    python tools/bench_ec.py --reads 1000000"""
import argparse, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fermi_b200 as fb
import helpers as H
ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1000000)
ap.add_argument("--err", type=float, default=0.01)
ap.add_argument("--no-ref", action="store_true")
a = ap.parse_args()
genome = fb.synth_genome(71, a.reads * 10)
reads = fb.synth_reads(72, genome, a.reads, 100, a.err)
fmd = fb.fm_build(fb.fmd_text(reads), 0)
fn = os.path.join(tempfile.gettempdir(), "bench_ec.fmd"); fmd.dump(fn)
idx = fb.FmdIndex(fmd, 0)
for it in range(2):
    t = time.time(); tri, cnt = fb.fm6_ec_collect(idx, -1, 3); dt = time.time() - t
res = {"reads": a.reads, "err": a.err, "kmers": int(len(tri)), "informative": cnt[1], "ambiguous": cnt[0] - cnt[1], "ours_s": dt, "ours_kmers_per_s": len(tri) / dt}
R = H.reference()
if R is not None and not a.no_ref:
    h = R.load(fn); t = time.time(); ref = R.ec_collect(h, -1, 3); dt = time.time() - t
    res.update({"ref_1thread_s": dt, "equal": bool(np.array_equal(ref[0], tri))})
print(json.dumps(res))
