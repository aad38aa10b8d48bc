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
ap.add_argument("--len", type=int, default=100)
ap.add_argument("--cov", type=float, default=10.0)
ap.add_argument("--bcr", action="store_true", help="build the index with BCR + the device RLD encoder (any size; BASELINE config 5 is 150 bp, 35x)")
a = ap.parse_args()
genome = fb.synth_genome(71, int(a.reads * a.len / a.cov))
reads = fb.synth_reads(72, genome, a.reads, a.len, a.err)
t = time.time()
if a.bcr:
    b = fb.Bcr(0)
    both = np.empty((2 * a.reads, a.len), np.uint8); both[0::2] = reads; both[1::2] = 5 - reads[:, ::-1]
    b.append_batch(both); del both
    fmd = b.build_fmd(); b.close()
else:
    fmd = fb.fm_build(fb.fmd_text(reads), 0)
t_build = time.time() - t
fn = os.path.join(tempfile.gettempdir(), "bench_ec.fmd")
if not a.no_ref:
    fmd.dump(fn)
idx = fb.FmdIndex(fmd, 0)
for it in range(2):
    t = time.time(); tri, cnt = fb.fm6_ec_collect(idx, -1, 3); dt = time.time() - t
res = {"reads": a.reads, "len": a.len, "cov": a.cov, "index_symbols": int(fmd.mcnt[0]), "index_hbm_mb": round(idx.nbytes / 1e6, 1), "build_s": t_build,
       "kmer_length": fb.ec_kmer_length(int(fmd.mcnt[0])) if hasattr(fb, "ec_kmer_length") else None, "err": a.err, "kmers": int(len(tri)), "informative": cnt[1], "ambiguous": cnt[0] - cnt[1], "ours_s": dt, "ours_kmers_per_s": len(tri) / dt}
R = H.reference()
if R is not None and not a.no_ref:
    h = R.load(fn); t = time.time(); ref = R.ec_collect(h, -1, 3); dt = time.time() - t
    res.update({"ref_1thread_s": dt, "equal": bool(np.array_equal(ref[0], tri))})
print(json.dumps(res))
