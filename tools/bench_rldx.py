#!/usr/bin/env python
"""Secondary measurement: rank2a throughput of the two device layouts on the same index -- occ blocks (one thread per query) and the
.fmd stream itself (fmg_rldx_*: one warp per query, TMA-staged blocks, pointer-doubling decode).
    python tools/bench_rldx.py --reads 2000000"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fermi_b200 as fb

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=2000000)
ap.add_argument("--queries", type=int, default=4000000)
a = ap.parse_args()
genome = fb.synth_genome(41, a.reads * 10)
reads = fb.synth_reads(42, genome, a.reads, 100, 0.0)
fmd = fb.fm_build(fb.fmd_text(reads), 0)
n = int(fmd.mcnt[0])
rng = np.random.RandomState(1)
k = rng.randint(0, n - 64, a.queries).astype(np.uint64)
l = k + rng.randint(0, 40, a.queries).astype(np.uint64)
res = {"index_symbols": n, "fmd_bytes": int(fmd.n_bytes), "queries": a.queries}
for name, make, call in (("occ_blocks", lambda: fb.FmdIndex(fmd, 0), lambda h: fb.rld_rank2a(h, k, l)), ("rld_stream", lambda: fb.RldIndex(fmd, 0), lambda h: h.rank2a(k, l))):
    h = make()
    call(h)
    t = time.time(); out = call(h); dt = time.time() - t
    res[name] = {"hbm_bytes": h.nbytes, "seconds_incl_copies": dt, "rank2a_per_s_incl_copies": a.queries / dt}
    res.setdefault("_chk", []).append(int(out[0].sum() % (1 << 61)))
    h.close()
res["equal"] = res["_chk"][0] == res["_chk"][1]
del res["_chk"]
print(json.dumps(res))
