#!/usr/bin/env python
"""k-mer collection of `fermi correct` over N GPUs of one box (trie subtrees sharded by suffix, one all-gather of the triples):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ec_multi.py --reads 10000000 --len 150 --cov 35"""
import argparse, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import fermi_b200 as fb
from fermi_b200 import parallel
ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1000000)
ap.add_argument("--len", type=int, default=100)
ap.add_argument("--cov", type=float, default=10.0)
ap.add_argument("--err", type=float, default=0.01)
ap.add_argument("--check", action="store_true", help="compare with the single-GPU collection on rank 0")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keep stdout for the JSON line (the banner is a printf)
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fn = os.path.join(tempfile.gettempdir(), "ec_multi_%d_%d.fmd" % (a.reads, a.len))
if rank == 0:
    genome = fb.synth_genome(71, int(a.reads * a.len / a.cov))
    reads = fb.synth_reads(72, genome, a.reads, a.len, a.err)
    b = fb.Bcr(local)
    both = np.empty((2 * a.reads, a.len), np.uint8); both[0::2] = reads; both[1::2] = 5 - reads[:, ::-1]
    b.append_batch(both); del both
    b.build_fmd().dump(fn); b.close()
dist.barrier()
idx = fb.FmdIndex(fb.Fmd.restore(fn), local)
for it in range(2):
    dist.barrier(); torch.cuda.synchronize(); t = time.time()
    tri, cnt = parallel.ec_collect_distributed(idx, -1, 3)
    dist.barrier(); dt = time.time() - t
if rank == 0:
    res = {"n_gpus": world, "reads": a.reads, "len": a.len, "index_symbols": int(idx.fmd.mcnt[0]), "kmers": int(len(tri)), "informative": cnt[1],
           "seconds": dt, "kmers_per_s": len(tri) / dt}
    if a.check:
        t = time.time(); one, c1 = fb.fm6_ec_collect(idx, -1, 3); res["single_gpu_s"] = time.time() - t
        res["equal_single_gpu"] = bool(np.array_equal(one, tri) and tuple(c1) == tuple(cnt))
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
