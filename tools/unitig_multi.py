#!/usr/bin/env python
"""`fermi unitig` over N GPUs of one box (one rank per GPU, NCCL all-gather of the overlap records):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/unitig_multi.py --reads 1000000 [--check]
Every rank holds the whole index; sequences are sharded by contiguous ranges (fermi_b200.parallel)."""
import argparse, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import fermi_b200 as fb
from fermi_b200 import parallel

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1000000)
ap.add_argument("--err", type=float, default=0.0)
ap.add_argument("--check", action="store_true", help="compare with a single-GPU fm6_unitig on rank 0")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--host-gather", action="store_true", help="the round-1 path: records to the host, all-gather, host walk on rank 0")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keep stdout for the JSON line (the banner is a printf)
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fn = os.path.join(tempfile.gettempdir(), "unitig_multi_%d_%g.fmd" % (a.reads, a.err))
if rank == 0:
    genome = fb.synth_genome(41, a.reads * 10)
    reads = fb.synth_reads(42, genome, a.reads, 100, a.err)
    fb.fm_build(fb.fmd_text(reads), local).dump(fn + ".tmp"); os.replace(fn + ".tmp", fn)
dist.barrier()
idx = fb.FmdIndex(fb.Fmd.restore(fn), local)
out = os.path.join(tempfile.gettempdir(), "unitig_multi.mag")
tm = {}
for it in range(a.iters):
    dist.barrier(); torch.cuda.synchronize(); t = time.time()
    n = parallel.unitig_distributed(idx, 50, out, 108) if a.host_gather else parallel.unitig_distributed_device(idx, 50, out, timings=tm)
    dist.barrier(); dt = time.time() - t
if rank == 0:
    res = {"n_gpus": world, "reads": a.reads, "err": a.err, "path": "host gather" if a.host_gather else "device (NCCL all-gather of record shards, every rank assembles and writes its part)", "stages_s_rank0": tm,
           "unitigs": n, "seconds": dt, "reads_per_s": a.reads / dt}
    if a.check:
        import helpers as H
        single = out + ".single"
        fb.fm6_unitig(idx, 50, single)
        res["set_equal_single_gpu"] = H.canonical_mag(H.parse_mag(open(out).read())) == H.canonical_mag(H.parse_mag(open(single).read()))
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
