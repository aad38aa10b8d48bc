#!/usr/bin/env python
"""bench.py -- reads/s through the SMEM path (fm6_smem, smem.c:397) on BASELINE.json config 2:
10 M x 100 bp synthetic reads (1 % substitutions) against the FMD-index of a 100 Mbp i.i.d. genome cut
into 10 kb records.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole read set of a rank.
  value        device-resident: reads and the index already in HBM, results left in HBM (kernels only)
  e2e          through the C-ABI with pinned HOST buffers: H2D of the reads, kernels, D2H of the records
  roofline     k_smem alone (CUDA events on its stream): algorithmic bytes (N_locate x 128 + L + 32 n_out
               per read, SURVEY.md 8d, N_locate counted by an instrumented oracle run on a sample of the
               same reads) / kernel time, against the measured HBM copy peak
  cpu_baseline the reference's own fm6_smem (oracle/_ref, unmodified, all host threads) on a bounded
               sample of the same reads and the same .fmd file
N > 1: every rank holds the whole index and its own 10 M reads (weak scaling); no data-path collective,
value = reads of all ranks / max-over-ranks device time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GENOME_SEED, READ_SEED = 21, 22
RECORD_LEN = 10000


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--err", type=float, default=0.01)
    ap.add_argument("--batch-reads", type=int, default=2_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--unitig-reads", type=int, default=10_000_000, help="reads of the unitig leg (BASELINE config 3); 0 disables it")
    ap.add_argument("--unitig-ref-reads", type=int, default=500_000, help="reads of the bounded reference sample of the unitig leg")
    ap.add_argument("--no-unitig-noisy", action="store_true", help="skip the 1 %% substitution variant of the unitig leg")
    ap.add_argument("--no-smem-hbm", action="store_true", help="skip the SMEM leg against the index that does not fit L2")
    return ap.parse_args()


def workload_name(a):
    return "fermi smem (fm6_smem): %d x %d bp reads, %.0f%% subst, vs FMD-index of a %d bp i.i.d. genome in 10 kb records" % (
        a.reads, a.read_len, a.err * 100, a.genome)


def index_path(a):
    return os.path.join(tempfile.gettempdir(), "fermi_b200_bench_g%d_s%d.fmd" % (a.genome, GENOME_SEED))


def genome_records(fb, a):
    g = fb.synth_genome(GENOME_SEED, a.genome)
    n_rec = a.genome // RECORD_LEN
    return g, g[: n_rec * RECORD_LEN].reshape(n_rec, RECORD_LEN)


def make_reads(fb, a, genome, rank, out=None):
    return fb.synth_reads(READ_SEED + 1000 * rank, genome, a.reads, a.read_len, a.err, out=out)


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------- CPU side (checker libs)
def cpu_lib():
    """the compiled reference if it travelled with the repo (oracle/_ref), else the oracle port"""
    import helpers as H
    R = H.reference()
    if R is not None:
        return R, "reference"
    return H.oracle(), "port"


def cpu_smem_rate(fmd_file, reads, seconds, cores):
    """fm6_smem of the reference (or port) with `cores` strided threads on a bounded sample of `reads`."""
    import helpers as H
    L, kind = cpu_lib()
    h = L.load(fmd_file)
    pilot = reads[: min(len(reads), 64 * cores)]
    seq, off = H.reads_to_flat(pilot)
    _, _, t, _, _ = L.smem(h, seq, off, 0, cores, want_records=False)
    rate = len(pilot) / max(t, 1e-6)
    n = int(min(len(reads), max(len(pilot), rate * seconds)))
    seq, off = H.reads_to_flat(reads[:n])
    _, mo, t, _, _ = L.smem(h, seq, off, 0, cores, want_records=False)
    L.destroy(h)
    return {"value": n / t, "unit": "reads/s", "cores": cores, "kind": kind,
            "sample": "first %d reads of rank 0's read set, fm6_smem(self_match=0), %d strided pthreads, %.1f s wall" % (n, cores, t)}, int(mo[-1]), n


def count_locates(fmd_file, reads, cores):
    """N_locate / N_extend / n_out per read from the instrumented oracle port."""
    import helpers as H
    O = H.oracle()
    h = O.load(fmd_file)
    seq, off = H.reads_to_flat(reads)
    _, mo, _, nloc, next_ = O.smem(h, seq, off, 0, cores, want_records=False)
    O.destroy(h)
    n = len(reads)
    return nloc / n, next_ / n, int(mo[-1]) / n


# ----------------------------------------------------------------------------------- unitig leg (BASELINE config 3)
UNITIG_GENOME_SEED, UNITIG_READ_SEED, UNITIG_COV, UNITIG_MIN = 41, 42, 10, 50
# block lookups (rld_locate_blk calls) per input read of `fermi unitig -l50 -t1` on error-free 10x 100 bp reads, measured with
# a counter in the reference at 1/100 scale: 306 rank2a + ~120 rank1a -> 423 lookups.  Only the fallback
# when the instrumented oracle cannot be run on the benchmark input.
UNITIG_LOCATES_PER_READ = 423.0


def unitig_index_path(n_reads, err):
    return os.path.join(tempfile.gettempdir(), "fermi_b200_bench_unitig_%d_%g.fmd" % (n_reads, err))


def unitig_build_index(fb, n_reads, read_len, err, device, fn):
    genome = fb.synth_genome(UNITIG_GENOME_SEED, n_reads * read_len // UNITIG_COV)
    reads = fb.synth_reads(UNITIG_READ_SEED, genome, n_reads, read_len, err)
    t = time.time()
    fmd = fb.fm_build(fb.fmd_text(reads), device)          # suffix sort, BWT and RLD encoding on the GPU
    dt = time.time() - t
    fmd.dump(fn + ".tmp")
    os.replace(fn + ".tmp", fn)
    return dt


def smem_hbm_leg(fb, a, idx, fn, device, peak, peak_src):
    """SMEM where the index does NOT fit L2: fm6_smem(self_match=0) of reads sampled from the unitig genome (1 % substitutions)
    against the config-3 read index (2.02e9 symbols, 1 GB of occ blocks >> 126 MB L2), reads and results resident in HBM.  The
    session picks the paired-gather kernel for such an index; the plain-load kernel is timed beside it."""
    import torch
    import helpers as H
    L, n = a.read_len, a.reads
    genome = fb.synth_genome(UNITIG_GENOME_SEED, a.unitig_reads * L // UNITIG_COV)
    reads = fb.synth_reads(READ_SEED + 7, genome, n, L, a.err)
    dev = torch.device("cuda", device)
    d_reads = torch.from_numpy(reads).to(dev)
    B = min(a.batch_reads, n)
    batches = [(s, min(B, n - s)) for s in range(0, n, B)]
    d_boff = [(torch.arange(m + 1, dtype=torch.int64, device=dev) * L).contiguous() for _, m in batches]
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    for label, env in (("paired_gathers", None), ("plain_loads", "0")):
        if env is None:
            os.environ.pop("FMG_SMEM_PAIR", None)
        else:
            os.environ["FMG_SMEM_PAIR"] = env
        sess = fb.SmemSession(idx, B, L)
        for _ in range(2):
            for (s0, m), bo in zip(batches, d_boff):
                sess.run(m, d_reads[s0:].data_ptr(), bo.data_ptr(), 0, stream)
        sess.result()
        sess.set_timing(True)
        sess.kernel_ms()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 3
        e0.record()
        for _ in range(steps):
            for (s0, m), bo in zip(batches, d_boff):
                sess.run(m, d_reads[s0:].data_ptr(), bo.data_ptr(), 0, stream)
        e1.record()
        torch.cuda.synchronize()
        k_ms, _ = sess.kernel_ms()
        res[label] = {"ms_per_step": e0.elapsed_time(e1) / steps, "kernel_ms_per_step": k_ms / steps}
        sess.close()
    os.environ.pop("FMG_SMEM_PAIR", None)
    sample_n = min(n, 10000)
    n_loc, n_ext, n_out = count_locates(fn, reads[:sample_n], min(os.cpu_count() or 1, 16))
    bytes_per_read = n_loc * 128 + L + 32 * n_out
    k = res["paired_gathers"]["kernel_ms_per_step"]
    achieved = n * bytes_per_read / (k / 1e3) / 1e9
    traffic = None
    try:
        import glob
        tr = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_smem_hbm_traffic.json")))[-1]))
        traffic = tr["traffic_bytes_per_launch"] * (B / tr["reads_per_launch"])
    except Exception:
        pass
    return {"workload": "fm6_smem: %d x %d bp reads (%g%% subst) vs the FMD-index of the config-3 reads (%d symbols, %.0f MB of occ blocks: HBM regime)"
                        % (n, L, a.err * 100, int(idx.fmd.mcnt[0]), idx.nbytes / 1e6),
            "value": n / (res["paired_gathers"]["ms_per_step"] / 1e3), "unit": "reads/s", "ms_per_step": res["paired_gathers"]["ms_per_step"],
            "plain_loads_value": n / (res["plain_loads"]["ms_per_step"] / 1e3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "k_smem<u32, paired gathers>", "kernel_ms_per_step": k, "plain_loads_kernel_ms_per_step": res["plain_loads"]["kernel_ms_per_step"],
                         "algorithmic_bytes_per_read": bytes_per_read, "algorithmic_bytes_per_launch": bytes_per_read * B,
                         "n_locate_per_read": round(n_loc, 2), "n_extend_per_read": round(n_ext, 2), "records_per_read": round(n_out, 3),
                         "counter_sample": "instrumented oracle on the first %d reads" % sample_n, "peak_source": peak_src}}


def unitig_count_locates(fn, n_seq, sample=3000):
    """N_locate per input read from the instrumented oracle on the ACTUAL index: the block lookups of the work the
    reference's walk does for one read it visits -- fm_retrieve of the seed, overlap_intv + fm6_get_nei, check_left_simple -- over a
    strided sample of the seed rows (odd rows, unitig.c:333-334)."""
    import helpers as H
    O = H.oracle()
    h = O.load(fn)
    n_seeds = n_seq // 2
    step = max(1, n_seeds // sample)
    seeds = (np.arange(0, n_seeds, step, dtype=np.uint64)[:sample] * 2 + 1)
    loc, _ = O.unitig_locates(h, UNITIG_MIN, seeds)
    O.destroy(h)
    n = len(seeds)
    return {"locates_per_read": sum(loc) / n, "by_stage": {"fm_retrieve": loc[0] / n, "fm6_is_contained": loc[1] / n, "fm6_get_nei": loc[2] / n,
                                                          "check_left_simple": loc[3] / n}, "sample_seeds": n}


def mag_set(path):
    import helpers as H
    with open(path) as fh:
        return H.canonical_mag(H.parse_mag(fh.read()))


def unitig_run(fb, idx, out, world, dist, timings):
    """one `fermi unitig -l50` over all ranks: the public call on one GPU, the NCCL path (fermi_b200.parallel) on several"""
    if world == 1:
        n = fb.fm6_unitig(idx, UNITIG_MIN, out)
        return n
    from fermi_b200 import parallel
    return parallel.unitig_distributed_device(idx, UNITIG_MIN, out, timings=timings)


def unitig_leg(fb, a, device, peak, peak_src, rank, world, dist, barrier, err=0.0, iters=3, check_single=True, cpu_baseline=True):
    """`fermi unitig -l50` (fm6_unitig, unitig.c:378) over the FMD-index of n x 100 bp 10x reads, end to end: overlap records of the
    rank's share of the sequences, (N > 1) the NCCL all-gather of the record shards, unitig assembly and the MAG text written to
    one file.  Timed with a barrier + device synchronisation on both sides, max over ranks.  Rank 0 returns the result object."""
    import torch
    import helpers as H
    L = a.read_len
    fn = unitig_index_path(a.unitig_reads, err)
    t_build = None
    if rank == 0 and not os.path.exists(fn):
        t_build = unitig_build_index(fb, a.unitig_reads, L, err, device, fn)
        fb.release_cache()
    barrier()
    fmd = fb.Fmd.restore(fn)
    idx = fb.FmdIndex(fmd, device)
    n_seq, n_sym = int(fmd.mcnt[1]), int(fmd.mcnt[0])
    # the MAG text goes to a memory-backed file where there is one (the page-cache copy is part of the timed call either way;
    # a journalling file system under /tmp adds its own per-page cost that has nothing to do with the path measured)
    out_dir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    out = os.path.join(out_dir, "fermi_b200_bench_unitig.mag")
    dev = torch.device("cuda", device)
    launches0 = fb.launch_count()
    tm = {}
    unitig_run(fb, idx, out, world, dist, tm)              # warm-up: scratch pool, pinned result buffers, NCCL channels
    launches = fb.launch_count() - launches0
    times, stages, n_u = [], [], 0
    for _ in range(iters):
        barrier()
        t = time.perf_counter()
        n_u = unitig_run(fb, idx, out, world, dist, tm)
        barrier()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt[0]))
        stages.append(dict(tm))
    secs = sum(times) / len(times)
    st = fb.overlap_stats()
    k_ms = torch.tensor([st["contained"] + st["neighbours"] + st["left_fix"], st["contained"], st["neighbours"], st["left_fix"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)
    k_ms = [float(x) for x in k_ms.tolist()]
    res = None
    if rank == 0:
        res = {"workload": "fermi unitig -l%d: FMD-index of %d x %d bp reads (%dx, %g%% substitutions), %d sequences, %d symbols" %
                           (UNITIG_MIN, a.unitig_reads, L, UNITIG_COV, err * 100, n_seq, n_sym),
               "value": a.unitig_reads / secs, "unit": "reads/s", "seconds": secs, "n_gpus": world, "scaling": "strong", "unitigs": int(n_u),
               "mag_bytes": os.path.getsize(out), "mag_path": out,
               "api": "fmg_unitig (records + assembly on the GPU, MAG text written to a file)" if world == 1 else
                      "fermi_b200.parallel.unitig_distributed_device: fmg_overlap_shard -> NCCL all-gather of the record shards -> fmg_overlap_merge + "
                      "fmg_overlap_left_fix -> fmg_unitig_part (every rank assembles, formats and writes the chains it owns)",
               "gpu_launches_per_call": int(launches)}
        if t_build is not None:
            res["setup"] = {"gpu_build_fmd_s": round(t_build, 2)}
        if world > 1:
            last = stages[-1]
            res["stages_ms_rank0"] = {k: round(1e3 * v, 2) for k, v in last.items() if isinstance(v, float)}
            res["collective_ms"] = round(1e3 * last.get("exchange", 0.0), 2)
            res["assembly_ms"] = round(1e3 * (last.get("assembly", 0.0) + last.get("write", 0.0)), 2)
            res["exchange_bytes_per_rank"] = last.get("exchange_bytes")
            res["host_walk"] = bool(last.get("host_walk"))
        # roofline of the overlap kernels: algorithmic bytes of the whole job / the slowest rank's kernel time
        loc = None
        try:
            loc = unitig_count_locates(fn, n_seq)
        except Exception as exc:
            log("unitig: instrumented oracle unavailable (%r); using the SURVEY constant" % (exc,))
        n_loc = loc["locates_per_read"] if loc else UNITIG_LOCATES_PER_READ
        bytes_per_read = n_loc * 128 + L
        achieved = a.unitig_reads * bytes_per_read / (k_ms[0] / 1e3) / 1e9 if k_ms[0] > 0 else 0.0
        traffic = None
        try:
            import glob
            tr = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_ov_traffic.json")))[-1]))
            traffic = tr["traffic_bytes_per_sequence"] * n_seq / world          # DRAM bytes of the phase kernels of one rank's pass (ncu --set full)
        except Exception:
            pass
        res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world), "traffic": traffic,
                           "kernel": "k_ov_chain<1> (fm_retrieve + fm6_is_contained) + k_ov_nei (fm6_get_nei) + deferred check_left_simple pass",
                           "kernel_ms": k_ms[0], "kernel_ms_by_phase_max_over_ranks": {"contained": k_ms[1], "neighbours": k_ms[2], "left_fix": k_ms[3]},
                           "left_rows": st["left_rows"], "algorithmic_bytes_per_read": bytes_per_read, "locates_per_read": n_loc,
                           "locates_source": ("instrumented oracle (oracle/fmd_oracle.c: fo_unitig_locates) on %d seed rows of this index: %s" %
                                              (loc["sample_seeds"], json.dumps({k: round(v, 1) for k, v in loc["by_stage"].items()}))) if loc else
                                             "SURVEY.md 8d: counter in the reference's rld_locate_blk, `unitig -l50 -t1`, error-free 10x, 1/100 scale",
                           "peak_source": peak_src + (" x %d GPUs" % world if world > 1 else "")}
        if world == 1 and err == 0.0 and not a.no_smem_hbm:
            try:
                res["smem_hbm"] = smem_hbm_leg(fb, a, idx, fn, device, peak, peak_src)
            except Exception as exc:
                res["smem_hbm"] = {"error": repr(exc)}
        if check_single and world > 1:
            single = out + ".single"
            fb.fm6_unitig(idx, UNITIG_MIN, single)          # untimed: the same index through the single-GPU call
            res["set_equal_single_gpu"] = mag_set(out) == mag_set(single)
    idx.close()
    del fmd
    fb.release_cache()
    ref_bin = H.ref_fermi_binary()
    if rank == 0 and cpu_baseline and ref_bin and a.unitig_ref_reads > 0 and not a.no_cpu_baseline:
        # the unmodified reference on a bounded sample generated the same way, and our MAG set against its own on that sample
        cores = os.cpu_count() or 1
        rfn = unitig_index_path(a.unitig_ref_reads, err)
        if not os.path.exists(rfn):
            unitig_build_index(fb, a.unitig_ref_reads, L, err, device, rfn)
        t = time.perf_counter()
        with open(out + ".ref", "wb") as fh:
            subprocess.run([ref_bin, "unitig", "-l", str(UNITIG_MIN), "-t", str(cores), rfn], stdout=fh, stderr=subprocess.DEVNULL, check=True)
        dt = time.perf_counter() - t
        res["cpu_baseline"] = {"value": a.unitig_ref_reads / dt, "unit": "reads/s", "cores": cores, "kind": "reference",
                               "sample": "fermi unitig -l%d -t%d on the index of %d reads generated the same way (incl. loading the .fmd), %.1f s wall"
                                         % (UNITIG_MIN, cores, a.unitig_ref_reads, dt)}
        ridx = fb.FmdIndex(fb.Fmd.restore(rfn), device)
        fb.fm6_unitig(ridx, UNITIG_MIN, out + ".ours")
        ridx.close()
        res["set_equal_reference_sample"] = mag_set(out + ".ours") == mag_set(out + ".ref")
        fb.release_cache()
    barrier()
    return res


# ----------------------------------------------------------------------------------- reference arm
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import helpers as H
    import fermi_b200 as fb
    L, kind = cpu_lib()
    cores = os.cpu_count() or 1
    fn = index_path(a)
    genome, recs = genome_records(fb, a)
    if not os.path.exists(fn):
        t0 = time.time()
        if kind == "reference":
            text = fb.fmd_text(recs)
            h = L.build_text(text)                    # the reference's own SA-IS build (build.c:33)
            L.dump(h, fn)
            L.destroy(h)
        else:                                          # no compiled reference on this box: set-up only, not timed
            fb.fm_build(fb.fmd_text(recs), 0).dump(fn)
        log("reference arm: index built in %.1f s (%s)" % (time.time() - t0, kind))
    n_sample = min(a.reads, 12000 * cores)
    reads = fb.synth_reads(READ_SEED, genome, n_sample, a.read_len, a.err)
    seq, off = H.reads_to_flat(reads)
    h = L.load(fn)
    times = []
    for it in range(a.warmup + a.steps):
        _, _, t, _, _ = L.smem(h, seq, off, 0, cores, want_records=False)
        if it >= a.warmup:
            times.append(t)
    L.destroy(h)
    ms = 1e3 * sum(times) / len(times)
    val = n_sample / (ms / 1e3)
    sample = "each step = fm6_smem over the first %d reads of the read set with %d strided pthreads" % (n_sample, cores)
    print(json.dumps({
        "impl": "reference", "metric": "reads/sec through SMEM (fm6_smem)", "value": val, "unit": "reads/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "reads/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ----------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import fermi_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; fermi_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the one JSON line only: NCCL_DEBUG=VERSION (set on the GPU boxes) makes NCCL printf its banner there
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- set-up (untimed): genome, index file (rank 0 builds, everyone loads), reads
    t0 = time.time()
    genome, recs = genome_records(fb, a)
    fn = index_path(a)
    if rank == 0 and not os.path.exists(fn):
        text = fb.fmd_text(recs)
        t1 = time.time()
        fmd = fb.fm_build(text, local)                 # suffix sort, BWT and RLD encoding on the GPU
        t2 = time.time()
        fmd.dump(fn + ".tmp")
        os.replace(fn + ".tmp", fn)
        log("index: %d symbols, GPU build (BWT + RLD encoding) %.1f s, write %.1f s" % (len(text), t2 - t1, time.time() - t2))
        del text, fmd
    barrier()
    fmd = fb.Fmd.restore(fn)
    t1 = time.time()
    idx = fb.FmdIndex(fmd, local)
    log("rank %d: index upload %.1f s, %.1f MB occ blocks; set-up so far %.1f s" % (rank, time.time() - t1, idx.nbytes / 1e6, time.time() - t0))

    L = a.read_len
    h_reads = torch.empty((a.reads, L), dtype=torch.uint8).pin_memory()
    make_reads(fb, a, genome, rank, out=h_reads.numpy())
    h_off = (torch.arange(a.reads + 1, dtype=torch.int64) * L).pin_memory()
    d_reads = h_reads.to(dev, non_blocking=True)
    d_off = h_off.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    B = min(a.batch_reads, a.reads)
    sess = fb.SmemSession(idx, B, L)
    stream = torch.cuda.current_stream().cuda_stream
    batches = [(s, min(B, a.reads - s)) for s in range(0, a.reads, B)]
    d_boff = [(d_off[s: s + n + 1] - s * L).contiguous() for s, n in batches]   # per-batch offsets, resident
    torch.cuda.synchronize()

    def device_step():
        tot = 0
        for (s, n), bo in zip(batches, d_boff):
            sess.run(n, d_reads[s:].data_ptr(), bo.data_ptr(), 0, stream)
        return tot

    launches0 = fb.launch_count()
    for _ in range(a.warmup):
        device_step()
    n_rec_last, _, _ = sess.result()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    sess.set_timing(True)
    sess.kernel_ms()
    launches1 = fb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        device_step()
    e1.record()
    barrier()
    step_ms = e0.elapsed_time(e1) / a.steps
    k_ms, k_n = sess.kernel_ms()
    sess.set_timing(False)
    launches_timed = fb.launch_count() - launches1
    clocks = sampler.stop()

    # ---- e2e: pinned host buffers through the C-ABI (H2D + kernels + D2H inside the timed region), with fmintv_t records
    # (fmg_smem_batch_into) and with the packed 16-byte records (fmg_smem_batch_into16: half the bytes back to the host)
    e2e = None
    e2e32_s = 0.0
    if not a.no_e2e:
        rec_cap = int(a.reads * 16)
        h_mem = torch.empty((rec_cap, 4), dtype=torch.int64).pin_memory()
        h_moff = torch.empty(a.reads + 1, dtype=torch.int64).pin_memory()
        sess.close()                                   # the host API owns its sessions

        def timed(call):
            n_rec, times = 0, []
            n_warm = max(1, min(a.warmup, 2))
            for it in range(n_warm + max(1, min(a.steps, 3))):
                barrier()
                t = time.perf_counter()
                n_rec = call(idx, a.reads, h_reads.data_ptr(), h_off.data_ptr(), h_mem.data_ptr(), rec_cap, h_moff.data_ptr(), 0, a.batch_reads)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t
                if it >= n_warm:
                    times.append(dt)
            return sum(times) / len(times), n_rec

        e2e32_s, n_rec = timed(fb.fm6_smem_raw)
        packed = int(fmd.mcnt[0]) < (1 << 32) and L < 32768
        if packed:
            e2e_s, n_rec16 = timed(fb.fm6_smem_raw16)
            assert n_rec16 == n_rec
        else:
            e2e_s = e2e32_s
        rb = 16 if packed else 32
        e2e = {"seconds": e2e_s, "n_rec": n_rec, "h2d": a.reads * L + (a.reads + 1) * 8, "d2h": n_rec * rb + (a.reads + 1) * 8, "record_bytes": rb,
               "d2h32": n_rec * 32 + (a.reads + 1) * 8}

    # ---- reduce over ranks: max time, summed reads
    t_dev = torch.tensor([step_ms, k_ms / max(1, a.steps), e2e["seconds"] if e2e else 0.0, e2e32_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    step_ms_max, k_ms_step_max, e2e_s_max, e2e32_s_max = [float(x) for x in t_dev.tolist()]
    total_reads = a.reads * world

    if rank == 0:
        cores = os.cpu_count() or 1
        sample_n = min(a.reads, 20000)
        n_loc, n_ext, n_out = count_locates(fn, h_reads.numpy()[:sample_n], min(cores, 16))
        bytes_per_read = n_loc * 128 + L + 32 * n_out
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        achieved = a.reads * bytes_per_read / (k_ms_step_max / 1e3) / 1e9
        traffic = None            # DRAM bytes per k_smem launch from the committed `ncu --set full` capture
        try:
            import glob
            tr = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_smem_traffic.json")))[-1]))
            traffic = tr["traffic_bytes_per_launch"] * (B / tr["reads_per_launch"])
        except Exception:
            pass
        out = {
            "metric": "reads/sec through SMEM (fm6_smem)", "value": total_reads / (step_ms_max / 1e3), "unit": "reads/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": workload_name(a), "reads_per_gpu": a.reads, "batch_reads": B, "index_symbols": int(fmd.mcnt[0]),
                "index_sequences": int(fmd.mcnt[1]), "index_hbm_mb": round(idx.nbytes / 1e6, 1), "parallelism": "replicated index, reads sharded x%d" % world,
                "l2": "each step streams %.2f GB of reads and %.2f GB of records (>> 126 MB L2); the %.0f MB index is re-read within a step by design"
                      % (a.reads * L / 1e9, n_out * a.reads * 32 / 1e9, idx.nbytes / 1e6),
                "records_per_read": round(n_out, 3), "n_locate_per_read": round(n_loc, 2), "n_extend_per_read": round(n_ext, 2),
                "counter_sample": "instrumented oracle on the first %d reads" % sample_n,
                "coordinates": "records are fmintv_t (u64), bit-exact; the kernels are instantiated with 32-bit coordinates when the index has < 2^32 symbols (this one), with 64-bit ones otherwise"},
            "clocks": clocks,
            "gpu_launches": int(launches_timed),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "k_smem", "kernel_ms_per_step": k_ms_step_max, "launches_per_step": k_n // max(1, a.steps),
                         "algorithmic_bytes_per_read": bytes_per_read, "algorithmic_bytes_per_launch": bytes_per_read * B,
                         "peak_source": peak_src,
                         "frac_of_8TBs": achieved / 8000.0},
        }
        if e2e:
            out["e2e"] = {"value": total_reads / e2e_s_max, "unit": "reads/s", "h2d_bytes_per_step": int(e2e["h2d"]),
                          "d2h_bytes_per_step": int(e2e["d2h"]), "seconds_per_step": e2e_s_max, "record_bytes": e2e["record_bytes"],
                          "api": ("fmg_smem_batch_into16 (pinned host buffers, 3-stream batch pipeline, 16-byte records packed on the device; "
                                  "fmg_intv16_expand gives fmintv_t)") if e2e["record_bytes"] == 16 else
                                 "fmg_smem_batch_into (pinned host buffers, 3-stream batch pipeline)",
                          "fmintv_t_records": {"value": total_reads / e2e32_s_max, "unit": "reads/s", "seconds_per_step": e2e32_s_max,
                                               "d2h_bytes_per_step": int(e2e["d2h32"]), "api": "fmg_smem_batch_into (32-byte fmintv_t records)"}}
        if world == 1 and not a.no_cpu_baseline:
            cb, _, _ = cpu_smem_rate(fn, h_reads.numpy(), a.cpu_seconds, cores)
            out["cpu_baseline"] = cb
    # ---- the second half of the metric (overlap / unitig, BASELINE config 3): its own index, so the SMEM buffers go first
    out_line = out if rank == 0 else None
    if a.unitig_reads > 0:
        del d_reads, d_off, d_boff, h_reads, h_off
        if e2e:
            del h_mem, h_moff
        else:
            sess.close()
        idx.close()
        fb.release_cache()
        torch.cuda.empty_cache()
        pk = float(peaks.get("hbm_gbs", 6650.0)) if rank == 0 else 0.0
        for key, err, iters in (("unitig", 0.0, 3), ("unitig_err1pct", 0.01, 1)):
            if err > 0 and a.no_unitig_noisy:
                continue
            try:
                r = unitig_leg(fb, a, local, pk, peak_src if rank == 0 else "", rank, world, dist, barrier, err=err, iters=iters,
                               check_single=(err == 0.0), cpu_baseline=(world == 1 and err == 0.0))
            except Exception as exc:                       # the SMEM line stands on its own
                r = {"error": repr(exc)}
                if world > 1:
                    raise
            if rank == 0:
                out_line[key] = r
    if rank == 0:
        print(json.dumps(out_line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
