"""Multi-GPU plumbing (one process per GPU, torch.distributed): the index is replicated in every GPU's
HBM and the units of work -- reads for SMEM, sequences for the unitig overlap records -- are sharded by
contiguous ranges, the way the reference stripes its threads (unitig.c:394-404, smem.c:346-381).

The only exchange on the path is the all-gather that reassembles the overlap records before the unitig
walk; SMEM results stay with the rank that owns the reads.  torch.distributed is used for
rendezvous and the collective only (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .api import INTV


def shard_range(n, rank, world):
    """contiguous [lo, hi) of n units for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _allgather_rows(a, counts, group=None):
    """all-gather of a 2-D array whose row count differs per rank (padded to the largest shard)."""
    world = dist.get_world_size(group)
    dev = _device()
    width = a.shape[1]
    pad = int(max(counts))
    t = torch.zeros((pad, width), dtype=torch.from_numpy(a[:0]).dtype, device=dev)
    if len(a):
        t[: len(a)] = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.concatenate([o[: int(c)].cpu().numpy() for o, c in zip(outs, counts)], axis=0)


def allgather_overlap_records(local, group=None):
    """Reassemble the per-sequence overlap records (fermi_b200.fm6_overlap) of all ranks, in rank order.
    `local` holds this rank's shard: rec[m,10], nei INTV[], nei_off[m+1], seq[m,L], ext[m,L]."""
    world = dist.get_world_size(group)
    dev = _device()
    m = len(local["rec"])
    n_nei = len(local["nei"])
    sizes = torch.tensor([m, n_nei], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)                    # (1) shard sizes
    rows = [int(s[0]) for s in all_sizes]
    neis = [int(s[1]) for s in all_sizes]
    rec = _allgather_rows(local["rec"], rows, group)                   # (2) padded payloads
    seq = _allgather_rows(local["seq"], rows, group)
    ext = _allgather_rows(local["ext"], rows, group)
    cnt = np.diff(local["nei_off"].astype(np.int64)).reshape(-1, 1)
    cnt = _allgather_rows(cnt, rows, group).reshape(-1)
    nei = _allgather_rows(np.ascontiguousarray(local["nei"]).view(np.int64).reshape(-1, 4), neis, group)
    nei = np.ascontiguousarray(nei).view(INTV).reshape(-1)
    nei_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
    return dict(rec=rec, nei=nei, nei_off=nei_off, seq=seq, ext=ext)


def unitig_distributed(idx, min_match, out_path, max_len, overlap_fn=None, group=None):
    """`fermi unitig` over all ranks: every rank computes the overlap records of its range of sequences on its
    GPU, one all-gather reassembles them, rank 0 walks the unitigs and writes the MAG file.  Returns the
    number of unitigs on rank 0 (None elsewhere)."""
    from . import api
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_seq = int(idx.mcnt[1])
    lo, hi = shard_range(n_seq, rank, world)
    fn = overlap_fn or (lambda first, n: api.fm6_overlap(idx, min_match, first=first, step=1, n=n, max_len=max_len))
    local = fn(lo, hi - lo)
    full = allgather_overlap_records(local, group)
    if rank == 0:
        return api.fm6_unitig_assemble(n_seq, min_match, full, out_path)
    return None


def unitig_distributed_device(idx, min_match, out_path, max_len=0, group=None):
    """`fermi unitig` over all ranks with everything in HBM: every rank computes the packed overlap records of its range of BWT
    rows on its GPU (fmg_overlap_shard), ONE exchange merges the shards -- an all-reduce(sum) of the rank-indexed 64-byte
    record array, which the shards fill disjointly, and all-gathers of the rank / appended-base / fork-neighbour shards --
    and rank 0 assembles the unitigs on its GPU (fmg_unitig_from_device) and writes the MAG file.  NCCL only.
    Returns the number of unitigs on rank 0 (None elsewhere)."""
    import ctypes as C
    from ._lib import lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _device()
    L = lib()
    n_seq = int(idx.mcnt[1])
    lo, hi = shard_range(n_seq // 2, rank, world)             # shards of whole reads: rows 2i (read) and 2i+1 (reverse complement)
    lo, hi = 2 * lo, (2 * hi if rank < world - 1 else n_seq)
    m = hi - lo
    pack = torch.zeros(n_seq * 8, dtype=torch.int64, device=dev)
    rnk = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    ext_cap, spill_cap = max(32 * m, 1 << 16), max(2 * m, 1 << 12)
    tot = (C.c_uint64 * 2)()
    while True:
        ext = torch.empty(ext_cap, dtype=torch.uint8, device=dev)
        spill = torch.empty(spill_cap * 4, dtype=torch.int64, device=dev)
        rc = L.fmg_overlap_shard(idx.h, int(min_match), int(max_len), lo, hi, pack.data_ptr(), rnk.data_ptr(), ext.data_ptr(), ext_cap,
                                 spill.data_ptr(), spill_cap, tot)
        if rc == 1:                                           # capacities too small: the call reports the need
            pack.zero_()
            ext_cap, spill_cap = max(ext_cap, int(tot[0])), max(spill_cap, int(tot[1]))
            continue
        if rc != 0:
            raise RuntimeError("fermi_b200: fmg_overlap_shard failed (see stderr)")
        break
    sizes = torch.tensor([m, int(tot[0]), int(tot[1])], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)                      # shard sizes
    all_sizes = all_sizes.view(world, 3).cpu()
    rows, exts, spills = all_sizes[:, 0].tolist(), all_sizes[:, 1].tolist(), all_sizes[:, 2].tolist()
    if L.fmg_overlap_rebase(idx.h, pack.data_ptr(), rnk.data_ptr(), m, sum(exts[:rank]), sum(spills[:rank])) != 0:
        raise RuntimeError("fermi_b200: fmg_overlap_rebase failed")
    dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)                        # the records: disjoint shards, so sum = union

    def gather(t, counts, width=1):
        pad = max(max(counts), 1) * width
        buf = torch.zeros(pad, dtype=t.dtype, device=dev)
        buf[: counts[rank] * width] = t[: counts[rank] * width]
        out = torch.empty(world * pad, dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, buf, group=group)
        return torch.cat([out[r * pad: r * pad + counts[r] * width] for r in range(world)]) if world > 1 else out[: counts[0] * width]

    rank_all = gather(rnk, rows)
    ext_all = gather(ext, exts)
    spill_all = gather(spill, spills, 4)
    n = None
    if rank == 0:
        nu = C.c_uint64()
        rc = L.fmg_unitig_from_device(idx.h, int(min_match), pack.data_ptr(), rank_all.data_ptr(), ext_all.data_ptr() if len(ext_all) else 0,
                                      sum(exts), spill_all.data_ptr() if len(spill_all) else 0, sum(spills), str(out_path).encode(), C.byref(nu))
        if rc == 1:                                           # irregular link graph: the single-GPU path with the host walk
            from . import api
            n = api.fm6_unitig(idx, min_match, out_path, max_len)
        elif rc != 0:
            raise RuntimeError("fermi_b200: fmg_unitig_from_device failed")
        else:
            n = int(nu.value)
    torch.cuda.synchronize()
    return n


def ec_collect_distributed(idx, w=-1, min_occ=3, group=None):
    """The k-mer collection of `fermi correct` over all ranks: rank r expands the trie subtrees of the suffixes s with
    s % world == r on its GPU (the reference stripes the same units over its threads, correct.c:346-350), one all-gather
    brings the (suffix, key, val) triples and the two counters together.  Returns (sorted triples, (cnt0, cnt1)) on every rank."""
    from . import api
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    tri, cnt = api.fm6_ec_collect(idx, w, min_occ, part=rank, n_parts=world)
    dev = _device()
    sizes = torch.tensor([len(tri), cnt[0], cnt[1]], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.view(world, 3).cpu()
    counts = all_sizes[:, 0].tolist()
    pad = max(max(counts), 1)
    buf = torch.zeros(pad, dtype=torch.int64, device=dev)
    buf[: len(tri)] = torch.from_numpy(tri.view(np.int64)).to(dev)
    out = torch.empty(world * pad, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    merged = torch.cat([out[r * pad: r * pad + counts[r]] for r in range(world)])
    res = np.sort(merged.cpu().numpy().view(np.uint64))
    return res, (int(all_sizes[:, 1].sum()), int(all_sizes[:, 2].sum()))


def allgather_counts(n_local, group=None):
    """per-rank unit counts -> (counts, exclusive offsets): global numbering of sharded SMEM results."""
    world = dist.get_world_size(group)
    t = torch.tensor([n_local], dtype=torch.int64, device=_device())
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    counts = [int(o[0]) for o in outs]
    return counts, [sum(counts[:i]) for i in range(world)]
