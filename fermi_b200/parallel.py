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


def unitig_shard_rows(n_seq, rank, world):
    """rows [lo, hi) of `rank`: shards of whole reads (row 2i = the read, row 2i+1 = its reverse complement)"""
    lo, hi = shard_range(n_seq // 2, rank, world)
    return 2 * lo, (2 * hi if rank < world - 1 else n_seq)


def overlap_shard_device(idx, min_match, lo, hi, max_len=0):
    """fmg_overlap_shard into torch buffers: (rec int64[m*8], rank int64[m], ext uint8[], spill int64[], ext_total, spill_total)"""
    import ctypes as C
    from ._lib import lib
    L = lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    m = hi - lo
    rec = torch.empty(max(m, 1) * 8, dtype=torch.int64, device=dev)
    rnk = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    ext_cap, spill_cap = max(32 * m, 1 << 16), max(2 * m, 1 << 12)
    tot = (C.c_uint64 * 2)()
    while True:
        ext = torch.empty(ext_cap, dtype=torch.uint8, device=dev)
        spill = torch.empty(spill_cap * 4, dtype=torch.int64, device=dev)
        rc = L.fmg_overlap_shard(idx.h, int(min_match), int(max_len), lo, hi, rec.data_ptr(), rnk.data_ptr(), ext.data_ptr(), ext_cap,
                                 spill.data_ptr(), spill_cap, tot)
        if rc == 1:                                           # capacities too small: the call reports the need
            ext_cap, spill_cap = max(ext_cap, int(tot[0])), max(spill_cap, int(tot[1]))
            continue
        if rc != 0:
            raise RuntimeError("fermi_b200: fmg_overlap_shard failed (see stderr)")
        return rec, rnk, ext, spill, int(tot[0]), int(tot[1])


def unitig_distributed_device(idx, min_match, out_path, max_len=0, group=None, timings=None):
    """`fermi unitig` over all ranks with everything in HBM (NCCL only).  Every rank computes the packed overlap records of its
    range of BWT rows on its GPU (fmg_overlap_shard); ONE all-gather (four tensors: records, ranks, appended bases, fork
    neighbour lists, each shard padded to the largest) gives every rank all records; every rank merges them into the
    rank-indexed array (fmg_overlap_merge), evaluates the deferred left checks (fmg_overlap_left_fix), builds the link graph
    and assembles + formats the unitigs whose chain head it owns (fmg_unitig_part: head rank % world == rank); after an
    all-gather of the text sizes every rank writes its text at its offset of the one MAG file.  Returns the total number of
    unitigs (on every rank).  `timings` (dict) receives wall-clock seconds of the stages, measured with device synchronisation."""
    import ctypes as C
    import time
    from ._lib import lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _device()
    L = lib()
    n_seq = int(idx.mcnt[1])
    t = [time.perf_counter()]

    def lap():
        torch.cuda.synchronize()
        t.append(time.perf_counter())

    lo, hi = unitig_shard_rows(n_seq, rank, world)
    m = hi - lo
    rec, rnk, ext, spill, ext_tot, spill_tot = overlap_shard_device(idx, min_match, lo, hi, max_len)
    lap()
    sizes = torch.tensor([m, ext_tot, spill_tot], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)                      # shard sizes
    all_sizes = all_sizes.view(world, 3).cpu()
    rows = all_sizes[:, 0].tolist()
    row_pad, ext_pad, spill_pad = max(int(all_sizes[:, 0].max()), 1), max(int(all_sizes[:, 1].max()), 1), max(int(all_sizes[:, 2].max()), 1)

    def gather(src, n_used, pad, width=1):
        buf = src
        if src.numel() < pad * width:                         # the shard buffer is smaller than the largest shard: pad a copy
            buf = torch.zeros(pad * width, dtype=src.dtype, device=dev)
            buf[: n_used * width] = src[: n_used * width]
        out = torch.empty(world * pad * width, dtype=src.dtype, device=dev)
        dist.all_gather_into_tensor(out, buf[: pad * width], group=group)
        return out

    rec_all = gather(rec, m, row_pad, 8)                      # the one exchange of the path: 64-byte records ...
    rank_all = gather(rnk, m, row_pad)                        # ... their ranks ...
    ext_all = gather(ext, ext_tot, ext_pad)                   # ... appended bases ...
    spill_all = gather(spill, spill_tot, spill_pad, 4)        # ... and the neighbour lists of forks
    lap()
    pack = torch.empty(n_seq * 8, dtype=torch.int64, device=dev)
    rank_of_row = torch.empty(n_seq, dtype=torch.int64, device=dev)
    c_rows = (C.c_uint64 * world)(*rows)
    if L.fmg_overlap_merge(idx.h, world, c_rows, row_pad, ext_pad, spill_pad, rec_all.data_ptr(), rank_all.data_ptr(), pack.data_ptr(),
                           rank_of_row.data_ptr()) != 0:
        raise RuntimeError("fermi_b200: fmg_overlap_merge failed")
    del rec_all, rank_all
    # the deferred left check, shared: every rank evaluates its own rows on the merged array, one all-gather of a byte per row
    # hands the flags round (the check of a row reads the records of other rows but patches only its own)
    n_left = C.c_uint64()
    if L.fmg_overlap_left_fix_rows(idx.h, int(min_match), int(max_len), pack.data_ptr(), rank_of_row.data_ptr(), lo, hi, C.byref(n_left)) != 0:
        raise RuntimeError("fermi_b200: fmg_overlap_left_fix_rows failed")
    if world > 1:
        flags = torch.zeros(row_pad, dtype=torch.int8, device=dev)
        if L.fmg_overlap_left_flags(idx.h, pack.data_ptr(), rank_of_row.data_ptr(), lo, hi, flags.data_ptr(), 0) != 0:
            raise RuntimeError("fermi_b200: fmg_overlap_left_flags failed")
        flags_all = torch.empty(world * row_pad, dtype=torch.int8, device=dev)
        dist.all_gather_into_tensor(flags_all, flags, group=group)
        first = 0
        for s in range(world):
            if s != rank and rows[s] and L.fmg_overlap_left_flags(idx.h, pack.data_ptr(), rank_of_row.data_ptr(), first, first + rows[s],
                                                                 flags_all.data_ptr() + s * row_pad, 1) != 0:
                raise RuntimeError("fermi_b200: fmg_overlap_left_flags failed")
            first += rows[s]
    lap()
    part, nu, nb = C.c_void_p(), C.c_uint64(), C.c_uint64()
    rc = L.fmg_unitig_part(idx.h, int(min_match), pack.data_ptr(), rank_of_row.data_ptr(), ext_all.data_ptr(), spill_all.data_ptr(), rank, world,
                           C.byref(part), C.byref(nu), C.byref(nb))
    lap()
    state = torch.tensor([rc, int(nu.value), int(nb.value)], dtype=torch.int64, device=dev)
    all_state = torch.empty(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_state, state, group=group)                      # text sizes (and whether the graph was regular)
    all_state = all_state.view(world, 3).cpu()
    if int(all_state[:, 0].max()) == 1 or int(all_state[:, 0].min()) < 0:
        # irregular link graph (the same on every rank: they hold the same records): the single-GPU path with the host walk
        if part:
            L.fmg_magpart_free(part)
        if int(all_state[:, 0].min()) < 0:
            raise RuntimeError("fermi_b200: fmg_unitig_part failed")
        n = None
        if rank == 0:
            from . import api
            n = api.fm6_unitig(idx, min_match, out_path, max_len)
        res = torch.tensor([n if n is not None else 0], dtype=torch.int64, device=dev)
        dist.broadcast(res, 0, group=group)
        if timings is not None:
            timings.update(records=t[1] - t[0], exchange=t[2] - t[1], merge_left=t[3] - t[2], assembly=t[4] - t[3], write=0.0, host_walk=True,
                           left_rows=int(n_left.value))
        return int(res[0])
    sizes_b = all_state[:, 2].tolist()
    t_sizes = time.perf_counter()
    path = str(out_path).encode()
    if rank == 0:                                             # the file has its final size before anyone writes into it (no truncation to
        import os                                             # zero first: dropping the pages of a previous output costs more than reusing them)
        fd = os.open(out_path, os.O_RDWR | os.O_CREAT, 0o644)
        os.ftruncate(fd, sum(sizes_b))
        os.close(fd)
    dist.barrier(group=group)
    t_sized = time.perf_counter()
    if L.fmg_magpart_write(part, path, sum(sizes_b[:rank]), 0) != 0:
        raise RuntimeError("fermi_b200: fmg_magpart_write failed")
    L.fmg_magpart_free(part)
    t_copied = time.perf_counter()
    dist.barrier(group=group)
    lap()
    if timings is not None:
        timings.update(records=t[1] - t[0], exchange=t[2] - t[1], merge_left=t[3] - t[2], assembly=t[4] - t[3], write=t[5] - t[4], host_walk=False,
                       write_wait_for_all_parts=t_sizes - t[4], write_size_file=t_sized - t_sizes, write_copy=t_copied - t_sized, write_final_barrier=t[5] - t_copied,
                       left_rows=int(n_left.value), exchange_bytes=int(8 * (world - 1) * (row_pad * 9) + (world - 1) * (ext_pad + spill_pad * 32)))
    return int(all_state[:, 1].sum())


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
