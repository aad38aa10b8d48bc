"""CPU, world_size 2, gloo: the sharding and the all-gather that reassembles overlap records before the
unitig walk (the N > 1 path of fermi_b200.parallel), fed by the host build of the overlap lane code."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_path):
    import torch.distributed as dist
    import emu_binding
    import fermi_b200 as fb
    from fermi_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(H.GOLDEN_DIR, case + ".npz"))
    E = emu_binding.load()
    x = E.index(os.path.join(H.GOLDEN_DIR, case + ".fmd"))
    max_len = int(g["ov_rec"][:, 1].max()) + 8

    def overlap(first, n):           # stands in for the GPU call on this CPU-only box (same lane code, host build)
        rec, nei, off, seq, ln, ext = E.overlap(x, int(g["ov_min"]), np.arange(first, first + n, dtype=np.uint64), max_len, nei_cap=32)
        return dict(rec=rec, nei=nei, nei_off=off, seq=seq, ext=ext)

    class Idx:                        # only mcnt is read by unitig_distributed
        mcnt = g["mcnt"]

    n = parallel.unitig_distributed(Idx, int(g["ov_min"]), out_path, max_len, overlap_fn=overlap)
    counts, offs = parallel.allgather_counts(10 + rank)
    assert counts == [10 + r for r in range(world)] and offs[rank] == sum(counts[:rank])
    if rank == 0:
        assert n > 0
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from fermi_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_unitig_row_shards_are_whole_reads():
    """fmg_overlap_shard wants shards of whole reads (row 2i = the read, 2i+1 = its reverse complement): even boundaries, every
    row exactly once, the last shard takes an odd tail (an index whose last read was given without its complement)."""
    from fermi_b200.parallel import unitig_shard_rows
    for n_seq in (0, 2, 6, 14, 15, 2000006, 2000007):
        for world in (1, 2, 3, 8):
            spans = [unitig_shard_rows(n_seq, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_seq
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(l % 2 == 0 for l, _ in spans) and all(h % 2 == 0 for _, h in spans[:-1])


def test_two_rank_unitig_equals_reference(product_lib, emu, tmp_path):
    case = "noisy"
    out = str(tmp_path / "u.mag")
    mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    ours = H.canonical_mag(H.parse_mag(open(out).read()))
    ref = H.canonical_mag(H.parse_mag(open(os.path.join(H.GOLDEN_DIR, case + ".mag")).read()))
    assert ours == ref
