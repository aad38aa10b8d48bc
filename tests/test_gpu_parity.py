"""GPU: the CUDA path (through the C-ABI of libfermi_b200.so) against the golden vectors generated
from the unmodified reference and against the oracle on fresh seeded inputs.  Bit-exact everywhere
(integer work): SA intervals, bi-intervals, SMEM records, BWT bytes."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from conftest import golden_cases

pytestmark = pytest.mark.gpu


def _load(case):
    return np.load(os.path.join(H.GOLDEN_DIR, case + ".npz")), os.path.join(H.GOLDEN_DIR, case + ".fmd")


@pytest.fixture(scope="module")
def fb(product_lib):
    import fermi_b200
    return fermi_b200


@pytest.mark.parametrize("case", golden_cases())
def test_golden_rank_extend_smem_search(fb, case):
    g, fmd = _load(case)
    idx = fb.FmdIndex(fb.Fmd.restore(fmd), 0)
    ok, ol = fb.rld_rank2a(idx, g["k"], g["l"])
    assert np.array_equal(ok, g["ok"]) and np.array_equal(ol, g["ol"])
    assert np.array_equal(fb.fm6_extend(idx, g["ik"], g["is_back"]), g["ext"])
    seq, off = H.reads_to_flat(g["q"])
    for sm, rk, ok_ in ((0, "smem0", "moff0"), (1, "smem1", "moff1")):
        rec, mo = fb.fm6_smem(idx, seq, off, sm)
        assert np.array_equal(mo, g[ok_]) and np.array_equal(rec, g[rk])
    b, e, s = fb.fm_backward_search(idx, seq, off)
    assert np.array_equal(b, g["sa_beg"]) and np.array_equal(e, g["sa_end"]) and np.array_equal(s, g["sa_size"])
    # rld_rank1a with its returned symbol (rld.c:424-446) against the oracle, incl. k = -1, and the `chkbwt -r` self-check (cmd.c:90-116)
    O = H.oracle()
    h = O.load(fmd)
    n_sym = int(idx.fmd.mcnt[0])
    rng = np.random.RandomState(3)
    ks = np.concatenate([rng.randint(0, n_sym, 5000).astype(np.uint64), np.array([0, n_sym - 1, 2**64 - 1], np.uint64)])
    ook, osym = O.rank1a(h, ks)
    ok1, sym1 = fb.rld_rank1a(idx, ks)
    assert np.array_equal(ok1, ook) and np.array_equal(sym1, osym)
    O.destroy(h)
    assert fb.check_rank(idx) == (0, 2**64 - 1)
    idx.close()


@pytest.mark.parametrize("case", golden_cases())
def test_gpu_bwt_build_reproduces_reference_fmd(fb, case, tmp_path):
    """fm_build on the GPU + host encoder == the reference's SA-IS build, byte for byte."""
    g, fmd = _load(case)
    out = str(tmp_path / "gpu.fmd")
    fb.fm_build(g["text"], 0).dump(out)
    assert open(out, "rb").read() == open(fmd, "rb").read()


def test_bwt_build_edge_cases(fb):
    rng = np.random.RandomState(4)
    texts = [
        np.array([0], np.uint8),
        np.array([1, 0, 4, 0], np.uint8),
        H.fmd_text([np.full(300, 1, np.uint8)] * 3),                       # long runs, identical sequences
        H.fmd_text([rng.randint(1, 5, size=rng.randint(1, 400)).astype(np.uint8) for _ in range(50)]),  # ragged
        H.fmd_text([np.tile(np.array([1, 2], np.uint8), 200)] * 2 + [np.tile(np.array([2, 1], np.uint8), 150)]),  # periodic
    ]
    for t in texts:
        assert np.array_equal(fb.fm_build_bwt(t, 0), H.naive_bwt(t))


def test_fresh_inputs_vs_oracle_100k_reads(fb, oracle, tmp_path):
    """BASELINE config 1: build from 100k x 100 bp reads, SA intervals + SMEM records bit-exact."""
    genome = fb.synth_genome(7, 1000000)
    reads = fb.synth_reads(8, genome, 100000, 100, 0.0)
    fmd = fb.fm_build(fb.fmd_text(reads), 0)
    fn = str(tmp_path / "r.fmd")
    fmd.dump(fn)
    ho = oracle.load(fn)
    assert np.array_equal(oracle.info(ho)["mcnt"], fmd.mcnt)
    idx = fb.FmdIndex(fmd, 0)
    q = np.concatenate([reads[:20000], fb.synth_reads(9, genome, 20000, 100, 0.01)])
    seq, off = H.reads_to_flat(q)
    b, e, s = fb.fm_backward_search(idx, seq, off)
    ob, oe, os_ = oracle.backward_search(ho, seq, off)
    assert np.array_equal(b, ob) and np.array_equal(e, oe) and np.array_equal(s, os_)
    assert (s[:20000] > 0).all()
    for sm in (0, 1):
        rec, mo = fb.fm6_smem(idx, seq, off, sm)
        orec, omo, _, _, _ = oracle.smem(ho, seq, off, sm, 8)
        assert np.array_equal(mo, omo) and np.array_equal(rec, orec)
    n = int(fmd.mcnt[0])
    rng = np.random.RandomState(1)
    k = rng.randint(0, n, size=200000).astype(np.uint64)
    l = np.minimum(k + rng.randint(0, 3000, size=200000).astype(np.uint64), np.uint64(n - 1))
    k[:100] = np.uint64(0xFFFFFFFFFFFFFFFF)
    ok, ol = fb.rld_rank2a(idx, k, l)
    ook, ool = oracle.rank2a(ho, k, l)
    assert np.array_equal(ok, ook) and np.array_equal(ol, ool)
    oracle.destroy(ho)
    idx.close()


def test_64bit_coordinate_kernel_matches_too(fb, monkeypatch):
    """indexes of >= 2^32 symbols use k_smem<uint64_t>; force it on a small index and compare with the vectors."""
    monkeypatch.setenv("FMG_FORCE_WIDE", "1")
    for case in golden_cases():
        g, fmd = _load(case)
        idx = fb.FmdIndex(fb.Fmd.restore(fmd), 0)
        seq, off = H.reads_to_flat(g["q"])
        for sm, rk, ok_ in ((0, "smem0", "moff0"), (1, "smem1", "moff1")):
            rec, mo = fb.fm6_smem(idx, seq, off, sm)
            assert np.array_equal(mo, g[ok_]) and np.array_equal(rec, g[rk])
        idx.close()


def test_ragged_empty_and_multibatch(fb, oracle):
    """empty / 1-base / ragged reads, and the multi-batch host pipeline (tiny batches force many)."""
    import ctypes as C
    fmd_path = os.path.join(H.GOLDEN_DIR, "reads10x.fmd")
    g = np.load(os.path.join(H.GOLDEN_DIR, "reads10x.npz"))
    rng = np.random.RandomState(9)
    reads = [g["q"][i % 300][: rng.randint(0, 101)] for i in range(3000)]
    reads[0] = reads[0][:0]
    reads[17] = g["q"][17][:1]
    reads[2900] = np.concatenate([g["q"][5], g["q"][6], g["q"][7][:37]])      # longer than anything before it: the pipeline, sized by
    seq = np.concatenate(reads).astype(np.uint8)                               # the batches seen so far, must rebuild itself mid-call
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    idx = fb.FmdIndex(fb.Fmd.restore(fmd_path), 0)
    h = oracle.load(fmd_path)
    for sm in (0, 1):
        orec, omo, _, _, _ = oracle.smem(h, seq, off, sm, 4)
        rec, mo = fb.fm6_smem(idx, seq, off, sm)
        assert np.array_equal(mo, omo) and np.array_equal(rec, orec)
        mem = np.zeros(len(orec) + 8, H.INTV)
        mo2 = np.zeros(len(off), np.uint64)
        n = fb.fm6_smem_raw(idx, len(off) - 1, seq.ctypes.data, off.ctypes.data, mem.ctypes.data, len(mem), mo2.ctypes.data,
                            sm, batch_reads=257)
        assert n == len(orec) and np.array_equal(mo2, omo) and np.array_equal(mem[:n], orec)
    oracle.destroy(h)
    idx.close()


def test_packed_records_paired_gathers_and_session_guard(fb, oracle, monkeypatch):
    """fmg_smem_batch_into16 (16-byte records packed on the device) expands to the records of fmg_smem_batch_into; the paired-gather
    kernel (chosen for indexes larger than L2, forced here) gives the same records; a read longer than a session was created
    for is reported, not run (it would overrun the lane's candidate lists)."""
    import torch
    fmd_path = os.path.join(H.GOLDEN_DIR, "noisy.fmd")
    g = np.load(os.path.join(H.GOLDEN_DIR, "noisy.npz"))
    seq, off = H.reads_to_flat(g["q"])
    h = oracle.load(fmd_path)
    idx = fb.FmdIndex(fb.Fmd.restore(fmd_path), 0)
    for pair in ("0", "1"):
        monkeypatch.setenv("FMG_SMEM_PAIR", pair)
        for sm in (0, 1):
            orec, omo, _, _, _ = oracle.smem(h, seq, off, sm, 4)
            mem = np.zeros((len(orec) + 8, 4), np.uint32)
            mo = np.zeros(len(off), np.uint64)
            n = fb.fm6_smem_raw16(idx, len(off) - 1, seq.ctypes.data, off.ctypes.data, mem.ctypes.data, len(mem), mo.ctypes.data, sm, batch_reads=97)
            assert n == len(orec) and np.array_equal(mo, omo)
            assert np.array_equal(fb.intv16_expand(mem[:n]), orec)
            rec, mo2 = fb.fm6_smem(idx, seq, off, sm)
            assert np.array_equal(mo2, omo) and np.array_equal(rec, orec)
        idx.close()                                            # the pipeline (and its kernel choice) lives with the handle
        idx = fb.FmdIndex(fb.Fmd.restore(fmd_path), 0)
    monkeypatch.delenv("FMG_SMEM_PAIR")
    oracle.destroy(h)
    L = g["q"].shape[1]
    sess = fb.SmemSession(idx, 16, L - 1)                       # one base too short for these reads
    d_seq = torch.from_numpy(seq[: 16 * L].copy()).cuda()
    d_off = torch.from_numpy(off[:17].astype(np.int64)).cuda()
    sess.run(16, d_seq.data_ptr(), d_off.data_ptr(), 0, 0)
    with pytest.raises(RuntimeError):
        sess.result()
    sess.close()
    idx.close()


def test_record_slot_overflow_is_rerun_not_truncated(fb, oracle, tmp_path):
    """a read with more SMEMs than the default 64 record slots: the session must grow and re-run."""
    rng = np.random.RandomState(2)
    # many short distinct sequences; one long query that stitches them together yields > 64 SMEMs
    pieces = [rng.randint(1, 5, size=24).astype(np.uint8) for _ in range(120)]
    text = H.fmd_text(pieces)
    fmd = fb.fm_build(text, 0)
    fn = str(tmp_path / "p.fmd")
    fmd.dump(fn)
    q = np.concatenate(pieces[:100]).astype(np.uint8)
    seq = np.concatenate([q, pieces[3]])
    off = np.array([0, len(q), len(q) + 24], np.uint64)
    h = oracle.load(fn)
    orec, omo, _, _, _ = oracle.smem(h, seq, off, 0, 1)
    assert omo[1] > 64
    idx = fb.FmdIndex(fmd, 0)
    rec, mo = fb.fm6_smem(idx, seq, off, 0)
    assert np.array_equal(mo, omo) and np.array_equal(rec, orec)
    oracle.destroy(h)
    idx.close()


@pytest.mark.timeout(180)
def test_smem_long_queries_vs_oracle(fb, oracle, tmp_path):
    """queries far longer than reads (contigs against a read index: what fm6_remap and `exact` on assemblies send): the lane count
    shrinks so that the per-lane candidate lists fit, the records stay bit-exact."""
    genome = fb.synth_genome(91, 400000)
    reads = fb.synth_reads(92, genome, 40000, 100, 0.0)
    fmd = fb.fm_build(fb.fmd_text(reads), 0)
    fn = str(tmp_path / "r.fmd")
    fmd.dump(fn)
    idx = fb.FmdIndex(fmd, 0)
    rng = np.random.RandomState(4)
    contigs = []
    for L in (200000, 60000, 15000, 100):
        s0 = rng.randint(0, len(genome) - L)
        c = genome[s0: s0 + L].copy()
        sub = rng.randint(0, L, size=L // 300 + 1)                      # substitutions break the matches (an N as the first base of a
        c[sub] = 1 + (c[sub] % 4)                                       # match is undefined behaviour in the reference: not used here)
        contigs.append(c)
    seq = np.concatenate(contigs)
    off = np.concatenate([[0], np.cumsum([len(c) for c in contigs])]).astype(np.uint64)
    h = oracle.load(fn)
    orec, omo, _, _, _ = oracle.smem(h, seq, off, 0, 4)
    rec, mo = fb.fm6_smem(idx, seq, off, 0)
    assert len(rec) > 1000 and np.array_equal(mo, omo) and np.array_equal(rec, orec)
    oracle.destroy(h)
    idx.close()


@pytest.mark.parametrize("case", golden_cases())
def test_overlap_records_and_unitigs_match_reference(fb, case, tmp_path, monkeypatch):
    """the four overlap phases against the reference's fm_retrieve / fm6_is_contained / fm6_get_nei records, and the
    unitig set of fmg_unitig -- device assembly and host walk -- against `fermi unitig -t1` (canonicalised MAG records)."""
    g, fmd = _load(case)
    idx = fb.FmdIndex(fb.Fmd.restore(fmd), 0)
    max_len = int(g["ov_rec"][:, 1].max()) + 8
    o = fb.fm6_overlap(idx, int(g["ov_min"]), ids=g["ov_seeds"], max_len=max_len)
    assert np.array_equal(o["rec"][:, :9], g["ov_rec"])
    assert np.array_equal(o["nei"], g["ov_nei"]) and np.array_equal(o["nei_off"], g["ov_off"])
    out = str(tmp_path / "u.mag")
    ref = H.parse_mag(open(os.path.join(H.GOLDEN_DIR, case + ".mag")).read())
    for host in ("0", "1"):
        monkeypatch.setenv("FMG_UNITIG_HOST", host)
        n = fb.fm6_unitig(idx, int(g["ov_min"]), out)
        ours = H.parse_mag(open(out).read())
        assert n == len(ref) and H.canonical_mag(ours) == H.canonical_mag(ref), "host walk" if host == "1" else "device assembly"
    idx.close()


@pytest.mark.skipif(H.ref_fermi_binary() is None, reason="oracle/_ref/fermi did not travel with the repo")
@pytest.mark.parametrize("err,cov,circular", [(0.0, 10, False), (0.01, 10, False), (0.02, 40, False), (0.0, 12, True)])
def test_unitig_20k_reads_vs_reference_binary(fb, tmp_path, monkeypatch, err, cov, circular):
    """BASELINE config 3 in miniature: unitig -l50 over 20k x 100 bp reads, set-equal to the reference -- through the device
    assembly (unitig_gpu.cu) and through the host walk.  The circular case (reads wrap around a 20 kb plasmid, error-free)
    makes the link graph one cycle, which the device path hands to the host walk."""
    n_reads = 20000 if not circular else 2400
    genome = fb.synth_genome(31, n_reads * 100 // cov)
    src = np.concatenate([genome, genome[:100]]) if circular else genome
    reads = fb.synth_reads(32, src, n_reads, 100, err)
    fmd = fb.fm_build(fb.fmd_text(reads), 0)
    fn = str(tmp_path / "r.fmd")
    fmd.dump(fn)
    ref = H.parse_mag(H.reference_unitig(fn, 50, 1))
    idx = fb.FmdIndex(fmd, 0)
    out = str(tmp_path / "u.mag")
    for host in ("0", "1"):
        monkeypatch.setenv("FMG_UNITIG_HOST", host)
        monkeypatch.setenv("FMG_THREADS", "1" if circular else "8")     # a cycle is cut where the first seed meets it
        n = fb.fm6_unitig(idx, 50, out)
        assert n == len(ref)
        assert H.canonical_mag(H.parse_mag(open(out).read())) == H.canonical_mag(ref), "host walk" if host == "1" else "device assembly"
    idx.close()


def test_unitig_small_cycles_take_the_host_walk(fb, tmp_path, monkeypatch):
    """Forty 300 bp plasmids tiled by reads every 25 bp, next to an ordinary linear genome: link-graph cycles of 12 reads.  Some hold
    a splitter of the list ranking (one read in 16) and keep the pointer jumping from converging, some hold none and are never
    reached by a walk (unitig_gpu.cu: k_rank_walk); either way the device assembly must notice (count check / round limit) and the
    records go to the host walk, whose result equals the reference's."""
    rng = np.random.RandomState(5)
    reads = [fb.synth_reads(52, fb.synth_genome(51, 40000), 4000, 100, 0.0)]
    for _ in range(40):
        pl = rng.randint(1, 5, 300).astype(np.uint8)
        ring = np.concatenate([pl, pl[:100]])
        reads.append(np.stack([ring[s: s + 100] for s in range(0, 300, 25)]))
    reads = np.concatenate(reads)
    fmd = fb.fm_build(fb.fmd_text(reads), 0)
    fn = str(tmp_path / "p.fmd")
    fmd.dump(fn)
    ref = H.parse_mag(H.reference_unitig(fn, 50, 1))
    idx = fb.FmdIndex(fmd, 0)
    out = str(tmp_path / "p.mag")
    monkeypatch.setenv("FMG_THREADS", "1")
    monkeypatch.setenv("FMG_UNITIG_HOST", "0")
    n = fb.fm6_unitig(idx, 50, out)
    assert n == len(ref)
    assert H.canonical_mag(H.parse_mag(open(out).read())) == H.canonical_mag(ref)
    idx.close()


@pytest.mark.parametrize("err,shards", [(0.0, 2), (0.01, 3)])
def test_sharded_records_merge_to_the_single_gpu_unitigs(fb, tmp_path, err, shards):
    """The multi-GPU data path on one GPU, stage by stage through the C-ABI: the records of row shards (fmg_overlap_shard) in
    separate device buffers, laid side by side with padding as the NCCL all-gather of fermi_b200.parallel.unitig_distributed_device
    leaves them, merged into the rank-indexed array (fmg_overlap_merge), the deferred left check (fmg_overlap_left_fix), and the
    unitigs assembled part by part (fmg_unitig_part + fmg_magpart_write into one file) == fm6_unitig."""
    import ctypes as C
    import torch
    from fermi_b200._lib import lib
    from fermi_b200.parallel import unitig_shard_rows
    L = lib()
    genome = fb.synth_genome(81, 300000)
    reads = fb.synth_reads(82, genome, 30000, 100, err)
    idx = fb.FmdIndex(fb.fm_build(fb.fmd_text(reads), 0), 0)
    n_seq = int(idx.fmd.mcnt[1])
    dev = torch.device("cuda", 0)
    recs, ranks, exts, spills, tots, rows = [], [], [], [], [], []
    for r in range(shards):
        lo, hi = unitig_shard_rows(n_seq, r, shards)
        rec = torch.empty((hi - lo) * 8, dtype=torch.int64, device=dev)
        rnk = torch.empty(hi - lo, dtype=torch.int64, device=dev)
        ext_cap, spill_cap = 64, 8                                     # too small on purpose: the call must report the need
        tot = (C.c_uint64 * 2)()
        for attempt in range(3):
            ext = torch.empty(ext_cap, dtype=torch.uint8, device=dev)
            spill = torch.empty(spill_cap * 4, dtype=torch.int64, device=dev)
            rc = L.fmg_overlap_shard(idx.h, 50, 0, lo, hi, rec.data_ptr(), rnk.data_ptr(), ext.data_ptr(), ext_cap, spill.data_ptr(), spill_cap, tot)
            if rc != 1:
                break
            ext_cap, spill_cap = int(tot[0]), int(tot[1])
        assert rc == 0 and attempt == 1
        recs.append(rec); ranks.append(rnk); exts.append(ext[: int(tot[0])]); spills.append(spill[: 4 * int(tot[1])]); tots.append((int(tot[0]), int(tot[1])))
        rows.append(hi - lo)
    row_pad, ext_pad, spill_pad = max(rows) + 3, max(t[0] for t in tots) + 5, max(t[1] for t in tots) + 1

    def side_by_side(parts, pad, width, dtype):
        out = torch.zeros(shards * pad * width, dtype=dtype, device=dev)
        for r, p in enumerate(parts):
            out[r * pad * width: r * pad * width + len(p)] = p
        return out

    rec_all, rank_all = side_by_side(recs, row_pad, 8, torch.int64), side_by_side(ranks, row_pad, 1, torch.int64)
    ext_all, spill_all = side_by_side(exts, ext_pad, 1, torch.uint8), side_by_side(spills, spill_pad, 4, torch.int64)
    assert sorted(torch.cat(ranks).cpu().tolist()) == list(range(n_seq))
    pack = torch.empty(n_seq * 8, dtype=torch.int64, device=dev)
    rank_of_row = torch.empty(n_seq, dtype=torch.int64, device=dev)
    assert L.fmg_overlap_merge(idx.h, shards, (C.c_uint64 * shards)(*rows), row_pad, ext_pad, spill_pad, rec_all.data_ptr(), rank_all.data_ptr(),
                               pack.data_ptr(), rank_of_row.data_ptr()) == 0
    assert torch.equal(rank_of_row, torch.cat(ranks))
    # the deferred left check shared between the ranks: each fixes its own rows on ITS copy of the merged array and hands one byte
    # per row round (fmg_overlap_left_fix_rows + fmg_overlap_left_flags) == the whole-array call
    merged = pack.clone()
    n_left = C.c_uint64()
    assert L.fmg_overlap_left_fix(idx.h, 50, 0, pack.data_ptr(), rank_of_row.data_ptr(), C.byref(n_left)) == 0
    flags, n_rows_fixed, first = [], 0, 0
    for r in range(shards):
        mine = merged.clone()
        k = C.c_uint64()
        assert L.fmg_overlap_left_fix_rows(idx.h, 50, 0, mine.data_ptr(), rank_of_row.data_ptr(), first, first + rows[r], C.byref(k)) == 0
        n_rows_fixed += int(k.value)
        f = torch.zeros(rows[r], dtype=torch.int8, device=dev)
        assert L.fmg_overlap_left_flags(idx.h, mine.data_ptr(), rank_of_row.data_ptr(), first, first + rows[r], f.data_ptr(), 0) == 0
        flags.append(f)
        first += rows[r]
    assert n_rows_fixed == int(n_left.value) and (err == 0.0 or n_rows_fixed > 0)
    first = 0
    for r in range(shards):
        assert L.fmg_overlap_left_flags(idx.h, merged.data_ptr(), rank_of_row.data_ptr(), first, first + rows[r], flags[r].data_ptr(), 1) == 0
        first += rows[r]
    torch.cuda.synchronize()
    assert torch.equal(merged, pack)
    out, single = str(tmp_path / "m.mag"), str(tmp_path / "s.mag")
    total, offset = 0, 0
    for part in range(shards):
        h, nu, nb = C.c_void_p(), C.c_uint64(), C.c_uint64()
        rc = L.fmg_unitig_part(idx.h, 50, pack.data_ptr(), rank_of_row.data_ptr(), ext_all.data_ptr(), spill_all.data_ptr(), part, shards,
                               C.byref(h), C.byref(nu), C.byref(nb))
        assert rc == 0
        assert L.fmg_magpart_write(h, out.encode(), offset, offset + nb.value) == 0      # the file grows part by part here
        L.fmg_magpart_free(h)
        total += nu.value
        offset += nb.value
    n = fb.fm6_unitig(idx, 50, single)
    assert os.path.getsize(out) == offset
    assert total == n and H.canonical_mag(H.parse_mag(open(out).read())) == H.canonical_mag(H.parse_mag(open(single).read()))
    idx.close()


def test_unitig_over_two_gpus_through_nccl(tmp_path):
    """The real multi-GPU path (one process per GPU, NCCL all-gather of the record shards, every rank assembling and writing its
    part of the unitigs): tools/unitig_multi.py --check compares the MAG set with the single-GPU fm6_unitig."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    for err in ("0.0", "0.01"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
               os.path.join(H.ROOT, "tools", "unitig_multi.py"), "--reads", "40000", "--err", err, "--check", "--iters", "1"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        assert res.returncode == 0, res.stderr[-2000:]
        line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
        assert line["n_gpus"] == 2 and line["set_equal_single_gpu"] is True


@pytest.mark.skipif(H.ref_fermi_binary() is None, reason="oracle/_ref/fermi did not travel with the repo")
@pytest.mark.parametrize("err,dup", [(0.0, 0), (0.01, 300)])
def test_seqrank_equals_reference_seqsort(fb, tmp_path, monkeypatch, err, dup):
    """fmg_seqsort (fm6_retrieve of every read on the GPU) == the bytes `fermi seqsort` writes (seqsort.c:12-70): ranks,
    contained and duplicate flags, with duplicated reads, palindromes and reads contained in longer ones in the input."""
    genome = fb.synth_genome(51, 150000)
    reads = list(fb.synth_reads(52, genome, 15000, 100, err))
    reads += [reads[i].copy() for i in range(dup)]                                   # exact duplicates
    reads += [r[10:70].copy() for r in reads[:200]]                                  # contained in a longer read
    pal = np.array([1, 2, 3, 4] * 10, np.uint8)
    reads.append(np.concatenate([pal, (5 - pal)[::-1]]))                             # equals its own reverse complement
    fmd = fb.fm_build(H.fmd_text(reads), 0)
    fn = str(tmp_path / "r.fmd")
    fmd.dump(fn)
    ref = np.frombuffer(subprocess.run([H.ref_fermi_binary(), "seqsort", "-t", "4", fn], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                       check=True).stdout, np.uint64)
    idx = fb.FmdIndex(fmd, 0)
    for wide in ("", "1"):
        if wide:
            monkeypatch.setenv("FMG_FORCE_WIDE", "1")
        ours, st = fb.fm6_seqsort(idx)
        assert np.array_equal(ours, ref)
        assert st == (int((ref == 0).sum()), int(((ref & 2) != 0).sum()), int((((ref & 2) == 0) & ((ref & 1) != 0) & (ref != 0)).sum()))
    idx.close()
    cli = os.path.join(H.ROOT, "fermi_b200", "bin", "fermi-b200")
    out = subprocess.run([cli, "seqrank", fn], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    assert out == ref.tobytes()


def _same_fmd(a, b, tmp_path):
    fa, fb_ = str(tmp_path / "a.fmd"), str(tmp_path / "b.fmd")
    a.dump(fa)
    b.dump(fb_)
    return open(fa, "rb").read() == open(fb_, "rb").read()


def test_device_rld_encoder_is_byte_identical(fb, tmp_path):
    """rld_enc.cu (speculative block chain + one thread per block) writes the same .fmd bytes as the serial host encoder
    (and hence as the reference, tests/test_oracle.py): golden BWTs, one-symbol and tiny inputs, runs >= 0x8000 symbols
    (7 x u32 block headers), 1-symbol runs only, and a stream longer than one 2^23-word chunk (shortened chunk-end blocks)."""
    rng = np.random.RandomState(9)
    cases = [np.array([0], np.uint8), np.array([3, 3, 3, 0], np.uint8), rng.randint(0, 6, size=777).astype(np.uint8)]
    for case in golden_cases():
        cases.append(fb.Fmd.restore(os.path.join(H.GOLDEN_DIR, case + ".fmd")).decode_bwt())
    long_runs = np.concatenate([np.full(l, s, np.uint8) for l, s in zip(rng.choice([1, 5, 300, 40000, 70000, 1 << 20], size=400), rng.randint(0, 6, size=400))])
    cases.append(long_runs)
    cases.append(np.tile(np.array([1, 2, 3, 4, 0, 5], np.uint8), 200000))                       # no run longer than 1: 4-bit codes
    cases.append(rng.randint(1, 5, size=150_000_000).astype(np.uint8))                           # ~80 MB of stream: crosses a chunk end
    for bwt in cases:
        dev, host = fb.Fmd.from_bwt_device(bwt, 0), fb.Fmd.from_bwt(bwt)
        assert dev.n_bytes == host.n_bytes and dev.n_frames == host.n_frames and list(dev.mcnt) == list(host.mcnt)
        assert _same_fmd(dev, host, tmp_path), "n=%d" % len(bwt)
    assert cases[-1].size and fb.Fmd.from_bwt(cases[-1]).n_bytes > (1 << 26)


def test_build_and_bcr_with_device_encoder(fb, tmp_path):
    """fm_build / Bcr.build_fmd keep the BWT on the device (suffix sort or BCR, then rld_enc.cu): same .fmd bytes as the
    host-encoded path and as the golden file the reference built."""
    g, fmd = _load("reads10x")
    text = g["text"]
    assert _same_fmd(fb.fm_build(text, 0), fb.fm_build(text, 0, host_encode=True), tmp_path)
    assert open(str(tmp_path / "a.fmd"), "rb").read() == open(fmd, "rb").read()
    reads = text[text != 0].reshape(-1, 100)[0::2]
    b = fb.Bcr(0)
    both = np.empty((2 * len(reads), 100), np.uint8)
    both[0::2], both[1::2] = reads, 5 - reads[:, ::-1]
    b.append_batch(both)
    e = b.build_fmd()
    e.dump(str(tmp_path / "c.fmd"))
    assert open(str(tmp_path / "c.fmd"), "rb").read() == open(fmd, "rb").read()


def test_bcr_bwt_equals_suffix_sort_and_reference(fb, tmp_path):
    """GPU BCR (fmg_bcr_*) == naive BWT on small/ragged inputs, == the GPU suffix-sort builder on 40k reads, and the
    RLD-encoded result is byte-identical to the golden .fmd the reference built with SA-IS."""
    rng = np.random.RandomState(3)
    # ragged lengths, duplicates, a 1-base sequence
    seqs = [rng.randint(1, 5, size=rng.randint(1, 60)).astype(np.uint8) for _ in range(300)]
    seqs[5] = seqs[3].copy()
    seqs[9] = np.array([2], np.uint8)
    b = fb.Bcr(0)
    for s_ in seqs:
        b.append(s_)
    b.build()
    text = np.concatenate([np.concatenate([s_, [0]]) for s_ in seqs]).astype(np.uint8)
    assert np.array_equal(b.bwt(), H.naive_bwt(text))
    # byte-RLE stream decodes to the same BWT (ropebwt.c:127-144 / rld.c:295-309)
    rle = b.rle()
    assert np.array_equal(fb.Fmd.from_rle6(rle).decode_bwt(), b.bwt())
    b.close()
    # golden: reads10x.fmd was built by the reference from r0 $ rc(r0) $ ...
    g, fmd = _load("reads10x")
    reads = g["text"][g["text"] != 0].reshape(-1, 100)[0::2]
    out = str(tmp_path / "bcr.fmd")
    fb.Fmd.from_bwt(fb.fm_ropebwt(reads, 0)).dump(out)
    assert open(out, "rb").read() == open(fmd, "rb").read()
    # 40k x 101 bp
    genome = fb.synth_genome(51, 400000)
    reads = fb.synth_reads(52, genome, 40000, 101, 0.005)
    assert np.array_equal(fb.fm_ropebwt(reads, 0), fb.fm_build_bwt(fb.fmd_text(reads), 0))


@pytest.mark.parametrize("case", golden_cases())
def test_ec_collect_matches_reference(fb, case, monkeypatch):
    """fm6_traverse + ec_collect (correct.c:35-87) as a frontier expansion: identical (suffix, key, val) triples and counters."""
    g, fmd = _load(case)
    idx = fb.FmdIndex(fb.Fmd.restore(fmd), 0)
    for w, mo, k in ((-1, 3, "ec_a"), (12, 2, "ec_b")):
        tri, cnt = fb.fm6_ec_collect(idx, w, mo)
        assert np.array_equal(tri, g[k]) and list(cnt) == list(g[k + "_cnt"])
    monkeypatch.setenv("FMG_FORCE_WIDE", "1")
    tri, cnt = fb.fm6_ec_collect(idx, -1, 3)
    assert np.array_equal(tri, g["ec_a"])
    idx.close()


def test_ec_collect_parts_add_up(fb):
    """the per-GPU shares of the k-mer collection (suffix s goes to part s % n_parts) are disjoint and their union -- triples and
    counters -- is the single-GPU result (which the tests above pin to the reference)."""
    genome = fb.synth_genome(73, 400000)
    reads = fb.synth_reads(74, genome, 40000, 100, 0.01)
    idx = fb.FmdIndex(fb.fm_build(fb.fmd_text(reads), 0), 0)
    full, cnt = fb.fm6_ec_collect(idx, 18, 2)
    assert len(full) > 100000
    for n_parts in (2, 3, 8):
        parts = [fb.fm6_ec_collect(idx, 18, 2, part=r, n_parts=n_parts) for r in range(n_parts)]
        tri = np.sort(np.concatenate([p[0] for p in parts]))
        assert np.array_equal(tri, full)
        assert sum(p[1][0] for p in parts) == cnt[0] and sum(p[1][1] for p in parts) == cnt[1]
        assert all(len(p[0]) > 0 for p in parts)
        suf = [set(((p[0] >> np.uint64(40)) % np.uint64(n_parts)).tolist()) for p in parts]
        assert all(s_ == {r} for r, s_ in enumerate(suf))
    idx.close()


@pytest.mark.skipif(H.ref_fermi_binary() is None, reason="oracle/_ref/fermi did not travel with the repo")
def test_merge_is_byte_identical_to_fermi_merge(fb, tmp_path):
    """`fermi merge a.fmd b.fmd` (fm_compute_gap_bits + fm_merge, merge.c:31-137): gap vector, interleaving and RLD encoding on the
    GPU give the reference's bytes; the merged index is the index of the concatenated collection (cmd.c:366-370); the command-line
    front end writes the same file."""
    genome = fb.synth_genome(95, 80000)
    a = fb.synth_reads(96, genome, 5000, 100, 0.01)
    b = fb.synth_reads(97, genome, 3000, 73, 0.0)                     # another read length, fewer reads
    fa, fbn, ref_out = str(tmp_path / "a.fmd"), str(tmp_path / "b.fmd"), str(tmp_path / "ref.fmd")
    ea, eb = fb.fm_build(fb.fmd_text(a), 0), fb.fm_build(fb.fmd_text(b), 0)
    ea.dump(fa); eb.dump(fbn)
    with open(ref_out, "wb") as fh:
        subprocess.run([H.ref_fermi_binary(), "merge", fa, fbn], stdout=fh, stderr=subprocess.DEVNULL, check=True)
    ours = str(tmp_path / "ours.fmd")
    fb.fm_merge(ea, eb, 0).dump(ours)
    assert open(ours, "rb").read() == open(ref_out, "rb").read()
    # the gap vector itself: as many set bits as the second index has symbols, and consistent with the merged BWT
    ia, ib = fb.FmdIndex(ea, 0), fb.FmdIndex(eb, 0)
    bits = fb.fm_gap_bits(ia, ib)
    n0, n1 = int(ea.mcnt[0]), int(eb.mcnt[0])
    flat = np.unpackbits(bits.view(np.uint8), bitorder="little")[: n0 + n1].astype(bool)
    assert flat.sum() == n1
    merged = fb.Fmd.restore(ours).decode_bwt()
    assert np.array_equal(merged[flat], eb.decode_bwt()) and np.array_equal(merged[~flat], ea.decode_bwt())
    ia.close(); ib.close()
    cli = os.path.join(H.ROOT, "fermi_b200", "bin", "fermi-b200")
    out2 = str(tmp_path / "cli.fmd")
    subprocess.run([cli, "merge", "-o", out2, fa, fbn], stderr=subprocess.DEVNULL, check=True)
    assert open(out2, "rb").read() == open(ref_out, "rb").read()


def test_rank_on_the_rld_stream_itself(fb):
    """fmg_rldx_*: the .fmd blocks stay run-length / Elias-delta coded in HBM and a warp decodes them (ballot search of the block,
    TMA-staged block + directory line, pointer doubling over the bit offsets, warp prefix sums): rank2a and fm6_extend equal the
    reference's golden vectors; on streams with 7 x u32 block headers (runs >= 0x8000), 4-bit codes only and a stream longer than
    one 2^23-word chunk (shortened chunk-end blocks) the ranks equal those of the occ-block path at random and boundary positions."""
    for case in golden_cases():
        g, fmd = _load(case)
        x = fb.RldIndex(fb.Fmd.restore(fmd), 0)
        ok, ol = x.rank2a(g["k"], g["l"])
        assert np.array_equal(ok, g["ok"]) and np.array_equal(ol, g["ol"]), case
        assert np.array_equal(x.extend(g["ik"], g["is_back"]), g["ext"]), case
        x.close()
    rng = np.random.RandomState(13)
    long_runs = np.concatenate([np.full(l, s, np.uint8) for l, s in zip(rng.choice([1, 5, 300, 40000, 70000, 1 << 20], size=300), rng.randint(0, 6, size=300))])
    cases = [long_runs, np.tile(np.array([1, 2, 3, 4, 0, 5], np.uint8), 100000), rng.randint(1, 5, size=150_000_000).astype(np.uint8)]
    for bwt in cases:
        e = fb.Fmd.from_bwt_device(bwt, 0)
        n = len(bwt)
        x, idx = fb.RldIndex(e, 0), fb.FmdIndex(e, 0)
        k = np.concatenate([rng.randint(0, n, 20000).astype(np.uint64), np.array([2**64 - 1, 0, 1, n - 2, n - 1], np.uint64)])
        l = np.minimum(k + rng.randint(0, 3000, len(k)).astype(np.uint64), np.uint64(n - 1))
        l[k == np.uint64(2**64 - 1)] = 17
        ok, ol = x.rank2a(k, l)
        ok2, ol2 = fb.rld_rank2a(idx, k, l)
        assert np.array_equal(ok, ok2) and np.array_equal(ol, ol2), n
        x.close(); idx.close()


@pytest.mark.skipif(H.reference() is None, reason="needs the compiled reference (oracle/_ref)")
def test_contrast_and_gap_bits_equal_the_reference_functions(fb, tmp_path):
    """fm6_contrast (cmp.c:94-126) on two read sets from genomes that differ by a few substitutions, and fm_compute_gap_bits
    (merge.c:68-94): the bitmaps of the GPU equal those of the reference's own functions (oracle/ref_harness.c)."""
    R = H.reference()
    g0 = fb.synth_genome(101, 40000)
    g1 = g0.copy()
    rng = np.random.RandomState(5)
    for p in rng.choice(len(g1), 25, replace=False):
        g1[p] = 1 + (g1[p] % 4)                                       # another base
    r0 = fb.synth_reads(102, g0, 12000, 100, 0.002)
    r1 = fb.synth_reads(103, g1, 12000, 100, 0.002)
    f0, f1 = str(tmp_path / "0.fmd"), str(tmp_path / "1.fmd")
    e0, e1 = fb.fm_build(fb.fmd_text(r0), 0), fb.fm_build(fb.fmd_text(r1), 0)
    e0.dump(f0); e1.dump(f1)
    i0, i1 = fb.FmdIndex(e0, 0), fb.FmdIndex(e1, 0)
    h0, h1 = R.load(f0), R.load(f1)
    for k, min_occ in ((25, 3), (31, 2)):
        s0, s1 = fb.fm6_contrast(i0, i1, k, min_occ)
        o0, o1 = R.contrast(h0, h1, int(e0.mcnt[1]), int(e1.mcnt[1]), k, min_occ, 3)
        assert np.array_equal(s0, o0) and np.array_equal(s1, o1)
        assert int(np.unpackbits(s0.view(np.uint8)).sum()) > 0 and int(np.unpackbits(s1.view(np.uint8)).sum()) > 0
    bits = fb.fm_gap_bits(i0, i1)
    assert np.array_equal(bits, R.gap_bits(h0, h1, int(e0.mcnt[0]) + int(e1.mcnt[0]), 2))
    R.destroy(h0); R.destroy(h1)
    i0.close(); i1.close()


DROP_BIN = os.path.join(H.ORACLE_DIR, "_ref", "fermi_drop")


@pytest.mark.skipif(not os.path.exists(DROP_BIN), reason="oracle/_ref/fermi_drop did not travel with the repo (make -C oracle drop)")
def test_reference_binary_with_the_library_dropped_in(fb, tmp_path):
    """The drop-in boundary, compiled: oracle/_ref/fermi_drop is the reference's own main.c, ropebwt.c (UNMODIFIED, linked against
    the bcr_* symbols of libfermi_b200 instead of bcr.c) and cmd.c with integration/main_exact.patch applied.  `exact` must print the
    bytes the reference printed (golden files), `ropebwt -a bcr` the BWT the reference prints."""
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(H.ROOT, "fermi_b200", "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    for case in ("genome", "noisy", "reads10x"):
        fmd, fa = os.path.join(H.GOLDEN_DIR, case + ".fmd"), os.path.join(H.GOLDEN_DIR, case + ".query.fa")
        res = subprocess.run([DROP_BIN, "exact", fmd, fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
        assert res.returncode == 0, res.stderr.decode()[-1500:]
        assert res.stdout == open(os.path.join(H.GOLDEN_DIR, case + ".exact.txt"), "rb").read(), case
    ref_bin = H.ref_fermi_binary()
    if ref_bin is None:
        pytest.skip("oracle/_ref/fermi did not travel with the repo")
    genome = fb.synth_genome(91, 60000)
    reads = fb.synth_reads(92, genome, 6000, 100, 0.005)
    reads[17, 40] = 5                                          # an N: ropebwt -N cuts the read there (ropebwt.c:107-116)
    fa = str(tmp_path / "r.fa")
    tab = np.array(list("$ACGTN"))
    with open(fa, "w") as fh:
        for i, r in enumerate(reads):
            fh.write(">%d\n%s\n" % (i, "".join(tab[r])))
    for flags in (["-a", "bcr", "-N"], ["-a", "bcr", "-N", "-b"], ["-a", "bcr", "-N", "-R"]):
        ours = subprocess.run([DROP_BIN, "ropebwt"] + flags + [fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
        ref = subprocess.run([ref_bin, "ropebwt"] + flags + [fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert ours.returncode == 0 and ref.returncode == 0, ours.stderr.decode()[-1500:]
        if "-b" in flags:                                      # binary run-length bytes: the runs may be cut differently, the BWT they spell may not
            def spell(b):
                assert b[:4] == b"RLE\6"
                a = np.frombuffer(b[4:], np.uint8)
                return np.repeat(a & 7, a >> 3)
            assert np.array_equal(spell(ours.stdout), spell(ref.stdout))
        else:
            assert ours.stdout == ref.stdout
    # `fermi remap` (fm6_remap / paircov, smem.c:139-394) with the SMEMs of the contigs from the GPU (integration/remap_smem.patch):
    # unpaired, paired with the rank table of seqrank, and paired + broken at low paired coverage (-c)
    frag = 300
    starts = np.random.RandomState(8).randint(0, len(genome) - frag, 4000)
    pe = np.empty((8000, 100), np.uint8)
    for i, st in enumerate(starts):                            # mates of a pair are consecutive reads, facing each other
        pe[2 * i] = genome[st: st + 100]
        pe[2 * i + 1] = (5 - genome[st + frag - 100: st + frag][::-1])
    pfmd, prank, pmag = str(tmp_path / "pe.fmd"), str(tmp_path / "pe.rank"), str(tmp_path / "pe.mag")
    fb.fm_build(fb.fmd_text(pe), 0).dump(pfmd)
    with open(prank, "wb") as fh:
        subprocess.run([ref_bin, "seqrank", pfmd], stdout=fh, stderr=subprocess.DEVNULL, check=True)
    with open(pmag, "wb") as fh:
        subprocess.run([ref_bin, "unitig", "-l", "40", pfmd], stdout=fh, stderr=subprocess.DEVNULL, check=True)
    assert os.path.getsize(pmag) > 50000
    for opts in ([], ["-r", prank], ["-r", prank, "-c", "2", "-D", "600"], ["-r", prank, "-t", "3"]):
        ours = subprocess.run([DROP_BIN, "remap"] + opts + [pfmd, pmag], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
        ref = subprocess.run([ref_bin, "remap"] + opts + [pfmd, pmag], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        assert ours.returncode == 0 and ref.returncode == 0, ours.stderr.decode()[-1500:]
        if "-t" in opts:                                       # several threads write whole records (four lines each) in any order
            def split(b):                                      # not on "\n@": a coverage line may start with '@' (33 + 31)
                ln = b.split(b"\n")
                assert ln[-1] == b"" and len(ln) % 4 == 1
                return sorted(tuple(ln[i:i + 4]) for i in range(0, len(ln) - 1, 4))
            assert split(ours.stdout) == split(ref.stdout)
        else:
            assert len(ref.stdout) > 50000 and ours.stdout == ref.stdout, opts
        pick = lambda e: [l for l in e.decode().splitlines() if "fm6_remap" in l]
        assert pick(ours.stderr) == pick(ref.stderr)           # "[M::fm6_remap] avg = ... std = ... cap = ..."
    # `fermi correct`: collect phase on the GPU (integration/correct_collect.patch), the reference's own fix phase
    reads = fb.synth_reads(93, genome[:25000], 10000, 100, 0.01)
    fq, fn = str(tmp_path / "c.fq"), str(tmp_path / "c.fmd")
    H.write_fastq(fq, reads)
    fb.fm_build(fb.fmd_text(reads), 0).dump(fn)
    ours = subprocess.run([DROP_BIN, "correct", "-t", "1", fn, fq], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
    ref = subprocess.run([ref_bin, "correct", "-t", "1", fn, fq], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert ours.returncode == 0 and ref.returncode == 0, ours.stderr.decode()[-1500:]
    assert len(ref.stdout) > 100000 and ours.stdout == ref.stdout


@pytest.mark.skipif(H.reference() is None or H.ref_fermi_binary() is None, reason="needs the compiled reference (oracle/_ref)")
def test_gpu_kmers_drive_the_reference_fix_phase_to_the_same_fastq(fb, tmp_path):
    """`fermi correct` with the collect phase on the GPU: fmg_ec_collect's triples are loaded into the
    reference's solid[] tables, the reference's own ec_fix corrects the reads (oracle/ref_harness_correct.c), and the FASTQ is
    byte-identical to `fermi correct -t1`."""
    R = H.reference()
    genome = fb.synth_genome(33, 30000)
    reads = fb.synth_reads(34, genome, 12000, 100, 0.01)
    fq, fn = str(tmp_path / "r.fq"), str(tmp_path / "r.fmd")
    H.write_fastq(fq, reads)
    fmd = fb.fm_build(fb.fmd_text(reads), 0)
    fmd.dump(fn)
    idx = fb.FmdIndex(fmd, 0)
    tri, cnt = fb.fm6_ec_collect(idx, -1, 3)
    idx.close()
    h = R.load(fn)
    out = str(tmp_path / "fixed.fq")
    R.ec_fix_from_triples(h, -1, 3, tri, fq, out)
    R.destroy(h)
    ref = subprocess.run([H.ref_fermi_binary(), "correct", "-t", "1", fn, fq], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    assert len(ref) > 100000 and open(out, "rb").read() == ref


def test_ec_collect_100k_reads_vs_oracle(fb, oracle, tmp_path):
    genome = fb.synth_genome(71, 500000)
    reads = fb.synth_reads(72, genome, 50000, 100, 0.01)
    fmd = fb.fm_build(fb.fmd_text(reads), 0)
    fn = str(tmp_path / "r.fmd")
    fmd.dump(fn)
    idx = fb.FmdIndex(fmd, 0)
    ho = oracle.load(fn)
    tri, cnt = fb.fm6_ec_collect(idx, -1, 3)
    otri, ocnt, _ = oracle.ec_collect(ho, -1, 3)
    assert len(tri) > 100000 and np.array_equal(tri, otri) and tuple(cnt) == tuple(ocnt)
    oracle.destroy(ho)
    idx.close()


def test_command_line_front_end(fb, tmp_path):
    """fermi-b200 exact / unitig / build / ropebwt+recode against the text the reference's commands print."""
    import subprocess
    exe = os.path.join(H.ROOT, "fermi_b200", "bin", "fermi-b200")
    case = "reads10x"
    g, fmd = _load(case)
    run = lambda *a: subprocess.run([exe] + list(a), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    # exact: byte-identical to `fermi exact` (cmd.c:292-333)
    assert run("exact", fmd, os.path.join(H.GOLDEN_DIR, case + ".query.fa")) == open(os.path.join(H.GOLDEN_DIR, case + ".exact.txt"), "rb").read()
    # unitig: same canonical MAG set as `fermi unitig -l50`
    ours = H.parse_mag(run("unitig", "-l", str(int(g["ov_min"])), fmd).decode())
    ref = H.parse_mag(open(os.path.join(H.GOLDEN_DIR, case + ".mag")).read())
    assert H.canonical_mag(ours) == H.canonical_mag(ref)
    # build and ropebwt -b | recode from FASTA: the reference's .fmd, byte for byte
    reads = g["text"][g["text"] != 0].reshape(-1, 100)[0::2]
    fa = str(tmp_path / "r.fa")
    tab = np.array(list("$ACGTN"))
    with open(fa, "w") as fh:
        for i, r in enumerate(reads):
            fh.write(">r%d\n%s\n" % (i, "".join(tab[r])))
    a, b, rle = str(tmp_path / "a.fmd"), str(tmp_path / "b.fmd"), str(tmp_path / "b.rle")
    run("build", "-fo", a, fa)
    run("ropebwt", "-a", "bcr", "-bN", "-o", rle, fa)
    run("recode", rle, b)
    c = str(tmp_path / "c.fmd")
    run("ropebwt", "-a", "bcr", "-rN", "-o", c, fa)            # BCR + RLD encoding on the GPU in one step
    want = open(fmd, "rb").read()
    assert open(a, "rb").read() == want and open(b, "rb").read() == want and open(c, "rb").read() == want


def test_gpu_transcode_equals_host_builder(fb, monkeypatch, tmp_path):
    """.fmd -> occ blocks on the GPU (RLD run decode per block) == the host transcoder, byte for byte;
    includes 32-bit block headers and very long runs."""
    rng = np.random.RandomState(5)
    parts = [np.full(rng.choice([1, 2, 7, 300, 40000, 70000]), rng.randint(0, 6), np.uint8) for _ in range(200)]
    long_runs = str(tmp_path / "long.fmd")
    fb.Fmd.from_bwt(np.concatenate(parts)).dump(long_runs)
    for fn in [os.path.join(H.GOLDEN_DIR, c + ".fmd") for c in golden_cases()] + [long_runs]:
        monkeypatch.delenv("FMG_HOST_OCC_BUILD", raising=False)
        dev = fb.FmdIndex(fb.Fmd.restore(fn), 0).export()
        monkeypatch.setenv("FMG_HOST_OCC_BUILD", "1")
        host = fb.FmdIndex(fb.Fmd.restore(fn), 0).export()
        assert np.array_equal(dev[0], host[0]) and np.array_equal(dev[1], host[1])


def test_full_size_smem_properties(fb, oracle, tmp_path):
    """BASELINE config 2 at full size (10 M x 100 bp reads, 100 Mbp genome index): the oracle cannot run all of it in
    seconds, so: exact equality on 20 k sampled reads, structural invariants on all ~86 M records, an independent
    re-derivation of 200 k records by backward search (a different kernel), and run-to-run identity."""
    import hashlib
    G, N, L = 100_000_000, 10_000_000, 100
    genome = fb.synth_genome(21, G)
    fmd = fb.fm_build(fb.fmd_text(genome[: G // 10000 * 10000].reshape(-1, 10000)), 0)
    fn = str(tmp_path / "g.fmd")
    fmd.dump(fn)
    idx = fb.FmdIndex(fmd, 0)
    reads = fb.synth_reads(22, genome, N, L, 0.01)
    seq, off = H.reads_to_flat(reads)
    cap = N * 16
    mem = np.zeros(cap, H.INTV)
    mo = np.zeros(N + 1, np.uint64)
    digests = []
    for _ in range(2):
        n = fb.fm6_smem_raw(idx, N, seq.ctypes.data, off.ctypes.data, mem.ctypes.data, cap, mo.ctypes.data, 0, 2_000_000)
        digests.append((n, hashlib.sha1(mem[:n].tobytes()).hexdigest(), hashlib.sha1(mo.tobytes()).hexdigest()))
    assert digests[0] == digests[1]
    rec = mem[:n]
    # (1) sampled reads, bit-exact vs the oracle
    ho = oracle.load(fn)
    for lo in (0, N - 10000, 4_999_000):
        sl = slice(lo, lo + 10000)
        orec, omo, _, _, _ = oracle.smem(ho, reads[sl].reshape(-1), (np.arange(10001, dtype=np.uint64) * np.uint64(L)), 0, 8)
        a, b = int(mo[lo]), int(mo[lo + 10000])
        assert np.array_equal(mo[lo: lo + 10001] - mo[lo], omo) and np.array_equal(rec[a:b], orec)
    oracle.destroy(ho)
    # (2) invariants over everything
    cnt = np.diff(mo.astype(np.int64))
    assert mo[0] == 0 and int(mo[-1]) == n and cnt.min() >= 1
    start = (rec["info"] >> np.uint64(32)) & np.uint64(0x3FFFFFFF)
    end = rec["info"] & np.uint64(0x3FFFFFFF)
    assert (start < end).all() and (end <= L).all() and (rec["x2"] >= 1).all()
    n_sym = np.uint64(fmd.mcnt[0])
    assert (rec["x0"] + rec["x2"] <= n_sym).all() and (rec["x1"] + rec["x2"] <= n_sym).all()
    # (3) 200 k records re-derived with fm_backward_search on the matched substring (exact.c:7): for a match that is not
    # closed by a sentinel on the right, [x0, x0+x2-1] is the SA interval of the substring
    rng = np.random.RandomState(1)
    pick = rng.randint(0, n, size=200000)
    pick = pick[rec["x1"][pick] >= np.uint64(fmd.mcnt[1])]
    owner = np.searchsorted(mo, pick, side="right") - 1
    s_, e_ = start[pick].astype(np.int64), end[pick].astype(np.int64)
    lens = e_ - s_
    sub_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    idxs = np.repeat(owner * L + s_, lens) + (np.arange(int(sub_off[-1])) - np.repeat(sub_off[:-1].astype(np.int64), lens))
    sub = seq[idxs]
    b, e, sz = fb.fm_backward_search(idx, sub, sub_off)
    assert np.array_equal(b, rec["x0"][pick]) and np.array_equal(sz, rec["x2"][pick])
    idx.close()


def test_plain_c_client_drives_the_device(tmp_path):
    """tests/c_abi/abi_check.c with "gpu": a C99 program (no Python, no torch in the process) uploads an index and calls
    fmg_rank1a_batch / fmg_check_rank / fmg_backward_search_batch through include/fermi_b200.h; the expected counts are taken
    from the BWT inside the C program."""
    import subprocess
    src = os.path.join(H.ROOT, "tests", "c_abi", "abi_check.c")
    exe = str(tmp_path / "abi_check")
    libdir = os.path.join(H.ROOT, "fermi_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(H.ROOT, "include"), "-o", exe, src,
                    "-L" + libdir, "-lfermi_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe, str(tmp_path / "t.fmd"), "gpu"], stdout=subprocess.PIPE)
    assert r.returncode == 0, "abi_check exit %d" % r.returncode
    out = r.stdout.decode()
    assert "gpu ok: 18 ranks" in out and " 0 kernel launches" not in out, out
