"""CPU: the product's host logic (.fmd reader/encoder, synthetic data, C-ABI surface) and the host
compilation of its device-side core (occ-line rank, fm6_extend, SMEM lane state machine) against the
golden vectors.  No GPU compute is called here."""
import os
import re

import numpy as np
import pytest

import helpers as H
from conftest import golden_cases


def _load(case):
    return np.load(os.path.join(H.GOLDEN_DIR, case + ".npz")), os.path.join(H.GOLDEN_DIR, case + ".fmd")


def test_plain_c_client_links_and_runs(product_lib, tmp_path):
    """include/fermi_b200.h is valid C99 and a C program (what fermi's own mains are) links every entry point and drives the
    host-side container calls; tests/c_abi/abi_check.c also lists every declared symbol, checked against the header here."""
    import subprocess
    src = os.path.join(H.ROOT, "tests", "c_abi", "abi_check.c")
    hdr = open(os.path.join(H.ROOT, "include", "fermi_b200.h")).read()
    declared = set(re.findall(r"\b(fmg_[a-z0-9_]+)\s*\(", hdr))
    listed = set(re.findall(r"\(anyfn_t\)(fmg_[a-z0-9_]+)", open(src).read()))
    assert declared == listed, sorted(declared ^ listed)
    exe = str(tmp_path / "abi_check")
    libdir = os.path.join(H.ROOT, "fermi_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(H.ROOT, "include"), "-o", exe, src,
                    "-L" + libdir, "-lfermi_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe, str(tmp_path / "t.fmd")], stdout=subprocess.PIPE, check=True).stdout.decode()
    assert out.startswith("ok %d symbols" % len(declared))


def test_library_exports_every_declared_symbol(product_lib):
    hdr = open(os.path.join(H.ROOT, "include", "fermi_b200.h")).read()
    declared = set(re.findall(r"\b(fmg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(product_lib, name), name + " is declared in include/fermi_b200.h but not exported"


def test_device_entry_points_fail_loudly_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import fermi_b200 as fb
    f = fb.Fmd.restore(os.path.join(H.GOLDEN_DIR, "reads10x.fmd"))
    with pytest.raises(RuntimeError):
        fb.FmdIndex(f, 0)
    with pytest.raises(RuntimeError):
        fb.fm_build_bwt(np.array([1, 2, 0, 3, 4, 0], np.uint8))
    # the builders that encode on the device, BCR and the device RLD encoder: no CPU path either
    with pytest.raises((RuntimeError, IOError)):
        fb.fm_build(np.array([1, 2, 0, 3, 4, 0], np.uint8))
    with pytest.raises((RuntimeError, IOError)):
        fb.Fmd.from_bwt_device(np.array([1, 2, 0, 3, 4, 0], np.uint8))
    b = fb.Bcr(0)
    b.append(np.array([1, 2, 3], np.uint8))
    with pytest.raises(RuntimeError):
        b.build()
    b.close()


def test_unitig_assemble_rejects_inconsistent_records(product_lib, tmp_path):
    """fmg_unitig_assemble packs the records it is handed by rank: a rank outside the index is an error, not a crash."""
    import fermi_b200 as fb
    n = 4
    rec = np.zeros((n, 10), np.int64)
    rec[:, 0] = [0, 1, 2, 7]                     # rank 7 of 4 sequences
    rec[:, 1] = 60
    o = dict(rec=rec, nei=np.zeros(0, fb.INTV), nei_off=np.zeros(n + 1, np.uint64), seq=np.ones((n, 64), np.uint8), ext=np.zeros((n, 64), np.uint8))
    with pytest.raises(RuntimeError):
        fb.fm6_unitig_assemble(n, 50, o, str(tmp_path / "x.mag"))


@pytest.mark.parametrize("case", golden_cases())
def test_fmd_reader_and_encoder_are_byte_identical(product_lib, case, tmp_path):
    import fermi_b200 as fb
    g, fmd = _load(case)
    f = fb.Fmd.restore(fmd)
    assert np.array_equal(f.mcnt, g["mcnt"]) and np.array_equal(f.cnt, g["cnt"])
    assert (f.n_bytes, f.n_frames, f.ibits) == (int(g["n_bytes"]), int(g["n_frames"]), int(g["ibits"]))
    bwt = f.decode_bwt()
    assert np.array_equal(bwt, H.naive_bwt(g["text"]))
    out = str(tmp_path / "o.fmd")
    fb.Fmd.from_bwt(bwt).dump(out)
    assert open(out, "rb").read() == open(fmd, "rb").read()
    f.dump(out)
    assert open(out, "rb").read() == open(fmd, "rb").read()


def test_encoder_long_runs_and_32bit_headers(product_lib, oracle, tmp_path):
    """runs >= 0x8000 symbols force the 7 x u32 block header (rld.c:119-124)."""
    import fermi_b200 as fb
    rng = np.random.RandomState(5)
    parts = []
    for _ in range(200):
        parts.append(np.full(rng.choice([1, 2, 7, 300, 40000, 70000]), rng.randint(0, 6), np.uint8))
    bwt = np.concatenate(parts)
    a, b = str(tmp_path / "a.fmd"), str(tmp_path / "b.fmd")
    fb.Fmd.from_bwt(bwt).dump(a)
    h = oracle.from_bwt(bwt)
    oracle.dump(h, b)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert np.array_equal(fb.Fmd.restore(a).decode_bwt(), bwt)
    assert np.array_equal(oracle.decode_bwt(h), bwt)
    R = H.reference()
    if R is not None:      # the reference itself must read it and agree on ranks
        hr = R.load(a)
        k = rng.randint(0, len(bwt), size=3000).astype(np.uint64)
        l = np.minimum(k + np.uint64(50000), np.uint64(len(bwt) - 1))
        ra, oa = R.rank2a(hr, k, l), oracle.rank2a(h, k, l)
        assert np.array_equal(ra[0], oa[0]) and np.array_equal(ra[1], oa[1])
        R.destroy(hr)
    oracle.destroy(h)


def test_synthetic_data_matches_numpy_mirror(product_lib):
    import fermi_b200 as fb
    g = fb.synth_genome(7, 100003)
    assert np.array_equal(g, H.synth_genome(100003, 7))
    r = fb.synth_reads(11, g, 5000, 100, 0.02)
    assert np.array_equal(r, H.synth_reads(g, 5000, 100, 0.02, 11))
    r[3] = np.concatenate([r[3, :50], H.revcomp(r[3, :50])])      # even-length rc-palindrome
    assert np.array_equal(fb.fmd_text(r), H.fmd_text(r))


@pytest.mark.parametrize("case", golden_cases())
def test_device_core_on_host_matches_reference_vectors(emu, case):
    g, fmd = _load(case)
    x = emu.index(fmd)
    ok, ol = emu.rank2a(x, g["k"], g["l"])
    assert np.array_equal(ok, g["ok"]) and np.array_equal(ol, g["ol"])
    assert np.array_equal(emu.extend(x, g["ik"], g["is_back"]), g["ext"])
    seq, off = H.reads_to_flat(g["q"])
    for sm, rk, ok_ in ((0, "smem0", "moff0"), (1, "smem1", "moff1")):
        for wide in (0, 1):     # 32-bit and 64-bit coordinate instantiations of the lane
            rec, mo, ov = emu.smem(x, seq, off, sm, n_lanes=5, out_cap=128, wide=wide)
            assert ov == 0
            assert np.array_equal(mo, g[ok_]) and np.array_equal(rec, g[rk])
    # a slot capacity that is too small must be reported, never silently truncated
    rec, mo, ov = emu.smem(x, seq, off, 0, n_lanes=3, out_cap=2)
    assert ov == 1
    emu.lib.emu_index_free(x)


def test_device_core_ragged_and_empty_reads(emu, oracle):
    fmd = os.path.join(H.GOLDEN_DIR, "reads10x.fmd")
    g = np.load(os.path.join(H.GOLDEN_DIR, "reads10x.npz"))
    rng = np.random.RandomState(9)
    reads = [g["q"][i][: rng.randint(0, 101)] for i in range(200)]
    reads[0] = reads[0][:0]
    reads[17] = g["q"][17][:1]
    seq = np.concatenate(reads).astype(np.uint8)
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    x = emu.index(fmd)
    h = oracle.load(fmd)
    for sm in (0, 1):
        rec, mo, ov = emu.smem(x, seq, off, sm)
        orec, omo, _, _, _ = oracle.smem(h, seq, off, sm, 1)
        assert ov == 0 and np.array_equal(mo, omo) and np.array_equal(rec, orec)
    oracle.destroy(h)
    emu.lib.emu_index_free(x)


@pytest.mark.parametrize("k_per,n_threads", [(16, 64), (32, 64), (16, 256), (32, 128)])
def test_bcr_merge_tile_arithmetic_on_host(emu, k_per, n_threads):
    """The per-thread arithmetic of k_bcr_merge (fermi_b200/csrc/bcr_tile.cuh: byte-permute spread of the old symbols, blend of the
    inserts, bit-plane base counts, in-tile insert ranks) compiled for the host, against the obvious per-symbol merge: tiles
    without inserts, sparse, dense, nothing but inserts, a short last tile, every byte offset of the staged symbols."""
    rng = np.random.RandomState(k_per * 1000 + n_threads)
    tile = k_per * n_threads
    cases = []
    for density in (0.0, 0.004, 0.05, 0.5, 1.0):
        for tile_len in (tile, tile - 1, tile // 2 + 3, 5):
            cases.append((density, tile_len, int(rng.randint(0, 16))))
    for density, tile_len, shift in cases:
        flags = (rng.rand(tile_len) < density).astype(np.uint8)
        n_ins = int(flags.sum())
        syms = np.where(flags == 1, rng.randint(0, 6, tile_len), 0xaa).astype(np.uint8)      # inserts: $ A C G T N
        old = rng.randint(0, 6, tile_len - n_ins).astype(np.uint8)
        out, hist, ranks = emu.bcr_tile(k_per, n_threads, flags, syms, old, shift)
        want = np.empty(tile_len, np.uint8)
        want[flags == 1] = syms[flags == 1]
        want[flags == 0] = old
        assert np.array_equal(out, want), (density, tile_len, shift)
        assert [int(h) for h in hist] == [int((want == c).sum()) for c in (1, 2, 3, 4)]
        pos = np.flatnonzero(flags)
        want_rank = [int((want[:p] == want[p]).sum()) if 1 <= want[p] <= 4 else 0 for p in pos]
        assert ranks.tolist() == want_rank, (density, tile_len, shift)


def test_committed_traffic_figures_follow_from_the_committed_launch_list(tmp_path):
    """roofline.traffic in the bench lines is read from profiles/r02_k_*_traffic.json; those files must be what tools/ncu_traffic.py
    derives from the committed ncu launch list (profiles/r02_launches.csv: 4 unitig passes of 20 M sequences in the capture)."""
    import json, subprocess, sys
    prof = os.path.join(H.ROOT, "profiles")
    subprocess.run([sys.executable, os.path.join(H.ROOT, "tools", "ncu_traffic.py"), os.path.join(prof, "r02_launches.csv"), str(tmp_path), "r02",
                    "20000000", "80000000"], check=True, stdout=subprocess.DEVNULL)
    for name in ("r02_k_smem_traffic.json", "r02_k_smem_hbm_traffic.json", "r02_k_ov_traffic.json"):
        assert json.load(open(os.path.join(str(tmp_path), name))) == json.load(open(os.path.join(prof, name))), name
    assert open(os.path.join(str(tmp_path), "r02_launch_shares.csv")).read() == open(os.path.join(prof, "r02_launch_shares.csv")).read()
    # and the bench line of the same round quotes them
    line = json.load(open(os.path.join(prof, "r02_bench_n1.json")))
    assert line["roofline"]["traffic"] == pytest.approx(json.load(open(os.path.join(prof, "r02_k_smem_traffic.json")))["traffic_bytes_per_launch"], rel=1e-3)
    ov = json.load(open(os.path.join(prof, "r02_k_ov_traffic.json")))["traffic_bytes_per_sequence"]
    assert line["unitig"]["roofline"]["traffic"] == pytest.approx(ov * 20000000, rel=1e-6)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_bench_lines_keep_the_contract(n):
    """The bench lines under profiles/ are what bench.py printed: they must carry every key of the measurement contract."""
    import json
    d = json.load(open(os.path.join(H.ROOT, "profiles", "r02_bench_n%d.json" % n)))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "gpu_launches", "roofline", "e2e", "unitig", "unitig_err1pct"):
        assert k in d, k
    assert d["n_gpus"] == n and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and 0 < d["roofline"]["frac"] <= 1.0
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "reference"
        assert d["unitig"]["set_equal_reference_sample"] is True
    else:
        assert d["unitig"]["set_equal_single_gpu"] is True and d["unitig"]["collective_ms"] > 0
