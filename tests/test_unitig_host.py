"""CPU: the overlap lane code (host build of fmd_overlap.cuh) and the product's unitig walker
(unitig_host.cpp, host code) against the records and the MAG output of the unmodified reference."""
import os

import numpy as np
import pytest

import helpers as H
from conftest import golden_cases


def _load(case):
    return np.load(os.path.join(H.GOLDEN_DIR, case + ".npz")), os.path.join(H.GOLDEN_DIR, case + ".fmd")


@pytest.mark.parametrize("case", golden_cases())
def test_overlap_records_match_reference(emu, case):
    g, fmd = _load(case)
    x = emu.index(fmd)
    max_len = int(g["ov_rec"][:, 1].max()) + 8
    for wide in (0, 1):
        rec, nei, off, seq, ln, ext = emu.overlap(x, int(g["ov_min"]), g["ov_seeds"], max_len, wide=wide)
        assert np.array_equal(rec[:, :9], g["ov_rec"])
        assert np.array_equal(nei, g["ov_nei"]) and np.array_equal(off, g["ov_off"])
    emu.lib.emu_index_free(x)


@pytest.mark.parametrize("case", golden_cases())
def test_unitig_walk_equals_reference_mag(emu, product_lib, case, tmp_path):
    """records of ALL sequences -> fmg_unitig_assemble -> the same unitig set as `fermi unitig -t1`."""
    import fermi_b200 as fb
    g, fmd = _load(case)
    x = emu.index(fmd)
    n_seq = int(g["mcnt"][1])
    max_len = int(g["ov_rec"][:, 1].max()) + 8
    rec, nei, off, seq, ln, ext = emu.overlap(x, int(g["ov_min"]), np.arange(n_seq, dtype=np.uint64), max_len, nei_cap=32)
    out = str(tmp_path / "u.mag")
    n = fb.fm6_unitig_assemble(n_seq, int(g["ov_min"]), dict(rec=rec, nei=nei, nei_off=off, seq=seq, ext=ext), out)
    ours = H.parse_mag(open(out).read())
    ref = H.parse_mag(open(os.path.join(H.GOLDEN_DIR, case + ".mag")).read())
    assert n == len(ours) == len(ref)
    assert H.canonical_mag(ours) == H.canonical_mag(ref)
    emu.lib.emu_index_free(x)


@pytest.mark.skipif(H.ref_fermi_binary() is None, reason="oracle/_ref/fermi not built")
@pytest.mark.parametrize("err,cov", [(0.0, 10), (0.01, 12), (0.0, 30)])
def test_unitig_walk_fresh_inputs_vs_reference_binary(emu, product_lib, tmp_path, err, cov):
    import fermi_b200 as fb
    R = H.reference()
    glen = 6000
    g = H.synth_genome(glen, 500 + cov)
    reads = H.synth_reads(g, glen * cov // 80, 80, err, 600 + cov)
    fmd = str(tmp_path / "f.fmd")
    h = R.build_text(H.fmd_text(reads))
    R.dump(h, fmd)
    R.destroy(h)
    ref_t1 = H.parse_mag(H.reference_unitig(fmd, 40, 1))
    ref_t4 = H.parse_mag(H.reference_unitig(fmd, 40, 4))
    assert H.canonical_mag(ref_t1) == H.canonical_mag(ref_t4)          # the parity rule of SURVEY.md section 4
    x = emu.index(fmd)
    n_seq = 2 * len(reads)
    rec, nei, off, seq, ln, ext = emu.overlap(x, 40, np.arange(n_seq, dtype=np.uint64), 96, nei_cap=32)
    out = str(tmp_path / "u.mag")
    fb.fm6_unitig_assemble(n_seq, 40, dict(rec=rec, nei=nei, nei_off=off, seq=seq, ext=ext), out)
    assert H.canonical_mag(H.parse_mag(open(out).read())) == H.canonical_mag(ref_t1)
    emu.lib.emu_index_free(x)


def test_threaded_walk_gives_the_same_unitig_set(emu, product_lib, tmp_path, monkeypatch):
    """many seeds per unitig (error-free 12x): 1 walker thread vs 7 must emit the same canonical set."""
    import fermi_b200 as fb
    g = fb.synth_genome(77, 40000)
    reads = fb.synth_reads(78, g, 6000, 80, 0.0)
    fmd = str(tmp_path / "f.fmd")
    fb.Fmd.from_bwt(H.naive_bwt(fb.fmd_text(reads))).dump(fmd)
    x = emu.index(fmd)
    n_seq = 2 * len(reads)
    rec, nei, off, seq, ln, ext = emu.overlap(x, 40, np.arange(n_seq, dtype=np.uint64), 88, nei_cap=32)
    o = dict(rec=rec, nei=nei, nei_off=off, seq=seq, ext=ext)
    sets = []
    for threads in ("1", "7", "7", "7"):
        monkeypatch.setenv("FMG_THREADS", threads)
        out = str(tmp_path / ("u%s.mag" % threads))
        fb.fm6_unitig_assemble(n_seq, 40, o, out)
        sets.append(H.canonical_mag(H.parse_mag(open(out).read())))
    assert len(sets[0]) > 10
    assert all(s == sets[0] for s in sets[1:])
    emu.lib.emu_index_free(x)
