"""CPU: the oracle (oracle/fmd_oracle.c) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py), and -- where oracle/_ref is built -- against the reference
itself on fresh seeded inputs."""
import os

import numpy as np
import pytest

import helpers as H
from conftest import golden_cases


def _load(case):
    return np.load(os.path.join(H.GOLDEN_DIR, case + ".npz")), os.path.join(H.GOLDEN_DIR, case + ".fmd")


@pytest.mark.parametrize("case", golden_cases())
def test_container_and_encoder(oracle, case, tmp_path):
    g, fmd = _load(case)
    h = oracle.load(fmd)
    info = oracle.info(h)
    assert np.array_equal(info["mcnt"], g["mcnt"]) and np.array_equal(info["cnt"], g["cnt"])
    assert (info["n_bytes"], info["n_frames"], info["ibits"]) == (int(g["n_bytes"]), int(g["n_frames"]), int(g["ibits"]))
    # decode -> re-encode must reproduce the reference's file byte for byte (rld_enc + rld_rank_index + rld_dump)
    bwt = oracle.decode_bwt(h)
    assert np.array_equal(bwt, H.naive_bwt(g["text"]))
    h2 = oracle.from_bwt(bwt)
    out = str(tmp_path / "re.fmd")
    oracle.dump(h2, out)
    assert open(out, "rb").read() == open(fmd, "rb").read()
    # the raw byte-RLE flavour (ropebwt.c:133, rld.c:295-309) re-encodes to the same file
    runs = []
    prev, n = bwt[0], 0
    for c in bwt:
        if c == prev and n < 31:
            n += 1
        else:
            runs.append(n << 3 | prev)
            prev, n = c, 1
    runs.append(n << 3 | prev)
    rle = str(tmp_path / "x.rle")
    with open(rle, "wb") as fh:
        fh.write(b"RLE\x06" + bytes(bytearray(runs)))
    h3 = oracle.load(rle)
    oracle.dump(h3, out)
    assert open(out, "rb").read() == open(fmd, "rb").read()
    for x in (h, h2, h3):
        oracle.destroy(x)


@pytest.mark.parametrize("case", golden_cases())
def test_queries_match_reference_vectors(oracle, case):
    g, fmd = _load(case)
    h = oracle.load(fmd)
    ok, ol = oracle.rank2a(h, g["k"], g["l"])
    assert np.array_equal(ok, g["ok"]) and np.array_equal(ol, g["ol"])
    ext = oracle.extend(h, g["ik"], g["is_back"])
    assert np.array_equal(ext, g["ext"])
    seq, off = H.reads_to_flat(g["q"])
    for sm, rk, ok_ in ((0, "smem0", "moff0"), (1, "smem1", "moff1")):
        rec, mo, _, nloc, next_ = oracle.smem(h, seq, off, sm, 3)
        assert np.array_equal(mo, g[ok_]) and np.array_equal(rec, g[rk])
        assert nloc >= next_ > 0
    b, e, s = oracle.backward_search(h, seq, off)
    assert np.array_equal(b, g["sa_beg"]) and np.array_equal(e, g["sa_end"]) and np.array_equal(s, g["sa_size"])
    for w, mo, k in ((-1, 3, "ec_a"), (12, 2, "ec_b")):
        tri, cnt, ww = oracle.ec_collect(h, w, mo)
        assert ww == int(g[k + "_w"]) and np.array_equal(tri, g[k]) and list(cnt) == list(g[k + "_cnt"])
    rec, nei, noff, _ = oracle.overlap(h, int(g["ov_min"]), g["ov_seeds"])
    assert np.array_equal(rec, g["ov_rec"]) and np.array_equal(nei, g["ov_nei"]) and np.array_equal(noff, g["ov_off"])
    oracle.destroy(h)


@pytest.mark.skipif(H.reference() is None, reason="oracle/_ref not built (reference sources absent)")
def test_oracle_vs_compiled_reference_fresh_inputs(oracle, tmp_path):
    R = H.reference()
    g = H.synth_genome(30000, 101)
    reads = H.synth_reads(g, 3000, 100, 0.005, 102)
    text = H.fmd_text(reads)
    hr = R.build_text(text)
    fn = str(tmp_path / "a.fmd")
    R.dump(hr, fn)
    ho = oracle.load(fn)
    n = int(oracle.info(ho)["mcnt"][0])
    rng = np.random.RandomState(3)
    k = rng.randint(0, n, size=50000).astype(np.uint64)
    l = np.minimum(k + rng.randint(0, 5000, size=50000).astype(np.uint64), np.uint64(n - 1))
    a, b = R.rank2a(hr, k, l), oracle.rank2a(ho, k, l)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    q = H.synth_reads(g, 1500, 100, 0.01, 103)
    seq, off = H.reads_to_flat(q)
    for sm in (0, 1):
        ra, oa = R.smem(hr, seq, off, sm, 2), oracle.smem(ho, seq, off, sm, 2)
        assert np.array_equal(ra[0], oa[0]) and np.array_equal(ra[1], oa[1])
    seeds = np.arange(1, 6000, 2).astype(np.uint64)
    ra, oa = R.overlap(hr, 50, seeds), oracle.overlap(ho, 50, seeds)
    assert all(np.array_equal(x, y) for x, y in zip(ra[:3], oa[:3]))
    R.destroy(hr)
    oracle.destroy(ho)


@pytest.mark.skipif(H.reference() is None or H.ref_fermi_binary() is None, reason="needs the compiled reference (oracle/_ref)")
def test_fix_phase_driver_reproduces_fermi_correct(tmp_path):
    """oracle/ref_harness_correct.c: refh_ec_fix_from_triples runs the reference's own worker2 / ec_fix from a given set of k-mer
    triples.  With the triples of the reference's own ec_collect its output must be `fermi correct -t1` byte for byte: that pins
    the driver the GPU hand-off test (tests/test_gpu_parity.py) relies on."""
    import subprocess
    import fermi_b200 as fb
    R = H.reference()
    genome = fb.synth_genome(31, 20000)
    reads = fb.synth_reads(32, genome, 8000, 100, 0.01)            # 40x, 1 % substitutions
    fq, fmd = str(tmp_path / "r.fq"), str(tmp_path / "r.fmd")
    H.write_fastq(fq, reads)
    h = R.build_text(fb.fmd_text(reads))
    R.dump(h, fmd)
    tri, _, w = R.ec_collect(h, -1, 3)
    out = str(tmp_path / "fixed.fq")
    assert R.ec_fix_from_triples(h, -1, 3, tri, fq, out) == w
    R.destroy(h)
    ref = subprocess.run([H.ref_fermi_binary(), "correct", "-t", "1", fmd, fq], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    ours = open(out, "rb").read()
    assert len(ref) > 100000 and ours == ref
    assert b"a" in ref or b"c" in ref or b"g" in ref or b"t" in ref   # some base was corrected (lower case, correct.c:247)
