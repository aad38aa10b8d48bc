"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run where /root/reference exists (after `make -C oracle`): it imports nothing from fermi_b200 and
calls only oracle/_ref/libfermi_ref.so (the compiled reference + ref_harness.c).  The fixtures pin
the oracle (tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py) on machines where
the reference is absent.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402


def case(name, genome_len, n_reads, L, err, seed, n_query, q_err, special=None, q_len=None):
    R = H.reference()
    assert R is not None, "build oracle/_ref first (make -C oracle)"
    g = H.synth_genome(genome_len, seed)
    reads = H.synth_reads(g, n_reads, L, err, seed + 1)
    if special is not None:
        reads = special(reads)
    text = H.fmd_text(reads)
    h = R.build_text(text)                       # fm_build: SA-IS + rld_enc (build.c:33)
    fmd = os.path.join(HERE, name + ".fmd")
    R.dump(h, fmd)
    info = R.info(h)
    n = int(info["mcnt"][0])
    rng = np.random.RandomState(seed)
    # rank2a: random pairs, the k == -1 case, the last symbol, pairs across block boundaries
    k = rng.randint(0, n, size=4000).astype(np.uint64)
    l = np.minimum(k + rng.randint(0, 2000, size=4000).astype(np.uint64), np.uint64(n - 1))
    k[:16] = np.uint64(0xFFFFFFFFFFFFFFFF)
    l[16:32] = np.uint64(n - 1)
    k[32:48] = l[32:48]
    ok, ol = R.rank2a(h, k, l)
    # extend: intervals reached by real searches (both directions) plus the six single-base intervals
    q = H.synth_reads(g, n_query, q_len or L, q_err, seed + 2)
    seq, off = H.reads_to_flat(q)
    smem0, moff0, _, _, _ = R.smem(h, seq, off, 0, 1)
    smem1, moff1, _, _, _ = R.smem(h, seq, off, 1, 1)
    ik = np.zeros(len(smem0) + 6, H.INTV)
    ik[: len(smem0)] = smem0
    cnt = info["cnt"]
    for c in range(6):
        comp = 5 - c if 1 <= c <= 4 else c
        ik[len(smem0) + c] = (cnt[c], cnt[comp], cnt[c + 1] - cnt[c], 0)
    ik["info"] = 0
    ik = ik[ik["x2"] > 0]
    is_back = (np.arange(len(ik)) & 1).astype(np.uint8)
    ext = R.extend(h, ik, is_back)
    sb, se, ss = R.backward_search(h, seq, off)
    seeds = np.arange(1, min(int(info["mcnt"][1]), 2 * 600), 2).astype(np.uint64)
    ov_rec, ov_nei, ov_off, _ = R.overlap(h, min(50, L // 2), seeds)
    ec_a, ec_a_cnt, ec_a_w = R.ec_collect(h, -1, 3)            # ec_collect over all suffixes (correct.c:35-87)
    ec_b, ec_b_cnt, ec_b_w = R.ec_collect(h, 12, 2)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        mcnt=info["mcnt"], cnt=info["cnt"], n_bytes=info["n_bytes"], n_frames=info["n_frames"], ibits=info["ibits"],
        text=text, k=k, l=l, ok=ok, ol=ol, ik=ik, is_back=is_back, ext=ext,
        q=q, smem0=smem0, moff0=moff0, smem1=smem1, moff1=moff1, sa_beg=sb, sa_end=se, sa_size=ss,
        ec_a=ec_a, ec_a_cnt=np.array(ec_a_cnt), ec_a_w=ec_a_w, ec_b=ec_b, ec_b_cnt=np.array(ec_b_cnt), ec_b_w=ec_b_w,
        ov_min=min(50, L // 2), ov_seeds=seeds, ov_rec=ov_rec, ov_nei=ov_nei, ov_off=ov_off)
    mag = H.reference_unitig(fmd, min(50, L // 2), 1)            # fermi unitig -l.. -t1 (cmd.c:184)
    with open(os.path.join(HERE, name + ".mag"), "w") as fh:
        fh.write(mag)
    # the text `fermi exact` prints for the queries (cmd.c:316-328): the CLI parity fixture
    import subprocess
    fa = os.path.join(HERE, name + ".query.fa")
    tab = np.array(list("$ACGTN"))
    with open(fa, "w") as fh:
        for i, r in enumerate(q):
            fh.write(">q%d\n%s\n" % (i, "".join(tab[r])))
    txt = subprocess.run([H.ref_fermi_binary(), "exact", fmd, fa], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    with open(os.path.join(HERE, name + ".exact.txt"), "wb") as fh:
        fh.write(txt)
    R.destroy(h)
    print(name, "unitigs", len(H.parse_mag(mag)), "symbols", n, "fmd bytes", os.path.getsize(fmd), "smem", len(smem0), len(smem1))


def with_dups_and_palindromes(reads):
    """duplicate reads, a read equal to another's reverse complement, and even-length rc-palindromes."""
    reads = reads.copy()
    reads[5] = reads[3]
    reads[9] = H.revcomp(reads[7])
    L = reads.shape[1]
    half = reads[11, : L // 2].copy()
    reads[11] = np.concatenate([half, H.revcomp(half)])
    reads[13] = reads[11]
    return reads


if __name__ == "__main__":
    # reads vs their own index (10x, error free): BASELINE config 1 in miniature
    case("reads10x", 8000, 800, 100, 0.0, 7, 300, 0.01)
    # noisy reads, duplicates / rc-duplicates / palindromes
    case("noisy", 3000, 600, 60, 0.02, 21, 300, 0.03, special=with_dups_and_palindromes)
    # genome-like index: few long sequences (10 records of 2 kb), short queries with errors: BASELINE config 2 in miniature
    case("genome", 20000, 10, 2000, 0.0, 33, 400, 0.01, q_len=100)
