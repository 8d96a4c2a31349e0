"""Parity on the benchmark's own workload (tools/synth: the generator bench.py uses for BASELINE.json configs[1]):
the CUDA path against the oracle per read, and the drop-in command line against the reference BINARY
(oracle/_ref/bgreat, built from the unmodified sources; it travels to the GPU box) on the same files."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools import synth  # noqa: E402

pytestmark = pytest.mark.gpu

BGREAT = os.path.join(ROOT, "bgreat_b200", "bin", "bgreat")


@pytest.fixture(scope="module")
def workload():
    g = synth.Graph(5_000_000, 31, seed=1, repeat_frac=0.005)
    n = 200_000
    seqs, off = g.make_reads(n, 150, 0.01)
    return g, seqs, off, n


def test_bench_workload_matches_oracle(workload):
    import bgreat_b200 as bg
    g, seqs, off, n = workload
    us, uo = g.unitig_arrays()
    o = O.Oracle(31, 2, 2)
    o.index_unitigs(g.unitig_list())
    b = seqs.tobytes()
    reads = [b[int(off[i]):int(off[i + 1])] for i in range(n)]
    for mode, m in ((O.GREEDY, 60_000), (O.EXHAUSTIVE, 30_000)):
        al = bg.Aligner(31, 2, 2)
        al.index_unitigs_packed(us, uo)
        got = al.align_packed(seqs, off[:m + 1].copy(), mode).as_tuples()
        o.reset_counters()
        want = [o.align(r, mode) for r in reads[:m]]
        assert got == want
        c, oc = al.counters(), o.counters()
        for name in ("reads", "aligned", "not_aligned", "no_overlap"):
            assert c[name] == oc[name], name
        al.close()
    # correction mode on a slice
    al = bg.Aligner(31, 2, 2)
    al.index_unitigs_packed(us, uo)
    _, corrected = al.correct_batch(reads[:5000])
    assert corrected == [o.correct(r) for r in reads[:5000]]
    al.close()


def test_whole_batch_is_deterministic_and_partitions(workload):
    """size-independent properties on the full sample: two runs agree bit for bit, classes partition the
    reads, every mapped path starts with a non-negative offset and names existing unitigs"""
    import bgreat_b200 as bg
    g, seqs, off, n = workload
    us, uo = g.unitig_arrays()
    al = bg.Aligner(31, 2, 2)
    al.index_unitigs_packed(us, uo)
    r1 = al.align_packed(seqs, off, 0)
    c1 = al.counters()
    r2 = al.align_packed(seqs, off, 0)
    assert np.array_equal(r1.flags, r2.flags) and np.array_equal(r1.path_len, r2.path_len)
    assert np.array_equal(r1.pool, r2.pool) and np.array_equal(r1.path_off, r2.path_off)
    assert c1["reads"] == n == c1["aligned"] + c1["not_aligned"] + c1["no_overlap"]
    assert int((r1.flags & 1).sum()) == c1["aligned"]
    mapped = np.nonzero(r1.flags & 1)[0]
    assert (r1.path_len[mapped] >= 2).all()
    first = r1.pool[r1.path_off[mapped].astype(np.int64)]
    assert (first >= 0).all()
    ids = np.abs(r1.pool)
    assert ids.max() <= g.n_unitigs
    al.close()


def _ref_runs():
    if not os.access(O.REF_BGREAT, os.X_OK):
        return False
    try:
        return subprocess.run([O.REF_BGREAT], capture_output=True, timeout=20).returncode == 0
    except Exception:
        return False


@pytest.mark.parametrize("fastq", [False, True], ids=["fasta", "fastq"])
def test_cli_equals_reference_binary_on_bench_files(workload, fastq, tmp_path):
    """The drop-in binary and the unmodified reference binary on the same 120 k-read files: identical
    paths / notAligned records (sorted: the reference's order with -t > 1 is unspecified) and counters."""
    if not _ref_runs():
        pytest.skip("compiled reference (oracle/_ref/bgreat) not runnable on this host")
    g, seqs, off, n = workload
    m = 120_000
    d = str(tmp_path)
    g.write_unitigs(os.path.join(d, "u.fa"))
    rf = "r.fq" if fastq else "r.fa"
    synth.write_reads(os.path.join(d, rf), seqs, off, m, fastq=fastq)
    common = ["-r", rf, "-g", "u.fa", "-k", "31", "-m", "2", "-t", "8"] + (["-q"] if fastq else [])
    ref = subprocess.run([O.REF_BGREAT] + common + ["-f", "ref.paths", "-a", "ref.na"], cwd=d, capture_output=True,
                         text=True, timeout=600)
    assert ref.returncode == 0
    ours = subprocess.run([BGREAT] + common + ["-f", "my.paths", "-a", "my.na"], cwd=d, capture_output=True, text=True,
                          timeout=600)
    assert ours.returncode == 0, ours.stderr

    def recs(fn):
        lines = open(os.path.join(d, fn), "rb").read().split(b"\n")
        return sorted(zip(lines[0::2], lines[1::2]))

    assert recs("my.paths") == recs("ref.paths")
    assert recs("my.na") == recs("ref.na")
    keep = ("Reads :", "No overlap :", "Got overlap :", "Overlap and aligned :", "Overlap but not aligned :")
    assert [l for l in ours.stdout.splitlines() if l.startswith(keep)] == \
           [l for l in ref.stdout.splitlines() if l.startswith(keep)]
    # -c on the same files: corrected reads equal the reference's
    refc = subprocess.run([O.REF_BGREAT] + common + ["-c", "-f", "ref.corr", "-a", "ref.na2"], cwd=d, capture_output=True,
                          text=True, timeout=600)
    oursc = subprocess.run([BGREAT] + common + ["-c", "-f", "my.corr", "-a", "my.na2"], cwd=d, capture_output=True,
                           text=True, timeout=600)
    assert refc.returncode == 0 and oursc.returncode == 0, oursc.stderr
    assert recs("my.corr") == recs("ref.corr")
