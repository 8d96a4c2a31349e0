"""Parity at the graph sizes of BASELINE.json configs[2] / [3] (100 Mbp: 2 M unitigs, 1.3 M border keys, the
large-index filter tier at its real size, no environment forcing) against the oracle, and -- at configs[4]'s 3 Gbp
(13 GB index, blob offsets beyond 4 GB, 64 MB of filter blocks) where the oracle would need minutes to index --
through properties that do not depend on size: both upload forms agree bit for bit, the classes partition the
reads, every mapped path spells its read within the mismatch budget. The oracle comparison at 1 Gbp / 3 Gbp is
tools/parity_large.py (run on the GPU box, digest under profiles/)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools import synth  # noqa: E402

pytestmark = pytest.mark.gpu

K = 31
COUNTERS = ("reads", "aligned", "not_aligned", "no_overlap", "overlaps")


def _reads(seqs, off, n):
    b = seqs.tobytes()
    return [b[int(off[i]):int(off[i + 1])] for i in range(n)]


@pytest.fixture(scope="module")
def g100():
    import bgreat_b200 as bg
    g = synth.Graph(100_000_000, K, seed=7, repeat_frac=0.002)
    us, uo = g.unitig_arrays()
    o = O.Oracle(K, 2, 2)
    o.index_arrays(us, uo)
    al = bg.Aligner(K, 2, 2)
    al.index_unitigs_packed(us, uo)
    assert al.index_stats()["keys"] > 1_000_000  # far beyond the shared-memory filter tier
    yield g, o, al
    al.close()


@pytest.mark.parametrize("m", [0, 1, 2, 3, 4, 5])
def test_greedy_mismatch_sweep_100mbp(g100, m):
    """configs[2]: -m 0..5, 1 % substitutions, 20 k reads per setting"""
    g, o, al = g100
    n = 20_000
    seqs, off = g.make_reads(n, 150, 0.01, first=1_000_000 * m)
    o.set_params(m, 2)
    o.reset_counters()
    want = [o.align(r, O.GREEDY) for r in _reads(seqs, off, n)]
    al.errors = m
    al.reset_counters()
    got = al.align_packed(seqs, off, O.GREEDY).as_tuples()
    assert got == want
    c, oc = al.counters(), o.counters()
    for name in COUNTERS:
        assert c[name] == oc[name], name
    assert c["aligned"] > (n // 4 if m >= 2 else 1000)  # with -m 0 only error-free reads map


@pytest.mark.parametrize("mode", [O.EXHAUSTIVE, O.EXHAUSTIVE_PATH, O.GREEDY_PATH])
def test_exhaustive_two_percent_100mbp(g100, mode):
    """configs[3]: the exhaustive aligners on reads with 2 % substitutions"""
    g, o, al = g100
    n = 20_000 if mode == O.EXHAUSTIVE else 6_000
    seqs, off = g.make_reads(n, 150, 0.02, first=50_000_000)
    for m in (2, 4):
        o.set_params(m, 2)
        o.reset_counters()
        want = [o.align(r, mode) for r in _reads(seqs, off, n)]
        al.errors = m
        al.reset_counters()
        got = al.align_packed(seqs, off, mode).as_tuples()
        assert got == want
        c, oc = al.counters(), o.counters()
        for name in COUNTERS:
            assert c[name] == oc[name], name


def test_packed_upload_and_correction_100mbp(g100):
    g, o, al = g100
    n = 20_000
    seqs, off = g.make_reads(n, 150, 0.01, n_rate=0.0003, first=70_000_000)
    reads = _reads(seqs, off, n)
    o.set_params(2, 2)
    al.errors = 2
    want = [o.align(r, O.GREEDY) for r in reads]
    assert al.align_batch_packed(reads, O.GREEDY, threads=4).as_tuples() == want
    _, corrected = al.correct_batch(reads[:3000])
    assert corrected == [o.correct(r) for r in reads[:3000]]


def _rc(b):
    return b[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))


def _spell(us, uo, k, path):
    """the sequence a path [offset, +-u1, ...] spells from its offset on (unitigs glued on their k-1 overlaps)"""
    s = b""
    for j, sid in enumerate(path[1:]):
        u = us[int(uo[abs(sid) - 1]):int(uo[abs(sid)])].tobytes()
        if sid < 0:
            u = _rc(u)
        if j:
            assert s[-(k - 1):] == u[:k - 1], "consecutive unitigs of a path do not overlap by k-1"
            u = u[k - 1:]
        s += u
    return s[path[0]:]


def _free_ram_gb():
    try:
        for l in open("/proc/meminfo"):
            if l.startswith("MemAvailable:"):
                return int(l.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def test_human_scale_graph_properties():
    """configs[4]: 3 Gbp graph (60 M unitigs, 40 M keys, 13 GB index). Size-independent properties of the mapping of
    1 M reads; the per-read oracle comparison at this size is tools/parity_large.py."""
    import torch
    import bgreat_b200 as bg
    if _free_ram_gb() < 40 or torch.cuda.mem_get_info()[0] < 40e9:
        pytest.skip("needs 40 GB of host RAM and of HBM")
    g = synth.Graph(3_000_000_000, K, seed=1, repeat_frac=0.001)
    us, uo = g.unitig_arrays()
    al = bg.Aligner(K, 2, 2)
    al.index_unitigs_packed(us, uo)
    st = al.index_stats()
    assert st["device_bytes"] > (1 << 32) and st["keys"] > 30_000_000
    n = 1_000_000
    seqs, off = g.make_reads(n, 150, 0.01)
    r1 = al.align_packed(seqs, off, O.GREEDY)
    c1 = al.counters()
    # a second run agrees bit for bit (results do not depend on the order in which warps finish)
    r2 = al.align_packed(seqs, off, O.GREEDY)
    assert np.array_equal(r1.flags, r2.flags) and np.array_equal(r1.path_len, r2.path_len)
    assert np.array_equal(r1.pool, r2.pool) and np.array_equal(r1.path_off, r2.path_off)
    assert c1["reads"] == n == c1["aligned"] + c1["not_aligned"] + c1["no_overlap"]
    assert int((r1.flags & 1).sum()) == c1["aligned"] > n // 3
    # every mapped path spells its read (or, for rc, its reverse complement) within the mismatch budget
    b = seqs.tobytes()
    mapped = np.nonzero(r1.flags & 1)[0]
    rng = np.random.default_rng(5)
    for i in rng.choice(mapped, 3000, replace=False):
        read = b[int(off[i]):int(off[i + 1])]
        if r1.flags[i] & 4:
            read = _rc(read)
        s = _spell(us, uo, K, r1.path(int(i)))
        assert len(s) >= len(read)
        assert sum(x != y for x, y in zip(read, s)) <= 2
    # the same reads as a 2-bit stream packed by the host (slice): identical records
    b_reads = [b[int(off[i]):int(off[i + 1])] for i in range(100_000)]
    rp = al.align_batch_packed(b_reads, O.GREEDY, threads=8)
    assert np.array_equal(rp.flags, r1.flags[:100_000]) and np.array_equal(rp.path_len, r1.path_len[:100_000])
    assert [rp.path(i) for i in range(0, 100_000, 37)] == [r1.path(i) for i in range(0, 100_000, 37)]
    # exhaustive mode on a slice: the classes partition the reads
    al.reset_counters()
    r3 = al.align_packed(seqs, off[:200_001].copy(), O.EXHAUSTIVE)
    c3 = al.counters()
    assert c3["reads"] == 200_000 == c3["aligned"] + c3["not_aligned"]
    assert int((r3.flags & 1).sum()) == c3["aligned"]
    al.close()
