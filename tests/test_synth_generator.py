"""The benchmark's workload generator (tools/bgr_synth.cpp) emits its unitigs analytically. Checked here on
small genomes: the unitigs hold exactly the k-mers of the two haplotypes, neighbours overlap by k-1 bases, and
the repeat families produce borders with more than two candidates (up to the reference's four-slot overflow)."""
import os
import sys
from collections import Counter

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

_RC = str.maketrans("ACGT", "TGCA")


def rc(s):
    return s[::-1].translate(_RC)


def canon(s):
    r = rc(s)
    return s if s <= r else r


@pytest.mark.parametrize("G,repeat_frac,seed", [(30000, 0.0, 3), (60000, 0.05, 3), (80000, 0.03, 7)])
def test_unitigs_are_the_graph_of_both_haplotypes(G, repeat_frac, seed):
    k = 31
    g = synth.Graph(G, k, seed=seed, repeat_frac=repeat_frac)
    haps = [g.haplotype(0), g.haplotype(1)]
    assert len(haps[0]) == G and sum(a != b for a, b in zip(*haps)) == g.n_snps
    want = {canon(s[i:i + k]) for s in haps for i in range(len(s) - k + 1)}
    us = g.unitig_list()
    got = Counter(canon(u[i:i + k]) for u in us for i in range(len(u) - k + 1))
    assert set(got) == want
    # a k-mer sits in one unitig only, except the k-mer that enters / leaves a repeat copy when two copies
    # happen to have the same flanking base (at most one on either side of every copy)
    assert sum(1 for v in got.values() if v > 1) <= 2 * g.n_repeat_copies
    borders = Counter()
    for u in us:
        borders[canon(u[:k - 1])] += 1
        borders[canon(u[-(k - 1):])] += 1
    if repeat_frac:
        assert g.n_repeat_copies >= 4 and max(borders.values()) >= 4  # degree > 2 exists
    else:
        assert max(borders.values()) <= 3  # bubbles only: one shared unitig + two branches


def test_reads_do_not_depend_on_how_they_are_generated():
    g = synth.Graph(200000, 31, seed=5, repeat_frac=0.01)
    lean = synth.Graph(200000, 31, seed=5, repeat_frac=0.01, unitigs=False)  # what a rank without the index builds
    assert lean.n_unitigs == 0 and lean.n_snps == g.n_snps
    a, _ = g.make_reads(3000, 150, 0.01, threads=1)
    b, _ = lean.make_reads(3000, 150, 0.01, threads=5)
    assert np.array_equal(a[:3000 * 150], b[:3000 * 150])
    c, _ = lean.make_reads(1000, 150, 0.01, first=2000, threads=3)
    assert np.array_equal(a[2000 * 150:3000 * 150], c[:1000 * 150])
    # an error-free read is a substring of a haplotype (or of its reverse complement)
    e, _ = g.make_reads(50, 150, 0.0)
    haps = [g.haplotype(0), g.haplotype(1)]
    for i in range(50):
        s = e[i * 150:(i + 1) * 150].tobytes().decode()
        assert any(s in h or rc(s) in h for h in haps)


def test_spiky_sequence_has_chunks_with_more_runs_than_slots():
    """tests/mm_spiky.py (input of the staged scan's overflow test): its plain-Python minimizers find, in the
    sequence it builds, spans of 32 windows with more than ten minimizer runs; a random sequence has next to none"""
    import random
    import mm_spiky
    rng = random.Random(77)
    g = mm_spiky.spiky_sequence(rng, 600, beam=48)
    m = mm_spiky.minimizers(g)
    assert sum(1 for j in range(0, len(m) - 32, 5) if mm_spiky.runs_in(m, j, j + 32) > 10) > 5
    r = "".join(rng.choice("ACGT") for _ in range(600))
    m = mm_spiky.minimizers(r)
    assert max(mm_spiky.runs_in(m, j, j + 32) for j in range(0, len(m) - 32, 5)) <= 12
