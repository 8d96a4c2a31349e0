"""The packed upload (bgr_pack_reads on the host + bgr_slot_upload_packed: reads travel at 2 bits per base)
must give bit-identical results to the oracle -- and therefore to the character upload -- on every kind of
batch: reads with N (they go through the side channel as characters), ragged and very long reads (tiles that
do not fit, the general scan kernel reading the 2-bit stream), empty reads, empty batches."""
import json
import os
import random

import pytest

import synth_small as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HARNESS = json.load(open(os.path.join(GOLDEN, "harness_cases.json")))
MODES = {"greedy": 0, "exhaustive": 1, "greedy_path": 2, "exhaustive_path": 3}
COUNTERS = ("reads", "aligned", "not_aligned", "no_overlap", "overlaps")


@pytest.fixture(scope="module")
def bg():
    import bgreat_b200
    return bgreat_b200


def _check(bg, k, m, e, unitigs, reads, modes, threads=3):
    for mode in modes:
        o = O.Oracle(k, m, e)
        o.index_unitigs(unitigs)
        want = [o.align(r, mode) for r in reads]
        wc = o.counters()
        al = bg.Aligner(k, m, e)
        al.index_unitigs(unitigs)
        got = al.align_batch_packed(reads, mode, threads=threads).as_tuples()
        bad = [i for i in range(len(reads)) if got[i] != want[i]]
        assert not bad, (mode, len(bad), bad[:5], [(reads[i], got[i], want[i]) for i in bad[:2]])
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name], (mode, name)
        al.close()


@pytest.mark.parametrize("case", HARNESS, ids=[c["name"] for c in HARNESS])
def test_packed_upload_golden_vectors_of_the_reference(bg, case):
    k = case["k"]
    for mode_name, mode in MODES.items():
        al = bg.Aligner(k, case["m"], case["effort"])
        al.index_unitigs(case["unitigs"])
        got = al.align_batch_packed(case["reads"], mode).as_tuples()
        want = case["results"][mode_name]
        for read, w, g in zip(case["reads"], want["reads"], got):
            if len(read) < k - 1:
                continue  # no defined reference result (short reads)
            assert [int(g[0]), int(g[1])] + g[2] == w, (mode_name, read)
        al.close()


@pytest.mark.parametrize("seed", range(2000, 2040))
def test_packed_upload_fuzz_against_oracle(bg, seed):
    c = S.fuzz_case(seed)
    _check(bg, c["k"], c["m"], c["effort"], c["unitigs"], c["reads"], (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH, O.EXHAUSTIVE_PATH))


def test_packed_upload_many_tiles_mixed_lengths_and_n(bg):
    rng = random.Random(22)
    unitigs, haps = S.make_graph(rng, 31, genome_len=30000, mean_unitig=70, n_variants=200)
    reads = []
    for ln in (150, 151, 100, 75, 250, 33, 30):
        reads += S.make_reads(rng, haps, 3000, ln, sub_rate=0.01, n_rate=0.0004)
    rng.shuffle(reads)
    reads += S.make_reads(rng, haps, 40000, 150, sub_rate=0.01, n_rate=0.0)
    reads += ["", "A", "N", "ACGT"]
    _check(bg, 31, 2, 2, unitigs, reads, (O.GREEDY,), threads=8)
    _check(bg, 31, 2, 2, unitigs, reads[:8000], (O.EXHAUSTIVE,), threads=2)


def test_packed_upload_ragged_and_long_reads(bg):
    """reads from k-1 up to 3000 bases in one batch: tiles overflow their buffer, the general kernel takes
    those reads out of the 2-bit stream; a batch of long reads only (no tile kernel at all)"""
    rng = random.Random(12)
    unitigs, haps = S.make_graph(rng, 21, genome_len=6000, mean_unitig=40, n_variants=30)
    reads = []
    for ln in [20, 21, 22, 31, 32, 33, 63, 64, 65, 127, 128, 129, 575, 576, 577, 1000, 3000]:
        reads += S.make_reads(rng, haps, 6, ln, sub_rate=0.002, n_rate=0.001, off_graph=0.0)
    rng.shuffle(reads)
    _check(bg, 21, 3, 2, unitigs, reads, (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH))
    # one huge read among short ones: its tile cannot fit
    mixed = S.make_reads(rng, haps, 600, 60, sub_rate=0.002, n_rate=0.0, off_graph=0.0)
    mixed[300:300] = S.make_reads(rng, haps, 2, 5500, sub_rate=0.0005, n_rate=0.0, off_graph=0.0)
    mixed[100:100] = S.make_reads(rng, haps, 1, 5000, sub_rate=0.0005, n_rate=0.0004, off_graph=0.0)
    _check(bg, 21, 3, 2, unitigs, mixed, (O.GREEDY,))
    long_only = S.make_reads(rng, haps, 40, 5900, sub_rate=0.0005, n_rate=0.0001, off_graph=0.0)
    _check(bg, 21, 3, 2, unitigs, long_only, (O.GREEDY, O.EXHAUSTIVE))


def test_packed_upload_empty_and_all_flagged_batches(bg):
    rng = random.Random(13)
    unitigs, haps = S.make_graph(rng, 15, genome_len=3000, mean_unitig=40, n_variants=10)
    _check(bg, 15, 2, 2, unitigs, [], (O.GREEDY,))
    _check(bg, 15, 2, 2, unitigs, ["", "", ""], (O.GREEDY,))
    reads = S.make_reads(rng, haps, 200, 80, sub_rate=0.01, n_rate=0.0)
    flagged = [r[:7] + "N" + r[8:] for r in reads]
    _check(bg, 15, 2, 2, unitigs, flagged, (O.GREEDY, O.EXHAUSTIVE))


def test_packed_and_character_uploads_agree_on_the_same_slot(bg):
    """alternating the two upload forms on one context: no state of one batch leaks into the next"""
    rng = random.Random(14)
    unitigs, haps = S.make_graph(rng, 31, genome_len=20000, mean_unitig=70, n_variants=130)
    al = bg.Aligner(31, 2, 2)
    al.index_unitigs(unitigs)
    for it in range(4):
        reads = S.make_reads(rng, haps, 5000 - 700 * it, 150, sub_rate=0.01, n_rate=0.0005 * (it % 2))
        a = al.align_batch_packed(reads, O.GREEDY, threads=4).as_tuples()
        b = al.align_batch(reads, O.GREEDY).as_tuples()
        assert a == b
    al.close()


@pytest.mark.parametrize("seed", range(7100, 7112))
def test_long_k_family_fuzz_both_upload_forms(bg, seed):
    """k around the compiled-in k-1 = 30 and its neighbours (run-time form of the scan), reads of k-1 .. 300 bases
    with and without N: character and packed upload against the oracle"""
    rng = random.Random(seed)
    k = rng.choice([31, 31, 32, 30, 21, 17])
    unitigs, haps = S.make_graph(rng, k, genome_len=rng.choice([3000, 8000]), mean_unitig=rng.choice([2 * k, 70, 150]),
                                 n_variants=rng.choice([0, 20, 60]))
    reads = S.make_reads(rng, haps, rng.choice([300, 1500]), rng.choice([k - 1, k, 60, 150, 151, 300]),
                         sub_rate=rng.choice([0, 0.01, 0.03]), n_rate=rng.choice([0, 0.001, 0.01]))
    m, e = rng.choice([0, 1, 2, 5]), rng.choice([0, 1, 2, 3])
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        o = O.Oracle(k, m, e)
        o.index_unitigs(unitigs)
        want = [o.align(r, mode) for r in reads]
        al = bg.Aligner(k, m, e)
        al.index_unitigs(unitigs)
        assert al.align_batch(reads, mode).as_tuples() == want, (k, mode, "characters")
        assert al.align_batch_packed(reads, mode, threads=3).as_tuples() == want, (k, mode, "packed")
        al.close()
