"""Device code logic on the CPU: the __host__ __device__ per-read bodies of bgr_kernels.cuh, single-stepped
by tests/emu (test-only), against the golden vectors of the compiled reference and against the oracle.
The real kernels are checked on the GPU in test_gpu_parity.py; this file keeps the logic covered where
there is no GPU."""
import json
import os
import random

import pytest

import emu_lib as E
import synth_small as S
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HARNESS = json.load(open(os.path.join(GOLDEN, "harness_cases.json")))
MODES = {"greedy": 0, "exhaustive": 1, "greedy_path": 2, "exhaustive_path": 3}


@pytest.mark.parametrize("case", HARNESS, ids=[c["name"] for c in HARNESS])
def test_emulated_device_code_matches_reference_golden(case):
    k = case["k"]
    reads = [r for r in case["reads"]]
    for mode_name, mode in MODES.items():
        want = case["results"][mode_name]
        got, counters, _, ovf = E.emu_map(case["unitigs"], k, reads, mode, case["m"], case["effort"], scratch_ints=4096)
        assert not ovf
        for read, w, g in zip(reads, want["reads"], got):
            if len(read) < k - 1:
                continue  # the reference has no defined result here (short reads)
            assert [int(g[0]), int(g[1])] + g[2] == w, (mode_name, read)
        if all(len(r) >= k - 1 for r in reads):
            for name in ("reads", "aligned", "not_aligned", "no_overlap", "overlaps"):
                assert counters[name] == want["counters"][name], (mode_name, name)


@pytest.mark.parametrize("seed", range(3000, 3040))
def test_emulated_device_code_matches_oracle_fuzz(seed):
    c = S.fuzz_case(seed)
    for mode in (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH, O.EXHAUSTIVE_PATH):
        o = O.Oracle(c["k"], c["m"], c["effort"])
        o.index_unitigs(c["unitigs"])
        want = [o.align(r, mode) for r in c["reads"]]
        got, counters, _, ovf = E.emu_map(c["unitigs"], c["k"], c["reads"], mode, c["m"], c["effort"], scratch_ints=4096)
        assert not ovf
        assert got == want
        wc = o.counters()
        for name in ("reads", "aligned", "not_aligned", "no_overlap", "overlaps"):
            assert counters[name] == wc[name]


@pytest.mark.parametrize("k,seed", [(31, 1), (31, 2), (31, 3), (21, 4), (17, 5), (12, 6)])
def test_emulated_large_index_filter_tiers(k, seed, monkeypatch):
    """The filter of large indexes (forced here on a small graph): minimizer-addressed 64-bit blocks for k = 31,
    two-hash words for any other k. The emulator also checks every chunk's filter answer against the filter's
    definition, window by window (return code 1003 / 1004 otherwise)."""
    monkeypatch.setenv("BGR_BLOOM", "global")
    rng = random.Random(seed)
    unitigs, haps = S.make_graph(rng, k, genome_len=30000, mean_unitig=70, n_variants=200)
    reads = S.make_reads(rng, haps, 1500, 150, sub_rate=0.01, n_rate=0.0005)
    reads += S.make_reads(rng, haps, 60, 40, sub_rate=0.02) + S.make_reads(rng, haps, 40, 333, sub_rate=0.01)
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        o = O.Oracle(k, 2, 2)
        o.index_unitigs(unitigs)
        want = [o.align(r, mode) for r in reads]
        for how in (0, 1):  # character and packed upload forms of the tile scan
            got, counters, _, ovf = E.emu_map(unitigs, k, reads, mode, 2, 2, scratch_ints=4096, how=how)
            assert not ovf
            assert got == want
            # the filter does filter: the windows it lets through are the true borders (a few per cent of all
            # windows in this graph) plus a small share of false positives
            assert counters["verified"] < 0.15 * counters["windows"]


def test_scratch_overflow_is_flagged_not_truncated():
    rng = random.Random(5)
    unitigs, haps = S.make_graph(rng, 10, genome_len=600, mean_unitig=2, n_variants=0)
    reads = S.make_reads(rng, haps, 30, 120, sub_rate=0.0, off_graph=0.0)
    got, _, _, ovf = E.emu_map(unitigs, 10, reads, 0, 2, 2, scratch_ints=8)
    assert ovf
    got, _, _, ovf = E.emu_map(unitigs, 10, reads, 0, 2, 2, scratch_ints=4096)
    assert not ovf
    o = O.Oracle(10, 2, 2)
    o.index_unitigs(unitigs)
    assert got == [o.align(r, O.GREEDY) for r in reads]
    assert max(len(p) for _, _, p in got) > 8


# ---- correction mode (-c): correct_read, the device function, against the oracle's recoverPath restatement
# (itself pinned on the reference binary's -c output, test_oracle_golden.py)

def _correct_case(k, m, effort, unitigs, reads):
    o = O.Oracle(k, m, effort)
    o.index_unitigs(unitigs)
    paths, lens, rcs, want = [], [], [], []
    for r in reads:
        of, rc, path = o.align(r, O.GREEDY)
        if len(path) >= 2:
            paths.append(path)
            lens.append(len(r))
            rcs.append(rc)
            want.append(o.correct(r))
    got = E.correct(unitigs, k, paths, lens, rcs)
    assert got == want
    return len(paths)


@pytest.mark.parametrize("seed", range(3000, 3030))
def test_emulated_correction_matches_oracle_fuzz(seed):
    c = S.fuzz_case(seed)
    _correct_case(c["k"], c["m"], c["effort"], c["unitigs"], c["reads"])


@pytest.mark.parametrize("k,rl", [(5, 40), (16, 70), (31, 150), (32, 200)])
def test_emulated_correction_long_reads(k, rl):
    rng = random.Random(k)
    unitigs, haps = S.make_graph(rng, k, genome_len=4000, mean_unitig=2 * k, n_variants=40)
    reads = S.make_reads(rng, haps, 400, rl, sub_rate=0.02, n_rate=0.001)
    assert _correct_case(k, 3, 2, unitigs, reads) > 10


# ---- packed upload (bgr_pack_reads + bgr_slot_upload_packed): the host packer and the scan forms that take
# their bases from the 2-bit stream must give what the character forms give

def _py_pack(reads):
    s = "".join(r if isinstance(r, str) else r.decode("latin-1") for r in reads)
    words = []
    for w in range((len(s) + 31) // 32):
        v = 0
        for j, c in enumerate(s[32 * w:32 * w + 32]):
            v |= ("ACGT".find(c) & 3 if c in "ACGT" else 0) << (62 - 2 * j)
        words.append(v)
    return words


@pytest.mark.parametrize("threads", [1, 2, 5, 16])
def test_host_packer_matches_plain_python(threads):
    rng = random.Random(100 + threads)
    reads = []
    for i in range(3000):
        ln = rng.choice([0, 1, 3, 31, 32, 33, 64, 100, 150, 151, 250])
        r = "".join(rng.choice("ACGT") for _ in range(ln))
        if ln and rng.random() < 0.02:
            p = rng.randrange(ln)
            r = r[:p] + rng.choice("NnXacgt*\x00\xff") + r[p + 1:]
        reads.append(r.encode("latin-1"))
    words, flagged = E.pack_reads(reads, threads)
    want_flagged = {i for i, r in enumerate(reads) if any(c not in b"ACGT" for c in r)}
    assert flagged == want_flagged and len(flagged) > 10
    want = _py_pack(reads)
    assert len(words) == len(want)
    # words that only cover clean reads must match exactly (codes under a flagged read are unspecified)
    off = [0]
    for r in reads:
        off.append(off[-1] + len(r))
    dirty = set()
    for i in flagged:
        dirty.update(range(off[i] // 32, (off[i + 1] + 31) // 32))
    checked = 0
    for w, (g, x) in enumerate(zip(words, want)):
        if w not in dirty:
            assert int(g) == x, w
            checked += 1
    assert checked > len(want) // 2


def test_host_packer_edge_batches():
    for reads in ([], [b""], [b"", b"", b""], [b"A"], [b"N"], [b"ACGT" * 8], [b"ACGT" * 8, b"", b"T"],
                  [b"ACGT" * 2048 + b"N"], [b"C" * 31, b"G" * 33, b"N" * 5, b"", b"T" * 64]):
        words, flagged = E.pack_reads(reads, 4)
        assert flagged == {i for i, r in enumerate(reads) if any(c not in b"ACGT" for c in r)}
        want = _py_pack(reads)
        if not flagged:
            assert [int(x) for x in words] == want


@pytest.mark.parametrize("seed", range(3000, 3030))
@pytest.mark.parametrize("how", [1, 3])
def test_emulated_packed_upload_forms_match_oracle_fuzz(seed, how):
    c = S.fuzz_case(seed)
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        o = O.Oracle(c["k"], c["m"], c["effort"])
        o.index_unitigs(c["unitigs"])
        want = [o.align(r, mode) for r in c["reads"]]
        got, counters, _, ovf = E.emu_map(c["unitigs"], c["k"], c["reads"], mode, c["m"], c["effort"], scratch_ints=4096,
                                          how=how)
        assert not ovf
        assert got == want
        wc = o.counters()
        for name in ("reads", "aligned", "not_aligned", "no_overlap", "overlaps"):
            assert counters[name] == wc[name]


@pytest.mark.parametrize("k,rl", [(16, 70), (31, 150), (32, 200), (31, 1000)])
def test_emulated_packed_upload_long_reads_with_n(k, rl):
    rng = random.Random(k + rl)
    unitigs, haps = S.make_graph(rng, k, genome_len=6000, mean_unitig=2 * k, n_variants=40)
    reads = S.make_reads(rng, haps, 300, rl, sub_rate=1.0 / rl, n_rate=0.3 / rl)
    ref, cref, _, _ = E.emu_map(unitigs, k, reads, 0, 2, 2, scratch_ints=4096, how=0)
    for how in (1, 3):
        got, cnt, _, _ = E.emu_map(unitigs, k, reads, 0, 2, 2, scratch_ints=4096, how=how)
        cnt["verified"] = cref["verified"]  # instrumentation of the filter: the general scan form has none
        assert got == ref and cnt == cref
    assert sum(1 for _, _, p in ref if p) > 20


@pytest.mark.parametrize("seed", range(7000, 7060))
def test_emulated_long_k_family_fuzz(seed):
    """k around the compiled-in k-1 = 30 (31, and its neighbours 30 / 32 which take the run-time form), reads of
    k-1 .. 300 bases with and without N, every upload / scan form, against the oracle"""
    rng = random.Random(seed)
    k = rng.choice([31, 31, 31, 32, 30, 21, 17])
    unitigs, haps = S.make_graph(rng, k, genome_len=rng.choice([3000, 8000]), mean_unitig=rng.choice([2 * k, 70, 150]),
                                 n_variants=rng.choice([0, 20, 60]))
    reads = S.make_reads(rng, haps, rng.choice([50, 200]), rng.choice([k - 1, k, 60, 150, 151, 300]),
                         sub_rate=rng.choice([0, 0.01, 0.03]), n_rate=rng.choice([0, 0.001, 0.01]))
    m, e = rng.choice([0, 1, 2, 5]), rng.choice([0, 1, 2, 3])
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        o = O.Oracle(k, m, e)
        o.index_unitigs(unitigs)
        want = [o.align(r, mode) for r in reads]
        for how in (0, 1, 3):
            got, _, _, ovf = E.emu_map(unitigs, k, reads, mode, m, e, scratch_ints=4096, how=how)
            assert not ovf and got == want, (k, mode, how)
