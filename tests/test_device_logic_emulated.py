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
