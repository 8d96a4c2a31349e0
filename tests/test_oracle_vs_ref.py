"""Live differential fuzz: CPU oracle vs the compiled reference (only where oracle/_ref exists,
i.e. in the build container; skipped on the GPU box, which relies on tests/golden/)."""
import os

import pytest

import synth_small as S
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not present")


@pytest.mark.parametrize("seed", range(1000, 1030))
def test_fuzz_against_reference_harness(seed, tmp_path):
    c = S.fuzz_case(seed)
    uf = str(tmp_path / "u.fa")
    S.write_unitigs(uf, c["unitigs"])
    for mode in (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH, O.EXHAUSTIVE_PATH):
        ref, ref_counters = O.ref_harness_align(uf, c["k"], c["m"], c["effort"], mode, c["reads"], str(tmp_path))
        o = O.Oracle(c["k"], c["m"], c["effort"])
        o.index_file(uf)
        got = [o.align(r, mode) for r in c["reads"]]
        assert got == ref
        cnt = o.counters()
        assert all(cnt[k] == v for k, v in ref_counters.items())


def test_whole_run_against_reference_binary(tmp_path):
    import random
    rng = random.Random(99)
    unitigs, haps = S.make_graph(rng, 9, genome_len=2000, mean_unitig=25, n_variants=30)
    reads = S.make_reads(rng, haps, 2000, 60, sub_rate=0.02, n_rate=0.002)
    S.write_unitigs(str(tmp_path / "u.fa"), unitigs)
    S.write_fasta(str(tmp_path / "r.fa"), reads)
    O.ref_bgreat(["-r", "r.fa", "-g", "u.fa", "-k", 9, "-m", 2, "-f", "ref.paths", "-a", "ref.na"], str(tmp_path))
    o = O.Oracle(9, 2, 2)
    o.index_file(str(tmp_path / "u.fa"))
    o.run(O.GREEDY, str(tmp_path / "r.fa"), False, str(tmp_path / "my.paths"), str(tmp_path / "my.na"))
    assert open(tmp_path / "my.paths").read() == open(tmp_path / "ref.paths").read()
    assert open(tmp_path / "my.na").read() == open(tmp_path / "ref.na").read()
