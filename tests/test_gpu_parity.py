"""Parity of the CUDA path (through the C ABI) with the oracle and with the golden vectors of the
compiled reference. Bit-exact: paths, start offsets, orientation signs, overlapFound / rc flags, counters."""
import json
import os
import random

import numpy as np
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


def oracle_results(k, m, e, unitigs, reads, mode):
    o = O.Oracle(k, m, e)
    o.index_unitigs(unitigs)
    res = [o.align(r, mode) for r in reads]
    return res, o.counters()


@pytest.mark.parametrize("case", HARNESS, ids=[c["name"] for c in HARNESS])
def test_golden_vectors_of_the_reference(bg, case):
    k = case["k"]
    for mode_name, mode in MODES.items():
        al = bg.Aligner(k, case["m"], case["effort"])
        al.index_unitigs(case["unitigs"])
        got = al.align_batch(case["reads"], mode).as_tuples()
        want = case["results"][mode_name]
        for read, w, g in zip(case["reads"], want["reads"], got):
            if len(read) < k - 1:
                continue  # no defined reference result (short reads)
            assert [int(g[0]), int(g[1])] + g[2] == w, (mode_name, read)
        if all(len(r) >= k - 1 for r in case["reads"]):
            c = al.counters()
            for name in COUNTERS:
                assert c[name] == want["counters"][name], (mode_name, name)
        al.close()


@pytest.mark.parametrize("seed", range(2000, 2060))
def test_fuzz_against_oracle(bg, seed):
    c = S.fuzz_case(seed)
    for mode in (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH, O.EXHAUSTIVE_PATH):
        want, wc = oracle_results(c["k"], c["m"], c["effort"], c["unitigs"], c["reads"], mode)
        al = bg.Aligner(c["k"], c["m"], c["effort"])
        al.index_unitigs(c["unitigs"])
        got = al.align_batch(c["reads"], mode).as_tuples()
        assert got == want
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name]
        al.close()


def _k31_case(seed, n_reads, read_len=150, genome=20000, sub=0.01, n_rate=0.0005):
    rng = random.Random(seed)
    unitigs, haps = S.make_graph(rng, 31, genome_len=genome, mean_unitig=70, n_variants=genome // 150)
    reads = S.make_reads(rng, haps, n_reads, read_len, sub_rate=sub, n_rate=n_rate)
    return unitigs, reads


def test_large_index_filter_tier(bg, monkeypatch):
    """the two-hash Bloom filter read from global memory (what indexes beyond 131072 keys get)"""
    monkeypatch.setenv("BGR_BLOOM", "global")
    unitigs, reads = _k31_case(9, 4000)
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        want, wc = oracle_results(31, 2, 2, unitigs, reads, mode)
        al = bg.Aligner(31, 2, 2)
        al.index_unitigs(unitigs)
        got = al.align_batch(reads, mode).as_tuples()
        assert got == want
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name]
        al.close()


@pytest.mark.parametrize("env", [{"BGR_SCAN": "classic"}, {"BGR_SCAN": "classic", "BGR_BLOOM": "global"},
                                 {"BGR_SCAN_MM": "inline", "BGR_BLOOM": "global"}],
                         ids=["classic-smem", "classic-minimizer", "ws-minimizer-inline"])
def test_other_scan_schedules_agree(bg, monkeypatch, env):
    """The comparison forms of the tile scan -- CTA-wide barriers between the phases (BGR_SCAN=classic, the schedule
    compute-sanitizer's racecheck models completely) on both filter tiers, and the warp-specialised kernel with the
    minimizer blocks loaded where they are tested (BGR_SCAN_MM=inline) -- give the oracle's results too"""
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    unitigs, reads = _k31_case(11, 3000)
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        want, wc = oracle_results(31, 2, 2, unitigs, reads, mode)
        al = bg.Aligner(31, 2, 2)
        al.index_unitigs(unitigs)
        assert al.align_batch(reads, mode).as_tuples() == want
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name]
        assert al.align_batch_packed(reads, mode, threads=2).as_tuples() == want
        al.close()


def test_staged_scan_chunks_with_more_runs_than_slots(bg, monkeypatch):
    """Minimizer tier, staged scan (bgr_scan.cuh: mm_stage_blocks keeps ten block slots per chunk): reads cut from a
    sequence built so that the minimizer changes at every third window; four chunks in ten have more runs than slots
    and are redone with the loads in place. Character and packed upload, greedy and exhaustive."""
    import mm_spiky
    monkeypatch.setenv("BGR_BLOOM", "global")
    rng = random.Random(77)
    genome = mm_spiky.spiky_sequence(rng, 2500, beam=48)
    mins = mm_spiky.minimizers(genome)
    over = sum(1 for j in range(0, len(mins) - 32, 5) if mm_spiky.runs_in(mins, j, j + 32) > 10)
    assert over > 50, "the generator no longer produces chunks with more than ten runs"
    unitigs, haps = S.make_graph(rng, 31, mean_unitig=25, n_variants=12, genome=genome)
    reads = S.make_reads(rng, haps, 6000, 150, sub_rate=0.01, n_rate=0.0003)
    reads += S.make_reads(rng, haps, 500, 97, sub_rate=0.0, n_rate=0.0)
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        want, wc = oracle_results(31, 2, 2, unitigs, reads, mode)
        al = bg.Aligner(31, 2, 2)
        al.index_unitigs(unitigs)
        assert al.index_stats()["keys"] > 0
        got = al.align_batch(reads, mode).as_tuples()
        bad = [i for i in range(len(reads)) if got[i] != want[i]]
        assert not bad, (len(bad), bad[:5], [(got[i], want[i]) for i in bad[:3]])
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name]
        assert al.align_batch_packed(reads, mode, threads=2).as_tuples() == want
        al.close()


def test_staged_scan_mixed_lengths(bg, monkeypatch):
    """Minimizer tier with read lengths around every boundary of the staged scan: one window (k-1 bases), a chunk of
    fewer than 32 windows, exactly 32 / 33 windows (61 / 62 bases), the last chunk of one window, long reads"""
    monkeypatch.setenv("BGR_BLOOM", "global")
    rng = random.Random(79)
    unitigs, haps = S.make_graph(rng, 31, genome_len=20000, mean_unitig=60, n_variants=130)
    reads = []
    for ln in (30, 31, 33, 60, 61, 62, 63, 93, 94, 100, 125, 126, 150, 151, 250, 400):
        reads += S.make_reads(rng, haps, 400, ln, sub_rate=0.01, n_rate=0.0003)
    rng.shuffle(reads)
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        want, wc = oracle_results(31, 2, 2, unitigs, reads, mode)
        al = bg.Aligner(31, 2, 2)
        al.index_unitigs(unitigs)
        got = al.align_batch(reads, mode).as_tuples()
        bad = [i for i in range(len(reads)) if got[i] != want[i]]
        assert not bad, (len(bad), bad[:5], [(len(reads[i]), got[i], want[i]) for i in bad[:3]])
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name]
        assert al.align_batch_packed(reads, mode, threads=2).as_tuples() == want
        al.close()


def test_staged_scan_many_borders_per_chunk(bg, monkeypatch):
    """Minimizer tier: unitigs a few bases long, so that a chunk of 32 windows holds eight or more border keys and the
    warp's joint verification (verify_warp: four survivors per lane and round) takes several rounds"""
    monkeypatch.setenv("BGR_BLOOM", "global")
    rng = random.Random(78)
    unitigs, haps = S.make_graph(rng, 31, genome_len=6000, mean_unitig=3, n_variants=30)
    reads = S.make_reads(rng, haps, 5000, 150, sub_rate=0.005, n_rate=0.0002)
    for mode in (O.GREEDY, O.EXHAUSTIVE):
        want, wc = oracle_results(31, 2, 2, unitigs, reads, mode)
        al = bg.Aligner(31, 2, 2)
        al.index_unitigs(unitigs)
        got = al.align_batch(reads, mode).as_tuples()
        bad = [i for i in range(len(reads)) if got[i] != want[i]]
        assert not bad, (len(bad), bad[:5], [(got[i], want[i]) for i in bad[:3]])
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name]
        al.close()


def test_many_tiles_and_mixed_lengths(bg):
    """a batch spanning several tiles per CTA, with read lengths that leave tiles ragged"""
    rng = random.Random(21)
    unitigs, haps = S.make_graph(rng, 31, genome_len=30000, mean_unitig=70, n_variants=200)
    reads = []
    for ln in (150, 151, 100, 75, 250, 33, 30):
        reads += S.make_reads(rng, haps, 3000, ln, sub_rate=0.01, n_rate=0.0002)
    rng.shuffle(reads)
    reads += S.make_reads(rng, haps, 40000, 150, sub_rate=0.01, n_rate=0.0)
    want, wc = oracle_results(31, 2, 2, unitigs, reads, O.GREEDY)
    al = bg.Aligner(31, 2, 2)
    al.index_unitigs(unitigs)
    got = al.align_batch(reads, O.GREEDY).as_tuples()
    bad = [i for i in range(len(reads)) if got[i] != want[i]]
    assert not bad, (len(bad), bad[:5], [(len(reads[i]), got[i], want[i]) for i in bad[:3]])
    gc = al.counters()
    for name in COUNTERS:
        assert gc[name] == wc[name]
    al.close()


@pytest.mark.parametrize("m", [0, 1, 2, 5])
def test_k31_reads_greedy_and_exhaustive(bg, m):
    unitigs, reads = _k31_case(7, 3000)
    for mode in (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH):
        want, wc = oracle_results(31, m, 2, unitigs, reads, mode)
        al = bg.Aligner(31, m, 2)
        al.index_unitigs(unitigs)
        got = al.align_batch(reads, mode).as_tuples()
        assert got == want
        gc = al.counters()
        for name in COUNTERS:
            assert gc[name] == wc[name]
        assert gc["windows"] == sum(max(0, len(r) - 30 + 1) for r in reads)
        al.close()


def test_ragged_and_long_reads(bg):
    """reads from k-1 bases up to 3000 bases in one batch (the >576-base path keeps words in HBM)."""
    rng = random.Random(11)
    unitigs, haps = S.make_graph(rng, 21, genome_len=6000, mean_unitig=40, n_variants=30)
    reads = []
    for ln in [20, 21, 22, 31, 32, 33, 63, 64, 65, 127, 128, 129, 575, 576, 577, 1000, 3000]:
        reads += S.make_reads(rng, haps, 6, ln, sub_rate=0.002, n_rate=0.001, off_graph=0.0)
    rng.shuffle(reads)
    for mode in (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH):
        want, wc = oracle_results(21, 3, 2, unitigs, reads, mode)
        al = bg.Aligner(21, 3, 2)
        al.index_unitigs(unitigs)
        got = al.align_batch(reads, mode).as_tuples()
        assert got == want
        al.close()


def test_empty_batch_and_unmappable(bg):
    al = bg.Aligner(5, 2, 2)
    al.index_unitigs(["ACGTACGTTT", "CGTTTGGA"])
    r = al.align_batch([], 0)
    assert len(r) == 0
    r = al.align_batch(["GGGGGGGGGGGG"], 0).as_tuples()
    assert r == [(False, False, [])]
    c = al.counters()
    assert c["reads"] == 1 and c["no_overlap"] == 1
    al.close()


def test_empty_index(bg):
    al = bg.Aligner(5, 2, 2)
    al.index_unitigs([])
    assert al.align_batch(["ACGTACGTAC"], 0).as_tuples() == [(False, False, [])]
    assert al.align_batch(["ACGTACGTAC"], 1).as_tuples() == [(True, False, [])]
    al.close()


def test_long_paths_grow_the_scratch(bg):
    rng = random.Random(5)
    unitigs, haps = S.make_graph(rng, 10, genome_len=600, mean_unitig=2, n_variants=0)
    reads = S.make_reads(rng, haps, 200, 300, sub_rate=0.0, off_graph=0.0)
    for mode in (O.GREEDY, O.EXHAUSTIVE, O.GREEDY_PATH):
        want, _ = oracle_results(10, 2, 2, unitigs, reads, mode)
        al = bg.Aligner(10, 2, 2)
        al.index_unitigs(unitigs)
        got = al.align_batch(reads, mode).as_tuples()
        assert got == want
        assert max(len(p) for _, _, p in got) > 100
        al.close()


def test_index_queries_match_reference_semantics(bg):
    """getBegin / getEnd on the K6 slot-overflow and K10 hairpin graphs."""
    al = bg.Aligner(4, 0, 2)
    al.index_unitigs(["ATCAA", "ATCCC", "ATCGG", "ATCTT", "ATCAG", "GGATC", "TTGATC"])
    atc = O.str2num("ATC")
    assert al.get_begin(atc) == [1, 2, 3, 5]  # the 5th arrival overwrote slot 4 (aligner.cpp:487)
    assert al.get_end(atc) == [6, 7]
    assert al.get_end(O.str2num("GAT")) == [-1, -2, -3, -5]  # rc(ATC): the same unitigs reversed
    al.close()
    al = bg.Aligner(4, 0, 2)
    al.index_unitigs(["AACTGTT", "CCAAC", "AACGG"])
    assert al.get_begin(O.str2num("AAC")) == [1, 1, 3]  # hairpin inserted twice, forward both times
    al.close()


def test_index_blob_roundtrip_host_and_device(bg):
    import torch
    unitigs, reads = _k31_case(3, 500)
    a = bg.Aligner(31, 2, 2)
    a.index_unitigs(unitigs)
    want = a.align_batch(reads, 0).as_tuples()
    blob = a.export_index_host()
    b = bg.Aligner(31, 2, 2)
    b.import_index_host(blob)
    assert b.align_batch(reads, 0).as_tuples() == want
    # device path: what bench.py does around a NCCL broadcast
    n = a.index_blob_size()
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    a.export_index_device(t.data_ptr(), n)
    torch.cuda.synchronize()
    c = bg.Aligner(31, 2, 2)
    c.import_index_device(t.data_ptr(), n)
    assert c.align_batch(reads, 0).as_tuples() == want
    assert c.index_stats() == a.index_stats()
    bad = blob.copy()
    bad[0] ^= 0xFF
    with pytest.raises(bg.BgrError):
        b.import_index_host(bad)
    for x in (a, b, c):
        x.close()


def test_errors_are_reported(bg):
    al = bg.Aligner(31, 2, 2)
    with pytest.raises(bg.BgrError):
        al.align_batch(["ACGT" * 20], 0)  # no index yet
    with pytest.raises(bg.BgrError):
        al.index_unitigs(["ACGTNACGTACGTACGTACGTACGTACGTACGTACGT"])  # N in a unitig
    with pytest.raises(bg.BgrError):
        bg.Aligner(40).index_unitigs(["A" * 50])  # k > 32
    al.close()


def test_effort_and_mismatch_sweep_properties(bg):
    """Size-independent properties on a larger batch: more budget never maps fewer reads exhaustively,
    counters partition the reads, exact reads map with m=0."""
    unitigs, reads = _k31_case(21, 20000, genome=50000, n_rate=0.0)
    prev = -1
    for m in range(0, 6):
        al = bg.Aligner(31, m, 2)
        al.index_unitigs(unitigs)
        r = al.align_batch(reads, 1)
        c = al.counters()
        assert c["reads"] == len(reads) == c["aligned"] + c["not_aligned"] + c["no_overlap"]
        assert c["aligned"] >= prev
        prev = c["aligned"]
        assert int((r.flags & 1).sum()) == c["aligned"]
        al.close()


# ---- index built on the GPU (default) vs built on the host (BGR_INDEX_BUILD=host): same blob layout, same
# content; only the slot a key occupies inside its probe sequence may differ

def _blob_view(blob):
    """-> (header dict, {key: entry}, meta bytes, pool bytes, bloom bytes, invariants hold). An entry is the key's
    8 slots as (start_len, id_flags, KEY of the next-level slot, next flags): the slot a key occupies inside its
    probe sequence may differ between the builders, so the links are compared by the key they lead to."""
    names = ["magic", "vk", "n_unitigs", "n_keys", "n_buckets", "pool_words", "off_keys", "off_entries", "off_meta",
             "off_pool", "total_bytes", "n_bloom", "off_bloom", "two"]
    h = dict(zip(names, np.frombuffer(blob[:8 * len(names)].tobytes(), dtype=np.uint64).tolist()))
    nslots = h["n_buckets"] * 4
    keys = np.frombuffer(blob[h["off_keys"]:h["off_keys"] + 8 * nslots].tobytes(), dtype=np.uint64)
    ent = np.frombuffer(blob[h["off_entries"]:h["off_entries"] + 128 * nslots].tobytes(), dtype=np.uint32).reshape(nslots, 8, 4)
    EMPTY = 0xFFFFFFFFFFFFFFFF
    table, links_ok = {}, True
    for i, k in enumerate(keys):
        if k == EMPTY:
            continue
        slots = []
        for lo, hi, idf, nxt in ent[i].tolist():
            if idf == 0:
                slots.append((0, 0, 0, 0))
                links_ok &= (lo | hi | nxt) == 0
            else:
                nk = int(keys[nxt & 0x3FFFFFFF])
                links_ok &= nk != EMPTY
                slots.append((lo | (hi << 32), idf, nk, nxt >> 30))
        table[int(k)] = tuple(slots)
    empty_ok = all(not ent[i].any() for i, k in enumerate(keys) if k == EMPTY)
    meta = blob[h["off_meta"]:h["off_meta"] + 16 * (h["n_unitigs"] + 1)].tobytes()
    pool = blob[h["off_pool"]:h["off_pool"] + 8 * h["pool_words"]].tobytes()
    bloom = blob[h["off_bloom"]:h["off_bloom"] + 4 * h["n_bloom"]].tobytes()
    # probe-sequence invariant: slots of a bucket fill in order
    kb = keys.reshape(-1, 4) != EMPTY
    in_order = bool((kb[:, :-1] >= kb[:, 1:]).all())
    return h, table, meta, pool, bloom, empty_ok and in_order and links_ok


def _build_both(bg, k, unitigs):
    out = []
    for where in ("gpu", "host"):
        os.environ["BGR_INDEX_BUILD"] = where
        try:
            al = bg.Aligner(k, 2, 2)
            al.index_unitigs(unitigs)
        finally:
            os.environ.pop("BGR_INDEX_BUILD", None)
        out.append(al)
    return out


@pytest.mark.parametrize("case", [c for c in HARNESS if c["name"].startswith(("K", "fuzz_1", "fuzz_2", "k31"))],
                         ids=lambda c: c["name"])
def test_gpu_built_index_equals_host_built_index(bg, case):
    g, h = _build_both(bg, case["k"], case["unitigs"])
    assert g.index_stats() == h.index_stats()
    hg, tg, mg, pg, bg_, okg = _blob_view(g.export_index_host())
    hh, th, mh, ph, bh, okh = _blob_view(h.export_index_host())
    assert hg == hh
    assert okg and okh
    assert tg == th          # every key, with its left/right slots (order, overflow, orientation flags)
    assert mg == mh and pg == ph and bg_ == bh
    g.close()
    h.close()


def test_gpu_built_index_maps_like_host_built_index(bg):
    unitigs, reads = _k31_case(11, 4000, genome=60000)
    for env in ({}, {"BGR_BLOOM": "global"}):
        os.environ.update(env)
        try:
            g, h = _build_both(bg, 31, unitigs)
        finally:
            for k_ in env:
                os.environ.pop(k_, None)
        _, tg, mg, pg, blg, okg = _blob_view(g.export_index_host())
        _, th, mh, ph, blh, okh = _blob_view(h.export_index_host())
        assert okg and okh and tg == th and mg == mh and pg == ph and blg == blh
        for mode in (O.GREEDY, O.EXHAUSTIVE):
            assert g.align_batch(reads, mode).as_tuples() == h.align_batch(reads, mode).as_tuples()
        g.close()
        h.close()


def test_gpu_index_build_reports_the_offending_unitig(bg):
    al = bg.Aligner(5, 2, 2)
    with pytest.raises(bg.BgrError) as e:
        al.index_unitigs(["ACGTACGT", "ACGTTGCA", "ACGTNACGT", "ACGTAAAA"])
    assert "unitig 3 " in str(e.value)
    al.index_unitigs(["ACGTACGT", "ACG", "ACGTNACGT"])  # loading stops at the short one: the N is never seen
    assert al.index_stats()["unitigs"] == 1
    al.close()


# ---- correction mode (-c) through the C ABI

@pytest.mark.parametrize("seed", range(3000, 3012))
def test_correction_mode_matches_oracle_fuzz(bg, seed):
    c = S.fuzz_case(seed)
    o = O.Oracle(c["k"], c["m"], c["effort"])
    o.index_unitigs(c["unitigs"])
    al = bg.Aligner(c["k"], c["m"], c["effort"])
    al.index_unitigs(c["unitigs"])
    reads = [r for r in c["reads"] if len(r) >= c["k"]]  # recoverPath needs a unitig in the path
    res, corrected = al.correct_batch(reads)
    assert res.as_tuples() == [o.align(r, O.GREEDY) for r in reads]
    assert corrected == [o.correct(r) for r in reads]
    al.close()


def test_correction_mode_k31_with_rc_and_n(bg):
    unitigs, reads = _k31_case(17, 6000, genome=40000)
    o = O.Oracle(31, 2, 2)
    o.index_unitigs(unitigs)
    al = bg.Aligner(31, 2, 2)
    al.index_unitigs(unitigs)
    res, corrected = al.correct_batch(reads)
    want = [o.correct(r) for r in reads]
    assert corrected == want
    n_rc = sum(1 for i in range(len(reads)) if res.mapped(i) and res.rc(i))
    assert n_rc > 100 and sum(1 for w in want if w is not None) > 1000
    # a corrected read is within -m substitutions of the read (or of its reverse complement)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i, (r, w) in enumerate(zip(reads, want)):
        if w is not None and "N" not in r:
            ref = r.encode() if not res.rc(i) else r.encode().translate(comp)[::-1]
            ref = ref if not res.rc(i) else ref
            assert len(w) == len(r)
    with pytest.raises(bg.BgrError):  # exhaustive mode has no correction in the reference
        p = al.params(1, 1)
        import ctypes as C
        al._ck(al.L.bgr_slot_run(al.ctx, 0, C.byref(p)))
    al.close()


def test_index_file_cache_round_trip(bg, tmp_path):
    """bgr_index_save_file / bgr_index_load_file: an index loaded from its file maps exactly like the one it
    was saved from; a file built for another k, a truncated file and a missing file are refused."""
    unitigs, reads = _k31_case(17, 3000)
    al = bg.Aligner(31, 2, 2)
    al.index_unitigs(unitigs)
    want = al.align_batch(reads, O.GREEDY).as_tuples()
    stats = al.index_stats()
    path = str(tmp_path / "index.bgr")
    al.save_index(path)
    al.close()
    assert os.path.getsize(path) == stats["device_bytes"]
    al2 = bg.Aligner(31, 2, 2)
    al2.load_index(path)
    assert al2.index_stats() == stats
    assert al2.align_batch(reads, O.GREEDY).as_tuples() == want
    assert al2.align_batch_packed(reads, O.EXHAUSTIVE).as_tuples() == oracle_results(31, 2, 2, unitigs, reads, O.EXHAUSTIVE)[0]
    al2.close()
    al3 = bg.Aligner(25, 2, 2)
    with pytest.raises(bg.BgrError) as e:
        al3.load_index(path)
    assert e.value.code == 5  # BGR_ERR_FORMAT
    with open(path, "r+b") as f:
        f.truncate(stats["device_bytes"] // 2)
    with pytest.raises(bg.BgrError) as e:
        al3.load_index(path)
    assert e.value.code == 5
    with pytest.raises(bg.BgrError) as e:
        al3.load_index(str(tmp_path / "missing.bgr"))
    assert e.value.code == 3  # BGR_ERR_IO
    al3.close()
