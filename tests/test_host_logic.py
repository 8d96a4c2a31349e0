"""Host-side logic of the product (bgreat_b200/csrc/bgr_host.cpp) on the CPU: the read parser must yield
exactly what the reference's getReads yields (golden vectors from the compiled reference), the unitig
loader must follow the reference's record loop."""
import json
import os

import pytest

import emu_lib as E
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARSE = json.load(open(os.path.join(GOLDEN, "parse_cases.json")))


@pytest.mark.parametrize("case", PARSE, ids=[c["name"] for c in PARSE])
def test_product_parser_matches_getReads(case, tmp_path):
    p = tmp_path / "reads"
    with open(p, "w", newline="") as f:
        f.write(case["content"])
    got = E.parse_reads(str(p), case["k"], bool(case["fastq"]), case["chunk"])
    assert [[h.decode("latin1"), r.decode("latin1")] for h, r in got] == case["records"]


def test_product_parser_missing_file():
    assert E.parse_reads("/nonexistent/reads.fa", 4) is None


@pytest.mark.parametrize("content,k,want", [
    (">1\nAAATC\n>2\nATCGGA\n>3\nGGATT\n", 4, ["AAATC", "ATCGGA", "GGATT"]),
    (">1\nAAATC\n>2\nATCGGA\n>3\nGGATT", 4, ["AAATC", "ATCGGA", "GGATT"]),
    (">1\nAAATC\n>2\nATC\n>3\nATCGGA\n", 4, ["AAATC"]),           # stops at the first line shorter than k
    (">1\nAAATC\n\n>2\nATCGGA\n", 4, ["AAATC"]),                    # blank line shifts the pairing: ">2" < k? no: ">2" is 2 chars
    ("", 4, []),
    (">1\nAAATC\n>a_header_without_sequence", 4, ["AAATC", ">a_header_without_sequence"]),  # failed 2nd getline keeps the header
])
def test_unitig_loader_follows_the_reference_loop(content, k, want, tmp_path):
    p = tmp_path / "u.fa"
    with open(p, "w", newline="") as f:
        f.write(content)
    opened, got = E.load_unitigs(str(p), k)
    assert opened
    # the oracle's loader restates the same loop and is pinned by the golden CLI cases
    o = O.Oracle(k)
    assert o.index_file(str(p)) == len(got)
    assert [g.decode() for g in got] == want
    for threads in (1, 2, 5):
        assert E.load_unitigs(str(p), k, threads=threads) == (True, got)


@pytest.mark.parametrize("seed", range(60))
def test_parallel_unitig_loader_equals_the_sequential_loop(seed, tmp_path):
    """random two-line records with short lines, blank lines, missing final newline, lone header at the end"""
    import random
    rng = random.Random(seed)
    k = rng.choice([3, 4, 7])
    lines = []
    for i in range(rng.randrange(0, 60)):
        lines.append(">%d" % i if rng.random() < 0.9 else "")
        ln = rng.randrange(k, 20) if rng.random() < 0.97 else rng.randrange(0, k)
        lines.append("".join(rng.choice("ACGT") for _ in range(ln)))
    if rng.random() < 0.3:
        lines.append(">lone_header_%d" % seed)
    text = "\n".join(lines) + "\n" * rng.choice([0, 1, 1, 2])
    p = tmp_path / "u.fa"
    with open(p, "w", newline="") as f:
        f.write(text)
    want = E.load_unitigs(str(p), k)
    for threads in (1, 2, 3, 8):
        assert E.load_unitigs(str(p), k, threads=threads, min_slice=rng.choice([1, 16, 64])) == want


def test_unitig_loader_missing_file():
    opened, got = E.load_unitigs("/nonexistent/u.fa", 4)
    assert not opened and got == []
    assert E.load_unitigs("/nonexistent/u.fa", 4, threads=3) == (False, [])


# ---- the parallel reader the CLI uses (bgr_host.cpp ParallelReader): same records as getReads driven in
# 10000-read chunks, whatever the number of workers and wherever the text is cut

import parse_gen as G

FUZZ = json.load(open(os.path.join(GOLDEN, "parse_fuzz_cases.json")))
BIG = json.load(open(os.path.join(GOLDEN, "parse_big_cases.json")))
SPLITS = [(1, 1 << 20, 4096), (3, 7, 1), (4, 64, 1), (8, 1 << 20, 1), (2, 33, 5)]


def _write(tmp_path, content):
    p = tmp_path / "reads"
    with open(p, "w", newline="") as f:
        f.write(content)
    return str(p)


@pytest.mark.parametrize("case", [c for c in PARSE if c["chunk"] == 10000 or not c["fastq"]],
                         ids=lambda c: c["name"])
def test_parallel_reader_on_the_getReads_goldens(case, tmp_path):
    p = _write(tmp_path, case["content"])
    for threads, rng, ms in SPLITS:
        got = E.parse_reads(p, case["k"], bool(case["fastq"]), threads=threads, range_bytes=rng, min_slice=ms)
        assert [[h.decode("latin1"), r.decode("latin1")] for h, r in got] == case["records"], (threads, rng, ms)


@pytest.mark.parametrize("case", FUZZ, ids=lambda c: c["name"])
def test_parsers_on_fuzzed_texts(case, tmp_path):
    """golden = the reference's getReads on the same text"""
    p = _write(tmp_path, case["content"])
    want = case["records"]
    got = E.parse_reads(p, case["k"], bool(case["fastq"]))
    assert [[h.decode("latin1"), r.decode("latin1")] for h, r in got] == want
    for threads, rng, ms in SPLITS:
        got = E.parse_reads(p, case["k"], bool(case["fastq"]), threads=threads, range_bytes=rng, min_slice=ms)
        assert [[h.decode("latin1"), r.decode("latin1")] for h, r in got] == want, (threads, rng, ms)


@pytest.mark.parametrize("case", BIG, ids=lambda c: c["name"])
def test_parsers_on_many_chunk_files(case, tmp_path):
    """> 10000 records: getReads chunk boundaries, the FASTQ tail rules; golden = digest of the reference's records"""
    p = _write(tmp_path, G.big_text(case["name"]))
    fastq = bool(case["fastq"])
    for kw in (dict(), dict(threads=1, range_bytes=1 << 30), dict(threads=5, range_bytes=100_000, min_slice=1),
               dict(threads=16, range_bytes=1 << 20), dict(threads=3, range_bytes=1000, min_slice=100)):
        got = E.parse_reads(p, case["k"], fastq, **kw)
        assert len(got) == case["n_records"], kw
        assert [[h.decode("latin1"), r.decode("latin1")] for h, r in got[-2:]] == case["last"], kw
        assert G.digest(got) == case["sha256"], kw


def test_read_validation_every_position():
    """acgtn_only (vector body, overlapping tail, scalar rest): one bad byte anywhere must be seen"""
    import ctypes as C
    import random
    L = E.lib()
    rng = random.Random(5)
    for n in list(range(0, 70)) + [95, 96, 97, 150, 151, 255, 256, 1000]:
        good = bytes(rng.choice(b"ACGTN") for _ in range(n))
        assert L.emu_acgtn_only(good, C.c_uint64(n)) == 1
        for pos in range(n):
            for badc in (b"a", b"X", b"\r", b"\x00", b"\xc1", b"B", b"@"):
                bad = good[:pos] + badc + good[pos + 1:]
                assert L.emu_acgtn_only(bad, C.c_uint64(n)) == 0, (n, pos, badc)
