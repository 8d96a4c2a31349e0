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


def test_unitig_loader_missing_file():
    opened, got = E.load_unitigs("/nonexistent/u.fa", 4)
    assert not opened and got == []
