"""The CPU oracle (oracle/bgreat_oracle.c) against golden vectors produced by the compiled,
unmodified reference (tests/golden/make_golden.py). This is what pins the oracle."""
import json
import os

import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HARNESS = json.load(open(os.path.join(GOLDEN, "harness_cases.json")))
CLI = json.load(open(os.path.join(GOLDEN, "cli_cases.json")))
PARSE = json.load(open(os.path.join(GOLDEN, "parse_cases.json"))) + json.load(open(os.path.join(GOLDEN, "parse_fuzz_cases.json")))
MODES = {"greedy": O.GREEDY, "exhaustive": O.EXHAUSTIVE, "greedy_path": O.GREEDY_PATH,
         "exhaustive_path": O.EXHAUSTIVE_PATH}


@pytest.mark.parametrize("case", HARNESS, ids=[c["name"] for c in HARNESS])
def test_per_read_results_match_reference(case):
    for mode_name, mode in MODES.items():
        o = O.Oracle(case["k"], case["m"], case["effort"])
        o.index_unitigs(case["unitigs"])
        want = case["results"][mode_name]
        for read, w in zip(case["reads"], want["reads"]):
            of, rc, path = o.align(read, mode)
            assert [int(of), int(rc)] + path == w, (mode_name, read)
        c = o.counters()
        for name, v in want["counters"].items():
            assert c[name] == v, (mode_name, name)


def _opt(args, flag, default):
    return args[args.index(flag) + 1] if flag in args else default


@pytest.mark.parametrize("case", [c for c in CLI if c["args"]], ids=[c["name"] for c in CLI if c["args"]])
def test_whole_run_files_match_reference_binary(case, tmp_path):
    args = case["args"]
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    k = int(_opt(args, "-k", "30"))
    m = int(_opt(args, "-m", "2"))
    e = int(_opt(args, "-e", "2"))
    brute = "-b" in args
    paths = _opt(args, "-f", "paths")
    na = _opt(args, "-a", "notAligned.fa")
    o = O.Oracle(k, m, e)
    o.index_file(str(tmp_path / _opt(args, "-g", "unitig.fa")))
    reads_csv = ",".join(str(tmp_path / f) for f in _opt(args, "-r", "").split(","))
    assert o.run(O.EXHAUSTIVE if brute else O.GREEDY, reads_csv, "-q" in args, str(tmp_path / paths),
                 str(tmp_path / na)) == 0
    c = o.counters()
    # stdout counters (aligner.cpp:589-593)
    want = {l.split(" : ")[0]: int(l.split(" : ")[1].split()[0]) for l in case["stdout"] if " : " in l}
    assert c["reads"] == want["Reads"]
    assert c["no_overlap"] == want["No overlap"]
    assert c["aligned"] == want["Overlap and aligned"]
    assert c["not_aligned"] == want["Overlap but not aligned"]
    if not brute:  # the reference writes nothing at all in -b mode (alignerExhaustive.cpp:285)
        assert open(tmp_path / paths, newline="").read() == case["outputs"][paths]
        assert open(tmp_path / na, newline="").read() == case["outputs"][na]


@pytest.mark.parametrize("case", PARSE, ids=[c["name"] for c in PARSE])
def test_read_parser_matches_getReads(case, tmp_path):
    p = tmp_path / "reads"
    with open(p, "w", newline="") as f:
        f.write(case["content"])
    got = O.parse_reads(str(p), case["k"], bool(case["fastq"]), case["chunk"])
    assert [[h.decode("latin1"), r.decode("latin1")] for h, r in got] == case["records"]


def test_missing_reads_file_is_an_error():
    assert O.parse_reads("/nonexistent/reads.fa", 4) is None


def test_primitives():
    assert O.str2num("ACGT") == 0b00011011
    assert O.str2num("ANGT") == 0b00111011  # non-ACG -> 3
    assert O.rcb(O.str2num("AAC"), 3) == O.str2num("GTT")
    assert O.reverse_complement("ACGTN") == b"AACGT"  # N -> A
    L = O.lib()
    assert L.bo_mismatch_number(b"ACGTACGT", b"ACCTACCA", 8, 5) == 3
    assert L.bo_mismatch_number(b"ACGTACGT", b"TTTTTTTT", 8, 2) == 3  # capped at n+1


# ---- correction mode (-c): the oracle's recoverPath restatement against the reference binary's -c runs
CORRECT = json.load(open(os.path.join(GOLDEN, "cli_correct_cases.json")))


@pytest.mark.parametrize("case", CORRECT, ids=[c["name"] for c in CORRECT])
def test_correction_mode_matches_reference_binary(case, tmp_path):
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    args = case["args"]
    k = int(args[args.index("-k") + 1])
    m = int(args[args.index("-m") + 1])
    o = O.Oracle(k, m, 2)
    o.index_file(str(tmp_path / args[args.index("-g") + 1]))
    assert o.run_correct(str(tmp_path / args[args.index("-r") + 1]), "-q" in args, str(tmp_path / "my.paths"),
                         str(tmp_path / "my.na")) == 0
    assert open(tmp_path / "my.paths", newline="").read() == case["outputs"]["paths"]
    assert open(tmp_path / "my.na", newline="").read() == case["outputs"]["notAligned.fa"]
