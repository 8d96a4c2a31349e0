"""The drop-in `bgreat` binary (bgreat_b200/bin/bgreat) against golden runs of the reference binary:
same files, same stdout lines (timing lines excluded), plus the documented -o / -b extensions."""
import json
import os
import subprocess

import pytest

import synth_small as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BGREAT = os.path.join(ROOT, "bgreat_b200", "bin", "bgreat")
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = json.load(open(os.path.join(GOLDEN, "cli_cases.json")))
TIMING = ("Indexing in seconds", "Reads/seconds", "Mapping in seconds")


def run_cli(args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([BGREAT] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=120, env=e)
    lines = [l for l in p.stdout.splitlines() if not l.startswith(TIMING)]
    return p.returncode, lines, p.stderr


def records(text):
    """header/body pairs as a sorted list (the reference's order is unspecified with -t > 1)."""
    lines = text.split("\n")
    return sorted(zip(lines[0::2], lines[1::2]))


@pytest.mark.parametrize("case", CLI, ids=[c["name"] for c in CLI])
def test_cli_matches_reference_binary(case, tmp_path):
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    rc, lines, err = run_cli(case["args"], str(tmp_path))
    assert rc == 0, err
    assert lines == case["stdout"]
    brute = "-b" in case["args"]
    with_o = "-o" in case["args"]
    for fn, want in case["outputs"].items():
        p = tmp_path / fn
        if want is None:
            if with_o and fn == case["args"][case["args"].index("-o") + 1]:
                continue  # documented extension: -o FILE is written here, ignored by the reference
            assert not p.exists(), fn
            continue
        got = open(p, newline="").read()
        if brute:
            continue  # the reference writes nothing in -b mode; checked against the harness below
        if with_o and fn.startswith("notAligned"):
            continue
        assert got == want, fn


@pytest.mark.parametrize("env", [{"BGR_RANGE_BYTES": "40"}, {"BGR_RANGE_BYTES": "1000", "BGR_OUT": "write"},
                                 {"BGR_OUT": "write"}], ids=["tiny_ranges", "ranges_write", "write"])
@pytest.mark.parametrize("name", ["K1_two_files", "K4_fastq_trailing_nl", "K4_fastq_no_trailing_nl", "K5_fasta_edges_m1",
                                  "random_k7_e2"])
def test_cli_many_batches_and_plain_writes(name, env, tmp_path):
    """The same golden runs with the input cut into many batches (several worker threads) and with the
    write() output path instead of the memory-mapped one."""
    case = next(c for c in CLI if c["name"] == name)
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    args = [a for a in case["args"]]
    if "-t" in args:
        args[args.index("-t") + 1] = "3"
    else:
        args += ["-t", "3"]
    rc, lines, err = run_cli(args, str(tmp_path), env)
    assert rc == 0, err
    assert lines == case["stdout"]
    for fn, want in case["outputs"].items():
        if want is not None:
            assert open(tmp_path / fn, newline="").read() == want, fn


def test_cli_output_to_a_pipe(tmp_path):
    """-f /dev/stdout style targets cannot be memory-mapped: plain writes."""
    case = next(c for c in CLI if c["name"] == "K5_fasta_edges_m1")
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    p = subprocess.run("%s %s -f /dev/stderr > /dev/null" % (BGREAT, " ".join(case["args"])), shell=True, cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0
    assert p.stderr == case["outputs"]["paths"]


CORRECT = json.load(open(os.path.join(GOLDEN, "cli_correct_cases.json")))


@pytest.mark.parametrize("env", [{}, {"BGR_RANGE_BYTES": "300", "BGR_OUT": "write"}], ids=["default", "small_batches"])
@pytest.mark.parametrize("case", CORRECT, ids=[c["name"] for c in CORRECT])
def test_cli_correction_mode_matches_reference_binary(case, env, tmp_path):
    """bgreat -c: corrected reads instead of paths, same files as the reference binary writes."""
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    rc, lines, err = run_cli(case["args"] + ["-t", "3"], str(tmp_path), env)
    assert rc == 0, err
    assert lines == case["stdout"]
    for fn, want in case["outputs"].items():
        assert open(tmp_path / fn, newline="").read() == want, fn


def test_cli_o_splits_no_overlap_reads(tmp_path):
    """D1: with -o the union of the two files equals the reference's -a file."""
    case = next(c for c in CLI if c["name"] == "K5_fasta_edges_m1")
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    rc, lines, err = run_cli(case["args"] + ["-o", "noOverlap.fa"], str(tmp_path))
    assert rc == 0, err
    na = open(tmp_path / "notAligned.fa").read()
    nov = open(tmp_path / "noOverlap.fa").read()
    assert records(na + nov) == records(case["outputs"]["notAligned.fa"])
    n_nov = int([l for l in case["stdout"] if l.startswith("No overlap")][0].split(" : ")[1].split()[0])
    assert len(records(nov)) == n_nov
    assert open(tmp_path / "paths").read() == case["outputs"]["paths"]


def test_cli_brute_writes_what_the_harness_computes(tmp_path):
    """D3: -b output = per-read results of the reference's alignReadExhaustive (golden harness)."""
    harness = json.load(open(os.path.join(GOLDEN, "harness_cases.json")))
    case = next(c for c in harness if c["name"] == "fuzz_3")
    reads = [r for r in case["reads"] if len(r) > case["k"]]  # the FASTA reader drops len <= k
    S.write_unitigs(str(tmp_path / "u.fa"), case["unitigs"])
    S.write_fasta(str(tmp_path / "r.fa"), reads)
    rc, lines, err = run_cli(["-r", "r.fa", "-g", "u.fa", "-k", case["k"], "-m", case["m"], "-e", case["effort"], "-b"],
                             str(tmp_path))
    assert rc == 0, err
    want = {r: w for r, w in zip(case["reads"], case["results"]["exhaustive"]["reads"])}
    got_paths = dict(records(open(tmp_path / "paths").read()))
    got_na = dict(records(open(tmp_path / "notAligned.fa").read()))
    for i, r in enumerate(reads):
        w = want[r][2:]
        h = ">r%d" % i
        if w:
            assert got_paths[h] == "".join("%d." % x for x in w)
        else:
            assert got_na[h] == r


def test_cli_missing_read_file_is_an_error(tmp_path):
    S.write_unitigs(str(tmp_path / "u.fa"), ["AAATC", "ATCGGA"])
    rc, lines, err = run_cli(["-r", "nope.fa", "-g", "u.fa", "-k", 4], str(tmp_path))
    assert rc != 0 and "cannot open" in err


def test_cli_threads_and_fastq_mid_size(tmp_path):
    """20 k reads, FASTA and FASTQ twins, -t 4: files equal the oracle's whole-run output."""
    import random
    rng = random.Random(77)
    unitigs, haps = S.make_graph(rng, 21, genome_len=20000, mean_unitig=50, n_variants=120)
    reads = S.make_reads(rng, haps, 20000, 100, sub_rate=0.01, n_rate=0.0005)
    S.write_unitigs(str(tmp_path / "u.fa"), unitigs)
    S.write_fasta(str(tmp_path / "r.fa"), reads)
    with open(tmp_path / "r.fq", "w") as f:
        for i, r in enumerate(reads):
            f.write("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
    for fastq in (False, True):
        rf = "r.fq" if fastq else "r.fa"
        args = ["-r", rf, "-g", "u.fa", "-k", 21, "-m", 2, "-t", 4, "-f", "my.paths", "-a", "my.na"] + (["-q"] if fastq else [])
        rc, lines, err = run_cli(args, str(tmp_path))
        assert rc == 0, err
        o = O.Oracle(21, 2, 2)
        o.index_file(str(tmp_path / "u.fa"))
        o.run(O.GREEDY, str(tmp_path / rf), fastq, str(tmp_path / "or.paths"), str(tmp_path / "or.na"))
        assert open(tmp_path / "my.paths").read() == open(tmp_path / "or.paths").read()
        assert open(tmp_path / "my.na").read() == open(tmp_path / "or.na").read()
        c = o.counters()
        assert "Reads : %d" % c["reads"] in lines
        assert any(l.startswith("Overlap and aligned : %d " % c["aligned"]) for l in lines)


def test_cli_index_cache_file(tmp_path):
    """BGR_INDEX_CACHE: the first run builds the index and writes the file, the second run maps from the file
    (the unitig file is gone by then) with identical outputs; a cache built for another k is not used."""
    case = next(c for c in CLI if "-b" not in c["args"] and "-o" not in c["args"] and "-c" not in c["args"]
                and c["outputs"].get("paths"))
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    cache = str(tmp_path / "graph.bgr")
    rc, lines, err = run_cli(case["args"], str(tmp_path), {"BGR_INDEX_CACHE": cache})
    assert rc == 0, err
    assert lines == case["stdout"] and os.path.getsize(cache) > 0
    first = {fn: open(tmp_path / fn, newline="").read() for fn, w in case["outputs"].items() if w is not None}
    unitig_file = case["args"][case["args"].index("-g") + 1] if "-g" in case["args"] else "unitig.fa"
    os.rename(tmp_path / unitig_file, tmp_path / "moved_away.fa")
    rc, lines, err = run_cli(case["args"], str(tmp_path), {"BGR_INDEX_CACHE": cache})
    assert rc == 0, err
    assert lines == case["stdout"]
    for fn, want in first.items():
        assert open(tmp_path / fn, newline="").read() == want == case["outputs"][fn], fn
    # another k: the cache is refused (message on stderr) and, the unitig file being gone, the index is empty
    args = list(case["args"])
    ki = args.index("-k") + 1
    args[ki] = str(int(args[ki]) + 1)
    rc, lines, err = run_cli(args, str(tmp_path), {"BGR_INDEX_CACHE": cache})
    assert "index cache not used" in err


# ---- several devices inside the binary (reads shard across the GPUs of a box, the index is replicated device to
# device, the writer restores input order). BGR_GPUS beyond the number of devices puts several contexts on one
# device, which exercises the same pipeline on a single-GPU box.

def _gpu_count():
    import bgreat_b200
    from bgreat_b200 import _abi as A
    import ctypes as C
    n = C.c_int(0)
    assert A.lib().bgr_device_count(C.byref(n)) == 0
    return n.value


@pytest.mark.parametrize("gpus", ["2", "3", "all"])
@pytest.mark.parametrize("name", ["K1_two_files", "K4_fastq_trailing_nl", "random_k7_e2"])
def test_cli_golden_runs_on_several_contexts(name, gpus, tmp_path):
    if gpus == "all" and _gpu_count() < 2:
        pytest.skip("one GPU: the oversubscribed forms cover the pipeline")
    case = next(c for c in CLI if c["name"] == name)
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    env = {"BGR_RANGE_BYTES": "60", "BGR_GPUS": gpus if gpus != "all" else str(_gpu_count())}
    rc, lines, err = run_cli(case["args"], str(tmp_path), env)
    assert rc == 0, err
    assert lines == case["stdout"]
    for fn, want in case["outputs"].items():
        if want is not None:
            assert open(tmp_path / fn, newline="").read() == want, fn


def test_cli_several_contexts_equal_one_on_a_larger_file(tmp_path):
    """40 k reads cut into ~80 batches: outputs of 1, 2 and 4 contexts are byte-equal (input order is kept)."""
    import random
    rng = random.Random(9)
    unitigs, haps = S.make_graph(rng, 31, genome_len=60000, mean_unitig=80, n_variants=400)
    reads = S.make_reads(rng, haps, 40000, 150, sub_rate=0.01, n_rate=0.0003)
    with open(tmp_path / "u.fa", "w") as f:
        for i, u in enumerate(unitigs):
            f.write(">%d\n%s\n" % (i + 1, u))
    with open(tmp_path / "r.fa", "w") as f:
        for i, r in enumerate(reads):
            f.write(">r%d\n%s\n" % (i, r))
    outs = {}
    for g in ("1", "2", "4"):
        rc, lines, err = run_cli(["-r", "r.fa", "-g", "u.fa", "-k", 31, "-m", 2, "-t", 4, "-f", "p" + g, "-a", "n" + g],
                                 str(tmp_path), {"BGR_GPUS": g, "BGR_RANGE_BYTES": "80000"})
        assert rc == 0, err
        outs[g] = (lines, open(tmp_path / ("p" + g)).read(), open(tmp_path / ("n" + g)).read())
    assert outs["1"] == outs["2"] == outs["4"]
    assert outs["1"][1].count("\n") > 20000
