"""The seam of INTEGRATION.md, compiled and run: oracle/_ref/bgreat_seam is the REFERENCE's own driver (its main(),
option parsing, class Aligner, alignAll, getReads, printPath, output files, summary) with Aligner::indexUnitigs and
Aligner::alignPartGreedy bound to libbgreat_b200.so (integration/aligner_gpu_binding.cpp, built by
integration/Makefile from a build-time copy of the reference sources). Checked against the golden runs of the
unmodified reference binary and, on the benchmark generator's files, against that binary itself."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

SEAM = os.path.join(ROOT, "oracle", "_ref", "bgreat_seam")
CLI = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_cases.json")))
TIMING = ("Indexing in seconds", "Reads/seconds", "Mapping in seconds")
GREEDY_CASES = ["K1", "K1_default_unitig_name", "K1_f_a", "K1_two_files", "K4_fastq_trailing_nl", "K4_fastq_no_trailing_nl",
                "K5_fasta_edges_m0", "K5_fasta_edges_m1", "K5_fasta_edges_m2", "missing_unitigs", "unitig_short_line_stops",
                "random_k7_e1", "random_k7_e2", "random_k7_e4"]


def _need_seam():
    if not os.access(SEAM, os.X_OK):
        pytest.skip("oracle/_ref/bgreat_seam not built (make -C integration needs the reference tree)")


def _records(text):
    lines = text.split("\n")
    return sorted(zip(lines[0::2], lines[1::2]))


@pytest.mark.parametrize("name", GREEDY_CASES)
def test_reference_driver_on_the_gpu_path_reproduces_the_reference(name, tmp_path):
    _need_seam()
    case = next(c for c in CLI if c["name"] == name)
    for fn, content in case["files"].items():
        with open(tmp_path / fn, "w", newline="") as f:
            f.write(content)
    p = subprocess.run([SEAM] + [str(a) for a in case["args"]], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert [l for l in p.stdout.splitlines() if not l.startswith(TIMING)] == case["stdout"]
    for fn, want in case["outputs"].items():
        if want is None:
            continue
        assert _records(open(tmp_path / fn, newline="").read()) == _records(want), fn


@pytest.mark.parametrize("extra", [[], ["-c"]], ids=["paths", "corrected"])
def test_reference_driver_on_the_gpu_path_on_bench_files(extra, tmp_path):
    """120 k reads of the benchmark generator, 8 reference threads feeding the one GPU context: records and summary
    equal those of the unmodified reference binary."""
    _need_seam()
    if not os.access(O.REF_BGREAT, os.X_OK):
        pytest.skip("compiled reference (oracle/_ref/bgreat) missing")
    from tools import synth
    g = synth.Graph(5_000_000, 31, seed=1, repeat_frac=0.005)
    n = 120_000
    seqs, off = g.make_reads(n, 150, 0.01)
    d = str(tmp_path)
    g.write_unitigs(os.path.join(d, "u.fa"))
    synth.write_reads(os.path.join(d, "r.fa"), seqs, off, n)
    common = ["-r", "r.fa", "-g", "u.fa", "-k", "31", "-m", "2", "-t", "8"] + extra
    ref = subprocess.run([O.REF_BGREAT] + common + ["-f", "ref.paths", "-a", "ref.na"], cwd=d, capture_output=True, text=True,
                         timeout=600)
    seam = subprocess.run([SEAM] + common + ["-f", "seam.paths", "-a", "seam.na"], cwd=d, capture_output=True, text=True,
                          timeout=600)
    assert ref.returncode == 0 and seam.returncode == 0, seam.stderr
    for a, b in (("seam.paths", "ref.paths"), ("seam.na", "ref.na")):
        assert _records(open(os.path.join(d, a)).read()) == _records(open(os.path.join(d, b)).read()), a
    keep = ("Reads :", "No overlap :", "Got overlap :", "Overlap and aligned :", "Overlap but not aligned :")
    assert [l for l in seam.stdout.splitlines() if l.startswith(keep)] == [l for l in ref.stdout.splitlines() if l.startswith(keep)]
