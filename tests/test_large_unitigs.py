"""bgreat_b200/bin/getLargeUnitigs against goldens of the reference's tool, and live against
oracle/_ref/getLargeUnitigs where that binary exists. Host only: no GPU needed."""
import base64
import json
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "bgreat_b200", "bin", "getLargeUnitigs")
REF = os.path.join(ROOT, "oracle", "_ref", "getLargeUnitigs")
CASES = json.load(open(os.path.join(ROOT, "tests", "golden", "large_unitigs_cases.json")))


def run(exe, text, length, d):
    src, dst = os.path.join(d, "in.fa"), os.path.join(d, "out.fa")
    with open(src, "wb") as f:
        f.write(text)
    if os.path.exists(dst):
        os.remove(dst)
    p = subprocess.run([exe, src, dst, str(length)], capture_output=True, timeout=60)
    return open(dst, "rb").read(), p.stdout.decode(), p.returncode


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_matches_reference_golden(case, tmp_path):
    out, stdout, rc = run(OURS, base64.b64decode(case["input"]), case["length"], str(tmp_path))
    assert rc == 0
    assert out == base64.b64decode(case["output"])
    assert stdout == case["stdout"]


def test_argument_count_behaviour(tmp_path):
    p = subprocess.run([OURS], capture_output=True)
    assert p.returncode == 0 and p.stdout == b"Need in, out and length...\n"
    p = subprocess.run([OURS, "a", "b", "3", "extra"], capture_output=True, cwd=str(tmp_path))
    assert p.returncode == 0 and p.stdout == b"" and not os.path.exists(tmp_path / "b")


def test_missing_input_writes_an_empty_record(tmp_path):
    """the reference spins forever on an input it cannot open (`while(!in.eof())` on a failed stream, like bgreat
    itself on a missing read file); here the output is the empty record"""
    p = subprocess.run([OURS, str(tmp_path / "nope.fa"), str(tmp_path / "out.fa"), "5"], capture_output=True, timeout=20)
    assert p.returncode == 0
    assert open(tmp_path / "out.fa", "rb").read() == b">N\n\n"


@pytest.mark.skipif(not os.access(REF, os.X_OK), reason="reference tool not built here")
@pytest.mark.parametrize("seed", range(20))
def test_live_fuzz_against_the_reference_tool(seed, tmp_path):
    rng = random.Random(500 + seed)
    lines = []
    for _ in range(rng.randrange(0, 200)):
        n = rng.choice([0, 1, 3, 29, 30, 31, 64, 300, 5000, 2 << 20 if rng.random() < 0.05 else 10])
        lines.append(bytes(rng.choice(b"ACGTacgtn;>_1 ") for _ in range(n)))
    text = b"\n".join(lines) + rng.choice([b"", b"\n", b"\n\n"])
    length = rng.choice([0, 1, 30, 31, 100, -5])
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir()
    d2.mkdir()
    assert run(OURS, text, length, str(d1))[:2] == run(REF, text, length, str(d2))[:2]
