#!/usr/bin/env python
"""Goldens of the reference's getLargeUnitigs (oracle/_ref/getLargeUnitigs, built from the unmodified source by
oracle/Makefile) -> tests/golden/large_unitigs_cases.json. Run in the build container (needs /root/reference)."""
import base64
import json
import os
import random
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "getLargeUnitigs")


def run(text, length):
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "in.fa"), os.path.join(d, "out.fa")
        with open(src, "wb") as f:
            f.write(text)
        p = subprocess.run([REF, src, dst, str(length)], capture_output=True)
        return open(dst, "rb").read(), p.stdout


def main():
    rng = random.Random(11)
    cases = []

    def add(name, text, length):
        out, stdout = run(text, length)
        cases.append({"name": name, "input": base64.b64encode(text).decode(), "length": length,
                      "output": base64.b64encode(out).decode(), "stdout": stdout.decode()})

    add("bundled_style", b">1\nACGTACGTAC;\n>2\nACG;\n>3\nTTTTGGGGCCCCAAAA;\n", 5)
    add("lowercase_and_headers", b">a_long_header_line_here\nacgtnacgtn;\nACGT\n", 4)
    add("no_trailing_newline", b">1\nACGTACGT;\n>2\nGGGGCCCCAA;", 3)
    add("crlf", b">1\r\nACGTACGT;\r\n>2\r\nAC;\r\n", 4)
    add("empty_file", b"", 0)
    add("only_newlines", b"\n\n\n", 0)
    add("length_zero", b"A\nAC\n\nACG\n", 0)
    add("negative_length", b"ACGTACGT;\nAC;\n", -1)
    add("huge_length", b"ACGTACGT;\nAC;\n", 1000)
    add("non_numeric_length", b"ACGTACGT;\nAC;\n", "x7")
    for i in range(12):
        lines = []
        for _ in range(rng.randrange(1, 25)):
            if rng.random() < 0.4:
                lines.append(b">" + str(rng.randrange(10 ** rng.randrange(1, 9))).encode())
            else:
                n = rng.choice([0, 1, 2, 5, 30, 31, 32, 100, 400])
                lines.append(bytes(rng.choice(b"ACGTacgtNn") for _ in range(n)) + rng.choice([b";", b"", b";"]))
        text = b"\n".join(lines) + rng.choice([b"", b"\n"])
        add("fuzz%d" % i, text, rng.choice([0, 1, 2, 5, 31, 32, 100]))
    json.dump(cases, open(os.path.join(ROOT, "tests", "golden", "large_unitigs_cases.json"), "w"), indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
