#!/usr/bin/env python
"""End-to-end benchmark of the drop-in `bgreat` command line against the reference binary, on files.

Synthetic config of BASELINE.json configs[1] (5 Mbp graph, 150 bp reads, 1 % substitutions, -m 2) written
as FASTA (or FASTQ with --fastq) to /dev/shm; runs bgreat_b200/bin/bgreat on all N reads and
oracle/_ref/bgreat -t <cores> on the first --ref-reads of them, times both from process start to exit
(index build included for both, reported separately from the tools' own stdout where available) and
checks that the two produce the same `paths` / `notAligned.fa` records on the common prefix
(sorted: the reference's order with -t > 1 is unspecified).

Prints one JSON line. This is a measurement aid for SURVEY 8(f) rank 1 (host parse + format at GPU rate),
not bench.py's metric."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

OURS = os.path.join(ROOT, "bgreat_b200", "bin", "bgreat")
REF = os.path.join(ROOT, "oracle", "_ref", "bgreat")


def run(exe, args, cwd):
    t0 = time.perf_counter()
    p = subprocess.run([exe] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if os.environ.get("BGR_VERBOSE"):
        sys.stderr.write(p.stderr)
    if p.returncode != 0:
        raise RuntimeError("%s failed rc=%d: %s" % (exe, p.returncode, p.stderr[-400:]))
    info = {}
    for l in p.stdout.splitlines():
        if " : " in l:
            k, v = l.split(" : ", 1)
            info[k.strip()] = v.split()[0]
    return dt, info


def records(path, limit_ids=None):
    """sorted (header, body) pairs of a two-lines-per-record file"""
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    return sorted(zip(lines[0::2], lines[1::2]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--ref-reads", type=int, default=2_000_000)
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--threads", type=int, default=0, help="-t for both tools (0 = all cores)")
    ap.add_argument("--fastq", action="store_true")
    ap.add_argument("--errors", type=int, default=2)
    ap.add_argument("--brute", action="store_true")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    cores = a.threads or min(len(os.sched_getaffinity(0)), 255)
    base = "/dev/shm" if os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    d = tempfile.mkdtemp(prefix="bgr_cli_", dir=base)
    try:
        g = synth.Graph(a.genome, 31, seed=1)
        g.write_unitigs(os.path.join(d, "unitigs.fa"))
        seqs, off = g.make_reads(a.reads, 150, 0.01)
        ext = "fq" if a.fastq else "fa"
        full, part = os.path.join(d, "reads." + ext), os.path.join(d, "reads_ref." + ext)
        synth.write_reads(full, seqs, off, a.reads, fastq=a.fastq)
        nref = min(a.ref_reads, a.reads)
        synth.write_reads(part, seqs, off, nref, fastq=a.fastq)
        common = ["-g", "unitigs.fa", "-k", 31, "-m", a.errors, "-t", cores] + (["-q"] if a.fastq else []) + \
                 (["-b"] if a.brute else [])
        best, info = None, None
        for _ in range(a.repeat):
            dt, info = run(OURS, ["-r", full, "-f", "paths.ours", "-a", "na.ours"] + common, d)
            best = dt if best is None else min(best, dt)
        out = {"tool": "bgreat_b200/bin/bgreat", "reads": a.reads, "format": ext, "threads": cores,
               "wall_s": best, "reads_per_s_wall": a.reads / best, "stdout": info,
               "input_bytes": os.path.getsize(full),
               "output_bytes": os.path.getsize(os.path.join(d, "paths.ours")) + os.path.getsize(os.path.join(d, "na.ours"))}
        if not a.no_ref and os.access(REF, os.X_OK):
            dt_r, info_r = run(REF, ["-r", part, "-f", "paths.ref", "-a", "na.ref"] + common, d)
            dt_e, _ = run(REF, ["-r", "/dev/null", "-f", "p0", "-a", "n0"] + common, d)
            out["reference"] = {"reads": nref, "wall_s": dt_r, "index_only_wall_s": dt_e,
                                "reads_per_s_wall": nref / dt_r, "reads_per_s_mapping": nref / max(dt_r - dt_e, 1e-6),
                                "stdout": info_r}
            out["speedup_wall"] = out["reads_per_s_wall"] / out["reference"]["reads_per_s_wall"]
            # parity on the common prefix, through the files
            dt_o, info_o = run(OURS, ["-r", part, "-f", "paths.ours_p", "-a", "na.ours_p"] + common, d)
            same_counts = all(info_o.get(k) == info_r.get(k) for k in
                              ("Reads", "No overlap", "Got overlap", "Overlap and aligned", "Overlap but not aligned"))
            if a.brute:  # the reference writes no files in -b mode
                out["parity"] = {"counters_equal": same_counts}
            else:
                same_p = records(os.path.join(d, "paths.ours_p")) == records(os.path.join(d, "paths.ref"))
                same_n = records(os.path.join(d, "na.ours_p")) == records(os.path.join(d, "na.ref"))
                out["parity"] = {"counters_equal": same_counts, "paths_equal_sorted": same_p,
                                 "not_aligned_equal_sorted": same_n}
        print(json.dumps(out), flush=True)
    finally:
        if not a.keep:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
