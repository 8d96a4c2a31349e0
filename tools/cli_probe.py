#!/usr/bin/env python
"""Where the drop-in command line spends its time: one synthetic read file in /dev/shm, the `bgreat` binary run
under a few settings with BGR_VERBOSE=1 (phase marks + busy time per pipeline stage on stderr).
Measurement aid, not part of the product.   python tools/cli_probe.py --reads 20000000"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--fastq", action="store_true")
    a = ap.parse_args()
    cores = min(len(os.sched_getaffinity(0)), 255)
    d = tempfile.mkdtemp(prefix="bgr_probe_", dir="/dev/shm")
    try:
        g = synth.Graph(a.genome, 31, seed=1)
        g.write_unitigs(os.path.join(d, "unitigs.fa"))
        seqs, off = g.make_reads(a.reads, 150, 0.01)
        ext = "fq" if a.fastq else "fa"
        synth.write_reads(os.path.join(d, "reads." + ext), seqs, off, a.reads, fastq=a.fastq)
        del seqs, off
        base = ["-r", "reads." + ext, "-g", "unitigs.fa", "-k", "31", "-m", "2"] + (["-q"] if a.fastq else [])
        runs = [("default -t %d" % cores, ["-t", str(cores), "-f", "paths", "-a", "na"], {})]
        if not os.environ.get("BGR_PROBE_ONLY_ENVS"):
            runs += [("again", ["-t", str(cores), "-f", "paths", "-a", "na"], {}),
                     ("outputs to /dev/null", ["-t", str(cores), "-f", "/dev/null", "-a", "/dev/null"], {}),
                     ("BGR_OUT=write", ["-t", str(cores), "-f", "paths", "-a", "na"], {"BGR_OUT": "write"}),
                     ("-t 8", ["-t", "8", "-f", "paths", "-a", "na"], {}),
                     ("-t 32", ["-t", "32", "-f", "paths", "-a", "na"], {})]
        for extra in os.environ.get("BGR_PROBE_ENVS", "").split(";"):
            if extra:
                kv = dict(x.split("=", 1) for x in extra.split(","))
                runs.append((extra, ["-t", str(cores), "-f", "paths", "-a", "na"], kv))
        for name, args, env in runs:
            e = dict(os.environ, BGR_VERBOSE="1", **env)
            t0 = time.perf_counter()
            p = subprocess.run([OURS] + base + args, cwd=d, capture_output=True, text=True, env=e)
            dt = time.perf_counter() - t0
            print("==== %s: rc=%d wall %.3f s (%.1f M reads/s)" % (name, p.returncode, dt, a.reads / dt / 1e6))
            print(p.stderr.strip())
            sys.stdout.flush()
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
