#!/usr/bin/env python
"""Small mapping workload for compute-sanitizer (memcheck / racecheck / synccheck): both upload forms, greedy
and exhaustive, reads with N, a batch of several tiles per CTA.
  compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_small as S  # noqa: E402
import bgreat_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
rng = random.Random(3)
unitigs, haps = S.make_graph(rng, 31, genome_len=20000, mean_unitig=70, n_variants=130)
reads = S.make_reads(rng, haps, n, 150, sub_rate=0.01, n_rate=0.0003)
reads += S.make_reads(rng, haps, 200, 700, sub_rate=0.002, n_rate=0.0)
al = bgreat_b200.Aligner(31, 2, 2)
al.index_unitigs(unitigs)
a = al.align_batch(reads, 0).as_tuples()
b = al.align_batch_packed(reads, 0, threads=4).as_tuples()
assert a == b
c = al.align_batch(reads[:20000], 1).as_tuples()
d = al.align_batch_packed(reads[:20000], 1, threads=4).as_tuples()
assert c == d
# the large-index filter tier (minimizer blocks) and a text batch (records parsed / output formatted on the device)
os.environ["BGR_BLOOM"] = "global"
al2 = bgreat_b200.Aligner(31, 2, 2)
al2.index_unitigs(unitigs)
e = al2.align_batch(reads[:30000], 0).as_tuples()
assert e == a[:30000]
text = "".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads[:30000]))
regular, n_text, streams = al2.map_text(text, False, 0, True)
assert regular and n_text > 0 and sum(len(x) for x in streams) > 0
print("ok", len(reads), sum(1 for x in a if x[2]), sum(1 for x in c if x[2]), n_text)
