"""Step-by-step GPU bring-up with timings (debug aid; prints flush immediately)."""
import os
import sys
import time
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import faulthandler
faulthandler.enable()
faulthandler.dump_traceback_later(100, exit=True)


def log(*a):
    print("[%.3f]" % time.time(), *a, flush=True)


t0 = time.time()
import bgreat_b200
from oracle import oracle as O
import synth_small as S
log("imports done")
al = bgreat_b200.Aligner(4, 2, 2)
log("ctx created")
al.index_unitigs(["AAATC", "ATCGGA", "GGATT"])
log("index built", al.index_stats())
log("getBegin(ATC)", al.get_begin(O.str2num("ATC")), "getEnd(ATC)", al.get_end(O.str2num("ATC")))
r = al.align_batch(["AAATCGGATT", "AAATCGCATT"], 0)
log("greedy", r.as_tuples(), al.counters())
r = al.align_batch(["AAATCGGATT", "AAATCGCATT"], 1)
log("exhaustive", r.as_tuples(), al.counters())
al.close()
log("closed")
for seed in range(5):
    c = S.fuzz_case(seed)
    for mode in (0, 1):
        al = bgreat_b200.Aligner(c["k"], c["m"], c["effort"])
        al.index_unitigs(c["unitigs"])
        t = time.time()
        got = al.align_batch(c["reads"], mode).as_tuples()
        dt = time.time() - t
        o = O.Oracle(c["k"], c["m"], c["effort"])
        o.index_unitigs(c["unitigs"])
        want = [o.align(x, mode) for x in c["reads"]]
        log("fuzz", seed, "mode", mode, "k", c["k"], "m", c["m"], "ok" if got == want else "MISMATCH", "%.3fs" % dt)
        if got != want:
            for i, (a, b) in enumerate(zip(want, got)):
                if a != b:
                    log("  read", i, c["reads"][i], "want", a, "got", b)
                    break
        al.close()
rng = random.Random(7)
unitigs, haps = S.make_graph(rng, 31, genome_len=20000, mean_unitig=70, n_variants=130)
reads = S.make_reads(rng, haps, 3000, 150, sub_rate=0.01, n_rate=0.0005)
for mode in (0, 1):
    al = bgreat_b200.Aligner(31, 2, 2)
    al.index_unitigs(unitigs)
    t = time.time()
    got = al.align_batch(reads, mode).as_tuples()
    dt = time.time() - t
    o = O.Oracle(31, 2, 2)
    o.index_unitigs(unitigs)
    want = [o.align(x, mode) for x in reads]
    log("k31 mode", mode, "ok" if got == want else "MISMATCH", "%.3fs" % dt, al.counters())
    al.close()
log("done in %.1fs" % (time.time() - t0))
