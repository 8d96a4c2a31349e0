#!/usr/bin/env python
"""Index build time: GPU builder (default) vs host builder (BGR_INDEX_BUILD=host), same unitigs.
usage: index_bench.py GENOME_BP [--no-host]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402
import bgreat_b200  # noqa: E402

genome = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
g = synth.Graph(genome, 31, seed=1)
us, uo = g.unitig_arrays()
out = {"genome_bp": genome, "unitigs": int(g.n_unitigs), "unitig_bases": int(g.unitig_bases)}
for where in ["gpu", "gpu"] + ([] if "--no-host" in sys.argv else ["host"]):
    os.environ["BGR_INDEX_BUILD"] = where
    al = bgreat_b200.Aligner(31, 2, 2, n_slots=1)
    t0 = time.perf_counter()
    al.index_unitigs_packed(us, uo)
    dt = time.perf_counter() - t0
    out[where + "_build_s"] = dt
    out["stats"] = al.index_stats()
    al.close()
print(json.dumps(out))
