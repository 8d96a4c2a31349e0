#!/usr/bin/env python
"""Per-read parity against the oracle on a graph of BASELINE.json's full size (1 Gbp / 3 Gbp: 4..13 GB index, blob
offsets beyond 4 GB, the filter at its cap). The oracle needs minutes and ~7 GB per Gbp to index such a graph, which
is why this is a script (run on the GPU box; its output is committed under profiles/) and not part of `pytest -m gpu`,
where the 3 Gbp graph is covered through size-independent properties (tests/test_gpu_large_graph.py).

  python tools/parity_large.py GENOME_BP [N_READS]
Prints one JSON line: sizes, per-mode agreement, sha256 of the GPU results (flags + paths, read order)."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import synth  # noqa: E402
import bgreat_b200 as bg  # noqa: E402

genome = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
K = 31
t0 = time.time()
g = synth.Graph(genome, K, seed=1, repeat_frac=0.001)
us, uo = g.unitig_arrays()
al = bg.Aligner(K, 2, 2)
al.index_unitigs_packed(us, uo)
t1 = time.time()
o = O.Oracle(K, 2, 2)
o.index_arrays(us, uo)
t2 = time.time()
out = {"genome_bp": genome, "reads_per_case": n, "index": al.index_stats(), "gpu_index_s": t1 - t0, "oracle_index_s": t2 - t1,
       "cases": []}
ok_all = True
for mode, m, sub, cnt in ((O.GREEDY, 2, 0.01, n), (O.GREEDY, 0, 0.01, n // 4), (O.GREEDY, 5, 0.01, n // 4),
                          (O.EXHAUSTIVE, 2, 0.02, n // 2), (O.EXHAUSTIVE_PATH, 2, 0.02, n // 10), (O.GREEDY_PATH, 2, 0.01, n // 20)):
    seqs, off = g.make_reads(cnt, 150, sub, n_rate=0.0002, first=17 * cnt)
    b = seqs.tobytes()
    o.set_params(m, 2)
    o.reset_counters()
    want = [o.align(b[int(off[i]):int(off[i + 1])], mode) for i in range(cnt)]
    al.errors = m
    al.reset_counters()
    res = al.align_packed(seqs, off, mode)
    got = res.as_tuples()
    c, oc = al.counters(), o.counters()
    same = got == want and all(c[x] == oc[x] for x in ("reads", "aligned", "not_aligned", "no_overlap", "overlaps"))
    ok_all &= same
    h = hashlib.sha256()
    h.update(res.flags.tobytes())
    h.update(res.path_len.tobytes())
    h.update(res.pool[:int(res.path_len.sum())].tobytes())
    out["cases"].append({"mode": O.MODE_NAMES[mode], "m": m, "sub_rate": sub, "reads": cnt, "identical_to_oracle": bool(same),
                         "aligned": c["aligned"], "not_aligned": c["not_aligned"], "no_overlap": c["no_overlap"],
                         "sha256": h.hexdigest()})
out["all_identical"] = bool(ok_all)
out["wall_s"] = time.time() - t0
print(json.dumps(out))
sys.exit(0 if ok_all else 1)
