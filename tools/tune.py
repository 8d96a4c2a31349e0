"""Kernel-time probe for tuning knobs (BGR_SCAN_MINB / BGR_EXT_MINB ...): resident reads, prints per-kernel ms."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tools import synth
import bgreat_b200
from bgreat_b200 import _abi as A

if os.environ.get("BGR_TUNE_LIB"):  # an experiment build of the library (tools only; the product loads its own)
    A.LIB_PATH = os.environ["BGR_TUNE_LIB"]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
genome = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = synth.Graph(genome, 31, seed=1)
al = bgreat_b200.Aligner(31, 2, 2, n_slots=1)
if os.environ.get("BGR_TUNE_L2GRAN"):  # experiment: cudaLimitMaxL2FetchGranularity (limit 5) of the primary context
    rt = C.CDLL("libcudart.so.12")
    print("cudaDeviceSetLimit ->", rt.cudaDeviceSetLimit(5, C.c_size_t(int(os.environ["BGR_TUNE_L2GRAN"]))), file=sys.stderr)
us, uo = g.unitig_arrays()
al.index_unitigs_packed(us, uo)
seqs, off = g.make_reads(n, 150, 0.01)
L = A.lib()
p = al.params(mode)
assert L.bgr_slot_upload(al.ctx, 0, seqs.ctypes.data, off.ctypes.data, n) == 0
ms = (C.c_float * 3)()
acc = np.zeros(3)
for it in range(6):
    assert L.bgr_slot_run(al.ctx, 0, C.byref(p)) == 0
    assert L.bgr_slot_wait(al.ctx, 0, None) == 0
    assert L.bgr_slot_kernel_ms(al.ctx, 0, ms) == 0
    if it >= 2:
        acc += np.array(list(ms))
acc /= 4
print(json.dumps({"scan_minb": os.environ.get("BGR_SCAN_MINB"), "ext_minb": os.environ.get("BGR_EXT_MINB"), "reads": n,
                  "genome": genome, "mode": mode, "scan_ms": acc[0], "ext_ms": acc[1], "total_ms": acc[2],
                  "Mreads_per_s": n / acc[2] / 1e3, "index": al.index_stats(), "counters": al.counters()}))
