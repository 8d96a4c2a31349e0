"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/_build/liboracle.so (the CPU restatement,
oracle/bgreat_oracle.c) and thin runners for the compiled reference in oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. Nothing under bgreat_b200/ does.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
REF_BGREAT = os.path.join(HERE, "_ref", "bgreat")
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")

GREEDY, EXHAUSTIVE, GREEDY_PATH, EXHAUSTIVE_PATH = 0, 1, 2, 3
MODE_NAMES = {GREEDY: "greedy", EXHAUSTIVE: "exhaustive", GREEDY_PATH: "greedy_path",
              EXHAUSTIVE_PATH: "exhaustive_path"}
COUNTER_NAMES = ["reads", "aligned", "not_aligned", "no_overlap", "overlaps", "windows", "lookups",
                 "queries", "compares", "cmp_bases"]


def build(with_ref=True):
    """(Re)build the oracle library, and the reference binaries when /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if with_ref else []))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build(with_ref=False)
        L = C.CDLL(LIB)
        L.bo_new.restype = C.c_void_p
        L.bo_new.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.bo_free.argtypes = [C.c_void_p]
        L.bo_set_params.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.bo_index_file.restype = C.c_int64
        L.bo_index_file.argtypes = [C.c_void_p, C.c_char_p]
        L.bo_index_mem.restype = C.c_int64
        L.bo_index_mem.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64), C.c_uint32]
        L.bo_align.restype = C.c_int
        L.bo_align.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_uint32, C.POINTER(C.c_int),
                               C.POINTER(C.c_int), C.POINTER(C.c_int32), C.c_int]
        L.bo_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bo_reset_counters.argtypes = [C.c_void_p]
        L.bo_parse_reads.restype = C.c_int64
        L.bo_parse_reads.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.bo_free_buf.argtypes = [C.c_void_p]
        L.bo_recover_path.restype = C.c_int64
        L.bo_recover_path.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_uint32, C.c_int, C.c_char_p]
        L.bo_run_correct.restype = C.c_int
        L.bo_run_correct.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]
        L.bo_run.restype = C.c_int
        L.bo_run.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]
        L.bo_str2num.restype = C.c_uint64
        L.bo_str2num.argtypes = [C.c_char_p, C.c_uint32]
        L.bo_rcb.restype = C.c_uint64
        L.bo_rcb.argtypes = [C.c_uint64, C.c_uint32]
        L.bo_reverse_complement.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.bo_mismatch_number.restype = C.c_uint32
        L.bo_mismatch_number.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


class Oracle:
    """The CPU restatement of `class Aligner` (reference aligner.h:58) for one (k, m, effort)."""

    PATH_CAP = 1 << 16

    def __init__(self, k, errors=2, effort=2):
        self.L = lib()
        self.h = self.L.bo_new(k, errors, effort)
        self.k = k
        self._path = (C.c_int32 * self.PATH_CAP)()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.bo_free(self.h)
            self.h = None

    def index_file(self, path):
        return self.L.bo_index_file(self.h, os.fsencode(path))

    def index_unitigs(self, unitigs):
        """unitigs: list of str/bytes in file order (ids 1..n)."""
        seqs = [u.encode() if isinstance(u, str) else bytes(u) for u in unitigs]
        off = (C.c_uint64 * (len(seqs) + 1))()
        t = 0
        for i, s in enumerate(seqs):
            off[i] = t
            t += len(s)
        off[len(seqs)] = t
        return self.L.bo_index_mem(self.h, b"".join(seqs), off, len(seqs))

    def index_arrays(self, seqs, off):
        """unitigs as one uint8 array of concatenated sequences + uint64 offsets[n+1] (no per-unitig objects:
        graphs of millions of unitigs)"""
        return self.L.bo_index_mem(self.h, seqs.ctypes.data_as(C.c_char_p), off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                   len(off) - 1)

    def set_params(self, errors, effort):
        """other -m / -e on the graph already indexed (the index does not depend on them)"""
        self.L.bo_set_params(self.h, errors, effort)

    def align(self, read, mode=GREEDY):
        """-> (overlap_found, rc, [path ints])"""
        r = read.encode() if isinstance(read, str) else bytes(read)
        of, rc = C.c_int(0), C.c_int(0)
        n = self.L.bo_align(self.h, mode, r, len(r), C.byref(of), C.byref(rc), self._path, self.PATH_CAP)
        if n < 0:
            raise RuntimeError("oracle path buffer too small")
        return bool(of.value), bool(rc.value), list(self._path[:n])

    def counters(self):
        out = (C.c_uint64 * len(COUNTER_NAMES))()
        self.L.bo_counters(self.h, out)
        return dict(zip(COUNTER_NAMES, [int(x) for x in out]))

    def reset_counters(self):
        self.L.bo_reset_counters(self.h)

    def correct(self, read):
        """-c mode for one read: the corrected read (bytes) or None when the read does not map
        (alignReadGreedy + recoverPath, alignerGreedy.cpp:391-398)."""
        r = read.encode() if isinstance(read, str) else bytes(read)
        of, rc, path = self.align(r, GREEDY)
        if not path:
            return None
        arr = (C.c_int32 * len(path))(*path)
        out = C.create_string_buffer(len(r) + 1)
        n = self.L.bo_recover_path(self.h, arr, len(path), len(r), int(rc), out)
        if n < 0:
            raise RuntimeError("reference aborts on this path (%d)" % n)
        return out.raw[:n]

    def run_correct(self, reads_csv, fastq, paths_file, not_aligned_file):
        return self.L.bo_run_correct(self.h, os.fsencode(reads_csv), int(fastq), os.fsencode(paths_file),
                                     os.fsencode(not_aligned_file))

    def run(self, mode, reads_csv, fastq, paths_file, not_aligned_file):
        return self.L.bo_run(self.h, mode, os.fsencode(reads_csv), int(fastq), os.fsencode(paths_file),
                             os.fsencode(not_aligned_file))


def parse_reads(path, k, fastq=False, chunk=10000):
    """getReads restatement -> list of (header, read) byte strings, or None if the file cannot be opened."""
    L = lib()
    ph, pho, pr, pro = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = L.bo_parse_reads(os.fsencode(path), k, int(fastq), chunk, C.byref(ph), C.byref(pho), C.byref(pr),
                         C.byref(pro))
    if n < 0:
        return None
    ho = C.cast(pho, C.POINTER(C.c_uint64))
    ro = C.cast(pro, C.POINTER(C.c_uint64))
    hb = C.string_at(ph, ho[n]) if n else b""
    rb = C.string_at(pr, ro[n]) if n else b""
    out = [(hb[ho[i]:ho[i + 1]], rb[ro[i]:ro[i + 1]]) for i in range(n)]
    for p in (ph, pho, pr, pro):
        L.bo_free_buf(p)
    return out


def str2num(s):
    b = s.encode() if isinstance(s, str) else s
    return lib().bo_str2num(b, len(b))


def rcb(x, n):
    return lib().bo_rcb(x, n)


def reverse_complement(s):
    b = s.encode() if isinstance(s, str) else s
    out = C.create_string_buffer(len(b))
    lib().bo_reverse_complement(b, len(b), out)
    return out.raw


# ------------------------------------------------------------------ compiled reference runners


def have_ref():
    return os.access(REF_BGREAT, os.X_OK) and os.access(REF_HARNESS, os.X_OK)


def ref_harness_align(unitig_file, k, m, effort, mode, reads, workdir):
    """Run the per-read harness (unmodified reference code) -> ([(overlap_found, rc, path)], counters)."""
    rf = os.path.join(workdir, "harness_reads.txt")
    with open(rf, "wb") as f:
        for r in reads:
            f.write((r.encode() if isinstance(r, str) else r) + b"\n")
    out = subprocess.run([REF_HARNESS, "align", unitig_file, str(k), str(m), str(effort), MODE_NAMES[mode], rf],
                         check=True, capture_output=True, cwd=workdir).stdout.decode()
    res, counters = [], None
    for line in out.splitlines():
        if line.startswith("#"):
            v = [int(x) for x in line[1:].split()]
            counters = dict(zip(COUNTER_NAMES[:5], v))
            continue
        t = line.split()
        n = int(t[2])
        res.append((bool(int(t[0])), bool(int(t[1])), [int(x) for x in t[3:3 + n]]))
    return res, counters


def ref_harness_parse(reads_file, k, fastq, chunk, workdir):
    out = subprocess.run([REF_HARNESS, "parse", reads_file, str(k), str(int(fastq)), str(chunk)], check=True,
                         capture_output=True, cwd=workdir, timeout=60).stdout
    res = []
    for line in out.split(b"\n"):
        if line.startswith(b"#chunk") or line == b"":
            continue
        h, _, r = line.partition(b"\t")
        res.append((h, r))
    return res


def ref_bgreat(args, workdir, timeout=600):
    """Run the reference CLI in `workdir`; returns its stdout (str)."""
    return subprocess.run([REF_BGREAT] + [str(a) for a in args], check=True, capture_output=True, cwd=workdir,
                          timeout=timeout).stdout.decode()
