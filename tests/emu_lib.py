"""TEST INFRASTRUCTURE ONLY: ctypes view of tests/emu/libbgr_emu.so, the host single-stepper of the
device code (see tests/emu/bgr_emu.cu). Built on demand with nvcc (host code only runs)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "bgr_emu.cu")
LIB = os.path.join(HERE, "emu", "libbgr_emu.so")
DEPS = [SRC] + [os.path.join(ROOT, "bgreat_b200", "csrc", f)
                for f in ("bgr_kernels.cuh", "bgr_dev.cuh", "bgr_scan.cuh", "bgr_extend.cuh", "bgr_emit.cuh", "bgr_common.h",
                          "bgr_host.cpp", "bgr_host.h", "bgr_pack.cpp")]

_lib = None


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler",
                           "-fPIC", "-shared", "-o", LIB, SRC,
                           os.path.join(ROOT, "bgreat_b200", "csrc", "bgr_host.cpp"),
                           os.path.join(ROOT, "bgreat_b200", "csrc", "bgr_pack.cpp")])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.emu_map.restype = C.c_int
    return _lib


def _concat(seqs):
    b = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(b) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in b], dtype=np.uint64)
    return b"".join(b), off


def pack_reads(reads, threads=1):
    """The product's host packer (bgr_pack.cpp) -> (words uint64 array, set of flagged read indices)."""
    L = lib()
    L.emu_pack_reads.restype = C.c_uint32
    rs, roff = _concat(reads)
    n = len(reads)
    words = np.full(((len(rs) + 31) >> 5) + 2, 0xDEADBEEFDEADBEEF, dtype=np.uint64)
    bad = np.full((n + 31) // 32 + 1, 0xFFFFFFFF, dtype=np.uint32)
    nb = L.emu_pack_reads(rs, roff.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(threads),
                          words.ctypes.data_as(C.c_void_p), bad.ctypes.data_as(C.c_void_p))
    flagged = {i for i in range(n) if (int(bad[i >> 5]) >> (i & 31)) & 1}
    assert nb == len(flagged)
    return words[:(len(rs) + 31) >> 5], flagged


def emu_map(unitigs, k, reads, mode, errors, effort, scratch_ints=64, how=0):
    """-> ([(overlap_found, rc, path)], counters dict, flags list, scratch_overflow).
    how: bit 0 = packed upload form, bit 1 = general scan form only (see bgr_emu.cu emu_map2)."""
    L = lib()
    us, uoff = _concat(unitigs)
    rs, roff = _concat(reads)
    n = len(reads)
    res = np.zeros((max(n, 1), 2), dtype=np.uint32)
    pool_cap = max(1024, (len(rs) + 2 * n) * 2)
    pool = np.zeros(pool_cap, dtype=np.int32)
    used = C.c_uint64(0)
    counters = (C.c_uint64 * 8)()
    ovf = C.c_int(0)
    L.emu_map2.restype = C.c_int
    rc = L.emu_map2(us, uoff.ctypes.data_as(C.c_void_p), C.c_uint32(len(unitigs)), C.c_uint32(k), rs,
                    roff.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(mode), C.c_uint32(errors),
                    C.c_uint32(effort), C.c_uint32(scratch_ints), res.ctypes.data_as(C.c_void_p),
                    pool.ctypes.data_as(C.c_void_p), C.c_uint64(pool_cap), C.byref(used), counters, C.byref(ovf),
                    C.c_uint32(how))
    if rc != 0:
        raise RuntimeError("emu_map failed: %d" % rc)
    out, flags = [], []
    for i in range(n):
        fl = int(res[i, 0]) & 0xFF
        ln = int(res[i, 0]) >> 8
        o = int(res[i, 1])
        out.append((bool(fl & 2), bool(fl & 4), [int(x) for x in pool[o:o + ln]]))
        flags.append(fl)
    names = ["reads", "aligned", "not_aligned", "no_overlap", "overlaps", "windows", "verified"]
    return out, dict(zip(names, [int(c) for c in counters[:7]])), flags, bool(ovf.value)


def parse_reads(path, k, fastq=False, chunk=10000, threads=0, range_bytes=1 << 20, min_slice=4096):
    """The PRODUCT's read parser -> [(header, read)] or None. threads == 0: the sequential ReadParser
    (bgr_host.cpp) driven in `chunk`-read getReads calls; threads > 0: the ParallelReader the CLI uses."""
    L = lib()
    L.emu_parse_reads.restype = C.c_long
    L.emu_parse_reads_parallel.restype = C.c_long
    ph, pho, ps, pso = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    if threads:
        n = L.emu_parse_reads_parallel(os.fsencode(path), C.c_uint32(k), C.c_int(int(fastq)), C.c_uint32(threads),
                                       C.c_uint64(range_bytes), C.c_uint64(min_slice), C.byref(ph), C.byref(pho), C.byref(ps), C.byref(pso))
    else:
        n = L.emu_parse_reads(os.fsencode(path), C.c_uint32(k), C.c_int(int(fastq)), C.c_uint32(chunk), C.byref(ph),
                              C.byref(pho), C.byref(ps), C.byref(pso))
    if n < 0:
        return None
    ho = C.cast(pho, C.POINTER(C.c_uint64))
    so = C.cast(pso, C.POINTER(C.c_uint64))
    hb = C.string_at(ph, ho[n])
    sb = C.string_at(ps, so[n])
    out = [(hb[ho[i]:ho[i + 1]], sb[so[i]:so[i + 1]]) for i in range(n)]
    L.emu_free.argtypes = [C.c_void_p]
    for p in (ph, pho, ps, pso):
        L.emu_free(p)
    return out


def load_unitigs(path, k, threads=0, min_slice=1):
    """The PRODUCT's unitig loader -> (opened, [sequences]). threads == 0: the sequential loop
    (bgr_host.cpp load_unitig_file); threads > 0: load_unitig_file_parallel, what bgr_index_build_file uses."""
    L = lib()
    L.emu_load_unitigs.restype = C.c_long
    L.emu_load_unitigs_parallel.restype = C.c_long
    ps, po, opened = C.c_void_p(), C.c_void_p(), C.c_int()
    if threads:
        n = L.emu_load_unitigs_parallel(os.fsencode(path), C.c_uint32(k), C.c_uint32(threads), C.c_uint64(min_slice),
                                        C.byref(ps), C.byref(po), C.byref(opened))
    else:
        n = L.emu_load_unitigs(os.fsencode(path), C.c_uint32(k), C.byref(ps), C.byref(po), C.byref(opened))
    off = C.cast(po, C.POINTER(C.c_uint64))
    sb = C.string_at(ps, off[n])
    out = [sb[off[i]:off[i + 1]] for i in range(n)]
    L.emu_free.argtypes = [C.c_void_p]
    L.emu_free(ps)
    L.emu_free(po)
    return bool(opened.value), out


def correct(unitigs, k, paths, lens, rcs):
    """correct_read (the device function of the -c mode, single-stepped on the host) -> [bytes]"""
    L = lib()
    us, uoff = _concat(unitigs)
    n = len(paths)
    ints = np.array([x for p in paths for x in p] + [0], dtype=np.int32)
    poff = np.zeros(n + 1, dtype=np.uint64)
    poff[1:] = np.cumsum([len(p) for p in paths], dtype=np.uint64)
    lens_a = np.array(list(lens) + [0], dtype=np.uint32)
    rcs_a = np.array([int(x) for x in rcs] + [0], dtype=np.int32)
    out = C.create_string_buffer(int(sum(lens)) + 1)
    got = np.zeros(n + 1, dtype=np.uint32)
    rc = L.emu_correct(us, uoff.ctypes.data_as(C.c_void_p), C.c_uint32(len(unitigs)), C.c_uint32(k),
                       ints.ctypes.data_as(C.c_void_p), poff.ctypes.data_as(C.c_void_p), lens_a.ctypes.data_as(C.c_void_p),
                       rcs_a.ctypes.data_as(C.c_void_p), C.c_uint32(n), out, got.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("emu_correct failed: %d" % rc)
    res, o = [], 0
    for i in range(n):
        res.append(out.raw[o:o + int(got[i])])
        o += int(lens[i])
    return res
