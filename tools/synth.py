"""ctypes wrapper of tools/libbgr_synth.so: the synthetic genome / unitig / read generator
 used by bench.py and the large parity tests."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbgr_synth.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "bgr_synth.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = C.CDLL(LIB)
        L.bgs_make_graph2.restype = C.c_void_p
        L.bgs_make_graph2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_double, C.c_double, C.c_int, C.c_int]
        L.bgs_free_graph.argtypes = [C.c_void_p]
        for f in ("bgs_n_unitigs", "bgs_n_snps", "bgs_n_repeat_copies", "bgs_unitig_bases", "bgs_genome_len"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.bgs_unitig_seqs.restype = C.c_void_p
        L.bgs_unitig_seqs.argtypes = [C.c_void_p]
        L.bgs_unitig_off.restype = C.c_void_p
        L.bgs_unitig_off.argtypes = [C.c_void_p]
        L.bgs_fill.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.bgs_write_unitigs.argtypes = [C.c_void_p, C.c_char_p]
        L.bgs_make_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_double,
                                     C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.bgs_write_reads.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
        _lib = L
    return _lib


class Graph:
    """Simulated diploid genome and its k-mer unitigs. repeat_frac: fraction of the genome covered by repeat
    families (2..6 identical copies: borders with more than two candidates). unitigs=False builds only what
    make_reads needs (a rank that does not build the index): nothing of genome size is stored."""

    def __init__(self, genome_len, k=31, seed=1, snp_rate=1.0 / 150, repeat_frac=0.0, unitigs=True, threads=0):
        self.L = lib()
        self.k = k
        threads = threads or min(32, os.cpu_count() or 1)
        self.h = self.L.bgs_make_graph2(genome_len, k, seed, snp_rate, repeat_frac, int(unitigs), threads)
        self.n_unitigs = self.L.bgs_n_unitigs(self.h)
        self.n_snps = self.L.bgs_n_snps(self.h)
        self.n_repeat_copies = self.L.bgs_n_repeat_copies(self.h)
        self.unitig_bases = self.L.bgs_unitig_bases(self.h)
        self.genome_len = self.L.bgs_genome_len(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.bgs_free_graph(self.h)
            self.h = None

    def unitig_arrays(self):
        """-> (uint8 array of concatenated unitigs, uint64 offsets[n+1]) -- views, valid while self lives."""
        seqs = np.ctypeslib.as_array(C.cast(self.L.bgs_unitig_seqs(self.h), C.POINTER(C.c_uint8)),
                                     shape=(max(1, self.unitig_bases),))
        off = np.ctypeslib.as_array(C.cast(self.L.bgs_unitig_off(self.h), C.POINTER(C.c_uint64)),
                                    shape=(self.n_unitigs + 1,))
        return seqs, off

    def unitig_list(self):
        seqs, off = self.unitig_arrays()
        b = seqs.tobytes()
        return [b[int(off[i]):int(off[i + 1])].decode() for i in range(self.n_unitigs)]

    def haplotype(self, h, pos=0, length=None):
        """bases [pos, pos+length) of haplotype h (regenerated: the genome is not stored)"""
        length = self.genome_len - pos if length is None else length
        buf = C.create_string_buffer(length)
        if self.L.bgs_fill(self.h, h, pos, length, buf) != 0:
            raise ValueError("range outside the genome")
        return buf.raw.decode()

    def write_unitigs(self, path):
        if self.L.bgs_write_unitigs(self.h, os.fsencode(path)) != 0:
            raise IOError(path)

    def make_reads(self, n, read_len=150, sub_rate=0.01, n_rate=0.0, seed=4, first=0, threads=0, seqs=None, off=None):
        """Fill (or allocate) seqs[n*read_len] uint8 and off[n+1] uint64 with reads first..first+n."""
        if seqs is None:
            seqs = np.empty(n * read_len + 16, dtype=np.uint8)
        if off is None:
            off = np.empty(n + 1, dtype=np.uint64)
        threads = threads or min(32, os.cpu_count() or 1)
        rc = self.L.bgs_make_reads(self.h, first, n, read_len, sub_rate, n_rate, seed, seqs.ctypes.data,
                                   off.ctypes.data, threads)
        if rc != 0:
            raise ValueError("bgs_make_reads failed")
        return seqs, off


def write_reads(path, seqs, off, n, first=0, fastq=False):
    if lib().bgs_write_reads(os.fsencode(path), seqs.ctypes.data, off.ctypes.data, n, first, int(fastq)) != 0:
        raise IOError(path)
