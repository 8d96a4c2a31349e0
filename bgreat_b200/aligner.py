"""Host-side mirror of the reference's `class Aligner` (reference aligner.h:58-179) over the C ABI.

Same vocabulary as the reference: an Aligner is built for (k, errorsMax, tryNumber), indexes unitigs
(indexUnitigs, aligner.cpp:537) and aligns reads (alignReadGreedy / alignReadExhaustive, aligner.h:127,135)
-- here a batch at a time on the GPU.
"""
import ctypes as C

import numpy as np

from . import _abi as A


class MappingResult:
    """Per-read results of one batch: flags (uint8), path offsets/lengths and the path pool."""

    def __init__(self, flags_len, path_off, pool):
        self.flags = (flags_len & 0xFF).astype(np.uint8)
        self.path_len = (flags_len >> 8).astype(np.uint32)
        self.path_off = path_off
        self.pool = pool

    def __len__(self):
        return len(self.flags)

    def path(self, i):
        o, n = int(self.path_off[i]), int(self.path_len[i])
        return [int(x) for x in self.pool[o:o + n]]

    def overlap_found(self, i):
        return bool(self.flags[i] & A.F_OVERLAP_FOUND)

    def rc(self, i):
        return bool(self.flags[i] & A.F_RC)

    def mapped(self, i):
        return bool(self.flags[i] & A.F_MAPPED)

    def as_tuples(self):
        """[(overlap_found, rc, path)] -- the shape the reference's per-read API returns."""
        return [(self.overlap_found(i), self.rc(i), self.path(i)) for i in range(len(self))]


def _concat(seqs):
    b = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(b) + 1, dtype=np.uint64)
    if b:
        off[1:] = np.cumsum([len(x) for x in b], dtype=np.uint64)
    return np.frombuffer(b"".join(b) + b"\0", dtype=np.uint8), off


class Aligner:
    """Aligner(k, errors=2, effort=2, device=0): reference ctor arguments kValue, errorsAllowed, effort
    (aligner.h:80). Modes: greedy (default, `bgreat`), exhaustive (`bgreat -b`)."""

    def __init__(self, k, errors=2, effort=2, device=0, n_slots=2):
        self.L = A.lib()
        self.k, self.errors, self.effort = k, errors, effort
        self.ctx = C.c_void_p()
        rc = self.L.bgr_ctx_create(device, n_slots, C.byref(self.ctx))
        if rc:
            raise A.BgrError(rc, self.L.bgr_last_error(None).decode())
        self.n_slots = n_slots

    def close(self):
        if getattr(self, "ctx", None):
            self.L.bgr_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise A.BgrError(rc, self.L.bgr_last_error(self.ctx).decode())

    # ---- index (indexUnitigs, aligner.cpp:537)
    def index_unitigs(self, unitigs):
        """unitigs: sequences in file order (ids 1..n)."""
        buf, off = _concat(unitigs)
        self._ck(self.L.bgr_index_build(self.ctx, buf.ctypes.data, off.ctypes.data, len(unitigs), self.k))

    def index_unitigs_packed(self, buf, off):
        """buf: uint8 array of concatenated sequences, off: uint64 array of n+1 offsets."""
        self._ck(self.L.bgr_index_build(self.ctx, buf.ctypes.data, off.ctypes.data, len(off) - 1, self.k))

    def index_file(self, path):
        self._ck(self.L.bgr_index_build_file(self.ctx, path.encode(), self.k))

    def index_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.L.bgr_index_stats(self.ctx, C.byref(a), C.byref(b), C.byref(c)))
        return dict(unitigs=a.value, keys=b.value, device_bytes=c.value)

    def index_blob_size(self):
        n = C.c_uint64()
        self._ck(self.L.bgr_index_blob_size(self.ctx, C.byref(n)))
        return n.value

    def export_index_device(self, dev_ptr, nbytes):
        self._ck(self.L.bgr_index_export_device(self.ctx, dev_ptr, nbytes))

    def import_index_device(self, dev_ptr, nbytes):
        self._ck(self.L.bgr_index_import_device(self.ctx, dev_ptr, nbytes))

    def export_index_host(self):
        n = self.index_blob_size()
        out = np.empty(n, dtype=np.uint8)
        self._ck(self.L.bgr_index_export_host(self.ctx, out.ctypes.data, n))
        return out

    def import_index_host(self, blob):
        self._ck(self.L.bgr_index_import_host(self.ctx, blob.ctypes.data, blob.nbytes))

    def save_index(self, path):
        """the index blob as a file (bgr_index_save_file)"""
        self._ck(self.L.bgr_index_save_file(self.ctx, path.encode()))

    def load_index(self, path):
        """index from a file written by save_index, for this aligner's k (bgr_index_load_file)"""
        self._ck(self.L.bgr_index_load_file(self.ctx, path.encode(), self.k))

    def get_begin(self, code):
        """getBegin (aligner.cpp:209): signed ids of the unitigs beginning with the (k-1)-mer code."""
        return self._query(code, 0)

    def get_end(self, code):
        """getEnd (aligner.cpp:147)."""
        return self._query(code, 1)

    def _query(self, code, want_end):
        ids = (C.c_int32 * 4)()
        n = self.L.bgr_index_query(self.ctx, code, want_end, ids)
        if n < 0:
            raise A.BgrError(-n, self.L.bgr_last_error(self.ctx).decode())
        return list(ids[:n])

    # ---- mapping
    def params(self, mode, flags=0):
        return A.Params(mode, self.errors, self.effort, flags)

    def correct_batch(self, reads):
        """Correction mode (`bgreat -c`, alignerGreedy.cpp:394-404): -> (MappingResult, [corrected read as
        bytes, or None for a read that does not map]). Uses the slot API (upload / run / download)."""
        buf, off = _concat(reads)
        n = len(off) - 1
        res = np.zeros((max(n, 1), 2), dtype=np.uint32)
        cap = max(1024, int(off[-1]) + 2 * n + 16)
        pool = np.zeros(cap, dtype=np.int32)
        nw = (int(off[-1]) >> 5) + n + 2
        words = np.zeros(nw, dtype=np.uint64)
        used = C.c_uint64()
        p = self.params(A.MODE_GREEDY, A.P_CORRECT)
        self._ck(self.L.bgr_slot_upload(self.ctx, 0, buf.ctypes.data, off.ctypes.data, n))
        self._ck(self.L.bgr_slot_run(self.ctx, 0, C.byref(p)))
        self._ck(self.L.bgr_slot_download(self.ctx, 0, res.ctypes.data, pool.ctypes.data, cap))
        self._ck(self.L.bgr_slot_download_corrected(self.ctx, 0, words.ctypes.data, nw))
        self._ck(self.L.bgr_slot_wait(self.ctx, 0, C.byref(used)))
        r = MappingResult(res[:n, 0], res[:n, 1], pool[:used.value])
        out = []
        for i in range(n):
            if not r.mapped(i):
                out.append(None)
                continue
            ln = int(off[i + 1] - off[i])
            w0 = (int(off[i]) >> 5) + i
            bits = words[w0:w0 + (ln + 31) // 32]
            s = bytearray()
            for j in range(ln):
                s.append(b"ACGT"[(int(bits[j >> 5]) >> (62 - 2 * (j & 31))) & 3])
            out.append(bytes(s))
        return r, out

    def align_batch(self, reads, mode=A.MODE_GREEDY):
        """reads: list of str/bytes. -> MappingResult (through bgr_map_batch, host buffers)."""
        buf, off = _concat(reads)
        return self.align_packed(buf, off, mode)

    def pack_reads(self, buf, off, threads=1):
        """bgr_pack_reads: the batch at 2 bits per base -> (words uint64, bad bitmap uint32, flagged reads)."""
        n = len(off) - 1
        words = np.zeros(max(1, (int(off[-1]) + 31) >> 5), dtype=np.uint64)
        bad = np.zeros(max(1, (n + 31) >> 5), dtype=np.uint32)
        n_bad = C.c_uint32()
        self._ck(self.L.bgr_pack_reads(buf.ctypes.data, off.ctypes.data, n, threads, words.ctypes.data, bad.ctypes.data,
                                       C.byref(n_bad)))
        return words, bad, n_bad.value

    def align_batch_packed(self, reads, mode=A.MODE_GREEDY, threads=1):
        """The same batch call with the reads travelling as a 2-bit stream (bgr_pack_reads on the host,
        bgr_slot_upload_packed); results are identical to align_batch's."""
        buf, off = _concat(reads)
        n = len(off) - 1
        words, bad, n_bad = self.pack_reads(buf, off, threads)
        res = np.zeros((max(n, 1), 2), dtype=np.uint32)
        cap = max(1024, int(off[-1]) + 2 * n + 16)
        pool = np.zeros(cap, dtype=np.int32)
        used = C.c_uint64()
        p = self.params(mode)
        self._ck(self.L.bgr_slot_upload_packed(self.ctx, 0, words.ctypes.data, off.ctypes.data, n, bad.ctypes.data, n_bad,
                                               buf.ctypes.data))
        self._ck(self.L.bgr_slot_run(self.ctx, 0, C.byref(p)))
        self._ck(self.L.bgr_slot_download(self.ctx, 0, res.ctypes.data, pool.ctypes.data, cap))
        self._ck(self.L.bgr_slot_wait(self.ctx, 0, C.byref(used)))
        return MappingResult(res[:n, 0], res[:n, 1], pool[:used.value])

    def align_packed(self, buf, off, mode=A.MODE_GREEDY):
        n = len(off) - 1
        res = np.zeros((max(n, 1), 2), dtype=np.uint32)
        cap = max(1024, int(off[-1]) + 2 * n + 16)  # a path never has more ints than the read has bases + 2
        pool = np.zeros(cap, dtype=np.int32)
        used = C.c_uint64()
        p = self.params(mode)
        self._ck(self.L.bgr_map_batch(self.ctx, buf.ctypes.data, off.ctypes.data, n, C.byref(p), res.ctypes.data,
                                      pool.ctypes.data, cap, C.byref(used)))
        return MappingResult(res[:n, 0], res[:n, 1], pool[:used.value])

    def map_text(self, text, fastq=False, mode=A.MODE_GREEDY, split_no_overlap=False, flags=0):
        """A run of whole FASTA / FASTQ records as text -> (regular, n_reads, [paths, notAligned, noOverlap] as
        bytes): records parsed and the output formatted on the device (bgr_slot_upload_text / bgr_slot_format_text).
        regular False: the text is not of the regular shape and nothing was mapped."""
        text = text.encode() if isinstance(text, str) else bytes(text)
        buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, dtype=np.uint8)
        n, regular = C.c_uint32(0), C.c_int(1)
        self._ck(self.L.bgr_slot_upload_text(self.ctx, 0, buf.ctypes.data, len(text), int(fastq), C.byref(n), C.byref(regular)))
        if not regular.value:
            return False, 0, [b"", b"", b""]
        p = self.params(mode, flags)
        self._ck(self.L.bgr_slot_run(self.ctx, 0, C.byref(p)))
        caps = [64, 64, 64]
        for _ in range(2):
            outs = [np.zeros(c, dtype=np.uint8) for c in caps]
            ptrs = (C.c_void_p * 3)(*[o.ctypes.data for o in outs])
            cap = (C.c_uint64 * 3)(*caps)
            used = (C.c_uint64 * 3)()
            rc = self.L.bgr_slot_format_text(self.ctx, 0, int(split_no_overlap), ptrs, cap, used)
            if rc == 6:  # BGR_ERR_LIMIT: used[] holds the sizes needed
                caps = [int(u) + 16 for u in used]
                continue
            self._ck(rc)
            return True, n.value, [outs[f][:int(used[f])].tobytes() for f in range(3)]
        raise A.BgrError(6, "output buffers")

    def counters(self):
        out = (C.c_uint64 * A.BGR_C_COUNT)()
        self._ck(self.L.bgr_counters(self.ctx, out))
        return dict(zip(A.COUNTER_NAMES, [int(x) for x in out]))

    def reset_counters(self):
        self._ck(self.L.bgr_counters_reset(self.ctx))

    def launch_count(self):
        return int(self.L.bgr_launch_count(self.ctx))
