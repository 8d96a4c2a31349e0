"""ctypes binding of libbgreat_b200.so (include/bgreat_b200.h). No fallback: if the CUDA library is
missing or no sm_100 device is present, importing / creating a context raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libbgreat_b200.so")

BGR_C_COUNT = 8
COUNTER_NAMES = ["reads", "aligned", "not_aligned", "no_overlap", "overlaps", "windows", "verified"]

MODE_GREEDY, MODE_EXHAUSTIVE, MODE_GREEDY_PATH, MODE_EXHAUSTIVE_PATH = 0, 1, 2, 3
F_MAPPED, F_OVERLAP_FOUND, F_RC = 0x01, 0x02, 0x04
F_CNT_ALIGNED, F_CNT_NOTALIGNED, F_CNT_NOOVERLAP = 0x10, 0x20, 0x40

ERR_NAMES = {0: "BGR_OK", 1: "BGR_ERR_CUDA", 2: "BGR_ERR_ARG", 3: "BGR_ERR_IO", 4: "BGR_ERR_NO_INDEX",
             5: "BGR_ERR_FORMAT", 6: "BGR_ERR_LIMIT", 7: "BGR_ERR_BUSY"}

# every symbol include/bgreat_b200.h declares
SYMBOLS = ["bgr_abi_version", "bgr_ctx_create", "bgr_ctx_destroy", "bgr_last_error", "bgr_host_alloc",
           "bgr_host_free", "bgr_index_build", "bgr_index_build_file", "bgr_index_stats", "bgr_index_blob_size",
           "bgr_index_export_device", "bgr_index_import_device", "bgr_index_export_host", "bgr_index_import_host",
           "bgr_index_query", "bgr_map_batch", "bgr_slot_upload", "bgr_slot_run", "bgr_slot_download",
           "bgr_slot_wait", "bgr_slot_download_corrected", "bgr_slot_kernel_ms", "bgr_counters", "bgr_counters_reset", "bgr_launch_count",
           "bgr_pack_reads", "bgr_slot_upload_packed", "bgr_index_save_file", "bgr_index_load_file", "bgr_device_count",
           "bgr_index_replicate", "bgr_slot_upload_text", "bgr_slot_format_text"]


P_CORRECT = 1  # bgr_params.flags: correction mode (-c)


class Params(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("errors", C.c_uint32), ("effort", C.c_uint32), ("flags", C.c_uint32)]


class BgrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(bgreat_b200 has no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.bgr_abi_version.restype = C.c_int
    L.bgr_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.bgr_ctx_destroy.argtypes = [vp]
    L.bgr_ctx_destroy.restype = None
    L.bgr_last_error.argtypes = [vp]
    L.bgr_last_error.restype = C.c_char_p
    L.bgr_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.bgr_host_free.argtypes = [vp]
    L.bgr_host_free.restype = None
    L.bgr_index_build.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32]
    L.bgr_index_build_file.argtypes = [vp, C.c_char_p, C.c_uint32]
    L.bgr_index_stats.argtypes = [vp, u64p, u64p, u64p]
    L.bgr_index_blob_size.argtypes = [vp, u64p]
    L.bgr_index_export_device.argtypes = [vp, vp, C.c_uint64]
    L.bgr_index_import_device.argtypes = [vp, vp, C.c_uint64]
    L.bgr_index_export_host.argtypes = [vp, vp, C.c_uint64]
    L.bgr_index_import_host.argtypes = [vp, vp, C.c_uint64]
    L.bgr_index_query.argtypes = [vp, C.c_uint64, C.c_int, C.POINTER(C.c_int32)]
    L.bgr_index_save_file.argtypes = [vp, C.c_char_p]
    L.bgr_index_load_file.argtypes = [vp, C.c_char_p, C.c_uint32]
    L.bgr_device_count.argtypes = [C.POINTER(C.c_int)]
    L.bgr_index_replicate.argtypes = [vp, vp]
    L.bgr_slot_upload_text.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.bgr_slot_format_text.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.bgr_map_batch.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(Params), vp, vp, C.c_uint64, u64p]
    L.bgr_slot_upload.argtypes = [vp, C.c_int, vp, vp, C.c_uint32]
    L.bgr_slot_run.argtypes = [vp, C.c_int, C.POINTER(Params)]
    L.bgr_pack_reads.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp, C.POINTER(C.c_uint32)]
    L.bgr_slot_upload_packed.argtypes = [vp, C.c_int, vp, vp, C.c_uint32, vp, C.c_uint32, vp]
    L.bgr_slot_download.argtypes = [vp, C.c_int, vp, vp, C.c_uint64]
    L.bgr_slot_wait.argtypes = [vp, C.c_int, u64p]
    L.bgr_slot_download_corrected.argtypes = [vp, C.c_int, vp, C.c_uint64]
    L.bgr_slot_kernel_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.bgr_counters.argtypes = [vp, u64p]
    L.bgr_counters_reset.argtypes = [vp]
    L.bgr_launch_count.argtypes = [vp]
    L.bgr_launch_count.restype = C.c_uint64
    for name in ("bgr_ctx_create", "bgr_host_alloc", "bgr_index_build", "bgr_index_build_file", "bgr_index_stats",
                 "bgr_index_blob_size", "bgr_index_export_device", "bgr_index_import_device",
                 "bgr_index_export_host", "bgr_index_import_host", "bgr_index_save_file", "bgr_index_load_file", "bgr_index_query", "bgr_map_batch",
                 "bgr_slot_upload", "bgr_pack_reads", "bgr_slot_upload_packed", "bgr_slot_run", "bgr_slot_download", "bgr_slot_wait", "bgr_slot_download_corrected", "bgr_slot_kernel_ms",
                 "bgr_counters", "bgr_counters_reset"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L
