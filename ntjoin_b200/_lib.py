"""ctypes binding of include/mxe.h.  Fails loudly when libmxe.so is missing: there is no fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# every symbol declared in include/mxe.h (tests check the library exports each one)
SYMBOLS = [
    "mxe_version", "mxe_last_error", "mxe_create", "mxe_destroy", "mxe_set_stream", "mxe_set_option",
    "mxe_fasta_read", "mxe_fasta_view", "mxe_fasta_name", "mxe_fasta_free",
    "mxe_sketch_file", "mxe_sketch_buffers", "mxe_prefetch_buffers", "mxe_sketch_device", "mxe_sketch_device_many", "mxe_sketch_from_arrays", "mxe_sketch_load_tsv", "mxe_sketch_view", "mxe_sketch_prefetch_host",
    "mxe_sketch_device_view", "mxe_sketch_record_start", "mxe_sketch_contig_name", "mxe_sketch_counts", "mxe_write_tsv", "mxe_sketch_free",
    "mxe_filter_and_edges", "mxe_filter_and_edges_device", "mxe_result_counts", "mxe_result_flags", "mxe_result_graph",
    "mxe_result_free", "mxe_write_dot", "mxe_timing", "mxe_timing_reset", "mxe_kernel_launches",
    "mxe_dist_mark", "mxe_dist_adjacency", "mxe_dist_edges", "mxe_dist_finish", "mxe_dist_free", "mxe_result_edge_keys",
    "mxe_a2a_partition", "mxe_a2a_mark", "mxe_a2a_sightings", "mxe_a2a_finish", "mxe_a2a_free",
    "mxe_p2p_create", "mxe_p2p_handle", "mxe_p2p_connect", "mxe_p2p_workspace", "mxe_p2p_connect_pointers", "mxe_p2p_scatter",
    "mxe_p2p_buckets", "mxe_p2p_adjacency", "mxe_p2p_edges", "mxe_p2p_finish", "mxe_p2p_free",
]


class MxeError(RuntimeError):
    """Raised when a libmxe call returns a negative code (message from mxe_last_error)."""

    def __init__(self, code, message):
        super().__init__(f"mxe error {code}: {message}")
        self.code = code


def library_path():
    return os.environ.get("MXE_LIBRARY", os.path.join(_HERE, "libmxe.so"))


def load_library():
    """Load libmxe.so (built in-tree by `make -C ntjoin_b200/csrc` or __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise MxeError(-2, f"{path} not found: build it with __graft_entry__.build() "
                           "(nvcc, sm_100a). The engine has no CPU fallback.")
    lib = C.CDLL(path)
    vp, u64p, u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    pp = C.POINTER(C.c_void_p)
    lib.mxe_version.restype = C.c_char_p
    lib.mxe_last_error.restype = C.c_char_p
    lib.mxe_create.argtypes = [C.c_int, pp]
    lib.mxe_destroy.argtypes = [vp]
    lib.mxe_destroy.restype = None
    lib.mxe_set_stream.argtypes = [vp, vp]
    lib.mxe_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    lib.mxe_sketch_file.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, pp]
    lib.mxe_sketch_buffers.argtypes = [vp, vp, u64p, C.c_uint32, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, pp]
    lib.mxe_fasta_read.argtypes = [C.c_char_p, pp]
    lib.mxe_fasta_view.argtypes = [vp, u32p, pp, pp]
    lib.mxe_fasta_name.argtypes = [vp, C.c_uint32, C.POINTER(C.c_char_p)]
    lib.mxe_fasta_free.argtypes = [vp]
    lib.mxe_fasta_free.restype = None
    lib.mxe_prefetch_buffers.argtypes = [vp, vp, C.c_uint64]
    lib.mxe_sketch_device.argtypes = [vp, vp, u64p, C.c_uint32, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, pp]
    lib.mxe_sketch_device_many.argtypes = [vp, C.c_int, pp, C.POINTER(u64p), u32p, C.c_int, C.c_int, C.c_int, pp]
    lib.mxe_sketch_from_arrays.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, C.POINTER(C.c_char_p), u64p, C.c_uint32, C.c_int, vp, pp]
    lib.mxe_sketch_load_tsv.argtypes = [vp, C.c_char_p, pp]
    lib.mxe_sketch_view.argtypes = [vp, u64p, pp, pp, pp, pp, pp]
    lib.mxe_sketch_prefetch_host.argtypes = [vp]
    lib.mxe_sketch_device_view.argtypes = [vp, u64p, pp, pp, pp]
    lib.mxe_sketch_record_start.argtypes = [vp, C.c_uint32, u64p]
    lib.mxe_sketch_contig_name.argtypes = [vp, C.c_uint32, C.POINTER(C.c_char_p)]
    lib.mxe_sketch_counts.argtypes = [vp, u64p, u64p, u64p, u64p, u32p]
    lib.mxe_write_tsv.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.mxe_sketch_free.argtypes = [vp]
    lib.mxe_sketch_free.restype = None
    lib.mxe_filter_and_edges.argtypes = [vp, pp, C.c_int, C.POINTER(C.c_double), pp]
    lib.mxe_filter_and_edges_device.argtypes = [vp, pp, pp, u64p, C.c_int, C.POINTER(C.c_double), pp]
    lib.mxe_result_counts.argtypes = [vp, u64p, u64p, u64p]
    lib.mxe_result_flags.argtypes = [vp, C.c_int, u64p, pp, pp]
    lib.mxe_result_graph.argtypes = [vp, u64p, pp, u64p, pp, pp, pp, pp]
    lib.mxe_result_free.argtypes = [vp]
    lib.mxe_result_free.restype = None
    lib.mxe_dist_mark.argtypes = [vp, vp, u64p, C.c_int, C.c_int, C.c_int, vp, pp, u64p]
    lib.mxe_dist_adjacency.argtypes = [vp, vp, u64p, u64p, u64p, pp, vp]
    lib.mxe_dist_edges.argtypes = [vp, vp, vp, u64p]
    lib.mxe_dist_finish.argtypes = [vp, vp, C.POINTER(C.c_double), pp]
    lib.mxe_dist_free.argtypes = [vp]
    lib.mxe_dist_free.restype = None
    lib.mxe_result_edge_keys.argtypes = [vp, u64p, pp]
    lib.mxe_a2a_partition.argtypes = [vp, pp, u64p, C.c_int, C.c_int, C.c_int, pp, u64p, pp]
    lib.mxe_a2a_mark.argtypes = [vp, vp, u64p, vp, u64p]
    lib.mxe_a2a_sightings.argtypes = [vp, vp, pp, u64p, u64p, pp]
    lib.mxe_a2a_finish.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), pp]
    lib.mxe_a2a_free.argtypes = [vp]
    lib.mxe_a2a_free.restype = None
    lib.mxe_p2p_create.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_int, pp]
    lib.mxe_p2p_handle.argtypes = [vp, vp, u64p]
    lib.mxe_p2p_connect.argtypes = [vp, vp]
    lib.mxe_p2p_workspace.argtypes = [vp, pp]
    lib.mxe_p2p_connect_pointers.argtypes = [vp, pp]
    lib.mxe_p2p_scatter.argtypes = [vp, pp, pp, u64p, C.c_int, C.POINTER(C.c_double)]
    lib.mxe_p2p_buckets.argtypes = [vp]
    lib.mxe_p2p_adjacency.argtypes = [vp]
    lib.mxe_p2p_edges.argtypes = [vp]
    lib.mxe_p2p_finish.argtypes = [vp, pp]
    lib.mxe_p2p_free.argtypes = [vp]
    lib.mxe_p2p_free.restype = None
    lib.mxe_write_dot.argtypes = [C.c_char_p, C.c_uint64, vp, C.c_int, C.POINTER(C.c_char_p), pp, pp, pp,
                                  C.c_uint64, vp, vp, vp, C.POINTER(C.c_char_p)]
    lib.mxe_timing.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), u64p]
    lib.mxe_timing_reset.argtypes = [vp]
    lib.mxe_kernel_launches.argtypes = [vp]
    lib.mxe_kernel_launches.restype = C.c_uint64
    _LIB = lib
    return lib


def check(lib, rc):
    if rc != 0:
        raise MxeError(rc, lib.mxe_last_error().decode("utf-8", "replace"))
