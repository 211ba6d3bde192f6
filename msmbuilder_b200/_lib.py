"""ctypes binding of libmsmb200.so (the C ABI declared in include/msmb200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 GPU is
visible when a compute entry point is called, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmsmb200.so")

c_i32, c_i64, c_int, c_sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
c_vp, c_dbl, c_flt = ctypes.c_void_p, ctypes.c_double, ctypes.c_float

VECTOR_METRICS = ("euclidean", "sqeuclidean", "cityblock", "chebyshev",
                  "canberra", "braycurtis", "hamming", "jaccard")
F32, F64 = 0, 1
TICA_AUTO, TICA_SIMT_F64, TICA_UMMA_3XTF32, TICA_UMMA_TF32 = 0, 1, 2, 3
TICA_UMMA_3XBF16, TICA_UMMA_6XBF16, TICA_UMMA_3XF16 = 4, 5, 6
ENGINES = {"auto": TICA_AUTO, "simt_f64": TICA_SIMT_F64,
           "umma_3xtf32": TICA_UMMA_3XTF32, "umma_tf32": TICA_UMMA_TF32,
           "umma_3xbf16": TICA_UMMA_3XBF16, "umma_6xbf16": TICA_UMMA_6XBF16,
           "umma_3xf16": TICA_UMMA_3XF16}

# name -> (restype, argtypes); kept in one table so tests can check that the
# library exports every symbol the header declares.
SIGNATURES = {
    "msmb200_abi_version": (c_int, []),
    "msmb200_last_error": (ctypes.c_char_p, []),
    "msmb200_launch_count": (ctypes.c_uint64, []),
    "msmb200_device_info": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                    ctypes.POINTER(c_int), ctypes.POINTER(c_sz)]),
    "msmb200_tica_acc_len": (c_sz, [c_int]),
    "msmb200_tica_workspace_bytes": (c_sz, [c_int, c_int]),
    "msmb200_tica_accumulate": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int, c_int,
                                        c_vp, c_vp, c_sz, c_vp]),
    "msmb200_tica_transform": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_vp, c_vp, c_vp,
                                       c_int, c_vp, c_vp]),
    "msmb200_candidate_bytes": (c_sz, [c_int, c_int]),
    "msmb200_kcenters_workspace_bytes": (c_sz, [c_int]),
    "msmb200_kcenters_pass": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_i32,
                                      c_vp, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "msmb200_candidate_select": (c_int, [c_vp, c_int, c_sz, c_int, c_int, c_vp, c_vp]),
    "msmb200_kcenters_lookahead_supported": (c_int, [c_int, c_i64, c_int, c_int]),
    "msmb200_kcenters_lane_bytes": (c_sz, [c_int]),
    "msmb200_kcenters_set_bytes": (c_sz, [c_int, c_int]),
    "msmb200_kcenters_centers_bytes": (c_sz, [c_int, c_int]),
    "msmb200_kcenters_multi_pass": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_int,
                                            c_int, c_i32, c_int, c_vp, c_vp, c_i64, c_vp, c_sz,
                                            c_vp]),
    "msmb200_kcenters_select": (c_int, [c_vp, c_i64, c_int, c_i64, c_i64, c_vp, c_int, c_vp, c_vp]),
    "msmb200_kcenters_chain": (c_int, [c_vp, c_int, c_sz, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "msmb200_candidate_from_row": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_i64, c_vp, c_vp]),
    "msmb200_assign_workspace_bytes": (c_sz, [c_i64, c_int, c_int]),
    "msmb200_assign_engine": (c_int, [c_i64, c_int, c_int]),
    "msmb200_assign_nearest": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_vp, c_int, c_int,
                                       c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "msmb200_dist": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_vp, c_int, c_vp, c_i64,
                             c_vp, c_vp]),
    "msmb200_cdist": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp]),
    "msmb200_pdist": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_i64, c_vp, c_vp]),
    "msmb200_sumdist": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_i64, c_vp,
                                c_vp]),
    "msmb200_rmsd_center": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "msmb200_rmsd_kcenters_pass": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i32, c_vp, c_vp,
                                           c_i64, c_vp, c_vp, c_sz, c_vp]),
    "msmb200_rmsd_kcenters_pass_pruned": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_sz, c_i32, c_vp,
                                                  c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "msmb200_rmsd_pass_workspace_bytes": (c_sz, [c_i64, c_i32]),
    "msmb200_rmsd_assign_nearest": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_vp,
                                            c_i64, c_vp, c_vp, c_vp, c_vp]),
    "msmb200_rmsd_dist": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_flt, c_vp, c_i64, c_vp,
                                  c_vp]),
    "msmb200_rmsd_pdist": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_vp]),
    "msmb200_kmedoids": (c_int, [c_i64, c_i64, c_vp, c_vp, ctypes.POINTER(c_dbl),
                                 ctypes.POINTER(c_i64)]),
    "msmb200_kmedoids_restarts": (c_int, [c_i64, c_i64, c_vp, c_i64, c_vp, c_vp,
                                          ctypes.POINTER(c_dbl), ctypes.POINTER(c_i64)]),
    "msmb200_contigify_ids": (c_int, [c_vp, c_i64, c_vp, ctypes.POINTER(c_i64)]),
    "msmb200_first_above": (c_int, [c_vp, c_i64, c_i64, c_dbl, c_vp, c_vp]),
    "msmb200_pooled_assign": (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_vp,
                                      c_vp, c_vp, c_vp]),
    "msmb200_label_range": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "msmb200_label_presence": (c_int, [c_vp, c_i64, c_int, c_i64, c_i64, c_vp, c_vp]),
    "msmb200_transition_counts": (c_int, [c_vp, c_int, c_vp, c_i64, c_i64, c_i64, c_vp,
                                          c_i64, c_i64, c_i32, c_vp, c_vp]),
}

_lib = None


class Msmb200Error(RuntimeError):
    pass


def load():
    """dlopen the shared library (no GPU needed for this) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Msmb200Error(
            "libmsmb200.so not found at %s. Build it with `python -m msmbuilder_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.msmb200_abi_version() != 1:
        raise Msmb200Error("libmsmb200 ABI version mismatch")
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().msmb200_last_error()
        raise Msmb200Error("%s failed (status %d): %s" % (
            what or "libmsmb200 call", status, msg.decode(errors="replace") if msg else ""))


_NVTX = os.environ.get("MSMB200_NVTX") == "1"     # NVTX range per C-ABI call (for nsys / ncu --nvtx timelines)


def call(name, *args):
    fn = getattr(load(), name)
    if not _NVTX:
        check(fn(*args), name)
        return
    import torch
    torch.cuda.nvtx.range_push(name)
    try:
        check(fn(*args), name)
    finally:
        torch.cuda.nvtx.range_pop()


def metric_id(metric):
    """ValueError for an unknown metric, worded like libdistance.pyx:122-124."""
    if isinstance(metric, bytes):
        metric = metric.decode()
    if metric not in VECTOR_METRICS:
        raise ValueError("metric must be one of %s" %
                         ", ".join("'%s'" % s for s in VECTOR_METRICS))
    return VECTOR_METRICS.index(metric)


def require_gpu():
    """Raise unless torch sees a CUDA device of compute capability 10.x."""
    import torch
    if not torch.cuda.is_available():
        raise Msmb200Error("msmbuilder_b200 needs an NVIDIA B200 (sm_100a) GPU; none is "
                           "visible and there is no CPU fallback.")
    major, _ = torch.cuda.get_device_capability()
    if major != 10:
        raise Msmb200Error("msmbuilder_b200 kernels are built for sm_100a only; found "
                           "compute capability %d.x" % major)
    load()
