"""ctypes binding of include/mcmcb200.h (libmcmcb200.so).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a only).  There is no
CPU fallback anywhere in this package: if the shared object is missing the import fails, and if no
B200 is visible every call that needs the device raises ``McmcError(ENODEVICE)``.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCMCB200_LIB") or os.path.join(HERE, "libmcmcb200.so")   # MCMCB200_LIB: kernel-variant experiments

OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, EUNSUPPORTED, ETAPE, ESTATE = 0, -1, -2, -3, -4, -5, -6, -7
PROPOSAL_UNIFORM, PROPOSAL_DYNAMIC = 0, 1
CONVERGE_VERTICES, CONVERGE_EDGES = 0, 1
FLAG_NO_FUSED_FINALIZE = 1
FLAG_NO_EARLY_STOP = 2
FLAG_FORCE_DIRECT = 4
FLAG_FORCE_BLOCKED = 8
FLAG_FORCE_BINNED = 16
FLAG_NO_OVERLAP = 32
VIEW_COLORS_CUR, VIEW_COLORS_NEXT, VIEW_COUNTERS = 0, 1, 2

# every symbol include/mcmcb200.h declares (tests/test_capi_symbols.py checks the export table against the header)
SYMBOLS = [
    "mcmcb200_create", "mcmcb200_create_partition", "mcmcb200_create_device_csr", "mcmcb200_destroy",
    "mcmcb200_init_colors", "mcmcb200_set_tape", "mcmcb200_sweep", "mcmcb200_status", "mcmcb200_get_colors",
    "mcmcb200_get_class_sizes", "mcmcb200_get_history", "mcmcb200_tailcut", "mcmcb200_conflicts_of",
    "mcmcb200_debug_occupancy", "mcmcb200_debug_all_occupancy", "mcmcb200_device_view", "mcmcb200_finalize_sweep",
    "mcmcb200_stream", "mcmcb200_synchronize", "mcmcb200_last_sweep_ms", "mcmcb200_launch_count", "mcmcb200_kernel_mode", "mcmcb200_csr_from_edges", "mcmcb200_csr_free",
    "mcmcb200_strerror", "mcmcb200_last_cuda_error", "mcmcb200_abi_version", "mcmcb200_luby_color",
    "mcmcb200_ipc_export", "mcmcb200_ipc_attach", "mcmcb200_ipc_detach", "mcmcb200_init_colors_slice", "mcmcb200_init_colors_finish",
    "mcmcb200_get_colors_slice", "mcmcb200_color_bytes", "mcmcb200_init_colors_narrow", "mcmcb200_get_colors_narrow",
    "mcmcb200_tailcut_dist_begin", "mcmcb200_tailcut_dist_mark", "mcmcb200_tailcut_dist_round", "mcmcb200_tailcut_dist_apply",
    "mcmcb200_tailcut_dist_recount", "mcmcb200_tailcut_dist_end", "mcmcb200_layout_bytes",
    "mcmcb200_init_colors_slice_narrow", "mcmcb200_get_colors_slice_narrow",
]


class Params(C.Structure):
    """mcmcb200_params (superset of ColoringMCMCParams, graph_coloring/coloring.h:65-74)."""
    _fields_ = [("nCol", C.c_uint32), ("epsilon", C.c_float), ("lambda_", C.c_float), ("numColorRatio", C.c_float),
                ("ratioFreezed", C.c_float), ("tabooIteration", C.c_uint32), ("maxRip", C.c_uint32),
                ("tailcut", C.c_uint32), ("proposal", C.c_uint32), ("convergence", C.c_uint32),
                ("seed", C.c_uint64), ("device", C.c_int32), ("flags", C.c_uint32),
                ("stageCapBytes", C.c_uint32), ("itemBits", C.c_uint32), ("stageBuffers", C.c_uint32), ("expectedSweeps", C.c_uint32)]


class Status(C.Structure):
    _fields_ = [("sweep", C.c_uint32), ("converged", C.c_int32), ("conflictEdges", C.c_uint64),
                ("violatingVertices", C.c_uint64), ("usedColors", C.c_uint32), ("countsSweep", C.c_uint32),
                ("z", C.c_uint64)]


class McmcError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__(f"{what}: error {code} ({detail})")


_lib = None


def lib():
    """Load libmcmcb200.so.  Raises ImportError (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built.  Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32p = C.c_void_p, C.POINTER(C.c_uint32)
    L.mcmcb200_create.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint64, vp, vp, C.POINTER(Params)]
    L.mcmcb200_create_partition.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.POINTER(Params)]
    L.mcmcb200_create_device_csr.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp,
                                             C.POINTER(Params)]
    L.mcmcb200_destroy.argtypes = [vp]
    L.mcmcb200_destroy.restype = None
    L.mcmcb200_init_colors.argtypes = [vp, vp]
    L.mcmcb200_set_tape.argtypes = [vp, vp, C.c_uint32]
    L.mcmcb200_sweep.argtypes = [vp, C.c_uint32]
    L.mcmcb200_status.argtypes = [vp, C.POINTER(Status)]
    L.mcmcb200_get_colors.argtypes = [vp, vp]
    L.mcmcb200_get_class_sizes.argtypes = [vp, vp]
    L.mcmcb200_get_history.argtypes = [vp, vp, C.c_uint32, u32p]
    L.mcmcb200_tailcut.argtypes = [vp, C.c_uint32, u32p]
    L.mcmcb200_conflicts_of.argtypes = [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.mcmcb200_debug_occupancy.argtypes = [vp, C.c_uint32, vp]
    L.mcmcb200_debug_all_occupancy.argtypes = [vp, vp, vp]
    L.mcmcb200_device_view.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64), u32p]
    L.mcmcb200_finalize_sweep.argtypes = [vp]
    L.mcmcb200_init_colors_slice.argtypes = [vp, vp]
    L.mcmcb200_init_colors_finish.argtypes = [vp]
    L.mcmcb200_get_colors_slice.argtypes = [vp, vp]
    L.mcmcb200_color_bytes.argtypes = [vp, u32p]
    L.mcmcb200_init_colors_narrow.argtypes = [vp, vp, C.c_uint32]
    L.mcmcb200_get_colors_narrow.argtypes = [vp, vp, C.c_uint32]
    L.mcmcb200_tailcut_dist_begin.argtypes = [vp, vp, vp, C.c_uint32, u32p]
    L.mcmcb200_tailcut_dist_mark.argtypes = [vp, vp, C.c_uint32]
    L.mcmcb200_tailcut_dist_round.argtypes = [vp, vp, vp, C.c_uint32, u32p, u32p, u32p]
    L.mcmcb200_tailcut_dist_apply.argtypes = [vp, vp, vp, C.c_uint32]
    L.mcmcb200_tailcut_dist_recount.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), u32p]
    L.mcmcb200_tailcut_dist_end.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32]
    L.mcmcb200_layout_bytes.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mcmcb200_init_colors_slice_narrow.argtypes = [vp, vp, C.c_uint32]
    L.mcmcb200_get_colors_slice_narrow.argtypes = [vp, vp, C.c_uint32]
    L.mcmcb200_ipc_export.argtypes = [vp, vp]
    L.mcmcb200_ipc_attach.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.mcmcb200_ipc_detach.argtypes = [vp]
    L.mcmcb200_stream.argtypes = [vp, C.POINTER(vp)]
    L.mcmcb200_synchronize.argtypes = [vp]
    L.mcmcb200_last_sweep_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.mcmcb200_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mcmcb200_kernel_mode.argtypes = [vp, C.POINTER(C.c_int)]
    L.mcmcb200_csr_from_edges.argtypes = [C.c_uint32, C.c_uint64, vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.mcmcb200_csr_free.argtypes = [vp, vp]
    L.mcmcb200_csr_free.restype = None
    L.mcmcb200_luby_color.argtypes = [C.c_uint32, C.c_uint64, vp, vp, C.c_uint64, C.c_int32, vp, u32p, u32p]
    L.mcmcb200_strerror.argtypes = [C.c_int]
    L.mcmcb200_strerror.restype = C.c_char_p
    L.mcmcb200_last_cuda_error.restype = C.c_char_p
    L.mcmcb200_abi_version.restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != OK:
        L = lib()
        detail = L.mcmcb200_strerror(rc).decode()
        if rc in (ECUDA, ENODEVICE, ENOMEM):
            detail += "; " + L.mcmcb200_last_cuda_error().decode()
        raise McmcError(rc, what, detail)


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.c_void_p)
