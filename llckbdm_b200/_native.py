"""ctypes binding of libllck.so (the C ABI declared in include/llck.h).

The product path has NO CPU fallback: if the CUDA extension is missing or no CUDA device is
present, every compute entry point raises.  ``load()`` itself only needs the shared object, so the
CPU test-suite can check that the library loads and exports every declared symbol.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libllck.so")

SYMBOLS = (
    "llck_version",
    "llck_release_resources",
    "llck_leading_dim",
    "llck_workspace_bytes",
    "llck_debug_offset",
    "llck_kbdm_batched",
    "llck_zgemm",
    "llck_bidiag_test",
    "llck_bdc_test",
    "llck_rmse_batched",
    "llck_silhouette_batched",
    "llck_pool_features",
    "llck_hdbscan_core_distances",
    "llck_hdbscan_mst",
    "llck_hdbscan_labels",
    "llck_multi_fid_batched",
)

FLAG_DEBUG_KEEP = 1
FLAG_TIMING = 2
FLAG_NO_GRAPH = 4
MST_SINGLE_CTA = 1
MST_DIM3 = 2

E_BADARG = 1
E_WORKSPACE = 2
E_SHORT_SIGNAL = 3
E_TOO_LARGE = 4
M_MAX = 2048

SVD_DC = 0
SVD_JACOBI = 1


class Options(ctypes.Structure):
    """``llck_options`` of include/llck.h: explicit tuning knobs of llck_kbdm_batched (the library reads no environment)."""
    _fields_ = [("struct_size", ctypes.c_int32), ("svd_mode", ctypes.c_int32), ("cluster_size", ctypes.c_int32),
                ("aed_window", ctypes.c_int32), ("aed_nibble", ctypes.c_int32), ("jacobi_max_sweeps", ctypes.c_int32),
                ("jacobi_conv", ctypes.c_double), ("hqr_profile", ctypes.c_void_p)]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = ctypes.sizeof(Options)

STATUS_OK = 0
STATUS_QR_NOCONV = 1
STATUS_SINGULAR = 2
STATUS_NONFINITE = 3
STATUS_SVD_NOCONV = 4

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load libllck.so (built in-tree by ``__graft_entry__.build()``); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). llckbdm_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    c_int, c_i64, c_sz, c_vp, c_dbl = ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_double
    lib.llck_version.restype = c_int
    lib.llck_version.argtypes = []
    lib.llck_release_resources.restype = c_int
    lib.llck_release_resources.argtypes = []
    lib.llck_leading_dim.restype = c_int
    lib.llck_leading_dim.argtypes = [c_int]
    lib.llck_workspace_bytes.restype = c_sz
    lib.llck_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.llck_debug_offset.restype = c_sz
    lib.llck_debug_offset.argtypes = [c_int, c_int, c_int]
    lib.llck_kbdm_batched.restype = c_int
    lib.llck_kbdm_batched.argtypes = [
        c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
        c_int, c_dbl, c_dbl, c_int,
        c_vp, c_i64,
        c_vp, c_vp, c_i64,
        c_vp, c_i64,
        c_vp, c_vp,
        c_vp, c_sz, c_int, ctypes.POINTER(Options),
        c_vp, ctypes.POINTER(c_int),
    ]
    lib.llck_zgemm.restype = c_int
    lib.llck_zgemm.argtypes = [c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]
    lib.llck_bidiag_test.restype = c_int
    lib.llck_bidiag_test.argtypes = [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.llck_bdc_test.restype = c_int
    lib.llck_bdc_test.argtypes = [c_vp, c_vp, ctypes.POINTER(c_int), c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.llck_rmse_batched.restype = c_int
    lib.llck_rmse_batched.argtypes = [c_vp, c_int, c_dbl, c_vp, c_i64, c_vp, c_int, c_int, c_dbl, c_vp, c_vp]
    lib.llck_silhouette_batched.restype = c_int
    lib.llck_silhouette_batched.argtypes = [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]
    lib.llck_pool_features.restype = c_int
    lib.llck_pool_features.argtypes = [c_vp, c_i64, c_vp, c_vp, c_int, c_dbl, c_dbl, c_vp, c_vp, c_vp]
    lib.llck_hdbscan_core_distances.restype = c_int
    lib.llck_hdbscan_core_distances.argtypes = [c_vp, c_int, c_int, c_vp, c_vp]
    lib.llck_hdbscan_mst.restype = c_int
    lib.llck_hdbscan_mst.argtypes = [c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]
    lib.llck_hdbscan_labels.restype = c_int
    lib.llck_hdbscan_labels.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]
    lib.llck_multi_fid_batched.restype = c_int
    lib.llck_multi_fid_batched.argtypes = [c_vp, c_i64, c_vp, c_int, c_int, c_dbl, c_vp, c_vp]
    _lib = lib
    return lib


_E_TEXT = {
    E_BADARG: "bad argument",
    E_WORKSPACE: "workspace too small",
    E_SHORT_SIGNAL: "a member's signal is shorter than the 2m + p - 1 points its Hankel matrices use",
    E_TOO_LARGE: f"Hankel dimension m above the supported maximum ({M_MAX})",
}


def check_rc(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: CUDA error {-rc}")
    raise ValueError(f"{what}: {_E_TEXT.get(rc, 'bad argument')} (code {rc})")
