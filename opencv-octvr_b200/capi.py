"""
ctypes binding of liboctvr_b200.so (include/octvr_b200.h).  Fails loudly when the library is
missing: there is no Python / CPU fallback for the stitch path.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCTVR_LIB") or os.path.join(_HERE, "liboctvr_b200.so")   # OCTVR_LIB: diagnostic A/B builds only

OK, ERR_INVALID, ERR_FORMAT, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3, -4


class OctvrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("octvr_b200 error %d: %s" % (code, msg))
        self.code = code


class Frame(C.Structure):
    """octvr_frame"""
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("y_pitch", C.c_size_t), ("u_pitch", C.c_size_t), ("v_pitch", C.c_size_t),
                ("uv_pixel_stride", C.c_int)]


SYMBOLS = [
    "octvr_last_error", "octvr_version",
    "octvr_template_load_dat", "octvr_template_load_file", "octvr_template_dump_file",
    "octvr_template_build_json", "octvr_template_create", "octvr_template_add_input", "octvr_template_from_arrays", "octvr_template_add_overlay", "octvr_template_create_masks", "octvr_debug_seam_backend", "octvr_fast_create", "octvr_fast_stitch_nv12", "octvr_fast_info", "octvr_fast_debug_table", "octvr_fast_destroy",
    "octvr_template_out_size", "octvr_template_num_inputs", "octvr_template_num_overlays",
    "octvr_template_input", "octvr_template_destroy",
    "octvr_mapper_create", "octvr_mapper_stitch", "octvr_mapper_stitch_packed", "octvr_mapper_set_keep_rgb", "octvr_mapper_result_rgb",
    "octvr_mapper_source_rows", "octvr_mapper_source_cols", "octvr_mapper_set_input_window", "octvr_crop_packed_frames", "octvr_shared_alloc", "octvr_shared_open", "octvr_shared_close", "octvr_mapper_debug_gain_ns", "octvr_mapper_debug_ring", "octvr_mapper_debug_gain_trace", "octvr_debug_fill_poly", "octvr_mapper_create_band", "octvr_mapper_create_window",
    "octvr_mapper_gains", "octvr_mapper_stats", "octvr_mapper_set_profiling", "octvr_mapper_stage_ms",
    "octvr_mapper_destroy",
    "octvr_async_create", "octvr_async_push", "octvr_async_pop", "octvr_async_fps", "octvr_async_preview", "octvr_async_destroy",
]

_lib = None


def build(verbose=False):
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")], stdout=out, stderr=out)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OctvrError(ERR_CUDA, "liboctvr_b200.so is not built (run __graft_entry__.build()); "
                                       "the stitch path has no fallback")
        L = C.CDLL(LIB_PATH)
        L.octvr_last_error.restype = C.c_char_p
        L.octvr_version.restype = C.c_char_p
        L.octvr_template_destroy.restype = None
        L.octvr_mapper_destroy.restype = None
        L.octvr_async_destroy.restype = None
        L.octvr_fast_destroy.restype = None
        L.octvr_fast_destroy.argtypes = [C.c_void_p]
        L.octvr_fast_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.octvr_fast_stitch_nv12.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.octvr_fast_debug_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.octvr_mapper_stitch.argtypes = [C.c_void_p, C.POINTER(Frame), C.c_int, C.POINTER(Frame), C.c_void_p, C.c_size_t,
                                          C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p]
        L.octvr_mapper_stitch_packed.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int,
                                                 C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_int, C.c_void_p]
        L.octvr_async_push.argtypes = [C.c_void_p, C.POINTER(Frame), C.c_int, C.POINTER(Frame)]
        L.octvr_async_pop.argtypes = [C.c_void_p]
        L.octvr_async_destroy.argtypes = [C.c_void_p]
        L.octvr_mapper_destroy.argtypes = [C.c_void_p]
        L.octvr_template_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise OctvrError(rc, lib().octvr_last_error().decode("utf-8", "replace"))
