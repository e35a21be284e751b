"""
oracle/refarm.py -- ctypes wrapper of oracle/_ref/libocvref.so, the REFERENCE CPU arm (test / measurement infrastructure).

The library is the unmodified reference CPU build behind oracle/refgen/ref_arm.cpp (recipe: oracle/build_ref.sh).  It only
exists where that build was available (the build container) or travelled to (the GPU box); available() says which.
Only tests/ and bench.py's CPU-baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libocvref.so")
_lib = None


def build():
    """Runs oracle/build_ref.sh when the reference CPU build is present; returns True when the library exists afterwards."""
    try:
        subprocess.run(["sh", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    except Exception:                    # noqa: BLE001
        pass
    return available()


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.refarm_create.restype = C.c_void_p
        L.refarm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.refarm_stitch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
        L.refarm_out_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.refarm_destroy.argtypes = [C.c_void_p]
        L.refarm_destroy.restype = None
        _lib = L
    return _lib


class RefArm:
    """The reference's CPU composition of one frame on a "VRv11" template file."""

    def __init__(self, dat_path, in_size, blend, gain):
        self.h = lib().refarm_create(dat_path.encode(), int(in_size[0]), int(in_size[1]), int(blend), int(bool(gain)))
        if not self.h:
            raise RuntimeError("refarm_create failed for " + dat_path)
        w, h = C.c_int(), C.c_int()
        self.n = lib().refarm_out_size(self.h, C.byref(w), C.byref(h))
        self.out_size = (w.value, h.value)
        self.in_size = tuple(in_size)

    def stitch(self, frames_i420):
        """frames_i420: n contiguous (1.5 h, w) u8 arrays (standard I420).  Returns (y, u, v) and the gains."""
        W, H = self.out_size
        fr = [np.ascontiguousarray(f, np.uint8) for f in frames_i420]
        ptrs = (C.c_void_p * self.n)(*[f.ctypes.data for f in fr])
        out = np.empty((H * 3 // 2, W), np.uint8)
        gains = np.ones(self.n, np.float64)
        rc = lib().refarm_stitch(self.h, ptrs, out.ctypes.data_as(C.c_void_p), gains.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("refarm_stitch failed (%d)" % rc)
        flat = out.reshape(-1)
        q = (W // 2) * (H // 2)
        return (flat[:W * H].reshape(H, W), flat[W * H:W * H + q].reshape(H // 2, W // 2), flat[W * H + q:W * H + 2 * q].reshape(H // 2, W // 2)), gains

    def threads(self):
        return int(lib().refarm_threads())

    def __del__(self):
        try:
            if self.h:
                lib().refarm_destroy(self.h)
                self.h = None
        except Exception:                # noqa: BLE001
            pass
