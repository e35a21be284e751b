"""
opencv-octvr_b200 -- B200-native (sm_100a) implementation of the octvr per-frame stitch path.

The product is liboctvr_b200.so (C ABI: include/octvr_b200.h; sources in csrc/).  This package is
the thin Python mirror of the reference's mapper API used by the tests and bench.py.  The directory
name has a hyphen; import it as `import octvr_b200` (alias module at the repo root).
"""
from .capi import OctvrError, Frame, lib, build, LIB_PATH, SYMBOLS  # noqa: F401
from .mapper import MapperTemplate, Mapper, AsyncMultiMapper, FastMapper, PackedRGB, frame_from_planes, split_packed  # noqa: F401
from . import sharding  # noqa: F401,E402
