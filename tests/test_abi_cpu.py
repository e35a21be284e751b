"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, and its host-side
template / init-time code agrees with the reference-generated golden tables (no GPU compute here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import octvr_b200 as vr
import oracle as O
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "octvr_b200.h")).read()
    declared = set(re.findall(r"\b(octvr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = vr.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared >= set(vr.SYMBOLS)
    assert b"sm_100a" in L.octvr_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    t = vr.MapperTemplate.from_arrays(*_arrays("rig2s"))
    with pytest.raises(vr.OctvrError) as e:
        vr.Mapper(t, [(192, 108)] * 2, blend=-3)
    assert e.value.code == vr.capi.ERR_CUDA


def _arrays(rig):
    t = util.template_from_gold(O, rig)
    return t.out_size, t.inputs, t.seam_masks


@pytest.mark.parametrize("rig", ["rig3", "masks", "models"])
def test_dat_roundtrip_and_host_seam_masks(rig, tmp_path):
    ot = util.template_from_gold(O, rig)
    p = str(tmp_path / "a.dat")
    O.dump_dat(ot, p)                         # writer = oracle (format pinned against the reference)
    t = vr.MapperTemplate.load(p)             # reader = product
    assert t.out_size == ot.out_size and t.num_inputs == len(ot.inputs)
    for i, d in enumerate(ot.inputs):
        e = t.input(i)
        assert e["roi"] == d["roi"]
        assert np.array_equal(e["map1"], d["map1"]) and np.array_equal(e["map2"], d["map2"])
        assert np.array_equal(e["mask"], d["mask"]) and np.array_equal(e["seam_mask"], ot.seam_masks[i])
        if d["vignette"] is not None:
            assert np.array_equal(e["vignette"], d["vignette"])
    # product writer -> oracle reader, byte-identical file
    q = str(tmp_path / "b.dat")
    t.dump(q)
    assert open(p, "rb").read() == open(q, "rb").read()
    # product's own DistanceSeamFinder (host init code) == reference-generated seam masks
    t.create_masks()
    for i in range(t.num_inputs):
        assert np.array_equal(t.input(i)["seam_mask"], ot.seam_masks[i]), (rig, i)


def test_bad_magic_and_truncation():
    with pytest.raises(vr.OctvrError) as e:
        vr.MapperTemplate.from_bytes(b"VRv10" + b"\0" * 64)
    assert e.value.code == vr.capi.ERR_FORMAT and "version does not match" in str(e.value)
    with pytest.raises(vr.OctvrError):
        vr.MapperTemplate.from_bytes(b"VRv11" + b"\1" * 11)


def test_from_arrays_validates_shapes():
    size, inputs, seams = _arrays("rig2s")
    bad = [dict(d) for d in inputs]
    bad[0]["roi"] = (0, 0, 4096, 64)
    with pytest.raises(vr.OctvrError):
        vr.MapperTemplate.from_arrays(size, bad, seams)


def test_fill_poly_matches_the_reference_build():
    """The host restatement of cv::fillPoly that draws the camera masks of a JSON config (selection rectangles, polygonal
    exclude / include masks; octvr/src/camera.cpp:96-167) against masks drawn by the reference build itself
    (tests/golden/fillpoly.npz, oracle/refgen/ref_fillpoly.cpp): convex, concave, self-intersecting, degenerate and partly /
    fully off-image polygons.  (cv2 4.13 is not a usable oracle here: it differs from the reference on every clipped edge.)"""
    L = vr.lib()
    g = np.load(os.path.join(util.GOLD, "fillpoly.npz"))
    n = int(g["n"])
    assert n >= 300
    for k in range(n):
        p = g["p%d" % k]
        w, h = int(p[0]), int(p[1])
        pts = np.ascontiguousarray(p[2:])
        want = np.unpackbits(g["m%d" % k])[:w * h].reshape(h, w).astype(bool)
        got = np.full((h, w), 7, np.uint8)
        assert L.octvr_debug_fill_poly(got.ctypes.data_as(C.c_void_p), w, h, pts.ctypes.data_as(C.c_void_p), len(pts) // 2, 200) == 0
        assert np.array_equal(got == 200, want) and set(np.unique(got)) <= {7, 200}, (k, w, h, pts.tolist())


def test_product_entry_points_fail_loudly_without_a_device():
    """No CPU fallback on the per-frame path: every constructor of a stitcher reports OCTVR_ERR_CUDA on a host without a GPU
    (this test only runs its assertions there), and argument errors are caught before any device work."""
    import torch
    L = vr.lib()
    assert L.octvr_crop_packed_frames(0, None, None, None, None, None, None, None, None) == vr.capi.ERR_INVALID
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    size, inputs, seams = _arrays("rig2s")
    t = vr.MapperTemplate.from_arrays(size, inputs, seams)
    n = len(inputs)
    for make in (lambda: vr.Mapper(t, [(192, 108)] * n, blend=-3), lambda: vr.Mapper(t, [(192, 108)] * n, blend=16),
                 lambda: vr.FastMapper(t, [(192, 108)] * n),
                 lambda: vr.MapperTemplate.from_json(util.rig_json("rig2s"), 128)):
        with pytest.raises(vr.OctvrError) as e:
            make()
        assert e.value.code == vr.capi.ERR_CUDA
    assert L.octvr_debug_seam_backend() in (-1, 0)          # seam masks, if any were made here, came from the host routine
