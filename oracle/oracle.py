"""
oracle/oracle.py -- CPU ORACLE wrapper (TEST INFRASTRUCTURE ONLY).

ctypes/numpy front-end for oracle/liborc.so, the plain-C restatement of the
reference's CPU algorithms (see oracle/orc.h for the file:line map).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product never does.

High-level entry points mirror the reference objects:
  build_template(cfg, width, height)  ~ octvr_dump: MapperTemplate + add_input + create_masks
                                         (modules/octvr/src/template.cpp:23-204, apps/octvr/dump.cpp:71-127)
  load_dat / dump_dat                  ~ "VRv11" (template.cpp:206-314)
  StitchOracle(...).stitch(frames)     ~ Mapper::stitch composition on the CPU contract
                                         (cvtColor -> remap -> gain -> blend -> cvtColor)
"""
import ctypes as C
import json
import math
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_image.c", "orc_blend.c", "orc_camera.c", "orc.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_fisheye_correction_radius.restype = C.c_double
        _LIB.orc_camera_aspect_ratio.restype = C.c_double
        _LIB.orc_bilinear_table.restype = C.POINTER(C.c_int16)
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def num_threads():
    return lib().orc_num_threads()


# ----------------------------------------------------------------------------- primitives
def bilinear_table():
    p = lib().orc_bilinear_table()
    return np.ctypeslib.as_array(p, shape=(32, 32, 4)).copy()


def yuv420_to_rgb(y, u, v):
    """y (H,W) u8; u,v (H/2,W/2) u8 views (any strides along x/y) -> (H,W,3) RGB."""
    h, w = y.shape
    rgb = np.empty((h, w, 3), np.uint8)
    lib().orc_yuv420_to_rgb(_p(y), C.c_ssize_t(y.strides[0]),
                            _p(u), C.c_ssize_t(u.strides[1]), C.c_ssize_t(u.strides[0]),
                            _p(v), C.c_ssize_t(v.strides[1]), C.c_ssize_t(v.strides[0]),
                            w, h, _p(rgb), C.c_ssize_t(rgb.strides[0]))
    return rgb


def rgb_to_yuv420(rgb):
    """(H,W,3) RGB -> y (H,W), u (H/2,W/2), v (H/2,W/2)."""
    rgb = np.ascontiguousarray(rgb)
    h, w = rgb.shape[:2]
    y = np.empty((h, w), np.uint8)
    u = np.empty((h // 2, w // 2), np.uint8)
    v = np.empty((h // 2, w // 2), np.uint8)
    lib().orc_rgb_to_yuv420(_p(rgb), C.c_ssize_t(rgb.strides[0]), w, h, _p(y), C.c_ssize_t(w),
                            _p(u), C.c_ssize_t(1), C.c_ssize_t(w // 2), _p(v), C.c_ssize_t(1), C.c_ssize_t(w // 2))
    return y, u, v


def scale_map(m, scale):
    m = np.ascontiguousarray(m, np.float32)
    out = np.empty_like(m)
    lib().orc_scale_map(_p(m), C.c_size_t(m.size), int(scale), _p(out))
    return out


def remap(src, mapx, mapy, linear=True):
    """cv::remap(src, mapx, mapy, INTER_LINEAR|INTER_NEAREST, BORDER_CONSTANT 0); maps in pixels."""
    src = np.ascontiguousarray(src)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    mapx = np.ascontiguousarray(mapx, np.float32)
    mapy = np.ascontiguousarray(mapy, np.float32)
    dh, dw = mapx.shape
    dst = np.empty((dh, dw) + ((cn,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_remap_u8(_p(src), C.c_ssize_t(src.strides[0]), sw, sh, cn, _p(mapx), _p(mapy), C.c_ssize_t(dw),
                       dw, dh, _p(dst), C.c_ssize_t(dst.strides[0]), 1 if linear else 0)
    return dst


def resize_nn(src, dw, dh):
    src = np.ascontiguousarray(src)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    dst = np.empty((dh, dw) + ((cn,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_resize_nn_u8(_p(src), C.c_ssize_t(src.strides[0]), sw, sh, cn, _p(dst), C.c_ssize_t(dst.strides[0]), dw, dh)
    return dst


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src)
    sh, sw = src.shape[:2]
    if src.dtype == np.float32:
        dst = np.empty((dh, dw), np.float32)
        lib().orc_resize_linear_f32(_p(src), C.c_ssize_t(sw), sw, sh, _p(dst), C.c_ssize_t(dw), dw, dh)
        return dst
    cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((dh, dw) + ((cn,) if src.ndim == 3 else ()), np.uint8)
    lib().orc_resize_linear_u8(_p(src), C.c_ssize_t(src.strides[0]), sw, sh, cn, _p(dst), C.c_ssize_t(dst.strides[0]), dw, dh)
    return dst


def resize_rgb(src, dw, dh):
    """cv::resize(src, dsize, 0, 0, INTER_LINEAR) for 8-bit images as Mapper::stitch calls it (mapper.cpp:290-294,308-312):
    equal sizes copy (imgwarp.cpp:3261-3265), exact 2x reductions take the INTER_AREA fast path (imgwarp.cpp:3299-3303,
    ResizeAreaFastVec :2349-2390: (a + b + c + d + 2) >> 2), everything else the 11-bit fixed-point bilinear."""
    src = np.ascontiguousarray(src)
    sh, sw = src.shape[:2]
    if (sw, sh) == (dw, dh):
        return src.copy()
    if sw == 2 * dw and sh == 2 * dh:
        s = src.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    return resize_linear(src, dw, dh)


def dist_l2_3x3(mask):
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    d = np.empty((h, w), np.float32)
    lib().orc_dist_l2_3x3(_p(mask), C.c_ssize_t(w), w, h, _p(d), C.c_ssize_t(w))
    return d


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def gain_feed(imgs, masks, corners):
    """GainCompensator::feed on working-scale RGB images; returns gains (n,) f64."""
    n = len(imgs)
    imgs = [np.ascontiguousarray(i, np.uint8) for i in imgs]
    masks = [np.ascontiguousarray(m, np.uint8) for m in masks]
    steps = (C.c_ssize_t * n)(*[i.strides[0] for i in imgs])
    msteps = (C.c_ssize_t * n)(*[m.strides[0] for m in masks])
    cor = np.ascontiguousarray(np.array(corners, np.int32).reshape(n, 2))
    siz = np.ascontiguousarray(np.array([[i.shape[1], i.shape[0]] for i in imgs], np.int32))
    g = np.zeros(n, np.float64)
    rc = lib().orc_gain_feed(n, _ptr_array(imgs), steps, _ptr_array(masks), msteps, _p(cor), _p(siz), _p(g))
    if rc != 0:
        raise RuntimeError("gain solve failed")
    return g


def mul_scalar(img, g):
    out = np.ascontiguousarray(img, np.uint8).copy()
    h = out.shape[0]
    lib().orc_mul_scalar_u8(_p(out), C.c_ssize_t(out.strides[0]), out.strides[0], h, C.c_double(g))
    return out


def feather_weights(masks, rois, border):
    n = len(masks)
    masks = [np.ascontiguousarray(m, np.uint8) for m in masks]
    r = np.ascontiguousarray(np.array(rois, np.int32).reshape(n, 4))
    ws = [np.empty(m.shape, np.float32) for m in masks]
    lib().orc_feather_weights(n, _ptr_array(masks), _p(r), int(border), _ptr_array(ws))
    return ws


def union_roi(rois):
    r = np.array(rois, np.int64).reshape(-1, 4)
    x0, y0 = r[:, 0].min(), r[:, 1].min()
    x1, y1 = (r[:, 0] + r[:, 2]).max(), (r[:, 1] + r[:, 3]).max()
    return int(x0), int(y0), int(x1 - x0), int(y1 - y0)


def feather_blend(imgs, weights, rois):
    n = len(imgs)
    imgs = [np.ascontiguousarray(i, np.uint8) for i in imgs]
    weights = [np.ascontiguousarray(w, np.float32) for w in weights]
    r = np.ascontiguousarray(np.array(rois, np.int32).reshape(n, 4))
    ox, oy, ow, oh = union_roi(rois)
    out = np.zeros((oh, ow, 3), np.uint8)
    lib().orc_feather_blend(n, _ptr_array(imgs), _ptr_array(weights), _p(r), _p(out), C.c_ssize_t(out.strides[0]), ox, oy, ow, oh)
    return out


def pyrdown_s16(a):
    a = np.ascontiguousarray(a, np.int16)
    h, w = a.shape[:2]
    cn = 1 if a.ndim == 2 else a.shape[2]
    d = np.empty(((h + 1) // 2, (w + 1) // 2) + ((cn,) if a.ndim == 3 else ()), np.int16)
    lib().orc_pyrdown_s16(_p(a), w, h, cn, _p(d))
    return d


def pyrup_s16(a):
    a = np.ascontiguousarray(a, np.int16)
    h, w = a.shape[:2]
    cn = 1 if a.ndim == 2 else a.shape[2]
    d = np.empty((2 * h, 2 * w) + ((cn,) if a.ndim == 3 else ()), np.int16)
    lib().orc_pyrup_s16(_p(a), w, h, cn, _p(d))
    return d


def pyrdown_f32(a):
    a = np.ascontiguousarray(a, np.float32)
    h, w = a.shape
    d = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().orc_pyrdown_f32(_p(a), w, h, _p(d))
    return d


def multiband_blend(imgs, masks, rois, num_bands):
    n = len(imgs)
    imgs = [np.ascontiguousarray(i, np.uint8) for i in imgs]
    masks = [np.ascontiguousarray(m, np.uint8) for m in masks]
    r = np.ascontiguousarray(np.array(rois, np.int32).reshape(n, 4))
    ox, oy, ow, oh = union_roi(rois)
    out = np.zeros((oh, ow, 3), np.uint8)
    omask = np.zeros((oh, ow), np.uint8)
    lib().orc_multiband_blend(n, _ptr_array(imgs), _ptr_array(masks), _p(r), int(num_bands),
                              _p(out), C.c_ssize_t(out.strides[0]), _p(omask), C.c_ssize_t(ow))
    return out, omask


def seam_masks(masks, rois, out_w, out_h):
    n = len(masks)
    masks = [np.ascontiguousarray(m, np.uint8) for m in masks]
    r = np.ascontiguousarray(np.array(rois, np.int32).reshape(n, 4))
    outs = [np.empty(m.shape, np.uint8) for m in masks]
    lib().orc_seam_masks(n, _ptr_array(masks), _p(r), int(out_w), int(out_h), _ptr_array(outs))
    return outs


# ----------------------------------------------------------------------------- cameras
class _Camera(C.Structure):
    _fields_ = [("type", C.c_int), ("rot", C.c_double * 9), ("min_lon", C.c_double), ("max_lon", C.c_double),
                ("p", C.c_double * 16), ("ip", C.c_int * 8),
                ("pol", C.c_double * 64), ("invpol", C.c_double * 64), ("n_pol", C.c_int), ("n_invpol", C.c_int),
                ("dist", C.c_double * 14), ("n_dist", C.c_int),
                ("exclude_mask", C.c_void_p), ("ex_w", C.c_int), ("ex_h", C.c_int),
                ("include_mask", C.c_void_p), ("in_w", C.c_int), ("in_h", C.c_int)]


CAM_TYPES = {"normal": 0, "perspective": 1, "pinhole": 2, "fisheye": 3, "equirectangular": 4,
             "fullframe_fisheye": 5, "ocam_fisheye": 6, "stupidoval": 7, "cubic": 8,
             "eqareanorthpole": 9, "eqareasouthpole": 10}


def fill_poly(mask, pts, val):
    """cv::fillPoly(mask, {pts}, val) in place (orc_fill_poly: drawing.cpp restated; pinned by tests/golden/fillpoly.npz).
    cv2 4.x is NOT equivalent: it differs from the reference on every edge that is clipped by the image."""
    assert mask.dtype == np.uint8 and mask.ndim == 2 and mask.strides[1] == 1
    p = np.ascontiguousarray(np.array(pts, np.int32).reshape(-1))
    lib().orc_fill_poly(_p(mask), C.c_ssize_t(mask.strides[0]), mask.shape[1], mask.shape[0], _p(p), len(p) // 2, int(val))
    return mask


_fill_poly = fill_poly


def read_ocam_file(path):
    """get_ocam_model (cameras/ocam_fisheye.cpp:19-80): comment lines start with '#'; then, in order: n + n direct
    coefficients, m + m inverse coefficients, centre "xc yc", affine "c d e", image "height width"."""
    nums = []
    for ln in open(path):
        ln = ln.strip()
        if ln and not ln.startswith("#"):
            nums += ln.split()
    it = iter(nums)
    n = int(next(it)); pol = [float(next(it)) for _ in range(n)]
    m = int(next(it)); inv = [float(next(it)) for _ in range(m)]
    xc, yc = float(next(it)), float(next(it))
    c, d, e = float(next(it)), float(next(it)), float(next(it))
    h, w = int(next(it)), int(next(it))
    return dict(pol=pol, invpol=inv, xc=xc, yc=yc, c=c, d=d, e=e, width=w, height=h)


def make_camera(typ, opts):
    """camera.cpp:33-135 + per-model constructors.  Returns (struct, keepalive list)."""
    if typ not in CAM_TYPES:
        raise ValueError("Invalid camera type " + typ)
    cam = _Camera()
    keep = []
    cam.type = CAM_TYPES[typ]
    R = (C.c_double * 9)()
    rot = opts.get("rotation", None)
    roll = rot["roll"] if rot else 0.0
    yaw = rot["yaw"] if rot else 0.0
    pitch = rot["pitch"] if rot else 0.0
    lib().orc_rotation_matrix(C.c_double(roll), C.c_double(yaw), C.c_double(pitch), R)
    if "rotation_matrix" in opts:
        for k in range(9):
            R[k] = float(opts["rotation_matrix"][k])
    cam.rot = R
    if "longitude_selection" in opts:
        cam.min_lon, cam.max_lon = float(opts["longitude_selection"][0]), float(opts["longitude_selection"][1])
        assert cam.max_lon > cam.min_lon
    else:
        cam.min_lon, cam.max_lon = -math.pi, math.pi
    ex = inc = None
    if "selection" in opts:
        w, h = int(opts["width"]), int(opts["height"])
        ex = np.full((h, w), 255, np.uint8)
        l, r, t, b = [int(v) for v in opts["selection"]]
        _fill_poly(ex, [(l, t), (l, b - 1), (r - 1, b - 1), (r - 1, t)], 0)
    for key, kind in (("exclude_masks", "exclude"), ("include_masks", "include")):
        if key not in opts:
            continue
        w, h = int(opts["width"]), int(opts["height"])
        if key == "exclude_masks" and ex is None:
            ex = np.zeros((h, w), np.uint8)
        if inc is None:
            inc = np.zeros((h, w), np.uint8)
        for area in opts[key]:
            if area["type"] == "polygonal":
                a = area["args"]
                pts = [(int(a[i]), int(a[i + 1])) for i in range(0, len(a), 2)]
                _fill_poly(inc if kind == "include" else ex, pts, 255)
            elif area["type"] == "png":
                # camera.cpp:169-187: cv::imdecode(args, 1) -> 8UC3 BGR; red channel != 0 -> exclude, green != 0 -> include,
                # whichever list the entry sits in.  (cv2's decoder stands in for the reference's: PNG is lossless.)
                import cv2
                img = cv2.imdecode(np.asarray(area["args"], np.uint8), 1)
                if ex is None or img is None or img.shape[:2] != ex.shape:
                    raise ValueError("png mask: needs an exclude mask of the same size (camera.cpp:175)")
                ex[img[:, :, 2] != 0] = 255
                inc[img[:, :, 1] != 0] = 255
            else:
                raise ValueError("unknown mask type")
    if ex is not None:
        keep.append(ex)
        cam.exclude_mask, cam.ex_w, cam.ex_h = ex.ctypes.data, ex.shape[1], ex.shape[0]
    if inc is not None:
        keep.append(inc)
        cam.include_mask, cam.in_w, cam.in_h = inc.ctypes.data, inc.shape[1], inc.shape[0]

    if typ == "normal":
        cam.p[0], cam.p[1] = float(opts["aspect_ratio"]), float(opts["cam_opt"])
    elif typ == "perspective":
        cam.p[0], cam.p[1] = float(opts["aspect_ratio"]), float(opts["sf"])
    elif typ in ("pinhole", "fisheye"):
        cam.p[0], cam.p[1], cam.p[2], cam.p[3] = [float(opts[k]) for k in ("fx", "fy", "cx", "cy")]
        cam.ip[0], cam.ip[1] = int(opts["width"]), int(opts["height"])
        d = [float(v) for v in opts["dist_coeffs"]]
        cam.n_dist = len(d)
        for i, v in enumerate(d):
            cam.dist[i] = v
    elif typ == "equirectangular":
        cam.p[0] = float(opts.get("min_lat", -math.pi / 2))
        cam.p[1] = float(opts.get("max_lat", math.pi / 2))
        cam.p[2] = float(opts.get("scale_lon", 1.0))
    elif typ == "fullframe_fisheye":
        w, h = int(opts["width"]), int(opts["height"])
        cx = cy = cw = ch = 0
        circ = False
        if "crop" in opts:
            a = [int(v) for v in opts["crop"]["rect"]]
            cx, cy, cw, ch = a[0], a[2], a[1] - a[0], a[3] - a[2]
            circ = bool(opts["crop"]["is_circular"])
        if cw * ch == 0:
            cx, cy, cw, ch, circ = 0, 0, w, h, False
        cam.ip[0], cam.ip[1], cam.ip[2], cam.ip[3], cam.ip[4], cam.ip[5], cam.ip[6] = w, h, cx, cy, cw, ch, int(circ)
        cam.p[0] = float(opts["hfov"])
        cam.p[1], cam.p[2] = float(opts["center_dx"]), float(opts["center_dy"])
        r = [float(v) for v in opts["radial"]]
        rd = (C.c_double * 4)(1.0 - r[0] - r[1] - r[2], r[2], r[1], r[0])
        for k in range(4):
            cam.p[3 + k] = rd[k]
        cam.p[3 + 4] = (cw if cw < ch else ch) / 2.0
        cam.p[3 + 5] = lib().orc_fisheye_correction_radius(rd)
    elif typ == "ocam_fisheye":
        if "file" in opts:                 # cameras/ocam_fisheye.cpp:19-80: Scaramuzza's calib_results.txt
            opts = dict(opts, **read_ocam_file(opts["file"]))
        pol, inv = [float(v) for v in opts["pol"]], [float(v) for v in opts["invpol"]]
        cam.n_pol, cam.n_invpol = len(pol), len(inv)
        for i, v in enumerate(pol):
            cam.pol[i] = v
        for i, v in enumerate(inv):
            cam.invpol[i] = v
        cam.p[0], cam.p[1], cam.p[2], cam.p[3], cam.p[4] = [float(opts[k]) for k in ("xc", "yc", "c", "d", "e")]
        cam.ip[0], cam.ip[1] = int(opts["width"]), int(opts["height"])
    elif typ == "eqareanorthpole":
        cam.p[0] = float(opts.get("arctic_circle", math.pi / 3))
    elif typ == "eqareasouthpole":
        cam.p[0] = float(opts.get("antarctic_circle", -math.pi / 3))
    return cam, keep


def vignette_map(opts, width=512, height=512):
    """vignette.cpp:18-54.  Returns None when the camera has no vignette."""
    if "vignette" not in opts:
        return None
    abcd = np.array([np.float32(v) for v in opts["vignette"]], np.float32)
    if "exposure" in opts:
        ev = np.float32(2.0 ** float(opts["exposure"]))
        abcd = (abcd / ev).astype(np.float32)
    out = np.empty((height, width), np.float32)
    lib().orc_vignette_map(_p(abcd), width, height, _p(out))
    return out


class Template:
    """Mirror of vr::MapperTemplate's data members (octvr.hpp:48-91)."""

    def __init__(self):
        self.out_size = (0, 0)   # (width, height)
        self.inputs = []         # dicts: roi (x,y,w,h), map1, map2, mask, vignette
        self.overlay_inputs = []
        self.seam_masks = []


def build_template(cfg, width, height=-1, use_roi=True, with_seams=True):
    """octvr_dump flow (apps/octvr/dump.cpp:71-127): MapperTemplate ctor + add_input per camera + create_masks()."""
    if isinstance(cfg, str):
        cfg = json.loads(cfg)
    ocam, okeep = make_camera(cfg["output"]["type"], cfg["output"].get("options", {}))
    ar = lib().orc_camera_aspect_ratio(C.byref(ocam))
    if height <= 0 and width <= 0:
        raise ValueError("Output width/height invalid")
    if height <= 0:
        height = int(float(width) / ar)
    if width <= 0:
        width = int(float(height) * ar)
    t = Template()
    t.out_size = (width, height)
    visible = np.zeros((height, width), np.uint8)
    for overlay, key in ((False, "inputs"), (True, "overlays")):
        for inp in cfg.get(key, []):
            icam, ikeep = make_camera(inp["type"], inp["options"])
            map1 = np.empty((height, width), np.float32)
            map2 = np.empty((height, width), np.float32)
            mask = np.empty((height, width), np.uint8)
            roi = (C.c_int * 4)()
            priors = [d["mask"] for d in t.inputs]
            prois = np.ascontiguousarray(np.array([d["roi"] for d in t.inputs], np.int32).reshape(-1, 4))
            rc = lib().orc_template_add_input(C.byref(ocam), C.byref(icam), width, height, _p(map1), _p(map2), _p(mask),
                                              _p(visible), int(use_roi), roi, len(priors),
                                              _ptr_array(priors) if priors else None, _p(prois))
            if rc == -1:
                raise NotImplementedError("camera model not usable in this direction")
            if rc != 0:
                raise RuntimeError("empty input (CV_Assert(min_h <= max_h ...), template.cpp:123)")
            x, y, w, h = [int(v) for v in roi]
            d = dict(roi=(x, y, w, h),
                     map1=np.ascontiguousarray(map1[y:y + h, x:x + w]),
                     map2=np.ascontiguousarray(map2[y:y + h, x:x + w]),
                     mask=np.ascontiguousarray(mask[y:y + h, x:x + w]),
                     vignette=vignette_map(inp["options"]))
            (t.overlay_inputs if overlay else t.inputs).append(d)
    if with_seams:
        t.seam_masks = seam_masks([d["mask"] for d in t.inputs], [d["roi"] for d in t.inputs], width, height)
    return t


def _w64(f, v):
    f.write(struct.pack("<q", int(v)))


def _wmat(f, m):
    if m is None or m.size == 0:
        _w64(f, 0), _w64(f, 0), _w64(f, 0)
        return
    _w64(f, 5 if m.dtype == np.float32 else 0), _w64(f, m.shape[0]), _w64(f, m.shape[1])
    f.write(np.ascontiguousarray(m).tobytes())


def dump_dat(t, path):
    """MapperTemplate::dump (template.cpp:206-256)."""
    with open(path, "wb") as f:
        f.write(b"VRv11")
        _w64(f, t.out_size[0]), _w64(f, t.out_size[1])
        _w64(f, len(t.inputs))
        for d in t.inputs:
            for v in d["roi"]:
                _w64(f, v)
            _wmat(f, d["map1"]), _wmat(f, d["map2"]), _wmat(f, d["mask"]), _wmat(f, d["vignette"])
        for m in t.seam_masks:
            _wmat(f, m)
        _w64(f, len(t.overlay_inputs))
        for d in t.overlay_inputs:
            for v in d["roi"]:
                _w64(f, v)
            _wmat(f, d["map1"]), _wmat(f, d["map2"]), _wmat(f, d["mask"]), _wmat(f, d["vignette"])


def load_dat(path):
    """MapperTemplate(std::ifstream&) (template.cpp:258-314)."""
    with open(path, "rb") as f:
        if f.read(5) != b"VRv11":
            raise ValueError("Invalid data file (version does not match)")
        r64 = lambda: struct.unpack("<q", f.read(8))[0]

        def rmat():
            typ, rows, cols = r64(), r64(), r64()
            if rows * cols == 0:
                return None
            dt = {0: np.uint8, 5: np.float32}[typ & 7]
            return np.frombuffer(f.read(rows * cols * np.dtype(dt).itemsize), dtype=dt).reshape(rows, cols).copy()

        def rinput():
            roi = tuple(r64() for _ in range(4))
            return dict(roi=roi, map1=rmat(), map2=rmat(), mask=rmat(), vignette=rmat())

        t = Template()
        w, h = r64(), r64()
        t.out_size = (w, h)
        t.inputs = [rinput() for _ in range(r64())]
        t.seam_masks = [rmat() for _ in range(len(t.inputs))]
        t.overlay_inputs = [rinput() for _ in range(r64())]
        return t


# ----------------------------------------------------------------------------- per-frame composition
def split_packed(frame, w, h):
    """Mapper's packed layout (mapper.hpp:75-83): (1.5h, w) u8, Y on top, U | V side by side below."""
    y = frame[:h, :w]
    u = frame[h:h + h // 2, :w // 2]
    v = frame[h:h + h // 2, w // 2:w]
    return y, u, v


class StitchOracle:
    """CPU contract for Mapper::Mapper + Mapper::stitch (mapper.cpp:47-323), stage by stage per SURVEY 8(c)."""

    def __init__(self, tmpl, in_sizes, blend=128, enable_gain=True, scale_output=(0, 0), texel_center=False):
        """in_sizes: blended inputs, then overlay inputs (mapper.cpp:84-127); scale_output: (0, 0) keeps the template size.
        texel_center: sample at map * W - 0.5, where the reference's CUDA Mapper samples (tex2D, normalised coordinates, linear
        filter: fast_remap.cu) instead of cv::remap's map * W -- the product's OCTVR_TEXEL_CENTER=1."""
        self.t = tmpl
        self.in_sizes = [tuple(s) for s in in_sizes]
        n = len(tmpl.inputs)
        self.overlays = list(getattr(tmpl, "overlay_inputs", []))
        assert len(self.in_sizes) == n + len(self.overlays)
        ov_sizes = self.in_sizes[n:]
        self.ov_mapx = [scale_map(d["map1"], s[0]) for d, s in zip(self.overlays, ov_sizes)]
        self.ov_mapy = [scale_map(d["map2"], s[1]) for d, s in zip(self.overlays, ov_sizes)]
        self.ov_vig = [resize_linear(d["vignette"], s[0], s[1]) if d.get("vignette") is not None else None
                       for d, s in zip(self.overlays, ov_sizes)]
        self.scale_output = tuple(scale_output) if scale_output[0] > 0 and scale_output[1] > 0 else tuple(tmpl.out_size)
        self.last_preview = None
        if n == 1:
            enable_gain, blend = False, 0
        self.blend, self.enable_gain, self.n = blend, enable_gain, n
        self.rois = [d["roi"] for d in tmpl.inputs]
        W, H = tmpl.out_size
        self.mapx = [scale_map(d["map1"], s[0]) for d, s in zip(tmpl.inputs, self.in_sizes)]
        self.mapy = [scale_map(d["map2"], s[1]) for d, s in zip(tmpl.inputs, self.in_sizes)]
        self.vig = [resize_linear(d["vignette"], s[0], s[1]) if d["vignette"] is not None else None
                    for d, s in zip(tmpl.inputs, self.in_sizes)]
        if texel_center:
            half = np.float32(0.5)
            self.mapx, self.mapy = [m - half for m in self.mapx], [m - half for m in self.mapy]
            self.ov_mapx, self.ov_mapy = [m - half for m in self.ov_mapx], [m - half for m in self.ov_mapy]
        # mapper.cpp:94-99,113-114
        self.working_scale = min(1.0, math.sqrt(0.1 * 1e6 / (W * H)))
        ws = self.working_scale
        self.srois = [(int(r[0] * ws), int(r[1] * ws), int(r[2] * ws), int(r[3] * ws)) for r in self.rois]
        if enable_gain:
            self.smasks = [resize_linear(d["mask"], sr[2], sr[3]) for d, sr in zip(tmpl.inputs, self.srois)]
        if blend > 0:
            self.bands = int(math.ceil(math.log(blend) / math.log(2.0)) - 1.0)
        elif blend < 0:
            self.weights = feather_weights([d["mask"] for d in tmpl.inputs], self.rois, -blend)
        self.last_gains = None

    def warp(self, frames_rgb):
        """frames_rgb: list of (h,w,3) RGB -> per-ROI remapped RGB."""
        out = []
        for i in range(self.n):
            src = frames_rgb[i]
            if self.vig[i] is not None:
                # cudaarithm mul_mat.cu:198-213: saturate_cast<uchar>(u8 * f32) (round to nearest even)
                src = np.clip(np.rint(src.astype(np.float32) * self.vig[i][:, :, None]), 0, 255).astype(np.uint8)
            out.append(remap(src, self.mapx[i], self.mapy[i], True))
        return out

    def gains_from(self, warped):
        imgs = [resize_nn(w, sr[2], sr[3]) for w, sr in zip(warped, self.srois)]
        return gain_feed(imgs, self.smasks, [(sr[0], sr[1]) for sr in self.srois])

    def blend_rgb(self, warped):
        W, H = self.t.out_size
        result = np.zeros((H, W, 3), np.uint8)
        ox, oy, ow, oh = union_roi(self.rois)
        if self.blend < 0:
            result[oy:oy + oh, ox:ox + ow] = feather_blend(warped, self.weights, self.rois)
        elif self.blend > 0:
            res, _ = multiband_blend(warped, self.t.seam_masks, self.rois, self.bands)
            result[oy:oy + oh, ox:ox + ow] = res
        else:
            for img, d in zip(warped, self.t.inputs):   # mapper.cpp:269-275: masked copy in camera order
                x, y, w, h = d["roi"]
                m = d["mask"] != 0
                result[y:y + h, x:x + w][m] = img[m]
        return result

    def stitch_rgb(self, frames_rgb, gains=None):
        """Mapper::result: blended inputs, then the overlays copied over it through their masks (mapper.cpp:279-282; the
        evident intent -- the reference copies from a buffer it never fills)."""
        warped = self.warp(frames_rgb[:self.n])
        if self.enable_gain:
            g = self.gains_from(warped) if gains is None else np.asarray(gains, np.float64)
            self.last_gains = g
            warped = [mul_scalar(w, float(gi)) for w, gi in zip(warped, g)]
        result = self.blend_rgb(warped)
        for k, d in enumerate(self.overlays):
            src = frames_rgb[self.n + k]
            if self.ov_vig[k] is not None:
                src = np.clip(np.rint(src.astype(np.float32) * self.ov_vig[k][:, :, None]), 0, 255).astype(np.uint8)
            w = remap(src, self.ov_mapx[k], self.ov_mapy[k], True)
            x, y, rw, rh = d["roi"]
            m = d["mask"] != 0
            result[y:y + rh, x:x + rw][m] = w[m]
        return result

    def stitch(self, frames_yuv, gains=None, preview_size=None):
        """frames_yuv: list of (y,u,v) plane triples.  Returns (y,u,v) of the stitched frame at scale_output size
        (mapper.cpp:290-306); preview_size=(w, h) also leaves the preview (mapper.cpp:308-312) in self.last_preview."""
        rgb = [yuv420_to_rgb(*f) for f in frames_yuv]
        result = self.stitch_rgb(rgb, gains)
        self.last_result = result
        if preview_size is not None:
            self.last_preview = resize_rgb(result, preview_size[0], preview_size[1])
        if self.scale_output != tuple(self.t.out_size):
            result = resize_rgb(result, self.scale_output[0], self.scale_output[1])
        return rgb_to_yuv420(result)


# ---------------------------------------------------------------------------------------------------------------------
# vr::FastMapper (modules/octvr/src/mapper_fast.cpp): NV12 in, NV12-shaped out, u8 feather weights, no RGB round trip.
# ---------------------------------------------------------------------------------------------------------------------
def convert_maps_16sc2(mapx, mapy):
    """cv::convertMaps(mapx, mapy, CV_16SC2, nninterpolate = false) (imgwarp.cpp:4900-4960): ix = cvRound(x * 32),
    map1 = saturate_cast<short>(ix >> 5), map2 = (iy & 31) * 32 + (ix & 31)."""
    ix = np.rint(np.asarray(mapx, np.float32) * np.float32(32)).astype(np.int64)
    iy = np.rint(np.asarray(mapy, np.float32) * np.float32(32)).astype(np.int64)
    ix = np.clip(ix, -2 ** 31, 2 ** 31 - 1)
    iy = np.clip(iy, -2 ** 31, 2 ** 31 - 1)
    m1 = np.stack([np.clip(ix >> 5, -32768, 32767), np.clip(iy >> 5, -32768, 32767)], -1).astype(np.int16)
    m2 = ((iy & 31) * 32 + (ix & 31)).astype(np.uint16)
    return m1, m2


def resize_half(src):
    """cv::resize(src, Size(cols / 2, rows / 2)) (INTER_LINEAR) as the FastMapper constructor calls it on maps, masks and
    feather masks (mapper_fast.cpp:57-58,70,98-99): an exact 2x reduction goes through the INTER_AREA fast path
    (imgwarp.cpp:3299-3303) -- u8: (a + b + c + d + 2) >> 2 (:2349-2390); f32: ((a + b) + (c + d)) * 0.25f in the SSE body
    (:2283-2318, dx <= w - 4) and (((a + b) + c) + d) * 0.25f in the scalar tail (:2425-2437) -- anything else through the
    bilinear resize."""
    src = np.ascontiguousarray(src)
    sh, sw = src.shape
    dw, dh = sw // 2, sh // 2
    if sw != 2 * dw or sh != 2 * dh:
        return resize_linear(src, dw, dh)
    a, b, c, d = src[0::2, 0::2], src[0::2, 1::2], src[1::2, 0::2], src[1::2, 1::2]
    if src.dtype == np.uint8:
        return ((a.astype(np.int32) + b + c + d + 2) >> 2).astype(np.uint8)
    out = ((a + b) + (c + d)) * np.float32(0.25)
    body = dw & ~3
    out[:, body:] = ((((np.float32(0) + a[:, body:]) + b[:, body:]) + c[:, body:]) + d[:, body:]) * np.float32(0.25)
    return out.astype(np.float32)


def remap_weighted(src, acc, map1, map2, weight):
    """cv::remap_weighted's OpenCL kernel (imgproc/src/opencl/remap_weighted.cl:20-77) for SRC_T = uchar, DST_T = ushort,
    WT = float: bilinear taps outside the source read 0, the 5-bit fractions weight them in float (exact), the product with
    the u8 weight is rounded once, converted with convert_ushort_sat_rte and ADDED to acc (16-bit wrap)."""
    sh, sw = src.shape
    ax = map1[..., 0].astype(np.int64)
    ay = map1[..., 1].astype(np.int64)
    m2 = map2.astype(np.int64) & 1023
    ux = ((m2 & 31).astype(np.float32)) / np.float32(32)
    uy = ((m2 >> 5).astype(np.float32)) / np.float32(32)

    def pix(gx, gy):
        ok = (gx >= 0) & (gy >= 0) & (gx < sw) & (gy < sh)
        v = src[np.clip(gy, 0, sh - 1), np.clip(gx, 0, sw - 1)].astype(np.float32)
        return np.where(ok, v, np.float32(0))
    one = np.float32(1)
    a, b, c, d = pix(ax, ay), pix(ax + 1, ay), pix(ax, ay + 1), pix(ax + 1, ay + 1)
    v = a * (one - ux) * (one - uy) + b * ux * (one - uy) + c * (one - ux) * uy + d * ux * uy
    p = (v * weight.astype(np.float32)).astype(np.float32)
    q = np.clip(np.rint(p), 0, 65535).astype(np.uint16)
    acc += q                                      # uint16: wraps like the kernel's `dst[0] += ...`


class FastMapperOracle:
    """vr::FastMapper: constructor mapper_fast.cpp:27-109, stitch_nv12 :153-195."""

    def __init__(self, tmpl, in_sizes):
        assert not getattr(tmpl, "overlay_inputs", [])                      # CV_Assert(mt.overlay_inputs.size() == 0)
        W, H = tmpl.out_size
        self.out_size, self.in_sizes = (W, H), [tuple(s) for s in in_sizes]
        self.map1, self.map2, self.hmap1, self.hmap2 = [], [], [], []
        for d, (iw, ih) in zip(tmpl.inputs, self.in_sizes):
            assert tuple(d["roi"]) == (0, 0, W, H), "FastMapper does not support ROI (mapper_fast.cpp:50-51)"
            m1, m2 = convert_maps_16sc2(scale_map(d["map1"], iw), scale_map(d["map2"], ih))
            self.map1.append(m1), self.map2.append(m2)
            h1, h2 = convert_maps_16sc2(scale_map(resize_half(d["map1"]), iw // 2), scale_map(resize_half(d["map2"]), ih // 2))
            self.hmap1.append(h1), self.hmap2.append(h2)
        # feather masks (mapper_fast.cpp:76-101): max(DT - 5, 0) / (1e-5 + sum), x 255 rounded to u8, and their half-size copies
        total = np.full((H, W), 1e-5, np.float32)
        ws = []
        for d in tmpl.inputs:
            w = dist_l2_3x3(d["mask"]) - np.float32(5)
            w = np.where(w > 0, w, np.float32(0)).astype(np.float32)
            total = (w + total).astype(np.float32)
            ws.append(w)
        self.feather = [np.clip(np.rint((w / total).astype(np.float32) * np.float32(255.0)), 0, 255).astype(np.uint8) for w in ws]
        self.hfeather = [resize_half(f) for f in self.feather]

    def stitch_nv12(self, frames):
        """frames: (1.5 h, w) u8 arrays, luma rows then interleaved chroma rows.  Returns the (1.5 H, W) output frame.
        As in the reference, output chroma channel 0 is remapped from input chroma channel 1 and vice versa (:179-180)."""
        W, H = self.out_size
        acc0 = np.zeros((H, W), np.uint16)
        acc1 = [np.zeros((H // 2, W // 2), np.uint16), np.zeros((H // 2, W // 2), np.uint16)]
        for i, (f, (iw, ih)) in enumerate(zip(frames, self.in_sizes)):
            assert f.shape == (ih + ih // 2, iw) and f.dtype == np.uint8
            c0 = f[:ih]
            c12 = f[ih:].reshape(ih // 2, iw // 2, 2)
            remap_weighted(c0, acc0, self.map1[i], self.map2[i], self.feather[i])
            remap_weighted(c12[..., 1], acc1[0], self.hmap1[i], self.hmap2[i], self.hfeather[i])
            remap_weighted(c12[..., 0], acc1[1], self.hmap1[i], self.hmap2[i], self.hfeather[i])
        self.last_acc0 = acc0
        k = np.float32(1.0 / 255.0)                                            # convertTo(CV_8U, 1.0 / 255.0): float scale, cvRound
        out = np.empty((H + H // 2, W), np.uint8)
        out[:H] = np.clip(np.rint(acc0.astype(np.float32) * k), 0, 255).astype(np.uint8)
        uv = np.stack(acc1, -1).astype(np.float32) * k
        out[H:] = np.clip(np.rint(uv), 0, 255).astype(np.uint8).reshape(H // 2, W)
        return out


def fast_noise_frame(cam, w, h):
    """The NV12-shaped noise frame of oracle/refgen/ref_fast.cpp."""
    o = np.arange((h + h // 2) * w, dtype=np.uint64)
    x = (np.uint64(0xFA57) ^ (np.uint64(cam) << np.uint64(32)) ^ o) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    return (x & np.uint64(0xFF)).astype(np.uint8).reshape(h + h // 2, w)
