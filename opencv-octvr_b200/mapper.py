"""
Python mirror of the octvr mapper API over the C ABI (used by tests and bench.py).
Names follow the reference (modules/octvr/include/octvr.hpp, src/mapper.hpp):
MapperTemplate, Mapper.stitch / gains, AsyncMultiMapper.push / pop.  Device memory is held in
torch tensors; every compute call goes through liboctvr_b200.so.
"""
import ctypes as C
import json

import numpy as np

from . import capi
from .capi import Frame, OctvrError, check, lib


def _np_ptr_array(arrs, ctype=C.c_void_p):
    return (ctype * len(arrs))(*[(a.ctypes.data if a is not None else None) for a in arrs])


class MapperTemplate:
    """vr::MapperTemplate (octvr.hpp:48-91)."""

    def __init__(self, handle):
        self._h = handle

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def load(cls, path):
        """explicit MapperTemplate(std::ifstream&) -- template.cpp:258-314"""
        h = C.c_void_p()
        check(lib().octvr_template_load_file(path.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def from_bytes(cls, data):
        h = C.c_void_p()
        check(lib().octvr_template_load_dat(data, C.c_size_t(len(data)), C.byref(h)))
        return cls(h)

    @classmethod
    def create(cls, to, to_opts, width, height=-1, device=0):
        """MapperTemplate(to, to_opts, width, height) (octvr.hpp:72-74); add inputs with add_input(), then create_masks()."""
        h = C.c_void_p()
        check(lib().octvr_template_create(to.encode(), json.dumps(to_opts or {}).encode(), int(width), int(height), int(device), C.byref(h)))
        return cls(h)

    def add_input(self, frm, from_opts, overlay=False, use_roi=True):
        """add_input(from, from_opts, overlay, use_roi) (octvr.hpp:75-78): the projection runs as a CUDA kernel."""
        check(lib().octvr_template_add_input(self._h, frm.encode(), json.dumps(from_opts or {}).encode(), int(overlay), int(use_roi)))

    @classmethod
    def from_json(cls, cfg, width, height=-1, use_roi=True, with_seam_masks=True, device=0):
        """MapperTemplate(to, to_opts, w, h) + add_input per camera (+ create_masks), maps generated on the GPU."""
        if not isinstance(cfg, str):
            cfg = json.dumps(cfg)
        h = C.c_void_p()
        check(lib().octvr_template_build_json(cfg.encode(), int(width), int(height), int(use_roi), int(with_seam_masks),
                                              int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def from_arrays(cls, out_size, inputs, seam_masks=None, overlays=()):
        """inputs / overlays: dicts with roi (x,y,w,h), map1, map2, mask[, vignette] -- the fields of MapperTemplate::Input
        (octvr.hpp:55-63: `inputs` and `overlay_inputs`)."""
        n = len(inputs)
        rois = np.ascontiguousarray(np.array([d["roi"] for d in inputs], np.int32).reshape(n, 4))
        m1 = [np.ascontiguousarray(d["map1"], np.float32) for d in inputs]
        m2 = [np.ascontiguousarray(d["map2"], np.float32) for d in inputs]
        mk = [np.ascontiguousarray(d["mask"], np.uint8) for d in inputs]
        sm = [np.ascontiguousarray(s, np.uint8) for s in seam_masks] if seam_masks else None
        vg = [np.ascontiguousarray(d["vignette"], np.float32) if d.get("vignette") is not None else None for d in inputs]
        vw = vh = 0
        for v in vg:
            if v is not None:
                vh, vw = v.shape
        h = C.c_void_p()
        check(lib().octvr_template_from_arrays(int(out_size[0]), int(out_size[1]), n, rois.ctypes.data_as(C.c_void_p),
                                               _np_ptr_array(m1), _np_ptr_array(m2), _np_ptr_array(mk),
                                               _np_ptr_array(sm) if sm else None,
                                               _np_ptr_array(vg) if any(v is not None for v in vg) else None,
                                               vw, vh, C.byref(h)))
        t = cls(h)
        for d in overlays:
            roi = (C.c_int * 4)(*[int(v) for v in d["roi"]])
            a1, a2, ak = (np.ascontiguousarray(d["map1"], np.float32), np.ascontiguousarray(d["map2"], np.float32),
                          np.ascontiguousarray(d["mask"], np.uint8))
            v = np.ascontiguousarray(d["vignette"], np.float32) if d.get("vignette") is not None else None
            check(lib().octvr_template_add_overlay(t._h, roi, a1.ctypes.data_as(C.c_void_p), a2.ctypes.data_as(C.c_void_p),
                                                   ak.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p) if v is not None else None,
                                                   v.shape[1] if v is not None else 0, v.shape[0] if v is not None else 0))
        return t

    # -- accessors -------------------------------------------------------------------------------
    @property
    def out_size(self):
        w, h = C.c_int(), C.c_int()
        check(lib().octvr_template_out_size(self._h, C.byref(w), C.byref(h)))
        return w.value, h.value

    @property
    def num_inputs(self):
        return lib().octvr_template_num_inputs(self._h)

    @property
    def num_overlays(self):
        return lib().octvr_template_num_overlays(self._h)

    def input(self, i):
        roi = (C.c_int * 4)()
        m1, m2 = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        mk, sm = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint8)()
        vg = C.POINTER(C.c_float)()
        vwh = (C.c_int * 2)()
        check(lib().octvr_template_input(self._h, i, roi, C.byref(m1), C.byref(m2), C.byref(mk), C.byref(sm), C.byref(vg), vwh))
        x, y, w, h = list(roi)
        d = dict(roi=(x, y, w, h),
                 map1=np.ctypeslib.as_array(m1, shape=(h, w)).copy(),
                 map2=np.ctypeslib.as_array(m2, shape=(h, w)).copy(),
                 mask=np.ctypeslib.as_array(mk, shape=(h, w)).copy(),
                 seam_mask=np.ctypeslib.as_array(sm, shape=(h, w)).copy() if sm else None,
                 vignette=np.ctypeslib.as_array(vg, shape=(vwh[1], vwh[0])).copy() if vg else None)
        return d

    @property
    def inputs(self):
        return [self.input(i) for i in range(self.num_inputs)]

    def create_masks(self):
        check(lib().octvr_template_create_masks(self._h))

    def dump(self, path):
        check(lib().octvr_template_dump_file(self._h, path.encode()))

    def __del__(self):
        try:
            if self._h:
                lib().octvr_template_destroy(self._h)
                self._h = None
        except Exception:
            pass


def frame_from_planes(y, u, v):
    """octvr_frame from three 2-D u8 torch tensors / numpy arrays (u, v may be strided views: NV12 = stride 2)."""
    def ptr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

    def strides(a):
        return tuple(a.stride()) if hasattr(a, "stride") else tuple(s // a.itemsize for s in a.strides)

    f = Frame()
    f.y, f.u, f.v = ptr(y), ptr(u), ptr(v)
    f.y_pitch, f.u_pitch, f.v_pitch = strides(y)[0], strides(u)[0], strides(v)[0]
    assert strides(y)[1] == 1 and strides(u)[1] == strides(v)[1] and strides(u)[1] in (1, 2)
    f.uv_pixel_stride = strides(u)[1]
    return f


def split_packed(t, w, h):
    """Mapper's packed layout (mapper.hpp:75-83): (1.5h, w), U and V side by side under Y."""
    return t[:h, :w], t[h:h + h // 2, :w // 2], t[h:h + h // 2, w // 2:w]


class PackedRGB:
    """An input frame as packed 8UC3: a (h, w, 3) u8 CUDA tensor in R,G,B (bgr=False) or B,G,R order (cv::Mat CV_8UC3 as
    cv::imread / cv::VideoCapture deliver it; include/octvr_b200.h OCTVR_FMT_RGB24 / OCTVR_FMT_BGR24).  Accepted by
    Mapper.stitch in place of a (y, u, v) triple."""

    def __init__(self, tensor, bgr=False):
        assert tensor.dim() == 3 and tensor.shape[2] == 3 and tensor.stride(2) == 1 and tensor.stride(1) == 3
        self.tensor, self.bgr = tensor, bgr

    def frame(self):
        f = Frame()
        f.y = self.tensor.data_ptr()
        f.u = f.v = None
        f.y_pitch, f.u_pitch, f.v_pitch = self.tensor.stride(0), 0, 0
        f.uv_pixel_stride = 4 if self.bgr else 3
        return f


class Mapper:
    """vr::Mapper (mapper.hpp:29-95).  blend > 0 multiband, < 0 feather border, 0 none."""

    def __init__(self, tmpl, in_sizes, blend=128, enable_gain_compensator=True, scale_output=(0, 0), device=0, band=None, cols=None):
        """band=(y0, y1): row-band mapper (multi-GPU partition of one frame, sharding.RowBandStitcher): only output rows
        [y0, y1) are produced; frames passed to stitch() stay full size.  cols=(x0, x1): the same for output columns
        (multiband only)."""
        import torch
        self._torch = torch
        self.tmpl = tmpl
        self.device = device
        self.band = band
        self.in_sizes = [tuple(s) for s in in_sizes]
        sz = np.ascontiguousarray(np.array(self.in_sizes, np.int32).reshape(-1, 2))
        h = C.c_void_p()
        self.cols = cols
        if cols is not None:
            y0, y1 = band if band is not None else (0, 0)
            check(lib().octvr_mapper_create_window(tmpl._h, sz.ctypes.data_as(C.c_void_p), len(self.in_sizes), int(blend),
                                                   int(bool(enable_gain_compensator)), int(cols[0]), int(cols[1]), int(y0), int(y1),
                                                   int(device), C.byref(h)))
        elif band is None:
            check(lib().octvr_mapper_create(tmpl._h, sz.ctypes.data_as(C.c_void_p), len(self.in_sizes), int(blend),
                                            int(bool(enable_gain_compensator)), int(scale_output[0]), int(scale_output[1]),
                                            int(device), C.byref(h)))
        else:
            check(lib().octvr_mapper_create_band(tmpl._h, sz.ctypes.data_as(C.c_void_p), len(self.in_sizes), int(blend),
                                                 int(bool(enable_gain_compensator)), int(band[0]), int(band[1]),
                                                 int(device), C.byref(h)))
        self._h = h
        so = tuple(int(v) for v in scale_output)
        self.out_size = so if so[0] > 0 and so[1] > 0 else tmpl.out_size      # scaled_output_size, mapper.cpp:68
        self.stitch_size = tmpl.out_size

    def stitch(self, inputs, output, gains=None, stream=None, preview=None):
        """inputs: list of (y,u,v) CUDA u8 tensors (blended inputs, then overlays); output: (y,u,v) CUDA u8 tensors (written
        in place); preview: optional (ph, pw, 3) CUDA u8 tensor, the result resized into it (mapper.cpp:308-312).
        Asynchronous on `stream` (default: torch's current stream)."""
        torch = self._torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        fin = (Frame * len(inputs))(*[p.frame() if isinstance(p, PackedRGB) else frame_from_planes(*p) for p in inputs])
        fout = frame_from_planes(*output) if output is not None else None
        g = None
        ng = 0
        if gains is not None:
            ng = len(gains)
            g = (C.c_double * ng)(*[float(v) for v in gains])
        pv, pp, pw, ph = None, 0, 0, 0
        if preview is not None:
            assert preview.dim() == 3 and preview.shape[2] == 3 and preview.stride(2) == 1 and preview.stride(1) == 3
            pv, pp, pw, ph = C.c_void_p(preview.data_ptr()), preview.stride(0), preview.shape[1], preview.shape[0]
        check(lib().octvr_mapper_stitch(self._h, fin, len(inputs), C.byref(fout) if fout is not None else None, pv, C.c_size_t(pp), pw, ph,
                                        g, ng, C.c_void_p(s.cuda_stream)))

    def stitch_packed(self, inputs, output, gains=None, stream=None):
        """Mapper::stitch's own layout: each input a (1.5h, w) CUDA u8 tensor, output likewise."""
        torch = self._torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        ptrs = (C.c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
        pit = (C.c_size_t * len(inputs))(*[t.stride(0) for t in inputs])
        g = None
        ng = 0
        if gains is not None:
            ng = len(gains)
            g = (C.c_double * ng)(*[float(v) for v in gains])
        check(lib().octvr_mapper_stitch_packed(self._h, ptrs, pit, len(inputs), C.c_void_p(output.data_ptr()),
                                               C.c_size_t(output.stride(0)), g, ng, C.c_void_p(s.cuda_stream)))

    def set_keep_rgb(self, on=True):
        check(lib().octvr_mapper_set_keep_rgb(self._h, int(on)))

    def result_rgb(self):
        w, h = self.stitch_size
        out = np.empty((h, w, 3), np.uint8)
        check(lib().octvr_mapper_result_rgb(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(w * 3)))
        return out

    def gains(self):
        n = self.tmpl.num_inputs             # overlay inputs are not gain-compensated (mapper.cpp:255-265)
        g = (C.c_double * n)()
        check(lib().octvr_mapper_gains(self._h, g, n))
        return np.array(list(g))

    def stats(self):
        pairs, roi, tb = C.c_int64(), C.c_int64(), C.c_int64()
        ln = C.c_int()
        check(lib().octvr_mapper_stats(self._h, C.byref(pairs), C.byref(roi), C.byref(tb), C.byref(ln)))
        return dict(pairs=pairs.value, roi_area=roi.value, table_bytes=tb.value, launches_per_stitch=ln.value)

    def set_input_window(self, cam, col0, width):
        """The frames of input `cam` passed to stitch() hold source columns [col0, col0 + width) only (a (1.5h, width) packed
        frame or planes whose first column is col0)."""
        check(lib().octvr_mapper_set_input_window(self._h, int(cam), int(col0), int(width)))

    def src_cols(self):
        """per blended input: the source columns (lo, hi) some table entry reads (only those are converted)."""
        n = self.tmpl.num_inputs
        a = (C.c_int * (2 * n))()
        check(lib().octvr_mapper_source_cols(self._h, a, n))
        return [(a[2 * i], a[2 * i + 1]) for i in range(n)]

    def src_rows(self):
        """per blended input: the source rows (lo, hi) this mapper converts and reads (all rows unless it is a row-band mapper)."""
        n = self.tmpl.num_inputs
        a = (C.c_int * (2 * n))()
        check(lib().octvr_mapper_source_rows(self._h, a, n))
        return [(a[2 * i], a[2 * i + 1]) for i in range(n)]

    def set_profiling(self, on=True):
        check(lib().octvr_mapper_set_profiling(self._h, int(on)))

    def stage_ms(self, stage):
        ms = C.c_float()
        check(lib().octvr_mapper_stage_ms(self._h, stage.encode(), C.byref(ms)))
        return ms.value

    def debug_gain_ns(self):
        """%globaltimer stamps of the gain kernel's last CTA (diagnostics)."""
        a = (C.c_ulonglong * 6)()
        check(lib().octvr_mapper_debug_gain_ns(self._h, a))
        return list(a)

    def debug_gain_trace(self, n=8 + 2 * 4096):
        a = (C.c_ulonglong * n)()
        check(lib().octvr_mapper_debug_gain_trace(self._h, a, n))
        return list(a)

    def debug_ring(self):
        """counters of K_blend_ring's TMA ring (diagnostics; needs a -DRING_DEBUG=1 build)."""
        a = (C.c_ulonglong * 8)()
        check(lib().octvr_mapper_debug_ring(self._h, a))
        return list(a)

    def __del__(self):
        try:
            if self._h:
                lib().octvr_mapper_destroy(self._h)
                self._h = None
        except Exception:
            pass


class AsyncMultiMapper:
    """vr::AsyncMultiMapper (octvr.hpp:103-121): host planes in, host planes out, N-deep pipeline."""

    def __init__(self, templates, in_sizes, out_size, blend_modes, gain_modes, output_regions, preview_size=(0, 0), device=0):
        n_out = len(templates)
        th = (C.c_void_p * n_out)(*[t._h for t in templates])
        sz = np.ascontiguousarray(np.array(in_sizes, np.int32).reshape(-1, 2))
        bm = (C.c_int * n_out)(*[int(b) for b in blend_modes])
        gm = (C.c_int * n_out)(*[int(g) for g in gain_modes])
        rg = (C.c_double * (4 * n_out))(*[float(v) for r in output_regions for v in r])
        h = C.c_void_p()
        check(lib().octvr_async_create(th, n_out, sz.ctypes.data_as(C.c_void_p), len(in_sizes), int(out_size[0]), int(out_size[1]),
                                       bm, gm, rg, int(preview_size[0]), int(preview_size[1]), int(device), C.byref(h)))
        self._h = h
        self._keep = []
        self.templates = templates

    @staticmethod
    def New(*a, **k):
        return AsyncMultiMapper(*a, **k)

    def push(self, inputs, output):
        """inputs: list of (y,u,v) host u8 numpy arrays; output: (y,u,v) host arrays to be filled by the matching pop."""
        fin = (Frame * len(inputs))(*[p.frame() if isinstance(p, PackedRGB) else frame_from_planes(*p) for p in inputs])
        fout = frame_from_planes(*output)
        check(lib().octvr_async_push(self._h, fin, len(inputs), C.byref(fout)))
        self._keep.append((inputs, output))           # only an accepted push owns a pop: keep the planes alive until then

    def pop(self):
        check(lib().octvr_async_pop(self._h))
        if self._keep:
            self._keep.pop(0)

    def fps(self):
        v = C.c_double()
        check(lib().octvr_async_fps(self._h, C.byref(v)))
        return v.value

    def preview(self):
        """(ph, pw, 3) RGB preview of the frame popped last (copy)."""
        p = C.POINTER(C.c_uint8)()
        pitch, w, h = C.c_size_t(), C.c_int(), C.c_int()
        check(lib().octvr_async_preview(self._h, C.byref(p), C.byref(pitch), C.byref(w), C.byref(h)))
        return np.ctypeslib.as_array(p, shape=(h.value, pitch.value))[:, :w.value * 3].reshape(h.value, w.value, 3).copy()

    def close(self):
        if self._h:
            lib().octvr_async_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FastMapper:
    """vr::FastMapper (octvr.hpp:123-144, mapper_fast.cpp): NV12 frames in, NV12-shaped frame out, one CUDA launch."""

    def __init__(self, tmpl, in_sizes, device=0):
        import torch
        self._torch = torch
        self.tmpl, self.device = tmpl, device
        self.in_sizes = [tuple(s) for s in in_sizes]
        sz = np.ascontiguousarray(np.array(self.in_sizes, np.int32).reshape(-1, 2))
        h = C.c_void_p()
        check(lib().octvr_fast_create(tmpl._h, sz.ctypes.data_as(C.c_void_p), len(self.in_sizes), int(device), C.byref(h)))
        self._h = h
        self.out_size = tmpl.out_size

    def stitch_nv12(self, inputs, output, stream=None):
        """inputs: (in_h + in_h // 2, in_w) u8 CUDA tensors (luma rows, then interleaved chroma rows); output:
        (H + H // 2, W) u8 CUDA tensor, written in place.  Asynchronous on `stream` (default: torch's current stream)."""
        torch = self._torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        W, H = self.out_size
        for t, (iw, ih) in zip(inputs, self.in_sizes):        # CV_Assert(inputs[i].rows == h + h / 2 && cols == w && type == CV_8U)
            if tuple(t.shape) != (ih + ih // 2, iw) or t.dtype != torch.uint8 or t.stride(1) != 1:
                raise OctvrError(capi.ERR_INVALID, "stitch_nv12: input must be a (h + h / 2, w) u8 frame")
        if tuple(output.shape) != (H + H // 2, W) or output.dtype != torch.uint8 or output.stride(1) != 1:
            raise OctvrError(capi.ERR_INVALID, "stitch_nv12: output must be a (H + H / 2, W) u8 frame")
        ptrs = (C.c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
        pitches = (C.c_size_t * len(inputs))(*[t.stride(0) for t in inputs])
        check(lib().octvr_fast_stitch_nv12(self._h, ptrs, pitches, len(inputs), C.c_void_p(output.data_ptr()),
                                           C.c_size_t(output.stride(0)), C.c_void_p(s.cuda_stream)))

    def info(self):
        w, h = C.c_int(), C.c_int()
        pl, pc, tb = C.c_longlong(), C.c_longlong(), C.c_longlong()
        check(lib().octvr_fast_info(self._h, C.byref(w), C.byref(h), C.byref(pl), C.byref(pc), C.byref(tb)))
        return dict(out_size=(w.value, h.value), pairs_luma=pl.value, pairs_chroma=pc.value, table_bytes=tb.value)

    def table(self, cam, name):
        """Host copy of one of the constructor's tables (tests): map1, half_map1, map2, half_map2, feather, half_feather."""
        which = ["map1", "half_map1", "map2", "half_map2", "feather", "half_feather"].index(name)
        W, H = self.out_size
        w, h = (W // 2, H // 2) if which % 2 else (W, H)
        a = np.empty((h, w, 2), np.int16) if which < 2 else np.empty((h, w), np.uint16 if which < 4 else np.uint8)
        check(lib().octvr_fast_debug_table(self._h, int(cam), which, a.ctypes.data_as(C.c_void_p)))
        return a

    def __del__(self):
        try:
            if self._h:
                lib().octvr_fast_destroy(self._h)
                self._h = None
        except Exception:
            pass
