"""
GPU parity of the stages after the blender (Mapper::stitch, mapper.cpp:279-312): overlay inputs, scale_output, preview.
Golden vectors tests/golden/post_*.npz come from the reference's own cv::remap / cv::resize / cv::cvtColor
(oracle/refgen/ref_stitch.cpp run against the unmodified reference build).  Bit-exact.
"""
import glob
import os

import numpy as np
import pytest

import octvr_b200 as vr
import oracle as O
import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

POST = sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(util.GOLD, "post_*.npz")))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(params=["ring", "staged", "fused", "direct"])
def blend_path(request, monkeypatch):
    monkeypatch.setenv("OCTVR_BLEND", request.param)
    return request.param


def _planes(flat, w, h):
    q = (w // 2) * (h // 2)
    return flat[:w * h].view(h, w), flat[w * h:w * h + q].view(h // 2, w // 2), flat[w * h + q:w * h + 2 * q].view(h // 2, w // 2)


def _vr_template(t):
    for d in t.inputs:
        d["vignette"] = None
    return vr.MapperTemplate.from_arrays(t.out_size, t.inputs, t.seam_masks, overlays=t.overlay_inputs)


@pytest.mark.parametrize("case", POST)
def test_overlay_scale_preview_vs_reference_golden(case, blend_path):
    g = np.load(os.path.join(util.GOLD, "post_%s.npz" % case))
    iw, ih, blend, gain, _, sw, sh, pw, ph = [int(v) for v in g["meta"]]
    t = util.template_from_gold(O, case.split("_")[0])
    vt = _vr_template(t)
    assert vt.num_overlays == 1
    n = len(t.inputs) + len(t.overlay_inputs)
    m = vr.Mapper(vt, [(iw, ih)] * n, blend=blend, enable_gain_compensator=bool(gain), scale_output=(sw, sh))
    m.set_keep_rgb(True)
    W, H = m.out_size
    assert (W, H) == ((sw, sh) if sw else t.out_size)
    ins = [_planes(dev(g["frame%d" % i]).view(-1), iw, ih) for i in range(n)]
    other = [_planes(dev(util.noise_frame(i, iw, ih, seed=77)).view(-1), iw, ih) for i in range(n)]
    out = torch.zeros(W * H * 3 // 2, dtype=torch.uint8, device="cuda")
    pv = torch.zeros((ph, pw, 3), dtype=torch.uint8, device="cuda") if pw else None
    m.stitch(other, _planes(out, W, H), preview=pv)      # a different frame first: nothing of it may survive
    m.stitch(ins, _planes(out, W, H), preview=pv)
    torch.cuda.synchronize()
    assert np.array_equal(m.result_rgb(), g["result_rgb"])
    assert np.array_equal(out.cpu().numpy(), g["result_yuv"].ravel())
    if pw:
        assert np.array_equal(pv.cpu().numpy(), g["preview_rgb"])
    if gain:
        assert np.allclose(m.gains(), g["gains"].ravel(), rtol=1e-9)


@pytest.mark.parametrize("scale,preview", [((96, 48), (40, 20)), ((256, 128), (256, 128)), ((300, 200), (128, 64)), ((128, 64), (31, 17))])
def test_scale_output_and_preview_sizes_vs_oracle(scale, preview):
    """Down, equal (copy), up and exact-2x sizes against the oracle's cv::resize restatement; no overlays; packed layout."""
    t = util.template_from_gold(O, "rig3")
    vt = _vr_template(t)
    n = len(t.inputs)
    iw, ih = 320, 240
    frames = [util.noise_frame(i, iw, ih) for i in range(n)]
    so = O.StitchOracle(t, [(iw, ih)] * n, blend=-3, enable_gain=True, scale_output=scale)
    ry, ru, rv = so.stitch([util.i420_planes(f, iw, ih) for f in frames], preview_size=preview)
    m = vr.Mapper(vt, [(iw, ih)] * n, blend=-3, enable_gain_compensator=True, scale_output=scale)
    W, H = m.out_size
    ins = [_planes(dev(f).view(-1), iw, ih) for f in frames]
    out = torch.zeros(W * H * 3 // 2, dtype=torch.uint8, device="cuda")
    pv = torch.zeros((preview[1], preview[0], 3), dtype=torch.uint8, device="cuda")
    m.stitch(ins, _planes(out, W, H), preview=pv)
    torch.cuda.synchronize()
    y, u, v = [p.cpu().numpy() for p in _planes(out, W, H)]
    assert np.array_equal(y, ry) and np.array_equal(u, ru) and np.array_equal(v, rv)
    assert np.array_equal(pv.cpu().numpy(), so.last_preview)


def test_overlay_with_json_template_and_wrong_input_count():
    """Overlay tables from the product's own map generation (octvr_template_build_json, "overlays" key)."""
    cfg = util.rig_json("rig3ov")
    ot = O.build_template(cfg, 256)
    vt = vr.MapperTemplate.from_json(cfg, 256)
    assert vt.num_inputs == 3 and vt.num_overlays == 1
    iw, ih = 320, 240
    frames = [util.noise_frame(i, iw, ih) for i in range(4)]
    so = O.StitchOracle(ot, [(iw, ih)] * 4, blend=-2, enable_gain=True)
    ry, ru, rv = so.stitch([util.i420_planes(f, iw, ih) for f in frames])
    m = vr.Mapper(vt, [(iw, ih)] * 4, blend=-2, enable_gain_compensator=True)
    W, H = m.out_size
    out = torch.zeros(W * H * 3 // 2, dtype=torch.uint8, device="cuda")
    m.stitch([_planes(dev(f).view(-1), iw, ih) for f in frames], _planes(out, W, H))
    torch.cuda.synchronize()
    y, u, v = [p.cpu().numpy() for p in _planes(out, W, H)]
    assert np.array_equal(y, ry) and np.array_equal(u, ru) and np.array_equal(v, rv)
    with pytest.raises(vr.OctvrError):
        m.stitch([_planes(dev(f).view(-1), iw, ih) for f in frames[:3]], _planes(out, W, H))     # the overlay frame is missing
    with pytest.raises(vr.OctvrError):
        vr.Mapper(vt, [(iw, ih)] * 3, blend=-2)


def test_async_preview_and_scaled_regions():
    """AsyncMultiMapper with a preview size and two output regions of a frame smaller than the templates: every region
    is resized to its share of the output and of the preview frame (async.cpp:76-84,247-259)."""
    t = util.template_from_gold(O, "rig3ov")
    vt = _vr_template(t)
    n = 4
    iw, ih = 320, 240
    OW, OH, PW, PH = 200, 180, 64, 48
    regions = [(0.0, 0.0, 1.0, 0.5), (0.0, 0.5, 1.0, 0.5)]
    am = vr.AsyncMultiMapper([vt, vt], [(iw, ih)] * n, (OW, OH), [-2, 0], [0, -1], regions, (PW, PH))
    frames = [util.noise_frame(i, iw, ih) for i in range(n)]
    hin = [util.i420_planes(f, iw, ih) for f in frames]
    hout = np.zeros(OW * OH * 3 // 2, np.uint8)
    am.push(hin, util.i420_planes(hout, OW, OH))
    am.pop()
    pv = am.preview()
    am.close()
    y, u, v = util.i420_planes(hout, OW, OH)
    for r, (blend, gain) in enumerate([(-2, True), (0, False)]):
        so = O.StitchOracle(t, [(iw, ih)] * n, blend=blend, enable_gain=gain, scale_output=(OW, OH // 2))
        ry, ru, rv = so.stitch(hin, preview_size=(PW, PH // 2))
        assert np.array_equal(y[r * OH // 2:(r + 1) * OH // 2], ry)
        assert np.array_equal(u[r * OH // 4:(r + 1) * OH // 4], ru) and np.array_equal(v[r * OH // 4:(r + 1) * OH // 4], rv)
        assert np.array_equal(pv[r * PH // 2:(r + 1) * PH // 2], so.last_preview)
