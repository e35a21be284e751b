"""
The CPU oracle (oracle/) pinned against golden vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, generator oracle/refgen/make_golden.py).  Integer/byte work is bit-exact.
"""
import glob
import json
import os

import numpy as np
import pytest

import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLD, "primitives.npz"))


@pytest.mark.parametrize("cn", [1, 3])
def test_remap_linear_and_nearest_bit_exact(prim, cn):
    p = "remap_c%d_" % cn
    src, mx, my = prim[p + "src"], prim[p + "mapx"], prim[p + "mapy"]
    assert np.array_equal(O.remap(src, mx, my, True), prim[p + "linear"])
    assert np.array_equal(O.remap(src, mx, my, False), prim[p + "nearest"])


def test_bilinear_table_closed_form():
    t = O.bilinear_table().astype(np.int64)
    fy, fx = np.mgrid[0:32, 0:32]
    cf = np.stack([(32 - fy) * (32 - fx), (32 - fy) * fx, fy * (32 - fx), fy * fx], -1) * 32
    # the one cell whose unit weight does not fit a short: {32767,0,0,1} in the reference's table,
    # pixel-equivalent to {32768,0,0,0} (see oracle/orc_image.c)
    assert list(t[0, 0]) == [32767, 0, 0, 1]
    t[0, 0] = cf[0, 0]
    assert np.array_equal(t, cf)
    a, d = np.mgrid[0:256, 0:256]
    assert np.array_equal((32767 * a + d + 16384) >> 15, a)


def test_map_scaling(prim):
    assert np.array_equal(O.scale_map(prim["scale_in"], 2704), prim["scale_2704"])
    assert np.array_equal(O.scale_map(prim["scale_in"], 1520), prim["scale_1520"])


def test_colour_bit_exact(prim):
    yuv = prim["color_yuv"]
    h = yuv.shape[0] * 2 // 3
    w = yuv.shape[1]
    y = yuv[:h]
    u = yuv[h:h + h // 4].reshape(h // 2, w // 2)
    v = yuv[h + h // 4:].reshape(h // 2, w // 2)
    rgb = O.yuv420_to_rgb(y, u, v)
    assert np.array_equal(rgb, prim["color_rgb_i420"])
    assert np.array_equal(rgb[:, :, ::-1], prim["color_bgr_i420"])
    uv = yuv[h:].reshape(h // 2, w // 2, 2)
    assert np.array_equal(O.yuv420_to_rgb(y, uv[:, :, 0], uv[:, :, 1]), prim["color_rgb_nv12"])
    yy, uu, vv = O.rgb_to_yuv420(prim["color_rgb_in"])
    back = prim["color_yuv_out"]
    assert np.array_equal(yy, back[:h])
    assert np.array_equal(uu.ravel(), back[h:h + h // 4].ravel())
    assert np.array_equal(vv.ravel(), back[h + h // 4:].ravel())


def test_distance_transform_bit_exact(prim):
    assert np.array_equal(O.dist_l2_3x3(prim["dt_mask"]), prim["dt_dist"])
    assert np.array_equal(O.dist_l2_3x3(np.full((40, 30), 255, np.uint8)), prim["dt_full_dist"])


def test_resize_bit_exact(prim):
    assert np.array_equal(O.resize_nn(prim["resize_nn_src"], 22, 13), prim["resize_nn_dst"])
    dn = O.resize_linear(prim["resize_lin_src"], 24, 12)
    assert np.array_equal(dn, prim["resize_lin_down"])
    assert np.array_equal(O.resize_linear(dn, 223, 117), prim["resize_lin_up"])
    assert np.array_equal(O.resize_linear(prim["resize_lin_noise_src"], 37, 29), prim["resize_lin_noise_dst"])
    assert np.array_equal(O.resize_linear(prim["resize_f32_src"], 75, 41), prim["resize_f32_dst"])


def test_pyramids_bit_exact(prim):
    dn = O.pyrdown_s16(prim["pyr_s16_src"])
    assert np.array_equal(dn, prim["pyr_s16_down"])
    assert np.array_equal(O.pyrup_s16(dn), prim["pyr_s16_up"])
    assert np.array_equal(O.pyrdown_f32(prim["pyr_f32_src"]), prim["pyr_f32_down"])
    assert np.array_equal(O.pyrdown_f32(prim["pyr_f32b_src"]), prim["pyr_f32b_down"])


@pytest.mark.parametrize("n", [2, 3, 4])
def test_gain_compensator(prim, n):
    imgs = [prim["gain%d_img%d" % (n, i)] for i in range(n)]
    masks = [prim["gain%d_mask%d" % (n, i)] for i in range(n)]
    g = O.gain_feed(imgs, masks, prim["gain%d_corners" % n])
    ref = prim["gain%d_gains" % n].ravel()
    assert np.allclose(g, ref, rtol=1e-12, atol=0)
    assert np.array_equal(O.mul_scalar(imgs[0], ref[0]), prim["gain%d_applied0" % n])


def test_multiply_scalar_bit_exact(prim):
    ramp = np.arange(256, dtype=np.uint8)[None, :]
    for k, g in enumerate(prim["mul_gains"].ravel()):
        assert np.array_equal(O.mul_scalar(ramp, float(g))[0], prim["mul_out"][k])


def test_feather_weights_and_narrowing_bit_exact(prim):
    rois = prim["feather_rois"]
    masks = [prim["feather_mask%d" % i] for i in range(3)]
    ws = O.feather_weights(masks, rois, 3)
    for i in range(3):
        assert np.array_equal(ws[i], prim["feather_w%d" % i])
    acc = prim["feather_acc"]
    alpha = np.float32(1.0 / 3)
    ours = np.clip(np.rint(acc.astype(np.float32) * alpha), 0, 255).astype(np.uint8)
    assert np.array_equal(ours, prim["feather_acc_u8"])


@pytest.mark.parametrize("variant", [0, 1])
def test_multiband_blender_bit_exact(prim, variant):
    p = "mb%d_" % variant
    rois = prim[p + "rois"]
    imgs = [prim[p + "img%d" % i] for i in range(3)]
    masks = [prim[p + "mask%d" % i] for i in range(3)]
    out, omask = O.multiband_blend(imgs, masks, rois, int(prim[p + "bands"][0, 0]))
    assert np.array_equal(omask, prim[p + "result_mask"])
    assert np.array_equal(out, prim[p + "result8"])


RIGS = sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(GOLD, "tmpl_*.npz")))


@pytest.mark.parametrize("rig", RIGS)
def test_template_tables_bit_exact(rig):
    """MapperTemplate + add_input + create_masks for all 11 camera models, in and out."""
    g = np.load(os.path.join(GOLD, "tmpl_%s.npz" % rig))
    import util
    cfg = util.rig_json(rig)
    w = json.load(open(os.path.join(GOLD, "rigs", "widths.json")))[rig]
    t = O.build_template(cfg, w)
    assert tuple(g["out_size"]) == t.out_size
    assert int(g["n"]) == len(t.inputs)
    for i, d in enumerate(t.inputs):
        assert tuple(g["roi%d" % i]) == d["roi"], (rig, i)
        assert np.array_equal(g["mask%d" % i], d["mask"]), (rig, i)
        W, H = t.out_size
        # contract: 1e-4 source px on fl32(map)*W; observed: bit-identical f32
        assert np.array_equal(g["map1_%d" % i], d["map1"]), (rig, i, np.abs(g["map1_%d" % i] - d["map1"]).max())
        assert np.array_equal(g["map2_%d" % i], d["map2"]), (rig, i)
        if "vig%d" % i in g:
            assert np.array_equal(g["vig%d" % i], d["vignette"])
        else:
            assert d["vignette"] is None
        assert np.array_equal(g["seam%d" % i], t.seam_masks[i]), (rig, i)


def test_template_empty_input_is_an_error():
    cfg = json.load(open(os.path.join(GOLD, "rigs", "out_perspective.json")))
    cfg["inputs"][1]["options"]["rotation"]["yaw"] = 0.3 + np.pi   # behind the perspective output
    with pytest.raises(RuntimeError):
        O.build_template(cfg, 160)


def test_dat_roundtrip(tmp_path):
    cfg = json.load(open(os.path.join(GOLD, "rigs", "rig2s.json")))
    t = O.build_template(cfg, 128)
    p = str(tmp_path / "t.dat")
    O.dump_dat(t, p)
    t2 = O.load_dat(p)
    assert t2.out_size == t.out_size
    for a, b in zip(t.inputs, t2.inputs):
        assert a["roi"] == tuple(b["roi"])
        assert np.array_equal(a["map1"], b["map1"]) and np.array_equal(a["mask"], b["mask"])
    with open(p, "r+b") as f:
        f.write(b"VRv10")
    with pytest.raises(ValueError):
        O.load_dat(p)


STITCH = sorted(os.path.basename(f)[7:-4] for f in glob.glob(os.path.join(GOLD, "stitch_*.npz")))


def _template_from_gold(rig):
    g = np.load(os.path.join(GOLD, "tmpl_%s.npz" % rig))
    t = O.Template()
    t.out_size = tuple(int(v) for v in g["out_size"])
    for i in range(int(g["n"])):
        t.inputs.append(dict(roi=tuple(int(v) for v in g["roi%d" % i]), map1=g["map1_%d" % i], map2=g["map2_%d" % i],
                             mask=g["mask%d" % i], vignette=None))
        t.seam_masks.append(g["seam%d" % i])
    return t


@pytest.mark.parametrize("case", STITCH)
def test_stitch_composition_bit_exact(case):
    """Whole-frame CPU contract (cvtColor -> remap -> gain -> blend -> cvtColor) vs the reference run."""
    g = np.load(os.path.join(GOLD, "stitch_%s.npz" % case))
    iw, ih, blend, gain, _ = [int(v) for v in g["meta"]]
    rig = case.split("_")[0]
    t = _template_from_gold(rig)
    n = len(t.inputs)
    so = O.StitchOracle(t, [(iw, ih)] * n, blend=blend, enable_gain=bool(gain))
    frames = []
    for i in range(n):
        f = g["frame%d" % i]
        frames.append((f[:ih], f[ih:ih + ih // 4].reshape(ih // 2, iw // 2), f[ih + ih // 4:].reshape(ih // 2, iw // 2)))
    y, u, v = so.stitch(frames)
    if gain:
        assert np.allclose(so.last_gains, g["gains"].ravel(), rtol=1e-12)
    H = t.out_size[1]
    ref = g["result_yuv"]
    assert np.array_equal(y, ref[:H])
    assert np.array_equal(u.ravel(), ref[H:H + H // 4].ravel())
    assert np.array_equal(v.ravel(), ref[H + H // 4:].ravel())


POST = sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(GOLD, "post_*.npz")))


@pytest.mark.parametrize("case", POST)
def test_after_blend_stages_bit_exact(case):
    """Overlay inputs, scale_output and preview (mapper.cpp:279-312) vs the reference's own cv::remap / cv::resize /
    cv::cvtColor run on the same template and frames (oracle/refgen/ref_stitch.cpp)."""
    import util
    g = np.load(os.path.join(GOLD, "post_%s.npz" % case))
    iw, ih, blend, gain, _, sw, sh, pw, ph = [int(v) for v in g["meta"]]
    t = util.template_from_gold(O, case.split("_")[0])
    for d in t.inputs:
        d["vignette"] = None
    n = len(t.inputs) + len(t.overlay_inputs)
    assert len(t.overlay_inputs) == 1
    so = O.StitchOracle(t, [(iw, ih)] * n, blend=blend, enable_gain=bool(gain), scale_output=(sw, sh))
    frames = [util.i420_planes(g["frame%d" % i], iw, ih) for i in range(n)]
    y, u, v = so.stitch(frames, preview_size=(pw, ph) if pw else None)
    assert np.array_equal(so.last_result, g["result_rgb"])
    W, H = (sw, sh) if sh else t.out_size
    ry, ru, rv = util.i420_planes(g["result_yuv"], W, H)
    assert np.array_equal(y, ry) and np.array_equal(u, ru) and np.array_equal(v, rv)
    if pw:
        assert np.array_equal(so.last_preview, g["preview_rgb"])


def test_fill_poly_bit_exact():
    """cv::fillPoly of the reference build (310 polygons incl. clipped, concave, degenerate; oracle/refgen/ref_fillpoly.cpp)."""
    g = np.load(os.path.join(GOLD, "fillpoly.npz"))
    for k in range(int(g["n"])):
        p = g["p%d" % k]
        w, h = int(p[0]), int(p[1])
        want = np.unpackbits(g["m%d" % k])[:w * h].reshape(h, w).astype(bool)
        got = O.fill_poly(np.full((h, w), 7, np.uint8), p[2:].reshape(-1, 2), 200)
        assert np.array_equal(got == 200, want), (k, p.tolist())


@pytest.mark.parametrize("rig", ["rig3", "models"])
def test_fast_mapper_oracle_vs_reference_fixture(rig):
    """vr::FastMapper: the oracle's tables == the tables the reference's own constructor built, and its frame == the frame
    the OpenCL kernel's arithmetic gives on them (oracle/refgen/ref_fast.cpp); second rig: odd output height, so the
    half-size tables go through the generic bilinear resize instead of the 2x area path."""
    g = np.load(os.path.join(GOLD, "fast_%s.npz" % rig))
    t = O.Template()
    t.out_size = tuple(int(v) for v in g["out_size"])
    W, H = t.out_size
    n = int(g["n"])
    for i in range(n):
        t.inputs.append(dict(roi=(0, 0, W, H), map1=g["t_map1_%d" % i], map2=g["t_map2_%d" % i], mask=g["t_mask%d" % i], vignette=None))
    iw, ih = (int(v) for v in g["in_size"])
    fo = O.FastMapperOracle(t, [(iw, ih)] * n)
    for i in range(n):
        for key, tab in (("map1_", fo.map1), ("map2_", fo.map2), ("hmap1_", fo.hmap1), ("hmap2_", fo.hmap2), ("feather", fo.feather), ("hfeather", fo.hfeather)):
            assert np.array_equal(g[key + str(i)], tab[i]), (rig, key, i)
    with np.errstate(over="ignore"):
        out = fo.stitch_nv12([O.fast_noise_frame(i, iw, ih) for i in range(n)])
    assert np.array_equal(fo.last_acc0, g["acc_c0"])
    assert np.array_equal(out, g["result"])


@pytest.mark.parametrize("blend,gain", [(-5, True), (16, True), (0, False)])
def test_reference_cpu_arm_equals_the_oracle(blend, gain, tmp_path):
    """oracle/_ref/libocvref.so (the unmodified reference CPU build behind oracle/refgen/ref_arm.cpp; bench.py's reference arm)
    composes the same frame as the oracle, byte for byte, gains to 1e-12.  Skipped where the reference build is absent."""
    import refarm
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import util
    if not refarm.available():
        pytest.skip("oracle/_ref/libocvref.so not built (needs the reference CPU build, oracle/build_ref.sh)")
    ot = util.template_from_gold(O, "rig3")
    for d in ot.inputs:
        d["vignette"] = None
    dat = str(tmp_path / "t.dat")
    O.dump_dat(ot, dat)
    arm = refarm.RefArm(dat, (320, 240), blend, gain)
    frames = [util.noise_frame(i, 320, 240) for i in range(3)]
    (y, u, v), g = arm.stitch(frames)
    so = O.StitchOracle(ot, [(320, 240)] * 3, blend=blend, enable_gain=gain)
    ry, ru, rv = so.stitch([util.i420_planes(f, 320, 240) for f in frames])
    assert np.array_equal(y, ry) and np.array_equal(u, ru) and np.array_equal(v, rv)
    if gain:
        assert np.allclose(g, so.last_gains, rtol=1e-12)
