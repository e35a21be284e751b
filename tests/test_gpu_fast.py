"""
vr::FastMapper (mapper_fast.cpp, remap_weighted.cl) -- csrc/fast.cu through the C ABI against (a) the fixtures written by
the reference's own FastMapper constructor + the OpenCL kernel's arithmetic (oracle/refgen/ref_fast.cpp) and (b) the CPU
oracle (oracle.FastMapperOracle, pinned to the same fixtures by tests/test_oracle_golden.py).  Bit-exact: tolerance 0.
"""
import os

import numpy as np
import pytest

import octvr_b200 as vr
import oracle as O
import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _template(g):
    W, H = (int(v) for v in g["out_size"])
    n = int(g["n"])
    inputs = [dict(roi=(0, 0, W, H), map1=g["t_map1_%d" % i], map2=g["t_map2_%d" % i], mask=g["t_mask%d" % i], vignette=None) for i in range(n)]
    return (W, H), inputs


@pytest.mark.parametrize("mode", ["staged", "direct"])
@pytest.mark.parametrize("rig", ["rig3", "models"])
def test_fast_mapper_matches_reference_fixture(rig, mode, monkeypatch):
    """Two device paths behind octvr_fast_stitch_nv12: k_fast_staged (source footprints copied into shared memory with cp.async;
    default for 16-byte aligned frames -- rig3) and the direct-gather k_fast_nv12 (OCTVR_FAST=direct, unaligned frames -- the
    200-pixel-wide frames of the `models` case -- or footprints larger than the stage)."""
    monkeypatch.setenv("OCTVR_FAST", mode)
    g = np.load(os.path.join(util.GOLD, "fast_%s.npz" % rig))
    (W, H), inputs = _template(g)
    iw, ih = (int(v) for v in g["in_size"])
    n = len(inputs)
    t = vr.MapperTemplate.from_arrays((W, H), inputs)
    fm = vr.FastMapper(t, [(iw, ih)] * n)
    for i in range(n):      # the constructor's tables == the reference constructor's
        for name, key in (("map1", "map1_"), ("map2", "map2_"), ("half_map1", "hmap1_"), ("half_map2", "hmap2_"), ("feather", "feather"), ("half_feather", "hfeather")):
            assert np.array_equal(fm.table(i, name), g[key + str(i)]), (rig, i, name)
    frames = [torch.from_numpy(O.fast_noise_frame(i, iw, ih)).cuda() for i in range(n)]
    out = torch.full((H + H // 2, W), 7, dtype=torch.uint8, device="cuda")
    fm.stitch_nv12(frames, out)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), g["result"])


@pytest.mark.parametrize("name,width,in_size", [("rig2", 1000, (1920, 1080)), ("rig6", 1536, (2704, 1520))])
def test_fast_mapper_vs_oracle_larger(name, width, in_size):
    """BASELINE rigs (real fisheye geometry, partial coverage, taps on the source border) at a size the numpy oracle handles
    in seconds, on frames with a pitch larger than the width."""
    cfg, _, _ = util.named_rig(name)
    ot = O.build_template(cfg, width, use_roi=False, with_seams=False)
    n = len(ot.inputs)
    fo = O.FastMapperOracle(ot, [in_size] * n)
    iw, ih = in_size
    host = [O.fast_noise_frame(i, iw, ih) for i in range(n)]
    with np.errstate(over="ignore"):
        ref = fo.stitch_nv12(host)
    t = vr.MapperTemplate.from_arrays(ot.out_size, ot.inputs)
    fm = vr.FastMapper(t, [in_size] * n)
    W, H = ot.out_size
    pad = [torch.zeros((ih + ih // 2, iw + 64), dtype=torch.uint8, device="cuda") for _ in range(n)]
    frames = []
    for p, h in zip(pad, host):
        p[:, :iw] = torch.from_numpy(h).cuda()
        frames.append(p[:, :iw])
    out = torch.zeros((H + H // 2, W), dtype=torch.uint8, device="cuda")
    fm.stitch_nv12(frames, out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got[:H], ref[:H]), "luma"
    assert np.array_equal(got[H:], ref[H:]), "chroma"
    info = fm.info()
    assert info["pairs_luma"] > 0 and info["pairs_chroma"] > 0


def test_fast_mapper_error_behaviour():
    tm = vr.MapperTemplate.from_json(util.rig_json("models"), 192)   # ROI-cropped inputs: "does not support ROI yet"
    with pytest.raises(vr.OctvrError) as e:
        vr.FastMapper(tm, [(320, 240)] * tm.num_inputs)
    assert e.value.code == vr.capi.ERR_UNSUPPORTED
    cfg = util.rig_json("rig3")
    t = vr.MapperTemplate.from_json(cfg, 256, use_roi=False)
    with pytest.raises(vr.OctvrError):
        vr.FastMapper(t, [(320, 240)] * 2)                        # input count
    fm = vr.FastMapper(t, [(320, 240)] * 3)
    W, H = t.out_size
    good = [torch.zeros((360, 320), dtype=torch.uint8, device="cuda") for _ in range(3)]
    with pytest.raises(vr.OctvrError):                            # CV_Assert(inputs[i].rows == h + h / 2), mapper_fast.cpp:156-160
        fm.stitch_nv12([torch.zeros((240, 320), dtype=torch.uint8, device="cuda")] * 3, torch.zeros((H + H // 2, W), dtype=torch.uint8, device="cuda"))
    with pytest.raises(vr.OctvrError):
        fm.stitch_nv12(good, torch.zeros((H, W), dtype=torch.uint8, device="cuda"))
    tov = util.template_from_gold(O, "rig3ov")                    # overlay inputs: CV_Assert(mt.overlay_inputs.size() == 0)
    tt = vr.MapperTemplate.from_arrays(tov.out_size, tov.inputs, tov.seam_masks, overlays=tov.overlay_inputs)
    with pytest.raises(vr.OctvrError):
        vr.FastMapper(tt, [(320, 240)] * (len(tov.inputs) + len(tov.overlay_inputs)))
