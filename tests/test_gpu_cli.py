"""CLI parity: tools/octvr_dump.py writes the same bytes as the reference's octvr_dump
(apps/octvr/dump.cpp) for the same config -- digests of the reference tool's files are in tests/golden/dat_sha256.json
(oracle/refgen/make_golden.py) -- and tools/octvr_map.py stitches still images through such a file."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

import octvr_b200 as vr
import oracle as O
import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
TOOLS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(TOOLS, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("rig", ["rig3", "rig3ov", "rig2s", "models", "masks"])
def test_octvr_dump_writes_the_reference_tools_bytes(rig, tmp_path):
    want = json.load(open(os.path.join(util.GOLD, "dat_sha256.json")))[rig]
    out = str(tmp_path / (rig + ".dat"))
    rc = _tool("octvr_dump").main(["-w", str(util.rig_width(rig)), "-o", out, os.path.join(util.GOLD, "rigs", rig + ".json")])
    assert rc == 0
    data = open(out, "rb").read()
    assert len(data) == want["bytes"]
    assert hashlib.sha256(data).hexdigest() == want["sha256"]


def test_octvr_map_one_shot(tmp_path):
    cv2 = pytest.importorskip("cv2")
    dat = str(tmp_path / "rig3.dat")
    assert _tool("octvr_dump").main(["-w", "256", "-o", dat, os.path.join(util.GOLD, "rigs", "rig3.json")]) == 0
    names = []
    rng = np.random.default_rng(5)
    for i in range(3):
        p = str(tmp_path / ("in%d.png" % i))
        cv2.imwrite(p, rng.integers(0, 256, (240, 320, 3), dtype=np.uint8))
        names.append(p)
    outp = str(tmp_path / "pano.png")
    assert _tool("octvr_map").main(["-b", "-3", "-g", dat, outp] + names) == 0
    got = cv2.imread(outp, 1)
    # the same through the oracle: BGR -> I420 (cv2, as the tool does) -> CPU contract -> BGR
    t = O.load_dat(dat)
    frames = []
    for p in names:
        yuv = cv2.cvtColor(cv2.imread(p, 1), cv2.COLOR_BGR2YUV_I420)
        frames.append(util.i420_planes(yuv, 320, 240))
    y, u, v = O.StitchOracle(t, [(320, 240)] * 3, blend=-3, enable_gain=True).stitch(frames)
    want = cv2.cvtColor(np.concatenate([y.ravel(), u.ravel(), v.ravel()]).reshape(128 * 3 // 2, 256), cv2.COLOR_YUV2BGR_I420)
    assert got.shape == (128, 256, 3) and np.array_equal(got, want)
