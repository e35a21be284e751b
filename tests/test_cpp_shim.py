"""include/octvr.hpp (header-only C++ shim with the reference's class names) compiles against the C ABI and
behaves like the reference: loads "VRv11" files, throws std::string on a bad magic; on a GPU the
AsyncMultiMapper push/pop path matches the Python path bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    import octvr_b200 as vr
    vr.lib()
    exe = str(tmp_path_factory.mktemp("cpp") / "stitch_demo")
    libdir = os.path.join(ROOT, "opencv-octvr_b200")
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "stitch_demo.cpp"), "-o", exe, "-L" + libdir, "-loctvr_b200",
                           "-L" + cuda_lib, "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib])
    return exe


def test_shim_loads_dat_and_throws_string_on_bad_magic(demo, tmp_path):
    ot = util.template_from_gold(O, "rig3")
    p = str(tmp_path / "t.dat")
    O.dump_dat(ot, p)
    out = subprocess.run([demo, "--dat", p], capture_output=True, text=True)
    assert out.returncode == 0 and "out 256x128 inputs 3" in out.stdout and out.stdout.count("seam 1") == 3
    bad = str(tmp_path / "bad.dat")
    open(bad, "wb").write(b"VRv10" + b"\0" * 40)
    out = subprocess.run([demo, "--dat", bad], capture_output=True, text=True)
    assert out.returncode == 10 and "version does not match" in out.stderr      # std::string, like template.cpp:262


@pytest.mark.gpu
def test_shim_async_stitch_matches_oracle(demo, tmp_path):
    cfg = os.path.join(util.GOLD, "rigs", "rig2s.json")
    outp = str(tmp_path / "o.i420")
    r = subprocess.run([demo, "--config", cfg, "128", "192", "108", "-3", outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(outp, np.uint8)
    ot = O.build_template(util.rig_json("rig2s"), 128)
    so = O.StitchOracle(ot, [(192, 108)] * 2, blend=-3, enable_gain=True)
    frames = [util.i420_planes(util.noise_frame(c, 192, 108, seed=1234 + 2), 192, 108) for c in range(2)]
    y, u, v = so.stitch(frames)
    assert np.array_equal(got, np.concatenate([y.ravel(), u.ravel(), v.ravel()]))


@pytest.mark.gpu
def test_shim_incremental_template_writes_the_reference_tools_bytes(demo, tmp_path):
    """MapperTemplate(to, to_opts, w, h) + add_input(...) per camera + create_masks() + dump(std::ofstream&) -- the call
    sequence of the reference's octvr_dump (apps/octvr/dump.cpp:98-127) -- yields the very file the reference tool wrote."""
    import hashlib
    import json
    cfg = util.rig_json("rig3")
    args = [demo, "--incremental", cfg["output"]["type"], json.dumps(cfg["output"].get("options", {})), str(util.rig_width("rig3")), str(tmp_path / "t.dat")]
    for inp in cfg["inputs"]:
        args += [inp["type"], json.dumps(inp["options"])]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "inputs 3 seams 3" in r.stdout
    ref = json.load(open(os.path.join(util.GOLD, "dat_sha256.json")))["rig3"]
    data = open(tmp_path / "t.dat", "rb").read()
    assert len(data) == ref["bytes"] and hashlib.sha256(data).hexdigest() == ref["sha256"]


@pytest.mark.gpu
def test_shim_fast_mapper_matches_oracle(demo, tmp_path):
    """vr::FastMapper(mt, in_sizes) + stitch_nv12 through the C++ shim == the FastMapper oracle."""
    cfg = os.path.join(util.GOLD, "rigs", "rig3.json")
    outp = str(tmp_path / "o.nv12")
    r = subprocess.run([demo, "--fast", cfg, "256", "320", "240", outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(outp, np.uint8).reshape(192, 256)
    assert np.array_equal(got, np.load(os.path.join(util.GOLD, "fast_rig3.npz"))["result"])
