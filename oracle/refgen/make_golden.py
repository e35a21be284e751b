#!/usr/bin/env python
"""
oracle/refgen/make_golden.py -- regenerate tests/golden/*.npz from the UNMODIFIED reference.

Needs the reference CPU build of SURVEY.md Appendix A (no-CUDA static libs + octvr_dump):
    REFBUILD (default /tmp/refbuild) = cmake build dir, REFSRC (default /tmp/refsrc) = the source copy
    it was configured from (identical to /root/reference except the nine CMake policy edits).
Run here (build container); the fixtures it writes are committed, so neither the GPU box nor the
test-suite ever needs /root/reference.

    python oracle/refgen/make_golden.py
"""
import json
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GOLD = os.path.join(ROOT, "tests", "golden")
B = os.environ.get("REFBUILD", "/tmp/refbuild")
S = os.environ.get("REFSRC", "/tmp/refsrc")
TMP = os.environ.get("GOLDTMP", "/tmp/gold")
sys.path.insert(0, os.path.join(ROOT, "oracle"))

MODS = "core imgproc stitching features2d flann calib3d imgcodecs videoio highgui ml objdetect octvr".split()
LIBS = "-lopencv_octvr -lopencv_stitching -lopencv_calib3d -lopencv_features2d -lopencv_flann -lopencv_imgcodecs " \
       "-lopencv_imgproc -lopencv_core".split()
DEPTH = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}


def compile_tool(name):
    exe = os.path.join(TMP, name)
    inc = ["-I" + B] + ["-I%s/modules/%s/include" % (S, m) for m in MODS]
    cmd = ["/usr/bin/g++", "-std=c++11", "-O2", "-w"] + inc + [os.path.join(HERE, name + ".cpp"), "-o", exe,
           "-L%s/lib" % B] + LIBS + ["-L%s/3rdparty/lib" % B, "-llibjpeg", "-llibpng", "-lzlib", "-lpthread", "-ldl"]
    subprocess.check_call(cmd)
    return exe


def read_container(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if not h:
                break
            nl, = struct.unpack("<I", h)
            name = f.read(nl).decode()
            depth, cn, rows, cols = struct.unpack("<IIQQ", f.read(24))
            dt = np.dtype(DEPTH[depth])
            a = np.frombuffer(f.read(rows * cols * cn * dt.itemsize), dtype=dt)
            out[name] = a.reshape(rows, cols, cn) if cn > 1 else a.reshape(rows, cols)
    return out


def main():
    only = sys.argv[1:]            # optional: regenerate only the named rigs / stitch cases (new fixtures without touching the others)
    os.makedirs(TMP, exist_ok=True)
    import oracle as O
    rigs_dir = os.path.join(GOLD, "rigs")
    widths = json.load(open(os.path.join(rigs_dir, "widths.json")))
    dump = os.path.join(B, "bin", "octvr_dump")
    for rig, w in widths.items():
        if only and rig not in only:
            continue
        dat = os.path.join(TMP, rig + ".dat")
        # cwd = the rigs directory: an ocam_fisheye "file" option names its calibration file relative to it
        subprocess.check_call([dump, "-w", str(w), "-o", dat, os.path.join(rigs_dir, rig + ".json")], cwd=rigs_dir,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t = O.load_dat(dat)
        arrs = {"out_size": np.array(t.out_size, np.int64), "n": np.array(len(t.inputs))}
        for i, d in enumerate(t.inputs):
            arrs["roi%d" % i] = np.array(d["roi"], np.int64)
            arrs["map1_%d" % i], arrs["map2_%d" % i], arrs["mask%d" % i] = d["map1"], d["map2"], d["mask"]
            if d["vignette"] is not None:
                arrs["vig%d" % i] = d["vignette"]
            arrs["seam%d" % i] = t.seam_masks[i]
        arrs["n_ov"] = np.array(len(t.overlay_inputs))
        for i, d in enumerate(t.overlay_inputs):
            arrs["ov_roi%d" % i] = np.array(d["roi"], np.int64)
            arrs["ov_map1_%d" % i], arrs["ov_map2_%d" % i], arrs["ov_mask%d" % i] = d["map1"], d["map2"], d["mask"]
        np.savez_compressed(os.path.join(GOLD, "tmpl_%s.npz" % rig), **arrs)
        print("template", rig, t.out_size, len(t.inputs))
        # digest of the reference tool's own file: the product's octvr_dump must write the same bytes (tests/test_gpu_cli.py)
        import hashlib
        hp = os.path.join(GOLD, "dat_sha256.json")
        hs = json.load(open(hp)) if os.path.exists(hp) else {}
        if rig in ("rig3", "rig3ov", "rig2s", "models", "masks"):
            hs[rig] = {"sha256": hashlib.sha256(open(dat, "rb").read()).hexdigest(), "bytes": os.path.getsize(dat)}
            json.dump(hs, open(hp, "w"), indent=1)

    if not only or "primitives" in only:
        gbin = os.path.join(TMP, "golden.bin")
        subprocess.check_call([compile_tool("ref_golden"), gbin])
        np.savez_compressed(os.path.join(GOLD, "primitives.npz"), **read_container(gbin))
        print("primitives ok")

    if not only or "fillpoly" in only:      # cv::fillPoly of the reference build on 310 polygons (camera masks)
        rng = np.random.default_rng(11)
        cases = [((64, 48), [(10, 10), (50, 12), (40, 40), (8, 30)]), ((64, 48), [(5, 5), (60, 5), (60, 40), (32, 15), (5, 40)]),
                 ((64, 48), [(5, 5), (60, 40), (60, 5), (5, 40)]), ((64, 48), [(30, 20), (30, 20), (30, 20)]), ((64, 48), [(3, 7), (50, 7)]),
                 ((64, 48), [(-20, -10), (90, 5), (70, 70), (-5, 40)]), ((64, 48), [(100, 100), (120, 100), (110, 130)]),
                 ((320, 240), [(30, 20), (30, 219), (289, 219), (289, 20)]), ((320, 240), [(10, 10), (150, 30), (120, 200), (20, 180)]),
                 ((320, 240), [(200, 60), (300, 80), (280, 200), (210, 170)])]
        for _ in range(300):
            w, h = int(rng.integers(8, 200)), int(rng.integers(8, 150))
            n = int(rng.integers(1, 9))
            cases.append(((w, h), [(int(rng.integers(-w // 2, w + w // 2)), int(rng.integers(-h // 2, h + h // 2))) for _ in range(n)]))
        txt = os.path.join(TMP, "fillpoly_cases.txt")
        with open(txt, "w") as f:
            for (w, h), pts in cases:
                f.write("%d %d %d %s\n" % (w, h, len(pts), " ".join("%d %d" % p for p in pts)))
        fbin = os.path.join(TMP, "fillpoly.bin")
        subprocess.check_call([compile_tool("ref_fillpoly"), txt, fbin])
        c = read_container(fbin)
        arrs = {"n": np.array(len(cases))}
        for k, ((w, h), pts) in enumerate(cases):
            arrs["p%d" % k] = np.array([w, h] + [v for p in pts for v in p], np.int32)
            arrs["m%d" % k] = np.packbits(c["m%d" % k] == 200)
        np.savez_compressed(os.path.join(GOLD, "fillpoly.npz"), **arrs)
        print("fillpoly", len(cases))

    # vr::FastMapper (mapper_fast.cpp): the reference constructor's tables + the NV12 frame (oracle/refgen/ref_fast.cpp) on
    # full-frame templates (octvr_dump -n); the second case has an odd output height, so the half-size tables come from the
    # generic bilinear resize instead of the 2x area path
    fast_cases = [("rig3", 256, 320, 240), ("models", 190, 200, 120)]
    if not only or "fast" in only:
        ft = compile_tool("ref_fast")
        for rig, w, iw, ih in fast_cases:
            dat = os.path.join(TMP, rig + "_n.dat")
            subprocess.check_call([dump, "-n", "-w", str(w), "-o", dat, os.path.join(rigs_dir, rig + ".json")], cwd=rigs_dir,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out = os.path.join(TMP, "fast_%s.bin" % rig)
            subprocess.check_call([ft, dat, str(iw), str(ih), out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            c = read_container(out)
            t = O.load_dat(dat)
            c["out_size"], c["n"], c["in_size"] = np.array(t.out_size, np.int64), np.array(len(t.inputs)), np.array([iw, ih], np.int64)
            for i, d in enumerate(t.inputs):
                assert tuple(d["roi"]) == (0, 0) + tuple(t.out_size)
                c["t_map1_%d" % i], c["t_map2_%d" % i], c["t_mask%d" % i] = d["map1"], d["map2"], d["mask"]
            np.savez_compressed(os.path.join(GOLD, "fast_%s.npz" % rig), **c)
            print("fast", rig, t.out_size, c["result"].shape)
    if only and set(only) <= {"fast"}:
        return

    st = compile_tool("ref_stitch")
    cases = [  # name, rig, in_w, in_h, blend, gain, kind
        ("rig3_feather5_gain_noise", "rig3", 320, 240, -5, 1, "noise"),
        ("rig3_feather1_nogain_smooth", "rig3", 320, 240, -1, 0, "smooth"),
        ("rig3_mb16_gain_smooth", "rig3", 320, 240, 16, 1, "smooth"),
        ("rig3_mb8_nogain_noise", "rig3", 320, 240, 8, 0, "noise"),
        ("rig3_noblend_noise", "rig3", 320, 240, 0, 0, "noise"),
        ("rig2s_feather3_gain_smooth", "rig2s", 192, 108, -3, 1, "smooth"),
        ("masks_feather2_gain_noise", "masks", 320, 240, -2, 1, "noise"),
        # after-blend stages (mapper.cpp:279-312): overlay input, scale_output (generic and exact 2x), preview
        ("rig3ov_feather2_gain_noise_scale200x90_prev64x32", "rig3ov", 320, 240, -2, 1, "noise", 200, 90, 64, 32),
        ("rig3ov_mb8_gain_smooth_scale128x64_prev100x60", "rig3ov", 320, 240, 8, 1, "smooth", 128, 64, 100, 60),
        ("rig3ov_noblend_noise_prev128x64", "rig3ov", 320, 240, 0, 0, "noise", 0, 0, 128, 64),
    ]
    for name, rig, iw, ih, blend, gain, kind, *post in cases:
        if only and name not in only:
            continue
        out = os.path.join(TMP, name + ".bin")
        subprocess.check_call([st, os.path.join(TMP, rig + ".dat"), str(iw), str(ih), str(blend), str(gain), kind, out] + [str(v) for v in post],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        c = read_container(out)
        keep = {k: v for k, v in c.items() if not k.startswith("warped") or name.startswith("rig3_feather5")}
        keep["meta"] = np.array([iw, ih, blend, gain, 1 if kind == "noise" else 0] + list(post), np.int64)
        np.savez_compressed(os.path.join(GOLD, ("post_%s.npz" if post else "stitch_%s.npz") % name), **keep)
        print("stitch", name, {k: v.shape for k, v in keep.items() if k.startswith("result") or k == "gains"})


if __name__ == "__main__":
    main()
