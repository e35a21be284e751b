"""Shared helpers for the test-suite: synthetic frames, golden loading, rigs."""
import json
import math
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXPO = [0.80, 0.90, 1.00, 1.10, 1.20, 0.95, 1.05, 0.85]


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def noise_frame(cam, w, h, seed=1234):
    """(1.5h, w) u8 standard I420 buffer: byte o = splitmix64(seed ^ cam<<32 ^ o) & 0xFF."""
    with np.errstate(over="ignore"):
        o = np.arange(w * h * 3 // 2, dtype=np.uint64)
        v = splitmix64(np.uint64(seed) ^ (np.uint64(cam) << np.uint64(32)) ^ o)
    return (v & np.uint64(0xFF)).astype(np.uint8).reshape(h * 3 // 2, w)


def smooth_frame(cam, w, h):
    e = EXPO[cam % 8]
    x = np.arange(w)[None, :]
    y = np.arange(h)[:, None]
    Y = (128 + 80 * np.sin(2 * math.pi * x / w * 3) * np.cos(2 * math.pi * y / h * 2)) * e
    f = np.empty((h * 3 // 2, w), np.uint8)
    f[:h] = np.clip(np.rint(Y), 0, 255).astype(np.uint8)
    xc = np.arange(w // 2)[None, :]
    yc = np.arange(h // 2)[:, None]
    U = np.clip(np.rint(128 + 40 * np.sin(2 * math.pi * xc / (w // 2) * 2) + 0 * yc), 0, 255).astype(np.uint8)
    V = np.clip(np.rint(128 - 40 * np.cos(2 * math.pi * yc / (h // 2) * 3) + 0 * xc), 0, 255).astype(np.uint8)
    a = f.reshape(-1)
    q = (w // 2) * (h // 2)
    a[w * h:w * h + q] = U.ravel()
    a[w * h + q:] = V.ravel()
    return f


def i420_planes(f, w, h):
    """standard I420 (1.5h, w) buffer -> y (h,w), u (h/2,w/2), v (h/2,w/2) views."""
    a = f.reshape(-1)
    q = (w // 2) * (h // 2)
    return a[:w * h].reshape(h, w), a[w * h:w * h + q].reshape(h // 2, w // 2), a[w * h + q:w * h + 2 * q].reshape(h // 2, w // 2)


def rig_json(name):
    cfg = json.load(open(os.path.join(GOLD, "rigs", name + ".json")))
    for cam in [cfg["output"]] + cfg["inputs"]:          # an ocam_fisheye "file" option is relative to the rigs directory
        if "file" in cam.get("options", {}):
            cam["options"]["file"] = os.path.join(GOLD, "rigs", cam["options"]["file"])
    return cfg


def rig_width(name):
    return json.load(open(os.path.join(GOLD, "rigs", "widths.json")))[name]


RIGS = os.path.join(os.path.dirname(GOLD[:-len("/tests/golden")] + "/x"), "rigs")


def named_rig(name):
    """SURVEY.md Appendix B rigs (rigs/*.json): returns (config, output width, (in_w, in_h))."""
    cfg = json.load(open(os.path.join(RIGS, name + ".json")))
    o = cfg["inputs"][0]["options"]
    width = {"rig2": 2048, "rig6": 4096, "rig8L": 7680, "rig8R": 7680}[name]
    return cfg, width, (o["width"], o["height"])


def template_from_gold(O, rig):
    g = np.load(os.path.join(GOLD, "tmpl_%s.npz" % rig))
    t = O.Template()
    t.out_size = tuple(int(v) for v in g["out_size"])
    for i in range(int(g["n"])):
        t.inputs.append(dict(roi=tuple(int(v) for v in g["roi%d" % i]), map1=g["map1_%d" % i], map2=g["map2_%d" % i],
                             mask=g["mask%d" % i], vignette=(g["vig%d" % i] if "vig%d" % i in g else None)))
        t.seam_masks.append(g["seam%d" % i])
    for i in range(int(g["n_ov"]) if "n_ov" in g else 0):
        t.overlay_inputs.append(dict(roi=tuple(int(v) for v in g["ov_roi%d" % i]), map1=g["ov_map1_%d" % i], map2=g["ov_map2_%d" % i],
                                     mask=g["ov_mask%d" % i], vignette=None))
    return t


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)
