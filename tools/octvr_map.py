#!/usr/bin/env python
"""
octvr_map (apps/octvr/map.cpp:75-140): one-shot stitch of still images through a template file.

    python tools/octvr_map.py [-b BLEND] [-g] map.dat output.png input0.jpg input1.jpg ...

The reference tool drives the OpenCL FastMapper; this one drives the CUDA Mapper (vr::Mapper semantics: BLEND > 0
multiband width, < 0 feather border, 0 none; -g enables gain compensation).  Images are converted BGR -> I420 on the
host (cv2), stitched on the GPU through the C ABI, and the 4:2:0 result is converted back for writing.
Overlay inputs of the template are taken from the trailing image arguments.
"""
import getopt
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv):
    opts, args = getopt.getopt(argv, "b:g")
    o = dict(opts)
    if len(args) < 3:
        sys.stderr.write(__doc__)
        return 1
    import cv2
    import torch
    import octvr_b200 as vr
    t = vr.MapperTemplate.load(args[0])
    n = t.num_inputs + t.num_overlays
    if len(args) - 2 != n:
        sys.stderr.write("octvr_map: the template needs %d images\n" % n)
        return 1
    ins, sizes = [], []
    for f in args[2:]:
        img = cv2.imread(f, 1)
        h, w = img.shape[:2]
        img = img[:h - h % 2, :w - w % 2]
        h, w = img.shape[:2]
        yuv = cv2.cvtColor(img, cv2.COLOR_BGR2YUV_I420).reshape(-1)
        q = (w // 2) * (h // 2)
        d = torch.from_numpy(yuv).cuda()
        ins.append((d[:w * h].view(h, w), d[w * h:w * h + q].view(h // 2, w // 2), d[w * h + q:].view(h // 2, w // 2)))
        sizes.append((w, h))
    m = vr.Mapper(t, sizes, blend=int(o.get("-b", 128)), enable_gain_compensator="-g" in o)
    W, H = m.out_size
    out = torch.zeros(W * H * 3 // 2, dtype=torch.uint8, device="cuda")
    q = (W // 2) * (H // 2)
    m.stitch(ins, (out[:W * H].view(H, W), out[W * H:W * H + q].view(H // 2, W // 2), out[W * H + q:].view(H // 2, W // 2)))
    torch.cuda.synchronize()
    bgr = cv2.cvtColor(out.cpu().numpy().reshape(H * 3 // 2, W), cv2.COLOR_YUV2BGR_I420)
    cv2.imwrite(args[1], bgr)
    if "-g" in o:
        sys.stderr.write("gains: %s\n" % np.array2string(m.gains(), precision=4))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
