#!/usr/bin/env python
"""Init-time split on the GPU box: map generation, seam masks (GPU / host), Mapper construction.  usage: time_init.py [rig]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import octvr_b200 as vr
import util

rig = sys.argv[1] if len(sys.argv) > 1 else "rig8L"
cfg, width, in_size = util.named_rig(rig)
h = 1920 if rig.startswith("rig8") else -1
for rep in range(2):
    t0 = time.time(); t = vr.MapperTemplate.from_json(cfg, width, h, with_seam_masks=False); t1 = time.time()
    t.create_masks(); t2 = time.time()
    print("rep %d: mapgen %.3f s, seam masks %.3f s (backend %d)" % (rep, t1 - t0, t2 - t1, vr.lib().octvr_debug_seam_backend()), flush=True)
for blend in (-1, 32):
    t0 = time.time(); m = vr.Mapper(t, [in_size] * t.num_inputs, blend=blend, enable_gain_compensator=True); t1 = time.time()
    print("Mapper(blend=%d) %.3f s" % (blend, t1 - t0), flush=True)
    del m
