"""tools/prof.py -- minimal driver for ncu / quick timing: builds the named workload and runs N stitches.
   python tools/prof.py [workload] [steps]     (diagnostic; bench.py is the measurement contract)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import octvr_b200 as vr
import util
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rig, blend, gain, desc = bench.WORKLOADS[wl]
cfg, width, in_size = util.named_rig(rig)
n = len(cfg["inputs"])
iw, ih = in_size
tmpl = bench.make_template(vr, cfg, width, 0)
t0 = time.time()
m = vr.Mapper(tmpl, [in_size] * n, blend=blend, enable_gain_compensator=gain, device=0)
print("mapper init %.2f s" % (time.time() - t0), m.stats())
W, H = tmpl.out_size
ring = []
for k in range(4):
    fr = []
    for c in range(n):
        y, u, v = util.i420_planes(util.noise_frame(c, iw, ih, seed=1234 + 7919 * k), iw, ih)
        fr.append(torch.from_numpy(np.concatenate([y, np.concatenate([u, v], 1)], 0)).cuda())
    ring.append(fr)
out = torch.zeros((H * 3 // 2, W), dtype=torch.uint8, device="cuda")
for k in range(5):
    m.stitch_packed(ring[k % 4], out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(steps):
    m.stitch_packed(ring[k % 4], out)
e1.record()
torch.cuda.synchronize()
print("ms/step %.4f" % (e0.elapsed_time(e1) / steps))
m.set_profiling(True)
for k in range(3):
    m.stitch_packed(ring[k % 4], out)
    print({s: round(m.stage_ms(s), 4) for s in ("convert", "gain", "blend", "total")})
try:
    d = m.debug_gain_ns()
    print("gain kernel last-CTA stamps (ns): stats %d, reduce %d, solve %d, tables %d; last CTA started %d ns after the first, chain ended %d ns after the first CTA started"
          % (d[1] - d[0], d[2] - d[1], d[3] - d[2], d[4] - d[3], d[0] - d[5], d[4] - d[5]))
except Exception as ex:
    print("no gain stamps:", ex)
try:
    m.debug_ring()
    for k in range(4):
        m.stitch_packed(ring[k % 4], out)
    d = m.debug_ring()
    if d[0]:
        print("ring: %d jobs (warp 0 view), mean wait %.0f ns, mean issue->use %.0f ns, %.1f %% of jobs waited > 200 ns, their mean issue->complete %.0f ns"
              % (d[0], d[1] / d[0], d[2] / d[0], 100.0 * d[3] / d[0], d[4] / max(d[3], 1)))
except Exception as ex:
    print("no ring counters:", ex)
