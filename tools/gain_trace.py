"""tools/gain_trace.py [workload] -- timeline of the gain CTAs inside k_convert_gain (diagnostic; OCTVR_GAIN_TRACE=1)."""
import os, sys
os.environ["OCTVR_GAIN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import octvr_b200 as vr, util, bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
rig, blend, gain, desc = bench.WORKLOADS[wl]
cfg, width, in_size = util.named_rig(rig)
n = len(cfg["inputs"]); iw, ih = in_size
tmpl = bench.make_template(vr, cfg, width, 0)
m = vr.Mapper(tmpl, [in_size] * n, blend=blend, enable_gain_compensator=gain, device=0)
W, H = tmpl.out_size
ring = []
for k in range(3):
    fr = []
    for c in range(n):
        y, u, v = util.i420_planes(util.noise_frame(c, iw, ih, seed=1234 + 7919 * k), iw, ih)
        fr.append(torch.from_numpy(np.concatenate([y, np.concatenate([u, v], 1)], 0)).cuda())
    ring.append(fr)
out = torch.zeros((H * 3 // 2, W), dtype=torch.uint8, device="cuda")
for k in range(6):
    m.stitch_packed(ring[k % 3], out)
m.debug_gain_trace()
for rep in range(3):
    m.stitch_packed(ring[rep], out)
    d = np.array(m.debug_gain_trace(), dtype=np.uint64).astype(np.int64)
    t0 = d[8::6]; t1 = d[12::6]
    k = int((t0 > 0).sum())
    t0, t1 = t0[:k], t1[:k]
    ph = [np.median(d[8 + j + 1::6][:k] - d[8 + j::6][:k]) / 1e3 for j in range(4)]
    print("phases (median us): sample load %.1f, taps + norms %.1f, pair sums %.1f, atomics + fence %.1f" % tuple(ph))
    base = min(t0.min(), d[6])
    print("gain CTAs %d | start: min %.1f p50 %.1f p90 %.1f max %.1f us | stats duration: p10 %.1f p50 %.1f p90 %.1f max %.1f us | last ticket at %.1f us | chain: ticket %.1f reduced %.1f solved %.1f done %.1f | conversion CTAs: first start %.1f, last end %.1f us"
          % (k, (t0.min() - base) / 1e3, (np.median(t0) - base) / 1e3, (np.percentile(t0, 90) - base) / 1e3, (t0.max() - base) / 1e3,
             np.percentile(t1 - t0, 10) / 1e3, np.median(t1 - t0) / 1e3, np.percentile(t1 - t0, 90) / 1e3, (t1 - t0).max() / 1e3, (t1.max() - base) / 1e3,
             (d[1] - base) / 1e3, (d[2] - base) / 1e3, (d[3] - base) / 1e3, (d[4] - base) / 1e3, (d[6] - base) / 1e3, (d[7] - base) / 1e3))
