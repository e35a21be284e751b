"""tools/overlap_probe.py -- does an NCCL broadcast issued ahead of compute overlap with it on this box? (diagnostic; torchrun, >= 2 ranks)"""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=os.environ.get("HP", "1") == "1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
bufs = [torch.zeros(100 * 1000 * 1000, dtype=torch.uint8, device="cuda") for _ in range(2)]
kind = os.environ.get("COMPUTE", "matmul")
stitch = None
if kind in ("c2", "c3"):
    import sys, numpy as np
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for q in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, q)
    import octvr_b200 as vr, util, bench
    rig, blend, gain, desc = bench.WORKLOADS[kind]
    cfg, width, in_size = util.named_rig(rig)
    n = len(cfg["inputs"]); iw, ih = in_size
    tmpl = bench.make_template(vr, cfg, width, local)
    m = vr.Mapper(tmpl, [in_size] * n, blend=blend, enable_gain_compensator=gain, device=local)
    W, H = tmpl.out_size
    fr = []
    for c in range(n):
        y, u, v = util.i420_planes(util.noise_frame(c, iw, ih), iw, ih)
        fr.append(torch.from_numpy(np.concatenate([y, np.concatenate([u, v], 1)], 0)).cuda())
    outp = torch.zeros((H * 3 // 2, W), dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream() if os.environ.get("SIDE", "0") == "1" else None
    def stitch():
        for _ in range(3):
            m.stitch_packed(fr, outp, stream=side)
a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
big = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
def compute():
    if stitch is not None:
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
        stitch()
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
    elif kind == "matmul":
        for _ in range(4): a @ a
    else:                      # many small-CTA elementwise kernels (like the stitch)
        for _ in range(12): big.add_(1.0)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_c = timed(compute)
t_b = timed(lambda: dist.broadcast(bufs[0], 0))
def both():
    w = dist.broadcast(bufs[1], 0, async_op=True)
    compute()
    w.wait()
t_both = timed(both)
if dist.get_rank() == 0:
    print("compute %.3f ms, broadcast %.3f ms, both (broadcast issued first, async) %.3f ms -> overlap %.0f %%" % (t_c, t_b, t_both, 100 * (t_c + t_b - t_both) / min(t_c, t_b)))
dist.destroy_process_group()
