#!/usr/bin/env python
"""tools/pcie_probe.py -- what the host <-> device link of this box delivers for the e2e path's copy pattern (diagnostic).
Pinned host memory; per "frame" 6 H2D copies of 6.2 MB (C2's inputs) and one D2H copy of 12.6 MB (its output)."""
import torch, time
torch.cuda.init()
n_in, in_b, out_b = 6, 2704 * 1520 * 3 // 2, 4096 * 2048 * 3 // 2
h_in = [torch.empty(in_b, dtype=torch.uint8).pin_memory() for _ in range(n_in)]
d_in = [torch.empty(in_b, dtype=torch.uint8, device="cuda") for _ in range(n_in)]
h_out = torch.empty(out_b, dtype=torch.uint8).pin_memory()
d_out = torch.empty(out_b, dtype=torch.uint8, device="cuda")
big_h = torch.empty(n_in * in_b, dtype=torch.uint8).pin_memory()
big_d = torch.empty(n_in * in_b, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

def run(name, fn, frames=60):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(frames): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / frames
    print("%-58s %.3f ms / frame  H2D %.1f GB/s  (%.0f frames/s)" % (name, dt * 1e3, n_in * in_b / dt / 1e9, 1 / dt))

def h2d_one_stream():
    with torch.cuda.stream(s1):
        for h, d in zip(h_in, d_in): d.copy_(h, non_blocking=True)
def h2d_big():
    with torch.cuda.stream(s1): big_d.copy_(big_h, non_blocking=True)
def h2d_two_streams():
    for k, (h, d) in enumerate(zip(h_in, d_in)):
        with torch.cuda.stream(s1 if k % 2 == 0 else s2): d.copy_(h, non_blocking=True)
def duplex():
    h2d_one_stream()
    with torch.cuda.stream(s3): h_out.copy_(d_out, non_blocking=True)
def duplex_two():
    h2d_two_streams()
    with torch.cuda.stream(s3): h_out.copy_(d_out, non_blocking=True)
run("H2D 6 x 6.2 MB, one stream", h2d_one_stream)
run("H2D 1 x 37 MB, one stream", h2d_big)
run("H2D 6 x 6.2 MB, two streams", h2d_two_streams)
run("H2D one stream + D2H 12.6 MB on another", duplex)
run("H2D two streams + D2H 12.6 MB on a third", duplex_two)
