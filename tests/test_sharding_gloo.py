"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes agree on stream ownership, the max-over-ranks
time and the whole-job frame count (the data path itself has no collective)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch.distributed as dist
    import octvr_b200 as vr
    dist.init_process_group("gloo")
    r, w = dist.get_rank(), dist.get_world_size()
    mine = vr.sharding.streams_of_rank(16, r, w)
    t = vr.sharding.max_over_ranks(1.0 + r)
    counts = vr.sharding.gather_frame_counts(100 * len(mine))
    bands = vr.sharding.row_bands(1920, w, 32)
    # row-band mode plumbing on CPU tensors: inputs broadcast from rank 0, each rank fills only its band, bands collected on rank 0
    import torch
    frames = [torch.full((6, 8), 10 * (k + 1) if r == 0 else 0, dtype=torch.uint8) for k in range(3)]
    vr.sharding.broadcast_frames(frames, src=0)
    got = [int(f[0, 0]) for f in frames]
    out = torch.zeros((64, 4), dtype=torch.uint8)
    y0, y1 = vr.sharding.row_bands(64, w, 32)[r]
    out[y0:y1] = r + 1
    vr.sharding.collect_bands(out, dst=0)
    # the same by point-to-point band transfers from one contiguous frame set, with the next step's broadcast in flight
    flat, views = vr.sharding.alloc_frame_set([(8, 4), (8, 4)], "cpu")
    flat2, views2 = vr.sharding.alloc_frame_set([(8, 4), (8, 4)], "cpu")
    if r == 0:
        flat.fill_(7); flat2.fill_(9)
    w1 = vr.sharding.broadcast_frames(flat, src=0, async_op=True)
    w2 = vr.sharding.broadcast_frames(flat2, src=0, async_op=True)
    for x in w1 + w2:
        x.wait()
    out2 = torch.full((96, 4), 99, dtype=torch.uint8)       # packed 64-row frame: 64 luma rows + 32 chroma rows
    shares = [[(a, b), (64 + a // 2, 64 + b // 2)] for a, b in vr.sharding.row_bands(64, w, 32)]
    for a, b in shares[r]:
        out2[a:b] = r + 1
    vr.sharding.collect_shares(out2, shares, dst=0)
    # FramePipeline with deferred band collection: three time steps through two output buffers
    class Fake:
        def shares(self):
            return shares
        def stitch_local(self, frames, o):
            for a, b in shares[r]:
                o[a:b] = int(frames[0][0, 0]) + r            # depends on the (broadcast) frame and on the rank
    sets = [vr.sharding.alloc_frame_set([(8, 4)], "cpu") for _ in range(3)]
    if r == 0:
        for k, (fl, _) in enumerate(sets):
            fl.fill_(10 * (k + 1))
    pipe = vr.sharding.FramePipeline(Fake(), src=0, defer_collect=True)
    bufs = [torch.zeros((96, 4), dtype=torch.uint8) for _ in range(2)]
    seen = []
    for k in range(3):
        pipe.step(sets[k][0], sets[k][1], bufs[k & 1], next_flat=sets[k + 1][0] if k < 2 else None)
        if k >= 1:                                            # frame k - 1 is complete once its buffer is flushed / reused
            pass
    pipe.flush()
    seen = [bufs[0][:, 0].tolist(), bufs[1][:, 0].tolist()]   # buffer 0 holds frame 2, buffer 1 frame 1
    # peer mode (bands stored straight into the owner's frame, no collection): one "frame complete" all-reduce per step on its
    # own group, waited for when the output buffer is next written or in flush() -- five steps through two buffers
    pipe2 = vr.sharding.FramePipeline(Fake(), src=0, peer=True)
    bufs2 = [torch.zeros((96, 4), dtype=torch.uint8) for _ in range(2)]
    in_flight = []
    for k in range(5):
        pipe2.step(sets[k %% 3][0], sets[k %% 3][1], bufs2[k & 1], next_flat=sets[(k + 1) %% 3][0] if k < 4 else None)
        in_flight.append(len(pipe2.signals))
    pipe2.flush()
    peer = {"in_flight": in_flight, "left": len(pipe2.signals), "tokens": sorted(int(t[0]) for t in pipe2.tokens.values()),
            "mine": [int(bufs2[0][shares[r][0][0], 0]), int(bufs2[1][shares[r][0][0], 0])]}
    # one write() per rank (< PIPE_BUF): the two ranks share the pipe and print()'s separate newline write can interleave
    os.write(1, (json.dumps({"rank": r, "mine": mine, "t": t, "counts": counts, "bands": bands, "bcast": got, "rows": out[:, 0].tolist(),
                             "set": [int(views[1][0, 0]), int(views2[0][5, 7])], "rows2": out2[:, 0].tolist(), "pipe": seen, "peer": peer}) + chr(10)).encode())
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    rows = sorted((json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")), key=lambda d: d["rank"])
    assert len(rows) == 2
    assert rows[0]["mine"] == list(range(0, 16, 2)) and rows[1]["mine"] == list(range(1, 16, 2))
    assert sorted(rows[0]["mine"] + rows[1]["mine"]) == list(range(16))          # every stream owned exactly once
    assert rows[0]["t"] == rows[1]["t"] == 2.0                                   # max over ranks
    assert rows[0]["counts"] == rows[1]["counts"] == [800, 800]
    assert rows[0]["bands"] == [[0, 960], [960, 1920]]
    assert rows[0]["bcast"] == rows[1]["bcast"] == [10, 20, 30]                   # every rank sees the ingest rank's frames
    assert rows[0]["rows"] == [1] * 32 + [2] * 32                                 # the bands assembled on rank 0
    assert rows[0]["set"] == rows[1]["set"] == [7, 9]
    assert rows[0]["rows2"] == [1] * 32 + [2] * 32 + [1] * 16 + [2] * 16          # luma and chroma bands, point to point
    # deferred collection: rank 0 ends with frame 2 (value 30) in buffer 0 and frame 1 (value 20) in buffer 1, both ranks' bands
    assert rows[0]["pipe"][0] == [30] * 32 + [31] * 32 + [30] * 16 + [31] * 16
    assert rows[0]["pipe"][1] == [20] * 32 + [21] * 32 + [20] * 16 + [21] * 16


    # peer mode: at most one signal per output buffer in flight, none left after flush(); each token was all-reduced (summed over
    # the 2 ranks) once per use of its buffer: 3 uses of buffer 0, 2 of buffer 1 -> 0 * 2^3 = 0 stays 0 (the token carries no data)
    for r in rows:
        assert r["peer"]["in_flight"] == [1, 2, 2, 2, 2] and r["peer"]["left"] == 0 and r["peer"]["tokens"] == [0, 0]
    assert rows[0]["peer"]["mine"] == [20, 10] and rows[1]["peer"]["mine"] == [21, 11]   # frames 4 (set 1) and 3 (set 0) of this rank's band


def test_row_bands_alignment():
    import octvr_b200 as vr
    b = vr.sharding.row_bands(1920, 8, 32)
    assert b[0][0] == 0 and b[-1][1] == 1920 and all(y0 % 32 == 0 and y1 % 32 == 0 for y0, y1 in b)
    assert all(b[i][1] == b[i + 1][0] for i in range(7))
    assert max(y1 - y0 for y0, y1 in b) - min(y1 - y0 for y0, y1 in b) <= 32
