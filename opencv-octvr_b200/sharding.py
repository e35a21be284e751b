"""
Multi-GPU partitioning of the stitch path.
  * Stream / frame sharding (default): stream s runs on rank s mod G with its own replicated tables; no data-path
    collective -- torch.distributed is used only to agree on the timing (max over ranks) and to gather results.
  * Row-band partition of ONE large frame (RowBandStitcher): rank r owns output rows row_bands(H, G, 32)[r] and only
    holds that band's tables; the input frames are broadcast from the ingest rank (the one exchange step of this mode,
    NCCL over NVLink on the GPU box), every rank computes the same gains, and the bands are collected on one rank.
  * Stereo top-bottom output (StereoRowBandStitcher, BASELINE config C4): two templates, one per eye, stacked in one
    output frame (regions (0,0,1,.5) and (0,.5,1,.5), projection_modes.cpp:26-46); the ranks are split between the eyes
    and, within an eye, by row bands.  Multiband band mappers carry their own halo rows (csrc/multiband.cu).
Works with any backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import os
import torch
import torch.distributed as dist


def streams_of_rank(n_streams, rank, world):
    """Round-robin ownership: stream s -> rank s mod world (video streams are independent)."""
    return [s for s in range(n_streams) if s % world == rank]


def row_bands(height, world, align):
    """Row-band partition of one large frame (C4): contiguous bands whose boundaries are multiples of `align`
    (2^bands for multiband, 2 for 4:2:0), as even as the alignment allows.  Returns [(y0, y1)] per rank."""
    units = height // align
    assert units * align == height and units >= world, "height must be a multiple of align and give every rank a band"
    out, y = [], 0
    for r in range(world):
        u = units // world + (1 if r < units % world else 0)
        out.append((y, y + u * align))
        y += u * align
    return out


def max_over_ranks(value, device="cpu"):
    """Timing rule: a multi-GPU number is the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frame_counts(n_frames, device="cpu"):
    """Whole-job throughput numerator: frames stitched by every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(n_frames)]
    t = torch.tensor([int(n_frames)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]


def alloc_frame_set(in_sizes, device, torch_mod=None):
    """All input frames of one time step in ONE contiguous buffer (Mapper's packed layout per camera), so that the
    row-band exchange is a single broadcast.  Returns (flat u8 tensor, [per-camera (1.5h, w) views])."""
    torch_mod = torch_mod or torch
    sizes = [w * h * 3 // 2 for w, h in in_sizes]
    flat = torch_mod.zeros(sum(sizes), dtype=torch_mod.uint8, device=device)
    views, o = [], 0
    for (w, h), n in zip(in_sizes, sizes):
        views.append(flat[o:o + n].view(h * 3 // 2, w))
        o += n
    return flat, views


def crop_packed_frames(src_views, dst_views, in_sizes, windows, stream=None):
    """Source columns [col0, col0 + width) of every packed (1.5 h, w) frame in src_views -> the packed (1.5 h, width) frames in
    dst_views, one kernel launch (octvr_crop_packed_frames).  windows: [(col0, width)] per camera, multiples of 32."""
    import ctypes as C
    from .capi import lib, check
    n = len(src_views)
    s = stream if stream is not None else torch.cuda.current_stream()
    srcp = (C.c_void_p * n)(*[t.data_ptr() for t in src_views])
    dstp = (C.c_void_p * n)(*[t.data_ptr() for t in dst_views])
    sp = (C.c_size_t * n)(*[t.stride(0) for t in src_views])
    dp = (C.c_size_t * n)(*[t.stride(0) for t in dst_views])
    wh = (C.c_int * (2 * n))(*[v for s_ in in_sizes for v in s_])
    c0 = (C.c_int * n)(*[w[0] for w in windows])
    cw = (C.c_int * n)(*[w[1] for w in windows])
    check(lib().octvr_crop_packed_frames(n, srcp, sp, wh, c0, cw, dstp, dp, C.c_void_p(s.cuda_stream)))


def broadcast_frames(frames, src=0, async_op=False):
    """Row-band mode: every rank needs every input frame.  `frames`: one flat tensor (alloc_frame_set) or a list of u8
    tensors, valid on `src`.  async_op=True returns the work handles (wait() before reading the frames)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return []
    if torch.is_tensor(frames):
        frames = [frames]
    works = [dist.broadcast(f, src=src, async_op=async_op) for f in frames]
    return [w for w in works if w is not None]


def band_slices(out, width, height, y0, y1, row0=0):
    """The two contiguous row ranges of a packed (1.5 H, W) frame that hold output rows [y0, y1) of a region starting at
    frame row `row0`: luma rows, and the chroma rows below the luma plane (U | V side by side, mapper.hpp:75-83)."""
    return out[row0 + y0:row0 + y1], out[height + (row0 + y0) // 2:height + (row0 + y1) // 2]


def collect_shares(out, shares, dst=0, wait=True):
    """Row-band mode: `shares[r]` = list of (row_a, row_b) ranges of `out` (a packed frame tensor, rows are contiguous)
    that rank r produced.  Every rank sends its ranges to `dst` point to point (batched isend / irecv: NCCL send/recv over
    NVLink on the GPU box); only the bands move, nothing is summed and `out` needs no zero fill.
    wait=False returns the work handles instead of waiting: the transfer then runs under whatever is enqueued next (the
    caller must not touch `out` before waiting on them)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return out if wait else []
    rank = dist.get_rank()
    ops = []
    for r, ranges in enumerate(shares):
        if r == dst:
            continue
        for a, b in ranges:
            if b <= a:
                continue
            if rank == r:
                ops.append(dist.P2POp(dist.isend, out[a:b], dst))
            elif rank == dst:
                ops.append(dist.P2POp(dist.irecv, out[a:b], r))
    works = dist.batch_isend_irecv(ops) if ops else []
    if not wait:
        return works
    for w in works:
        w.wait()
    return out


def collect_bands(out, dst=0):
    """Row-band mode, simple form: each rank's output buffer is zero outside its own band, so the full frame is the
    element-wise sum (u8; bands are disjoint, nothing overflows).  The result is valid on `dst`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM)
    return out


class PeerFrame:
    """A packed output frame (rows x width u8) that lives on rank `owner` and is mapped into every rank of the node
    (octvr_shared_alloc / octvr_shared_open: CUDA IPC, NVLink peer access).  `tensor` is a torch view of it on every rank;
    a band mapper that is given this view as its output stores its rows straight into the owner's memory, so the row-band
    mode needs no collection step.  Collective constructor (the handle travels by all_gather_object)."""

    def __init__(self, rows, width, device, owner=0):
        import ctypes as C
        from .capi import lib, check
        self._C, self._lib, self._check = C, lib, check
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        self.owner, self.opened = owner, self.rank != owner
        self.nbytes = rows * width
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        if self.rank == owner:
            check(lib().octvr_shared_alloc(C.c_size_t(self.nbytes), int(device), C.byref(ptr), handle))
        blobs = [None] * world
        if world > 1:
            dist.all_gather_object(blobs, bytes(handle) if self.rank == owner else None)
        if self.rank != owner:
            handle = (C.c_ubyte * 64).from_buffer_copy(blobs[owner])
            check(lib().octvr_shared_open(handle, int(device), C.byref(ptr)))
        self.ptr = ptr.value

        class _Iface:                                   # torch.as_tensor reads __cuda_array_interface__
            pass
        holder = _Iface()
        holder.__cuda_array_interface__ = {"shape": (rows, width), "typestr": "|u1", "data": (self.ptr, False), "version": 3, "strides": None}
        self._holder = holder
        self.tensor = torch.as_tensor(holder, device=torch.device("cuda", device))

    def close(self):
        if self.ptr:
            torch.cuda.synchronize()
            if dist.is_initialized() and dist.get_world_size() > 1:
                dist.barrier()                          # nobody may still be storing into the owner's memory
            self.tensor = None
            self._check(self._lib().octvr_shared_close(self._C.c_void_p(self.ptr), int(self.opened)))
            self.ptr = None


_SIGNAL_GROUP = {}


def signal_group():
    """A second communicator (its own NCCL stream) for the 4-byte "frame complete" all-reduce: on the default group it
    would queue behind the broadcast of the NEXT step's input frames, which is issued first and takes ~0.2 ms."""
    if "g" not in _SIGNAL_GROUP:
        _SIGNAL_GROUP["g"] = dist.new_group() if dist.is_initialized() and dist.get_world_size() > 1 else None
    return _SIGNAL_GROUP["g"]


def frame_complete(token, async_op=False):
    """Row-band mode with PeerFrame outputs: the frame on the owner is complete when every rank's stitch has finished.
    One 4-byte all-reduce enqueued behind the stitch on every rank; the owner's stream is past it only when all are.
    async_op=True returns the work handle: the all-reduce then runs on its communicator's stream behind this rank's stitch
    while the compute stream goes on with the next frame; wait() on the handle before the frame is read or its buffer reused."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(token, group=signal_group(), async_op=async_op)
    return None


class RowBandStitcher:
    """One large frame split over the ranks by output row bands (BASELINE config C4's partition; feather / no blend).
    Every rank constructs it with the same template; stitch() is collective."""

    def __init__(self, vr, tmpl, in_sizes, blend, enable_gain, device, align=32):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.bands = row_bands(tmpl.out_size[1], self.world, align)
        self.mapper = vr.Mapper(tmpl, in_sizes, blend=blend, enable_gain_compensator=enable_gain, device=device,
                                band=self.bands[self.rank])

    def shares(self):
        H = self.mapper.out_size[1]
        return [[(y0, y1), (H + y0 // 2, H + y1 // 2)] for (y0, y1) in self.bands]

    def stitch(self, frames_packed, out_packed, src=0, collect=True, flat=None):
        """frames_packed: per-camera packed (1.5h, w) CUDA tensors (contents valid on `src`; views of `flat` when the
        frame set was made by alloc_frame_set: one broadcast instead of one per camera); out_packed: full-size packed
        output; after the call rank `src` holds the whole frame if collect."""
        broadcast_frames(flat if flat is not None else frames_packed, src)
        self.mapper.stitch_packed(frames_packed, out_packed)
        if collect:
            collect_shares(out_packed, self.shares(), src)
        return out_packed


def stereo_assignment(world):
    """rank -> [(eye, band index, bands per eye)].  One rank does both eyes; otherwise the ranks are split evenly between
    the eyes and each eye's rows are cut into world / 2 bands."""
    if world == 1:
        return [[(0, 0, 1), (1, 0, 1)]]
    assert world % 2 == 0, "stereo row bands need an even number of ranks"
    per = world // 2
    return [[(r // per, r % per, per)] for r in range(world)]


class StereoRowBandStitcher:
    """Two eye templates of equal size -> one top-bottom frame, partitioned over the ranks by (eye, row band).
    Every rank constructs it with the same templates; stitch() is collective (input broadcast, band collection)."""

    def __init__(self, vr, tmpls, in_sizes, blend, enable_gain, device, align=32, rank=None, world=None, split="auto"):
        """split: how an eye is cut when several ranks share it -- "rows" (bands of output rows), "cols" (bands of output
        columns; multiband only) or "auto" (rows unless OCTVR_C4_SPLIT says otherwise).  Every band carries a halo of
        4 * 2^bands pixels either side of the cut (C4, 8 GPUs: 480 + 2 * 128 of 1920 rows, or 1920 + 2 * 128 of 7680 columns)."""
        assert len(tmpls) == 2 and tuple(tmpls[0].out_size) == tuple(tmpls[1].out_size)
        self.vr = vr
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.eye_w, self.eye_h = tmpls[0].out_size
        self.align = align
        self.in_sizes = [tuple(s) for s in in_sizes]
        if split == "auto":
            # measured on 8 B200s (C4): column bands cut the multiband stage per rank from 0.292 to 0.259 ms but every rank then
            # converts all rows of (nearly) all cameras (0.068 -> 0.107 ms): 0.511 vs 0.512 ms per frame -- rows stay the default
            split = os.environ.get("OCTVR_C4_SPLIT", "rows")
        assert split in ("rows", "cols")
        self.split = split
        from concurrent.futures import ThreadPoolExecutor

        def make(job):
            eye, b, per = job
            band, cols = (0, self.eye_h), None
            if per > 1 and split == "rows":
                band = row_bands(self.eye_h, per, align)[b]
            elif per > 1:
                cols = row_bands(self.eye_w, per, align)[b]
            m = vr.Mapper(tmpls[eye], in_sizes, blend=blend, enable_gain_compensator=enable_gain, device=device,
                          band=None if (per == 1 or cols is not None) else band, cols=cols)
            return (eye, band if cols is None else ("cols",) + tuple(cols), m)
        mine = stereo_assignment(self.world)[self.rank]
        with ThreadPoolExecutor(max_workers=max(1, len(mine))) as ex:      # one rank, both eyes: the two mappers are built side by side
            self.jobs = list(ex.map(make, mine))

    def source_cols(self):
        """per camera: the source columns (lo, hi) this rank's mappers read (tables and gain samples)."""
        cols = None
        for _, _, m in self.jobs:
            c = m.src_cols()
            cols = c if cols is None else [(min(a[0], b[0]), max(a[1], b[1])) for a, b in zip(cols, c)]
        return cols

    def set_input_windows(self, windows):
        """The frames passed to stitch_local() from now on are packed (1.5 h, width) frames holding source columns
        [col0, col0 + width) of every camera only (octvr_mapper_set_input_window): windows = [(col0, width)] per camera."""
        for _, _, m in self.jobs:
            for c, (col0, width) in enumerate(windows):
                m.set_input_window(c, col0, width)
        self.frame_sizes = [(width, h) for (col0, width), (w, h) in zip(windows, self.in_sizes)]

    def stitch_local(self, frames_packed, out_packed, stream=None):
        """This rank's share: frames in Mapper's packed layout, out_packed the full top-bottom frame (W x 1.5 * 2 * eye_h)."""
        W, He = self.eye_w, self.eye_h
        oy, ou, ov = self.vr.split_packed(out_packed, W, 2 * He)
        ins = [self.vr.split_packed(f, w, h) for f, (w, h) in zip(frames_packed, getattr(self, "frame_sizes", self.in_sizes))]
        for eye, band, m in self.jobs:
            m.stitch(ins, (oy[eye * He:(eye + 1) * He], ou[eye * He // 2:(eye + 1) * He // 2], ov[eye * He // 2:(eye + 1) * He // 2]),
                     stream=stream)

    def shares(self):
        """Row ranges of the packed top-bottom frame produced by every rank (for collect_shares)."""
        if self.split != "rows" and self.world > 2:
            raise NotImplementedError("column bands are stored straight into the collecting rank's frame (PeerFrame); there is no send / recv collection for them")
        He, H = self.eye_h, 2 * self.eye_h
        out = []
        for jobs in stereo_assignment(self.world):
            rr = []
            for eye, b, per in jobs:
                y0, y1 = row_bands(He, per, self.align)[b]
                rr += [(eye * He + y0, eye * He + y1), (H + (eye * He + y0) // 2, H + (eye * He + y1) // 2)]
            out.append(rr)
        return out

    def stitch(self, frames_packed, out_packed, src=0, collect=True, flat=None):
        """Collective: broadcast (one call when `flat`, the buffer behind frames_packed, is given), this rank's share,
        bands sent to `src`.  FramePipeline overlaps the broadcast of the next time step with the stitch."""
        broadcast_frames(flat if flat is not None else frames_packed, src)
        self.stitch_local(frames_packed, out_packed)
        if collect:
            collect_shares(out_packed, self.shares(), src)
        return out_packed


class FramePipeline:
    """Row-band exchange overlapped with the stitch ("overlap with previous frame"): the input broadcast of
    time step k + 1 runs under the stitch of step k, and -- with defer_collect -- so does the band collection of step k - 1.
    stitcher: RowBandStitcher or StereoRowBandStitcher.  With defer_collect the caller alternates between (at least) two
    output buffers: a buffer is being read by its collection until the step after next begins (or flush() returns)."""

    def __init__(self, stitcher, src=0, defer_collect=False, peer=False):
        """peer=True: `out` passed to step() is a PeerFrame.tensor (the owner's memory mapped on every rank): the bands are
        stored there directly, and the collection is replaced by frame_complete()."""
        self.st, self.src, self.pending, self.defer, self.peer = stitcher, src, {}, defer_collect, peer
        self.collecting = {}                            # output buffer -> work handles of its band collection
        self.token = None
        # peer mode: the completion signal of frame k is waited for when its output buffer is next written (or in flush()), not
        # right behind the stitch: it then costs the compute stream nothing (OCTVR_C4_SIGNAL=sync: wait at once, 0.02 ms per step)
        self.signal_sync = os.environ.get("OCTVR_C4_SIGNAL", "deferred") == "sync"
        self.signals, self.tokens = {}, {}
        self.prepare, self.side = None, None           # prepare(flat): fills `flat` on the source rank right before it is broadcast
        self.trace = [] if os.environ.get("OCTVR_C4_TRACE") else None
        if peer:
            signal_group()                              # collective: every rank builds the pipeline

    def _wait_collect(self, out):
        for w in self.collecting.pop(out.data_ptr(), []):
            w.wait()

    def step(self, flat, frames, out, next_flat=None, collect=True):
        key = flat.data_ptr()
        works = self.pending.pop(key, None)
        if works is None:
            works = self._broadcast(flat)
        for w in works:
            w.wait()                                   # the compute stream waits for the broadcast, the host does not
        if self.trace is not None:
            self._mark("inputs")
        if next_flat is not None:
            self.pending[next_flat.data_ptr()] = self._broadcast(next_flat)
        self._wait_collect(out)                        # an earlier frame may still be leaving this buffer
        w = self.signals.pop(out.data_ptr(), None)     # ... or its "frame complete" signal may still be in flight
        if w is not None:
            w.wait()
        if hasattr(self.st, "stitch_local"):
            self.st.stitch_local(frames, out)
        else:
            self.st.mapper.stitch_packed(frames, out)
        if self.trace is not None:
            self._mark("stitch")
        if collect and self.peer:
            if self.signal_sync:
                if self.token is None:
                    self.token = torch.zeros(1, dtype=torch.int32, device=out.device)
                frame_complete(self.token)
            else:
                tok = self.tokens.setdefault(out.data_ptr(), torch.zeros(1, dtype=torch.int32, device=out.device))
                self.signals[out.data_ptr()] = frame_complete(tok, async_op=True)
        elif collect:
            if self.defer:
                self.collecting[out.data_ptr()] = collect_shares(out, self.st.shares(), self.src, wait=False)
            else:
                collect_shares(out, self.st.shares(), self.src)
        if self.trace is not None:
            self._mark("done")
        return out

    def _broadcast(self, flat):
        """Asynchronous broadcast of a frame set; on the source rank `prepare` (e.g. the crop of the full frames to the
        columns that are read) runs first, on a side stream, so that neither it nor the broadcast waits behind the stitch."""
        rank = dist.get_rank() if dist.is_initialized() else 0
        if self.prepare is None or rank != self.src:
            return broadcast_frames(flat, self.src, async_op=True)
        if self.side is None:
            self.side = torch.cuda.Stream()
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            self.prepare(flat)
            return broadcast_frames(flat, self.src, async_op=True)

    def _mark(self, what):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.trace.append((what, e))

    def trace_report(self):
        """OCTVR_C4_TRACE: per step, ms the compute stream spent waiting for the input broadcast, in the stitch, and in the
        completion signal (CUDA events on the compute stream; call after a synchronize)."""
        rows, prev = [], None
        ev = self.trace
        for i in range(0, len(ev) - 2, 3):
            a, b, c = ev[i][1], ev[i + 1][1], ev[i + 2][1]
            rows.append((prev.elapsed_time(a) if prev is not None else 0.0, a.elapsed_time(b), b.elapsed_time(c)))
            prev = c
        return rows

    def flush(self):
        """Wait (on the current stream) for every band collection and completion signal still in flight."""
        for key in list(self.collecting):
            for w in self.collecting.pop(key):
                w.wait()
        for key in list(self.signals):
            w = self.signals.pop(key)
            if w is not None:
                w.wait()
