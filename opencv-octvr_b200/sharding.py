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


def broadcast_frames(frames, src=0):
    """Row-band mode: every rank needs every input frame.  `frames`: list of equally shaped u8 tensors, valid on `src`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for f in frames:
            dist.broadcast(f, src=src)
    return frames


def collect_bands(out, dst=0):
    """Row-band mode: each rank's output buffer is zero outside its own band, so the full frame is the element-wise sum.
    (u8 sum; bands are disjoint, nothing overflows.)  The result is valid on `dst`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM)
    return out


class RowBandStitcher:
    """One large frame split over the ranks by output row bands (BASELINE config C4's partition; feather / no blend).
    Every rank constructs it with the same template; stitch() is collective."""

    def __init__(self, vr, tmpl, in_sizes, blend, enable_gain, device, align=32):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.bands = row_bands(tmpl.out_size[1], self.world, align)
        self.mapper = vr.Mapper(tmpl, in_sizes, blend=blend, enable_gain_compensator=enable_gain, device=device,
                                band=self.bands[self.rank])

    def stitch(self, frames_packed, out_packed, src=0, collect=True):
        """frames_packed: per-camera packed (1.5h, w) CUDA tensors (contents valid on `src`); out_packed: full-size
        packed output, zero-initialised by the caller; after the call rank `src` holds the whole frame if collect."""
        broadcast_frames(frames_packed, src)
        self.mapper.stitch_packed(frames_packed, out_packed)
        if collect:
            collect_bands(out_packed, src)
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

    def __init__(self, vr, tmpls, in_sizes, blend, enable_gain, device, align=32, rank=None, world=None):
        assert len(tmpls) == 2 and tuple(tmpls[0].out_size) == tuple(tmpls[1].out_size)
        self.vr = vr
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.eye_w, self.eye_h = tmpls[0].out_size
        self.in_sizes = [tuple(s) for s in in_sizes]
        self.jobs = []
        for eye, b, per in stereo_assignment(self.world)[self.rank]:
            band = row_bands(self.eye_h, per, align)[b]
            m = vr.Mapper(tmpls[eye], in_sizes, blend=blend, enable_gain_compensator=enable_gain, device=device,
                          band=None if per == 1 else band)
            self.jobs.append((eye, band, m))

    def stitch_local(self, frames_packed, out_packed, stream=None):
        """This rank's share: frames in Mapper's packed layout, out_packed the full top-bottom frame (W x 1.5 * 2 * eye_h)."""
        W, He = self.eye_w, self.eye_h
        oy, ou, ov = self.vr.split_packed(out_packed, W, 2 * He)
        ins = [self.vr.split_packed(f, w, h) for f, (w, h) in zip(frames_packed, self.in_sizes)]
        for eye, band, m in self.jobs:
            m.stitch(ins, (oy[eye * He:(eye + 1) * He], ou[eye * He // 2:(eye + 1) * He // 2], ov[eye * He // 2:(eye + 1) * He // 2]),
                     stream=stream)

    def stitch(self, frames_packed, out_packed, src=0, collect=True):
        broadcast_frames(frames_packed, src)
        self.stitch_local(frames_packed, out_packed)
        if collect:
            collect_bands(out_packed, src)
        return out_packed
