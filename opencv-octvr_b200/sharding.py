"""
Multi-GPU partitioning of the stitch path.
  * Stream / frame sharding (default): stream s runs on rank s mod G with its own replicated tables; no data-path
    collective -- torch.distributed is used only to agree on the timing (max over ranks) and to gather results.
  * Row-band partition of ONE large frame (RowBandStitcher): rank r owns output rows row_bands(H, G, 32)[r] and only
    holds that band's tables; the input frames are broadcast from the ingest rank (the one exchange step of this mode,
    NCCL over NVLink on the GPU box), every rank computes the same gains, and the bands are collected on one rank.
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
