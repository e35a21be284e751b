"""
Multi-GPU partitioning of the stitch path.  The path shards by stream / frame: stream s runs on
rank s mod G with its own replicated tables, and there is no data-path collective -- torch.distributed is used
only to agree on the timing (max over ranks) and to gather per-rank results on the host.  Works with any backend
(NCCL on the GPU box, gloo in the CPU tests).
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
