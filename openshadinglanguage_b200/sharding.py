"""Point-range / tile sharding across the GPUs of one box (host-side logic).

Shading points are independent, so each rank gets a
contiguous range of shade indices (grid rows for testshade) and no data-path
collective is needed; the only exchange is gathering output strips to rank 0
(NCCL on GPUs, gloo in the CPU tests).
"""


def point_range(npoints, rank, world, align=256):
    """Contiguous [begin, end) of shade indices for `rank`, aligned to the CTA
    tile so every rank's outputs start on a tile boundary."""
    tiles = (npoints + align - 1) // align
    per, extra = divmod(tiles, world)
    t0 = rank * per + min(rank, extra)
    t1 = t0 + per + (1 if rank < extra else 0)
    return min(npoints, t0 * align), min(npoints, t1 * align)


def gather_strips(local, npoints, rank, world, align=256, floats_per_point=None):
    """Gather each rank's output strip (a [n_local, F] tensor on any device)
    to rank 0 with torch.distributed; returns the full [npoints, F] tensor on
    rank 0 and None elsewhere.  Strips differ in length, so this is a gather of
    padded strips (the framebuffer gather of the design: 12 B/pixel, once)."""
    import torch
    import torch.distributed as dist
    F = local.shape[1] if floats_per_point is None else floats_per_point
    ranges = [point_range(npoints, r, world, align) for r in range(world)]
    longest = max(e - b for b, e in ranges)
    pad = torch.zeros((longest, F), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if rank == 0:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=0)
        out = torch.empty((npoints, F), dtype=local.dtype, device=local.device)
        for (b, e), buf in zip(ranges, bufs):
            out[b:e] = buf[: e - b]
        return out
    dist.gather(pad, None, dst=0)
    return None
