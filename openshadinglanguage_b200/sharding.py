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


def tile_list(xres, yres, tile=64):
    """Row-major list of the image's tiles (x0, y0, w, h); edge tiles are smaller."""
    return [(x, y, min(tile, xres - x), min(tile, yres - y))
            for y in range(0, yres, tile) for x in range(0, xres, tile)]


def rank_tiles(xres, yres, rank, world, tile=64):
    """The interleaved tile work set of `rank`: tile k of the row-major
    list goes to rank k % world, so every GPU gets an even sample of cheap (sky) and expensive
    (glass, interreflection) image regions - contiguous bands do not (the top band of
    render-mx-layer is all sky)."""
    return tile_list(xres, yres, tile)[rank::world]


def tile_pixel_index(tiles, xres):
    """Flat image index y*xres + x of every pixel of a tile work set, in work-set order
    (tile after tile, row-major inside a tile) as an int64 torch tensor."""
    import torch
    parts = []
    for x0, y0, w, h in tiles:
        ys = torch.arange(y0, y0 + h, dtype=torch.int64).unsqueeze(1)
        xs = torch.arange(x0, x0 + w, dtype=torch.int64).unsqueeze(0)
        parts.append((ys * xres + xs).reshape(-1))
    return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64)


def gather_plan(xres, yres, world, device, tile=64):
    """What rank 0 needs to reassemble a tile-sharded frame, computed once per (resolution, world):
    per rank the work-set size and the flat image index of each of its pixels (on `device`)."""
    sets = [rank_tiles(xres, yres, r, world, tile) for r in range(world)]
    sizes = [sum(w * h for _, _, w, h in s) for s in sets]
    return sizes, [tile_pixel_index(s, xres).to(device) for s in sets]


def gather_tiles(local, xres, yres, rank, world, tile=64, plan=None):
    """Framebuffer gather of a tile-sharded render: every rank contributes its work set's
    [npix_local, 3] strip (device memory on GPUs: no host round trip); rank 0 receives the
    padded strips with one gather and scatters them into the [yres, xres, 3] image on its
    device.  Returns the image on rank 0, None elsewhere.  `plan`: gather_plan(...) result,
    so that a renderer does not rebuild the (static) index tables every frame."""
    import torch
    import torch.distributed as dist
    if plan is None:
        plan = gather_plan(xres, yres, world, local.device if rank == 0 else "cpu", tile)
    sizes, index = plan
    longest = max(sizes)
    if local.shape[0] == longest:
        pad = local.contiguous()
    else:
        pad = torch.zeros((longest, 3), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    if rank != 0:
        dist.gather(pad, None, dst=0)
        return None
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.gather(pad, bufs, dst=0)
    img = torch.empty((yres * xres, 3), dtype=local.dtype, device=local.device)
    for idx, n, buf in zip(index, sizes, bufs):
        img[idx] = buf[:n]
    return img.reshape(yres, xres, 3)
