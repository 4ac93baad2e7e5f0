"""
Multi-GPU decomposition of the hot path: contiguous x-slabs (axis 0 is the slowest
axis, so a slab is one contiguous block of cells and its sky tile is a contiguous
(nx_slab, nz) block).  Every ray, channel and epoch is independent and cell values are
analytic in the indices, so there is no halo and no reduction: the single exchange is
an all-gather of the finished image/cube tiles (NCCL over NVLink on GPUs; gloo in the
CPU tests).  SURVEY.md section 8(e).
"""


def slab_bounds(nx, rank, world):
    """[x_lo, x_hi) of `rank` when nx planes are dealt to `world` ranks as evenly as
    possible (the first nx % world ranks get one extra plane)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad (rank, world)")
    if world > nx:
        raise ValueError("more ranks than x-planes")
    base, extra = divmod(nx, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def epoch_shares(n_epochs, rank, world):
    """Indices of the epochs `rank` integrates when a time series is sharded by epoch
    (BASELINE config 4): round-robin keeps the per-rank cost even."""
    return list(range(rank, n_epochs, world))


def gather_x(tile, nx, rank, world, dim=0, group=None):
    """All-gather x-slab tiles along dimension `dim` into the full array (every rank gets
    the full result).  `tile` is a torch tensor on the device NCCL/gloo is bound to."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return tile
    if not dist.is_initialized():
        raise RuntimeError("sharded JetModel needs torch.distributed to be initialised")
    sizes = [hi - lo for lo, hi in (slab_bounds(nx, r, world) for r in range(world))]
    tile = tile.contiguous()
    if dim != 0:
        tile = tile.movedim(dim, 0).contiguous()
    big = max(sizes)
    if tile.shape[0] < big:  # uneven split: pad to the largest slab (collectives need equal sizes)
        pad = torch.zeros([big - tile.shape[0]] + list(tile.shape[1:]), dtype=tile.dtype,
                          device=tile.device)
        tile = torch.cat([tile, pad], dim=0)
    out = torch.empty([world * big] + list(tile.shape[1:]), dtype=tile.dtype,
                      device=tile.device)
    dist.all_gather_into_tensor(out, tile, group=group)
    if min(sizes) != big:
        out = torch.cat([out[r * big: r * big + sizes[r]] for r in range(world)], dim=0)
    if dim != 0:
        out = out.movedim(0, dim).contiguous()
    return out
