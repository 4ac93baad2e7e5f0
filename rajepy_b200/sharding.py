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
    the full result).  `tile` is a torch tensor on the device NCCL/gloo is bound to.

    Equal slabs (the normal case) need no staging copy at all: the tiles are gathered
    as they are into a (world, ...) buffer and the result is returned as a VIEW of it with
    the rank axis folded into x (for dim > 0 the view is non-contiguous; consumers that
    need contiguity copy once, e.g. straight into pinned host memory)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return tile
    if not dist.is_initialized():
        raise RuntimeError("sharded JetModel needs torch.distributed to be initialised")
    sizes = [hi - lo for lo, hi in (slab_bounds(nx, r, world) for r in range(world))]
    tile = tile.contiguous()
    big = max(sizes)
    if min(sizes) == big:
        flat = torch.empty(world * tile.numel(), dtype=tile.dtype, device=tile.device)
        dist.all_gather_into_tensor(flat, tile.view(-1), group=group)
        out = flat.view([world] + list(tile.shape))
        # (world, d0, .., nxs, ..) -> (d0, .., world, nxs, ..) -> fold world into x
        out = out.movedim(0, dim)
        shape = list(tile.shape)
        shape[dim] = world * big
        return out.reshape(shape) if dim == 0 else _fold(out, dim)
    # uneven split: pad to the largest slab (collectives need equal sizes)
    if dim != 0:
        tile = tile.movedim(dim, 0).contiguous()
    if tile.shape[0] < big:
        pad = torch.zeros([big - tile.shape[0]] + list(tile.shape[1:]), dtype=tile.dtype,
                          device=tile.device)
        tile = torch.cat([tile, pad], dim=0)
    out = torch.empty([world * big] + list(tile.shape[1:]), dtype=tile.dtype,
                      device=tile.device)
    dist.all_gather_into_tensor(out, tile, group=group)
    out = torch.cat([out[r * big: r * big + sizes[r]] for r in range(world)], dim=0)
    if dim != 0:
        out = out.movedim(0, dim).contiguous()
    return out


class _FoldedView:
    """(d0, .., world, nxs, ..) strided device tensor presented as (d0, .., world*nxs, ..)
    without copying.  Supports what the callers need: shape, copy into a host buffer,
    materialisation, indexing of the leading axis."""

    def __init__(self, t, dim):
        self._t, self._dim = t, dim
        shp = list(t.shape)
        self.shape = tuple(shp[:dim] + [shp[dim] * shp[dim + 1]] + shp[dim + 2:])
        self.dtype, self.device = t.dtype, t.device

    def contiguous(self):
        return self._t.reshape(self.shape)

    def to_host(self):
        import torch
        host = torch.empty(self.shape, dtype=self.dtype, pin_memory=self.device.type == "cuda")
        host.view(self._t.shape).copy_(self._t, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        return host

    def __getitem__(self, i):
        return self._t[i].reshape(self.shape[1:]) if self._dim > 0 else self.contiguous()[i]


def _fold(t, dim):
    return _FoldedView(t, dim)
