"""
Multi-GPU decomposition of the hot path: contiguous x-slabs (axis 0 is the slowest
axis, so a slab is one contiguous block of cells and its sky tile is a contiguous
(nx_slab, nz) block).  Every ray, channel and epoch is independent and cell values are
analytic in the indices, so there is no halo and no reduction: the single exchange is
an all-gather of the finished image/cube tiles (NCCL over NVLink on GPUs; gloo in the
CPU tests).  SURVEY.md section 8(e).

Sky images (8 MB each at 1024^2) are gathered densely (`gather_x`).  Cubes are gathered
SPARSELY (`exchange_ray_columns`): 94 % of the rays of the BASELINE jet miss the jet and
carry constants (tau_L = 0, flux = NaN) that every rank writes itself from the
all-gathered per-ray extents, so only the cube columns of jet-crossing rays travel over
NVLink (0.55 GB instead of 8.6 GB at 1024^2 x 512 channels x 2 cubes).
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


def balanced_bounds(weights, world, plane_cost=0.0, overlap=1.0):
    """Contiguous x-slabs [(lo, hi)] * world that minimise the cost of the most expensive slab,
    cost(slab) = max(L, W) + overlap * min(L, W),  L = sum of `weights` over its planes,
    W = plane_cost * number of planes.

    `weights`: one per x-plane, an estimate of its in-jet cells (the ray walk / channel loop).
    `plane_cost`: what a plane costs even when it is empty sky -- its rows of constants in
    every cube plane -- in the same unit.  `overlap`: how much of the shorter of the two a
    slab pays on top of the longer one: the constant writer runs beside the channel loop
    (different resources: store bandwidth / instruction issue), but on SMs it shares with it;
    1 = the costs add, 0 = they hide each other completely (measured ~0.3,
    tools/slab_probe.py).  The jet occupies a narrow range of x, so equal-width slabs would
    leave most ranks without any ray to integrate.  Every rank gets >= 1 plane."""
    import numpy as np
    w = np.asarray(weights, dtype=np.float64)
    nx = w.size
    if world < 1 or world > nx:
        raise ValueError("more ranks than x-planes")
    w = w + max(w.sum() + plane_cost * nx, 1.0) * 1e-6 / nx   # empty planes still cost a little
    cum = np.concatenate([[0.0], np.cumsum(w)])
    pc, ov = float(plane_cost), float(overlap)

    def cost(lo, hi):
        a, b = cum[hi] - cum[lo], pc * (hi - lo)
        return max(a, b) + ov * min(a, b)

    def reach(lo, limit, hi_max):
        """Largest hi <= hi_max with cost(lo, hi) <= limit (cost is monotone in hi), >= lo + 1."""
        a, b = lo + 1, hi_max
        if cost(lo, b) <= limit:
            return b
        while b - a > 1:                           # invariant: cost(lo, a) <= limit or a = lo + 1
            m = (a + b) // 2
            if cost(lo, m) <= limit:
                a = m
            else:
                b = m
        return a

    def cut(limit):
        """Greedy slabs of cost <= limit; None if more than `world` are needed."""
        cuts, lo = [0], 0
        for k in range(world):
            hi = reach(lo, limit, nx - (world - k - 1))   # slabs still to come need a plane each
            if cost(lo, hi) > limit:
                return None
            cuts.append(hi)
            lo = hi
        return cuts if cuts[-1] >= nx else None

    hi_t = cost(0, nx)
    lo_t = 0.0
    best = cut(hi_t)
    for _ in range(60):
        mid = 0.5 * (lo_t + hi_t)
        c = cut(mid)
        if c is None:
            lo_t = mid
        else:
            best, hi_t = c, mid
    # the greedy cut packs the leading slabs full and leaves the remainder to the last ones:
    # move the cuts towards the right where that does not raise the maximum, so that the
    # slack is spread over the slabs
    limit = max(cost(best[i], best[i + 1]) for i in range(world)) * (1 + 1e-12)
    for k in range(world - 1, 0, -1):
        # slab k-1 = [best[k-1], best[k]) may grow up to the limit; slab k keeps >= 1 plane
        far = reach(best[k - 1], limit, best[k + 1] - 1)
        if far > best[k]:
            # half way: leaves both neighbours below the limit
            best[k] = best[k] + (far - best[k]) // 2
    return [(best[i], best[i + 1]) for i in range(world)]


def even_bounds(nx, world):
    return [slab_bounds(nx, r, world) for r in range(world)]


def epoch_shares(n_epochs, rank, world):
    """Indices of the epochs `rank` integrates when a time series is sharded by epoch
    (BASELINE config 4): round-robin keeps the per-rank cost even."""
    return list(range(rank, n_epochs, world))


def chan_bounds(nchan, rank, world):
    """[c_lo, c_hi) of the channels `rank` integrates when a line cube is sharded by CHANNEL:
    contiguous blocks (a block of channel planes is a contiguous piece of the (nchan, nx, nz)
    cube, i.e. of the FITS product), as even as possible.  Every rank walks all jet-crossing
    rays for its channels and writes only its planes: no exchange of cube data at all."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad (rank, world)")
    base, extra = divmod(int(nchan), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_channel_totals(local, nchan, rank, world, group=None):
    """All-gather per-channel scalars (e.g. the sky-summed flux of every channel, what
    Pipeline stores as results['flux'], classes.py:2468-2472) of a channel-sharded cube:
    `local` holds the values of chan_bounds(nchan, rank, world); returns (nchan,) on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    nmax = (nchan + world - 1) // world
    pad = torch.zeros(nmax, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    buf = torch.empty(world * nmax, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = chan_bounds(nchan, r, world)
        parts.append(buf[r * nmax: r * nmax + (hi - lo)])
    return torch.cat(parts)


def gather_epochs(local, n_epochs, rank, world, group=None):
    """All-gather per-epoch results of a time series whose epochs were dealt round-robin
    (`epoch_shares`): `local` is (len(epoch_shares(n_epochs, rank, world)), ...) in the order of
    the rank's own epochs; returns (n_epochs, ...) in epoch order on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    nmax = (n_epochs + world - 1) // world
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    buf = torch.empty((world,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(-1), pad.view(-1), group=group)
    # epoch e sits at buf[e % world, e // world]
    out = buf.transpose(0, 1).reshape((nmax * world,) + tuple(local.shape[1:]))
    return out[:n_epochs].contiguous()


def gather_x(tile, nx, rank, world, dim=0, group=None, bounds=None):
    """All-gather x-slab tiles along dimension `dim` into the full array (every rank gets
    the full result).  `tile` is a torch tensor on the device NCCL/gloo is bound to.

    Equal slabs need no staging copy at all: the tiles are gathered
    as they are into a (world, ...) buffer and the result is returned as a VIEW of it with
    the rank axis folded into x (for dim > 0 the view is non-contiguous; consumers that
    need contiguity copy once, e.g. straight into pinned host memory)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return tile
    if not dist.is_initialized():
        raise RuntimeError("sharded JetModel needs torch.distributed to be initialised")
    sizes = [hi - lo for lo, hi in (bounds or even_bounds(nx, world))]
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
    # uneven split (work-balanced slabs differ by up to two orders of magnitude in width, so
    # padding every tile to the widest slab would move mostly padding): every rank drops its
    # tile into a zero-filled full-size array and the arrays are summed.  x + 0 is exact (and
    # NaN + 0 stays NaN), so the result is bit-identical to a gather, up to the sign of zero.
    los = [0]
    for sz in sizes:
        los.append(los[-1] + sz)
    shape = list(tile.shape)
    shape[dim] = los[-1]
    out = torch.zeros(shape, dtype=tile.dtype, device=tile.device)
    out.narrow(dim, los[rank], sizes[rank]).copy_(tile)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
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


# ------------------------------------------------------------------ sparse cube exchange
class TorchColumnOps:
    """pack / scatter / constant fill with plain torch indexing: the CPU (gloo) tests and any
    non-CUDA tensor.  On CUDA tensors JetModel passes the C-ABI kernels instead
    (rjp_pack_rays / rjp_scatter_rays / rjp_fill_missed)."""

    @staticmethod
    def pack(cube, ids, out):
        out[:, :ids.numel()] = cube[:, ids.long()]

    @staticmethod
    def scatter(src, ids, cube):
        cube[:, ids.long()] = src[:, :ids.numel()]

    @staticmethod
    def fill_missed(extents, offset, cubes, values):
        miss = (extents[:, 0] >= extents[:, 1]).nonzero().view(-1) + offset
        for cube, v in zip(cubes, values):
            if cube is not None:
                cube[:, miss] = v


def build_ray_meta(extents_tile, ray_list_tile, x_lo, nx, nz, rank, world, group=None,
                   bounds=None):
    """Exchange, once per model, what the sparse cube gather needs: the per-ray extents of
    every slab and the sorted GLOBAL ids (x * nz + z) of every slab's jet-crossing rays.
    `extents_tile` (nxs*nz, 2) int32, `ray_list_tile` (n_active,) int32 slab-local ids."""
    import torch
    import torch.distributed as dist
    nxs = extents_tile.shape[0] // nz
    ext = gather_x(extents_tile.view(nxs, nz, 2), nx, rank, world, dim=0, group=group,
                   bounds=bounds)
    ext = ext.contiguous().view(nx * nz, 2)
    ids = torch.sort(ray_list_tile.to(torch.int32))[0] + x_lo * nz
    cnt = torch.tensor([ids.numel()], dtype=torch.int64, device=ids.device)
    cnts = torch.empty(world, dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = [int(c) for c in cnts.cpu()]
    nmax = max(max(counts), 1)
    padded = torch.full((nmax,), -1, dtype=torch.int32, device=ids.device)
    padded[:ids.numel()] = ids
    allids = torch.empty((world, nmax), dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(allids.view(-1), padded, group=group)
    return {"extents": ext, "counts": counts, "nmax": nmax,
            "ids": [allids[r, :counts[r]].contiguous() for r in range(world)]}


def exchange_ray_columns(cubes, values, meta, nx, nz, rank, world, ops=TorchColumnOps,
                         group=None, bounds=None):
    """Complete full-size cubes (nchan, nx*nz) whose rows of the own slab are final: all-gather
    the columns of the jet-crossing rays of every slab and write the constants `values[i]`
    (tau: 0, flux: NaN) of the rays that miss the jet in the other slabs.  Entries of `cubes`
    may be None (product not requested)."""
    import torch
    import torch.distributed as dist
    live = [c for c in cubes if c is not None]
    if not live or world == 1:
        return
    nch, nmax = live[0].shape[0], meta["nmax"]
    send = torch.empty((len(live), nch, nmax), dtype=live[0].dtype, device=live[0].device)
    for i, cube in enumerate(live):
        ops.pack(cube, meta["ids"][rank], send[i])
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    for r in range(world):
        if r == rank:
            continue
        lo, hi = (bounds or even_bounds(nx, world))[r]
        ops.fill_missed(meta["extents"][lo * nz: hi * nz], lo * nz, cubes, values)
        for i, cube in enumerate(live):
            ops.scatter(recv[r, i], meta["ids"][r], cube)
