"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed for plumbing.

The path shards only where it shards naturally -- there is no exchange step inside a frame:
  * batched views (config C5): views are independent units -> contiguous blocks of views per rank, the 1.5 MB mesh is
    replicated, every rank renders into its own slab; an optional final gather (NCCL over NVLink, or gloo on CPU).
  * very high resolution (config C4): screen-row bands -> rank g owns rows [row0,row1) (tile aligned), sees all the
    triangles, clamps every bounding box to its band; per-pixel arithmetic does not depend on the band, so the
    concatenated bands equal the single-GPU frame bit for bit.  Row bands are contiguous in row-major [H,W,*] buffers,
    so the gather is a plain concatenation along dim 0.
Nothing here touches CUDA directly; the gathers work on whatever device the tensors live on (nccl: cuda, gloo: cpu).
"""
import torch
import torch.distributed as dist

TILE_ROWS = 32   # k_raster tile height: bands are aligned to it so no tile straddles two ranks


def view_shard(n_views, rank, world):
    """Contiguous block of views for `rank`: (first, count).  Remainders go to the lowest ranks."""
    base, rem = divmod(int(n_views), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def view_shard_strided(n_views, rank, world):
    """Round-robin alternative: the views k with k % world == rank (what bench.py uses -- frames of neighbouring view
    angles cost alike, so striding balances the ranks where contiguous arcs do not).  Returns the list of view indices."""
    return list(range(int(rank), int(n_views), int(world)))


def band_shard(h, rank, world, align=TILE_ROWS):
    """Row band [row0,row1) of an h-row image for `rank`, aligned to `align` rows (last band takes the ragged end).
    Bands partition [0,h) exactly; ranks beyond the number of tile rows get an empty band."""
    tiles = (int(h) + align - 1) // align
    base, rem = divmod(tiles, int(world))
    t0 = rank * base + min(rank, rem)
    t1 = t0 + base + (1 if rank < rem else 0)
    return min(t0 * align, h), min(t1 * align, h)


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def gather_views(local, n_views, dst=None):
    """Gathers per-rank view slabs [count_r, ...] (block layout of view_shard) into [n_views, ...].

    dst=None: every rank gets the result (all_gather); dst=k: only rank k does (gather), others get None.
    Ragged shards (n_views not divisible by the world size) are padded to the largest shard for the collective."""
    world = _world()
    if world == 1:
        return local
    rank = dist.get_rank()
    counts = [view_shard(n_views, r, world)[1] for r in range(world)]
    cmax = max(counts)
    if local.shape[0] != cmax:
        pad = torch.zeros((cmax - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    if dst is None:
        out = torch.empty((world * cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
    else:
        parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local, parts, dst=dst)
        if rank != dst:
            return None
        out = torch.cat(parts, dim=0)
    if all(c == cmax for c in counts):
        return out
    return torch.cat([out[r * cmax:r * cmax + counts[r]] for r in range(world)], dim=0)


def gather_bands(local, h, dst=None, align=TILE_ROWS):
    """Gathers per-rank row bands [rows_r, W, ...] (layout of band_shard) into the full [h, W, ...] buffer."""
    world = _world()
    if world == 1:
        return local
    rank = dist.get_rank()
    rows = [band_shard(h, r, world, align) for r in range(world)]
    rmax = max(b - a for a, b in rows)
    if local.shape[0] != rmax:
        pad = torch.zeros((rmax - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    if dst is None:
        out = torch.empty((world * rmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
    else:
        parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local, parts, dst=dst)
        if rank != dst:
            return None
        out = torch.cat(parts, dim=0)
    return torch.cat([out[r * rmax:r * rmax + (b - a)] for r, (a, b) in enumerate(rows)], dim=0)
