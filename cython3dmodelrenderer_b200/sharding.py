"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed for plumbing.

The path shards only where it shards naturally -- there is no exchange step inside a frame:
  * batched views (config C5): views are independent units -> contiguous blocks of views per rank, the 1.5 MB mesh is
    replicated, every rank renders into its own slab; an optional final gather (NCCL over NVLink, or gloo on CPU).
  * very high resolution (config C4): screen-row bands -> rank g owns rows [row0,row1) (tile aligned), sees all the
    triangles, clamps every bounding box to its band; per-pixel arithmetic does not depend on the band, so the
    concatenated bands equal the single-GPU frame bit for bit.  Row bands are contiguous in row-major [H,W,*] buffers,
    so the gather is a plain concatenation along dim 0.
The gathers work on whatever device the tensors live on (nccl: cuda, gloo: cpu).  PeerFrame is the gather-free variant for
one node: the destination rank's frame is mapped into every rank (CUDA IPC through the C ABI's crb_shared_*), each rank's
filler renders its band straight into it over NVLink, and the "gather" is the rasterizer's own stores.
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


def tile_row_costs(v_by_tri, n_by_tri, h, w, fov=90.0, align=TILE_ROWS, per_row=2.0, per_kilo_triangle=0.57):
    """Estimated cost (microseconds on a B200) of every `align`-row strip of an h x w frame of the [T,3,3] camera-space
    arrays: `per_kilo_triangle` per 1000 drawn triangles whose bounding rows touch the strip + `per_row` for its pixels
    (measured on the 10 M-triangle sphere at 8192^2: the work that depends on the band is dominated by the triangles in
    it).  A load-balancing heuristic only -- float32 torch arithmetic on whatever device the tensors live on, not the
    reference's rounding; the rendered result never depends on where the bands are cut."""
    import math
    v = torch.as_tensor(v_by_tri)
    n = torch.as_tensor(n_by_tri)
    f = 1.0 / math.tan(float(fov) / 2.0 / 180.0 * math.pi)
    y = (v[..., 1] * f / v[..., 2] + 1.0) * (h / 2.0)                       # [T,3] screen rows (pyx:116-130)
    x = (v[..., 0] * (f * h / w) / v[..., 2] + 1.0) * (w / 2.0)
    drawn = ~(n[..., 2].sum(dim=1) >= 0)                                      # pyx:202-204
    drawn &= (x.amin(dim=1) < w) & (x.amax(dim=1) > 0) & (y.amin(dim=1) < h) & (y.amax(dim=1) > 0)
    strips = (int(h) + align - 1) // align
    lo = (y.amin(dim=1).clamp(0, h - 1) / align).floor().long()[drawn]
    hi = (y.amax(dim=1).clamp(0, h - 1) / align).floor().long()[drawn]
    diff = torch.zeros(strips + 1, dtype=torch.float64, device=v.device)
    diff.index_add_(0, lo, torch.ones_like(lo, dtype=torch.float64))
    diff.index_add_(0, hi + 1, -torch.ones_like(hi, dtype=torch.float64))
    tris = diff.cumsum(0)[:strips]
    return (tris * (per_kilo_triangle / 1000.0) + per_row).cpu()


def balanced_bands(costs, world, h, align=TILE_ROWS):
    """Cuts the strips of `costs` (one entry per `align` rows, e.g. tile_row_costs) into `world` contiguous row bands of
    nearly equal summed cost: [(row0,row1)] * world, partitioning [0,h) exactly, every cut on a strip boundary.  With
    more ranks than strips the last ranks get empty bands."""
    costs = [float(c) for c in costs]
    strips = len(costs)
    assert strips == (int(h) + align - 1) // align
    parts = min(int(world), strips)
    pre = [0.0]
    for c in costs:
        pre.append(pre[-1] + c)
    # linear partition: best[k][i] = smallest possible maximum band cost when the first i strips form k non-empty bands
    INF = float("inf")
    best = [[INF] * (strips + 1) for _ in range(parts + 1)]
    cut = [[0] * (strips + 1) for _ in range(parts + 1)]
    best[0][0] = 0.0
    for k in range(1, parts + 1):
        for i in range(k, strips - (parts - k) + 1):
            for j in range(k - 1, i):
                m = max(best[k - 1][j], pre[i] - pre[j])
                if m < best[k][i]:
                    best[k][i], cut[k][i] = m, j
    ends, i = [], strips
    for k in range(parts, 0, -1):
        ends.append(i)
        i = cut[k][i]
    ends = ends[::-1] + [strips] * (int(world) - parts)
    starts = [0] + ends[:-1]
    return [(min(a * align, h), min(b * align, h)) for a, b in zip(starts, ends)]


def rebalance_bands(costs, bands, measured, world, h, align=TILE_ROWS):
    """One step of measured load balancing: the estimated cost of every strip inside band b is scaled by measured[b] /
    estimated(b) (what the band's frame really took against what the strips were thought to cost), and the frame is cut again.
    Returns (new bands, scaled strip costs -- the input of the next step).  Two or three steps settle the bands where the
    heuristic of tile_row_costs is off (its coefficients were fitted to one build of the kernels)."""
    costs = [float(c) for c in costs]
    scaled = list(costs)
    for (r0, r1), t in zip(bands, measured):
        s0, s1 = r0 // align, (r1 + align - 1) // align
        est = sum(costs[s0:s1])
        if est > 0 and t > 0 and s1 > s0:
            for i in range(s0, s1):
                scaled[i] = costs[i] * float(t) / est
    return balanced_bands(scaled, world, h, align), scaled


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


def gather_views_overlapped(produce, n_local, chunk, dst=0, out=None):
    """View-sharded render with the gather running beside it ("chunk the gather so it overlaps rendering").

    `produce(first, count)` renders the local views [first, first+count) and returns them as one contiguous tensor
    [count, ...] (queued on the current stream).  After every chunk an asynchronous gather to rank `dst` is started: the
    collective waits for that chunk only, and the next chunk is rendered while it travels (NCCL runs on its own stream;
    gloo on its worker thread).  Every rank must hold the same n_local.  Returns, on `dst`, a tensor [world, n_local, ...]
    (rank-major: entry [r, i] is local view i of rank r; pass `out` to reuse it) and None elsewhere -- after all
    collectives have completed on the current stream."""
    world = _world()
    rank = dist.get_rank() if world > 1 else 0
    works, first = [], 0
    while first < n_local:
        count = min(int(chunk), n_local - first)
        part = produce(first, count)
        if world == 1:
            if out is None:
                out = torch.empty((1, n_local) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
            out[0, first:first + count].copy_(part)
        else:
            if rank == dst and out is None:
                out = torch.empty((world, n_local) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
            dests = [out[r, first:first + count] for r in range(world)] if rank == dst else None
            works.append((dist.gather(part.contiguous(), dests, dst=dst, async_op=True), part))   # `part` stays alive until done
        first += count
    for w, _ in works:
        w.wait()
    return out if rank == dst else None


def exchange_rows_overlapped(produce, n_local, chunk, out=None):
    """View-sharded render delivered ROW-sharded: after every chunk of views an asynchronous all-to-all is started in which
    rank r receives rows [r*H/N, (r+1)*H/N) of the chunk's views from every rank, and the next chunk is rendered while it
    travels.  Unlike a gather into one rank, whose NVLink ingress (900 GB/s) caps the whole job, every rank here receives only
    1/N of the frames' bytes, so the delivery scales with the number of GPUs -- the layout for a consumer that is itself
    sharded by image region (an encoder or a loss evaluated per row band).  `produce(first, count)` as in
    gather_views_overlapped; the image height must be divisible by the world size.  Returns, on every rank, a tensor
    [world, n_local, H/N, ...] (entry [s, i] = this rank's row band of local view i of rank s; pass `out` to reuse it)."""
    world = _world()
    rank = dist.get_rank() if world > 1 else 0
    works, first = [], 0
    while first < n_local:
        count = min(int(chunk), n_local - first)
        part = produce(first, count)                                  # [count, H, ...]
        H = part.shape[1]
        if H % world:
            raise ValueError(f"image height {H} is not divisible by the world size {world}")
        hb = H // world
        if out is None:
            out = torch.empty((world, n_local, hb) + tuple(part.shape[2:]), dtype=part.dtype, device=part.device)
        if world == 1:
            out[0, first:first + count].copy_(part)
        else:
            # [count, N, hb, ...] -> [N, count, hb, ...]: block d goes to rank d
            send = part.reshape((count, world, hb) + tuple(part.shape[2:])).transpose(0, 1).contiguous()
            recv = torch.empty_like(send)
            w = dist.all_to_all_single(recv, send, async_op=True)
            works.append((w, send, recv, first, count))
        first += count
    for w, _, recv, f0, cnt in works:
        w.wait()
        out[:, f0:f0 + cnt].copy_(recv)
    return out


def gather_bands(local, h, dst=None, align=TILE_ROWS, bands=None):
    """Gathers per-rank row bands [rows_r, W, ...] into the full [h, W, ...] buffer.  `bands` = the [(row0,row1)] list
    every rank used (e.g. balanced_bands); default: the uniform layout of band_shard."""
    world = _world()
    if world == 1:
        return local
    rank = dist.get_rank()
    rows = list(bands) if bands is not None else [band_shard(h, r, world, align) for r in range(world)]
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


class PeerFrame:
    """One h x w frame (z [h,w], colour [h,w,3], normals [h,w,3], float32) in the memory of rank `dst`, mapped into every
    rank of the node.  Rank r builds its filler with `AdvancedPixelBufferFiller(h, w, ..., band=(row0,row1),
    out_ptrs=frame.band_pointers(row0))` and renders; after `frame.complete()` (stream synchronisation + barrier) the
    tensors of `frame.tensors()` on rank `dst` hold the whole frame -- bit-identical to a single-GPU render, with no
    collective and no staging copy.  Collective constructor: every rank of the process group must call it.
    `local_device`: CUDA device of this rank; ranks may share a device (CPU-side tests of the plumbing use gloo)."""

    def __init__(self, h, w, dst=0, local_device=None):
        import ctypes
        from . import _lib
        self._L = _lib.load_library()
        self._check = _lib.check
        self.h, self.w, self.dst = int(h), int(w), int(dst)
        self.world = _world()
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.device = torch.cuda.current_device() if local_device is None else int(local_device)
        pix = self.h * self.w
        self._off_c = (4 * pix + 255) // 256 * 256
        self._off_n = self._off_c + (12 * pix + 255) // 256 * 256
        self.nbytes = self._off_n + 12 * pix
        self.owner = self.rank == self.dst
        ptr = ctypes.c_void_p()
        box = [None]
        if self.owner:
            handle = ctypes.create_string_buffer(64)
            self._check(self._L.crb_shared_alloc(self.device, max(self.nbytes, 256), ctypes.byref(ptr), handle))
            box = [handle.raw]
        if self.world > 1:
            dist.broadcast_object_list(box, src=self.dst)
        if not self.owner:
            self._check(self._L.crb_shared_open(self.device, ctypes.create_string_buffer(box[0], 64), ctypes.byref(ptr)))
        self.base = int(ptr.value)

    def band_pointers(self, row0):
        """Device addresses of row `row0` in the three arrays: what a band filler takes as `out_ptrs`."""
        r = int(row0) * self.w
        return self.base + 4 * r, self.base + self._off_c + 12 * r, self.base + self._off_n + 12 * r

    def tensors(self):
        """(z, colour, normals) torch views of the whole frame (any rank; on other ranks than `dst` they read over NVLink)."""
        from .pixel_buffer_filler import wrap_device_pointer
        dev = torch.device("cuda", self.device)
        return (wrap_device_pointer(torch, self.base, (self.h, self.w), dev),
                wrap_device_pointer(torch, self.base + self._off_c, (self.h, self.w, 3), dev),
                wrap_device_pointer(torch, self.base + self._off_n, (self.h, self.w, 3), dev))

    def complete(self):
        """Every rank's stores have landed in the frame when this returns on all ranks."""
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()

    def close(self):
        """Collective: unmaps on the other ranks first, then the owner frees."""
        if self.base:
            if not self.owner:
                self._check(self._L.crb_shared_close(self.device, self.base))
            if self.world > 1:
                dist.barrier()
            if self.owner:
                self._check(self._L.crb_shared_free(self.device, self.base))
            self.base = 0


class RowExchange:
    """Delivery of a view-sharded batch (config C5) fused into the rasterizer: every rank's k_raster stores the flipped uint8
    image of its views (run.py:26) row band by row band straight into the receive buffers of the ranks that own the bands --
    peer memory mapped through CUDA IPC (crb_shared_*), stores over NVLink / NVSwitch while the frame is rasterized, no NCCL
    call and no staging copy (crb_set_u8_exchange).  Two layouts:
      * all-to-all (default): rank d receives rows [d*H/N, (d+1)*H/N) of ALL N*V views: `tensor()` = [N, V, H/N, W, 3] uint8,
        entry [s, i] = this rank's row band of local view i of rank s.  Every rank takes in 1/N of the bytes, so the
        delivery scales with the number of GPUs (a gather into one rank is capped by that rank's 900 GB/s NVLink ingress);
      * `gather_to=d`: whole images of all views into rank d: `tensor()` on d = [N, V, H, W, 3].
    Use: `f.render_views(v, c, n, views, want=(), u8_exchange=xch.plan(first_view))`, then `xch.complete()`.
    Collective constructor / close (every rank of the process group).  Ranks may share a device (CPU-side plumbing tests)."""

    def __init__(self, n_local_views, h, w, local_device=None, gather_to=None):
        import ctypes
        from . import _lib
        self._L = _lib.load_library()
        self._check = _lib.check
        self.world = _world()
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.device = torch.cuda.current_device() if local_device is None else int(local_device)
        self.V, self.h, self.w = int(n_local_views), int(h), int(w)
        self.gather_to = gather_to
        self.n_bands = 1 if gather_to is not None else self.world
        if self.h % self.n_bands:
            raise ValueError(f"image height {self.h} is not divisible by the {self.n_bands} row bands")
        if self.n_bands > 8:
            raise ValueError("at most 8 row bands (CRB_MAX_EXCHANGE)")
        self.hb = self.h // self.n_bands
        self.view_bytes = self.hb * self.w * 3
        self.receives = gather_to is None or self.rank == gather_to
        self.nbytes = self.world * self.V * self.view_bytes
        ptr = ctypes.c_void_p()
        handle = None
        if self.receives:
            hb_ = ctypes.create_string_buffer(64)
            self._check(self._L.crb_shared_alloc(self.device, max(self.nbytes, 256), ctypes.byref(ptr), hb_))
            handle = hb_.raw
        self.own = int(ptr.value or 0)
        handles = [handle]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle)
        owners = [gather_to] if gather_to is not None else list(range(self.world))
        self.base, self._opened = [], []
        for d in owners:
            if d == self.rank:
                self.base.append(self.own)
            else:
                p = ctypes.c_void_p()
                self._check(self._L.crb_shared_open(self.device, ctypes.create_string_buffer(handles[d], 64), ctypes.byref(p)))
                self.base.append(int(p.value))
                self._opened.append(int(p.value))

    def plan(self, first_view=0):
        """What render_views takes as `u8_exchange` for a call whose views are this rank's local views [first_view, ...)."""
        off = (self.rank * self.V + int(first_view)) * self.view_bytes
        return self.hb, [b + off for b in self.base]

    def tensor(self):
        """This rank's receive buffer [world, V, rows_per_band, W, 3] uint8 (None on ranks that receive nothing)."""
        if not self.receives:
            return None
        from .pixel_buffer_filler import wrap_device_pointer
        return wrap_device_pointer(torch, self.own, (self.world, self.V, self.hb, self.w, 3), torch.device("cuda", self.device),
                                   dtype=torch.uint8)

    def complete(self):
        """Every rank's stores have landed when this returns on all ranks."""
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()

    def close(self):
        """Collective: unmap the peers' buffers, then free the own one."""
        for p in self._opened:
            self._check(self._L.crb_shared_close(self.device, p))
        self._opened = []
        if self.world > 1:
            dist.barrier()
        if self.own:
            self._check(self._L.crb_shared_free(self.device, self.own))
            self.own = 0


class PeerImage:
    """One h x w run.py:26 image (uint8, rows flipped) in the memory of rank `dst`, assembled by band-sharded fillers (config
    C4): rank r renders its row band with `render_views(..., want=(), u8_exchange=img.plan(band))` and k_raster stores the
    band's image rows straight into rank `dst`'s memory over NVLink (crb_set_u8_exchange with one band).  3 bytes per pixel
    instead of the 28 of a PeerFrame: 201 MB for 8192^2, far below one GPU's NVLink ingress per frame time, so -- unlike the
    float32 frame -- the complete image on ONE rank scales with the number of GPUs.  Collective constructor / close."""

    def __init__(self, h, w, dst=0, local_device=None):
        import ctypes
        from . import _lib
        self._L = _lib.load_library()
        self._check = _lib.check
        self.h, self.w, self.dst = int(h), int(w), int(dst)
        self.world = _world()
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.device = torch.cuda.current_device() if local_device is None else int(local_device)
        self.owner = self.rank == self.dst
        ptr = ctypes.c_void_p()
        box = [None]
        if self.owner:
            handle = ctypes.create_string_buffer(64)
            self._check(self._L.crb_shared_alloc(self.device, max(self.h * self.w * 3, 256), ctypes.byref(ptr), handle))
            box = [handle.raw]
        if self.world > 1:
            dist.broadcast_object_list(box, src=self.dst)
        if not self.owner:
            self._check(self._L.crb_shared_open(self.device, ctypes.create_string_buffer(box[0], 64), ctypes.byref(ptr)))
        self.base = int(ptr.value)

    def plan(self, band):
        """`u8_exchange` argument for the filler that owns image rows [row0, row1) (unflipped): its flipped rows are the
        image's rows [h - row1, h - row0)."""
        row0, row1 = int(band[0]), int(band[1])
        return row1 - row0, [self.base + (self.h - row1) * self.w * 3]

    def tensor(self):
        """The image as a torch uint8 tensor [h, w, 3] (on other ranks than `dst` it reads over NVLink)."""
        from .pixel_buffer_filler import wrap_device_pointer
        return wrap_device_pointer(torch, self.base, (self.h, self.w, 3), torch.device("cuda", self.device), dtype=torch.uint8)

    def complete(self):
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()

    def close(self):
        if self.base:
            if not self.owner:
                self._check(self._L.crb_shared_close(self.device, self.base))
            if self.world > 1:
                dist.barrier()
            if self.owner:
                self._check(self._L.crb_shared_free(self.device, self.base))
            self.base = 0
