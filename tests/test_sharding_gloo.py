"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo process groups.  The shards are rendered with the oracle (the
CUDA path needs a GPU); what is under test is the partitioning (view blocks, tile-aligned row bands) and the gathers of
cython3dmodelrenderer_b200.sharding -- the results must equal the unsharded frame bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import bits_equal, load_indexed
from cython3dmodelrenderer_b200 import sharding
from cython3dmodelrenderer_b200 import views as VW


def test_view_shard_partitions_exactly():
    for n in (0, 1, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [sharding.view_shard(n, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == n
            pos = 0
            for first, count in blocks:
                assert first == pos
                pos += count
    assert sharding.view_shard(1024, 3, 8) == (384, 128)


def test_band_shard_partitions_exactly_and_tile_aligned():
    for h in (1, 31, 32, 33, 300, 1024, 8192):
        for world in (1, 2, 3, 4, 8):
            bands = [sharding.band_shard(h, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(a % 32 == 0 or a == h for a, _ in bands)
    assert sharding.band_shard(8192, 1, 8) == (1024, 2048)


def test_balanced_bands_partition_exactly_and_minimise_the_heaviest_band():
    rng = np.random.default_rng(0)
    for h in (1, 31, 64, 300, 1024, 8192):
        strips = (h + 31) // 32
        for world in (1, 2, 3, 8):
            costs = rng.uniform(0.1, 10.0, strips) ** 3
            bands = sharding.balanced_bands(costs, world, h)
            assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(a % 32 == 0 or a == h for a, _ in bands)
            if strips >= world:
                assert all(b > a for a, b in bands)      # nobody idles while there are strips to hand out
            heaviest = max(costs[a // 32:(b + 31) // 32].sum() for a, b in bands)
            uniform = max(costs[a // 32:(b + 31) // 32].sum() for a, b in (sharding.band_shard(h, r, world) for r in range(world)))
            assert heaviest <= uniform * (1 + 1e-12)
    assert sharding.balanced_bands([10, 1, 1, 1, 1, 1, 1, 10], 4, 256) == [(0, 32), (32, 64), (64, 224), (224, 256)]


def test_tile_row_costs_follow_the_triangles():
    from cython3dmodelrenderer_b200 import synthetic
    m = synthetic.uv_sphere(200, 98)
    c = sharding.tile_row_costs(m._vertices_by_triangles, m._normals_by_triangles, 512, 512, 45.0, per_row=0.0,
                                per_kilo_triangle=1000.0)
    drawn = ~(m._normals_by_triangles[..., 2].sum(axis=1) >= 0)
    assert c.shape == (16,) and float(c.sum()) >= drawn.sum()          # every drawn triangle counted in >= 1 strip
    assert float(c[0]) == 0.0 and float(c[8]) > 0.0                      # nothing above the sphere, plenty at the equator


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        m = load_indexed("trex")
        # ---- view sharding: 5 views over `world` ranks (ragged), all_gather and gather-to-0 -----------------------
        h, w, V = 48, 64, 5
        views = VW.orbit_views(V)
        first, count = sharding.view_shard(V, rank, world)
        z = np.zeros((count, h, w), np.float32)
        col = np.zeros((count, h, w, 3), np.float32)
        for i in range(count):
            vk, nk = VW.transform_arrays_host(views[first + i], m._vertices_by_triangles, m._normals_by_triangles)
            f = O.OracleFiller(h, w, fov=45.0)
            f.render_arrays(vk, m._colors_by_triangles, nk)
            z[i], col[i] = f.get_z_buffer(), f.get_color_buffer()
        z_all = sharding.gather_views(torch.from_numpy(z), V)
        col_0 = sharding.gather_views(torch.from_numpy(col), V, dst=0)
        assert z_all.shape == (V, h, w)
        assert (col_0 is None) == (rank != 0)
        # ---- band sharding: 100-row frame, tile-aligned bands -------------------------------------------------------
        H, W = 100, 72
        r0, r1 = sharding.band_shard(H, rank, world)
        full = O.OracleFiller(H, W, fov=45.0)
        full.render_model(m)
        zb = torch.from_numpy(full.get_z_buffer()[r0:r1].copy())     # what a band filler produces (tests/test_gpu_parity)
        nb = torch.from_numpy(full.get_normals_buffer()[r0:r1].copy())
        z_band = sharding.gather_bands(zb, H)
        n_band = sharding.gather_bands(nb, H, dst=0)
        # ---- the same with bands cut by cost (unequal heights) -------------------------------------------------------
        bands = sharding.balanced_bands([1.0, 5.0, 1.0, 0.5], world, H)
        b0, b1 = bands[rank]
        z_bal = sharding.gather_bands(torch.from_numpy(full.get_z_buffer()[b0:b1].copy()), H, bands=bands)
        assert bits_equal(z_bal.numpy(), full.get_z_buffer())
        # ---- the chunked gather that runs beside the producer: 5 local "views" per rank in chunks of 2 ------------------
        made = []
        def produce(first, count):
            made.append((first, count))
            return torch.arange(first, first + count, dtype=torch.float32)[:, None, None].repeat(1, 3, 4) + 100.0 * rank
        ov = sharding.gather_views_overlapped(produce, 5, 2, dst=0)
        assert made == [(0, 2), (2, 2), (4, 1)]
        if rank == 0:
            assert ov.shape == (world, 5, 3, 4)
            for r in range(world):
                for i in range(5):
                    assert bool((ov[r, i] == 100.0 * r + i).all())
        else:
            assert ov is None
        # ---- the row exchange (all-to-all per chunk): every rank ends with its row band of ALL views --------------------
        Hx = 6 * world

        def produce_img(first, count):          # value = 1000 * rank + 10 * view + row
            rows = torch.arange(Hx, dtype=torch.float32)[None, :, None]
            vs = torch.arange(first, first + count, dtype=torch.float32)[:, None, None]
            return (1000.0 * rank + 10.0 * vs + rows).repeat(1, 1, 5)
        ex = sharding.exchange_rows_overlapped(produce_img, 5, 2)
        assert ex.shape == (world, 5, 6, 5)
        for src in range(world):
            for i in range(5):
                for rr_ in range(6):
                    assert bool((ex[src, i, rr_] == 1000.0 * src + 10.0 * i + (6 * rank + rr_)).all())
        if rank == 0:
            np.savez(os.path.join(out_dir, "r0.npz"), z_all=z_all.numpy(), col_0=col_0.numpy(), z_band=z_band.numpy(),
                     n_band=n_band.numpy(), z_full=full.get_z_buffer(), n_full=full.get_normals_buffer())
        else:
            np.savez(os.path.join(out_dir, f"r{rank}.npz"), z_all=z_all.numpy(), z_band=z_band.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_results_equal_unsharded(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    m = load_indexed("trex")
    r0 = np.load(tmp_path / "r0.npz")
    views = VW.orbit_views(5)
    for k in range(5):
        vk, nk = VW.transform_arrays_host(views[k], m._vertices_by_triangles, m._normals_by_triangles)
        f = O.OracleFiller(48, 64, fov=45.0)
        f.render_arrays(vk, m._colors_by_triangles, nk)
        assert bits_equal(r0["z_all"][k], f.get_z_buffer())
        assert bits_equal(r0["col_0"][k], f.get_color_buffer())
    assert bits_equal(r0["z_band"], r0["z_full"]) and bits_equal(r0["n_band"], r0["n_full"])
    for r in range(1, world):
        rr = np.load(tmp_path / f"r{r}.npz")
        assert bits_equal(rr["z_all"], r0["z_all"]) and bits_equal(rr["z_band"], r0["z_full"])


def test_row_exchange_plan_addresses_match_the_kernel_formula():
    """sharding.RowExchange.plan / PeerImage.plan hand crb_set_u8_exchange one base address per row band; k_raster stores image
    row r (after the flip) of the call's view k at base[d] + ((k * hb + r - d * hb) * w + x) * 3 with d = r // hb
    (csrc/crender_b200.cu u8_pixel).  Emulated here on a flat byte array for every (source rank, view, row): each byte of every
    receive buffer is written exactly once, at the [source][view][row][x][c] position `tensor()` exposes.  No GPU needed."""
    import numpy as np
    from cython3dmodelrenderer_b200 import sharding
    world, V, h, w = 4, 3, 16, 5
    hb = h // world
    view_bytes = hb * w * 3
    bufs = [np.zeros(world * V * view_bytes, np.int64) for _ in range(world)]      # receive buffer of rank d, as write counters
    for s in range(world):
        xch = sharding.RowExchange.__new__(sharding.RowExchange)           # the arithmetic only: no allocation, no process group
        xch.rank, xch.V, xch.view_bytes, xch.hb = s, V, view_bytes, hb
        xch.base = [d * 10 ** 9 for d in range(world)]                     # fake device addresses, one per destination rank
        for first in (0, 2):                                               # two calls: views [0, 2) and [2, 3)
            rows_per_band, bases = xch.plan(first)
            assert rows_per_band == hb
            for k in range(V - first if first else 2):
                for r in range(h):
                    d = r // hb
                    off = bases[d] + ((k * hb + (r - d * hb)) * w) * 3 - d * 10 ** 9
                    bufs[d][off:off + w * 3] += 1
                    # the position tensor() gives this row: [source s][view first + k][row r - d * hb]
                    assert off == ((s * V + first + k) * hb + (r - d * hb)) * w * 3
    assert all((b == 1).all() for b in bufs)
    img = sharding.PeerImage.__new__(sharding.PeerImage)
    img.base, img.h, img.w = 4096, 64, 8
    rows, bases = img.plan((32, 48))                                        # unflipped rows [32, 48) are image rows [16, 32) after the flip
    assert rows == 16 and bases == [4096 + 16 * 8 * 3]


def test_rebalance_bands_moves_rows_away_from_the_slow_rank():
    """sharding.rebalance_bands: a band that took longer than its estimate gets fewer strips in the next cut; bands stay a
    partition of the frame on strip boundaries, and equal measurements leave a balanced cut alone."""
    from cython3dmodelrenderer_b200 import sharding
    h, world = 32 * 16, 4
    costs = [1.0] * 16
    bands = sharding.balanced_bands(costs, world, h)
    assert bands == [(0, 128), (128, 256), (256, 384), (384, 512)]
    same, _ = sharding.rebalance_bands(costs, bands, [1.0, 1.0, 1.0, 1.0], world, h)
    assert same == bands
    new, scaled = sharding.rebalance_bands(costs, bands, [1.0, 3.0, 1.0, 1.0], world, h)
    assert new[0][0] == 0 and new[-1][1] == h and all(a[1] == b[0] for a, b in zip(new, new[1:]))
    assert all(r0 % 32 == 0 and r1 % 32 == 0 for r0, r1 in new)
    rows = [r1 - r0 for r0, r1 in new]
    assert scaled[4:8] == [0.75] * 4 and scaled[:4] == [0.25] * 4          # cost x measured / estimated, band by band
    # the slow band's old rows [128, 256) are now shared by more ranks: no new band contains all of them
    assert not any(r0 <= 128 and r1 >= 256 for r0, r1 in new)
    assert sum(rows) == h
