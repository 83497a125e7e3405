"""PeerFrame: band-sharded fillers of several PROCESSES render straight into the frame of rank 0, mapped into
every rank through CUDA IPC (crb_shared_*).  The frame on rank 0 must equal the oracle's single frame bit for bit.  Ranks are
placed on cuda:(rank % device_count), so the test runs on a one-GPU box as well (the mapping then crosses processes, not
GPUs); torch.distributed (gloo) only carries the 64-byte handle and the barriers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import bits_equal, load_indexed, random_scene

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, balanced):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, sharding
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        h, w = 320, 256
        m, m2 = load_indexed("bunny"), random_scene(3, T=3000)
        if balanced:
            bands = sharding.balanced_bands([1.0, 1.0, 6.0, 6.0, 2.0, 1.0, 1.0, 1.0, 1.0, 1.0], world, h)
        else:
            bands = [sharding.band_shard(h, r, world) for r in range(world)]
        frame = sharding.PeerFrame(h, w, dst=0, local_device=dev)
        r0, r1 = bands[rank]
        f = AdvancedPixelBufferFiller(h, w, fov=45.0, device=dev, band=(r0, r1), out_ptrs=frame.band_pointers(r0))
        for rep in range(2):                       # a fresh frame (fused clear through TMA boxes where the layout allows) ...
            f.clear()
            f.render_model(m)
        f.render_model(m2)                         # ... and a second model composited into it (plain stores, z read back)
        frame.complete()
        if rank == 0:
            z, c, n = (t.cpu().numpy() for t in frame.tensors())
            np.savez(os.path.join(out_dir, "frame.npz"), z=z, c=c, n=n)
        del f
        frame.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,balanced", [(2, False), (3, True)])
def test_band_fillers_render_into_rank0_frame(world, balanced, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), balanced), nprocs=world, join=True)
    from oracle import oracle as O
    o = O.OracleFiller(320, 256, fov=45.0)
    o.render_model(load_indexed("bunny"))
    o.render_model(random_scene(3, T=3000))
    got = np.load(tmp_path / "frame.npz")
    assert bits_equal(got["z"], o.get_z_buffer())
    assert bits_equal(got["c"], o.get_color_buffer())
    assert bits_equal(got["n"], o.get_normals_buffer())


def _exchange_worker(rank, world, port, out_dir, gather_to):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, sharding, views as VW
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        h, w, V = 192, 160, 5            # 192 rows = 2 bands of 96 / 3 bands of 64; 160 = 5 tiles of 32 (W % 16 == 0: vector clear)
        m = load_indexed("trex")
        dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
        views = VW.orbit_views(world * V, first=rank, count=V, stride=world)       # view k -> rank k mod N, like bench.py
        f = AdvancedPixelBufferFiller(h, w, fov=45.0, device=dev)
        xch = sharding.RowExchange(V, h, w, local_device=dev, gather_to=gather_to)
        # two calls (3 + 2 views), the first split into launches of 2 views: the receive addresses advance with the view index
        f.render_views(dv, dc, dn, views[:3], want=(), chunk=2, u8_exchange=xch.plan(0))
        f.render_views(dv, dc, dn, views[3:], want=(), chunk=2, u8_exchange=xch.plan(3))
        xch.complete()
        ref = f.render_views(dv, dc, dn, views, want=(), color_u8_out=True, chunk=4)["color_u8"]    # the same images, kept local
        np.save(os.path.join(out_dir, f"ref{rank}.npy"), ref.cpu().numpy())
        t = xch.tensor()
        if t is not None:
            np.save(os.path.join(out_dir, f"got{rank}.npy"), t.cpu().numpy())
        dist.barrier()
        del f
        xch.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,gather_to", [(2, None), (3, None), (2, 1)])
def test_row_exchange_from_inside_the_rasterizer(world, gather_to, tmp_path):
    """sharding.RowExchange / crb_set_u8_exchange: after the exchange rank d holds its row band of every rank's uint8
    images (or, gather_to=d, the whole images), equal to what each rank renders into a local color_u8 array."""
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path), gather_to), nprocs=world, join=True)
    refs = [np.load(tmp_path / f"ref{r}.npy") for r in range(world)]            # [V, h, w, 3] per source rank
    assert all(int((r.sum(axis=-1) > 0).sum()) > 1000 for r in refs)
    if gather_to is None:
        hb = refs[0].shape[1] // world
        for d in range(world):
            got = np.load(tmp_path / f"got{d}.npy")                              # [world, V, hb, w, 3]
            for s in range(world):
                assert np.array_equal(got[s], refs[s][:, d * hb:(d + 1) * hb]), f"band {d} of rank {s}'s views"
    else:
        got = np.load(tmp_path / f"got{gather_to}.npy")                           # [world, V, h, w, 3]
        for s in range(world):
            assert np.array_equal(got[s], refs[s])
        assert not (tmp_path / f"got{1 - gather_to}.npy").exists()


def _image_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, sharding, views as VW
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        h, w = 320, 256
        m = load_indexed("bunny")
        dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
        ident = torch.from_numpy(VW.view_matrix()[None, :]).cuda()
        bands = sharding.balanced_bands([1.0, 1.0, 6.0, 6.0, 2.0, 1.0, 1.0, 1.0, 1.0, 1.0], world, h)
        img = sharding.PeerImage(h, w, dst=0, local_device=dev)
        f = AdvancedPixelBufferFiller(h, w, fov=45.0, device=dev, band=bands[rank])
        f.render_views(dv, dc, dn, ident, want=(), chunk=1, guro_light=[0, 0, 1], u8_exchange=img.plan(bands[rank]))
        img.complete()
        if rank == 0:
            np.save(os.path.join(out_dir, "img.npy"), img.tensor().cpu().numpy())
            full = AdvancedPixelBufferFiller(h, w, fov=45.0, device=dev)
            ref = full.render_views(dv, dc, dn, ident, want=(), chunk=1, guro_light=[0, 0, 1], color_u8_out=True)["color_u8"][0]
            np.save(os.path.join(out_dir, "ref.npy"), ref.cpu().numpy())
        dist.barrier()
        del f
        img.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_fillers_assemble_the_uint8_image_on_rank0(world, tmp_path):
    """sharding.PeerImage: band-sharded fillers store their rows of run.py's (lit, flipped, uint8) image straight into rank 0's
    memory from inside the rasterizer; the assembled image equals a single filler's."""
    mp.spawn(_image_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got, ref = np.load(tmp_path / "img.npy"), np.load(tmp_path / "ref.npy")
    assert int((ref.sum(axis=-1) > 0).sum()) > 5000
    assert np.array_equal(got, ref)
