"""Indexed UV sphere (the C4 mesh family of SURVEY 8d, here WITH an index buffer for the ingest tests)."""
import numpy as np


def indexed_sphere(n_lon, n_lat, radius=0.5, centre=(0.0, 0.0, 1.5)):
    """-> vertices [V,3] f32, triangles [2*n_lon*(n_lat-1), 3] int32; poles are single vertices of valence n_lon."""
    th = np.linspace(0, np.pi, n_lat + 1)[1:-1]
    ph = np.arange(n_lon) * (2 * np.pi / n_lon)
    ring = np.stack([np.outer(np.sin(th), np.cos(ph)), np.outer(np.sin(th), np.sin(ph)),
                     np.outer(np.cos(th), np.ones(n_lon))], -1).reshape(-1, 3)
    v = (np.concatenate([[[0, 0, 1]], ring, [[0, 0, -1]]]) * radius + np.asarray(centre)).astype(np.float32)
    j = np.arange(n_lon)

    def idx(i, jj):
        return 1 + i * n_lon + (jj % n_lon)
    tris = [np.stack([np.zeros(n_lon, np.int64), idx(0, j), idx(0, j + 1)], 1)]
    for i in range(n_lat - 2):
        tris.append(np.stack([idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)], 1))
        tris.append(np.stack([idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)], 1))
    tris.append(np.stack([np.full(n_lon, len(v) - 1), idx(n_lat - 2, j + 1), idx(n_lat - 2, j)], 1))
    tri = np.ascontiguousarray(np.concatenate(tris), dtype=np.int32)
    assert len(tri) == 2 * n_lon * (n_lat - 1)
    return v, tri
