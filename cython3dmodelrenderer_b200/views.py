"""Camera views for batched rendering (config C5).

The reference has no camera stage: a new view is made by mutating the model on the host
(crender/cy/data_structures/model.py:238-256, `rotate` = vertices @ R.T followed by an O(T) Python normal
recomputation).  Here a view is 16 floats {R (9, row-major), p (3), q (3), pad} applied on the GPU inside the
setup and shading kernels:   v' = R (v - p) + q,   n' = R n,
each component evaluated as ((r0*d0 + r1*d1) + r2*d2) + q in float32 with one rounding per operation.
`transform_arrays_host` restates exactly that arithmetic in NumPy so the same camera-space arrays can be handed
to the reference / the oracle in parity tests.
"""
import numpy as np


def view_matrix(R=None, p=(0.0, 0.0, 0.0), q=None):
    """16 float32: R row-major, pivot p, translation q (defaults to p: rotate about p)."""
    M = np.zeros(16, dtype=np.float32)
    M[:9] = (np.eye(3) if R is None else np.asarray(R, dtype=np.float64)).astype(np.float32).ravel()
    M[9:12] = np.asarray(p, dtype=np.float32)
    M[12:15] = np.asarray(p if q is None else q, dtype=np.float32)
    return M


def rot_y(theta):
    """Rotation about y in the reference's convention (model.py:229-236,246-247: mat_rot_y[::2, ::2] = [[c,s],[-s,c]])."""
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def orbit_views(n_total, first=0, count=None, center=(0.0, 0.0, 1.0), stride=1):
    """Views first, first+stride, ... (count of them) of an n_total-view orbit about the y axis through `center`
    (theta = 2*pi*k/n_total).  stride = world size, first = rank gives every rank an evenly spread sample of the orbit
    (view k -> GPU k mod N): frames of different view angles cost differently, contiguous arcs would not balance."""
    count = (n_total - first + stride - 1) // stride if count is None else count
    out = np.zeros((count, 16), dtype=np.float32)
    for i in range(count):
        out[i] = view_matrix(rot_y(2.0 * np.pi * (first + i * stride) / n_total), p=center)
    return out


def transform_arrays_host(view, v, n):
    """NumPy float32 restatement of the GPU view transform: returns (v', n') for [T,3,3] float32 inputs."""
    M = np.asarray(view, dtype=np.float32).ravel()
    v = np.asarray(v, dtype=np.float32)
    n = np.asarray(n, dtype=np.float32)
    d = [v[..., k] - M[9 + k] for k in range(3)]
    vo = np.empty_like(v)
    no = np.empty_like(n)
    for r in range(3):
        vo[..., r] = ((M[3 * r] * d[0] + M[3 * r + 1] * d[1]) + M[3 * r + 2] * d[2]) + M[12 + r]
        no[..., r] = (M[3 * r] * n[..., 0] + M[3 * r + 1] * n[..., 1]) + M[3 * r + 2] * n[..., 2]
    return vo, no
