"""Deterministic synthetic shapes for benchmarks and parity tests.

``synth_mesh(n, seed)``: n Fibonacci-lattice points on the unit sphere, triangulated
by their convex hull (2n-4 outward-oriented triangles, uniform size -- the
"well-connected and uniform" regime of the reference's README.md:62-63), then
displaced radially by six random sinusoids and scaled per axis.
"""
import functools

import numpy as np


@functools.lru_cache(maxsize=8)
def _fibonacci_sphere(n):
    from scipy.spatial import ConvexHull

    i = np.arange(n, dtype=np.float64)
    z = 1.0 - (2.0 * i + 1.0) / n
    phi = i * np.pi * (3.0 - np.sqrt(5.0))
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    u = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    hull = ConvexHull(u, qhull_options="Qt")
    F = hull.simplices.astype(np.int32)
    a, b, c = u[F[:, 0]], u[F[:, 1]], u[F[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0
    F[flip] = F[flip][:, [0, 2, 1]]
    F = F[np.lexsort((F[:, 2], F[:, 1], F[:, 0]))]
    u.setflags(write=False)
    F.setflags(write=False)
    return u, F


def synth_params(seed):
    rng = np.random.default_rng(seed)
    f = rng.normal(0.0, 2.0, size=(6, 3))
    a = rng.uniform(-0.08, 0.08, size=6)
    ph = rng.uniform(0.0, 2.0 * np.pi, size=6)
    s = rng.uniform(0.6, 1.0, size=3)
    return f, a, ph, s


def synth_mesh(n, seed, axis_scale=None):
    """Returns (V float32 [n,3], F int32 [2n-4,3])."""
    u, F = _fibonacci_sphere(int(n))
    f, a, ph, s = synth_params(seed)
    if axis_scale is not None:
        s = np.asarray(axis_scale, dtype=np.float64)
    r = 1.0 + (a[None, :] * np.sin(u @ f.T + ph[None, :])).sum(axis=1)
    V = (u * r[:, None]) * s[None, :]
    return np.ascontiguousarray(V, dtype=np.float32), np.array(F, dtype=np.int32, order="C")


def synth_pair(i, n_src=5000, n_tar=5000):
    """Pair i of the cfg4 batch: (srcV, srcF, tarV, tarF); the source re-uses the target's axis scale."""
    s = synth_params(2 * i + 1)[3]
    tarV, tarF = synth_mesh(n_tar, 2 * i + 1)
    srcV, srcF = synth_mesh(n_src, 2 * i, axis_scale=s)
    return srcV, srcF, tarV, tarF


def unique_edges(F):
    """Undirected edge list [e,2] int32 of a triangle mesh (3n-6 edges for a closed genus-0 mesh)."""
    E = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
    E = np.sort(E, axis=1)
    E = np.unique(E, axis=0)
    return np.ascontiguousarray(E, dtype=np.int32)
