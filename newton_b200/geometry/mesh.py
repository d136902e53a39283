"""Host-side triangle / convex-hull mesh container (subset of reference ``newton.Mesh``, ``geometry/types.py``) - input
construction for ``ModelBuilder.add_shape_convex_hull``: the vertices feed the CONVEX_MESH support map of the collide kernel
(``csrc/nb2_convex.cuh``), the triangles only the mass properties computed once at build time."""

from __future__ import annotations

import numpy as np


class Mesh:
    def __init__(self, vertices, indices=None, *, compute_inertia: bool = True, is_solid: bool = True, maxhullvert: int = 64):
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))
        self.indices = None if indices is None else np.ascontiguousarray(np.asarray(indices, dtype=np.int32).reshape(-1))
        self.is_solid = is_solid
        self.maxhullvert = maxhullvert
        self.has_inertia = False
        self.mass, self.com, self.inertia = 1.0, np.zeros(3), np.eye(3)
        if compute_inertia:
            self.mass, self.com, self.inertia, _ = compute_inertia_mesh(1.0, self.vertices, self._triangles())
            self.has_inertia = True

    def _triangles(self) -> np.ndarray:
        """Outward-oriented triangles: the given ``indices``, or the faces of the vertices' convex hull."""
        if self.indices is not None and self.indices.size:
            return self.indices.reshape(-1, 3)
        from scipy.spatial import ConvexHull

        hull = ConvexHull(self.vertices.astype(np.float64))
        tris = hull.simplices.copy()
        centre = self.vertices[hull.vertices].mean(axis=0)
        v = self.vertices.astype(np.float64)
        for t in tris:  # orient every face away from the interior point
            n = np.cross(v[t[1]] - v[t[0]], v[t[2]] - v[t[0]])
            if np.dot(n, v[t[0]] - centre) < 0.0:
                t[1], t[2] = t[2], t[1]
        return tris

    @staticmethod
    def create_sphere(radius: float = 1.0, *, num_latitudes: int = 32, num_longitudes: int = 32, compute_inertia: bool = True) -> "Mesh":
        """UV sphere, vertex order of the reference generator (``utils/mesh.py:1026-1076``)."""
        pts = []
        for i in range(num_latitudes + 1):
            theta = i * np.pi / num_latitudes
            for j in range(num_longitudes + 1):
                phi = j * 2 * np.pi / num_longitudes
                pts.append([np.cos(phi) * np.sin(theta) * radius, np.cos(theta) * radius, np.sin(phi) * np.sin(theta) * radius])
        return Mesh(np.asarray(pts, dtype=np.float32), compute_inertia=compute_inertia)

    @staticmethod
    def create_box(hx: float, hy: float, hz: float) -> "Mesh":
        c = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float32)
        return Mesh(c)


def compute_inertia_mesh(density: float, vertices, triangles):
    """Mass, centre of mass, inertia about the COM and signed volume of a closed triangle mesh, by signed tetrahedra against the
    origin (the solid branch of reference ``geometry/inertia.py:473-567``: volume, first and second moments summed per face)."""
    v = np.asarray(vertices, dtype=np.float64)
    t = np.asarray(triangles, dtype=np.int64).reshape(-1, 3)
    a, b, c = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))  # 6 x signed tetra volume
    volume = vol6.sum() / 6.0
    first = ((a + b + c) * vol6[:, None]).sum(axis=0) / 24.0
    # second moments of a tetrahedron (0, a, b, c): integral x x^T = vol/20 * (sum_i p_i p_i^T + sum_{i<=j} ...) closed form
    s = a + b + c
    second = (np.einsum("i,ij,ik->jk", vol6, a, a) + np.einsum("i,ij,ik->jk", vol6, b, b) + np.einsum("i,ij,ik->jk", vol6, c, c)
              + np.einsum("i,ij,ik->jk", vol6, s, s)) / 120.0
    if volume <= 0.0:
        raise ValueError("mesh has non-positive volume (triangles must be outward oriented and closed)")
    mass = density * volume
    com = first / volume
    cov = density * second - mass * np.outer(com, com)  # covariance about the COM
    inertia = np.trace(cov) * np.eye(3) - cov
    return float(mass), com, inertia, float(volume)
