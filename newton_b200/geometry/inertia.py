"""Mass / centre of mass / inertia of primitive shapes (host side, runs once at build time).

Same closed forms as the reference (``newton/_src/geometry/inertia.py:78-300`` primitives,
``:570-605`` ``transform_inertia``, ``:608-766`` ``compute_inertia_shape``) so that models built
here carry the same body constants as models built by the reference builder.
"""

from __future__ import annotations

import math

import numpy as np

from ..sim.enums import GeoType
from ..utils.xform import quat_to_matrix


def compute_inertia_sphere(density, r):
    m = density * 4.0 / 3.0 * math.pi * r * r * r
    Ia = 2.0 / 5.0 * m * r * r
    return m, np.zeros(3), np.diag([Ia, Ia, Ia])


def compute_inertia_capsule(density, r, hh):
    h = 2.0 * hh
    ms = density * (4.0 / 3.0) * math.pi * r * r * r
    mc = density * math.pi * r * r * h
    m = ms + mc
    Ia = mc * (0.25 * r * r + (1.0 / 12.0) * h * h) + ms * (0.4 * r * r + 0.375 * r * h + 0.25 * h * h)
    Ib = (mc * 0.5 + ms * 0.4) * r * r
    return m, np.zeros(3), np.diag([Ia, Ia, Ib])


_BARREL_NODES, _BARREL_WEIGHTS = np.polynomial.legendre.leggauss(32)  # reference inertia.py:22


def compute_inertia_cylinder(density, r, hh, barrel_radius=0.0):
    """Solid cylinder along z; ``barrel_radius`` != 0: the side profile is the circular arc of that radius through the end rims
    (reference ``inertia.py:151-192``: 32-point Gauss-Legendre over the disc stack r(z))."""
    if barrel_radius != 0.0:
        if barrel_radius < hh:
            raise ValueError("barrel_radius must be zero or at least half_height")
        z = hh * _BARREL_NODES
        w = hh * _BARREL_WEIGHTS
        end_offset = math.sqrt(barrel_radius * barrel_radius - hh * hh)
        profile_offset = np.sqrt(np.maximum(barrel_radius * barrel_radius - z * z, 0.0))
        rz = r + (hh * hh - z * z) / (profile_offset + end_offset)
        r2 = rz * rz
        r4 = r2 * r2
        m = float(density * np.pi * np.dot(w, r2))
        Ib = float(0.5 * density * np.pi * np.dot(w, r4))
        Ia = float(density * np.pi * np.dot(w, 0.25 * r4 + r2 * z * z))
        return m, np.zeros(3), np.diag([Ia, Ia, Ib])
    h = 2.0 * hh
    m = density * math.pi * r * r * h
    Ia = 1.0 / 12.0 * m * (3.0 * r * r + h * h)
    Ib = 0.5 * m * r * r
    return m, np.zeros(3), np.diag([Ia, Ia, Ib])


def compute_inertia_cone(density, r, hh):
    """Solid cone along z, base at -hh, apex at +hh (reference ``inertia.py:195-223``): COM a quarter height above the base."""
    h = 2.0 * hh
    m = density * math.pi * r * r * h / 3.0
    Ia = 3.0 / 20.0 * m * r * r + 3.0 / 80.0 * m * h * h
    Ib = 3.0 / 10.0 * m * r * r
    return m, np.array([0.0, 0.0, -h / 4.0]), np.diag([Ia, Ia, Ib])


def compute_inertia_ellipsoid(density, rx, ry, rz):
    m = density * (4.0 / 3.0) * math.pi * rx * ry * rz
    return (
        m,
        np.zeros(3),
        np.diag([0.2 * m * (ry * ry + rz * rz), 0.2 * m * (rx * rx + rz * rz), 0.2 * m * (rx * rx + ry * ry)]),
    )


def compute_inertia_box(density, hx, hy, hz):
    m = density * 8.0 * hx * hy * hz
    return (
        m,
        np.zeros(3),
        np.diag(
            [
                1.0 / 3.0 * m * (hy * hy + hz * hz),
                1.0 / 3.0 * m * (hx * hx + hz * hz),
                1.0 / 3.0 * m * (hx * hx + hy * hy),
            ]
        ),
    )


def compute_inertia_shape(geo_type, scale, density, is_solid=True, thickness=0.001):
    """(mass, com, inertia-about-com) of a primitive (reference ``inertia.py:608-766``)."""
    if density == 0.0 or geo_type == GeoType.PLANE:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    fns = {
        GeoType.SPHERE: lambda s: compute_inertia_sphere(density, s[0]),
        GeoType.BOX: lambda s: compute_inertia_box(density, s[0], s[1], s[2]),
        GeoType.CAPSULE: lambda s: compute_inertia_capsule(density, s[0], s[1]),
        GeoType.CYLINDER: lambda s: compute_inertia_cylinder(density, s[0], s[1], s[2] if len(s) > 2 else 0.0),
        GeoType.ELLIPSOID: lambda s: compute_inertia_ellipsoid(density, s[0], s[1], s[2]),
        GeoType.CONE: lambda s: compute_inertia_cone(density, s[0], s[1]),
    }
    if geo_type not in fns:
        raise NotImplementedError(f"inertia of shape type {geo_type} is outside the hot-path scope")
    solid = fns[geo_type](scale)
    if is_solid:
        return solid
    # hollow: outer solid minus an inner solid shrunk by `thickness` in every dimension (inertia.py:634-722)
    limited = {GeoType.SPHERE: ("radius",), GeoType.BOX: ("hx", "hy", "hz"), GeoType.CAPSULE: ("radius", "half_height"),
               GeoType.CYLINDER: ("radius", "half_height"), GeoType.CONE: ("radius", "half_height"),
               GeoType.ELLIPSOID: ("rx", "ry", "rz")}[geo_type]
    thickness = _hollow_thickness(GeoType(geo_type).name.lower(), thickness, [(n, scale[i]) for i, n in enumerate(limited)])
    if thickness == 0.0:
        return 0.0, solid[1], np.zeros((3, 3))
    inner = [x - thickness for x in scale]
    if geo_type == GeoType.CYLINDER:
        inner = [scale[0] - thickness, scale[1] - thickness, scale[2] - thickness if len(scale) > 2 and scale[2] > 0.0 else 0.0]
    hollow = fns[geo_type](inner)
    m_shell = solid[0] - hollow[0]
    if geo_type != GeoType.CONE:
        return m_shell, solid[1], solid[2] - hollow[2]
    # the two cones have different centres of mass: the shell's is their mass-weighted difference, and both tensors move there
    # (parallel axes) before they are subtracted (inertia.py:692-707)
    com_shell = (solid[0] * solid[1] - hollow[0] * hollow[1]) / m_shell

    def about(mass, inertia, com):
        d = com_shell - com
        return inertia + mass * (np.dot(d, d) * np.eye(3) - np.outer(d, d))

    return m_shell, com_shell, about(solid[0], solid[2], solid[1]) - about(hollow[0], hollow[2], hollow[1])


def _hollow_thickness(shape_name, thickness, limits):
    """Wall thickness of a hollow primitive must be a finite real in [0, every listed dimension) (reference ``inertia.py:48-76``);
    zero is allowed (massless shell) with a warning."""
    import numbers
    import warnings

    if isinstance(thickness, bool) or not isinstance(thickness, numbers.Real):
        raise TypeError(f"thickness must be a real scalar for a hollow {shape_name} geom")
    thickness = float(thickness)
    if not math.isfinite(thickness):
        raise ValueError(f"thickness must be finite for a hollow {shape_name} geom; got {thickness}")
    if thickness < 0.0:
        raise ValueError(f"thickness must be >= 0 for a hollow {shape_name} geom; got {thickness}")
    if thickness == 0.0:
        warnings.warn(f"A hollow {shape_name} geom with zero thickness has zero mass and inertia.", stacklevel=3)
        return thickness
    for name, value in limits:
        if thickness >= float(value):
            raise ValueError(f"thickness ({thickness}) must be smaller than {name} ({float(value)}) for a hollow {shape_name} geom")
    return thickness


def transform_inertia(mass, inertia, offset, quat):
    """Rotate by ``quat`` then parallel-axis shift by ``offset`` (reference ``inertia.py:570-605``)."""
    R = quat_to_matrix(quat)
    offset = np.asarray(offset, dtype=np.float64)
    return R @ np.asarray(inertia, dtype=np.float64) @ R.T + mass * (
        np.dot(offset, offset) * np.eye(3) - np.outer(offset, offset)
    )


def compute_shape_radius(geo_type, scale, src=None):
    """Bounding-sphere radius (reference ``geometry/utils.py:73-127``)."""
    s = np.abs(np.asarray(scale, dtype=np.float64))
    if geo_type in (GeoType.CONVEX_MESH, GeoType.MESH) and src is not None:  # bounding sphere of the scaled local AABB (utils.py:86-97)
        v = np.asarray(src.vertices, dtype=np.float64) * np.asarray(scale, dtype=np.float64)
        return float(0.5 * np.linalg.norm(v.max(axis=0) - v.min(axis=0)))
    if geo_type == GeoType.SPHERE:
        return float(s[0])
    if geo_type == GeoType.BOX:
        return float(np.linalg.norm(s))
    if geo_type in (GeoType.CAPSULE, GeoType.CYLINDER, GeoType.CONE):
        return float(s[0] + s[1])
    if geo_type == GeoType.ELLIPSOID:
        return float(max(s))
    if geo_type == GeoType.PLANE:
        if s[0] > 0.0 and s[1] > 0.0:
            return float(np.linalg.norm(s)) * 0.5
        return 1.0e6
    return 10.0
