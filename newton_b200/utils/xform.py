"""Host-side (NumPy, float64) quaternion / transform helpers used by the scene builder.

Conventions follow the reference data model (``newton/_src/core/types.py:57-64``):
quaternions are stored ``xyzw``; a transform is ``[px, py, pz, qx, qy, qz, qw]``.
These run once at model-construction time and are not on the hot path.
"""

from __future__ import annotations

import math

import numpy as np


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def quat_identity():
    return np.array([0.0, 0.0, 0.0, 1.0])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array(
        [
            aw * bx + bw * ax + ay * bz - by * az,
            aw * by + bw * ay + az * bx - bz * ax,
            aw * bz + bw * az + ax * by - bx * ay,
            aw * bw - ax * bx - ay * by - az * bz,
        ]
    )


def quat_inverse(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_rotate(q, v):
    q = np.asarray(q, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    u = q[:3]
    w = q[3]
    return v * (2.0 * w * w - 1.0) + u * (2.0 * (u[0] * v[0] + u[1] * v[1] + u[2] * v[2])) + cross(u, v) * (2.0 * w)


def quat_rotate_inv(q, v):
    return quat_rotate(quat_inverse(q), v)


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n > 0.0:
        axis = axis / n
    s = math.sin(angle * 0.5)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, math.cos(angle * 0.5)])


def quat_rpy(roll, pitch, yaw):
    """Warp ``quat_rpy`` (intrinsic ZYX / extrinsic XYZ as used by URDF)."""
    cy = math.cos(yaw * 0.5)
    sy = math.sin(yaw * 0.5)
    cr = math.cos(roll * 0.5)
    sr = math.sin(roll * 0.5)
    cp = math.cos(pitch * 0.5)
    sp = math.sin(pitch * 0.5)
    w = cy * cr * cp + sy * sr * sp
    x = cy * sr * cp - sy * cr * sp
    y = cy * cr * sp + sy * sr * cp
    z = sy * cr * cp - cy * sr * sp
    return np.array([x, y, z, w])


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def quat_between_vectors(a, b):
    """Shortest-arc rotation taking unit vector ``a`` to unit vector ``b``."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a = a / np.linalg.norm(a)
    b = b / np.linalg.norm(b)
    d = float(np.dot(a, b))
    if d > 1.0 - 1e-12:
        return quat_identity()
    if d < -1.0 + 1e-12:
        ref = np.array([1.0, 0.0, 0.0]) if abs(a[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
        axis = np.cross(a, ref)
        return quat_from_axis_angle(axis, math.pi)
    c = np.cross(a, b)
    q = np.array([c[0], c[1], c[2], 1.0 + d])
    return q / np.linalg.norm(q)


def transform(p=(0.0, 0.0, 0.0), q=(0.0, 0.0, 0.0, 1.0)):
    return np.concatenate([np.asarray(p, dtype=np.float64), np.asarray(q, dtype=np.float64)])


def transform_identity():
    return transform()


def transform_mul(a, b):
    return np.concatenate([quat_rotate(a[3:], b[:3]) + a[:3], quat_mul(a[3:], b[3:])])


def transform_inverse(t):
    qi = quat_inverse(t[3:])
    return np.concatenate([-quat_rotate(qi, t[:3]), qi])


def transform_point(t, p):
    return quat_rotate(t[3:], p) + t[:3]


def transform_vector(t, v):
    return quat_rotate(t[3:], v)
