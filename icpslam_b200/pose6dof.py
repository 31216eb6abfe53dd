"""Pose6DOF algebra the path uses on the host (reference src/utils/pose6DOF.cpp):
compose :98-105, inverse :117-122, fromEigenMatrix :185-190.  pose7 = [px, py, pz, qw, qx, qy, qz]."""
from __future__ import annotations

import numpy as np


def identity() -> np.ndarray:
    return np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


def _qmul(a, b):
    w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3]
    x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2]
    y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1]
    z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]
    return np.array([w, x, y, z])


def rotation_matrix(q) -> np.ndarray:
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def compose(p1, p2) -> np.ndarray:
    """pos = p1.pos + R(p1) p2.pos; rot = q1 q2; normalize (pose6DOF.cpp:98-105); `operator+`."""
    p1, p2 = np.asarray(p1, np.float64), np.asarray(p2, np.float64)
    q = _qmul(p1[3:], p2[3:])
    q /= np.linalg.norm(q)
    return np.concatenate([p1[:3] + rotation_matrix(p1[3:]) @ p2[:3], q])


def inverse(p) -> np.ndarray:
    """pos = -(q^-1 pos); rot = q^-1 (pose6DOF.cpp:117-122)."""
    p = np.asarray(p, np.float64)
    qi = np.array([p[3], -p[4], -p[5], -p[6]]) / float(p[3:] @ p[3:])
    return np.concatenate([-(rotation_matrix(qi) @ p[:3]), qi])


def from_matrix(T) -> np.ndarray:
    """pos = T[0:3,3]; rot = Quaterniond(T[0:3,0:3]); normalize (pose6DOF.cpp:8-13,185-190)."""
    T = np.asarray(T, np.float64).reshape(4, 4)
    m = T[:3, :3]
    q = np.zeros(4)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1:] = [(m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t]
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    q /= np.linalg.norm(q)
    return np.concatenate([T[:3, 3], q])


def to_matrix(p) -> np.ndarray:
    p = np.asarray(p, np.float64)
    T = np.eye(4)
    T[:3, :3] = rotation_matrix(p[3:])
    T[:3, 3] = p[:3]
    return T
