"""Small numpy geometry helpers for the synthetic generator and the tests.

These are *input-side* utilities (building trajectories, Pluecker lines, perturbations); the
solver itself lives in csrc/. Conventions follow the reference's boundary types
(/root/reference/include/g2o_optimization/types.h): quaternions are stored x,y,z,w (Eigen), a 3-D
line is g2o::Line3D storage [w(3) moment, d(3) direction] with |d| = 1
(/root/reference/src/line_processor.cc:427-441 builds them with Line3D::fromCartesian).
"""
from __future__ import annotations

import numpy as np


def quat_to_R(q: np.ndarray) -> np.ndarray:
    """(..., 4) x,y,z,w -> (..., 3, 3)."""
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def R_to_quat(R: np.ndarray) -> np.ndarray:
    """(3, 3) -> (4,) x,y,z,w with w >= 0."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        q = np.array([(R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s, w])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        v = np.zeros(3)
        v[i] = 0.5 * s
        s = 0.5 / s
        w = (R[k, j] - R[j, k]) * s
        v[j] = (R[j, i] + R[i, j]) * s
        v[k] = (R[k, i] + R[i, k]) * s
        q = np.array([v[0], v[1], v[2], w])
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def rotvec_to_R(r: np.ndarray) -> np.ndarray:
    """Rodrigues, (3,) -> (3, 3)."""
    r = np.asarray(r, dtype=np.float64)
    th = np.linalg.norm(r)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / (th * th) * (K @ K)


def rot_angle(Ra: np.ndarray, Rb: np.ndarray) -> float:
    """Relative rotation angle (rad) between two rotation matrices."""
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def quat_angle(qa: np.ndarray, qb: np.ndarray) -> np.ndarray:
    """Relative rotation angle (rad) between quaternions x,y,z,w (sign-insensitive). (...,4) -> (...)."""
    qa = np.asarray(qa, dtype=np.float64)
    qb = np.asarray(qb, dtype=np.float64)
    qa = qa / np.linalg.norm(qa, axis=-1, keepdims=True)
    qb = qb / np.linalg.norm(qb, axis=-1, keepdims=True)
    ax, ay, az, aw = -qa[..., 0], -qa[..., 1], -qa[..., 2], qa[..., 3]  # conj(qa)
    bx, by, bz, bw = qb[..., 0], qb[..., 1], qb[..., 2], qb[..., 3]
    rw = aw * bw - ax * bx - ay * by - az * bz
    rx = aw * bx + ax * bw + ay * bz - az * by
    ry = aw * by + ay * bw + az * bx - ax * bz
    rz = aw * bz + az * bw + ax * by - ay * bx
    return 2.0 * np.arctan2(np.sqrt(rx * rx + ry * ry + rz * rz), np.abs(rw))


def line_from_cartesian(p: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Pluecker line through p with direction v. (...,3),(...,3) -> (...,6) [w, d], |d| = 1."""
    p = np.asarray(p, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    d = v / np.linalg.norm(v, axis=-1, keepdims=True)
    p = p - d * np.sum(d * p, axis=-1, keepdims=True)
    w = np.cross(p, p + d)
    return np.concatenate([w, d], axis=-1)


def line_oplus(L: np.ndarray, v: np.ndarray) -> np.ndarray:
    """4-DoF orthonormal-representation update of a Pluecker line (one line). (6,),(4,) -> (6,)."""
    L = np.asarray(L, dtype=np.float64)
    w, d = L[:3], L[3:]
    mx, my = np.linalg.norm(d), np.linalg.norm(w)
    n = np.hypot(mx, my)
    W = np.array([[my / n, -mx / n], [mx / n, my / n]])
    c = np.cross(w, d)
    U = np.stack([w / my, d / mx, c / np.linalg.norm(c)], axis=1)
    q = np.array([v[0], v[1], v[2], np.sqrt(1.0 - (v[0] ** 2 + v[1] ** 2 + v[2] ** 2))])
    q /= np.linalg.norm(q)
    U = U @ quat_to_R(q)
    W = W @ np.array([[np.cos(v[3]), -np.sin(v[3])], [np.sin(v[3]), np.cos(v[3])]])
    out = np.concatenate([U[:, 0] * W[0, 0], U[:, 1] * W[1, 0]])
    return out / np.linalg.norm(out[3:])


def line_transform(R: np.ndarray, t: np.ndarray, L: np.ndarray) -> np.ndarray:
    """Rigid transform of a Pluecker line: w' = R w + t x (R d), d' = R d."""
    w, d = L[..., :3], L[..., 3:]
    Rd = d @ R.T
    return np.concatenate([w @ R.T + np.cross(t, Rd), Rd], axis=-1)
