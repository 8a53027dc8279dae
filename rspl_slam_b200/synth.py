"""Seeded synthetic inputs for the BA path.

There is no dataset or network: every benchmark / parity input is generated here, EuRoC-shaped:
the rectified EuRoC camera (/root/reference/configs/euroc.yaml:7,36), keyframes 0.25 m apart on a
gently curving path starting at the reference's initial pose (/root/reference/src/map_builder.cc:368-371),
points at 1-10 m depth (euroc.yaml:9-10), 85 % stereo / 15 % mono point observations, 70 % stereo
line observations whose endpoints slide along the 3-D segment, N(0, 1 px) noise and 5 % gross
outliers. Constraint arrays are emitted the way /root/reference/src/map.cc:609-707 emits them:
landmark-major in landmark *discovery* order, observers ascending by frame id.

``seed = 20261018 + 1000 * config + instance`` as fixed by SURVEY §8d.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .geometry import R_to_quat, line_from_cartesian, quat_to_R, rotvec_to_R
from .problem import (EUROC_CAMERA, EUROC_IMAGE_WH, F64, I32, U8, FrameBatch, FrameProblem, LocalBatch,
                      LocalProblem)

BASE_SEED = 20261018
#: Twc of the first frame (map_builder.cc:368-371): camera z (forward) = world +y, camera y = world -z
R_WC0 = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]])
T_WC0 = np.array([0.0, 0.0, 1.0])


def config_seed(config: int, instance: int = 0) -> int:
    return BASE_SEED + 1000 * config + instance


def _rz(psi: np.ndarray) -> np.ndarray:
    c, s = np.cos(psi), np.sin(psi)
    R = np.zeros(psi.shape + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1], R[..., 2, 2] = c, -s, s, c, 1.0
    return R


def make_trajectory(rng: np.random.Generator, n_kf: int, step: float = 0.25, max_yaw_deg: float = 5.0,
                    loops: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Returns Rwc (n,3,3), twc (n,3). ``loops`` > 0 bends the path into that many full turns
    (global-BA config)."""
    if loops > 0:
        dpsi = np.full(n_kf, 2 * np.pi * loops / n_kf) + np.deg2rad(rng.uniform(-0.5, 0.5, n_kf))
    else:
        dpsi = np.deg2rad(rng.uniform(-max_yaw_deg, max_yaw_deg, n_kf)) * 0.5
        dpsi = np.convolve(dpsi, np.ones(3) / 3.0, mode="same")
    dpsi[0] = 0.0
    psi = np.cumsum(dpsi)
    Rwc = _rz(psi) @ R_WC0
    fwd = np.stack([-np.sin(psi), np.cos(psi), np.zeros(n_kf)], axis=1)
    twc = T_WC0 + np.concatenate([np.zeros((1, 3)), np.cumsum(step * fwd[1:], axis=0)], axis=0)
    return Rwc, twc


def _project(Xc: np.ndarray, cam: np.ndarray):
    fx, fy, cx, cy, bf = cam
    z = Xc[..., 2]
    u = fx * Xc[..., 0] / z + cx
    v = fy * Xc[..., 1] / z + cy
    return u, v, u - bf / z


def _world_to_cams(Rwc: np.ndarray, twc: np.ndarray, Xw: np.ndarray) -> np.ndarray:
    """(K,3,3),(K,3),(N,3) -> (N,K,3) camera-frame coordinates."""
    d = Xw[:, None, :] - twc[None, :, :]
    return np.einsum("kji,nkj->nki", Rwc, d)


def _sample_in_frusta(rng, n, Rwc, twc, cam, zmin, zmax, wh):
    a = rng.integers(0, len(twc), n)
    u = rng.uniform(0, wh[0], n)
    v = rng.uniform(0, wh[1], n)
    z = rng.uniform(zmin, zmax, n)
    Xc = np.stack([(u - cam[2]) / cam[0] * z, (v - cam[3]) / cam[1] * z, z], axis=1)
    return np.einsum("nij,nj->ni", Rwc[a], Xc) + twc[a]


def _choose_observers(rng, vis: np.ndarray, cap_lo: int, cap_hi: int) -> np.ndarray:
    """vis (N,K) bool -> bool mask with each row capped to a random U{lo..hi} of its visible KFs."""
    n, k = vis.shape
    cap = rng.integers(cap_lo, cap_hi + 1, n)
    score = rng.random((n, k))
    score[~vis] = 2.0
    rank = np.argsort(np.argsort(score, axis=1), axis=1)
    return vis & (rank < cap[:, None])


def _ids_with_gaps(rng, n: int, start: int = 0) -> np.ndarray:
    return (start + np.cumsum(rng.integers(1, 4, n))).astype(I32) if n else np.zeros(0, dtype=I32)


def _line_oplus_batch(L: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Vectorised g2o-style orthonormal update of Pluecker lines. (n,6),(n,4) -> (n,6)."""
    w, d = L[:, :3], L[:, 3:]
    mx, my = np.linalg.norm(d, axis=1), np.linalg.norm(w, axis=1)
    n = np.hypot(mx, my)
    c = np.cross(w, d)
    U = np.stack([w / my[:, None], d / mx[:, None], c / np.linalg.norm(c, axis=1, keepdims=True)], axis=2)
    q = np.concatenate([v[:, :3], np.sqrt(1.0 - np.sum(v[:, :3] ** 2, axis=1))[:, None]], axis=1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    U = U @ quat_to_R(q)
    phi = np.arctan2(mx / n, my / n) + v[:, 3]
    out = np.concatenate([U[:, :, 0] * np.cos(phi)[:, None], U[:, :, 1] * np.sin(phi)[:, None]], axis=1)
    return out / np.linalg.norm(out[:, 3:], axis=1, keepdims=True)


def make_local_problem(seed: int, n_kf: int = 10, n_points: int = 3000, n_lines: int = 300,
                       first_kf_id: int = 0, stereo_point_frac: float = 0.85, stereo_line_frac: float = 0.70,
                       outlier_frac: float = 0.05, pixel_sigma: float = 1.0, loops: int = 0,
                       point_cap=(2, 6), line_cap=(2, 5), extra_fixed: bool = True,
                       trajectory=None, kf_ids=None) -> LocalProblem:
    """One local-BA window (config C1 with the defaults; C3: n_kf=20, n_points=10000, n_lines=1000).
    ``trajectory`` = (Rwc, twc) places the window on a given path (used by make_global_problem)."""
    rng = np.random.default_rng(seed)
    cam = EUROC_CAMERA
    wh = EUROC_IMAGE_WH
    b = cam[4] / cam[0]
    Rwc, twc = trajectory if trajectory is not None else make_trajectory(rng, n_kf, loops=loops)
    pose_id = np.arange(first_kf_id, first_kf_id + n_kf, dtype=I32) if kf_ids is None else np.asarray(kf_ids, dtype=I32)
    assert len(pose_id) == n_kf  # (kf_ids: explicit, ascending keyframe ids of a block that is not one run: loop closures)
    # fixed: KF 0 if in the window, else one extra fixed KF (map.cc:559,593) = the oldest one here
    pose_fixed = np.zeros(n_kf, dtype=U8)
    if first_kf_id == 0 or extra_fixed:
        pose_fixed[0] = 1

    def inside(u, v, z, zmin=0.5):
        return (z > zmin) & (u >= 0) & (u < wh[0]) & (v >= 0) & (v < wh[1])

    # ---- points ----
    Xw = np.zeros((0, 3))
    vis = np.zeros((0, n_kf), dtype=bool)
    while len(Xw) < n_points:
        cand = _sample_in_frusta(rng, max(int(1.5 * n_points), 64), Rwc, twc, cam, 1.0, 10.0, wh)
        Xc = _world_to_cams(Rwc, twc, cand)
        u, v, ur = _project(Xc, cam)
        vv = inside(u, v, Xc[..., 2]) & (ur > 1.0)
        ok = vv.sum(axis=1) >= 2
        Xw = np.concatenate([Xw, cand[ok]])
        vis = np.concatenate([vis, vv[ok]])
    Xw, vis = Xw[:n_points], vis[:n_points]
    obs = _choose_observers(rng, vis, *point_cap)
    pi, ki = np.nonzero(obs)  # row-major: landmark-major, observers ascending
    Xc = np.einsum("nji,nj->ni", Rwc[ki], Xw[pi] - twc[ki])
    u, v, ur = _project(Xc, cam)
    m = np.stack([u, v, ur], axis=1) + rng.normal(0.0, pixel_sigma, (len(pi), 3))
    gross = rng.random(len(pi)) < outlier_frac
    m[gross] += rng.uniform(-40.0, 40.0, (int(gross.sum()), 3))
    stereo = (rng.random(len(pi)) < stereo_point_frac) & (m[:, 2] > 0)
    point_id = _ids_with_gaps(rng, n_points)
    disc = rng.permutation(n_points)  # discovery order of map.cc:570-592
    rank = np.empty(n_points, dtype=np.int64)
    rank[disc] = np.arange(n_points)
    order = np.lexsort((ki, rank[pi]))
    pi, ki, m, stereo = pi[order], ki[order], m[order], stereo[order]
    # a point needs >=1 stereo or >=2 mono observations (map.cc:651)
    n_st = np.bincount(pi[stereo], minlength=n_points)
    n_mo = np.bincount(pi[~stereo], minlength=n_points)
    lonely = (n_st == 0) & (n_mo < 2)
    keep = ~lonely[pi]
    pi, ki, m, stereo = pi[keep], ki[keep], m[keep], stereo[keep]
    used = np.zeros(n_points, dtype=bool)
    used[pi] = True
    remap = np.cumsum(used) - 1
    point_id, Xw_used = point_id[used], Xw[used]
    pi = remap[pi]

    # ---- lines ----
    P1 = np.zeros((0, 3))
    P2 = np.zeros((0, 3))
    lvis = np.zeros((0, n_kf), dtype=bool)
    while len(P1) < n_lines and n_lines > 0:
        nc = max(int(2.5 * n_lines), 64)
        a = _sample_in_frusta(rng, nc, Rwc, twc, cam, 1.5, 8.0, wh)
        dirs = rng.normal(size=(nc, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        bpt = a + dirs * rng.uniform(0.5, 3.0, nc)[:, None]
        ok_all = np.ones((nc, n_kf), dtype=bool)
        for s in (-0.1, 0.0, 1.0, 1.1):
            Q = a + s * (bpt - a)
            Qc = _world_to_cams(Rwc, twc, Q)
            for shift in (0.0, b):
                Qs = Qc.copy()
                Qs[..., 0] -= shift
                uu, vv_, _ = _project(Qs, cam)
                ok_all &= inside(uu, vv_, Qs[..., 2], 0.3)
        ok = ok_all.sum(axis=1) >= 2
        P1, P2, lvis = np.concatenate([P1, a[ok]]), np.concatenate([P2, bpt[ok]]), np.concatenate([lvis, ok_all[ok]])
    P1, P2, lvis = P1[:n_lines], P2[:n_lines], lvis[:n_lines]
    if n_lines > 0:
        lobs = _choose_observers(rng, lvis, *line_cap)
        li, lk = np.nonzero(lobs)
    else:
        li = lk = np.zeros(0, dtype=np.int64)
    nlo = len(li)
    s = np.stack([rng.uniform(-0.1, 0.1, (nlo, 2)), 1.0 + rng.uniform(-0.1, 0.1, (nlo, 2))], axis=2)  # (n, L/R, end)
    lm = np.zeros((nlo, 8))
    for side, shift in ((0, 0.0), (1, b)):
        for end in (0, 1):
            Q = P1[li] + s[:, side, end][:, None] * (P2[li] - P1[li])
            Qc = np.einsum("nji,nj->ni", Rwc[lk], Q - twc[lk])
            Qc[:, 0] -= shift
            uu, vv_, _ = _project(Qc, cam)
            lm[:, 4 * side + 2 * end] = uu
            lm[:, 4 * side + 2 * end + 1] = vv_
    lm += rng.normal(0.0, pixel_sigma, lm.shape)
    lgross = rng.random(nlo) < outlier_frac
    lm[lgross] += rng.uniform(-40.0, 40.0, (int(lgross.sum()), 8))
    lstereo = rng.random(nlo) < stereo_line_frac
    line_id = _ids_with_gaps(rng, n_lines)
    ldisc = rng.permutation(n_lines)
    lrank = np.empty(n_lines, dtype=np.int64)
    lrank[ldisc] = np.arange(n_lines)
    lorder = np.lexsort((lk, lrank[li])) if nlo else np.zeros(0, dtype=np.int64)
    li, lk, lm, lstereo = li[lorder], lk[lorder], lm[lorder], lstereo[lorder]
    l_st = np.bincount(li[lstereo], minlength=n_lines)
    l_mo = np.bincount(li[~lstereo], minlength=n_lines)
    llonely = (l_st == 0) & (l_mo < 2)
    lkeep = ~llonely[li]
    li, lk, lm, lstereo = li[lkeep], lk[lkeep], lm[lkeep], lstereo[lkeep]
    lused = np.zeros(n_lines, dtype=bool)
    lused[li] = True
    lremap = np.cumsum(lused) - 1
    line_id = line_id[lused]
    L_true = line_from_cartesian(P1[lused], (P2 - P1)[lused]) if lused.any() else np.zeros((0, 6))
    li = lremap[li]

    # ---- initial estimates ----
    p0 = twc + rng.normal(0.0, 0.02, twc.shape) * (1 - pose_fixed)[:, None]
    q0 = np.zeros((n_kf, 4))
    for k in range(n_kf):
        Rk = Rwc[k] if pose_fixed[k] else Rwc[k] @ rotvec_to_R(rng.normal(0.0, np.deg2rad(0.5), 3))
        q0[k] = R_to_quat(Rk)
    X0 = Xw_used + rng.normal(0.0, 0.05, Xw_used.shape)
    L0 = _line_oplus_batch(L_true, rng.normal(0.0, 0.01, (len(L_true), 4))) if len(L_true) else L_true

    zc = lambda n: np.zeros(n, dtype=I32)
    one = lambda n: np.ones(n, dtype=U8)
    mo, st = ~stereo, stereo
    lmo, lst = ~lstereo, lstereo
    prob = LocalProblem(
        pose_id=pose_id, pose_p=p0, pose_q=q0, pose_fixed=pose_fixed,
        point_id=point_id, point_p=X0, line_id=line_id, line_L=L0, cams=cam[None, :].copy(),
        mp_id_pose=pose_id[ki[mo]], mp_id_point=point_id[pi[mo]], mp_id_cam=zc(int(mo.sum())),
        mp_kp=m[mo, :2], mp_inlier=one(int(mo.sum())),
        sp_id_pose=pose_id[ki[st]], sp_id_point=point_id[pi[st]], sp_id_cam=zc(int(st.sum())),
        sp_kp=m[st], sp_inlier=one(int(st.sum())),
        ml_id_pose=pose_id[lk[lmo]], ml_id_line=line_id[li[lmo]], ml_id_cam=zc(int(lmo.sum())),
        ml_l2d=lm[lmo, :4], ml_inlier=one(int(lmo.sum())),
        sl_id_pose=pose_id[lk[lst]], sl_id_line=line_id[li[lst]], sl_id_cam=zc(int(lst.sum())),
        sl_l2d=lm[lst], sl_inlier=one(int(lst.sum())),
        truth=dict(Rwc=Rwc, twc=twc, points=Xw_used, lines=L_true, point_gross=gross, seed=seed))
    return prob.normalise()


def make_global_problem(seed: int, n_kf: int = 2000, n_points: int = 1_000_000, n_lines: int = 100_000,
                        loops: int = 3, block: int = 16, closure_every: int = 0) -> LocalProblem:
    """Config C5: ONE problem over a long multi-loop trajectory, keyframe 0 fixed. Built from
    overlapping blocks of ``block`` consecutive keyframes (stride block/2): the landmarks of a block are
    sampled in its frusta and observed by 2-6 (points) / 2-5 (lines) of its keyframes, exactly like a
    local window, so the cost is linear in the problem size; the overlap chains the blocks together.
    Poses are perturbed once, globally (2 cm, 0.5 deg, §8d).
    ``closure_every`` = c > 0 adds LOOP CLOSURES: every c-th block gets a sibling block made of its first half and the
    keyframes at the same place one lap later, whose landmarks are seen from both laps (cross-loop covisibility: the
    reduced camera system is then no longer banded). 0 (the default and the C5 bench): none."""
    rng = np.random.default_rng(seed)
    Rwc, twc = make_trajectory(rng, n_kf, loops=loops)
    stride = max(block // 2, 1)
    starts = list(range(0, max(n_kf - stride, 1), stride))
    lap = n_kf // loops if loops > 0 else 0
    half = max(block // 2, 2)
    closures = [s0 for bi, s0 in enumerate(starts)
                if closure_every > 0 and lap > block and bi % closure_every == 0 and s0 + lap + half <= n_kf]
    nb = len(starts) + len(closures)
    parts = []
    for bi, s0 in enumerate(starts):
        e0 = min(s0 + block, n_kf)
        npb = n_points // nb + (1 if bi < n_points % nb else 0)
        nlb = n_lines // nb + (1 if bi < n_lines % nb else 0)
        parts.append(make_local_problem(seed + 7919 * (bi + 1), n_kf=e0 - s0, n_points=npb, n_lines=nlb, first_kf_id=s0,
                                        trajectory=(Rwc[s0:e0], twc[s0:e0])))
    for ci, s0 in enumerate(closures):
        bi = len(starts) + ci
        ids = np.concatenate([np.arange(s0, s0 + half), np.arange(s0 + lap, s0 + lap + half)])
        npb = n_points // nb + (1 if bi < n_points % nb else 0)
        nlb = n_lines // nb + (1 if bi < n_lines % nb else 0)
        parts.append(make_local_problem(seed + 7919 * (bi + 1), n_kf=len(ids), n_points=npb, n_lines=nlb, kf_ids=ids,
                                        extra_fixed=False, trajectory=(Rwc[ids], twc[ids])))
    pt_space = max(int(p.point_id.max()) + 1 if len(p.point_id) else 1 for p in parts)
    ln_space = max(int(p.line_id.max()) + 1 if len(p.line_id) else 1 for p in parts)
    cat = lambda name: np.concatenate([getattr(p, name) for p in parts])
    cat_ids = lambda name, space: np.concatenate([getattr(p, name).astype(np.int64) + bi * space for bi, p in enumerate(parts)])
    if (nb * pt_space) >= 2**31 or (nb * ln_space) >= 2**31:
        raise ValueError("landmark id space exceeds int32")
    pose_fixed = np.zeros(n_kf, dtype=U8)
    pose_fixed[0] = 1
    p0 = twc + rng.normal(0.0, 0.02, twc.shape) * (1 - pose_fixed)[:, None]
    q0 = np.zeros((n_kf, 4))
    for k in range(n_kf):
        Rk = Rwc[k] if pose_fixed[k] else Rwc[k] @ rotvec_to_R(rng.normal(0.0, np.deg2rad(0.5), 3))
        q0[k] = R_to_quat(Rk)
    kw = dict(pose_id=np.arange(n_kf, dtype=I32), pose_p=p0, pose_q=q0, pose_fixed=pose_fixed,
              point_id=cat_ids("point_id", pt_space), point_p=cat("point_p"),
              line_id=cat_ids("line_id", ln_space), line_L=cat("line_L"), cams=parts[0].cams.copy())
    for pre, lm, key, space in (("mp", "id_point", "kp", pt_space), ("sp", "id_point", "kp", pt_space),
                                ("ml", "id_line", "l2d", ln_space), ("sl", "id_line", "l2d", ln_space)):
        kw[f"{pre}_id_pose"] = cat(f"{pre}_id_pose")
        kw[f"{pre}_{lm}"] = cat_ids(f"{pre}_{lm}", space)
        kw[f"{pre}_id_cam"] = cat(f"{pre}_id_cam")
        kw[f"{pre}_{key}"] = cat(f"{pre}_{key}")
        kw[f"{pre}_inlier"] = cat(f"{pre}_inlier")
    kw["truth"] = dict(Rwc=Rwc, twc=twc, seed=seed)
    return LocalProblem(**kw).normalise()


def make_local_batch(config: int, n_windows: int, first_instance: int = 0, **kw) -> Tuple[LocalBatch, List[LocalProblem]]:
    probs = [make_local_problem(config_seed(config, first_instance + i), **kw) for i in range(n_windows)]
    return LocalBatch.from_problems(probs), probs


def _frame_arrays(seed: int, n_points: int, stereo_frac: float, outlier_frac: float, pixel_sigma: float):
    rng = np.random.default_rng(seed)
    cam = EUROC_CAMERA
    wh = EUROC_IMAGE_WH
    psi = rng.uniform(-np.pi, np.pi)
    Rwc = _rz(np.array(psi)) @ R_WC0 @ rotvec_to_R(rng.normal(0.0, np.deg2rad(2.0), 3))
    twc = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), 1.0 + rng.uniform(-0.5, 0.5)])
    u = rng.uniform(0, wh[0], n_points)
    v = rng.uniform(0, wh[1], n_points)
    z = rng.uniform(1.0, 10.0, n_points)
    Xc = np.stack([(u - cam[2]) / cam[0] * z, (v - cam[3]) / cam[1] * z, z], axis=1)
    Xw = Xc @ Rwc.T + twc
    m = np.stack([u, v, u - cam[4] / z], axis=1) + rng.normal(0.0, pixel_sigma, (n_points, 3))
    gross = rng.random(n_points) < outlier_frac
    m[gross] += rng.uniform(-40.0, 40.0, (int(gross.sum()), 3))
    stereo = (rng.random(n_points) < stereo_frac) & (m[:, 2] > 0)
    # initial pose = truth (+) N(3 cm, 1 deg)
    R0 = Rwc @ rotvec_to_R(rng.normal(0.0, np.deg2rad(1.0), 3))
    t0 = twc + rng.normal(0.0, 0.03, 3)
    return Rwc, twc, R0, t0, Xw, m, stereo, gross


def _frame_line_arrays(seed: int, Rwc: np.ndarray, twc: np.ndarray, n_lines: int, stereo_frac: float = 0.7,
                       outlier_frac: float = 0.05, pixel_sigma: float = 1.0):
    """Fixed world lines seen by the frame (the "+60 lines" of config C2, an extension of the reference's
    pose-only path): segments of 0.5-3 m in front of the camera, g2o::Line3D [w, d] in world coordinates,
    measurements = endpoints slid +-10 % along the segment, projected into the left and right image, + noise.
    Own random stream, so the point data of the frame does not depend on n_lines."""
    rng = np.random.default_rng(seed ^ 0x11E5)
    cam, wh = EUROC_CAMERA, EUROC_IMAGE_WH
    b = cam[4] / cam[0]

    def proj(Pc, shift):
        x = Pc[:, 0] - shift
        return np.stack([cam[0] * x / Pc[:, 2] + cam[2], cam[1] * Pc[:, 1] / Pc[:, 2] + cam[3]], axis=1)

    def inside(uv, z):
        return (z > 0.5) & (uv[:, 0] >= 0) & (uv[:, 0] < wh[0]) & (uv[:, 1] >= 0) & (uv[:, 1] < wh[1])

    P1 = np.zeros((0, 3))
    P2 = np.zeros((0, 3))
    while len(P1) < n_lines:
        nc = max(4 * n_lines, 32)
        u, v, z = rng.uniform(40, wh[0] - 40, nc), rng.uniform(40, wh[1] - 40, nc), rng.uniform(1.5, 8.0, nc)
        a = np.stack([(u - cam[2]) / cam[0] * z, (v - cam[3]) / cam[1] * z, z], axis=1)
        d = rng.normal(size=(nc, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        e = a + d * rng.uniform(0.5, 3.0, nc)[:, None]
        ok = np.ones(nc, dtype=bool)
        for s_ in (-0.1, 0.0, 1.0, 1.1):
            Q = a + s_ * (e - a)
            for shift in (0.0, b):
                ok &= inside(proj(Q, shift), Q[:, 2])
        P1, P2 = np.concatenate([P1, a[ok]]), np.concatenate([P2, e[ok]])
    P1, P2 = P1[:n_lines], P2[:n_lines]
    slide = np.stack([rng.uniform(-0.1, 0.1, (n_lines, 2)), 1.0 + rng.uniform(-0.1, 0.1, (n_lines, 2))], axis=2)  # (n, L/R, end)
    meas = np.zeros((n_lines, 8))
    for side, shift in ((0, 0.0), (1, b)):
        for end in (0, 1):
            Q = P1 + slide[:, side, end][:, None] * (P2 - P1)
            meas[:, 4 * side + 2 * end:4 * side + 2 * end + 2] = proj(Q, shift)
    meas += rng.normal(0.0, pixel_sigma, meas.shape)
    gross = rng.random(n_lines) < outlier_frac
    meas[gross] += rng.uniform(-40.0, 40.0, (int(gross.sum()), 8))
    stereo = rng.random(n_lines) < stereo_frac
    P1w, P2w = P1 @ Rwc.T + twc, P2 @ Rwc.T + twc
    Lw = line_from_cartesian(P1w, P2w - P1w) if n_lines else np.zeros((0, 6))
    return Lw, meas, stereo, gross


def make_frame_problem(seed: int, n_points: int = 400, stereo_frac: float = 1.0, outlier_frac: float = 0.05,
                       pixel_sigma: float = 1.0, n_lines: int = 0) -> FrameProblem:
    """One pose-only frame (config C2: 400 stereo points; n_lines = 60 adds the line extension)."""
    Rwc, twc, R0, t0, Xw, m, stereo, gross = _frame_arrays(seed, n_points, stereo_frac, outlier_frac, pixel_sigma)
    rng = np.random.default_rng(seed ^ 0x5EED)
    ids = _ids_with_gaps(rng, n_points, start=100)
    mo = ~stereo
    lines = {}
    if n_lines > 0:
        Lw, lm, lst, _ = _frame_line_arrays(seed, Rwc, twc, n_lines, outlier_frac=outlier_frac, pixel_sigma=pixel_sigma)
        lid = np.arange(7, 7 + 3 * n_lines, 3, dtype=I32)
        lmo = ~lst
        lines = dict(line_id=lid, line_L=Lw,
                     ml_id_line=lid[lmo], ml_id_cam=np.zeros(int(lmo.sum()), dtype=I32), ml_l2d=lm[lmo, :4],
                     ml_inlier=np.ones(int(lmo.sum()), dtype=U8),
                     sl_id_line=lid[lst], sl_id_cam=np.zeros(int(lst.sum()), dtype=I32), sl_l2d=lm[lst],
                     sl_inlier=np.ones(int(lst.sum()), dtype=U8))
    return FrameProblem(
        **lines,
        pose_p=t0, pose_q=R_to_quat(R0), point_id=ids, point_p=Xw, cams=EUROC_CAMERA[None, :].copy(),
        mp_id_point=ids[mo], mp_id_cam=np.zeros(int(mo.sum()), dtype=I32), mp_kp=m[mo, :2],
        mp_inlier=np.ones(int(mo.sum()), dtype=U8),
        sp_id_point=ids[stereo], sp_id_cam=np.zeros(int(stereo.sum()), dtype=I32), sp_kp=m[stereo],
        sp_inlier=np.ones(int(stereo.sum()), dtype=U8),
        truth=dict(Rwc=Rwc, twc=twc, gross=gross, stereo=stereo, seed=seed)).normalise()


def make_frame_batch(config: int, n_frames: int, first_instance: int = 0, n_points: int = 400,
                     stereo_frac: float = 1.0, outlier_frac: float = 0.05, pixel_sigma: float = 1.0,
                     n_lines: int = 0) -> FrameBatch:
    """n_frames independent pose-only frames as one flat batch (same per-frame content as
    ``make_frame_problem(config_seed(config, first_instance + f), n_lines=n_lines)``)."""
    pose = np.zeros((7, n_frames))
    mm, mx, sm, sx, mb, sb = [], [], [], [], [0], [0]
    lml, lmm, lsl, lsm, lmb, lsb = [], [], [], [], [0], [0]
    for f in range(n_frames):
        Rwc, twc, R0, t0, Xw, m, stereo, _ = _frame_arrays(config_seed(config, first_instance + f), n_points,
                                                           stereo_frac, outlier_frac, pixel_sigma)
        if n_lines > 0:
            Lw, lm, lst, _ = _frame_line_arrays(config_seed(config, first_instance + f), Rwc, twc, n_lines,
                                                outlier_frac=outlier_frac, pixel_sigma=pixel_sigma)
            lml.append(Lw[~lst])
            lmm.append(lm[~lst, :4])
            lsl.append(Lw[lst])
            lsm.append(lm[lst])
            lmb.append(lmb[-1] + int((~lst).sum()))
            lsb.append(lsb[-1] + int(lst.sum()))
        pose[:3, f] = t0
        pose[3:, f] = R_to_quat(R0)
        mo = ~stereo
        mm.append(m[mo, :2])
        mx.append(Xw[mo])
        sm.append(m[stereo])
        sx.append(Xw[stereo])
        mb.append(mb[-1] + int(mo.sum()))
        sb.append(sb[-1] + int(stereo.sum()))
    c = lambda xs, d: np.ascontiguousarray(np.concatenate(xs, axis=0).T if xs else np.zeros((d, 0)), dtype=F64)
    nm, ns = mb[-1], sb[-1]
    return FrameBatch(
        cameras=EUROC_CAMERA[None, :].copy(), pose_twc=np.ascontiguousarray(pose),
        mono_begin=np.asarray(mb, dtype=I32), stereo_begin=np.asarray(sb, dtype=I32),
        mono_meas=c(mm, 2), mono_xw=c(mx, 3), mono_cam=np.zeros(nm, dtype=I32), mono_inlier=np.ones(nm, dtype=U8),
        stereo_meas=c(sm, 3), stereo_xw=c(sx, 3), stereo_cam=np.zeros(ns, dtype=I32),
        stereo_inlier=np.ones(ns, dtype=U8),
        **(dict(mline_begin=np.asarray(lmb, dtype=I32), sline_begin=np.asarray(lsb, dtype=I32),
                mline_lw=c(lml, 6), mline_meas=c(lmm, 4), mline_cam=np.zeros(lmb[-1], dtype=I32),
                mline_inlier=np.ones(lmb[-1], dtype=U8),
                sline_lw=c(lsl, 6), sline_meas=c(lsm, 8), sline_cam=np.zeros(lsb[-1], dtype=I32),
                sline_inlier=np.ones(lsb[-1], dtype=U8)) if n_lines > 0 else {}))


def make_triangulation_batch(seed: int, n_points: int = 4096, n_frames: int = 12, max_obs: int = 6, pixel_sigma: float = 1.0,
                             degenerate_frac: float = 0.05):
    """Inputs of the batched Map::TriangulateMappoint: keyframes 0.25 m apart on a gently curving
    path looking forward (8(d) trajectory), points at 1-10 m depth in front of them, each observed by 0..max_obs
    keyframes with N(0, pixel_sigma) pixel noise. A fraction of the points is made rank-deficient on purpose (all
    observations from one keyframe: parallel rays through one centre). Returns a dict of the C-ABI arrays + `truth`."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, _ = EUROC_CAMERA
    twc = np.zeros((7, n_frames))
    Rs = []
    yaw = 0.0
    pos = np.zeros(3)
    for f in range(n_frames):
        yaw += np.deg2rad(rng.uniform(-5, 5))
        Rwc = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        Rs.append(Rwc)
        twc[:3, f] = pos
        twc[3:, f] = R_to_quat(Rwc)
        pos = pos + Rwc @ np.array([0.0, 0.0, 0.25])
    truth = np.zeros((n_points, 3))
    begin, frames, us, vs = [0], [], [], []
    for i in range(n_points):
        f0 = int(rng.integers(0, n_frames))
        depth = rng.uniform(1.0, 10.0)
        pc = np.array([rng.uniform(-0.6, 0.6) * depth, rng.uniform(-0.4, 0.4) * depth, depth])
        X = Rs[f0] @ pc + twc[:3, f0]
        truth[i] = X
        n_obs = int(rng.integers(0, max_obs + 1))
        if rng.random() < degenerate_frac:
            cand = [f0] * max(n_obs, 2)  # every ray through the same centre and direction: rank 2
        else:
            cand = list(rng.permutation(n_frames)[:n_obs])
        for f in cand:
            xc = Rs[f].T @ (X - twc[:3, f])
            if xc[2] < 0.2:
                continue
            frames.append(int(f))
            us.append(fx * xc[0] / xc[2] + cx + rng.normal(0, pixel_sigma))
            vs.append(fy * xc[1] / xc[2] + cy + rng.normal(0, pixel_sigma))
        begin.append(len(frames))
    return dict(obs_begin=np.asarray(begin, dtype=I32), obs_frame=np.asarray(frames, dtype=I32),
                obs_uv=np.ascontiguousarray(np.stack([np.asarray(us, dtype=F64), np.asarray(vs, dtype=F64)])),
                frame_twc=np.ascontiguousarray(twc), cam5=EUROC_CAMERA.copy(), truth=truth)


def make_mapline_batch(seed: int, n_lines: int = 4096, max_pts: int = 24, outlier_frac: float = 0.15, box: float = 6.0,
                       lines_per_window: int = 300):
    """Inputs of the batched Map::UppdateMapline: optimised lines as g2o::Line3D [w, d] (d not
    unit, as the BA leaves it) with 0..max_pts map points each: most within a few cm of the line, a fraction
    0.1 - 1 m off it (the 0.2 m gate of map.cc:155 splits those), plus references to points of OTHER lines (far away).
    Lines come in local maps of `lines_per_window`: the point array is ordered by map and shuffled inside each (a
    line's points are map points of its own local map). The scene straddles the origin (box [-box / 3, box]^3), so a
    share of the lines has only non-positive main coordinates (the reference's DBL_MIN quirk). Returns a dict of the
    C-ABI arrays + the true segment ends `p1`, `p2`."""
    rng = np.random.default_rng(seed)
    p1 = rng.uniform(-box / 3.0, box, (n_lines, 3))
    v = rng.normal(size=(n_lines, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    p2 = p1 + v * rng.uniform(0.5, 3.0, (n_lines, 1))
    wd = line_from_cartesian(p1, p2 - p1) * rng.uniform(0.5, 2.0, (n_lines, 1))
    cnt = rng.integers(0, max_pts + 1, n_lines)
    cnt[rng.random(n_lines) < 0.03] = 0
    begin = np.zeros(n_lines + 1, dtype=I32)
    np.cumsum(cnt, out=begin[1:])
    n_own = int(begin[-1])
    owner = np.repeat(np.arange(n_lines), cnt)
    t = rng.uniform(0.0, 1.0, (n_own, 1))
    off = rng.normal(size=(n_own, 3))
    off -= (off * v[owner]).sum(1, keepdims=True) * v[owner]  # perpendicular to the line
    off /= np.maximum(np.linalg.norm(off, axis=1, keepdims=True), 1e-12)
    far = rng.random(n_own) < outlier_frac
    radius = np.where(far, rng.uniform(0.1, 1.0, n_own), np.abs(rng.normal(0.0, 0.03, n_own)))
    xyz = p1[owner] + t * (p2 - p1)[owner] + off * radius[:, None]
    win = owner // max(lines_per_window, 1)
    index = np.argsort(win + rng.random(n_own), kind="stable").astype(I32)  # a permutation inside every local map
    pts = np.zeros((n_own, 3))
    pts[index] = xyz
    # every eighth reference points at some other point of the same local map instead
    if n_own:
        n_win = int(win[-1]) + 1
        w_start = np.searchsorted(win, np.arange(n_win), side="left")
        w_len = np.searchsorted(win, np.arange(n_win), side="right") - w_start
        other = w_start[win] + np.minimum((rng.random(n_own) * w_len[win]).astype(np.int64), w_len[win] - 1)
        index = np.where(rng.random(n_own) < 0.125, other, index).astype(I32)
    return dict(line_wd=np.ascontiguousarray(wd.T), pt_begin=begin, pt_index=index,
                point_xyz=np.ascontiguousarray(pts.T), p1=p1, p2=p2)
