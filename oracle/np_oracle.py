"""Independent numpy restatement of the reference's BA algorithm — TEST INFRASTRUCTURE ONLY.

Second opinion on oracle/oracle.cc (SURVEY.md §7 "two independent restatements agreeing is the only
substitute for the missing g2o"). Written from SURVEY.md §9 and the reference sources
(/root/reference/src/g2o_optimization/g2o_optimization.cc:21-397, edge_project_line.cc:21-42,
edge_project_stereo_line.cc:22-51) with deliberately different machinery than the C++ oracle:

* poses are (R, t) matrices updated with the closed-form Rodrigues exponential (no quaternions),
* every Jacobian — points included — is a central difference quotient (delta = 1e-6) of the residual
  through the manifold update, so it also cross-checks the analytic point Jacobians of §9.3,
* the damped normal equations are assembled as ONE dense matrix over all free vertices and solved
  with numpy's Cholesky (no Schur complement, no block bookkeeping).

It shares only the *specification* with the C++ oracle: edge order, Huber with a float-rounded delta,
lambda initialisation, the accept / reject rule, the stale-error flagging. Small problems only.
"""
from __future__ import annotations

import numpy as np

DELTA = 1e-6


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def quat_to_R(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    if w > 1e-6:
        q = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    else:  # not needed for the near-identity-free test trajectories, kept for completeness
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1)
        v = np.zeros(3)
        v[i] = s / 2
        v[j] = (R[j, i] + R[i, j]) / (2 * s)
        v[k] = (R[k, i] + R[i, k]) / (2 * s)
        q = np.array([v[0], v[1], v[2], (R[k, j] - R[j, k]) / (2 * s)])
    return q if q[3] >= 0 else -q


def se3_exp(u):
    """[omega, upsilon] -> (R, t) with t = V upsilon (g2o SE3Quat::exp, §9.2)."""
    om, up = u[:3], u[3:]
    th = np.linalg.norm(om)
    K = _skew(om)
    if th < 1e-5:
        R = np.eye(3) + K + 0.5 * K @ K
        V = np.eye(3) + 0.5 * K + K @ K / 6
    else:
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * K @ K
    return R, V @ up


def pose_oplus(T, u):
    R, t = T
    dR, dt = se3_exp(u)
    return dR @ R, dR @ t + dt


def line_oplus(L, v):
    w, d = L[:3], L[3:]
    a, b = np.linalg.norm(d), np.linalg.norm(w)
    phi = np.arctan2(a, b)
    U = np.stack([w / b, d / a, np.cross(w, d) / np.linalg.norm(np.cross(w, d))], axis=1)
    q = np.array([v[0], v[1], v[2], np.sqrt(1 - v[:3] @ v[:3])])
    U = U @ quat_to_R(q)
    phi = phi + v[3]
    out = np.concatenate([U[:, 0] * np.cos(phi), U[:, 1] * np.sin(phi)])
    return out / np.linalg.norm(out[3:])


def _project_line(cam, w):
    fx, fy, cx, cy = cam[:4]
    return np.array([fy * w[0], fx * w[1], -fy * cx * w[0] - fx * cy * w[1] + fx * fy * w[2]])


class Edge:
    __slots__ = ("kind", "pose", "lm", "meas", "cam", "info", "delta", "thr", "level", "robust", "err", "Xw", "bf_float")

    def residual(self, T, X):
        R, t = T
        cam = self.cam
        if self.kind in ("mp", "sp", "mo", "so"):
            c = R @ X + t
            u = cam[0] * c[0] / c[2] + cam[2]
            v = cam[1] * c[1] / c[2] + cam[3]
            if self.kind in ("mp", "mo"):
                return self.meas - np.array([u, v])
            bf = float(np.float32(cam[4])) if (self.kind == "sp" and self.bf_float) else cam[4]
            return self.meas - np.array([u, v, u - bf / c[2]])
        w, d = X[:3], X[3:]
        Rd = R @ d
        wl = R @ w + np.cross(t, Rd)
        out = []
        m = self.meas
        l2 = _project_line(cam, wl)
        n = np.hypot(l2[0], l2[1])
        out += [(m[0] * l2[0] + m[1] * l2[1] + l2[2]) / n, (m[2] * l2[0] + m[3] * l2[1] + l2[2]) / n]
        if self.kind == "sl":
            tr = t - np.array([cam[4] / cam[0], 0, 0])
            wr = R @ w + np.cross(tr, Rd)
            r2 = _project_line(cam, wr)
            nr = np.hypot(r2[0], r2[1])
            out += [(m[4] * r2[0] + m[5] * r2[1] + r2[2]) / nr, (m[6] * r2[0] + m[7] * r2[1] + r2[2]) / nr]
        return np.array(out)

    def chi2(self):
        return float(self.info * self.err @ self.err)


def _huber(e, delta):
    d2 = delta * delta
    if e <= d2:
        return e, 1.0
    s = np.sqrt(e)
    return 2 * s * delta - d2, delta / s


class Problem:
    """Vertices: poses (dim 6), then points (3), then lines (4) — g2o's id order (:39-70)."""

    def __init__(self):
        self.poses, self.pose_fixed, self.lms, self.lm_kind, self.edges = [], [], [], [], []

    def lm_oplus(self, i, v):
        return self.lms[i] + v if self.lm_kind[i] == 0 else line_oplus(self.lms[i], v)

    def evaluate(self, active):
        for e in active:
            e.err = e.residual(self.poses[e.pose], self.lms[e.lm] if e.lm >= 0 else e.Xw)

    def robust_chi2(self, active):
        return sum(_huber(e.chi2(), e.delta)[0] if e.robust else e.chi2() for e in active)

    def optimize(self, iters, level, trace):
        active = [e for e in self.edges if e.level == level and not (self.pose_fixed[e.pose] and e.lm < 0)]
        if not active:
            return
        used_p = sorted({e.pose for e in active if not self.pose_fixed[e.pose]})
        used_l = sorted({e.lm for e in active if e.lm >= 0})
        pidx = {p: 6 * k for k, p in enumerate(used_p)}
        base = 6 * len(used_p)
        lidx, off = {}, base
        for l in used_l:
            lidx[l] = off
            off += 3 if self.lm_kind[l] == 0 else 4
        n = off
        lam, ni = 0.0, 2.0
        for it in range(iters):
            self.evaluate(active)
            cur = self.robust_chi2(active)
            H, b = np.zeros((n, n)), np.zeros(n)
            for e in active:
                T = self.poses[e.pose]
                X = self.lms[e.lm] if e.lm >= 0 else e.Xw
                blocks = []
                if not self.pose_fixed[e.pose]:
                    J = np.zeros((len(e.err), 6))
                    for d in range(6):
                        u = np.zeros(6)
                        u[d] = DELTA
                        J[:, d] = (e.residual(pose_oplus(T, u), X) - e.residual(pose_oplus(T, -u), X)) / (2 * DELTA)
                    blocks.append((pidx[e.pose], J))
                if e.lm >= 0:
                    dl = 3 if self.lm_kind[e.lm] == 0 else 4
                    J = np.zeros((len(e.err), dl))
                    for d in range(dl):
                        v = np.zeros(dl)
                        v[d] = DELTA
                        J[:, d] = (e.residual(T, self.lm_oplus(e.lm, v)) - e.residual(T, self.lm_oplus(e.lm, -v))) / (2 * DELTA)
                    blocks.append((lidx[e.lm], J))
                w = _huber(e.chi2(), e.delta)[1] if e.robust else 1.0
                for (i0, Ji) in blocks:
                    b[i0:i0 + Ji.shape[1]] -= w * e.info * Ji.T @ e.err
                    for (j0, Jj) in blocks:
                        H[i0:i0 + Ji.shape[1], j0:j0 + Jj.shape[1]] += w * e.info * Ji.T @ Jj
            if it == 0:
                lam, ni = 1e-5 * np.abs(np.diag(H)).max(), 2.0
            q, rho = 0, 0.0
            while True:
                backup = ([(R.copy(), t.copy()) for R, t in self.poses], [x.copy() for x in self.lms])
                try:
                    Lc = np.linalg.cholesky(H + lam * np.eye(n))
                    x = np.linalg.solve(Lc.T, np.linalg.solve(Lc, b))
                    ok = True
                except np.linalg.LinAlgError:
                    x, ok = np.zeros(n), False
                if ok:
                    for p in used_p:
                        self.poses[p] = pose_oplus(self.poses[p], x[pidx[p]:pidx[p] + 6])
                    for l in used_l:
                        dl = 3 if self.lm_kind[l] == 0 else 4
                        self.lms[l] = self.lm_oplus(l, x[lidx[l]:lidx[l] + dl])
                self.evaluate(active)
                tmp = self.robust_chi2(active) if ok else np.finfo(float).max
                rho = (cur - tmp) / (float(x @ (lam * x + b)) + 1e-3)
                lam_used = lam
                accepted = rho > 0 and np.isfinite(tmp)
                if accepted:
                    lam *= max(1 / 3, min(2 / 3, 1 - (2 * rho - 1) ** 3))
                    ni = 2.0
                else:
                    lam *= ni
                    ni *= 2
                    self.poses, self.lms = backup
                trace.append((level, it, q, int(accepted), cur, tmp, lam_used, rho))
                if accepted:
                    cur = tmp
                if not np.isfinite(lam):
                    break
                q += 1
                if not (rho < 0 and q < 10):
                    break
            if q == 10 or rho == 0 or not np.isfinite(lam):
                break


def _mk_edge(kind, pose, lm, meas, cam, thr, bf_float=True, Xw=None):
    e = Edge()
    e.kind, e.pose, e.lm, e.meas, e.cam = kind, pose, lm, np.asarray(meas, dtype=float), np.asarray(cam, dtype=float)
    e.info = 0.1 if kind in ("ml", "sl") else 1.0
    e.thr, e.delta = thr, float(np.float32(np.sqrt(thr)))
    e.level, e.robust, e.err, e.Xw, e.bf_float = 0, True, np.zeros(len(meas) // (2 if kind in ("ml", "sl") else 1)), Xw, bf_float
    return e


def _T_from_twc(p, q):
    Rwc = quat_to_R(np.asarray(q, dtype=float))
    return Rwc.T, -Rwc.T @ np.asarray(p, dtype=float)


def _twc_from_T(T):
    R, t = T
    return -R.T @ t, R_to_quat(R.T)


def local_ba(p, thr=(50.0, 75.0, 50.0, 75.0), iters=(10, 5)):
    """LocalmapOptimization on a rspl_slam_b200.problem.LocalProblem, IN PLACE. Returns the LM trace."""
    P = Problem()
    pose_of = {int(i): k for k, i in enumerate(p.pose_id)}
    for k in range(len(p.pose_id)):
        P.poses.append(_T_from_twc(p.pose_p[k], p.pose_q[k]))
        P.pose_fixed.append(bool(p.pose_fixed[k]))
    pt_of = {int(i): k for k, i in enumerate(p.point_id)}
    for k in range(len(p.point_id)):
        P.lms.append(p.point_p[k].copy())
        P.lm_kind.append(0)
    ln_of = {int(i): len(P.lms) + k for k, i in enumerate(p.line_id)}
    for k in range(len(p.line_id)):
        P.lms.append(p.line_L[k].copy())
        P.lm_kind.append(1)
    groups = []
    for kind, idp, idl, meas, table, th in (("mp", p.mp_id_pose, p.mp_id_point, p.mp_kp, pt_of, thr[0]),
                                            ("sp", p.sp_id_pose, p.sp_id_point, p.sp_kp, pt_of, thr[1]),
                                            ("ml", p.ml_id_pose, p.ml_id_line, p.ml_l2d, ln_of, thr[2]),
                                            ("sl", p.sl_id_pose, p.sl_id_line, p.sl_l2d, ln_of, thr[3])):
        es = [_mk_edge(kind, pose_of[int(a)], table[int(b)], m, p.cams[0], th) for a, b, m in zip(idp, idl, meas)]
        groups.append(es)
        P.edges += es
    trace = []
    P.optimize(iters[0], 0, trace)
    for e in P.edges:  # :176-206
        depth_ok = True
        if e.kind in ("mp", "sp"):
            R, t = P.poses[e.pose]
            depth_ok = (R @ P.lms[e.lm] + t)[2] > 0
        if e.chi2() > e.thr or not depth_ok:
            e.level = 1
        e.robust = False
    P.optimize(iters[1], 0, trace)
    for es, out in zip(groups, (p.mp_inlier, p.sp_inlier, p.ml_inlier, p.sl_inlier)):  # :213-231
        for k, e in enumerate(es):
            ok = e.chi2() <= e.thr
            if e.kind in ("mp", "sp"):
                R, t = P.poses[e.pose]
                ok = ok and (R @ P.lms[e.lm] + t)[2] > 0
            out[k] = 1 if ok else 0
    for k in range(len(p.pose_id)):
        p.pose_p[k], p.pose_q[k] = _twc_from_T(P.poses[k])
    for k in range(len(p.point_id)):
        p.point_p[k] = P.lms[k]
    for k in range(len(p.line_id)):
        p.line_L[k] = P.lms[len(p.point_id) + k]
    return trace


def frame_opt(p, thr=(50.0, 75.0, 50.0, 75.0), rounds=4, iters=10):
    """FrameOptimization on a rspl_slam_b200.problem.FrameProblem, IN PLACE. Returns (ret, trace).
    Line constraints on fixed lines (the extension of oracle.h; none in the reference) are edges whose
    landmark is a constant, like the world point of the pose-only point edges."""
    P = Problem()
    T0 = _T_from_twc(p.pose_p, p.pose_q)
    P.poses, P.pose_fixed = [T0], [False]
    xw = {int(i): p.point_p[k] for k, i in enumerate(p.point_id)}
    mono = [_mk_edge("mo", 0, -1, m, p.cams[0], thr[0], Xw=xw[int(i)]) for i, m in zip(p.mp_id_point, p.mp_kp)]
    stereo = [_mk_edge("so", 0, -1, m, p.cams[0], thr[1], Xw=xw[int(i)]) for i, m in zip(p.sp_id_point, p.sp_kp)]
    lw = {int(i): p.line_L[k] for k, i in enumerate(p.line_id)}
    mline = [_mk_edge("ml", 0, -1, m, p.cams[0], thr[2], Xw=lw[int(i)]) for i, m in zip(p.ml_id_line, p.ml_l2d)]
    sline = [_mk_edge("sl", 0, -1, m, p.cams[0], thr[3], Xw=lw[int(i)]) for i, m in zip(p.sl_id_line, p.sl_l2d)]
    P.edges = mono + stereo + mline + sline
    trace, n_out = [], 0
    for rnd in range(rounds):
        P.poses[0] = (T0[0].copy(), T0[1].copy())
        P.optimize(iters, 0, trace)
        n_out = 0
        for es, inl in ((mono, p.mp_inlier), (stereo, p.sp_inlier), (mline, p.ml_inlier), (sline, p.sl_inlier)):
            for k, e in enumerate(es):
                if not inl[k]:
                    e.err = e.residual(P.poses[0], e.Xw)
                if float(np.float32(e.chi2())) > e.thr:
                    inl[k], e.level = 0, 1
                    n_out += 1
                else:
                    inl[k], e.level = 1, 0
                if rnd == 2:
                    e.robust = False
        if len(P.edges) < 10:
            break
    p.pose_p[:], p.pose_q[:] = _twc_from_T(P.poses[0])
    return len(P.edges) - n_out, trace
