"""Flat float64 serialisation of a synthetic problem for tests/shim/shim_driver.cpp (shared by tests/test_shim.py
and bench.py's single-call latency leg)."""
import numpy as np


def dump_problem(kind, p, path, cfg=(50.0, 75.0, 50.0, 75.0)):
    local = kind == 0
    with_lines = kind in (0, 2)
    v = [kind, len(p.pose_id) if local else 1, len(p.point_id), len(p.line_id) if with_lines else 0,
         len(p.mp_id_point), len(p.sp_id_point), len(p.ml_id_line) if with_lines else 0, len(p.sl_id_line) if with_lines else 0]
    v += list(cfg) + list(p.cams[0])
    if local:
        for i in range(len(p.pose_id)):
            v += [p.pose_id[i], p.pose_fixed[i], *p.pose_p[i], *p.pose_q[i]]
    else:
        v += [7, 0, *p.pose_p, *p.pose_q]
    for i in range(len(p.point_id)):
        v += [p.point_id[i], *p.point_p[i]]
    if with_lines:
        for i in range(len(p.line_id)):
            v += [p.line_id[i], *p.line_L[i]]
    pose_of = (lambda a, i: a[i]) if local else (lambda a, i: 7)
    for i in range(len(p.mp_id_point)):
        v += [pose_of(getattr(p, "mp_id_pose", None), i), p.mp_id_point[i], p.mp_inlier[i], *p.mp_kp[i]]
    for i in range(len(p.sp_id_point)):
        v += [pose_of(getattr(p, "sp_id_pose", None), i), p.sp_id_point[i], p.sp_inlier[i], *p.sp_kp[i]]
    if with_lines:
        for i in range(len(p.ml_id_line)):
            v += [p.ml_id_pose[i] if local else 7, p.ml_id_line[i], p.ml_inlier[i], *p.ml_l2d[i]]
        for i in range(len(p.sl_id_line)):
            v += [p.sl_id_pose[i] if local else 7, p.sl_id_line[i], p.sl_inlier[i], *p.sl_l2d[i]]
    np.asarray(v, dtype=np.float64).tofile(path)
