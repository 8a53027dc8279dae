"""ctypes binding of the CPU oracle (oracle/oracle.h) — TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package (rspl_slam_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class OrcTraceRow(C.Structure):
    _fields_ = [("pass_", C.c_int32), ("iter", C.c_int32), ("trial", C.c_int32), ("accepted", C.c_int32),
                ("chi_before", C.c_double), ("chi_after", C.c_double), ("lambda_", C.c_double), ("rho", C.c_double)]


class OrcStats(C.Structure):
    _fields_ = [("iters", C.c_int32 * 4), ("trials", C.c_int32 * 4), ("edges_linearized", C.c_int64),
                ("edges_evaluated", C.c_int64), ("final_chi2", C.c_double), ("n_trace", C.c_int32),
                ("trace_cap", C.c_int32), ("trace", C.POINTER(OrcTraceRow))]


class OrcConfig(C.Structure):
    _fields_ = [("mono_point", C.c_double), ("stereo_point", C.c_double), ("mono_line", C.c_double),
                ("stereo_line", C.c_double), ("iters_pass1", C.c_int32), ("iters_pass2", C.c_int32),
                ("rounds", C.c_int32), ("iters_round", C.c_int32), ("stereo_bf_float", C.c_int32),
                ("reserved", C.c_int32), ("numeric_delta", C.c_double)]


class OrcLocalProblem(C.Structure):
    _fields_ = [
        ("n_poses", C.c_int32), ("pose_id", c_i32p), ("pose_p", c_f64p), ("pose_q", c_f64p), ("pose_fixed", c_u8p),
        ("n_points", C.c_int32), ("point_id", c_i32p), ("point_p", c_f64p),
        ("n_lines", C.c_int32), ("line_id", c_i32p), ("line_L", c_f64p),
        ("n_cams", C.c_int32), ("cams", c_f64p),
        ("n_mono_pt", C.c_int32), ("mp_id_pose", c_i32p), ("mp_id_point", c_i32p), ("mp_id_cam", c_i32p),
        ("mp_kp", c_f64p), ("mp_inlier", c_u8p),
        ("n_stereo_pt", C.c_int32), ("sp_id_pose", c_i32p), ("sp_id_point", c_i32p), ("sp_id_cam", c_i32p),
        ("sp_kp", c_f64p), ("sp_inlier", c_u8p),
        ("n_mono_ln", C.c_int32), ("ml_id_pose", c_i32p), ("ml_id_line", c_i32p), ("ml_id_cam", c_i32p),
        ("ml_l2d", c_f64p), ("ml_inlier", c_u8p),
        ("n_stereo_ln", C.c_int32), ("sl_id_pose", c_i32p), ("sl_id_line", c_i32p), ("sl_id_cam", c_i32p),
        ("sl_l2d", c_f64p), ("sl_inlier", c_u8p),
    ]


class OrcFrameProblem(C.Structure):
    _fields_ = [
        ("pose_p", c_f64p), ("pose_q", c_f64p),
        ("n_points", C.c_int32), ("point_id", c_i32p), ("point_p", c_f64p),
        ("n_cams", C.c_int32), ("cams", c_f64p),
        ("n_mono_pt", C.c_int32), ("mp_id_point", c_i32p), ("mp_id_cam", c_i32p), ("mp_kp", c_f64p), ("mp_inlier", c_u8p),
        ("n_stereo_pt", C.c_int32), ("sp_id_point", c_i32p), ("sp_id_cam", c_i32p), ("sp_kp", c_f64p), ("sp_inlier", c_u8p),
        # extension: constraints on fixed lines (oracle.h)
        ("n_lines", C.c_int32), ("line_id", c_i32p), ("line_L", c_f64p),
        ("n_mono_ln", C.c_int32), ("ml_id_line", c_i32p), ("ml_id_cam", c_i32p), ("ml_l2d", c_f64p), ("ml_inlier", c_u8p),
        ("n_stereo_ln", C.c_int32), ("sl_id_line", c_i32p), ("sl_id_cam", c_i32p), ("sl_l2d", c_f64p), ("sl_inlier", c_u8p),
    ]


def build(force: bool = False) -> str:
    """Compiles oracle/liboracle.so with the committed Makefile (building the checker is not using it)."""
    src = [os.path.join(_HERE, f) for f in ("oracle.cc", "oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_local_ba.argtypes = [C.POINTER(OrcLocalProblem), C.POINTER(OrcConfig), C.POINTER(OrcStats)]
        L.orc_local_ba.restype = C.c_int
        L.orc_frame_opt.argtypes = [C.POINTER(OrcFrameProblem), C.POINTER(OrcConfig), C.POINTER(OrcStats)]
        L.orc_frame_opt.restype = C.c_int
        L.orc_frame_opt_batch.argtypes = [C.c_int32, C.POINTER(OrcFrameProblem), C.POINTER(OrcConfig),
                                          C.POINTER(OrcStats), c_i32p, C.c_int32]
        L.orc_frame_opt_batch.restype = C.c_int
        L.orc_local_ba_batch.argtypes = [C.c_int32, C.POINTER(OrcLocalProblem), C.POINTER(OrcConfig),
                                         C.POINTER(OrcStats), C.c_int32]
        L.orc_local_ba_batch.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        for name, n in (("orc_pose_from_twc", 3), ("orc_pose_to_twc", 3), ("orc_se3_exp", 2), ("orc_pose_oplus", 3),
                        ("orc_line_oplus", 3), ("orc_line_from_cartesian", 2), ("orc_line_transform", 3)):
            getattr(L, name).argtypes = [c_f64p] * n
            getattr(L, name).restype = None
        L.orc_edge_eval.argtypes = [C.c_int, c_f64p, c_f64p, c_f64p, c_f64p, C.c_int, c_f64p, c_f64p, c_f64p]
        L.orc_edge_eval.restype = C.c_int
        L.orc_huber.argtypes = [C.c_double, C.c_double, c_f64p]
        L.orc_huber.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ct)


def make_config(cfg=None, iters=(10, 5), rounds=4, iters_round=10, stereo_bf_float=1, numeric_delta=0.0) -> OrcConfig:
    mp, sp, ml, sl = (50.0, 75.0, 50.0, 75.0) if cfg is None else (cfg.mono_point, cfg.stereo_point, cfg.mono_line, cfg.stereo_line)
    return OrcConfig(mp, sp, ml, sl, iters[0], iters[1], rounds, iters_round, stereo_bf_float, 0, numeric_delta)


def _local_struct(p) -> OrcLocalProblem:
    return OrcLocalProblem(
        len(p.pose_id), _p(p.pose_id, c_i32p), _p(p.pose_p, c_f64p), _p(p.pose_q, c_f64p), _p(p.pose_fixed, c_u8p),
        len(p.point_id), _p(p.point_id, c_i32p), _p(p.point_p, c_f64p),
        len(p.line_id), _p(p.line_id, c_i32p), _p(p.line_L, c_f64p),
        len(p.cams), _p(p.cams, c_f64p),
        len(p.mp_id_pose), _p(p.mp_id_pose, c_i32p), _p(p.mp_id_point, c_i32p), _p(p.mp_id_cam, c_i32p),
        _p(p.mp_kp, c_f64p), _p(p.mp_inlier, c_u8p),
        len(p.sp_id_pose), _p(p.sp_id_pose, c_i32p), _p(p.sp_id_point, c_i32p), _p(p.sp_id_cam, c_i32p),
        _p(p.sp_kp, c_f64p), _p(p.sp_inlier, c_u8p),
        len(p.ml_id_pose), _p(p.ml_id_pose, c_i32p), _p(p.ml_id_line, c_i32p), _p(p.ml_id_cam, c_i32p),
        _p(p.ml_l2d, c_f64p), _p(p.ml_inlier, c_u8p),
        len(p.sl_id_pose), _p(p.sl_id_pose, c_i32p), _p(p.sl_id_line, c_i32p), _p(p.sl_id_cam, c_i32p),
        _p(p.sl_l2d, c_f64p), _p(p.sl_inlier, c_u8p))


def _frame_struct(p) -> OrcFrameProblem:
    return OrcFrameProblem(
        _p(p.pose_p, c_f64p), _p(p.pose_q, c_f64p), len(p.point_id), _p(p.point_id, c_i32p), _p(p.point_p, c_f64p),
        len(p.cams), _p(p.cams, c_f64p),
        len(p.mp_id_point), _p(p.mp_id_point, c_i32p), _p(p.mp_id_cam, c_i32p), _p(p.mp_kp, c_f64p), _p(p.mp_inlier, c_u8p),
        len(p.sp_id_point), _p(p.sp_id_point, c_i32p), _p(p.sp_id_cam, c_i32p), _p(p.sp_kp, c_f64p), _p(p.sp_inlier, c_u8p),
        len(p.line_id), _p(p.line_id, c_i32p), _p(p.line_L, c_f64p),
        len(p.ml_id_line), _p(p.ml_id_line, c_i32p), _p(p.ml_id_cam, c_i32p), _p(p.ml_l2d, c_f64p), _p(p.ml_inlier, c_u8p),
        len(p.sl_id_line), _p(p.sl_id_line, c_i32p), _p(p.sl_id_cam, c_i32p), _p(p.sl_l2d, c_f64p), _p(p.sl_inlier, c_u8p))


def _stats_dict(s: OrcStats, trace_buf=None) -> dict:
    d = dict(iters=list(s.iters), trials=list(s.trials), edges_linearized=int(s.edges_linearized),
             edges_evaluated=int(s.edges_evaluated), final_chi2=float(s.final_chi2))
    if trace_buf is not None:
        d["trace"] = [dict(pass_=r.pass_, iter=r.iter, trial=r.trial, accepted=r.accepted, chi_before=r.chi_before,
                           chi_after=r.chi_after, lambda_=r.lambda_, rho=r.rho) for r in trace_buf[:s.n_trace]]
    return d


def local_ba(prob, cfg: Optional[OrcConfig] = None, trace: bool = False) -> dict:
    """Runs LocalmapOptimization on ``prob`` IN PLACE (like the reference). Returns stats."""
    cfg = cfg or make_config()
    st = OrcStats()
    buf = None
    if trace:
        buf = (OrcTraceRow * 512)()
        st.trace, st.trace_cap = buf, 512
    s = _local_struct(prob)
    rc = lib().orc_local_ba(C.byref(s), C.byref(cfg), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"orc_local_ba failed: {rc}")
    return _stats_dict(st, buf)


def frame_opt(prob, cfg: Optional[OrcConfig] = None, trace: bool = False) -> dict:
    """Runs FrameOptimization on ``prob`` IN PLACE. Returns stats incl. ``ret`` (#inliers)."""
    cfg = cfg or make_config()
    st = OrcStats()
    buf = None
    if trace:
        buf = (OrcTraceRow * 512)()
        st.trace, st.trace_cap = buf, 512
    s = _frame_struct(prob)
    rc = lib().orc_frame_opt(C.byref(s), C.byref(cfg), C.byref(st))
    if rc < 0:
        raise RuntimeError(f"orc_frame_opt failed: {rc}")
    d = _stats_dict(st, buf)
    d["ret"] = rc
    return d


def frame_opt_batch(probs: Sequence, cfg: Optional[OrcConfig] = None, n_threads: int = 0) -> List[dict]:
    cfg = cfg or make_config()
    n = len(probs)
    arr = (OrcFrameProblem * n)(*[_frame_struct(p) for p in probs])
    st = (OrcStats * n)()
    ret = np.zeros(n, dtype=np.int32)
    lib().orc_frame_opt_batch(n, arr, C.byref(cfg), st, _p(ret, c_i32p), n_threads)
    out = []
    for i in range(n):
        d = _stats_dict(st[i])
        d["ret"] = int(ret[i])
        out.append(d)
    return out


def local_ba_batch(probs: Sequence, cfg: Optional[OrcConfig] = None, n_threads: int = 0) -> List[dict]:
    cfg = cfg or make_config()
    n = len(probs)
    arr = (OrcLocalProblem * n)(*[_local_struct(p) for p in probs])
    st = (OrcStats * n)()
    rc = lib().orc_local_ba_batch(n, arr, C.byref(cfg), st, n_threads)
    if rc != 0:
        raise RuntimeError("orc_local_ba_batch failed")
    return [_stats_dict(st[i]) for i in range(n)]


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---- unit-level helpers ----
def _arr(x, n):
    a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    assert a.size >= n
    return a


def pose_from_twc(p3, q4) -> np.ndarray:
    out = np.zeros(7)
    lib().orc_pose_from_twc(_p(_arr(p3, 3), c_f64p), _p(_arr(q4, 4), c_f64p), _p(out, c_f64p))
    return out


def pose_to_twc(pose7):
    p, q = np.zeros(3), np.zeros(4)
    lib().orc_pose_to_twc(_p(_arr(pose7, 7), c_f64p), _p(p, c_f64p), _p(q, c_f64p))
    return p, q


def se3_exp(u6) -> np.ndarray:
    out = np.zeros(7)
    lib().orc_se3_exp(_p(_arr(u6, 6), c_f64p), _p(out, c_f64p))
    return out


def pose_oplus(pose7, u6) -> np.ndarray:
    out = np.zeros(7)
    lib().orc_pose_oplus(_p(_arr(pose7, 7), c_f64p), _p(_arr(u6, 6), c_f64p), _p(out, c_f64p))
    return out


def line_oplus(L6, v4) -> np.ndarray:
    out = np.zeros(6)
    lib().orc_line_oplus(_p(_arr(L6, 6), c_f64p), _p(_arr(v4, 4), c_f64p), _p(out, c_f64p))
    return out


def line_from_cartesian(pv6) -> np.ndarray:
    out = np.zeros(6)
    lib().orc_line_from_cartesian(_p(_arr(pv6, 6), c_f64p), _p(out, c_f64p))
    return out


def line_transform(pose7, L6) -> np.ndarray:
    out = np.zeros(6)
    lib().orc_line_transform(_p(_arr(pose7, 7), c_f64p), _p(_arr(L6, 6), c_f64p), _p(out, c_f64p))
    return out


_DIMS = (2, 3, 2, 4, 2, 3)


def edge_eval(edge_type: int, pose7, lm, meas, cam5, stereo_bf_float: int = 1):
    """Returns (err[dim], Jl[dim, ld], Jp[dim, 6]) of one edge as g2o would compute them."""
    err, Jl, Jp = np.zeros(4), np.zeros(16), np.zeros(24)
    lm6 = np.zeros(6)
    lm = np.asarray(lm, dtype=np.float64).reshape(-1)
    lm6[:lm.size] = lm
    m8 = np.zeros(8)
    meas = np.asarray(meas, dtype=np.float64).reshape(-1)
    m8[:meas.size] = meas
    dim = lib().orc_edge_eval(edge_type, _p(_arr(pose7, 7), c_f64p), _p(lm6, c_f64p), _p(m8, c_f64p),
                              _p(_arr(cam5, 5), c_f64p), stereo_bf_float, _p(err, c_f64p), _p(Jl, c_f64p), _p(Jp, c_f64p))
    ld = 0 if edge_type >= 4 else (3 if edge_type < 2 else 4)
    return err[:dim].copy(), Jl[:dim * ld].reshape(dim, ld).copy() if ld else np.zeros((dim, 0)), Jp[:dim * 6].reshape(dim, 6).copy()


def triangulate_points(obs_begin, obs_frame, obs_uv, frame_twc, cam5, xyz_init=None):
    """Map::TriangulateMappoint (map.cc:292-339) for a batch of points; same layout as rspl_ba_triangulate_points.
    Returns (xyz [3][n], ok [n] uint8, number triangulated)."""
    obs_begin = np.ascontiguousarray(obs_begin, dtype=np.int32)
    obs_frame = np.ascontiguousarray(obs_frame, dtype=np.int32)
    obs_uv = np.ascontiguousarray(obs_uv, dtype=np.float64)
    frame_twc = np.ascontiguousarray(frame_twc, dtype=np.float64)
    cam5 = np.ascontiguousarray(cam5, dtype=np.float64)
    n = len(obs_begin) - 1
    xyz = np.zeros((3, n)) if xyz_init is None else np.ascontiguousarray(xyz_init, dtype=np.float64).copy()
    ok = np.zeros(n, dtype=np.uint8)
    L = lib()
    c_i32p, c_u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    L.orc_triangulate_points.argtypes = [C.c_int32, c_i32p, c_i32p, c_f64p, C.c_int32, c_f64p, C.c_int32, c_f64p, c_f64p, c_u8p]
    L.orc_triangulate_points.restype = C.c_int
    cnt = L.orc_triangulate_points(n, _p(obs_begin, c_i32p), _p(obs_frame, c_i32p), _p(obs_uv, c_f64p), obs_uv.shape[1],
                                   _p(frame_twc, c_f64p), frame_twc.shape[1], _p(cam5, c_f64p), _p(xyz, c_f64p), _p(ok, c_u8p))
    return xyz, ok, int(cnt)


def line_to_cartesian(wd6) -> np.ndarray:
    """g2o::Line3D::toCartesian: anchor (3), unit direction (3)."""
    out = np.zeros(6)
    L = lib()
    L.orc_line_to_cartesian.argtypes = [c_f64p, c_f64p]
    L.orc_line_to_cartesian.restype = None
    L.orc_line_to_cartesian(_p(_arr(wd6, 6), c_f64p), _p(out, c_f64p))
    return out


def update_maplines(line_wd, pt_begin, pt_index, point_xyz, endpoints_init=None):
    """Map::UppdateMapline (map.cc:121-177) for a batch of lines; same layout as rspl_ba_update_maplines.
    Returns (endpoints [6][n], ok [n] uint8, number refreshed)."""
    line_wd = np.ascontiguousarray(line_wd, dtype=np.float64)
    pt_begin = np.ascontiguousarray(pt_begin, dtype=np.int32)
    pt_index = np.ascontiguousarray(pt_index, dtype=np.int32)
    point_xyz = np.ascontiguousarray(point_xyz, dtype=np.float64)
    n = len(pt_begin) - 1
    ends = np.zeros((6, n)) if endpoints_init is None else np.ascontiguousarray(endpoints_init, dtype=np.float64).copy()
    ok = np.zeros(n, dtype=np.uint8)
    L = lib()
    c_i32p, c_u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    L.orc_update_maplines.argtypes = [C.c_int32, c_f64p, c_i32p, c_i32p, c_f64p, C.c_int32, c_f64p, c_u8p]
    L.orc_update_maplines.restype = C.c_int
    cnt = L.orc_update_maplines(n, _p(line_wd, c_f64p), _p(pt_begin, c_i32p), _p(pt_index, c_i32p), _p(point_xyz, c_f64p),
                                point_xyz.shape[1], _p(ends, c_f64p), _p(ok, c_u8p))
    return ends, ok, int(cnt)


def huber(chi2: float, thr: float) -> np.ndarray:
    out = np.zeros(3)
    lib().orc_huber(chi2, thr, _p(out, c_f64p))
    return out
