"""Host-side problem containers.

Two levels, mirroring the two sides of the drop-in boundary:

* ``LocalProblem`` / ``FrameProblem`` hold what the reference passes to ``LocalmapOptimization`` /
  ``FrameOptimization`` (/root/reference/include/g2o_optimization/g2o_optimization.h:15-22): vertices
  addressed by *id* in ascending (``std::map``) order, constraints carrying ``id_pose`` /
  ``id_point`` / ``id_camera`` and an in/out ``inlier`` flag (types.h:19-174).
* ``LocalBatch`` / ``FrameBatch`` are the flat structure-of-arrays batches of ``include/rspl_ba.h``
  (component-major planes, window-local indices) — what the C++ shim
  (include/rspl_ba/g2o_optimization_shim.hpp) emits and the CUDA library consumes.

``LocalBatch.from_problems`` / ``FrameBatch.from_problems`` do in numpy what the shim does in C++
(id compaction through the ordered maps, plane transposition); ``scatter_back`` writes results
into the reference-style containers the way the shim mutates the caller's maps in place.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

F64 = np.float64
I32 = np.int32
U8 = np.uint8


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=F64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i(a):
    return np.ascontiguousarray(a, dtype=I32)


def _u(a):
    return np.ascontiguousarray(a, dtype=U8)


@dataclass
class OptimizationConfig:
    """include/read_configs.h:50-56; defaults = configs/configs_euroc.yaml:57-61."""
    mono_point: float = 50.0
    stereo_point: float = 75.0
    mono_line: float = 50.0
    stereo_line: float = 75.0
    rate: float = 0.5  # never read by the optimiser


#: EuRoC rectified camera fx, fy, cx, cy, bf (configs/euroc.yaml:7,36)
EUROC_CAMERA = np.array([435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297,
                         47.90639384423901], dtype=F64)
EUROC_IMAGE_WH = (752, 480)


@dataclass
class LocalProblem:
    """Arguments of LocalmapOptimization (g2o_optimization.cc:21-24)."""
    pose_id: np.ndarray      # (NP,) int32 ascending
    pose_p: np.ndarray       # (NP,3) Twc translation
    pose_q: np.ndarray       # (NP,4) Twc quaternion x,y,z,w
    pose_fixed: np.ndarray   # (NP,) uint8
    point_id: np.ndarray     # (NL,)
    point_p: np.ndarray      # (NL,3)
    line_id: np.ndarray      # (NLn,)
    line_L: np.ndarray       # (NLn,6) g2o::Line3D [w,d]
    cams: np.ndarray         # (NC,5)
    mp_id_pose: np.ndarray
    mp_id_point: np.ndarray
    mp_id_cam: np.ndarray
    mp_kp: np.ndarray        # (n,2)
    mp_inlier: np.ndarray    # (n,) uint8
    sp_id_pose: np.ndarray
    sp_id_point: np.ndarray
    sp_id_cam: np.ndarray
    sp_kp: np.ndarray        # (n,3)
    sp_inlier: np.ndarray
    ml_id_pose: np.ndarray
    ml_id_line: np.ndarray
    ml_id_cam: np.ndarray
    ml_l2d: np.ndarray       # (n,4)
    ml_inlier: np.ndarray
    sl_id_pose: np.ndarray
    sl_id_line: np.ndarray
    sl_id_cam: np.ndarray
    sl_l2d: np.ndarray       # (n,8)
    sl_inlier: np.ndarray
    truth: dict = field(default_factory=dict)  # generator ground truth (not part of the boundary)

    def normalise(self) -> "LocalProblem":
        self.pose_id, self.point_id, self.line_id = _i(self.pose_id), _i(self.point_id), _i(self.line_id)
        self.pose_p, self.pose_q = _f(self.pose_p, (-1, 3)), _f(self.pose_q, (-1, 4))
        self.pose_fixed = _u(self.pose_fixed)
        self.point_p, self.line_L = _f(self.point_p, (-1, 3)), _f(self.line_L, (-1, 6))
        self.cams = _f(self.cams, (-1, 5))
        for pre, dim, key in (("mp", 2, "kp"), ("sp", 3, "kp"), ("ml", 4, "l2d"), ("sl", 8, "l2d")):
            for name in ("id_pose", "id_point" if pre in ("mp", "sp") else "id_line", "id_cam"):
                setattr(self, f"{pre}_{name}", _i(getattr(self, f"{pre}_{name}")))
            setattr(self, f"{pre}_{key}", _f(getattr(self, f"{pre}_{key}"), (-1, dim)))
            setattr(self, f"{pre}_inlier", _u(getattr(self, f"{pre}_inlier")))
        return self

    def copy(self) -> "LocalProblem":
        kw = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()}
        return LocalProblem(**kw)

    @property
    def n_edges(self) -> int:
        return len(self.mp_id_pose) + len(self.sp_id_pose) + len(self.ml_id_pose) + len(self.sl_id_pose)


@dataclass
class FrameProblem:
    """Arguments of FrameOptimization (g2o_optimization.cc:256-258); poses.size() == 1."""
    pose_p: np.ndarray       # (3,)
    pose_q: np.ndarray       # (4,) x,y,z,w
    point_id: np.ndarray     # (N,)
    point_p: np.ndarray      # (N,3)
    cams: np.ndarray         # (NC,5)
    mp_id_point: np.ndarray
    mp_id_cam: np.ndarray
    mp_kp: np.ndarray        # (n,2)
    mp_inlier: np.ndarray
    sp_id_point: np.ndarray
    sp_id_cam: np.ndarray
    sp_kp: np.ndarray        # (n,3)
    sp_inlier: np.ndarray
    truth: dict = field(default_factory=dict)
    # Extension (not in the reference, whose FrameOptimization takes no lines, g2o_optimization.cc:284-285):
    # constraints of the frame on FIXED 3-D lines (SURVEY §8a note); empty by default.
    line_id: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=I32))
    line_L: np.ndarray = field(default_factory=lambda: np.zeros((0, 6)))
    ml_id_line: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=I32))
    ml_id_cam: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=I32))
    ml_l2d: np.ndarray = field(default_factory=lambda: np.zeros((0, 4)))
    ml_inlier: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=U8))
    sl_id_line: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=I32))
    sl_id_cam: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=I32))
    sl_l2d: np.ndarray = field(default_factory=lambda: np.zeros((0, 8)))
    sl_inlier: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=U8))

    def normalise(self) -> "FrameProblem":
        self.pose_p, self.pose_q = _f(self.pose_p, (3,)), _f(self.pose_q, (4,))
        self.point_id, self.point_p = _i(self.point_id), _f(self.point_p, (-1, 3))
        self.cams = _f(self.cams, (-1, 5))
        self.mp_id_point, self.mp_id_cam = _i(self.mp_id_point), _i(self.mp_id_cam)
        self.sp_id_point, self.sp_id_cam = _i(self.sp_id_point), _i(self.sp_id_cam)
        self.mp_kp, self.sp_kp = _f(self.mp_kp, (-1, 2)), _f(self.sp_kp, (-1, 3))
        self.mp_inlier, self.sp_inlier = _u(self.mp_inlier), _u(self.sp_inlier)
        self.line_id, self.line_L = _i(self.line_id), _f(self.line_L, (-1, 6))
        self.ml_id_line, self.ml_id_cam, self.sl_id_line, self.sl_id_cam = (_i(self.ml_id_line), _i(self.ml_id_cam),
                                                                            _i(self.sl_id_line), _i(self.sl_id_cam))
        self.ml_l2d, self.sl_l2d = _f(self.ml_l2d, (-1, 4)), _f(self.sl_l2d, (-1, 8))
        self.ml_inlier, self.sl_inlier = _u(self.ml_inlier), _u(self.sl_inlier)
        return self

    @property
    def n_edges(self) -> int:
        return len(self.mp_id_point) + len(self.sp_id_point) + len(self.ml_id_line) + len(self.sl_id_line)

    def copy(self) -> "FrameProblem":
        kw = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()}
        return FrameProblem(**kw)


def _offsets(counts: Sequence[int]) -> np.ndarray:
    o = np.zeros(len(counts) + 1, dtype=I32)
    np.cumsum(np.asarray(counts, dtype=np.int64), out=o[1:])
    return o


def _local_index(ids_sorted: np.ndarray, ids: np.ndarray, what: str) -> np.ndarray:
    """id -> position in the ascending id array (what std::map iteration order gives the shim)."""
    if len(ids) == 0:
        return np.zeros(0, dtype=I32)
    pos = np.searchsorted(ids_sorted, ids)
    pos = np.clip(pos, 0, max(len(ids_sorted) - 1, 0))
    if len(ids_sorted) == 0 or not np.array_equal(ids_sorted[pos], ids):
        raise KeyError(f"constraint references a missing {what} id")
    return pos.astype(I32)


@dataclass
class FrameBatch:
    """RsplFrameBatch (include/rspl_ba.h): SoA planes over all frames."""
    cameras: np.ndarray       # (NC,5)
    pose_twc: np.ndarray      # (7,F)
    mono_begin: np.ndarray    # (F+1,)
    stereo_begin: np.ndarray  # (F+1,)
    mono_meas: np.ndarray     # (2,Nm)
    mono_xw: np.ndarray       # (3,Nm)
    mono_cam: np.ndarray      # (Nm,)
    mono_inlier: np.ndarray   # (Nm,)
    stereo_meas: np.ndarray   # (3,Ns)
    stereo_xw: np.ndarray     # (3,Ns)
    stereo_cam: np.ndarray    # (Ns,)
    stereo_inlier: np.ndarray  # (Ns,)
    # line extension (None: the batch has no line constraints, the reference's case)
    mline_begin: Optional[np.ndarray] = None   # (F+1,)
    sline_begin: Optional[np.ndarray] = None
    mline_lw: Optional[np.ndarray] = None      # (6,Nml) fixed world line [w, d] copied into the edge
    mline_meas: Optional[np.ndarray] = None    # (4,Nml)
    mline_cam: Optional[np.ndarray] = None
    mline_inlier: Optional[np.ndarray] = None
    sline_lw: Optional[np.ndarray] = None      # (6,Nsl)
    sline_meas: Optional[np.ndarray] = None    # (8,Nsl)
    sline_cam: Optional[np.ndarray] = None
    sline_inlier: Optional[np.ndarray] = None

    _LINE_FIELDS = ("mline_begin", "sline_begin", "mline_lw", "mline_meas", "mline_cam", "mline_inlier",
                    "sline_lw", "sline_meas", "sline_cam", "sline_inlier")

    @property
    def n_frames(self) -> int:
        return self.pose_twc.shape[1]

    @property
    def has_lines(self) -> bool:
        return self.mline_begin is not None and (int(self.mline_begin[-1]) + int(self.sline_begin[-1])) > 0

    @property
    def n_edges(self) -> int:
        n = int(self.mono_begin[-1]) + int(self.stereo_begin[-1])
        if self.mline_begin is not None:
            n += int(self.mline_begin[-1]) + int(self.sline_begin[-1])
        return n

    @staticmethod
    def from_problems(probs: List[FrameProblem]) -> "FrameBatch":
        cams = probs[0].cams
        pose = np.stack([np.concatenate([p.pose_p, p.pose_q]) for p in probs], axis=1)
        mb = _offsets([len(p.mp_id_point) for p in probs])
        sb = _offsets([len(p.sp_id_point) for p in probs])
        mxw, sxw = [], []
        for p in probs:
            if not np.array_equal(p.cams, cams):
                raise ValueError("all frames of a batch share one camera list")
            order = np.argsort(p.point_id, kind="stable")
            ids_sorted = p.point_id[order]
            mxw.append(p.point_p[order][_local_index(ids_sorted, p.mp_id_point, "point")])
            sxw.append(p.point_p[order][_local_index(ids_sorted, p.sp_id_point, "point")])
        cat = lambda xs, d: (np.concatenate(xs, axis=0) if xs else np.zeros((0, d)))
        lines = {}
        if any(len(p.ml_id_line) + len(p.sl_id_line) for p in probs):
            mlw, slw = [], []
            for p in probs:
                order = np.argsort(p.line_id, kind="stable")
                ids_sorted = p.line_id[order]
                mlw.append(p.line_L[order][_local_index(ids_sorted, p.ml_id_line, "line")])
                slw.append(p.line_L[order][_local_index(ids_sorted, p.sl_id_line, "line")])
            lines = dict(
                mline_begin=_offsets([len(p.ml_id_line) for p in probs]), sline_begin=_offsets([len(p.sl_id_line) for p in probs]),
                mline_lw=_f(cat(mlw, 6).T), mline_meas=_f(cat([p.ml_l2d for p in probs], 4).T),
                mline_cam=_i(np.concatenate([p.ml_id_cam for p in probs])), mline_inlier=_u(np.concatenate([p.ml_inlier for p in probs])),
                sline_lw=_f(cat(slw, 6).T), sline_meas=_f(cat([p.sl_l2d for p in probs], 8).T),
                sline_cam=_i(np.concatenate([p.sl_id_cam for p in probs])), sline_inlier=_u(np.concatenate([p.sl_inlier for p in probs])))
        return FrameBatch(
            **lines,
            cameras=_f(cams), pose_twc=_f(pose), mono_begin=mb, stereo_begin=sb,
            mono_meas=_f(cat([p.mp_kp for p in probs], 2).T), mono_xw=_f(cat(mxw, 3).T),
            mono_cam=_i(np.concatenate([p.mp_id_cam for p in probs])),
            mono_inlier=_u(np.concatenate([p.mp_inlier for p in probs])),
            stereo_meas=_f(cat([p.sp_kp for p in probs], 3).T), stereo_xw=_f(cat(sxw, 3).T),
            stereo_cam=_i(np.concatenate([p.sp_id_cam for p in probs])),
            stereo_inlier=_u(np.concatenate([p.sp_inlier for p in probs])))

    def frame_problem(self, f: int) -> FrameProblem:
        """Re-expands frame f into the reference-style container (fresh point ids 0..n-1)."""
        m0, m1 = int(self.mono_begin[f]), int(self.mono_begin[f + 1])
        s0, s1 = int(self.stereo_begin[f]), int(self.stereo_begin[f + 1])
        nm, ns = m1 - m0, s1 - s0
        pts = np.concatenate([self.mono_xw[:, m0:m1].T, self.stereo_xw[:, s0:s1].T], axis=0)
        return FrameProblem(
            pose_p=self.pose_twc[:3, f].copy(), pose_q=self.pose_twc[3:, f].copy(),
            point_id=np.arange(nm + ns, dtype=I32), point_p=pts, cams=self.cameras.copy(),
            mp_id_point=np.arange(nm, dtype=I32), mp_id_cam=self.mono_cam[m0:m1].copy(),
            mp_kp=self.mono_meas[:, m0:m1].T.copy(), mp_inlier=self.mono_inlier[m0:m1].copy(),
            sp_id_point=np.arange(nm, nm + ns, dtype=I32), sp_id_cam=self.stereo_cam[s0:s1].copy(),
            sp_kp=self.stereo_meas[:, s0:s1].T.copy(), sp_inlier=self.stereo_inlier[s0:s1].copy(),
            **self._frame_lines(f)).normalise()

    def _frame_lines(self, f: int) -> dict:
        if self.mline_begin is None:
            return {}
        a0, a1 = int(self.mline_begin[f]), int(self.mline_begin[f + 1])
        b0, b1 = int(self.sline_begin[f]), int(self.sline_begin[f + 1])
        na, nb = a1 - a0, b1 - b0
        return dict(
            line_id=np.arange(na + nb, dtype=I32),
            line_L=np.concatenate([self.mline_lw[:, a0:a1].T, self.sline_lw[:, b0:b1].T], axis=0),
            ml_id_line=np.arange(na, dtype=I32), ml_id_cam=self.mline_cam[a0:a1].copy(),
            ml_l2d=self.mline_meas[:, a0:a1].T.copy(), ml_inlier=self.mline_inlier[a0:a1].copy(),
            sl_id_line=np.arange(na, na + nb, dtype=I32), sl_id_cam=self.sline_cam[b0:b1].copy(),
            sl_l2d=self.sline_meas[:, b0:b1].T.copy(), sl_inlier=self.sline_inlier[b0:b1].copy())

    def slice(self, f0: int, f1: int) -> "FrameBatch":
        """Frames [f0, f1) as an independent batch (how a rank takes its shard)."""
        m0, m1 = int(self.mono_begin[f0]), int(self.mono_begin[f1])
        s0, s1 = int(self.stereo_begin[f0]), int(self.stereo_begin[f1])
        return FrameBatch(
            cameras=self.cameras, pose_twc=_f(self.pose_twc[:, f0:f1]),
            mono_begin=_i(self.mono_begin[f0:f1 + 1] - m0), stereo_begin=_i(self.stereo_begin[f0:f1 + 1] - s0),
            mono_meas=_f(self.mono_meas[:, m0:m1]), mono_xw=_f(self.mono_xw[:, m0:m1]),
            mono_cam=_i(self.mono_cam[m0:m1]), mono_inlier=_u(self.mono_inlier[m0:m1]),
            stereo_meas=_f(self.stereo_meas[:, s0:s1]), stereo_xw=_f(self.stereo_xw[:, s0:s1]),
            stereo_cam=_i(self.stereo_cam[s0:s1]), stereo_inlier=_u(self.stereo_inlier[s0:s1]),
            **self._slice_lines(f0, f1))

    def _slice_lines(self, f0: int, f1: int) -> dict:
        if self.mline_begin is None:
            return {}
        a0, a1 = int(self.mline_begin[f0]), int(self.mline_begin[f1])
        b0, b1 = int(self.sline_begin[f0]), int(self.sline_begin[f1])
        return dict(
            mline_begin=_i(self.mline_begin[f0:f1 + 1] - a0), sline_begin=_i(self.sline_begin[f0:f1 + 1] - b0),
            mline_lw=_f(self.mline_lw[:, a0:a1]), mline_meas=_f(self.mline_meas[:, a0:a1]),
            mline_cam=_i(self.mline_cam[a0:a1]), mline_inlier=_u(self.mline_inlier[a0:a1]),
            sline_lw=_f(self.sline_lw[:, b0:b1]), sline_meas=_f(self.sline_meas[:, b0:b1]),
            sline_cam=_i(self.sline_cam[b0:b1]), sline_inlier=_u(self.sline_inlier[b0:b1]))

    def h2d_bytes(self) -> int:
        """bytes the library copies to the device: every array except, with a single camera, the per-edge camera
        indices (validated on the host, never read by the kernels)"""
        single = len(self.cameras) == 1
        return sum(int(getattr(self, k).nbytes) for k in (
            "cameras", "pose_twc", "mono_begin", "stereo_begin", "mono_meas", "mono_xw", "mono_cam",
            "mono_inlier", "stereo_meas", "stereo_xw", "stereo_cam", "stereo_inlier") + self._LINE_FIELDS
            if getattr(self, k) is not None and not (single and k.endswith("_cam")))


@dataclass
class FrameBatchResult:
    pose_twc: np.ndarray       # (7,F)
    mono_inlier: np.ndarray
    stereo_inlier: np.ndarray
    num_inliers: np.ndarray    # (F,)
    stats: np.ndarray          # structured (F,)
    mline_inlier: Optional[np.ndarray] = None  # line extension
    sline_inlier: Optional[np.ndarray] = None

    def d2h_bytes(self) -> int:
        return sum(int(getattr(self, k).nbytes) for k in ("pose_twc", "mono_inlier", "stereo_inlier", "num_inliers", "stats",
                                                          "mline_inlier", "sline_inlier") if getattr(self, k) is not None)


_EDGE_CLASSES = (("mp", "point", 2, "kp"), ("sp", "point", 3, "kp"), ("ml", "line", 4, "l2d"), ("sl", "line", 8, "l2d"))


@dataclass
class LocalBatch:
    """RsplLocalBatch (include/rspl_ba.h)."""
    cameras: np.ndarray
    pose_begin: np.ndarray
    point_begin: np.ndarray
    line_begin: np.ndarray
    mono_pt_begin: np.ndarray
    stereo_pt_begin: np.ndarray
    mono_ln_begin: np.ndarray
    stereo_ln_begin: np.ndarray
    pose_twc: np.ndarray     # (7,NP)
    pose_fixed: np.ndarray
    point_xyz: np.ndarray    # (3,NL)
    line_wd: np.ndarray      # (6,NLn)
    mp_pose: np.ndarray
    mp_point: np.ndarray
    mp_cam: np.ndarray
    mp_meas: np.ndarray      # (2,n)
    sp_pose: np.ndarray
    sp_point: np.ndarray
    sp_cam: np.ndarray
    sp_meas: np.ndarray      # (3,n)
    ml_pose: np.ndarray
    ml_line: np.ndarray
    ml_cam: np.ndarray
    ml_meas: np.ndarray      # (4,n)
    sl_pose: np.ndarray
    sl_line: np.ndarray
    sl_cam: np.ndarray
    sl_meas: np.ndarray      # (8,n)

    @property
    def n_windows(self) -> int:
        return len(self.pose_begin) - 1

    @property
    def n_edges(self) -> int:
        return int(self.mono_pt_begin[-1] + self.stereo_pt_begin[-1] + self.mono_ln_begin[-1] + self.stereo_ln_begin[-1])

    def window_edges(self) -> np.ndarray:
        return (np.diff(self.mono_pt_begin) + np.diff(self.stereo_pt_begin)
                + np.diff(self.mono_ln_begin) + np.diff(self.stereo_ln_begin)).astype(np.int64)

    @staticmethod
    def from_problems(probs: List[LocalProblem]) -> "LocalBatch":
        cams = probs[0].cams
        for p in probs:
            if not np.array_equal(p.cams, cams):
                raise ValueError("all windows of a batch share one camera list")
            for ids in (p.pose_id, p.point_id, p.line_id):
                if len(ids) > 1 and not np.all(np.diff(ids) > 0):
                    raise ValueError("vertex ids must be strictly ascending (std::map order)")
        cat1 = lambda xs, dt: (np.concatenate(xs) if xs else np.zeros(0)).astype(dt)
        catT = lambda xs, d: _f((np.concatenate(xs, axis=0) if xs else np.zeros((0, d))).T)
        kw = dict(
            cameras=_f(cams),
            pose_begin=_offsets([len(p.pose_id) for p in probs]),
            point_begin=_offsets([len(p.point_id) for p in probs]),
            line_begin=_offsets([len(p.line_id) for p in probs]),
            mono_pt_begin=_offsets([len(p.mp_id_pose) for p in probs]),
            stereo_pt_begin=_offsets([len(p.sp_id_pose) for p in probs]),
            mono_ln_begin=_offsets([len(p.ml_id_pose) for p in probs]),
            stereo_ln_begin=_offsets([len(p.sl_id_pose) for p in probs]),
            pose_twc=catT([np.concatenate([p.pose_p, p.pose_q], axis=1) for p in probs], 7),
            pose_fixed=cat1([p.pose_fixed for p in probs], U8),
            point_xyz=catT([p.point_p for p in probs], 3),
            line_wd=catT([p.line_L for p in probs], 6))
        for pre, lm, dim, key in _EDGE_CLASSES:
            kw[f"{pre}_pose"] = cat1([_local_index(p.pose_id, getattr(p, f"{pre}_id_pose"), "pose") for p in probs], I32)
            kw[f"{pre}_{lm}"] = cat1([_local_index(getattr(p, f"{lm}_id"), getattr(p, f"{pre}_id_{lm}"), lm) for p in probs], I32)
            kw[f"{pre}_cam"] = cat1([getattr(p, f"{pre}_id_cam") for p in probs], I32)
            kw[f"{pre}_meas"] = catT([getattr(p, f"{pre}_{key}") for p in probs], dim)
        return LocalBatch(**kw)

    def slice(self, w0: int, w1: int) -> "LocalBatch":
        """Windows [w0, w1) as an independent batch (rank shard)."""
        def rng(b):
            return int(b[w0]), int(b[w1])
        kw = dict(cameras=self.cameras)
        for name in ("pose", "point", "line", "mono_pt", "stereo_pt", "mono_ln", "stereo_ln"):
            b = getattr(self, f"{name}_begin")
            kw[f"{name}_begin"] = _i(b[w0:w1 + 1] - b[w0])
        a, b = rng(self.pose_begin)
        kw["pose_twc"], kw["pose_fixed"] = _f(self.pose_twc[:, a:b]), _u(self.pose_fixed[a:b])
        a, b = rng(self.point_begin)
        kw["point_xyz"] = _f(self.point_xyz[:, a:b])
        a, b = rng(self.line_begin)
        kw["line_wd"] = _f(self.line_wd[:, a:b])
        for (pre, lm, dim, key), bname in zip(_EDGE_CLASSES, ("mono_pt", "stereo_pt", "mono_ln", "stereo_ln")):
            a, b = rng(getattr(self, f"{bname}_begin"))
            kw[f"{pre}_pose"] = _i(getattr(self, f"{pre}_pose")[a:b])
            kw[f"{pre}_{lm}"] = _i(getattr(self, f"{pre}_{lm}")[a:b])
            kw[f"{pre}_cam"] = _i(getattr(self, f"{pre}_cam")[a:b])
            kw[f"{pre}_meas"] = _f(getattr(self, f"{pre}_meas")[:, a:b])
        return LocalBatch(**kw)

    def h2d_bytes(self) -> int:
        """bytes the library copies to the device (with a single camera the per-edge camera indices stay on the host)"""
        single = len(self.cameras) == 1
        return sum(int(v.nbytes) for k, v in self.__dict__.items()
                   if isinstance(v, np.ndarray) and not (single and k.endswith("_cam")))


@dataclass
class LocalBatchResult:
    pose_twc: np.ndarray
    point_xyz: np.ndarray
    line_wd: np.ndarray
    mp_inlier: np.ndarray
    sp_inlier: np.ndarray
    ml_inlier: np.ndarray
    sl_inlier: np.ndarray
    stats: np.ndarray

    def d2h_bytes(self) -> int:
        return sum(int(v.nbytes) for v in self.__dict__.values() if isinstance(v, np.ndarray))

    def scatter_back(self, batch: LocalBatch, probs: List[LocalProblem]) -> None:
        """Mutates the reference-style containers in place (g2o_optimization.cc:213-251)."""
        for w, p in enumerate(probs):
            a, b = int(batch.pose_begin[w]), int(batch.pose_begin[w + 1])
            p.pose_p[:] = self.pose_twc[:3, a:b].T
            p.pose_q[:] = self.pose_twc[3:, a:b].T
            a, b = int(batch.point_begin[w]), int(batch.point_begin[w + 1])
            p.point_p[:] = self.point_xyz[:, a:b].T
            a, b = int(batch.line_begin[w]), int(batch.line_begin[w + 1])
            p.line_L[:] = self.line_wd[:, a:b].T
            for pre, bname in (("mp", "mono_pt"), ("sp", "stereo_pt"), ("ml", "mono_ln"), ("sl", "stereo_ln")):
                beg = getattr(batch, f"{bname}_begin")
                a, b = int(beg[w]), int(beg[w + 1])
                getattr(p, f"{pre}_inlier")[:] = getattr(self, f"{pre}_inlier")[a:b]


def shard_range(n_units: int, rank: int, world: int) -> tuple:
    """Block partition of independent units (frames / windows) over ranks: no data-path collective
    (SURVEY §8e). Returns [begin, end)."""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


# ------------------------------------------------------------------------------------------------
# Global BA: landmark partition of ONE problem over ranks
# ------------------------------------------------------------------------------------------------
@dataclass
class LandmarkShard:
    """One rank's share of a LocalProblem: all poses, a block of the points and lines with all their
    constraints, and the index maps that scatter the rank's results back into the full problem."""
    problem: LocalProblem
    point_idx: np.ndarray                       # positions of the shard's points in the full problem
    line_idx: np.ndarray
    edge_idx: dict                              # {"mp" | "sp" | "ml" | "sl": positions in the full arrays}


def _balanced_blocks(weights: np.ndarray, world: int) -> np.ndarray:
    """Boundaries [world+1] of contiguous blocks with about equal total weight."""
    n = len(weights)
    if n == 0:
        return np.zeros(world + 1, dtype=np.int64)
    cum = np.cumsum(weights, dtype=np.float64)
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left") + 1
    return np.concatenate([[0], np.minimum(cuts, n), [n]]).astype(np.int64)


def shard_landmarks(p: LocalProblem, rank: int, world: int) -> LandmarkShard:
    """Points and lines (each with ALL its constraints) are split into `world` contiguous blocks of
    about equal constraint count; every rank keeps every pose. Deterministic, no communication."""
    def block(ids, id_a, id_b):
        order = np.argsort(ids, kind="stable")
        w = np.ones(len(ids), dtype=np.int64)  # +1: landmarks without constraints still count
        pos_of = []
        for arr in (id_a, id_b):
            pos = order[np.searchsorted(ids, arr, sorter=order)] if len(arr) else np.zeros(0, dtype=np.int64)
            pos_of.append(pos)
            np.add.at(w, pos, 1)
        cuts = _balanced_blocks(w, world)
        lo, hi = cuts[rank], cuts[rank + 1]
        return np.arange(lo, hi, dtype=np.int64), [np.nonzero((q >= lo) & (q < hi))[0] for q in pos_of]

    pt_idx, (mp_sel, sp_sel) = block(p.point_id, p.mp_id_point, p.sp_id_point)
    ln_idx, (ml_sel, sl_sel) = block(p.line_id, p.ml_id_line, p.sl_id_line)
    kw = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in p.__dict__.items()
          if not k.startswith(("mp_", "sp_", "ml_", "sl_", "point_", "line_"))}
    kw["point_id"], kw["point_p"] = p.point_id[pt_idx].copy(), p.point_p[pt_idx].copy()
    kw["line_id"], kw["line_L"] = p.line_id[ln_idx].copy(), p.line_L[ln_idx].copy()
    kw["truth"] = {}
    edge_idx = {"mp": mp_sel, "sp": sp_sel, "ml": ml_sel, "sl": sl_sel}
    for pre, lm, key in (("mp", "id_point", "kp"), ("sp", "id_point", "kp"), ("ml", "id_line", "l2d"), ("sl", "id_line", "l2d")):
        for name in ("id_pose", lm, "id_cam", key, "inlier"):
            kw[f"{pre}_{name}"] = getattr(p, f"{pre}_{name}")[edge_idx[pre]].copy()
    return LandmarkShard(LocalProblem(**kw).normalise(), pt_idx, ln_idx, edge_idx)


def merge_landmark_shards(full: LocalProblem, shards: Sequence[LandmarkShard]) -> LocalProblem:
    """Scatter the solved shards back into a copy of the full problem (poses are identical on every
    rank; rank 0's are taken)."""
    out = full.copy()
    out.pose_p[:], out.pose_q[:] = shards[0].problem.pose_p, shards[0].problem.pose_q
    for s in shards:
        out.point_p[s.point_idx] = s.problem.point_p
        out.line_L[s.line_idx] = s.problem.line_L
        for pre in ("mp", "sp", "ml", "sl"):
            getattr(out, f"{pre}_inlier")[s.edge_idx[pre]] = getattr(s.problem, f"{pre}_inlier")
    return out
