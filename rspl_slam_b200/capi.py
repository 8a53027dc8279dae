"""ctypes binding of include/rspl_ba.h (the C-ABI a C++ caller links against).

The shared library is built in-tree by ``rspl_slam_b200.build.build_library`` (nvcc, sm_100a) and
must exist: there is no Python / CPU fallback — a missing library or a machine without a CUDA
device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

import numpy as np

from .problem import (FrameBatch, FrameBatchResult, LocalBatch, LocalBatchResult, OptimizationConfig)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librspl_ba.so")

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)

RSPL_BA_OK = 0
RSPL_BA_ERR_INVALID = -1
RSPL_BA_ERR_CUDA = -2
RSPL_BA_ERR_UNSUPPORTED = -3
RSPL_BA_ERR_STATE = -4


class RsplBaOptions(C.Structure):
    _fields_ = [("thr_mono_point", C.c_double), ("thr_stereo_point", C.c_double), ("thr_mono_line", C.c_double),
                ("thr_stereo_line", C.c_double), ("local_iters_pass1", C.c_int32), ("local_iters_pass2", C.c_int32),
                ("frame_rounds", C.c_int32), ("frame_iters", C.c_int32), ("stereo_bf_float", C.c_int32),
                ("frame_latency_mode", C.c_int32)]


class RsplBaStats(C.Structure):
    _fields_ = [("iters", C.c_int32 * 4), ("trials", C.c_int32 * 4), ("edges_linearized", C.c_int64),
                ("edges_evaluated", C.c_int64), ("final_chi2", C.c_double), ("final_lambda", C.c_double)]


STATS_DTYPE = np.dtype([("iters", np.int32, 4), ("trials", np.int32, 4), ("edges_linearized", np.int64),
                        ("edges_evaluated", np.int64), ("final_chi2", np.float64), ("final_lambda", np.float64)])
assert STATS_DTYPE.itemsize == C.sizeof(RsplBaStats)


class RsplFrameBatch(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("n_cameras", C.c_int32), ("cameras", c_f64p), ("pose_twc", c_f64p),
                ("mono_begin", c_i32p), ("stereo_begin", c_i32p),
                ("mono_meas", c_f64p), ("mono_xw", c_f64p), ("mono_cam", c_i32p), ("mono_inlier", c_u8p),
                ("stereo_meas", c_f64p), ("stereo_xw", c_f64p), ("stereo_cam", c_i32p), ("stereo_inlier", c_u8p),
                # line extension (all NULL: the reference's case)
                ("mono_line_begin", c_i32p), ("stereo_line_begin", c_i32p),
                ("mono_line_lw", c_f64p), ("mono_line_meas", c_f64p), ("mono_line_cam", c_i32p), ("mono_line_inlier", c_u8p),
                ("stereo_line_lw", c_f64p), ("stereo_line_meas", c_f64p), ("stereo_line_cam", c_i32p),
                ("stereo_line_inlier", c_u8p)]


class RsplFrameBatchResult(C.Structure):
    _fields_ = [("pose_twc", c_f64p), ("mono_inlier", c_u8p), ("stereo_inlier", c_u8p), ("num_inliers", c_i32p),
                ("stats", C.POINTER(RsplBaStats)), ("mono_line_inlier", c_u8p), ("stereo_line_inlier", c_u8p)]


class RsplLocalBatch(C.Structure):
    _fields_ = [("n_windows", C.c_int32), ("n_cameras", C.c_int32), ("cameras", c_f64p),
                ("pose_begin", c_i32p), ("point_begin", c_i32p), ("line_begin", c_i32p),
                ("mono_pt_begin", c_i32p), ("stereo_pt_begin", c_i32p), ("mono_ln_begin", c_i32p),
                ("stereo_ln_begin", c_i32p),
                ("pose_twc", c_f64p), ("pose_fixed", c_u8p), ("point_xyz", c_f64p), ("line_wd", c_f64p),
                ("mp_pose", c_i32p), ("mp_point", c_i32p), ("mp_cam", c_i32p), ("mp_meas", c_f64p),
                ("sp_pose", c_i32p), ("sp_point", c_i32p), ("sp_cam", c_i32p), ("sp_meas", c_f64p),
                ("ml_pose", c_i32p), ("ml_line", c_i32p), ("ml_cam", c_i32p), ("ml_meas", c_f64p),
                ("sl_pose", c_i32p), ("sl_line", c_i32p), ("sl_cam", c_i32p), ("sl_meas", c_f64p)]


class RsplLocalBatchResult(C.Structure):
    _fields_ = [("pose_twc", c_f64p), ("point_xyz", c_f64p), ("line_wd", c_f64p),
                ("mp_inlier", c_u8p), ("sp_inlier", c_u8p), ("ml_inlier", c_u8p), ("sl_inlier", c_u8p),
                ("stats", C.POINTER(RsplBaStats))]


#: every symbol include/rspl_ba.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    "rspl_ba_version", "rspl_ba_default_options", "rspl_ba_create", "rspl_ba_destroy", "rspl_ba_last_error",
    "rspl_ba_stream", "rspl_ba_device", "rspl_ba_frame_batch", "rspl_ba_frame_batch_upload",
    "rspl_ba_frame_batch_solve", "rspl_ba_frame_batch_download", "rspl_ba_local_batch",
    "rspl_ba_local_batch_upload", "rspl_ba_local_batch_solve", "rspl_ba_local_batch_download",
    "rspl_ba_alloc_pinned", "rspl_ba_free_pinned", "rspl_ba_launch_count", "rspl_ba_sync",
    "rspl_ba_eval_edges", "rspl_ba_oplus", "rspl_ba_triangulate_points", "rspl_ba_update_maplines", "rspl_ba_local_batch_update_maplines", "rspl_ba_unit_math",
    "rspl_ba_set_profiling", "rspl_ba_get_profile",
    "rspl_ba_comm_unique_id", "rspl_ba_comm_init", "rspl_ba_comm_destroy", "rspl_ba_comm_size", "rspl_ba_comm_rank",
    "rspl_ba_collective_count", "rspl_ba_global_upload", "rspl_ba_global_solve", "rspl_ba_global_download")

_lib = None
LOCAL_BA_READY = True
COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """Rank 0: a fresh NCCL unique id to hand to the other ranks (e.g. with torch.distributed.broadcast)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = load_library().rspl_ba_comm_unique_id(buf)
    if rc != 0:
        raise RsplBaError(rc, "rspl_ba_comm_unique_id failed (is libnccl.so.2 loadable?)")
    return buf.raw


class RsplBaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rspl_ba error {code}: {msg}")
        self.code = code


def load_library() -> C.CDLL:
    """Loads the in-tree CUDA library. Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RSPL_BA_LIB", LIB_PATH)  # A/B builds of the same library (profiles/scripts)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    ctx = C.c_void_p
    L.rspl_ba_version.restype = C.c_int
    L.rspl_ba_default_options.argtypes = [C.POINTER(RsplBaOptions)]
    L.rspl_ba_default_options.restype = None
    L.rspl_ba_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(ctx)]
    L.rspl_ba_create.restype = C.c_int
    L.rspl_ba_destroy.argtypes = [ctx]
    L.rspl_ba_destroy.restype = None
    L.rspl_ba_last_error.argtypes = [ctx]
    L.rspl_ba_last_error.restype = C.c_char_p
    L.rspl_ba_stream.argtypes = [ctx]
    L.rspl_ba_stream.restype = C.c_void_p
    L.rspl_ba_device.argtypes = [ctx]
    L.rspl_ba_device.restype = C.c_int
    L.rspl_ba_launch_count.argtypes = [ctx]
    L.rspl_ba_launch_count.restype = C.c_int64
    L.rspl_ba_sync.argtypes = [ctx]
    L.rspl_ba_sync.restype = C.c_int
    L.rspl_ba_alloc_pinned.argtypes = [C.c_size_t]
    L.rspl_ba_alloc_pinned.restype = C.c_void_p
    L.rspl_ba_free_pinned.argtypes = [C.c_void_p]
    L.rspl_ba_free_pinned.restype = None
    opt = C.POINTER(RsplBaOptions)
    L.rspl_ba_frame_batch.argtypes = [ctx, C.POINTER(RsplFrameBatch), opt, C.POINTER(RsplFrameBatchResult)]
    L.rspl_ba_frame_batch_upload.argtypes = [ctx, C.POINTER(RsplFrameBatch)]
    L.rspl_ba_frame_batch_solve.argtypes = [ctx, opt]
    L.rspl_ba_frame_batch_download.argtypes = [ctx, C.POINTER(RsplFrameBatchResult)]
    L.rspl_ba_local_batch.argtypes = [ctx, C.POINTER(RsplLocalBatch), opt, C.POINTER(RsplLocalBatchResult)]
    L.rspl_ba_local_batch_upload.argtypes = [ctx, C.POINTER(RsplLocalBatch)]
    L.rspl_ba_local_batch_solve.argtypes = [ctx, opt]
    L.rspl_ba_local_batch_download.argtypes = [ctx, C.POINTER(RsplLocalBatchResult)]
    for n in ("rspl_ba_frame_batch", "rspl_ba_frame_batch_upload", "rspl_ba_frame_batch_solve",
              "rspl_ba_frame_batch_download", "rspl_ba_local_batch", "rspl_ba_local_batch_upload",
              "rspl_ba_local_batch_solve", "rspl_ba_local_batch_download"):
        getattr(L, n).restype = C.c_int
    L.rspl_ba_eval_edges.argtypes = [ctx, C.c_int, C.c_int32, c_f64p, c_f64p, c_f64p, c_f64p, C.c_int32,
                                     c_f64p, c_f64p, c_f64p, c_f64p]
    L.rspl_ba_eval_edges.restype = C.c_int
    L.rspl_ba_triangulate_points.argtypes = [ctx, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_f64p, C.c_int32,
                                             c_f64p, c_f64p, c_f64p, C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
    L.rspl_ba_triangulate_points.restype = C.c_int
    L.rspl_ba_update_maplines.argtypes = [ctx, C.c_int32, c_f64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, c_f64p,
                                          c_f64p, C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
    L.rspl_ba_update_maplines.restype = C.c_int
    L.rspl_ba_local_batch_update_maplines.argtypes = [ctx, C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_f64p,
                                                      C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
    L.rspl_ba_local_batch_update_maplines.restype = C.c_int
    L.rspl_ba_unit_math.argtypes = [ctx, C.c_int, C.c_int32, c_f64p, c_f64p]
    L.rspl_ba_unit_math.restype = C.c_int
    L.rspl_ba_oplus.argtypes = [ctx, C.c_int, C.c_int32, c_f64p, c_f64p, c_f64p]
    L.rspl_ba_oplus.restype = C.c_int
    L.rspl_ba_set_profiling.argtypes = [ctx, C.c_int]
    L.rspl_ba_set_profiling.restype = C.c_int
    L.rspl_ba_get_profile.argtypes = [ctx, c_f64p, C.POINTER(C.c_int64)]
    L.rspl_ba_get_profile.restype = C.c_int
    L.rspl_ba_comm_unique_id.argtypes = [C.c_void_p]
    L.rspl_ba_comm_unique_id.restype = C.c_int
    L.rspl_ba_comm_init.argtypes = [ctx, C.c_int, C.c_int, C.c_void_p]
    L.rspl_ba_comm_init.restype = C.c_int
    for n in ("rspl_ba_comm_destroy", "rspl_ba_comm_size", "rspl_ba_comm_rank"):
        getattr(L, n).argtypes = [ctx]
        getattr(L, n).restype = C.c_int
    L.rspl_ba_collective_count.argtypes = [ctx]
    L.rspl_ba_collective_count.restype = C.c_int64
    L.rspl_ba_global_upload.argtypes = [ctx, C.POINTER(RsplLocalBatch)]
    L.rspl_ba_global_solve.argtypes = [ctx, opt]
    L.rspl_ba_global_download.argtypes = [ctx, C.POINTER(RsplLocalBatchResult)]
    for n in ("rspl_ba_global_upload", "rspl_ba_global_solve", "rspl_ba_global_download"):
        getattr(L, n).restype = C.c_int
    _lib = L
    return L


def _p(a: Optional[np.ndarray], ct):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ct)


def make_options(cfg: Optional[OptimizationConfig] = None, local_iters=(10, 5), frame_rounds: int = 4,
                 frame_iters: int = 10, stereo_bf_float: int = 1, frame_latency_mode: int = 0) -> RsplBaOptions:
    o = RsplBaOptions()
    load_library().rspl_ba_default_options(C.byref(o))
    if cfg is not None:
        o.thr_mono_point, o.thr_stereo_point = cfg.mono_point, cfg.stereo_point
        o.thr_mono_line, o.thr_stereo_line = cfg.mono_line, cfg.stereo_line
    o.local_iters_pass1, o.local_iters_pass2 = local_iters
    o.frame_rounds, o.frame_iters, o.stereo_bf_float = frame_rounds, frame_iters, stereo_bf_float
    o.frame_latency_mode = frame_latency_mode
    return o


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory from rspl_ba_alloc_pinned (kept alive by the array)."""
    L = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = L.rspl_ba_alloc_pinned(max(n, 1))
    if not ptr:
        raise RsplBaError(RSPL_BA_ERR_CUDA, "cudaHostAlloc failed")
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    weakref.finalize(buf, L.rspl_ba_free_pinned, ptr)  # freed when the last view of buf dies
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Context:
    """One solver context = one CUDA stream + device workspaces (one per calling thread)."""

    def __init__(self, device: int = -1, stream: int = 0):
        self._L = load_library()
        self._ctx = C.c_void_p()
        rc = self._L.rspl_ba_create(device, C.c_void_p(stream) if stream else None, C.byref(self._ctx))
        if rc != RSPL_BA_OK:
            self._ctx = None
            raise RsplBaError(rc, "rspl_ba_create failed (no CUDA device? there is no CPU fallback)")
        self._keep = None

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.rspl_ba_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != RSPL_BA_OK:
            raise RsplBaError(rc, self._L.rspl_ba_last_error(self._ctx).decode())

    @property
    def stream(self) -> int:
        return int(self._L.rspl_ba_stream(self._ctx) or 0)

    @property
    def device(self) -> int:
        return int(self._L.rspl_ba_device(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self._L.rspl_ba_launch_count(self._ctx))

    def sync(self):
        self._check(self._L.rspl_ba_sync(self._ctx))

    # ---------------- FrameOptimization ----------------
    @staticmethod
    def _frame_struct(b: FrameBatch) -> RsplFrameBatch:
        return RsplFrameBatch(
            b.n_frames, len(b.cameras), _p(b.cameras, c_f64p), _p(b.pose_twc, c_f64p),
            _p(b.mono_begin, c_i32p), _p(b.stereo_begin, c_i32p),
            _p(b.mono_meas, c_f64p), _p(b.mono_xw, c_f64p), _p(b.mono_cam, c_i32p), _p(b.mono_inlier, c_u8p),
            _p(b.stereo_meas, c_f64p), _p(b.stereo_xw, c_f64p), _p(b.stereo_cam, c_i32p), _p(b.stereo_inlier, c_u8p),
            _p(b.mline_begin, c_i32p), _p(b.sline_begin, c_i32p),
            _p(b.mline_lw, c_f64p), _p(b.mline_meas, c_f64p), _p(b.mline_cam, c_i32p), _p(b.mline_inlier, c_u8p),
            _p(b.sline_lw, c_f64p), _p(b.sline_meas, c_f64p), _p(b.sline_cam, c_i32p), _p(b.sline_inlier, c_u8p))

    @staticmethod
    def alloc_frame_result(b: FrameBatch, pinned: bool = False) -> FrameBatchResult:
        mk = pinned_empty if pinned else (lambda shape, dt: np.zeros(shape, dtype=dt))
        return FrameBatchResult(
            pose_twc=mk((7, b.n_frames), np.float64), mono_inlier=mk((int(b.mono_begin[-1]),), np.uint8),
            stereo_inlier=mk((int(b.stereo_begin[-1]),), np.uint8), num_inliers=mk((b.n_frames,), np.int32),
            stats=mk((b.n_frames,), STATS_DTYPE),
            mline_inlier=mk((int(b.mline_begin[-1]),), np.uint8) if b.mline_begin is not None else None,
            sline_inlier=mk((int(b.sline_begin[-1]),), np.uint8) if b.sline_begin is not None else None)

    @staticmethod
    def _frame_result_struct(r: FrameBatchResult) -> RsplFrameBatchResult:
        return RsplFrameBatchResult(_p(r.pose_twc, c_f64p), _p(r.mono_inlier, c_u8p), _p(r.stereo_inlier, c_u8p),
                                    _p(r.num_inliers, c_i32p), C.cast(r.stats.ctypes.data, C.POINTER(RsplBaStats)),
                                    _p(r.mline_inlier, c_u8p), _p(r.sline_inlier, c_u8p))

    def frame_batch(self, b: FrameBatch, opt: Optional[RsplBaOptions] = None,
                    out: Optional[FrameBatchResult] = None) -> FrameBatchResult:
        """rspl_ba_frame_batch: upload + solve + download with host buffers."""
        opt = opt or make_options()
        out = out or self.alloc_frame_result(b)
        s, r = self._frame_struct(b), self._frame_result_struct(out)
        self._check(self._L.rspl_ba_frame_batch(self._ctx, C.byref(s), C.byref(opt), C.byref(r)))
        return out

    def frame_batch_upload(self, b: FrameBatch):
        s = self._frame_struct(b)
        self._check(self._L.rspl_ba_frame_batch_upload(self._ctx, C.byref(s)))

    def frame_batch_solve(self, opt: Optional[RsplBaOptions] = None):
        opt = opt or make_options()
        self._check(self._L.rspl_ba_frame_batch_solve(self._ctx, C.byref(opt)))

    def frame_batch_download(self, out: FrameBatchResult) -> FrameBatchResult:
        r = self._frame_result_struct(out)
        self._check(self._L.rspl_ba_frame_batch_download(self._ctx, C.byref(r)))
        return out

    # ---------------- LocalmapOptimization ----------------
    @staticmethod
    def _local_struct(b: LocalBatch) -> RsplLocalBatch:
        return RsplLocalBatch(
            b.n_windows, len(b.cameras), _p(b.cameras, c_f64p),
            _p(b.pose_begin, c_i32p), _p(b.point_begin, c_i32p), _p(b.line_begin, c_i32p),
            _p(b.mono_pt_begin, c_i32p), _p(b.stereo_pt_begin, c_i32p), _p(b.mono_ln_begin, c_i32p),
            _p(b.stereo_ln_begin, c_i32p),
            _p(b.pose_twc, c_f64p), _p(b.pose_fixed, c_u8p), _p(b.point_xyz, c_f64p), _p(b.line_wd, c_f64p),
            _p(b.mp_pose, c_i32p), _p(b.mp_point, c_i32p), _p(b.mp_cam, c_i32p), _p(b.mp_meas, c_f64p),
            _p(b.sp_pose, c_i32p), _p(b.sp_point, c_i32p), _p(b.sp_cam, c_i32p), _p(b.sp_meas, c_f64p),
            _p(b.ml_pose, c_i32p), _p(b.ml_line, c_i32p), _p(b.ml_cam, c_i32p), _p(b.ml_meas, c_f64p),
            _p(b.sl_pose, c_i32p), _p(b.sl_line, c_i32p), _p(b.sl_cam, c_i32p), _p(b.sl_meas, c_f64p))

    @staticmethod
    def alloc_local_result(b: LocalBatch, pinned: bool = False) -> LocalBatchResult:
        mk = pinned_empty if pinned else (lambda shape, dt: np.zeros(shape, dtype=dt))
        return LocalBatchResult(
            pose_twc=mk(b.pose_twc.shape, np.float64), point_xyz=mk(b.point_xyz.shape, np.float64),
            line_wd=mk(b.line_wd.shape, np.float64),
            mp_inlier=mk((len(b.mp_pose),), np.uint8), sp_inlier=mk((len(b.sp_pose),), np.uint8),
            ml_inlier=mk((len(b.ml_pose),), np.uint8), sl_inlier=mk((len(b.sl_pose),), np.uint8),
            stats=mk((b.n_windows,), STATS_DTYPE))

    @staticmethod
    def _local_result_struct(r: LocalBatchResult) -> RsplLocalBatchResult:
        return RsplLocalBatchResult(_p(r.pose_twc, c_f64p), _p(r.point_xyz, c_f64p), _p(r.line_wd, c_f64p),
                                    _p(r.mp_inlier, c_u8p), _p(r.sp_inlier, c_u8p), _p(r.ml_inlier, c_u8p),
                                    _p(r.sl_inlier, c_u8p), C.cast(r.stats.ctypes.data, C.POINTER(RsplBaStats)))

    def local_batch(self, b: LocalBatch, opt: Optional[RsplBaOptions] = None,
                    out: Optional[LocalBatchResult] = None) -> LocalBatchResult:
        opt = opt or make_options()
        out = out or self.alloc_local_result(b)
        s, r = self._local_struct(b), self._local_result_struct(out)
        self._check(self._L.rspl_ba_local_batch(self._ctx, C.byref(s), C.byref(opt), C.byref(r)))
        return out

    def local_batch_upload(self, b: LocalBatch):
        s = self._local_struct(b)
        self._check(self._L.rspl_ba_local_batch_upload(self._ctx, C.byref(s)))

    def local_batch_solve(self, opt: Optional[RsplBaOptions] = None):
        opt = opt or make_options()
        self._check(self._L.rspl_ba_local_batch_solve(self._ctx, C.byref(opt)))

    def local_batch_download(self, out: LocalBatchResult) -> LocalBatchResult:
        r = self._local_result_struct(out)
        self._check(self._L.rspl_ba_local_batch_download(self._ctx, C.byref(r)))
        return out

    # ---- global BA: one problem, landmarks partitioned over the ranks of a communicator ----
    def comm_init(self, n_ranks: int, rank: int, unique_id: Optional[bytes] = None):
        """Joins the NCCL communicator of the global-BA path (`unique_id` from `comm_unique_id()` on rank 0)."""
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES) if unique_id is not None else None
        self._check(self._L.rspl_ba_comm_init(self._ctx, n_ranks, rank, buf))

    def comm_destroy(self):
        self._check(self._L.rspl_ba_comm_destroy(self._ctx))

    def comm_size(self) -> int:
        return int(self._L.rspl_ba_comm_size(self._ctx))

    def collective_count(self) -> int:
        return int(self._L.rspl_ba_collective_count(self._ctx))

    def global_upload(self, shard: LocalBatch):
        s = self._local_struct(shard)
        self._check(self._L.rspl_ba_global_upload(self._ctx, C.byref(s)))

    def global_solve(self, opt: Optional[RsplBaOptions] = None):
        opt = opt or make_options()
        self._check(self._L.rspl_ba_global_solve(self._ctx, C.byref(opt)))

    def global_download(self, out: LocalBatchResult) -> LocalBatchResult:
        r = self._local_result_struct(out)
        self._check(self._L.rspl_ba_global_download(self._ctx, C.byref(r)))
        return out

    def global_ba(self, shard: LocalBatch, opt: Optional[RsplBaOptions] = None) -> LocalBatchResult:
        """Collective: every rank passes its shard (all poses + its landmarks) of the same problem."""
        self.global_upload(shard)
        self.global_solve(opt)
        return self.global_download(self.alloc_local_result(shard))

    PROFILE_CLASSES = ("frame_opt", "local_setup", "unused", "init_pairs", "linearize", "pose_blocks",
                       "schur_prep", "schur_reduce", "reduced_solve", "backsub_update_eval", "lm_control", "flag_writeback",
                       "collectives", "dense_assemble", "schur_tile")

    def set_profiling(self, enabled: bool):
        self._check(self._L.rspl_ba_set_profiling(self._ctx, 1 if enabled else 0))

    def get_profile(self) -> dict:
        """{class: (milliseconds, launches)} accumulated since the last call (CUDA events on the context stream)."""
        ms = np.zeros(16)
        n = np.zeros(16, dtype=np.int64)
        self._check(self._L.rspl_ba_get_profile(self._ctx, _p(ms, c_f64p), n.ctypes.data_as(C.POINTER(C.c_int64))))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.PROFILE_CLASSES)}

    # ---------------- unit-level ----------------
    def eval_edges(self, edge_type: int, pose7: np.ndarray, lm: np.ndarray, meas: np.ndarray, cam5: np.ndarray,
                   stereo_bf_float: int = 1):
        n = len(pose7)
        pose7 = np.ascontiguousarray(pose7, dtype=np.float64).reshape(n, 7)
        lm6 = np.zeros((n, 6))
        lm6[:, :lm.shape[1]] = lm
        m8 = np.zeros((n, 8))
        m8[:, :meas.shape[1]] = meas
        cam5 = np.ascontiguousarray(cam5, dtype=np.float64)
        err, Jl, Jp, chi2 = np.zeros((n, 4)), np.zeros((n, 16)), np.zeros((n, 24)), np.zeros(n)
        self._check(self._L.rspl_ba_eval_edges(self._ctx, edge_type, n, _p(pose7, c_f64p), _p(lm6, c_f64p),
                                               _p(m8, c_f64p), _p(cam5, c_f64p), stereo_bf_float, _p(err, c_f64p),
                                               _p(Jl, c_f64p), _p(Jp, c_f64p), _p(chi2, c_f64p)))
        return err, Jl, Jp, chi2

    def triangulate_points(self, obs_begin, obs_frame, obs_uv, frame_twc, cam5, xyz_init=None, out=None):
        """Batched Map::TriangulateMappoint (map.cc:292-339). obs_uv [2][n_obs], frame_twc [7][n_frames] (p, q xyzw).
        Returns (xyz [3][n_points], ok [n_points] uint8, number triangulated); xyz of a failed point keeps xyz_init.
        out = (xyz, ok): caller-owned result arrays (e.g. page-locked ones from pinned_empty), used in place."""
        obs_begin = np.ascontiguousarray(obs_begin, dtype=np.int32)
        obs_frame = np.ascontiguousarray(obs_frame, dtype=np.int32)
        obs_uv = np.ascontiguousarray(obs_uv, dtype=np.float64)
        frame_twc = np.ascontiguousarray(frame_twc, dtype=np.float64)
        cam5 = np.ascontiguousarray(cam5, dtype=np.float64)
        n = len(obs_begin) - 1
        if out is not None:
            xyz, ok = out
            assert xyz.shape == (3, n) and xyz.dtype == np.float64 and xyz.flags.c_contiguous and ok.shape == (n,) and ok.dtype == np.uint8
        else:
            xyz = np.zeros((3, n)) if xyz_init is None else np.ascontiguousarray(xyz_init, dtype=np.float64).copy()
            ok = np.zeros(n, dtype=np.uint8)
        cnt = C.c_int32(0)
        c_i32p, c_u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        self._check(self._L.rspl_ba_triangulate_points(
            self._ctx, n, _p(obs_begin, c_i32p), _p(obs_frame, c_i32p), _p(obs_uv, c_f64p), frame_twc.shape[1],
            _p(frame_twc, c_f64p), _p(cam5, c_f64p), _p(xyz, c_f64p), _p(ok, c_u8p), C.byref(cnt)))
        return xyz, ok, int(cnt.value)

    def update_maplines(self, line_wd, pt_begin, pt_index, point_xyz, endpoints_init=None, out=None):
        """Batched Map::UppdateMapline (map.cc:121-177): endpoint refresh of optimised lines from their map points.
        line_wd [6][n_lines], point_xyz [3][n_points], CSR pt_begin / pt_index. Returns (endpoints [6][n_lines],
        ok [n_lines] uint8, number refreshed); endpoints of a line that is not refreshed keep endpoints_init.
        out = (endpoints, ok): caller-owned result arrays (e.g. page-locked ones from pinned_empty), used in place."""
        line_wd = np.ascontiguousarray(line_wd, dtype=np.float64)
        pt_begin = np.ascontiguousarray(pt_begin, dtype=np.int32)
        pt_index = np.ascontiguousarray(pt_index, dtype=np.int32)
        point_xyz = np.ascontiguousarray(point_xyz, dtype=np.float64)
        n = len(pt_begin) - 1
        if out is not None:
            ends, ok = out
            assert ends.shape == (6, n) and ends.dtype == np.float64 and ends.flags.c_contiguous and ok.shape == (n,) and ok.dtype == np.uint8
        else:
            ends = np.zeros((6, n)) if endpoints_init is None else np.ascontiguousarray(endpoints_init, dtype=np.float64).copy()
            ok = np.zeros(n, dtype=np.uint8)
        cnt = C.c_int32(0)
        c_i32p, c_u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        self._check(self._L.rspl_ba_update_maplines(
            self._ctx, n, _p(line_wd, c_f64p), _p(pt_begin, c_i32p), _p(pt_index, c_i32p), point_xyz.shape[1],
            _p(point_xyz, c_f64p), _p(ends, c_f64p), _p(ok, c_u8p), C.byref(cnt)))
        return ends, ok, int(cnt.value)

    def local_update_maplines(self, pt_begin, pt_index, out=None):
        """The endpoint refresh on the resident result of the local BA just solved on this context: pt_begin
        [n_lines + 1] over all lines of the batch, pt_index batch-wide point indices. Returns (endpoints [6][n_lines]
        (0 where not refreshed), ok [n_lines] uint8, number refreshed)."""
        pt_begin = np.ascontiguousarray(pt_begin, dtype=np.int32)
        pt_index = np.ascontiguousarray(pt_index, dtype=np.int32)
        n = len(pt_begin) - 1
        if out is not None:
            ends, ok = out
            assert ends.shape == (6, n) and ends.dtype == np.float64 and ends.flags.c_contiguous and ok.shape == (n,) and ok.dtype == np.uint8
        else:
            ends, ok = np.zeros((6, n)), np.zeros(n, dtype=np.uint8)
        cnt = C.c_int32(0)
        c_i32p, c_u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        self._check(self._L.rspl_ba_local_batch_update_maplines(self._ctx, _p(pt_begin, c_i32p), _p(pt_index, c_i32p),
                                                                _p(ends, c_f64p), _p(ok, c_u8p), C.byref(cnt)))
        return ends, ok, int(cnt.value)

    def unit_math(self, op: int, x: np.ndarray) -> np.ndarray:
        """rcp_nr (0) / rsqrt_nr (1) / sqrt_nr (2) of ba_math.cuh on the device."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        self._check(self._L.rspl_ba_unit_math(self._ctx, op, len(x), _p(x, c_f64p), _p(out, c_f64p)))
        return out

    def oplus(self, kind: int, state: np.ndarray, upd: np.ndarray) -> np.ndarray:
        n = len(state)
        s7, u6 = np.zeros((n, 7)), np.zeros((n, 6))
        s7[:, :state.shape[1]] = state
        u6[:, :upd.shape[1]] = upd
        out = np.zeros((n, 7))
        self._check(self._L.rspl_ba_oplus(self._ctx, kind, n, _p(s7, c_f64p), _p(u6, c_f64p), _p(out, c_f64p)))
        return out
