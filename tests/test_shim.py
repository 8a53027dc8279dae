"""The C++ shim (include/rspl_ba/g2o_optimization_shim.hpp): the reference's two entry points with
their exact signatures on top of the C-ABI. CPU: it compiles against layout-compatible mock
types (Eigen / g2o are not installable here). GPU: driving the shim with the reference-style
containers gives bit-identical results to the direct C-ABI call."""
import os
import subprocess
import sys

import numpy as np
import pytest

from rspl_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    from rspl_slam_b200 import build
    build.build_library()
    exe = str(tmp_path_factory.mktemp("shim") / "shim_driver")
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", f"-I{ROOT}/include", f"-I{ROOT}/tests/shim",
                    f"{ROOT}/tests/shim/shim_driver.cpp", "-o", exe, f"-L{ROOT}/rspl_slam_b200", "-lrspl_ba",
                    f"-Wl,-rpath,{ROOT}/rspl_slam_b200"], check=True, env=env)
    return exe


sys.path.insert(0, os.path.join(ROOT, "tests", "shim"))
from shim_dump import dump_problem as _dump  # noqa: E402


def test_shim_compiles_with_reference_signatures(driver):
    assert os.path.exists(driver)
    hdr = open(os.path.join(ROOT, "include", "rspl_ba", "g2o_optimization_shim.hpp")).read()
    # the two reference declarations (g2o_optimization.h:15-22), verbatim parameter lists
    assert "inline void LocalmapOptimization(MapOfPoses& poses, MapOfPoints3d& points, MapOfLine3d& lines," in hdr
    assert "inline int FrameOptimization(MapOfPoses& poses, MapOfPoints3d& points, std::vector<CameraPtr>& camera_list," in hdr


@pytest.mark.gpu
def test_shim_local_equals_direct_cabi(driver, gpu_ctx, tmp_path):
    from rspl_slam_b200.problem import LocalBatch
    p = synth.make_local_problem(synth.config_seed(1, 400), n_kf=6, n_points=300, n_lines=30, first_kf_id=12)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _dump(0, p, fin)
    subprocess.run([driver, fin, fout], check=True)
    out = np.fromfile(fout, dtype=np.float64)
    batch = LocalBatch.from_problems([p])
    res = gpu_ctx.local_batch(batch)
    k = 1
    npose, npt, nln = len(p.pose_id), len(p.point_id), len(p.line_id)
    poses = out[k:k + 7 * npose].reshape(npose, 7); k += 7 * npose
    pts = out[k:k + 3 * npt].reshape(npt, 3); k += 3 * npt
    lns = out[k:k + 6 * nln].reshape(nln, 6); k += 6 * nln
    assert np.array_equal(poses.view(np.uint64), np.ascontiguousarray(res.pose_twc.T).view(np.uint64))
    assert np.array_equal(pts.view(np.uint64), np.ascontiguousarray(res.point_xyz.T).view(np.uint64))
    assert np.array_equal(lns.view(np.uint64), np.ascontiguousarray(res.line_wd.T).view(np.uint64))
    for name in ("mp_inlier", "sp_inlier", "ml_inlier", "sl_inlier"):
        n = len(getattr(res, name))
        assert np.array_equal(out[k:k + n].astype(np.uint8), getattr(res, name)); k += n
    assert k == len(out)


@pytest.mark.gpu
def test_shim_frame_equals_direct_cabi(driver, gpu_ctx, tmp_path):
    from rspl_slam_b200.problem import FrameBatch
    p = synth.make_frame_problem(synth.config_seed(2, 400), n_points=250, stereo_frac=0.8)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _dump(1, p, fin)
    subprocess.run([driver, fin, fout], check=True)
    out = np.fromfile(fout, dtype=np.float64)
    batch = FrameBatch.from_problems([p])
    from rspl_slam_b200 import capi
    res = gpu_ctx.frame_batch(batch, capi.make_options(frame_latency_mode=1))  # (the shim's single-frame calls use it)
    assert int(out[0]) == int(res.num_inliers[0])
    assert np.array_equal(out[1:8].view(np.uint64), np.ascontiguousarray(res.pose_twc[:, 0]).view(np.uint64))
    k = 8 + 3 * len(p.point_id)
    nm = len(p.mp_inlier)
    assert np.array_equal(out[k:k + nm].astype(np.uint8), res.mono_inlier)
    assert np.array_equal(out[k + nm:].astype(np.uint8), res.stereo_inlier)


@pytest.mark.gpu
def test_shim_frame_with_lines_equals_direct_cabi(driver, gpu_ctx, tmp_path):
    """The shim's extension entry point (FrameOptimizationWithLinesImpl: the reference's line containers next to
    its FrameOptimization arguments) gives the bits of the direct C-ABI call."""
    from rspl_slam_b200.problem import FrameBatch
    p = synth.make_frame_problem(synth.config_seed(2, 401), n_points=120, stereo_frac=0.8, n_lines=30)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _dump(2, p, fin)
    subprocess.run([driver, fin, fout], check=True)
    out = np.fromfile(fout, dtype=np.float64)
    batch = FrameBatch.from_problems([p])
    from rspl_slam_b200 import capi
    res = gpu_ctx.frame_batch(batch, capi.make_options(frame_latency_mode=1))  # (the shim's single-frame calls use it)
    assert int(out[0]) == int(res.num_inliers[0])
    assert np.array_equal(out[1:8].view(np.uint64), np.ascontiguousarray(res.pose_twc[:, 0]).view(np.uint64))
    k = 8 + 3 * len(p.point_id) + 6 * len(p.line_id)
    nm, ns, nml = len(p.mp_inlier), len(p.sp_inlier), len(p.ml_inlier)
    assert np.array_equal(out[k:k + nm].astype(np.uint8), res.mono_inlier)
    assert np.array_equal(out[k + nm:k + nm + ns].astype(np.uint8), res.stereo_inlier)
    assert np.array_equal(out[k + nm + ns:k + nm + ns + nml].astype(np.uint8), res.mline_inlier)
    assert np.array_equal(out[k + nm + ns + nml:].astype(np.uint8), res.sline_inlier)
