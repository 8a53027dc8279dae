"""Pins the CPU oracle against fixtures produced by the REFERENCE itself (real g2o), when they exist.

tests/golden/g2o_*.npz are written by oracle/g2o_validation/make_fixtures.py on a machine that has g2o, Eigen and
OpenCV (the recipe compiles the unmodified reference sources). The build container of this repository has none of
them, so until somebody has run the recipe these tests skip and DESIGN.md keeps saying "parity unpinned"."""
import glob
import os

import numpy as np
import pytest

from rspl_slam_b200.geometry import quat_angle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "g2o_*.npz")))


def test_recipe_is_present():
    for f in ("CMakeLists.txt", "make_fixtures.py"):
        assert os.path.exists(os.path.join(ROOT, "oracle", "g2o_validation", f))
    src = open(os.path.join(ROOT, "tests", "shim", "shim_driver.cpp")).read()
    assert "RSPL_REFERENCE_BUILD" in src and '#include "g2o_optimization/g2o_optimization.h"' in src


@pytest.mark.skipif(not FIX, reason="no g2o-generated fixtures committed (g2o is not installable here): parity unpinned")
@pytest.mark.parametrize("path", FIX)
def test_oracle_matches_g2o_fixture(path):
    from oracle import orc
    from rspl_slam_b200.problem import FrameProblem, LocalProblem
    z = np.load(path)
    local = "in_pose_id" in z.files
    cls = LocalProblem if local else FrameProblem
    p = cls(**{k[3:]: z[k] for k in z.files if k.startswith("in_")})
    (orc.local_ba if local else orc.frame_opt)(p)
    for f in ("mp_inlier", "sp_inlier") + (("ml_inlier", "sl_inlier") if local else ()):
        assert np.array_equal(getattr(p, f), z["out_" + f]), f
    ref_p, ref_q = np.atleast_2d(z["out_pose_p"]), np.atleast_2d(z["out_pose_q"])
    got_p, got_q = np.atleast_2d(p.pose_p), np.atleast_2d(p.pose_q)
    assert np.abs(got_p - ref_p).max() < 1e-5
    assert max(quat_angle(a, b) for a, b in zip(got_q, ref_q)) < 1e-5
