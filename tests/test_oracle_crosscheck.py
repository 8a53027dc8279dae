"""CPU: the C++ oracle against an independently written numpy restatement (oracle/np_oracle.py:
(R, t) poses with Rodrigues updates, difference-quotient Jacobians for every edge type, one dense
damped normal-equation solve instead of block Schur). The reference gives no golden vectors and
g2o cannot be built here, so two restatements agreeing step by step is the strongest check of the
oracle available. Agreement bar: identical inlier sets, identical accept / reject
pattern, chi2 of every LM trial within 1e-6 relative, poses within 1e-7."""
import numpy as np
import pytest

from oracle import np_oracle
from rspl_slam_b200 import synth
from rspl_slam_b200.geometry import quat_angle


@pytest.mark.parametrize("inst,kw", [
    (0, dict(n_kf=4, n_points=40, n_lines=6)),
    (1, dict(n_kf=5, n_points=60, n_lines=8, first_kf_id=7)),
    (2, dict(n_kf=3, n_points=30, n_lines=0, stereo_point_frac=0.5)),
    (3, dict(n_kf=4, n_points=0, n_lines=14)),
])
def test_local_ba_two_restatements_agree(orc, inst, kw):
    p = synth.make_local_problem(synth.config_seed(1, 7000 + inst), **kw)
    a, b = p.copy(), p.copy()
    st = orc.local_ba(a, orc.make_config(numeric_delta=1e-6), trace=True)
    tr = np_oracle.local_ba(b)
    for f in ("mp_inlier", "sp_inlier", "ml_inlier", "sl_inlier"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert len(tr) == len(st["trace"])
    for r, t in zip(st["trace"], tr):
        assert (r["iter"], r["trial"], r["accepted"]) == (t[1], t[2], t[3])
        assert abs(r["chi_before"] - t[4]) <= 1e-6 * max(t[4], 1.0)
        assert abs(r["chi_after"] - t[5]) <= 1e-6 * max(t[5], 1.0)
        assert abs(r["lambda_"] - t[6]) <= 1e-5 * max(t[6], 1e-12)
    assert np.abs(a.pose_p - b.pose_p).max() < 1e-7
    assert quat_angle(a.pose_q, b.pose_q).max() < 1e-7
    if len(p.point_id):
        assert np.quantile(np.abs(a.point_p - b.point_p).max(axis=1), 0.9) < 1e-6


@pytest.mark.parametrize("inst,kw", [(0, dict(n_points=60)), (1, dict(n_points=40, stereo_frac=0.5)),
                                      (2, dict(n_points=7)), (3, dict(n_points=80, outlier_frac=0.2)),
                                      # the line extension of the pose-only path (constraints on fixed lines)
                                      (4, dict(n_points=50, n_lines=20)), (5, dict(n_points=6, n_lines=30, outlier_frac=0.15))])
def test_frame_optimization_two_restatements_agree(orc, inst, kw):
    p = synth.make_frame_problem(synth.config_seed(2, 7000 + inst), **kw)
    a, b = p.copy(), p.copy()
    st = orc.frame_opt(a, trace=True)
    ret, tr = np_oracle.frame_opt(b)
    assert ret == st["ret"]
    assert np.array_equal(a.sp_inlier, b.sp_inlier) and np.array_equal(a.mp_inlier, b.mp_inlier)
    assert np.array_equal(a.sl_inlier, b.sl_inlier) and np.array_equal(a.ml_inlier, b.ml_inlier)
    assert np.linalg.norm(a.pose_p - b.pose_p) < 1e-7 and quat_angle(a.pose_q, b.pose_q) < 1e-7
    # until the iteration has converged to rounding noise both restatements take the same decisions
    for r, t in list(zip(st["trace"], tr))[:3]:
        assert (r["iter"], r["trial"], r["accepted"]) == (t[1], t[2], t[3])
        assert abs(r["chi_after"] - t[5]) <= 1e-6 * max(t[5], 1.0)
