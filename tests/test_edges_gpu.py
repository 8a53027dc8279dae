"""GPU parity, unit level: per-edge residuals / Jacobians / manifold updates of the CUDA path
(through the C-ABI: rspl_ba_eval_edges, rspl_ba_oplus) against the CPU oracle.

Tolerances: residuals and point Jacobians are the same closed forms in fp64 -> 1e-9 relative.
Line Jacobians are analytic on the device but numeric central differences (delta = 1e-9, as g2o,
SURVEY §9.8) in the oracle, whose own rounding noise is ~1e-16/1e-9 * |r| ~ 1e-6 absolute -> they
are compared at 2e-5 of the Jacobian scale, and additionally against a delta = 1e-6 difference
quotient of the oracle's residual (truncation error O(1e-12)) at 1e-7.
"""
import numpy as np
import pytest

from rspl_slam_b200 import synth
from rspl_slam_b200.geometry import R_to_quat, line_from_cartesian, rotvec_to_R
from rspl_slam_b200.problem import EUROC_CAMERA

pytestmark = pytest.mark.gpu


def _random_edges(orc, rng, n):
    cam = EUROC_CAMERA
    b = cam[4] / cam[0]
    poses, pts, lines, mpt, mln = [], [], [], [], []
    for _ in range(n):
        Rwc = rotvec_to_R(rng.normal(0, 0.4, 3)) @ synth.R_WC0
        twc = rng.normal(0, 1.0, 3)
        poses.append(orc.pose_from_twc(twc, R_to_quat(Rwc)))
        z = rng.uniform(1, 10)
        Xc = np.array([rng.uniform(-0.8, 0.8) * z, rng.uniform(-0.5, 0.5) * z, z])
        Xw = Rwc @ Xc + twc
        pts.append(Xw)
        u = cam[0] * Xc[0] / z + cam[2]
        v = cam[1] * Xc[1] / z + cam[3]
        mpt.append(np.array([u, v, u - cam[4] / z]) + rng.normal(0, 2.0, 3))
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        P2c = Xc + d * rng.uniform(0.5, 2.0)
        P2c[2] = max(P2c[2], 0.8)
        lines.append(line_from_cartesian(Xw, Rwc @ (P2c - Xc)))
        m = []
        for shift in (0.0, b):
            for P in (Xc, P2c):
                m += [cam[0] * (P[0] - shift) / P[2] + cam[2], cam[1] * P[1] / P[2] + cam[3]]
        mln.append(np.array(m) + rng.normal(0, 2.0, 8))
    return np.array(poses), np.array(pts), np.array(lines), np.array(mpt), np.array(mln)


@pytest.mark.parametrize("edge_type", [0, 1, 2, 3])
def test_edge_residual_and_jacobian_match_oracle(gpu_ctx, orc, edge_type):
    rng = np.random.default_rng(100 + edge_type)
    n = 64
    poses, pts, lines, mpt, mln = _random_edges(orc, rng, n)
    lm = pts if edge_type < 2 else lines
    meas = {0: mpt[:, :2], 1: mpt, 2: mln[:, :4], 3: mln}[edge_type]
    err, Jl, Jp, chi2 = gpu_ctx.eval_edges(edge_type, poses, lm, meas, EUROC_CAMERA)
    dim = (2, 3, 2, 4)[edge_type]
    ld = 3 if edge_type < 2 else 4
    info = 1.0 if edge_type < 2 else 0.1
    for i in range(n):
        e, jl, jp = orc.edge_eval(edge_type, poses[i], lm[i], meas[i], EUROC_CAMERA)
        np.testing.assert_allclose(err[i, :dim], e, rtol=1e-9, atol=1e-9)
        assert abs(chi2[i] - info * float(e @ e)) <= 1e-9 * max(1.0, chi2[i])
        g_jl = Jl[i, :dim * ld].reshape(dim, ld)
        g_jp = Jp[i, :dim * 6].reshape(dim, 6)
        if edge_type < 2:
            np.testing.assert_allclose(g_jl, jl, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(g_jp, jp, rtol=1e-9, atol=1e-9)
        else:
            scale = max(np.abs(jp).max(), np.abs(jl).max(), 1.0)
            assert np.abs(g_jl - jl).max() <= 2e-5 * scale
            assert np.abs(g_jp - jp).max() <= 2e-5 * scale
            # sharper check against a delta = 1e-6 central difference of the oracle residual
            h = 1e-6
            for d in range(6):
                u = np.zeros(6)
                u[d] = h
                ep = orc.edge_eval(edge_type, orc.pose_oplus(poses[i], u), lm[i], meas[i], EUROC_CAMERA)[0]
                em = orc.edge_eval(edge_type, orc.pose_oplus(poses[i], -u), lm[i], meas[i], EUROC_CAMERA)[0]
                assert np.abs((ep - em) / (2 * h) - g_jp[:, d]).max() <= 1e-7 * scale
            for d in range(4):
                v = np.zeros(4)
                v[d] = h
                ep = orc.edge_eval(edge_type, poses[i], orc.line_oplus(lm[i], v), meas[i], EUROC_CAMERA)[0]
                em = orc.edge_eval(edge_type, poses[i], orc.line_oplus(lm[i], -v), meas[i], EUROC_CAMERA)[0]
                assert np.abs((ep - em) / (2 * h) - g_jl[:, d]).max() <= 1e-7 * scale


@pytest.mark.parametrize("edge_type", [4, 5])
def test_pose_only_edges_match_oracle(gpu_ctx, orc, edge_type):
    """g2o::EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose (the edges of FrameOptimization,
    g2o_optimization.cc:288-333): residual and 2x6 / 3x6 pose Jacobian with the world point held in the edge."""
    rng = np.random.default_rng(200 + edge_type)
    n = 64
    poses, pts, _, mpt, _ = _random_edges(orc, rng, n)
    meas = mpt[:, :2] if edge_type == 4 else mpt
    err, Jl, Jp, chi2 = gpu_ctx.eval_edges(edge_type, poses, pts, meas, EUROC_CAMERA)
    dim = 2 if edge_type == 4 else 3
    assert not Jl.any()
    for i in range(n):
        e, _, jp = orc.edge_eval(edge_type, poses[i], pts[i], meas[i], EUROC_CAMERA)
        np.testing.assert_allclose(err[i, :dim], e, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(Jp[i, :dim * 6].reshape(dim, 6), jp, rtol=1e-9, atol=1e-9)
        assert abs(chi2[i] - float(e @ e)) <= 1e-9 * max(1.0, chi2[i])
        # and against a central difference of the oracle's residual through exp(u) * T
        h = 1e-6
        for d in range(6):
            u = np.zeros(6)
            u[d] = h
            ep = orc.edge_eval(edge_type, orc.pose_oplus(poses[i], u), pts[i], meas[i], EUROC_CAMERA)[0]
            em = orc.edge_eval(edge_type, orc.pose_oplus(poses[i], -u), pts[i], meas[i], EUROC_CAMERA)[0]
            assert np.abs((ep - em) / (2 * h) - Jp[i, :dim * 6].reshape(dim, 6)[:, d]).max() <= 1e-6 * max(np.abs(jp).max(), 1.0)


def test_stereo_bf_float_switch(gpu_ctx, orc):
    rng = np.random.default_rng(7)
    poses, pts, _, mpt, _ = _random_edges(orc, rng, 8)
    for flag in (0, 1):
        err, *_ = gpu_ctx.eval_edges(1, poses, pts, mpt, EUROC_CAMERA, stereo_bf_float=flag)
        for i in range(8):
            e, _, _ = orc.edge_eval(1, poses[i], pts[i], mpt[i], EUROC_CAMERA, stereo_bf_float=flag)
            np.testing.assert_allclose(err[i, :3], e, rtol=1e-12, atol=1e-11)


def test_manifold_updates_match_oracle(gpu_ctx, orc):
    rng = np.random.default_rng(11)
    n = 50
    poses, pts, lines, _, _ = _random_edges(orc, rng, n)
    u = rng.normal(0, 0.05, (n, 6))
    u[:5] *= 1e-7  # small-angle branch of SE3Quat::exp (theta < 1e-5)
    out = gpu_ctx.oplus(0, poses, u)
    for i in range(n):
        np.testing.assert_allclose(out[i], orc.pose_oplus(poses[i], u[i]), rtol=0, atol=1e-13)
    v = rng.normal(0, 0.02, (n, 4))
    out = gpu_ctx.oplus(2, lines, v)
    for i in range(n):
        np.testing.assert_allclose(out[i, :6], orc.line_oplus(lines[i], v[i]), rtol=0, atol=1e-12)
    out = gpu_ctx.oplus(1, pts, u[:, :3])
    np.testing.assert_allclose(out[:, :3], pts + u[:, :3], rtol=0, atol=0)


@pytest.mark.gpu
def test_range_test_free_reciprocal_rsqrt_sqrt_accuracy(gpu_ctx):
    """ba_math.cuh: rcp_nr / rsqrt_nr / sqrt_nr (MUFU seed + FMA refinement without the CUDA library's range test and
    slow path, used in every hot loop) against the correctly rounded IEEE results over 560 decades: the reciprocal is
    correctly rounded on every sample, rsqrt / sqrt are within 2 ulp (like the CUDA library's own rsqrt)."""
    rng = np.random.default_rng(20261018)
    mant = rng.uniform(1.0, 2.0, 200000)
    expo = rng.integers(-930, 930, 200000)
    x = np.ldexp(mant, expo)
    x[:1000] = np.linspace(0.5, 20.0, 1000)  # depths
    for op, ref in ((0, 1.0 / x), (1, 1.0 / np.sqrt(x)), (2, np.sqrt(x))):
        got = gpu_ctx.unit_math(op, x)
        ulp = np.abs(got - ref) / np.spacing(np.abs(ref))
        assert np.isfinite(got).all() and ulp.max() <= (0.0 if op == 0 else 2.0), (op, float(ulp.max()))
    assert gpu_ctx.unit_math(2, np.array([0.0]))[0] == 0.0  # sqrt_nr is exact at 0
    neg = gpu_ctx.unit_math(0, -x[:100])
    assert np.abs(neg + 1.0 / x[:100]).max() <= np.spacing(1.0 / x[:100]).max()
