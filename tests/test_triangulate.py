"""batched Map::TriangulateMappoint (/root/reference/src/map.cc:292-339).
CPU: the oracle against closed-form known answers and an independent numpy restatement (the reference ships no
fixtures for it; Eigen is not vendored: parity unpinned like the rest of the oracle). GPU: the kernel behind
rspl_ba_triangulate_points against the oracle on a seeded batch with missing, single and degenerate observations."""
import numpy as np
import pytest

from oracle import orc
from rspl_slam_b200 import synth
from rspl_slam_b200.geometry import quat_to_R


def _np_triangulate(b):
    """independent restatement: normal equations of map.cc:320-326 solved by numpy, rank by singular values"""
    fx, fy, cx, cy, _ = b["cam5"]
    n = len(b["obs_begin"]) - 1
    xyz, ok, ratio = np.zeros((3, n)), np.zeros(n, dtype=np.uint8), np.zeros(n)
    for i in range(n):
        o0, o1 = b["obs_begin"][i], b["obs_begin"][i + 1]
        if o1 - o0 < 2:
            continue
        A, rhs = (o1 - o0) * np.eye(3), np.zeros(3)
        for o in range(o0, o1):
            f = b["obs_frame"][o]
            R = quat_to_R(b["frame_twc"][3:, f])
            c = b["frame_twc"][:3, f]
            bv = R @ np.array([(b["obs_uv"][0, o] - cx) / fx, (b["obs_uv"][1, o] - cy) / fy, 1.0])
            A -= np.outer(bv, bv) / (bv @ bv)
            rhs += c - bv * (bv @ c) / (bv @ bv)
        sv = np.linalg.svd(A, compute_uv=False)
        ratio[i] = sv[-1] / sv[0]
        if ratio[i] <= 1e-5:
            continue
        xyz[:, i] = np.linalg.solve(A, rhs)
        ok[i] = 1
    return xyz, ok, ratio


def test_oracle_recovers_exact_intersections_and_rejects_degenerate_points():
    b = synth.make_triangulation_batch(20261018, n_points=400, pixel_sigma=0.0, degenerate_frac=0.1)
    sentinel = np.full((3, 400), 123.5)
    xyz, ok, cnt = orc.triangulate_points(b["obs_begin"], b["obs_frame"], b["obs_uv"], b["frame_twc"], b["cam5"], xyz_init=sentinel)
    nobs = np.diff(b["obs_begin"])
    assert cnt == int(ok.sum()) and 0 < cnt < 400
    assert not ok[nobs < 2].any()  # map.cc:317
    good = ok.astype(bool)
    # noise-free rays meet in the point: known answer
    assert np.abs(xyz[:, good].T - b["truth"][good]).max() < 1e-8
    # a failed point keeps its position (the reference does not call SetPosition)
    assert np.array_equal(xyz[:, ~good], sentinel[:, ~good])
    # every point whose observations all come from one keyframe has parallel rays: rank 2, rejected
    for i in np.flatnonzero(nobs >= 2):
        fr = b["obs_frame"][b["obs_begin"][i]:b["obs_begin"][i + 1]]
        if len(set(fr.tolist())) == 1:
            assert not ok[i]


def test_oracle_matches_the_numpy_restatement():
    b = synth.make_triangulation_batch(20261019, n_points=600, pixel_sigma=1.0)
    xyz, ok, _ = orc.triangulate_points(b["obs_begin"], b["obs_frame"], b["obs_uv"], b["frame_twc"], b["cam5"])
    xyz2, ok2, ratio = _np_triangulate(b)
    # the pivot ratio |R_33| / max |R_ii| of the QR and the singular-value ratio agree within a small factor: the rank
    # decisions must coincide away from the threshold
    clear = (ratio > 1e-4) | (ratio < 1e-6)
    assert clear.sum() > 400 and np.array_equal(ok[clear], ok2[clear])
    good = ok.astype(bool) & ok2.astype(bool)
    cond = 1.0 / ratio[good]
    err = np.abs(xyz[:, good] - xyz2[:, good]).max(axis=0) / np.maximum(1.0, np.abs(xyz2[:, good]).max(axis=0))
    assert (err < 1e-13 * cond + 1e-12).all()  # two exact solvers of the same 3 x 3 system: rounding x condition number


def test_oracle_two_view_known_answer():
    # two cameras 1 m apart on the x axis, both looking down +z; the point (0.5, 0, 4) projects symmetrically
    cam5 = np.array([400.0, 400.0, 320.0, 240.0, 40.0])
    twc = np.zeros((7, 2))
    twc[6] = 1.0
    twc[0, 1] = 1.0
    uv = np.array([[320.0 + 400.0 * 0.5 / 4, 320.0 - 400.0 * 0.5 / 4], [240.0, 240.0]])
    xyz, ok, cnt = orc.triangulate_points([0, 2], [0, 1], uv, twc, cam5)
    assert cnt == 1 and ok[0] == 1 and np.abs(xyz[:, 0] - [0.5, 0.0, 4.0]).max() < 1e-12
    # one observation only
    _, ok1, cnt1 = orc.triangulate_points([0, 1], [0], uv[:, :1], twc, cam5)
    assert cnt1 == 0 and ok1[0] == 0


@pytest.mark.gpu
def test_triangulate_points_matches_oracle(gpu_ctx):
    b = synth.make_triangulation_batch(20261020, n_points=20000, n_frames=16, max_obs=8, pixel_sigma=1.0)
    init = np.full((3, 20000), -7.25)
    ref_xyz, ref_ok, ref_cnt = orc.triangulate_points(b["obs_begin"], b["obs_frame"], b["obs_uv"], b["frame_twc"], b["cam5"], xyz_init=init)
    xyz, ok, cnt = gpu_ctx.triangulate_points(b["obs_begin"], b["obs_frame"], b["obs_uv"], b["frame_twc"], b["cam5"], xyz_init=init)
    assert np.array_equal(ok, ref_ok) and cnt == ref_cnt  # bit-exact flags
    good = ok.astype(bool)
    scale = np.maximum(1.0, np.abs(ref_xyz[:, good]))
    assert (np.abs(xyz[:, good] - ref_xyz[:, good]) / scale).max() < 1e-9  # fp64, same algorithm: rounding only
    assert np.array_equal(xyz[:, ~good], init[:, ~good])


@pytest.mark.gpu
def test_triangulate_points_edge_cases(gpu_ctx):
    cam5 = np.array([400.0, 400.0, 320.0, 240.0, 40.0])
    twc = np.zeros((7, 2))
    twc[6] = 1.0
    twc[0, 1] = 1.0
    # empty batch, a point without observations, and the two-view known answer
    xyz, ok, cnt = gpu_ctx.triangulate_points([0], np.zeros(0, np.int32), np.zeros((2, 0)), twc, cam5)
    assert xyz.shape == (3, 0) and cnt == 0
    uv = np.array([[370.0, 270.0], [240.0, 240.0]])
    xyz, ok, cnt = gpu_ctx.triangulate_points([0, 0, 2], [0, 1], uv, twc, cam5)
    assert list(ok) == [0, 1] and cnt == 1 and np.abs(xyz[:, 1] - [0.5, 0.0, 4.0]).max() < 1e-12
    from rspl_slam_b200.capi import RsplBaError
    with pytest.raises(RsplBaError):
        gpu_ctx.triangulate_points([0, 2], [0, 5], uv, twc, cam5)  # keyframe index out of range
