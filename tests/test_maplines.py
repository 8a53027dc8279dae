"""batched Map::UppdateMapline (/root/reference/src/map.cc:121-177), the endpoint refresh that
follows the local BA. CPU: the oracle against hand-worked answers, the reference's DBL_MIN quirk and an independent
numpy restatement (closed-form anchor instead of the LDLT solve; the reference ships no fixtures and g2o / Eigen are
not vendored: parity unpinned like the rest of the oracle). GPU: the kernel behind rspl_ba_update_maplines against the
oracle, BIT-EXACT (both sides are written in the reference's operation order without FMA contraction)."""
import numpy as np
import pytest

from oracle import orc
from rspl_slam_b200 import synth
from rspl_slam_b200.geometry import line_from_cartesian


def _np_update(b):
    """independent restatement: anchor = d x w / (|d|^2 + 1e-9), distances and ends vectorised per line"""
    wd = b["line_wd"].T
    n = len(wd)
    ends, ok, margin = np.zeros((6, n)), np.zeros(n, dtype=np.uint8), np.full(n, np.inf)
    for l in range(n):
        idx = b["pt_index"][b["pt_begin"][l]:b["pt_begin"][l + 1]]
        if len(idx) == 0:
            continue
        w, d = wd[l, :3], wd[l, 3:]
        v = d / np.linalg.norm(d)
        anchor = np.cross(d, w) / (d @ d + 1e-9)
        X = b["point_xyz"][:, idx].T
        dist = np.linalg.norm(np.cross(v, X - anchor), axis=1)
        margin[l] = np.abs(dist - 0.2).min()
        near = X[dist <= 0.2]
        md = int(np.argmax(np.abs(v)))
        hi = near[:, md][near[:, md] > np.finfo(np.float64).tiny]  # (sic) running maximum starts at DBL_MIN
        if len(near) == 0 or len(hi) == 0:
            continue
        ends[:3, l] = anchor + (hi.max() - anchor[md]) / v[md] * v
        ends[3:, l] = anchor + (near[:, md].min() - anchor[md]) / v[md] * v
        ok[l] = 1
    return ends, ok, margin


def test_line_to_cartesian_is_the_closest_point_and_unit_direction():
    rng = np.random.default_rng(5)
    for _ in range(200):
        p, v = rng.uniform(-8, 8, 3), rng.normal(size=3)
        scale = rng.uniform(0.3, 3.0)
        wd = line_from_cartesian(p, v) * scale
        cart = orc.line_to_cartesian(wd)
        u = v / np.linalg.norm(v)
        assert np.abs(cart[3:] - u).max() < 1e-14  # d / |d|
        # the point of the line closest to the origin, shrunk by the 1e-9 damping of the normal equations
        foot = (p - u * (u @ p)) * scale**2 / (scale**2 + 1e-9)
        # those equations are singular along the direction up to the damping: the LDLT solve may slide the anchor
        # along the line by rounding / 1e-9, never off it
        off = cart[:3] - foot
        assert np.linalg.norm(off - u * (u @ off)) < 1e-12 and abs(u @ off) < 1e-5


def test_oracle_hand_worked_segment_and_gate():
    # the line through (0, 1, 2) along +x; points at x = 1, 3, 5 on it, one at x = 9 but 0.5 m off, one 0.15 m off at x = 7
    wd = line_from_cartesian(np.array([0.0, 1.0, 2.0]), np.array([1.0, 0.0, 0.0]))
    pts = np.array([[1, 1, 2], [3, 1, 2], [5, 1, 2], [9, 1.5, 2], [7, 1, 2.15]], dtype=np.float64).T
    ends, ok, cnt = orc.update_maplines(wd.reshape(6, 1), [0, 5], [3, 0, 4, 2, 1], pts)
    assert cnt == 1 and ok[0] == 1
    # first the end at the largest main coordinate; 1e-8: toCartesian's 1e-9 damping pulls the anchor towards the origin
    assert np.abs(ends[:, 0] - [7, 1, 2, 1, 1, 2]).max() < 1e-8
    # without the point at x = 7 the far end is x = 5: the 0.5 m point is gated out (dist > 0.2, map.cc:155)
    ends, ok, _ = orc.update_maplines(wd.reshape(6, 1), [0, 4], [3, 0, 2, 1], pts)
    assert np.abs(ends[:, 0] - [5, 1, 2, 1, 1, 2]).max() < 1e-8


def test_oracle_keeps_the_reference_quirks():
    sentinel = np.full((6, 3), 42.0)
    wd = np.stack([line_from_cartesian(np.array([0.0, 1.0, 2.0]), np.array([1.0, 0.0, 0.0]))] * 3).T
    # line 0: all main coordinates negative -> `di > DBL_MIN` never holds, no maximum, not refreshed (map.cc:150-166)
    # line 1: no points at all (map.cc:127 / :166); line 2: one negative and one positive coordinate -> refreshed
    pts = np.array([[-4, 1, 2], [-2, 1, 2], [3, 1, 2]], dtype=np.float64).T
    ends, ok, cnt = orc.update_maplines(wd, [0, 2, 2, 4], [0, 1, 1, 2], pts, endpoints_init=sentinel)
    assert list(ok) == [0, 0, 1] and cnt == 1
    assert np.array_equal(ends[:, :2], sentinel[:, :2])  # SetEndpoints is not called
    assert np.abs(ends[:, 2] - [3, 1, 2, -2, 1, 2]).max() < 1e-8
    # a single usable point gives both ends
    ends, ok, _ = orc.update_maplines(wd[:, :1], [0, 1], [2], pts)
    assert ok[0] == 1 and np.abs(ends[:, 0] - [3, 1, 2, 3, 1, 2]).max() < 1e-8


def test_oracle_matches_the_numpy_restatement():
    b = synth.make_mapline_batch(20261021, n_lines=800)
    ends, ok, cnt = orc.update_maplines(b["line_wd"], b["pt_begin"], b["pt_index"], b["point_xyz"])
    ends2, ok2, margin = _np_update(b)
    clear = margin > 1e-9  # no point within rounding of the 0.2 m gate
    assert clear.sum() > 700 and np.array_equal(ok[clear], ok2[clear]) and cnt == int(ok.sum())
    assert 0.5 < ok.mean() < 0.98  # the batch has refreshed lines, empty lines and quirk lines
    good = clear & ok.astype(bool)
    assert np.abs(ends[:, good] - ends2[:, good]).max() < 1e-9
    # the refreshed segment lies on the true line and inside the true segment's span (+ the gate)
    p1, p2 = b["p1"][good], b["p2"][good]
    u = (p2 - p1) / np.linalg.norm(p2 - p1, axis=1, keepdims=True)
    for e in (ends[:3, good].T, ends[3:, good].T):
        rel = e - p1
        assert np.abs(rel - u * (rel * u).sum(1, keepdims=True)).max() < 1e-6  # (the damping again, |d| down to 0.5)


@pytest.mark.gpu
def test_update_maplines_matches_oracle_bit_exactly(gpu_ctx):
    b = synth.make_mapline_batch(20261022, n_lines=30000, max_pts=40)
    init = np.full((6, 30000), -3.5)
    ref_ends, ref_ok, ref_cnt = orc.update_maplines(b["line_wd"], b["pt_begin"], b["pt_index"], b["point_xyz"], endpoints_init=init)
    ends, ok, cnt = gpu_ctx.update_maplines(b["line_wd"], b["pt_begin"], b["pt_index"], b["point_xyz"], endpoints_init=init)
    assert np.array_equal(ok, ref_ok) and cnt == ref_cnt and 0 < cnt < 30000
    assert np.array_equal(ends, ref_ends)  # bit-exact, the untouched entries included


@pytest.mark.gpu
def test_update_maplines_edge_cases(gpu_ctx):
    wd = np.stack([line_from_cartesian(np.array([0.0, 1.0, 2.0]), np.array([1.0, 0.0, 0.0]))] * 3).T
    pts = np.array([[-4, 1, 2], [-2, 1, 2], [3, 1, 2]], dtype=np.float64).T
    sentinel = np.full((6, 3), 42.0)
    ends, ok, cnt = gpu_ctx.update_maplines(wd, [0, 2, 2, 4], [0, 1, 1, 2], pts, endpoints_init=sentinel)
    assert list(ok) == [0, 0, 1] and cnt == 1 and np.array_equal(ends[:, :2], sentinel[:, :2])
    assert np.abs(ends[:, 2] - [3, 1, 2, -2, 1, 2]).max() < 1e-8
    # empty batch; lines without any point array
    ends, ok, cnt = gpu_ctx.update_maplines(np.zeros((6, 0)), [0], np.zeros(0, np.int32), np.zeros((3, 0)))
    assert ends.shape == (6, 0) and cnt == 0
    ends, ok, cnt = gpu_ctx.update_maplines(wd, [0, 0, 0, 0], np.zeros(0, np.int32), np.zeros((3, 0)))
    assert cnt == 0 and not ok.any()
    from rspl_slam_b200.capi import RsplBaError
    with pytest.raises(RsplBaError):
        gpu_ctx.update_maplines(wd, [0, 2, 2, 4], [0, 1, 1, 7], pts)  # point index out of range
    with pytest.raises(RsplBaError):
        gpu_ctx.update_maplines(wd, [0, 2, 1, 4], [0, 1, 1, 2], pts)  # offsets not monotone


def _nearest_point_lists(batch, res, k=12):
    """for every line of a solved local batch: the k map points of its own window closest to the optimised line, as
    batch-wide indices (what Map::UppdateMapline gathers are the map points on the line in its observers)"""
    begin, index = [0], []
    for w in range(batch.n_windows):
        p0, p1 = int(batch.point_begin[w]), int(batch.point_begin[w + 1])
        X = res.point_xyz[:, p0:p1].T
        for l in range(int(batch.line_begin[w]), int(batch.line_begin[w + 1])):
            cart = orc.line_to_cartesian(res.line_wd[:, l])
            dist = np.linalg.norm(np.cross(cart[3:], X - cart[:3]), axis=1)
            near = np.argsort(dist)[:min(k, len(X)) if l % 7 else 0]  # every seventh line has no points
            index.extend((p0 + near).tolist())
            begin.append(len(index))
    return np.asarray(begin, dtype=np.int32), np.asarray(index, dtype=np.int32)


@pytest.mark.gpu
def test_update_maplines_on_the_resident_local_ba_result(gpu_ctx):
    """rspl_ba_local_batch_update_maplines reads the optimised lines and points where the solve left them in HBM: the
    same bits as the host-array call on the downloaded result (and as the oracle), zeros where not refreshed."""
    from rspl_slam_b200.problem import LocalBatch
    probs = [synth.make_local_problem(synth.config_seed(1, 700 + i), n_kf=4 + i, n_points=400 + 100 * i, n_lines=40 + 10 * i) for i in range(4)]
    batch = LocalBatch.from_problems(probs)
    out = gpu_ctx.alloc_local_result(batch)
    gpu_ctx.local_batch_upload(batch)
    gpu_ctx.local_batch_solve()
    gpu_ctx.local_batch_download(out)
    begin, index = _nearest_point_lists(batch, out)
    ends, ok, cnt = gpu_ctx.local_update_maplines(begin, index)
    ref_ends, ref_ok, ref_cnt = orc.update_maplines(out.line_wd, begin, index, out.point_xyz)
    assert np.array_equal(ok, ref_ok) and cnt == ref_cnt and 0 < cnt < len(ok)
    assert np.array_equal(ends, ref_ends)  # (the oracle wrapper starts from zeros as well)
    host_ends, host_ok, _ = gpu_ctx.update_maplines(out.line_wd, begin, index, out.point_xyz)
    assert np.array_equal(host_ends, ends) and np.array_equal(host_ok, ok)
    from rspl_slam_b200.capi import RsplBaError
    bad = index.copy()
    bad[0] = out.point_xyz.shape[1]
    with pytest.raises(RsplBaError):
        gpu_ctx.local_update_maplines(begin, bad)
