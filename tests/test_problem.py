"""CPU: host logic — flattening of reference-style containers into the C-ABI batches (what the
C++ shim does), sharding of independent units, generator determinism."""
import numpy as np
import pytest

from rspl_slam_b200 import synth
from rspl_slam_b200.problem import FrameBatch, LocalBatch, shard_range


def test_generator_is_seed_deterministic_and_c1_shaped():
    a = synth.make_local_problem(synth.config_seed(1, 0))
    b = synth.make_local_problem(synth.config_seed(1, 0))
    for k, v in a.__dict__.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, getattr(b, k)), k
    assert len(a.pose_id) == 10 and a.pose_fixed.sum() == 1 and a.pose_fixed[0] == 1
    assert 2900 <= len(a.point_id) <= 3000 and 280 <= len(a.line_id) <= 300
    n_pt = len(a.mp_id_pose) + len(a.sp_id_pose)
    n_ln = len(a.ml_id_pose) + len(a.sl_id_pose)
    assert 10500 <= n_pt <= 13500 and 800 <= n_ln <= 1200  # ~4 obs/point, ~3.5 obs/line
    assert 0.80 < len(a.sp_id_pose) / n_pt < 0.90  # 85 % stereo
    # landmark-major emission (map.cc:609-658): each point's stereo observations are contiguous
    ids = a.sp_id_point
    change = np.nonzero(np.diff(ids) != 0)[0]
    assert len(change) + 1 == len(np.unique(ids))
    # every point keeps >= 1 stereo or >= 2 mono observations (map.cc:651)
    st = np.isin(a.point_id, a.sp_id_point)
    mono_cnt = np.array([(a.mp_id_point == i).sum() for i in a.point_id[~st]])
    assert (mono_cnt >= 2).all()


def test_local_flatten_compacts_ids_and_roundtrips():
    probs = [synth.make_local_problem(synth.config_seed(1, 10 + i), n_kf=4 + i, n_points=60, n_lines=8,
                                      first_kf_id=3 * i) for i in range(3)]
    batch = LocalBatch.from_problems(probs)
    assert batch.n_windows == 3 and batch.pose_twc.shape[0] == 7 and batch.sl_meas.shape[0] == 8
    for w, p in enumerate(probs):
        a, b = batch.stereo_pt_begin[w], batch.stereo_pt_begin[w + 1]
        # local indices address the window's id-sorted vertex slices
        assert np.array_equal(p.pose_id[batch.sp_pose[a:b]], p.sp_id_pose)
        assert np.array_equal(p.point_id[batch.sp_point[a:b]], p.sp_id_point)
        np.testing.assert_array_equal(batch.sp_meas[:, a:b].T, p.sp_kp)
        pa, pb = batch.pose_begin[w], batch.pose_begin[w + 1]
        np.testing.assert_array_equal(batch.pose_twc[:3, pa:pb].T, p.pose_p)
        np.testing.assert_array_equal(batch.pose_twc[3:, pa:pb].T, p.pose_q)
    sub = batch.slice(1, 3)
    ref = LocalBatch.from_problems(probs[1:])
    for k, v in sub.__dict__.items():
        assert np.array_equal(v, getattr(ref, k)), k
    assert batch.n_edges == sum(p.n_edges for p in probs)
    bad = probs[0].copy()
    bad.sp_id_point[0] = 10**6
    with pytest.raises(KeyError):
        LocalBatch.from_problems([bad])


def test_frame_flatten_and_reexpand():
    probs = [synth.make_frame_problem(synth.config_seed(2, i), n_points=50 + 7 * i, stereo_frac=0.7) for i in range(4)]
    batch = FrameBatch.from_problems(probs)
    direct = synth.make_frame_batch(2, 4, n_points=50, stereo_frac=0.7)
    assert direct.n_frames == 4
    for f, p in enumerate(probs):
        q = batch.frame_problem(f)
        np.testing.assert_array_equal(q.sp_kp, p.sp_kp)
        # Xw of every edge is the point the constraint's id refers to
        idx = np.searchsorted(p.point_id, p.sp_id_point)
        np.testing.assert_array_equal(batch.stereo_xw[:, batch.stereo_begin[f]:batch.stereo_begin[f + 1]].T, p.point_p[idx])
    # make_frame_batch == from_problems(make_frame_problem) on the same seeds
    same = FrameBatch.from_problems([synth.make_frame_problem(synth.config_seed(2, i), n_points=50, stereo_frac=0.7) for i in range(4)])
    for k, v in direct.__dict__.items():
        assert np.array_equal(v, getattr(same, k)), k


@pytest.mark.parametrize("n,world", [(4096, 8), (1024, 4), (10, 3), (3, 8), (0, 2)])
def test_shard_range_partitions_units(n, world):
    ranges = [shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1
