"""GPU parity: FrameOptimization (g2o_optimization.cc:256-397) through the C-ABI against the oracle.

Bar (BASELINE.json north_star): inlier/outlier index sets bit-exact, poses within 1e-5 m and
1e-5 rad, return value (#inliers) equal.
"""
import numpy as np
import pytest

from rspl_slam_b200 import capi, synth
from rspl_slam_b200.geometry import quat_angle
from rspl_slam_b200.problem import FrameBatch

pytestmark = pytest.mark.gpu

POS_TOL = 1e-5
ROT_TOL = 1e-5


def _check_against_oracle(orc, probs, batch, res, cfg=None):
    for f, p in enumerate(probs):
        q = p.copy()
        st = orc.frame_opt(q, cfg)
        m0, m1 = batch.mono_begin[f], batch.mono_begin[f + 1]
        s0, s1 = batch.stereo_begin[f], batch.stereo_begin[f + 1]
        assert np.array_equal(q.mp_inlier, res.mono_inlier[m0:m1]), f"frame {f}: mono inlier set differs"
        assert np.array_equal(q.sp_inlier, res.stereo_inlier[s0:s1]), f"frame {f}: stereo inlier set differs"
        if batch.mline_begin is not None:
            a0, a1 = batch.mline_begin[f], batch.mline_begin[f + 1]
            b0, b1 = batch.sline_begin[f], batch.sline_begin[f + 1]
            assert np.array_equal(q.ml_inlier, res.mline_inlier[a0:a1]), f"frame {f}: mono line inlier set differs"
            assert np.array_equal(q.sl_inlier, res.sline_inlier[b0:b1]), f"frame {f}: stereo line inlier set differs"
        assert int(res.num_inliers[f]) == st["ret"]
        assert np.linalg.norm(q.pose_p - res.pose_twc[:3, f]) < POS_TOL
        assert quat_angle(q.pose_q, res.pose_twc[3:, f]) < ROT_TOL


def test_frame_batch_matches_oracle_c2_shape(gpu_ctx, orc):
    """64 frames of the C2 shape (400 stereo points, 5 % gross outliers)."""
    probs = [synth.make_frame_problem(synth.config_seed(2, i)) for i in range(64)]
    batch = FrameBatch.from_problems(probs)
    res = gpu_ctx.frame_batch(batch)
    _check_against_oracle(orc, probs, batch, res)
    assert (res.stats["edges_linearized"] > 0).all()


def test_frame_latency_mode_matches_oracle(gpu_ctx, orc):
    """frame_latency_mode = 1 (one CTA per frame, what the shim's single-frame FrameOptimization uses): same parity
    bar, with and without the line extension, ragged / tiny frames included; agrees with the one-warp-per-frame
    mode to rounding."""
    opt = capi.make_options(frame_latency_mode=1)
    probs = [synth.make_frame_problem(synth.config_seed(2, 300 + i), n_points=n, stereo_frac=0.7)
             for i, n in enumerate([400, 37, 9, 3, 600, 250])]
    batch = FrameBatch.from_problems(probs)
    res = gpu_ctx.frame_batch(batch, opt)
    _check_against_oracle(orc, probs, batch, res)
    thr = gpu_ctx.frame_batch(batch)
    assert np.array_equal(thr.stereo_inlier, res.stereo_inlier) and np.abs(thr.pose_twc - res.pose_twc).max() < 1e-9
    lprobs = [synth.make_frame_problem(synth.config_seed(2, 320 + i), n_points=200, n_lines=40) for i in range(4)]
    lbatch = FrameBatch.from_problems(lprobs)
    _check_against_oracle(orc, lprobs, lbatch, gpu_ctx.frame_batch(lbatch, opt))
    # bitwise reproducible within the mode
    again = gpu_ctx.frame_batch(batch, opt)
    assert np.array_equal(again.pose_twc.view(np.uint64), res.pose_twc.view(np.uint64))


def test_frame_batch_mixed_mono_stereo_and_ragged(gpu_ctx, orc):
    """Ragged edge counts, mono+stereo mixes, tiny frames (<10 edges -> single round, :387), an
    empty frame, and caller-provided ->inlier = false flags."""
    rng = np.random.default_rng(5)
    probs = []
    for i, n in enumerate([400, 37, 9, 3, 123, 600, 1, 250]):
        p = synth.make_frame_problem(synth.config_seed(2, 100 + i), n_points=n, stereo_frac=float(rng.uniform(0.3, 1.0)))
        if i % 2 == 1 and len(p.sp_inlier):
            p.sp_inlier[rng.random(len(p.sp_inlier)) < 0.3] = 0
            p.mp_inlier[rng.random(len(p.mp_inlier)) < 0.3] = 0
        probs.append(p)
    empty = synth.make_frame_problem(synth.config_seed(2, 199), n_points=4)
    for k in ("mp_id_point", "mp_id_cam", "mp_inlier", "sp_id_point", "sp_id_cam", "sp_inlier"):
        setattr(empty, k, getattr(empty, k)[:0])
    empty.mp_kp, empty.sp_kp = empty.mp_kp[:0], empty.sp_kp[:0]
    probs.insert(3, empty)
    batch = FrameBatch.from_problems(probs)
    res = gpu_ctx.frame_batch(batch)
    _check_against_oracle(orc, probs, batch, res)
    # the empty frame returns its (normalised) input pose and 0 inliers
    assert res.num_inliers[3] == 0
    assert np.linalg.norm(res.pose_twc[:3, 3] - empty.pose_p) < 1e-12


def test_frame_batch_uma_thresholds_and_schedule(gpu_ctx, orc):
    """Other OptimizationConfig values (configs_uma_bumblebee_indoor.yaml: 25/37) and a shorter schedule."""
    from rspl_slam_b200.problem import OptimizationConfig
    cfg = OptimizationConfig(mono_point=25.0, stereo_point=37.0)
    probs = [synth.make_frame_problem(synth.config_seed(2, 300 + i), n_points=300, stereo_frac=0.8) for i in range(16)]
    batch = FrameBatch.from_problems(probs)
    res = gpu_ctx.frame_batch(batch, capi.make_options(cfg, frame_rounds=3, frame_iters=6))
    _check_against_oracle(orc, probs, batch, res, orc.make_config(cfg, rounds=3, iters_round=6))


def test_frame_batch_is_bitwise_deterministic_and_staged_equals_oneshot(gpu_ctx):
    batch = synth.make_frame_batch(2, 256, first_instance=500)
    a = gpu_ctx.frame_batch(batch)
    gpu_ctx.frame_batch_upload(batch)
    outs = []
    for _ in range(3):  # solve restarts from the uploaded inputs every time
        gpu_ctx.frame_batch_solve()
        outs.append(gpu_ctx.frame_batch_download(gpu_ctx.alloc_frame_result(batch)))
    for o in outs:
        assert np.array_equal(o.pose_twc.view(np.uint64), a.pose_twc.view(np.uint64))
        assert np.array_equal(o.stereo_inlier, a.stereo_inlier)
        assert np.array_equal(o.num_inliers, a.num_inliers)


def test_frame_batch_sharding_is_invariant(gpu_ctx):
    """Frames are independent units: solving shards [0,n/2) and [n/2,n) separately (what two ranks
    do) gives bit-identical results to the whole batch."""
    from rspl_slam_b200.problem import shard_range
    batch = synth.make_frame_batch(2, 96, first_instance=700)
    whole = gpu_ctx.frame_batch(batch)
    for world in (2, 4):
        poses, inl = [], []
        for r in range(world):
            a, b = shard_range(batch.n_frames, r, world)
            part = gpu_ctx.frame_batch(batch.slice(a, b))
            poses.append(part.pose_twc)
            inl.append(part.stereo_inlier)
        assert np.array_equal(np.concatenate(poses, axis=1).view(np.uint64), whole.pose_twc.view(np.uint64))
        assert np.array_equal(np.concatenate(inl), whole.stereo_inlier)


def test_frame_batch_full_c2_size_properties(gpu_ctx, orc):
    """BASELINE config C2 at full size (4096 frames x 400 stereo points): size-independent
    properties + oracle parity on a strided sample of frames."""
    batch = synth.make_frame_batch(2, 4096)
    res = gpu_ctx.frame_batch(batch)
    n = np.diff(batch.stereo_begin) + np.diff(batch.mono_begin)
    # return value = #edges - #outliers of the last round, and equals the inlier flag count
    # (points whose noisy right coordinate is <= 0 are mono observations, frame.cc:115)
    cs = np.concatenate([[0], np.cumsum(res.stereo_inlier, dtype=np.int64)])
    cm = np.concatenate([[0], np.cumsum(res.mono_inlier, dtype=np.int64)])
    cnt = np.diff(cs[batch.stereo_begin]) + np.diff(cm[batch.mono_begin])
    assert np.array_equal(cnt, res.num_inliers)
    assert (res.num_inliers <= n).all() and (res.num_inliers > 0.85 * n).all()
    # unit quaternions with w >= 0 (SE3Quat normalisation, g2o_optimization.cc:391-393)
    qn = np.linalg.norm(res.pose_twc[3:], axis=0)
    assert np.abs(qn - 1).max() < 1e-12 and (res.pose_twc[6] >= 0).all()
    # idempotence: re-optimising from the optimum with the found inlier flags keeps pose and flags
    again = FrameBatch(**{**batch.__dict__, "pose_twc": res.pose_twc.copy(), "stereo_inlier": res.stereo_inlier.copy()})
    res2 = gpu_ctx.frame_batch(again)
    assert np.abs(res2.pose_twc[:3] - res.pose_twc[:3]).max() < 1e-6
    assert (res2.stereo_inlier != res.stereo_inlier).mean() < 1e-4
    sample = list(range(0, 4096, 128))
    probs = [batch.frame_problem(f) for f in sample]
    for f, p in zip(sample, probs):
        st = orc.frame_opt(p)
        s0, s1 = batch.stereo_begin[f], batch.stereo_begin[f + 1]
        assert np.array_equal(p.sp_inlier, res.stereo_inlier[s0:s1])
        assert np.array_equal(p.mp_inlier, res.mono_inlier[batch.mono_begin[f]:batch.mono_begin[f + 1]])
        assert st["ret"] == res.num_inliers[f]
        assert np.linalg.norm(p.pose_p - res.pose_twc[:3, f]) < POS_TOL
        assert quat_angle(p.pose_q, res.pose_twc[3:, f]) < ROT_TOL


def test_invalid_inputs_are_rejected(gpu_ctx):
    batch = synth.make_frame_batch(2, 4, first_instance=800, n_points=20)
    bad = FrameBatch(**{**batch.__dict__, "stereo_begin": batch.stereo_begin[::-1].copy()})
    with pytest.raises(capi.RsplBaError) as e:
        gpu_ctx.frame_batch(bad)
    assert e.value.code == capi.RSPL_BA_ERR_INVALID
    bad = FrameBatch(**{**batch.__dict__, "stereo_cam": np.full_like(batch.stereo_cam, 3)})
    with pytest.raises(capi.RsplBaError) as e:
        gpu_ctx.frame_batch(bad)
    assert e.value.code == capi.RSPL_BA_ERR_INVALID
    with pytest.raises(capi.RsplBaError) as e:
        capi.Context(device=0).frame_batch_solve()
    assert e.value.code == capi.RSPL_BA_ERR_STATE
    # line extension: offsets must come in pairs and be monotone, arrays must be present, cameras in range
    lb = synth.make_frame_batch(2, 4, first_instance=810, n_points=20, n_lines=6)
    for patch in (dict(sline_begin=None), dict(mline_begin=lb.mline_begin[::-1].copy()), dict(sline_lw=None),
                  dict(mline_cam=np.full_like(lb.mline_cam, 7))):
        bad = FrameBatch(**{**lb.__dict__, **patch})
        with pytest.raises(capi.RsplBaError) as e:
            gpu_ctx.frame_batch(bad, out=gpu_ctx.alloc_frame_result(lb))
        assert e.value.code == capi.RSPL_BA_ERR_INVALID, patch
    res = gpu_ctx.alloc_frame_result(lb)
    res.sline_inlier = None  # result buffers of the line edges are mandatory when the batch has lines
    with pytest.raises(capi.RsplBaError) as e:
        gpu_ctx.frame_batch(lb, out=res)
    assert e.value.code == capi.RSPL_BA_ERR_INVALID


def test_frame_batch_with_line_extension(gpu_ctx, orc):
    """BASELINE config C2 in full: 400 stereo points + 60 lines per frame. The reference's FrameOptimization
    takes no lines (g2o_optimization.cc:284-285); the extension treats them as EdgeSE3ProjectLine /
    EdgeStereoSE3ProjectLine with the line vertex fixed (SURVEY 8a note), and the oracle does the same with
    g2o's numeric pose Jacobians. Also ragged / mono-only / line-only frames, and pipelined chunks."""
    probs = [synth.make_frame_problem(synth.config_seed(2, 300 + i), n_lines=60) for i in range(24)]
    probs.append(synth.make_frame_problem(synth.config_seed(2, 330), n_points=40, n_lines=7))
    probs.append(synth.make_frame_problem(synth.config_seed(2, 331), n_points=12, n_lines=25, stereo_frac=0.4))
    only_lines = synth.make_frame_problem(synth.config_seed(2, 332), n_points=4, n_lines=30)
    for k in ("mp_id_point", "mp_id_cam", "mp_inlier", "sp_id_point", "sp_id_cam", "sp_inlier"):
        setattr(only_lines, k, getattr(only_lines, k)[:0])
    only_lines.mp_kp, only_lines.sp_kp = only_lines.mp_kp[:0], only_lines.sp_kp[:0]
    probs.append(only_lines)
    probs.append(synth.make_frame_problem(synth.config_seed(2, 333)))  # a frame without lines in a batch with lines
    batch = FrameBatch.from_problems(probs)
    assert batch.has_lines
    res = gpu_ctx.frame_batch(batch)
    _check_against_oracle(orc, probs, batch, res)
    # upload / solve / download path gives the same bits as the one-shot call
    gpu_ctx.frame_batch_upload(batch)
    gpu_ctx.frame_batch_solve()
    res2 = gpu_ctx.frame_batch_download(gpu_ctx.alloc_frame_result(batch))
    assert np.array_equal(res.pose_twc, res2.pose_twc) and np.array_equal(res.sline_inlier, res2.sline_inlier)
    # a frame's result does not depend on the batch around it
    solo = gpu_ctx.frame_batch(batch.slice(3, 4))
    assert np.array_equal(solo.pose_twc[:, 0], res.pose_twc[:, 3])


def test_frame_batch_lines_large_pipelined(gpu_ctx, orc):
    """2048 frames with lines go through the chunked H2D / compute / D2H pipeline of the one-shot call."""
    batch = synth.make_frame_batch(2, 2048, first_instance=5000, n_points=60, n_lines=12)
    res = gpu_ctx.frame_batch(batch)
    for f in (0, 511, 512, 1300, 2047):
        p = batch.frame_problem(f)
        st = orc.frame_opt(p)
        assert int(res.num_inliers[f]) == st["ret"]
        assert np.linalg.norm(p.pose_p - res.pose_twc[:3, f]) < POS_TOL
        b0, b1 = batch.sline_begin[f], batch.sline_begin[f + 1]
        assert np.array_equal(p.sl_inlier, res.sline_inlier[b0:b1])
