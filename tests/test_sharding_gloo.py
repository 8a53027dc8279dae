"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py / multi-GPU use. Independent units
(frames, windows) are block-partitioned across ranks with NO data-path collective;
the only collectives are the timing MAX and the unit-count SUM of the benchmark. Each rank's
shard, generated from its own instance range, is exactly the corresponding slice of the global
batch, so results cannot depend on the GPU count."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rspl_slam_b200 import synth
from rspl_slam_b200.problem import shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, n_windows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_frames, rank, world)
    fb = synth.make_frame_batch(2, b - a, first_instance=a, n_points=40)      # the rank's own frames
    wa, wb = shard_range(n_windows, rank, world)
    lb, _ = synth.make_local_batch(4, wb - wa, first_instance=wa, n_kf=4, n_points=40, n_lines=6)
    # what bench.py reduces: max of times, sum of processed units / edges
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    n = torch.tensor([float(fb.n_edges), float(lb.n_edges), float(b - a), float(wb - wa)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, a, b, wa, wb, fb.pose_twc.tobytes(), fb.stereo_meas.tobytes(), lb.sp_meas.tobytes(), float(t[0]), n.tolist()))
    dist.destroy_process_group()


def test_two_ranks_partition_the_units_and_reduce_only_scalars():
    world, n_frames, n_windows = 2, 9, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, n_windows, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole_f = synth.make_frame_batch(2, n_frames, n_points=40)
    whole_l, _ = synth.make_local_batch(4, n_windows, n_kf=4, n_points=40, n_lines=6)
    for rank, a, b, wa, wb, pose, smeas, lmeas, tmax, sums in out:
        part = whole_f.slice(a, b)
        assert pose == part.pose_twc.tobytes() and smeas == part.stereo_meas.tobytes()
        assert lmeas == whole_l.slice(wa, wb).sp_meas.tobytes()
        assert tmax == float(world)  # MAX over ranks
        assert sums == [float(whole_f.n_edges), float(whole_l.n_edges), float(n_frames), float(n_windows)]
    assert out[0][2] == out[1][1] and out[0][1] == 0 and out[1][2] == n_frames


# ------------------------------------------------------------------------------------------------
# Global BA: ONE problem, landmarks partitioned over the ranks. Host logic only
# (no GPU here): the shards cover every landmark and constraint exactly once, the 128-byte
# communicator id reaches every rank, per-rank results merge back into the full problem.
# ------------------------------------------------------------------------------------------------
def _global_worker(rank, world, port, q):
    from oracle import orc
    from rspl_slam_b200.problem import shard_landmarks

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = synth.make_local_problem(synth.config_seed(5, 7), n_kf=9, n_points=300, n_lines=40)  # same seed on every rank
    ident = [bytes(range(128)) if rank == 0 else None]  # stands for rspl_ba_comm_unique_id() on rank 0
    dist.broadcast_object_list(ident, src=0)
    shard = shard_landmarks(full, rank, world)
    # classification only (0 LM iterations): an edge's flag depends on its own pose and landmark, so the
    # shards' flags must merge to the flags of the undivided problem
    st = orc.local_ba(shard.problem, orc.make_config(iters=(0, 0)))
    n = torch.tensor([float(shard.problem.n_edges), float(len(shard.point_idx)), float(len(shard.line_idx))], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, shard)
    dist.barrier()
    if rank == 0:
        q.put((ident[0], n.tolist(), gathered, st["final_chi2"]))
    else:
        q.put((ident[0], None, None, st["final_chi2"]))
    dist.destroy_process_group()


def test_two_ranks_partition_the_landmarks_of_one_problem():
    from oracle import orc
    from rspl_slam_b200.problem import merge_landmark_shards

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_global_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[0] == bytes(range(128)) for o in out)
    sums, shards = next((o[1], o[2]) for o in out if o[1] is not None)
    full = synth.make_local_problem(synth.config_seed(5, 7), n_kf=9, n_points=300, n_lines=40)
    assert sums == [float(full.n_edges), float(len(full.point_id)), float(len(full.line_id))]
    # disjoint cover
    pts = np.concatenate([s.point_idx for s in shards])
    assert np.array_equal(np.sort(pts), np.arange(len(full.point_id)))
    for pre in ("mp", "sp", "ml", "sl"):
        e = np.concatenate([s.edge_idx[pre] for s in shards])
        assert np.array_equal(np.sort(e), np.arange(len(getattr(full, f"{pre}_inlier"))))
    for s in shards:  # every rank keeps every pose
        # (the oracle's write-back went Twc -> Tcw -> Twc: equal up to rounding)
        assert np.array_equal(s.problem.pose_id, full.pose_id) and np.allclose(s.problem.pose_p, full.pose_p, rtol=0, atol=1e-12)
    ref = full.copy()
    st = orc.local_ba(ref, orc.make_config(iters=(0, 0)))
    merged = merge_landmark_shards(full, shards)
    for pre in ("mp", "sp", "ml", "sl"):
        assert np.array_equal(getattr(merged, f"{pre}_inlier"), getattr(ref, f"{pre}_inlier"))
    assert abs(sum(o[3] for o in out) - st["final_chi2"]) <= 1e-9 * st["final_chi2"]
