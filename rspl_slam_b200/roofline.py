"""Algorithmic-byte accounting of SURVEY.md §8(d) (the contract figures the roofline is quoted on).

fp64 values, int32 indices, structure-of-arrays, each array touched once per pass
("materialised-W" formulation):
  linearise pass, per edge   : edge record read + chi2 write + pose x landmark block write
      stereo point 188 B, mono point 180 B, stereo line 276 B, mono line 244 B
  linearise pass, per vertex : free point 120 B, free line 208 B, free pose 392 B
  Schur pass                 : per edge 144 / 192 B (W read); per landmark 96/160 B read + 96/160 B write;
                               288 B per non-zero block of the reduced system
  back-substitution / update : per edge 144 / 192 B (W read); per landmark 48 / 80 B
  evaluation pass            : edge record read + chi2 write (44 / 36 / 84 / 52 B), landmark state read
A pass count comes from the device-side statistics (edges_linearized = active edges x linearise
passes, trials = Schur + back-substitution passes).
"""
from __future__ import annotations

import numpy as np

LIN_EDGE = {"sp": 188, "mp": 180, "sl": 276, "ml": 244}
LIN_VERTEX = {"point": 120, "line": 208, "pose": 392}
SCHUR_EDGE = {"point": 144, "line": 192}
SCHUR_LM = {"point": 192, "line": 320}
SCHUR_BLOCK = 288
BACK_EDGE = {"point": 144, "line": 192}
BACK_LM = {"point": 48, "line": 80}


def local_pass_bytes(batch) -> dict:
    """Algorithmic bytes of ONE linearise / Schur / back-substitution pass over the whole batch."""
    n = {"mp": len(batch.mp_pose), "sp": len(batch.sp_pose), "ml": len(batch.ml_pose), "sl": len(batch.sl_pose)}
    n_pt, n_ln = batch.point_xyz.shape[1], batch.line_wd.shape[1]
    n_free = int((batch.pose_fixed == 0).sum())
    free_per_win = np.add.reduceat((batch.pose_fixed == 0).astype(np.int64), batch.pose_begin[:-1])
    blocks = int((free_per_win * (free_per_win + 1) // 2).sum())
    lin = sum(LIN_EDGE[k] * v for k, v in n.items()) + LIN_VERTEX["point"] * n_pt + LIN_VERTEX["line"] * n_ln + LIN_VERTEX["pose"] * n_free
    pe, le = n["mp"] + n["sp"], n["ml"] + n["sl"]
    schur = SCHUR_EDGE["point"] * pe + SCHUR_EDGE["line"] * le + SCHUR_LM["point"] * n_pt + SCHUR_LM["line"] * n_ln + SCHUR_BLOCK * blocks
    back = BACK_EDGE["point"] * pe + BACK_EDGE["line"] * le + BACK_LM["point"] * n_pt + BACK_LM["line"] * n_ln
    return {"linearize": float(lin), "schur": float(schur), "backsub": float(back), "edges": pe + le}


def local_algorithmic_bytes(batch, stats) -> float:
    """Algorithmic bytes of one whole solve launch: pass counts from the device statistics.
    Outlier edges dropped in pass 2 are accounted through edges_linearized (active edges only)."""
    pb = local_pass_bytes(batch)
    edges = max(pb["edges"], 1)
    lin_passes = float(stats["edges_linearized"].sum()) / edges  # batch-average number of linearise passes
    win_edges = batch.window_edges().astype(np.float64)
    trials = stats["trials"].sum(axis=1).astype(np.float64)
    trial_passes = float((trials * win_edges).sum()) / edges     # edge-weighted number of Schur/back-sub passes
    return pb["linearize"] * lin_passes + (pb["schur"] + pb["backsub"]) * trial_passes


def local_class_bytes(batch, stats) -> dict:
    """Algorithmic bytes per solve of the three kernel classes (linearise / Schur / back-substitution)."""
    pb = local_pass_bytes(batch)
    edges = max(pb["edges"], 1)
    lin_passes = float(stats["edges_linearized"].sum()) / edges
    win_edges = batch.window_edges().astype(np.float64)
    trials = stats["trials"].sum(axis=1).astype(np.float64)
    trial_passes = float((trials * win_edges).sum()) / edges
    return {"linearize": pb["linearize"] * lin_passes, "schur": pb["schur"] * trial_passes,
            "backsub": pb["backsub"] * trial_passes}
