#!/usr/bin/env python
"""Generates tests/golden/*.npz: small seeded problems + the outputs of the CPU oracle on them.

PROVENANCE: the reference (RSPL-SLAM / g2o) cannot be built or imported in the build container
(g2o, Eigen and OpenCV are absent), and it ships no golden vectors of its own, so
these fixtures are produced by oracle/ (the CPU restatement of the reference's algorithm), not by
the reference itself. They pin the oracle against regressions and let the GPU tests check the
CUDA path without executing anything under oracle/. Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from rspl_slam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL_FIELDS = ["pose_id", "pose_p", "pose_q", "pose_fixed", "point_id", "point_p", "line_id", "line_L", "cams",
                "mp_id_pose", "mp_id_point", "mp_id_cam", "mp_kp", "mp_inlier", "sp_id_pose", "sp_id_point", "sp_id_cam",
                "sp_kp", "sp_inlier", "ml_id_pose", "ml_id_line", "ml_id_cam", "ml_l2d", "ml_inlier", "sl_id_pose",
                "sl_id_line", "sl_id_cam", "sl_l2d", "sl_inlier"]
FRAME_FIELDS = ["pose_p", "pose_q", "point_id", "point_p", "cams", "mp_id_point", "mp_id_cam", "mp_kp", "mp_inlier",
                "sp_id_point", "sp_id_cam", "sp_kp", "sp_inlier"]

LOCAL_CASES = {
    "local_a": dict(seed=synth.config_seed(1, 9000), n_kf=5, n_points=120, n_lines=16),
    "local_b": dict(seed=synth.config_seed(1, 9001), n_kf=7, n_points=200, n_lines=24, first_kf_id=40),
    "local_c": dict(seed=synth.config_seed(1, 9002), n_kf=4, n_points=80, n_lines=0, stereo_point_frac=0.5),
}
FRAME_CASES = {
    "frame_a": dict(seed=synth.config_seed(2, 9000), n_points=120),
    "frame_b": dict(seed=synth.config_seed(2, 9001), n_points=60, stereo_frac=0.6),
    "frame_c": dict(seed=synth.config_seed(2, 9002), n_points=8),
    # line extension of the pose-only path (constraints on fixed lines; absent in the reference)
    "frame_d_lines": dict(seed=synth.config_seed(2, 9003), n_points=100, n_lines=24),
    "frame_e_lines": dict(seed=synth.config_seed(2, 9004), n_points=30, n_lines=40, stereo_frac=0.5),
}
FRAME_LINE_FIELDS = ["line_id", "line_L", "ml_id_line", "ml_id_cam", "ml_l2d", "ml_inlier", "sl_id_line", "sl_id_cam", "sl_l2d",
                     "sl_inlier"]


def main():
    for name, kw in LOCAL_CASES.items():
        p = synth.make_local_problem(**kw)
        q = p.copy()
        st = orc.local_ba(q, trace=True)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            **{"in_" + f: getattr(p, f) for f in LOCAL_FIELDS}, **{"out_" + f: getattr(q, f) for f in LOCAL_FIELDS},
            iters=np.asarray(st["iters"]), trials=np.asarray(st["trials"]), final_chi2=st["final_chi2"],
            edges_linearized=st["edges_linearized"],
            trace=np.asarray([[r["pass_"], r["iter"], r["trial"], r["accepted"], r["chi_before"], r["chi_after"],
                               r["lambda_"], r["rho"]] for r in st["trace"]]))
    for name, kw in FRAME_CASES.items():
        p = synth.make_frame_problem(**kw)
        q = p.copy()
        st = orc.frame_opt(q, trace=True)
        fields = FRAME_FIELDS + (FRAME_LINE_FIELDS if kw.get("n_lines") else [])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            **{"in_" + f: getattr(p, f) for f in fields}, **{"out_" + f: getattr(q, f) for f in fields},
            iters=np.asarray(st["iters"]), trials=np.asarray(st["trials"]), final_chi2=st["final_chi2"], ret=st["ret"],
            edges_linearized=st["edges_linearized"])
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
