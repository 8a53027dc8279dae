#!/usr/bin/env python
"""Runs the reference's own LocalmapOptimization / FrameOptimization (oracle/_ref/ref_driver, built by the CMake recipe
in this directory against a real g2o) on the seeded problems of tests/golden/make_golden.py and writes
tests/golden/g2o_<case>.npz (inputs + the REFERENCE's outputs). tests/test_g2o_fixtures.py then pins the CPU oracle
against them at the parity tolerances (index sets bit-exact, poses 1e-5 m / 1e-5 rad).

Needs a machine with g2o, Eigen, OpenCV and yaml-cpp; the build container of this repository has none of them, so no
g2o_*.npz is committed yet and the oracle's header still says "parity unpinned".

    RSPL_REF_CAMERA_YAML=<RSPL-SLAM>/configs/euroc.yaml python oracle/g2o_validation/make_fixtures.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "shim"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from rspl_slam_b200 import synth  # noqa: E402
from shim_dump import dump_problem  # noqa: E402
from make_golden import FRAME_CASES, FRAME_FIELDS, LOCAL_CASES, LOCAL_FIELDS  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(kind, p, tmp):
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    dump_problem(kind, p, fin)
    subprocess.run([DRIVER, fin, fout], check=True)
    return np.fromfile(fout, dtype=np.float64)


def main():
    if not os.path.exists(DRIVER):
        raise SystemExit("build oracle/_ref/ref_driver first (see CMakeLists.txt in this directory)")
    if "RSPL_REF_CAMERA_YAML" not in os.environ:
        raise SystemExit("set RSPL_REF_CAMERA_YAML to <RSPL-SLAM>/configs/euroc.yaml")
    tmp = tempfile.mkdtemp(prefix="rspl_g2o_")
    for name, kw in LOCAL_CASES.items():
        p = synth.make_local_problem(**kw)
        out = run(0, p, tmp)
        q = p.copy()
        k = 1
        npose, npt, nln = len(p.pose_id), len(p.point_id), len(p.line_id)
        poses = out[k:k + 7 * npose].reshape(npose, 7); k += 7 * npose
        q.pose_p, q.pose_q = poses[:, :3].copy(), poses[:, 3:].copy()
        q.point_p = out[k:k + 3 * npt].reshape(npt, 3).copy(); k += 3 * npt
        q.line_L = out[k:k + 6 * nln].reshape(nln, 6).copy(); k += 6 * nln
        for f in ("mp_inlier", "sp_inlier", "ml_inlier", "sl_inlier"):
            n = len(getattr(p, f))
            setattr(q, f, out[k:k + n].astype(getattr(p, f).dtype)); k += n
        np.savez_compressed(os.path.join(GOLDEN, "g2o_" + name + ".npz"),
                            **{"in_" + f: getattr(p, f) for f in LOCAL_FIELDS}, **{"out_" + f: getattr(q, f) for f in LOCAL_FIELDS})
    for name, kw in FRAME_CASES.items():
        if kw.get("n_lines"):
            continue  # the line extension has no counterpart in the reference
        p = synth.make_frame_problem(**kw)
        out = run(1, p, tmp)
        q = p.copy()
        q.pose_p, q.pose_q = out[1:4].copy(), out[4:8].copy()
        k = 8 + 3 * len(p.point_id)
        for f in ("mp_inlier", "sp_inlier"):
            n = len(getattr(p, f))
            setattr(q, f, out[k:k + n].astype(getattr(p, f).dtype)); k += n
        np.savez_compressed(os.path.join(GOLDEN, "g2o_" + name + ".npz"),
                            **{"in_" + f: getattr(p, f) for f in FRAME_FIELDS}, **{"out_" + f: getattr(q, f) for f in FRAME_FIELDS},
                            ret=int(out[0]))
    print("wrote", sorted(f for f in os.listdir(GOLDEN) if f.startswith("g2o_")))


if __name__ == "__main__":
    main()
