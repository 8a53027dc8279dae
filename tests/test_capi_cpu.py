"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/rspl_ba.h
declares, and refuses to run without a CUDA device (no CPU fallback). No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rspl_slam_b200 import build, capi
    build.build_library()
    return capi.load_library()


def test_every_declared_symbol_is_exported(lib):
    from rspl_slam_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "rspl_ba.h")).read()
    declared = set(re.findall(r"\b(rspl_ba_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_version_and_default_options(lib):
    from rspl_slam_b200 import capi
    assert lib.rspl_ba_version() == 100
    o = capi.make_options()
    assert (o.thr_mono_point, o.thr_stereo_point, o.thr_mono_line, o.thr_stereo_line) == (50.0, 75.0, 50.0, 75.0)
    assert (o.local_iters_pass1, o.local_iters_pass2, o.frame_rounds, o.frame_iters) == (10, 5, 4, 10)
    assert o.stereo_bf_float == 1


def test_struct_layouts_match_header(lib):
    from rspl_slam_b200 import capi
    assert ctypes.sizeof(capi.RsplBaStats) == 64 and capi.STATS_DTYPE.itemsize == 64
    assert ctypes.sizeof(capi.RsplBaOptions) == 56
    # pointer-heavy structs: 2 ints + 12 (+ 10 of the line extension) pointers / 2 ints + 28 pointers
    assert ctypes.sizeof(capi.RsplFrameBatch) == 8 + 22 * 8
    assert ctypes.sizeof(capi.RsplFrameBatchResult) == 7 * 8
    assert ctypes.sizeof(capi.RsplLocalBatch) == 8 + 28 * 8


def test_no_cpu_fallback(lib):
    import torch
    from rspl_slam_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only machines")
    with pytest.raises(capi.RsplBaError) as e:
        capi.Context(device=0)
    assert e.value.code == capi.RSPL_BA_ERR_CUDA


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under rspl_slam_b200/ or include/ references it."""
    for base in ("rspl_slam_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".inl")):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "liboracle" not in src and "import orc" not in src and "from oracle" not in src, f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle on the host cores) needs no GPU: one JSON line with the
    keys the driver reads."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "point+line BA edges linearized/sec" and line["unit"] == "edges/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
