"""B200-native point+line bundle adjustment: drop-in for RSPL-SLAM's g2o_optimization module.

The product is the CUDA library behind include/rspl_ba.h (rspl_slam_b200/csrc); this package is
the thin host layer around it: the ctypes binding, the flat batch containers that mirror the C++
shim, the synthetic-input generator and the build helper.
"""
from .problem import (EUROC_CAMERA, FrameBatch, FrameBatchResult, FrameProblem, LocalBatch, LocalBatchResult,
                      LocalProblem, OptimizationConfig, shard_range)

__all__ = ["EUROC_CAMERA", "FrameBatch", "FrameBatchResult", "FrameProblem", "LocalBatch", "LocalBatchResult",
           "LocalProblem", "OptimizationConfig", "shard_range"]
