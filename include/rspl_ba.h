/*
 * rspl_ba.h — C-ABI of the B200-native point+line bundle-adjustment solver.
 *
 * Drop-in boundary for the reference's g2o_optimization module (RSPL-SLAM):
 *   LocalmapOptimization(...)  include/g2o_optimization/g2o_optimization.h:15-18  ->  rspl_ba_local_batch()
 *   FrameOptimization(...)     include/g2o_optimization/g2o_optimization.h:20-22  ->  rspl_ba_frame_batch()
 * The reference has no FFI for this path (plain C++ free functions linked into air_vo_lib,
 * CMakeLists.txt:57-77); the binding a maintainer adds is the header-only C++ shim
 * include/rspl_ba/g2o_optimization_shim.hpp, which keeps the two signatures and flattens the
 * reference containers (include/g2o_optimization/types.h:19-174) into the structure-of-arrays
 * batches below. See INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes, fp64 values, int32 indices, uint8 flags; caller owns every buffer;
 *    the library never keeps a caller pointer after a call returns.
 *  - "planes": an array documented as [k][n] is k contiguous planes of n values (component-major
 *    structure-of-arrays), so device loads coalesce and H2D is a straight copy.
 *  - poses cross the boundary as the caller's Twc (Pose3d: p, q) with Eigen's quaternion storage
 *    order x,y,z,w (types.h:19-35); inversion to the optimiser's Tcw and back happens on the
 *    device exactly as g2o_optimization.cc:42,237-239,271,391-393 do.
 *  - vertex references inside a window/frame are *local indices* into that window's slice of the
 *    vertex arrays, in ascending-id (std::map) order; the shim compacts ids.
 *  - every function returns RSPL_BA_OK (0) or a negative RsplBaStatus; no exceptions, no aborts.
 *  - one context per calling thread; a context owns a non-blocking CUDA stream and grow-only
 *    device workspaces; no call performs a device-wide synchronisation.
 *  - there is NO CPU fallback: without a CUDA device every entry point fails with
 *    RSPL_BA_ERR_CUDA.
 */
#ifndef RSPL_BA_H_
#define RSPL_BA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSPL_BA_VERSION 100 /* 0.1.0 */

typedef enum RsplBaStatus {
  RSPL_BA_OK = 0,
  RSPL_BA_ERR_INVALID = -1,     /* malformed input (null pointer, index out of range, bad offsets) */
  RSPL_BA_ERR_CUDA = -2,        /* CUDA runtime error / no device */
  RSPL_BA_ERR_UNSUPPORTED = -3, /* problem exceeds a documented limit of this path (g2o itself accepts these inputs):
                                 *   - two constraints joining the same (pose, landmark) pair;
                                 *   - a landmark with more than 254 observations in a window of <= 64 free poses
                                 *     (larger windows: no limit);
                                 *   - more than 255 cameras, 65535 poses in one window, 65535 windows in one call.
                                 * Reported by the solve (host-driven paths) or by the download (graph path). */
  RSPL_BA_ERR_STATE = -4        /* staged call out of order (solve before upload, ...) */
} RsplBaStatus;

typedef struct RsplBaContext RsplBaContext;

/* OptimizationConfig (include/read_configs.h:50-56; `rate` is never read by the optimiser) plus the
 * schedules the reference hard-codes. */
typedef struct RsplBaOptions {
  double thr_mono_point;   /* cfg.mono_point   (chi2 threshold; Huber delta = (float)sqrt) */
  double thr_stereo_point; /* cfg.stereo_point */
  double thr_mono_line;    /* cfg.mono_line    */
  double thr_stereo_line;  /* cfg.stereo_line  */
  int32_t local_iters_pass1; /* 10, g2o_optimization.cc:173 */
  int32_t local_iters_pass2; /*  5, g2o_optimization.cc:210 */
  int32_t frame_rounds;      /*  4, g2o_optimization.cc:339 */
  int32_t frame_iters;       /* 10, g2o_optimization.cc:336 */
  int32_t stereo_bf_float;   /* 1: g2o's EdgeStereoSE3ProjectXYZ::cam_project(xyz, const float& bf) */
  int32_t frame_latency_mode; /* FrameOptimization batches: 0 = one warp per frame (throughput, default); 1 = one CTA
                               * per frame (latency: single calls, what the reference makes). Both are deterministic;
                               * they agree to rounding, not bitwise (different summation trees). */
} RsplBaOptions;

/* Per-problem statistics (optional outputs). */
typedef struct RsplBaStats {
  int32_t iters[4];          /* outer LM iterations per pass (local: 2 used) / per round (frame) */
  int32_t trials[4];         /* inner LM trials per pass / round */
  int64_t edges_linearized;  /* sum over buildSystem passes of active edges */
  int64_t edges_evaluated;   /* sum over error-evaluation passes of active edges */
  double final_chi2;         /* robust chi2 of the last accepted state of the last pass */
  double final_lambda;
} RsplBaStats;

/* ---------------------------------------------------------------------------------------------
 * FrameOptimization batch (pose-only): n_frames independent frames.
 * Per frame f: mono edges  [mono_begin[f],   mono_begin[f+1])   (VectorOfMonoPointConstraints order)
 *              stereo edges[stereo_begin[f], stereo_begin[f+1]) (VectorOfStereoPointConstraints order)
 * Each edge carries its world point Xw (the reference copies points[id].p into the edge,
 * g2o_optimization.cc:305,328).
 * ------------------------------------------------------------------------------------------- */
typedef struct RsplFrameBatch {
  int32_t n_frames;
  int32_t n_cameras;
  const double* cameras;        /* [n_cameras][5] fx, fy, cx, cy, bf (camera.h:25-29) */
  const double* pose_twc;       /* [7][n_frames] planes px,py,pz,qx,qy,qz,qw */
  const int32_t* mono_begin;    /* [n_frames+1] */
  const int32_t* stereo_begin;  /* [n_frames+1] */
  const double* mono_meas;      /* [2][n_mono]   x_left, y_left */
  const double* mono_xw;        /* [3][n_mono] */
  const int32_t* mono_cam;      /* [n_mono] or NULL (all 0) */
  const uint8_t* mono_inlier;   /* [n_mono] initial ->inlier flags, or NULL (all 1; map_builder.cc:569) */
  const double* stereo_meas;    /* [3][n_stereo] x_left, y_left, x_right */
  const double* stereo_xw;      /* [3][n_stereo] */
  const int32_t* stereo_cam;    /* [n_stereo] or NULL */
  const uint8_t* stereo_inlier; /* [n_stereo] or NULL */
  /* Optional extension, NOT in the reference (its FrameOptimization takes no line containers and leaves
   * deltaMonoLine / deltaStereoLine unused, g2o_optimization.cc:284-285): constraints of a frame on FIXED
   * 3-D lines, i.e. EdgeSE3ProjectLine / EdgeStereoSE3ProjectLine (edge_project_line.cc:21-42,
   * edge_project_stereo_line.cc:22-51) with the VertexLine3D held fixed. Information 0.1 I and Huber
   * delta = (float)sqrt(threshold) as in LocalmapOptimization (:128-169); classified per round like the
   * point edges ((float)chi2 > thr_mono_line / thr_stereo_line). All NULL / absent: the reference's case. */
  const int32_t* mono_line_begin;    /* [n_frames+1] or NULL */
  const int32_t* stereo_line_begin;  /* [n_frames+1] or NULL (both or neither) */
  const double* mono_line_lw;        /* [6][n_ml] world line g2o::Line3D [w, d], copied into the edge */
  const double* mono_line_meas;      /* [4][n_ml] x1,y1,x2,y2 (left) */
  const int32_t* mono_line_cam;      /* [n_ml] or NULL */
  const uint8_t* mono_line_inlier;   /* [n_ml] or NULL (all 1) */
  const double* stereo_line_lw;      /* [6][n_sl] */
  const double* stereo_line_meas;    /* [8][n_sl] left x1,y1,x2,y2, right x1,y1,x2,y2 */
  const int32_t* stereo_line_cam;    /* [n_sl] or NULL */
  const uint8_t* stereo_line_inlier; /* [n_sl] or NULL */
} RsplFrameBatch;

typedef struct RsplFrameBatchResult {
  double* pose_twc;        /* [7][n_frames] optimised Twc */
  uint8_t* mono_inlier;    /* [n_mono] */
  uint8_t* stereo_inlier;  /* [n_stereo] */
  int32_t* num_inliers;    /* [n_frames] FrameOptimization's return value, or NULL */
  RsplBaStats* stats;      /* [n_frames] or NULL */
  uint8_t* mono_line_inlier;   /* [n_ml] (line extension; may be NULL when the batch has no lines) */
  uint8_t* stereo_line_inlier; /* [n_sl] */
} RsplFrameBatchResult;

/* ---------------------------------------------------------------------------------------------
 * LocalmapOptimization batch: n_windows independent local windows.
 * Window w owns poses [pose_begin[w], pose_begin[w+1]) etc.; edge vertex indices are local to
 * the window. Edges may arrive in any order (the reference happens to emit them landmark-major,
 * map.cc:609-707; the library builds its own landmark-major and pose-major CSR on the device).
 * Line3d is g2o::Line3D storage [w(3), d(3)] (types.h:108-121).
 * ------------------------------------------------------------------------------------------- */
typedef struct RsplLocalBatch {
  int32_t n_windows;
  int32_t n_cameras;
  const double* cameras;          /* [n_cameras][5] */
  const int32_t* pose_begin;      /* [n_windows+1] */
  const int32_t* point_begin;     /* [n_windows+1] */
  const int32_t* line_begin;      /* [n_windows+1] */
  const int32_t* mono_pt_begin;   /* [n_windows+1] */
  const int32_t* stereo_pt_begin; /* [n_windows+1] */
  const int32_t* mono_ln_begin;   /* [n_windows+1] */
  const int32_t* stereo_ln_begin; /* [n_windows+1] */
  const double* pose_twc;         /* [7][n_poses] */
  const uint8_t* pose_fixed;      /* [n_poses] Pose3d::fixed */
  const double* point_xyz;        /* [3][n_points] */
  const double* line_wd;          /* [6][n_lines] */
  /* mono point edges */
  const int32_t* mp_pose;  const int32_t* mp_point; const int32_t* mp_cam; /* mp_cam may be NULL */
  const double* mp_meas;          /* [2][n_mp] */
  /* stereo point edges */
  const int32_t* sp_pose;  const int32_t* sp_point; const int32_t* sp_cam;
  const double* sp_meas;          /* [3][n_sp] */
  /* mono line edges */
  const int32_t* ml_pose;  const int32_t* ml_line;  const int32_t* ml_cam;
  const double* ml_meas;          /* [4][n_ml] x1,y1,x2,y2 (left) */
  /* stereo line edges */
  const int32_t* sl_pose;  const int32_t* sl_line;  const int32_t* sl_cam;
  const double* sl_meas;          /* [8][n_sl] left x1,y1,x2,y2, right x1,y1,x2,y2 (map.cc:686) */
} RsplLocalBatch;

typedef struct RsplLocalBatchResult {
  double* pose_twc;    /* [7][n_poses] */
  double* point_xyz;   /* [3][n_points] */
  double* line_wd;     /* [6][n_lines] */
  uint8_t* mp_inlier;  /* [n_mp] */
  uint8_t* sp_inlier;  /* [n_sp] */
  uint8_t* ml_inlier;  /* [n_ml] */
  uint8_t* sl_inlier;  /* [n_sl] */
  RsplBaStats* stats;  /* [n_windows] or NULL */
} RsplLocalBatchResult;

/* --- lifecycle ------------------------------------------------------------------------------ */
int rspl_ba_version(void);
void rspl_ba_default_options(RsplBaOptions* opt); /* thresholds 50/75/50/75 (configs_euroc.yaml:57-60), 10/5, 4x10 */
/* device < 0: current device. stream: a cudaStream_t to run on, or NULL to create an own
 * non-blocking stream. */
int rspl_ba_create(int device, void* stream, RsplBaContext** out);
void rspl_ba_destroy(RsplBaContext* ctx);
const char* rspl_ba_last_error(const RsplBaContext* ctx);
void* rspl_ba_stream(const RsplBaContext* ctx); /* the cudaStream_t all work of this context is ordered on */
int rspl_ba_device(const RsplBaContext* ctx);

/* --- FrameOptimization ---------------------------------------------------------------------- */
/* One call = upload + solve + download with host buffers (what the shim uses). */
int rspl_ba_frame_batch(RsplBaContext* ctx, const RsplFrameBatch* in, const RsplBaOptions* opt,
                        RsplFrameBatchResult* out);
/* Staged variant: inputs stay resident in HBM; solve may be repeated (it restarts from the
 * uploaded inputs every time); all three are ordered on the context stream, upload/download
 * block the host until their copies finished, solve is asynchronous. */
int rspl_ba_frame_batch_upload(RsplBaContext* ctx, const RsplFrameBatch* in);
int rspl_ba_frame_batch_solve(RsplBaContext* ctx, const RsplBaOptions* opt);
int rspl_ba_frame_batch_download(RsplBaContext* ctx, RsplFrameBatchResult* out);

/* --- LocalmapOptimization ------------------------------------------------------------------- */
/* One call = upload + solve + download. Large batches (>= 128 windows, >= 64 MB) are processed in window chunks on
 * child contexts so that copies overlap the solves (same results); such a call leaves the batch of an earlier
 * rspl_ba_local_batch_upload resident, a small one replaces it. */
int rspl_ba_local_batch(RsplBaContext* ctx, const RsplLocalBatch* in, const RsplBaOptions* opt,
                        RsplLocalBatchResult* out);
int rspl_ba_local_batch_upload(RsplBaContext* ctx, const RsplLocalBatch* in);
int rspl_ba_local_batch_solve(RsplBaContext* ctx, const RsplBaOptions* opt);
int rspl_ba_local_batch_download(RsplBaContext* ctx, RsplLocalBatchResult* out);

/* --- pinned host memory (optional; any host pointer is accepted, pinned ones copy faster) ---- */
void* rspl_ba_alloc_pinned(size_t bytes);
void rspl_ba_free_pinned(void* p);

/* --- introspection for benchmarks ----------------------------------------------------------- */
/* Number of kernel launches issued by this context since creation. */
int64_t rspl_ba_launch_count(const RsplBaContext* ctx);
/* Block the host until all work queued on the context stream has finished. */
int rspl_ba_sync(RsplBaContext* ctx);

/* Per-kernel-class timing with CUDA events on the context stream (what bench.py's roofline uses).
 * Classes: 0 frame_opt, 1 local_setup, 2 (unused), 3 init + pair lists, 4 linearize,
 * 5 pose blocks, 6 Schur prep, 7 Schur reduce, 8 reduced solve, 9 back-substitution / update /
 * evaluation, 10 LM control kernels, 11 flagging + write-back, 12 collectives of the global-BA path,
 * 13 assembly of the dense reduced system + pose update (14-15 reserved, zero). get_profile synchronises the stream, returns milliseconds and launch counts
 * (arrays of RSPL_BA_PROFILE_CLASSES entries) accumulated since the last call and resets them. */
#define RSPL_BA_PROFILE_CLASSES 16
int rspl_ba_set_profiling(RsplBaContext* ctx, int enabled);
int rspl_ba_get_profile(RsplBaContext* ctx, double* ms16, int64_t* launches16);

/* --- global BA: ONE problem distributed over the ranks of a communicator ----------------------
 * Replaces nothing in the reference (its only BA is the local one, map.cc:709); it is the scale-out
 * configuration C5 of SURVEY.md 8(e): landmarks, with all their constraints, are partitioned over the
 * ranks (one process per GPU), poses are replicated. Every rank uploads ONE window that holds all poses
 * (identical arrays on every rank) and its own share of the points / lines and their constraints, then
 * calls rspl_ba_global_solve collectively. Per LM trial the ranks all-reduce (sum, fp64) the pose blocks,
 * the rank-local pieces of the Schur complement and five scalars; the reduced camera system is
 * factorised redundantly on every rank, so all ranks take identical LM decisions and end with
 * identical poses. Results: poses on every rank, landmarks and inlier flags of the rank's own share.
 * The communicator is NCCL, loaded at run time; the 128-byte id comes from rank 0
 * (rspl_ba_comm_unique_id) and reaches the other ranks by whatever channel the host has.
 * n_ranks == 1 needs no id and no NCCL: the collectives become device copies. */
#define RSPL_BA_COMM_ID_BYTES 128
int rspl_ba_comm_unique_id(void* id128);
int rspl_ba_comm_init(RsplBaContext* ctx, int n_ranks, int rank, const void* id128);
int rspl_ba_comm_destroy(RsplBaContext* ctx);
int rspl_ba_comm_size(const RsplBaContext* ctx);
int rspl_ba_comm_rank(const RsplBaContext* ctx);
int64_t rspl_ba_collective_count(const RsplBaContext* ctx);
int rspl_ba_global_upload(RsplBaContext* ctx, const RsplLocalBatch* shard);
int rspl_ba_global_solve(RsplBaContext* ctx, const RsplBaOptions* opt);
int rspl_ba_global_download(RsplBaContext* ctx, RsplLocalBatchResult* out);

/* --- triangulation of new map points -------------------------------------
 * Replaces the per-point body of Map::TriangulateMappoint (/root/reference/src/map.cc:292-339) for a whole batch of
 * points: point i has the observations obs_begin[i] .. obs_begin[i+1] - 1 (its observers with a valid keypoint,
 * map.cc:299-314): keyframe index obs_frame[o] into frame_twc and left-image pixel (obs_uv[o], obs_uv[n_obs + o]).
 * frame_twc [7][n_frames] = Frame::GetPose() as p, q (x, y, z, w); cam5 = fx, fy, cx, cy, bf. out_ok[i] is the
 * reference's bool: 0 with fewer than two observations or when the 3 x 3 normal matrix has rank < 3 under
 * ColPivHouseholderQR with threshold 1e-5; out_xyz [3][n_points] of such a point is left untouched (the reference
 * does not call SetPosition then). *n_done (may be null): number of points triangulated. All pointers are host. */
int rspl_ba_triangulate_points(RsplBaContext* ctx, int32_t n_points, const int32_t* obs_begin, const int32_t* obs_frame,
                               const double* obs_uv, int32_t n_frames, const double* frame_twc, const double* cam5,
                               double* out_xyz, uint8_t* out_ok, int32_t* n_done);

/* --- endpoint refresh of the optimised map lines ----------------------------
 * Replaces the body of Map::UppdateMapline (/root/reference/src/map.cc:121-177), which Map::LocalMapOptimization
 * calls for every line the local BA returned (map.cc:790-797), for a whole batch of lines. line_wd [6][n_lines] =
 * g2o::Line3D [w, d] (what rspl_ba_local_batch_download returns); line l has the map points
 * pt_index[pt_begin[l] .. pt_begin[l+1] - 1] (the valid map points on the line in its observing keyframes,
 * map.cc:128-139) into point_xyz [3][n_points]. out_ok[l] is the reference's bool: 1 when a point within 0.2 of the
 * line gave both ends, with the reference's quirk that the running maximum starts at DBL_MIN (no end from points
 * whose main coordinate is not positive). endpoints [6][n_lines] (first the end at the largest main coordinate) is
 * left untouched where out_ok[l] = 0, as the reference does not call SetEndpoints then. *n_done (may be null):
 * number of lines refreshed. All pointers are host. */
int rspl_ba_update_maplines(RsplBaContext* ctx, int32_t n_lines, const double* line_wd, const int32_t* pt_begin,
                            const int32_t* pt_index, int32_t n_points, const double* point_xyz, double* endpoints,
                            uint8_t* out_ok, int32_t* n_done);

/* The same refresh on the RESIDENT result of the local BA just solved on this context (after rspl_ba_local_batch_solve
 * or an unchunked rspl_ba_local_batch): the optimised Line3Ds and map points are read where the solve left them in
 * HBM, only the point lists travel. Line l = the l-th line of the batch (RsplLocalBatch::line_begin order),
 * pt_index = batch-wide point indices (RsplLocalBatch::point_begin[w] + index in window w). endpoints [6][n_lines],
 * out_ok [n_lines]; unlike above, endpoints of a line with out_ok[l] = 0 are written as 0. */
int rspl_ba_local_batch_update_maplines(RsplBaContext* ctx, const int32_t* pt_begin, const int32_t* pt_index,
                                        double* endpoints, uint8_t* out_ok, int32_t* n_done);

/* --- unit-level device entry points (used by the parity tests) ------------------------------- */
/* Evaluates n edges of one type on the device. edge_type: 0 mono point, 1 stereo point, 2 mono
 * line, 3 stereo line, 4 / 5 mono / stereo pose-only point edge (lm = the fixed world point Xw, Jl = 0). pose7 [n][7] = optimiser pose Tcw as qx,qy,qz,qw,tx,ty,tz; lm [n][6]
 * (3 used for points); meas [n][8]; cam5 [5]. Outputs (host): err [n][4], Jl [n][16], Jp [n][24]
 * (row-major dim x ld / dim x 6), chi2 [n]. */
int rspl_ba_eval_edges(RsplBaContext* ctx, int edge_type, int32_t n, const double* pose7,
                       const double* lm, const double* meas, const double* cam5,
                       int32_t stereo_bf_float, double* err, double* Jl, double* Jp, double* chi2);
/* Applies the manifold updates on the device: kind 0 pose (state 7, update 6), 1 point (3,3),
 * 2 line (6,4). state [n][7], upd [n][6], out [n][7]. */
int rspl_ba_oplus(RsplBaContext* ctx, int kind, int32_t n, const double* state, const double* upd,
                  double* out);

/* The kernels' reciprocal (op 0), reciprocal square root (op 1) and square root (op 2) without the CUDA library's
 * range test (ba_math.cuh: rcp_nr / rsqrt_nr / sqrt_nr), evaluated on n host values (tests/test_edges_gpu.py: the
 * reciprocal correctly rounded, rsqrt / sqrt within 2 ulp for normal arguments). */
int rspl_ba_unit_math(RsplBaContext* ctx, int op, int32_t n, const double* in, double* out);

#ifdef __cplusplus
}
#endif
#endif /* RSPL_BA_H_ */
