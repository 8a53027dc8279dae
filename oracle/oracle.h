/*
 * oracle/oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C interface of the CPU oracle: a dependency-free restatement of the algorithm the
 * reference's g2o_optimization module runs (reference: src/g2o_optimization/g2o_optimization.cc:21-397,
 * edge_project_line.cc:21-42, edge_project_stereo_line.cc:22-51, vertex_line3d.h:26-43) plus the
 * slice of g2o it exercises (un-vendored by the reference; semantics per SURVEY.md §9).
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors for this path and g2o + Eigen
 * are not installable here, so this oracle is checked only against (i) closed-form known-answer
 * geometry and (ii) an independent numpy restatement (oracle/np_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (rspl_slam_b200/, include/rspl_ba.h) never links or calls it.
 *
 * Data model mirrors the reference's boundary types (include/g2o_optimization/types.h): vertices
 * are addressed by *id* (std::map keys), constraints carry id_pose / id_point / id_camera and an
 * in/out `inlier` flag, everything is mutated in place.
 */
#ifndef RSPL_ORACLE_H_
#define RSPL_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcTraceRow {
  int32_t pass;      /* local BA: 0/1; pose-only: round index */
  int32_t iter;      /* outer LM iteration */
  int32_t trial;     /* inner trial */
  int32_t accepted;  /* 1 = step kept */
  double chi_before; /* currentChi */
  double chi_after;  /* tempChi */
  double lambda;     /* lambda used for this trial */
  double rho;
} OrcTraceRow;

typedef struct OrcStats {
  int32_t iters[4];        /* outer iterations executed per pass / round */
  int32_t trials[4];       /* inner trials per pass / round */
  int64_t edges_linearized;/* sum over buildSystem calls of active edges */
  int64_t edges_evaluated; /* sum over computeActiveErrors calls of active edges */
  double final_chi2;       /* activeRobustChi2 of the last accepted state of the last pass */
  int32_t n_trace;         /* rows written to trace (<= trace_cap) */
  int32_t trace_cap;
  OrcTraceRow* trace;      /* optional, caller-owned */
} OrcStats;

/* Mirrors OptimizationConfig (include/read_configs.h:50-56) + the schedules hard-coded in
 * g2o_optimization.cc:173,210,336,339. */
typedef struct OrcConfig {
  double mono_point, stereo_point, mono_line, stereo_line;
  int32_t iters_pass1, iters_pass2; /* 10, 5 */
  int32_t rounds, iters_round;      /* 4, 10 */
  int32_t stereo_bf_float;          /* 1: EdgeStereoSE3ProjectXYZ::cam_project takes bf as float (g2o) */
  int32_t reserved;
  double numeric_delta;             /* step of the numeric line Jacobians; <= 0 means g2o's 1e-9.
                                       Tests use 1e-6 to separate difference-quotient noise from real deviations. */
} OrcConfig;

typedef struct OrcLocalProblem {
  /* MapOfPoses (types.h:19-35): id -> {fixed, p, q(x,y,z,w)}; ids ascending */
  int32_t n_poses;
  const int32_t* pose_id;
  double* pose_p; /* [n][3] in/out, Twc */
  double* pose_q; /* [n][4] x,y,z,w in/out */
  const uint8_t* pose_fixed;
  /* MapOfPoints3d (types.h:38-51) */
  int32_t n_points;
  const int32_t* point_id;
  double* point_p; /* [n][3] in/out */
  /* MapOfLine3d (types.h:108-121): g2o::Line3D = [w(3), d(3)] */
  int32_t n_lines;
  const int32_t* line_id;
  double* line_L; /* [n][6] in/out */
  /* camera_list: fx, fy, cx, cy, bf */
  int32_t n_cams;
  const double* cams; /* [n][5] */
  /* constraints (types.h:54-174), insertion order */
  int32_t n_mono_pt;
  const int32_t *mp_id_pose, *mp_id_point, *mp_id_cam;
  const double* mp_kp; /* [n][2] */
  uint8_t* mp_inlier;
  int32_t n_stereo_pt;
  const int32_t *sp_id_pose, *sp_id_point, *sp_id_cam;
  const double* sp_kp; /* [n][3] */
  uint8_t* sp_inlier;
  int32_t n_mono_ln;
  const int32_t *ml_id_pose, *ml_id_line, *ml_id_cam;
  const double* ml_l2d; /* [n][4] */
  uint8_t* ml_inlier;
  int32_t n_stereo_ln;
  const int32_t *sl_id_pose, *sl_id_line, *sl_id_cam;
  const double* sl_l2d; /* [n][8] */
  uint8_t* sl_inlier;
} OrcLocalProblem;

typedef struct OrcFrameProblem {
  /* poses.size()==1 (g2o_optimization.cc:259) */
  double* pose_p; /* [3] in/out */
  double* pose_q; /* [4] x,y,z,w in/out */
  int32_t n_points;
  const int32_t* point_id;
  const double* point_p; /* [n][3] */
  int32_t n_cams;
  const double* cams;
  int32_t n_mono_pt;
  const int32_t *mp_id_point, *mp_id_cam;
  const double* mp_kp;
  uint8_t* mp_inlier;
  int32_t n_stereo_pt;
  const int32_t *sp_id_point, *sp_id_cam;
  const double* sp_kp;
  uint8_t* sp_inlier;
  /* Extension, NOT in the reference (its FrameOptimization takes no line containers and leaves
   * deltaMonoLine / deltaStereoLine unused, g2o_optimization.cc:284-285): constraints of the frame on
   * FIXED 3-D lines, i.e. EdgeSE3ProjectLine / EdgeStereoSE3ProjectLine whose VertexLine3D is fixed
   * (SURVEY 8a note). Set-up as in LocalmapOptimization (:128-169): information 0.1 I, Huber delta =
   * (float)sqrt(threshold); classified per round like the point edges ((float)chi2 > threshold). */
  int32_t n_lines;
  const int32_t* line_id;
  const double* line_L; /* [n][6] */
  int32_t n_mono_ln;
  const int32_t *ml_id_line, *ml_id_cam;
  const double* ml_l2d; /* [n][4] */
  uint8_t* ml_inlier;
  int32_t n_stereo_ln;
  const int32_t *sl_id_line, *sl_id_cam;
  const double* sl_l2d; /* [n][8] */
  uint8_t* sl_inlier;
} OrcFrameProblem;

/* LocalmapOptimization (g2o_optimization.cc:21-252). Returns 0, or <0 on malformed input. */
int orc_local_ba(OrcLocalProblem* prob, const OrcConfig* cfg, OrcStats* stats);
/* FrameOptimization (g2o_optimization.cc:256-397). Returns the reference's int (#inliers), <0 on error. */
int orc_frame_opt(OrcFrameProblem* prob, const OrcConfig* cfg, OrcStats* stats);

/* Batched drivers for the CPU baseline: one problem per OpenMP thread (n_threads<=0: all). */
int orc_frame_opt_batch(int32_t n, OrcFrameProblem* probs, const OrcConfig* cfg, OrcStats* stats,
                        int32_t* ret, int32_t n_threads);
int orc_local_ba_batch(int32_t n, OrcLocalProblem* probs, const OrcConfig* cfg, OrcStats* stats,
                       int32_t n_threads);
int orc_max_threads(void);

/* ---- unit-level entry points (per-edge KATs and GPU unit tests) ---- */
/* pose7 = g2o vertex estimate Tcw as [qx,qy,qz,qw,tx,ty,tz]. */
void orc_pose_from_twc(const double* p3, const double* q4, double* pose7); /* SE3Quat(q,p).inverse() */
void orc_pose_to_twc(const double* pose7, double* p3, double* q4);
void orc_se3_exp(const double* u6, double* pose7);
void orc_pose_oplus(const double* pose7, const double* u6, double* out7); /* exp(u)*T */
void orc_line_oplus(const double* L6, const double* v4, double* out6);
void orc_line_from_cartesian(const double* pv6, double* out6);
void orc_line_transform(const double* pose7, const double* L6, double* out6);
/* edge_type: 0 mono point, 1 stereo point, 2 mono line, 3 stereo line (binary edges);
 * 4 mono pose-only, 5 stereo pose-only (landmark = Xw constant).
 * cam5 = fx,fy,cx,cy,bf. err[4]; Jl[dim*ld] (ld = 3 points / 4 lines, row-major), Jp[dim*6].
 * Jacobians follow g2o: analytic for points, numeric central differences
 * delta=1e-9 for lines (§9.8). Returns the residual dimension. */
int orc_edge_eval(int edge_type, const double* pose7, const double* lm, const double* meas,
                  const double* cam5, int stereo_bf_float, double* err, double* Jl, double* Jp);
/* Huber (§9.5): delta = (float)sqrt(thr); out rho[3] */
void orc_huber(double chi2, double thr, double* rho3);


/* ---- triangulation of new map points (Map::TriangulateMappoint, src/map.cc:292-339) ----
 * Point i has the observations obs_begin[i] .. obs_begin[i+1]: keyframe index obs_frame[o] and left-image pixel
 * obs_uv[o], obs_uv[n_obs + o]. frame_twc [7][n_frames]: p (3), q (x, y, z, w) of Frame::GetPose(). cam5 = fx, fy, cx,
 * cy, bf (bearing = R * ((u - cx) / fx, (v - cy) / fy, 1), camera.cc:150-155). out_ok[i] = the reference's bool
 * (fewer than two observations or rank < 3 of the 3 x 3 normal matrix under Eigen's ColPivHouseholderQR with
 * threshold 1e-5: false, out_xyz untouched). Returns the number of points triangulated. */
int orc_triangulate_points(int32_t n_points, const int32_t* obs_begin, const int32_t* obs_frame, const double* obs_uv,
                           int32_t n_obs, const double* frame_twc, int32_t n_frames, const double* cam5, double* out_xyz,
                           uint8_t* out_ok);

/* --- Map::UppdateMapline (/root/reference/src/map.cc:121-177), the line endpoint refresh that follows the local BA
 * (map.cc:797), for a batch of lines. line_wd [6][n_lines] g2o::Line3D [w, d]; line l has the map points
 * pt_index[pt_begin[l] .. pt_begin[l+1] - 1] (the valid map points on the line in its observers, map.cc:128-139) in
 * point_xyz [3][n_points]. endpoints [6][n_lines] (first the end at the largest main-direction coordinate) is written
 * only where out_ok[l] = 1, the reference's return value. Returns the number of lines refreshed. */
void orc_line_to_cartesian(const double* wd6, double* cart6);
int orc_update_maplines(int32_t n_lines, const double* line_wd, const int32_t* pt_begin, const int32_t* pt_index,
                        const double* point_xyz, int32_t n_points, double* endpoints, uint8_t* out_ok);

#ifdef __cplusplus
}
#endif
#endif /* RSPL_ORACLE_H_ */
