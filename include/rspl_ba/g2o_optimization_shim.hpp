// g2o_optimization_shim.hpp — header-only C++ binding of the reference's optimisation entry points
// onto the C-ABI of include/rspl_ba.h.
//
// Replaces (signatures unchanged, so src/map.cc:709 and src/map_builder.cc:583 compile as they are):
//   void LocalmapOptimization(MapOfPoses&, MapOfPoints3d&, MapOfLine3d&, std::vector<CameraPtr>&,
//                             VectorOfMonoPointConstraints&, VectorOfStereoPointConstraints&,
//                             VectorOfMonoLineConstraints&, VectorOfStereoLineConstraints&,
//                             const OptimizationConfig&)          include/g2o_optimization/g2o_optimization.h:15-18
//   int  FrameOptimization(MapOfPoses&, MapOfPoints3d&, std::vector<CameraPtr>&,
//                          VectorOfMonoPointConstraints&, VectorOfStereoPointConstraints&,
//                          const OptimizationConfig&)             include/g2o_optimization/g2o_optimization.h:20-22
//
// The shim is written against the *shape* of the reference's boundary types
// (include/g2o_optimization/types.h:19-174), not against Eigen / g2o themselves, so it compiles
// both inside the reference tree (Eigen::Vector3d, Eigen::Quaterniond, g2o::Line3D) and against the
// layout-identical mock types of tests/shim/mock_types.h (neither Eigen nor g2o is installed in
// the build container). Required of the types:
//   Pose3d        .fixed, .p(i) read/write, .q.x()/.y()/.z()/.w() read/write
//   Position3d    .p(i)
//   Line3d        .line_3d(i), i = 0..5  (g2o::Line3D is a Vector6d: [w, d])
//   *Constraint   ->id_pose, ->id_point | ->id_line, ->id_camera, ->inlier, ->keypoint(i) | ->line_2d(i)
//   Camera        ->Fx(), ->Fy(), ->Cx(), ->Cy(), ->BF()          (include/camera.h:25-29)
//   cfg           .mono_point, .stereo_point, .mono_line, .stereo_line (include/read_configs.h:50-56)
//
// What it does, per call: id compaction through the ordered maps (vertex index = position in
// std::map order, which is also g2o's id-sorted vertex order, g2o_optimization.cc:39-70), flattening
// into structure-of-arrays planes, one rspl_ba_local_batch / rspl_ba_frame_batch call (n = 1), and
// in-place write-back of p, q, line_3d and ->inlier exactly where the reference writes them
// (g2o_optimization.cc:213-251, :345-396). Errors: the reference surfaces none (void / int); the
// shim returns the C-ABI status from the *Impl functions and the reference-signature wrappers
// ignore it (FrameOptimization returns 0 inliers on failure, so tracking treats the frame as lost).
//
// To generate the two reference-signature functions, define RSPL_BA_DEFINE_REFERENCE_ENTRY_POINTS
// before including this header in ONE translation unit that has already included the reference's
// "g2o_optimization/types.h", "camera.h" and "read_configs.h" (see INTEGRATION.md).
#ifndef RSPL_BA_G2O_OPTIMIZATION_SHIM_HPP_
#define RSPL_BA_G2O_OPTIMIZATION_SHIM_HPP_

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <vector>

#include "../rspl_ba.h"

namespace rspl_ba {

// One context per calling thread (the reference calls both entry points from _tracking_thread only,
// src/map_builder.cc:49); created on first use, destroyed with the thread.
struct ThreadContext {
  RsplBaContext* ctx = nullptr;
  int status = RSPL_BA_OK;
  ThreadContext() { status = rspl_ba_create(-1, nullptr, &ctx); }
  ~ThreadContext() {
    if (ctx) rspl_ba_destroy(ctx);
  }
  ThreadContext(const ThreadContext&) = delete;
  ThreadContext& operator=(const ThreadContext&) = delete;
};
// What happens when the library refuses or fails a call. The reference has no error channel (void / int returns,
// asserts), and g2o accepts inputs this library does not (limits below), so a failure must not pass silently as
// "nothing optimised" / "0 inliers": the default handler prints rspl_ba_last_error and aborts, like the reference's
// own assert(poses.size() == 1) (g2o_optimization.cc:259). Install another handler to log and carry on instead.
//   Limits (RSPL_BA_ERR_UNSUPPORTED): at most one constraint per (pose, landmark) pair; at most 254 observations of
//   a landmark in windows of up to 64 free poses; at most 255 cameras, 65535 poses per window, 65535 windows per call.
//   RSPL_BA_ERR_CUDA: no CUDA device / driver error (there is no CPU fallback). RSPL_BA_ERR_INVALID: an id that is
//   not in the containers (the reference would dereference a null vertex, g2o_optimization.cc:83).
using FailureHandler = void (*)(const char* entry_point, int code, const char* message);
inline void default_failure_handler(const char* entry_point, int code, const char* message) {
  std::fprintf(stderr, "rspl_ba: %s failed with status %d: %s\n", entry_point, code, message ? message : "");
  std::abort();
}
inline FailureHandler& failure_handler() {
  static FailureHandler h = default_failure_handler;
  return h;
}
inline void set_failure_handler(FailureHandler h) { failure_handler() = h ? h : default_failure_handler; }

inline RsplBaContext* thread_context() {
  static thread_local ThreadContext tc;
  return tc.ctx;
}

namespace detail {

// id -> position in the ordered container (the compact index the C-ABI takes). The ids of a window are nearly
// contiguous (frame / mappoint / mapline counters), so a dense table answers in one load; sparse id sets fall back to a
// binary search over the sorted keys. (A std::map lookup per constraint was 1 ms of a 13 k-constraint window.)
class IdIndex {
 public:
  template <class Map>
  explicit IdIndex(const Map& m) {
    keys_.reserve(m.size());
    for (const auto& kv : m) keys_.push_back(kv.first); // ascending: std::map order
    if (keys_.empty()) return;
    lo_ = keys_.front();
    const long long span = (long long)keys_.back() - lo_ + 1;
    if (span <= 8LL * (long long)keys_.size() + 4096) {
      dense_.assign((size_t)span, -1);
      for (size_t i = 0; i < keys_.size(); ++i) dense_[(size_t)((long long)keys_[i] - lo_)] = (int32_t)i;
    }
  }
  // position of id, or -1
  int32_t find(int id) const {
    if (!dense_.empty()) {
      const long long k = (long long)id - lo_;
      return k < 0 || k >= (long long)dense_.size() ? -1 : dense_[(size_t)k];
    }
    const auto it = std::lower_bound(keys_.begin(), keys_.end(), id);
    return it == keys_.end() || *it != id ? -1 : (int32_t)(it - keys_.begin());
  }

 private:
  std::vector<int> keys_;
  std::vector<int32_t> dense_;
  long long lo_ = 0;
};

template <class CameraList>
std::vector<double> flatten_cameras(const CameraList& cams) {
  std::vector<double> out;
  out.reserve(cams.size() * 5);
  for (const auto& c : cams) {
    out.push_back(c->Fx());
    out.push_back(c->Fy());
    out.push_back(c->Cx());
    out.push_back(c->Cy());
    out.push_back(c->BF());
  }
  return out;
}

template <class Cfg>
RsplBaOptions make_options(const Cfg& cfg) {
  RsplBaOptions o;
  rspl_ba_default_options(&o);
  o.thr_mono_point = cfg.mono_point;
  o.thr_stereo_point = cfg.stereo_point;
  o.thr_mono_line = cfg.mono_line;
  o.thr_stereo_line = cfg.stereo_line;
  return o;
}

// one constraint class -> pose / landmark index arrays + measurement planes
template <class Vec, class GetLm, class GetMeas>
int flatten_edges(const Vec& cons, int dim, const IdIndex& pose_idx, const IdIndex& lm_idx,
                  GetLm get_lm, GetMeas get_meas, std::vector<int32_t>& pose, std::vector<int32_t>& lm,
                  std::vector<int32_t>& cam, std::vector<double>& meas) {
  const size_t n = cons.size();
  pose.resize(n);
  lm.resize(n);
  cam.resize(n);
  meas.assign(n * dim, 0.0);
  for (size_t i = 0; i < n; ++i) {
    const auto& c = cons[i];
    const int32_t pi = pose_idx.find(c->id_pose), li = lm_idx.find(get_lm(*c));
    if (pi < 0 || li < 0) return RSPL_BA_ERR_INVALID; // the reference would null-deref here
    pose[i] = pi;
    lm[i] = li;
    cam[i] = c->id_camera;
    for (int k = 0; k < dim; ++k) meas[(size_t)k * n + i] = get_meas(*c, k);
  }
  return RSPL_BA_OK;
}

} // namespace detail

// ------------------------------------------------------------------------------------------------
// LocalmapOptimization
// ------------------------------------------------------------------------------------------------
template <class MapOfPosesT, class MapOfPointsT, class MapOfLinesT, class CameraListT, class MonoPtT, class StereoPtT,
          class MonoLnT, class StereoLnT, class CfgT>
int LocalmapOptimizationImpl(RsplBaContext* ctx, MapOfPosesT& poses, MapOfPointsT& points, MapOfLinesT& lines,
                             CameraListT& camera_list, MonoPtT& mono_point_constraints,
                             StereoPtT& stereo_point_constraints, MonoLnT& mono_line_constraints,
                             StereoLnT& stereo_line_constraints, const CfgT& cfg) {
  if (!ctx) return RSPL_BA_ERR_CUDA;
  const int np = (int)poses.size(), npt = (int)points.size(), nln = (int)lines.size();
  const detail::IdIndex pose_idx(poses), point_idx(points), line_idx(lines);
  std::vector<double> pose_twc((size_t)7 * np), point_xyz((size_t)3 * npt), line_wd((size_t)6 * nln);
  std::vector<uint8_t> pose_fixed(np);
  {
    int i = 0;
    for (const auto& kv : poses) {
      for (int k = 0; k < 3; ++k) pose_twc[(size_t)k * np + i] = kv.second.p(k);
      pose_twc[(size_t)3 * np + i] = kv.second.q.x();
      pose_twc[(size_t)4 * np + i] = kv.second.q.y();
      pose_twc[(size_t)5 * np + i] = kv.second.q.z();
      pose_twc[(size_t)6 * np + i] = kv.second.q.w();
      pose_fixed[i] = kv.second.fixed ? 1 : 0;
      ++i;
    }
    i = 0;
    for (const auto& kv : points) {
      for (int k = 0; k < 3; ++k) point_xyz[(size_t)k * npt + i] = kv.second.p(k);
      ++i;
    }
    i = 0;
    for (const auto& kv : lines) {
      for (int k = 0; k < 6; ++k) line_wd[(size_t)k * nln + i] = kv.second.line_3d(k);
      ++i;
    }
  }
  const std::vector<double> cams = detail::flatten_cameras(camera_list);
  std::vector<int32_t> mp_pose, mp_lm, mp_cam, sp_pose, sp_lm, sp_cam, ml_pose, ml_lm, ml_cam, sl_pose, sl_lm, sl_cam;
  std::vector<double> mp_meas, sp_meas, ml_meas, sl_meas;
  int rc = detail::flatten_edges(
      mono_point_constraints, 2, pose_idx, point_idx, [](const auto& c) { return c.id_point; },
      [](const auto& c, int k) { return c.keypoint(k); }, mp_pose, mp_lm, mp_cam, mp_meas);
  if (rc == RSPL_BA_OK)
    rc = detail::flatten_edges(
        stereo_point_constraints, 3, pose_idx, point_idx, [](const auto& c) { return c.id_point; },
        [](const auto& c, int k) { return c.keypoint(k); }, sp_pose, sp_lm, sp_cam, sp_meas);
  if (rc == RSPL_BA_OK)
    rc = detail::flatten_edges(
        mono_line_constraints, 4, pose_idx, line_idx, [](const auto& c) { return c.id_line; },
        [](const auto& c, int k) { return c.line_2d(k); }, ml_pose, ml_lm, ml_cam, ml_meas);
  if (rc == RSPL_BA_OK)
    rc = detail::flatten_edges(
        stereo_line_constraints, 8, pose_idx, line_idx, [](const auto& c) { return c.id_line; },
        [](const auto& c, int k) { return c.line_2d(k); }, sl_pose, sl_lm, sl_cam, sl_meas);
  if (rc != RSPL_BA_OK) return rc;

  const int32_t pose_begin[2] = {0, np}, point_begin[2] = {0, npt}, line_begin[2] = {0, nln};
  const int32_t mp_begin[2] = {0, (int32_t)mp_pose.size()}, sp_begin[2] = {0, (int32_t)sp_pose.size()};
  const int32_t ml_begin[2] = {0, (int32_t)ml_pose.size()}, sl_begin[2] = {0, (int32_t)sl_pose.size()};
  RsplLocalBatch in{};
  in.n_windows = 1;
  in.n_cameras = (int32_t)camera_list.size();
  in.cameras = cams.data();
  in.pose_begin = pose_begin;
  in.point_begin = point_begin;
  in.line_begin = line_begin;
  in.mono_pt_begin = mp_begin;
  in.stereo_pt_begin = sp_begin;
  in.mono_ln_begin = ml_begin;
  in.stereo_ln_begin = sl_begin;
  in.pose_twc = pose_twc.data();
  in.pose_fixed = pose_fixed.data();
  in.point_xyz = point_xyz.data();
  in.line_wd = line_wd.data();
  in.mp_pose = mp_pose.data();
  in.mp_point = mp_lm.data();
  in.mp_cam = mp_cam.data();
  in.mp_meas = mp_meas.data();
  in.sp_pose = sp_pose.data();
  in.sp_point = sp_lm.data();
  in.sp_cam = sp_cam.data();
  in.sp_meas = sp_meas.data();
  in.ml_pose = ml_pose.data();
  in.ml_line = ml_lm.data();
  in.ml_cam = ml_cam.data();
  in.ml_meas = ml_meas.data();
  in.sl_pose = sl_pose.data();
  in.sl_line = sl_lm.data();
  in.sl_cam = sl_cam.data();
  in.sl_meas = sl_meas.data();

  std::vector<double> o_pose((size_t)7 * np), o_pt((size_t)3 * npt), o_ln((size_t)6 * nln);
  std::vector<uint8_t> o_mp(mp_pose.size() + 1), o_sp(sp_pose.size() + 1), o_ml(ml_pose.size() + 1), o_sl(sl_pose.size() + 1);
  RsplLocalBatchResult out{};
  out.pose_twc = o_pose.data();
  out.point_xyz = o_pt.data();
  out.line_wd = o_ln.data();
  out.mp_inlier = o_mp.data();
  out.sp_inlier = o_sp.data();
  out.ml_inlier = o_ml.data();
  out.sl_inlier = o_sl.data();
  const RsplBaOptions opt = detail::make_options(cfg);
  rc = rspl_ba_local_batch(ctx, &in, &opt, &out);
  if (rc != RSPL_BA_OK) return rc;

  // write-back (g2o_optimization.cc:213-251)
  for (size_t i = 0; i < mono_point_constraints.size(); ++i) mono_point_constraints[i]->inlier = o_mp[i] != 0;
  for (size_t i = 0; i < stereo_point_constraints.size(); ++i) stereo_point_constraints[i]->inlier = o_sp[i] != 0;
  for (size_t i = 0; i < mono_line_constraints.size(); ++i) mono_line_constraints[i]->inlier = o_ml[i] != 0;
  for (size_t i = 0; i < stereo_line_constraints.size(); ++i) stereo_line_constraints[i]->inlier = o_sl[i] != 0;
  {
    int i = 0;
    for (auto& kv : poses) {
      for (int k = 0; k < 3; ++k) kv.second.p(k) = o_pose[(size_t)k * np + i];
      kv.second.q.x() = o_pose[(size_t)3 * np + i];
      kv.second.q.y() = o_pose[(size_t)4 * np + i];
      kv.second.q.z() = o_pose[(size_t)5 * np + i];
      kv.second.q.w() = o_pose[(size_t)6 * np + i];
      ++i;
    }
    i = 0;
    for (auto& kv : points) {
      for (int k = 0; k < 3; ++k) kv.second.p(k) = o_pt[(size_t)k * npt + i];
      ++i;
    }
    i = 0;
    for (auto& kv : lines) {
      for (int k = 0; k < 6; ++k) kv.second.line_3d(k) = o_ln[(size_t)k * nln + i];
      ++i;
    }
  }
  return RSPL_BA_OK;
}

// ------------------------------------------------------------------------------------------------
// FrameOptimization: returns the reference's int (number of inliers) through *num_inliers
// ------------------------------------------------------------------------------------------------
namespace detail {
// stand-ins for "no line containers" (the reference's FrameOptimization has none)
struct NoLineConstraint {
  int id_line = 0, id_camera = 0;
  bool inlier = true;
  double line_2d(int) const { return 0.0; }
};
struct NoLine {
  double line_3d(int) const { return 0.0; }
};
} // namespace detail

// FrameOptimization plus constraints on FIXED map lines (extension: the reference's FrameOptimization takes
// no line containers, g2o_optimization.h:20-22). `lines` is the reference's MapOfLine3d, the constraint
// vectors its VectorOfMonoLineConstraints / VectorOfStereoLineConstraints (types.h:124-174); their
// id_pose is ignored (there is one pose), ->inlier is read and written like the point constraints'.
template <class MapOfPosesT, class MapOfPointsT, class MapOfLinesT, class CameraListT, class MonoPtT, class StereoPtT,
          class MonoLnT, class StereoLnT, class CfgT>
int FrameOptimizationWithLinesImpl(RsplBaContext* ctx, MapOfPosesT& poses, MapOfPointsT& points, MapOfLinesT& lines,
                                   CameraListT& camera_list, MonoPtT& mono_point_constraints,
                                   StereoPtT& stereo_point_constraints, MonoLnT& mono_line_constraints,
                                   StereoLnT& stereo_line_constraints, const CfgT& cfg, int* num_inliers);

template <class MapOfPosesT, class MapOfPointsT, class CameraListT, class MonoPtT, class StereoPtT, class CfgT>
int FrameOptimizationImpl(RsplBaContext* ctx, MapOfPosesT& poses, MapOfPointsT& points, CameraListT& camera_list,
                          MonoPtT& mono_point_constraints, StereoPtT& stereo_point_constraints, const CfgT& cfg,
                          int* num_inliers) {
  std::map<int, detail::NoLine> no_lines;
  std::vector<std::shared_ptr<detail::NoLineConstraint>> no_ml, no_sl;
  return FrameOptimizationWithLinesImpl(ctx, poses, points, no_lines, camera_list, mono_point_constraints,
                                        stereo_point_constraints, no_ml, no_sl, cfg, num_inliers);
}

template <class MapOfPosesT, class MapOfPointsT, class MapOfLinesT, class CameraListT, class MonoPtT, class StereoPtT,
          class MonoLnT, class StereoLnT, class CfgT>
int FrameOptimizationWithLinesImpl(RsplBaContext* ctx, MapOfPosesT& poses, MapOfPointsT& points, MapOfLinesT& lines,
                                   CameraListT& camera_list, MonoPtT& mono_point_constraints,
                                   StereoPtT& stereo_point_constraints, MonoLnT& mono_line_constraints,
                                   StereoLnT& stereo_line_constraints, const CfgT& cfg, int* num_inliers) {
  if (num_inliers) *num_inliers = 0;
  if (!ctx) return RSPL_BA_ERR_CUDA;
  if (poses.size() != 1) return RSPL_BA_ERR_INVALID; // assert(poses.size() == 1), g2o_optimization.cc:259
  auto& pose = poses.begin()->second;
  const double pose_twc[7] = {pose.p(0), pose.p(1), pose.p(2), pose.q.x(), pose.q.y(), pose.q.z(), pose.q.w()};
  const std::vector<double> cams = detail::flatten_cameras(camera_list);
  const size_t nm = mono_point_constraints.size(), ns = stereo_point_constraints.size();
  std::vector<double> m_meas(2 * nm), m_xw(3 * nm), s_meas(3 * ns), s_xw(3 * ns);
  std::vector<int32_t> m_cam(nm), s_cam(ns);
  std::vector<uint8_t> m_inl(nm + 1), s_inl(ns + 1);
  for (size_t i = 0; i < nm; ++i) {
    const auto& c = mono_point_constraints[i];
    auto it = points.find(c->id_point); // Position3d point = points[mpc->id_point] (:289)
    if (it == points.end()) return RSPL_BA_ERR_INVALID;
    for (int k = 0; k < 2; ++k) m_meas[k * nm + i] = c->keypoint(k);
    for (int k = 0; k < 3; ++k) m_xw[k * nm + i] = it->second.p(k);
    m_cam[i] = c->id_camera;
    m_inl[i] = c->inlier ? 1 : 0;
  }
  for (size_t i = 0; i < ns; ++i) {
    const auto& c = stereo_point_constraints[i];
    auto it = points.find(c->id_point); // (:314)
    if (it == points.end()) return RSPL_BA_ERR_INVALID;
    for (int k = 0; k < 3; ++k) s_meas[k * ns + i] = c->keypoint(k);
    for (int k = 0; k < 3; ++k) s_xw[k * ns + i] = it->second.p(k);
    s_cam[i] = c->id_camera;
    s_inl[i] = c->inlier ? 1 : 0;
  }
  // line constraints (empty in the reference's path): the fixed world line is copied into the edge
  const size_t nml = mono_line_constraints.size(), nsl = stereo_line_constraints.size();
  std::vector<double> ml_lw(6 * nml), ml_meas(4 * nml), sl_lw(6 * nsl), sl_meas(8 * nsl);
  std::vector<int32_t> ml_cam(nml), sl_cam(nsl);
  std::vector<uint8_t> ml_inl(nml + 1), sl_inl(nsl + 1);
  for (size_t i = 0; i < nml; ++i) {
    const auto& c = mono_line_constraints[i];
    auto it = lines.find(c->id_line);
    if (it == lines.end()) return RSPL_BA_ERR_INVALID;
    for (int k = 0; k < 6; ++k) ml_lw[k * nml + i] = it->second.line_3d(k);
    for (int k = 0; k < 4; ++k) ml_meas[k * nml + i] = c->line_2d(k);
    ml_cam[i] = c->id_camera;
    ml_inl[i] = c->inlier ? 1 : 0;
  }
  for (size_t i = 0; i < nsl; ++i) {
    const auto& c = stereo_line_constraints[i];
    auto it = lines.find(c->id_line);
    if (it == lines.end()) return RSPL_BA_ERR_INVALID;
    for (int k = 0; k < 6; ++k) sl_lw[k * nsl + i] = it->second.line_3d(k);
    for (int k = 0; k < 8; ++k) sl_meas[k * nsl + i] = c->line_2d(k);
    sl_cam[i] = c->id_camera;
    sl_inl[i] = c->inlier ? 1 : 0;
  }
  const int32_t mline_begin[2] = {0, (int32_t)nml}, sline_begin[2] = {0, (int32_t)nsl};
  const int32_t mono_begin[2] = {0, (int32_t)nm}, stereo_begin[2] = {0, (int32_t)ns};
  RsplFrameBatch in{};
  in.n_frames = 1;
  in.n_cameras = (int32_t)camera_list.size();
  in.cameras = cams.data();
  in.pose_twc = pose_twc;
  in.mono_begin = mono_begin;
  in.stereo_begin = stereo_begin;
  in.mono_meas = m_meas.data();
  in.mono_xw = m_xw.data();
  in.mono_cam = m_cam.data();
  in.mono_inlier = m_inl.data();
  in.stereo_meas = s_meas.data();
  in.stereo_xw = s_xw.data();
  in.stereo_cam = s_cam.data();
  in.stereo_inlier = s_inl.data();
  if (nml + nsl > 0) {
    in.mono_line_begin = mline_begin;
    in.stereo_line_begin = sline_begin;
    in.mono_line_lw = ml_lw.data();
    in.mono_line_meas = ml_meas.data();
    in.mono_line_cam = ml_cam.data();
    in.mono_line_inlier = ml_inl.data();
    in.stereo_line_lw = sl_lw.data();
    in.stereo_line_meas = sl_meas.data();
    in.stereo_line_cam = sl_cam.data();
    in.stereo_line_inlier = sl_inl.data();
  }
  double o_pose[7];
  int32_t n_inl = 0;
  std::vector<uint8_t> o_m(nm + 1), o_s(ns + 1), o_ml(nml + 1), o_sl(nsl + 1);
  RsplFrameBatchResult out{};
  out.pose_twc = o_pose;
  out.mono_inlier = o_m.data();
  out.stereo_inlier = o_s.data();
  out.mono_line_inlier = o_ml.data();
  out.stereo_line_inlier = o_sl.data();
  out.num_inliers = &n_inl;
  RsplBaOptions opt = detail::make_options(cfg);
  opt.frame_latency_mode = 1; // one frame per call (map_builder.cc:583-584): the whole CTA works on it
  const int rc = rspl_ba_frame_batch(ctx, &in, &opt, &out);
  if (rc != RSPL_BA_OK) return rc;
  for (size_t i = 0; i < nm; ++i) mono_point_constraints[i]->inlier = o_m[i] != 0;
  for (size_t i = 0; i < ns; ++i) stereo_point_constraints[i]->inlier = o_s[i] != 0;
  for (size_t i = 0; i < nml; ++i) mono_line_constraints[i]->inlier = o_ml[i] != 0;
  for (size_t i = 0; i < nsl; ++i) stereo_line_constraints[i]->inlier = o_sl[i] != 0;
  for (int k = 0; k < 3; ++k) pose.p(k) = o_pose[k]; // :391-393
  pose.q.x() = o_pose[3];
  pose.q.y() = o_pose[4];
  pose.q.z() = o_pose[5];
  pose.q.w() = o_pose[6];
  if (num_inliers) *num_inliers = n_inl;
  return RSPL_BA_OK;
}

} // namespace rspl_ba

#ifdef RSPL_BA_DEFINE_REFERENCE_ENTRY_POINTS
// The reference's own declarations (include/g2o_optimization/g2o_optimization.h:15-22), defined here.
inline void LocalmapOptimization(MapOfPoses& poses, MapOfPoints3d& points, MapOfLine3d& lines,
                                 std::vector<CameraPtr>& camera_list,
                                 VectorOfMonoPointConstraints& mono_point_constraints,
                                 VectorOfStereoPointConstraints& stereo_point_constraints,
                                 VectorOfMonoLineConstraints& mono_line_constraints,
                                 VectorOfStereoLineConstraints& stereo_line_constraints, const OptimizationConfig& cfg) {
  RsplBaContext* ctx = rspl_ba::thread_context();
  const int rc = rspl_ba::LocalmapOptimizationImpl(ctx, poses, points, lines, camera_list, mono_point_constraints,
                                                   stereo_point_constraints, mono_line_constraints, stereo_line_constraints, cfg);
  if (rc != RSPL_BA_OK) rspl_ba::failure_handler()("LocalmapOptimization", rc, ctx ? rspl_ba_last_error(ctx) : "no CUDA context");
}

inline int FrameOptimization(MapOfPoses& poses, MapOfPoints3d& points, std::vector<CameraPtr>& camera_list,
                             VectorOfMonoPointConstraints& mono_point_constraints,
                             VectorOfStereoPointConstraints& stereo_point_constraints, const OptimizationConfig& cfg) {
  int n = 0;
  RsplBaContext* ctx = rspl_ba::thread_context();
  const int rc = rspl_ba::FrameOptimizationImpl(ctx, poses, points, camera_list, mono_point_constraints, stereo_point_constraints,
                                                cfg, &n);
  if (rc != RSPL_BA_OK) rspl_ba::failure_handler()("FrameOptimization", rc, ctx ? rspl_ba_last_error(ctx) : "no CUDA context");
  return n;
}
#endif // RSPL_BA_DEFINE_REFERENCE_ENTRY_POINTS

#endif // RSPL_BA_G2O_OPTIMIZATION_SHIM_HPP_
