// shim_driver.cpp — calls the reference-signature LocalmapOptimization / FrameOptimization provided
// by include/rspl_ba/g2o_optimization_shim.hpp on a problem read from a flat float64 file and writes
// the mutated containers back. TEST INFRASTRUCTURE (see tests/test_shim.py).
// With a third argument N it also times N further calls on fresh copies of the containers (bench.py's
// single-call latency through the reference signatures) and prints {"median_us", "min_us"} to stdout.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#ifdef RSPL_REFERENCE_BUILD
// oracle/g2o_validation: the SAME driver compiled against the reference's own headers and linked with the unmodified
// g2o_optimization.cc + edge sources and a real g2o (fixtures that pin the oracle against the reference itself)
#include "g2o_optimization/g2o_optimization.h"
#else
#include "mock_types.h"
#define RSPL_BA_DEFINE_REFERENCE_ENTRY_POINTS
#include "rspl_ba/g2o_optimization_shim.hpp"
#endif

static std::vector<double> read_all(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<double> v(n / 8);
  if (fread(v.data(), 8, v.size(), f) != v.size()) exit(2);
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::vector<double> in = read_all(argv[1]);
  size_t k = 0;
  auto next = [&]() { return in[k++]; };
  const int kind = (int)next();
  const int np = (int)next(), npt = (int)next(), nln = (int)next();
  const int nmp = (int)next(), nsp = (int)next(), nml = (int)next(), nsl = (int)next();
  OptimizationConfig cfg{next(), next(), next(), next(), 0.5};
#ifdef RSPL_REFERENCE_BUILD
  // the reference's Camera reads its intrinsics from a calibration file: RSPL_REF_CAMERA_YAML = <reference>/configs/euroc.yaml,
  // whose LEFT.P / bf are the EuRoC values the synthetic generator uses (the five numbers in the dump are checked against it)
  const double cam5[5] = {next(), next(), next(), next(), next()};
  const char* cam_yaml = getenv("RSPL_REF_CAMERA_YAML");
  if (!cam_yaml) { fprintf(stderr, "set RSPL_REF_CAMERA_YAML\n"); return 2; }
  std::vector<CameraPtr> cams{std::make_shared<Camera>(std::string(cam_yaml))};
  if (std::abs(cams[0]->Fx() - cam5[0]) > 1e-9 || std::abs(cams[0]->Fy() - cam5[1]) > 1e-9 || std::abs(cams[0]->Cx() - cam5[2]) > 1e-9 ||
      std::abs(cams[0]->Cy() - cam5[3]) > 1e-9 || std::abs(cams[0]->BF() - cam5[4]) > 1e-9) {
    fprintf(stderr, "camera file does not match the dumped intrinsics\n");
    return 2;
  }
#else
  std::vector<CameraPtr> cams{std::make_shared<Camera>(Camera{next(), next(), next(), next(), next()})};
#endif
  MapOfPoses poses;
  MapOfPoints3d points;
  MapOfLine3d lines;
  for (int i = 0; i < np; ++i) {
    const int id = (int)next();
    Pose3d p;
    p.fixed = next() != 0;
    for (int c = 0; c < 3; ++c) p.p(c) = next();
    p.q.x() = next(); p.q.y() = next(); p.q.z() = next(); p.q.w() = next();
    poses[id] = p;
  }
  for (int i = 0; i < npt; ++i) {
    const int id = (int)next();
    Position3d p;
    p.fixed = false;
    for (int c = 0; c < 3; ++c) p.p(c) = next();
    points[id] = p;
  }
  for (int i = 0; i < nln; ++i) {
    const int id = (int)next();
    Line3d l;
    l.fixed = false;
    for (int c = 0; c < 6; ++c) l.line_3d(c) = next();
    lines[id] = l;
  }
  VectorOfMonoPointConstraints mp;
  VectorOfStereoPointConstraints sp;
  VectorOfMonoLineConstraints ml;
  VectorOfStereoLineConstraints sl;
  for (int i = 0; i < nmp; ++i) {
    auto c = std::make_shared<MonoPointConstraint>();
    c->id_pose = (int)next(); c->id_point = (int)next(); c->id_camera = 0; c->inlier = next() != 0; c->pixel_sigma = 0.8;
    for (int q = 0; q < 2; ++q) c->keypoint(q) = next();
    mp.push_back(c);
  }
  for (int i = 0; i < nsp; ++i) {
    auto c = std::make_shared<StereoPointConstraint>();
    c->id_pose = (int)next(); c->id_point = (int)next(); c->id_camera = 0; c->inlier = next() != 0; c->pixel_sigma = 0.8;
    for (int q = 0; q < 3; ++q) c->keypoint(q) = next();
    sp.push_back(c);
  }
  for (int i = 0; i < nml; ++i) {
    auto c = std::make_shared<MonoLineConstraint>();
    c->id_pose = (int)next(); c->id_line = (int)next(); c->id_camera = 0; c->inlier = next() != 0; c->pixel_sigma = 0.8;
    for (int q = 0; q < 4; ++q) c->line_2d(q) = next();
    ml.push_back(c);
  }
  for (int i = 0; i < nsl; ++i) {
    auto c = std::make_shared<StereoLineConstraint>();
    c->id_pose = (int)next(); c->id_line = (int)next(); c->id_camera = 0; c->inlier = next() != 0; c->pixel_sigma = 0.8;
    for (int q = 0; q < 8; ++q) c->line_2d(q) = next();
    sl.push_back(c);
  }
  double ret = 0;
  auto call = [&](MapOfPoses& P, MapOfPoints3d& X, MapOfLine3d& L, VectorOfMonoPointConstraints& a,
                  VectorOfStereoPointConstraints& b, VectorOfMonoLineConstraints& c, VectorOfStereoLineConstraints& e) {
    double r = 0;
    if (kind == 0) {
      LocalmapOptimization(P, X, L, cams, a, b, c, e, cfg);
    } else if (kind == 1) {
      r = FrameOptimization(P, X, cams, a, b, cfg);
    } else { // kind 2: pose-only with constraints on fixed lines (the extension; no reference signature)
#ifndef RSPL_REFERENCE_BUILD
      int n = 0;
      (void)rspl_ba::FrameOptimizationWithLinesImpl(rspl_ba::thread_context(), P, X, L, cams, a, b, c, e, cfg, &n);
      r = n;
#else
      fprintf(stderr, "kind 2 (line extension) has no counterpart in the reference\n");
      exit(2);
#endif
    }
    return r;
  };
  const int reps = argc > 3 ? atoi(argv[3]) : 0;
  if (reps > 0) { // latency of one call through the reference signature (containers copied outside the timed region)
    auto deep = [](const auto& v) {
      typename std::decay<decltype(v)>::type o;
      for (auto& p : v) o.push_back(std::make_shared<typename std::decay<decltype(*p)>::type>(*p));
      return o;
    };
    std::vector<double> us;
    for (int i = 0; i < reps + 3; ++i) {
      MapOfPoses P = poses;
      MapOfPoints3d X = points;
      MapOfLine3d L = lines;
      auto a = deep(mp);
      auto b = deep(sp);
      auto c = deep(ml);
      auto e = deep(sl);
      const auto t0 = std::chrono::steady_clock::now();
      call(P, X, L, a, b, c, e);
      const auto t1 = std::chrono::steady_clock::now();
      if (i >= 3) us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
    }
    std::sort(us.begin(), us.end());
    printf("{\"median_us\": %.2f, \"min_us\": %.2f, \"reps\": %d}\n", us[us.size() / 2], us[0], reps);
  }
  ret = call(poses, points, lines, mp, sp, ml, sl);
  std::vector<double> out;
  out.push_back(ret);
  for (auto& kv : poses) {
    for (int c = 0; c < 3; ++c) out.push_back(kv.second.p(c));
    out.push_back(kv.second.q.x()); out.push_back(kv.second.q.y()); out.push_back(kv.second.q.z()); out.push_back(kv.second.q.w());
  }
  for (auto& kv : points) for (int c = 0; c < 3; ++c) out.push_back(kv.second.p(c));
  for (auto& kv : lines) for (int c = 0; c < 6; ++c) out.push_back(kv.second.line_3d(c));
  for (auto& c : mp) out.push_back(c->inlier ? 1 : 0);
  for (auto& c : sp) out.push_back(c->inlier ? 1 : 0);
  for (auto& c : ml) out.push_back(c->inlier ? 1 : 0);
  for (auto& c : sl) out.push_back(c->inlier ? 1 : 0);
  FILE* f = fopen(argv[2], "wb");
  if (!f) return 2;
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  return 0;
}
