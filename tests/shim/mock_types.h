// mock_types.h — layout-compatible stand-ins for the reference's boundary types
// (/root/reference/include/g2o_optimization/types.h:19-174, include/camera.h:25-29,
// include/read_configs.h:50-56) so the shim can be compiled and exercised without Eigen / g2o,
// neither of which is installed here. Same type names, same member names, same accessor syntax
// (v(i), q.x() ... q.w()) as the Eigen types the reference uses. TEST INFRASTRUCTURE.
#ifndef RSPL_BA_MOCK_TYPES_H_
#define RSPL_BA_MOCK_TYPES_H_

#include <map>
#include <memory>
#include <vector>

template <int N>
struct MockVec { // Eigen::Matrix<double, N, 1> stand-in
  double d[N] = {};
  double& operator()(int i) { return d[i]; }
  const double& operator()(int i) const { return d[i]; }
};
struct MockQuat { // Eigen::Quaterniond stand-in; storage order x, y, z, w like Eigen
  double c[4] = {0, 0, 0, 1};
  double& x() { return c[0]; }
  double& y() { return c[1]; }
  double& z() { return c[2]; }
  double& w() { return c[3]; }
  const double& x() const { return c[0]; }
  const double& y() const { return c[1]; }
  const double& z() const { return c[2]; }
  const double& w() const { return c[3]; }
};

struct Pose3d {
  bool fixed;
  MockVec<3> p;
  MockQuat q;
};
typedef std::map<int, Pose3d> MapOfPoses;
struct Position3d {
  bool fixed;
  MockVec<3> p;
};
typedef std::map<int, Position3d> MapOfPoints3d;
struct Line3d {
  bool fixed;
  MockVec<6> line_3d; // g2o::Line3D is a Vector6d [w, d]
};
typedef std::map<int, Line3d> MapOfLine3d;

struct MonoPointConstraint {
  int id_pose, id_point, id_camera;
  bool inlier;
  MockVec<2> keypoint;
  double pixel_sigma;
};
typedef std::shared_ptr<MonoPointConstraint> MonoPointConstraintPtr;
typedef std::vector<MonoPointConstraintPtr> VectorOfMonoPointConstraints;
struct StereoPointConstraint {
  int id_pose, id_point, id_camera;
  bool inlier;
  MockVec<3> keypoint;
  double pixel_sigma;
};
typedef std::shared_ptr<StereoPointConstraint> StereoPointConstraintPtr;
typedef std::vector<StereoPointConstraintPtr> VectorOfStereoPointConstraints;
struct MonoLineConstraint {
  int id_pose, id_line, id_camera;
  bool inlier;
  MockVec<4> line_2d;
  double pixel_sigma;
};
typedef std::shared_ptr<MonoLineConstraint> MonoLineConstraintPtr;
typedef std::vector<MonoLineConstraintPtr> VectorOfMonoLineConstraints;
struct StereoLineConstraint {
  int id_pose, id_line, id_camera;
  bool inlier;
  MockVec<8> line_2d;
  double pixel_sigma;
};
typedef std::shared_ptr<StereoLineConstraint> StereoLineConstraintPtr;
typedef std::vector<StereoLineConstraintPtr> VectorOfStereoLineConstraints;

struct Camera {
  double fx, fy, cx, cy, bf;
  double Fx() const { return fx; }
  double Fy() const { return fy; }
  double Cx() const { return cx; }
  double Cy() const { return cy; }
  double BF() const { return bf; }
};
typedef std::shared_ptr<Camera> CameraPtr;

struct OptimizationConfig {
  double mono_point, stereo_point, mono_line, stereo_line, rate;
};

#endif // RSPL_BA_MOCK_TYPES_H_
