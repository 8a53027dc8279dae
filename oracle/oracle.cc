/*
 * oracle/oracle.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see oracle.h).
 *
 * CPU restatement of the reference's bundle-adjustment path:
 *   LocalmapOptimization  — /root/reference/src/g2o_optimization/g2o_optimization.cc:21-252
 *   FrameOptimization     — /root/reference/src/g2o_optimization/g2o_optimization.cc:256-397
 *   line edges            — edge_project_line.cc:21-42, edge_project_stereo_line.cc:22-51
 *   line vertex           — include/g2o_optimization/vertex_line3d.h:26-43
 * and of the g2o machinery those call (g2o is a system dependency of the reference, located by
 * cmake/FindG2O.cmake, un-vendored and un-pinned; semantics follow SURVEY.md §9: SparseOptimizer
 * active sets, OptimizationAlgorithmLevenberg, BlockSolver Schur path, LinearSolverEigen (restated
 * as a dense upper LLT), SE3Quat / VertexSE3Expmap, EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ (+OnlyPose),
 * RobustKernelHuber, Line3D, numeric central-difference linearizeOplus).
 *
 * It is deliberately structured like g2o (graph of vertices and edges, per-edge computeError /
 * linearizeOplus / constructQuadraticForm in insertion order, block Schur, back-substitution) so
 * that (a) summation orders follow the reference and (b) timing it is a fair "g2o-equivalent CPU
 * restatement" baseline. It shares no code with the CUDA product.
 */
#include "oracle.h"

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include <atomic>
#include <thread>

namespace {

// ----------------------------------------------------------------------------------------------
// small fixed-size helpers
// ----------------------------------------------------------------------------------------------
struct V3 {
  double v[3];
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
};
inline V3 mk(double a, double b, double c) { return V3{{a, b, c}}; }
inline V3 operator+(const V3& a, const V3& b) { return mk(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3& a, const V3& b) { return mk(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator*(double s, const V3& a) { return mk(s * a[0], s * a[1], s * a[2]); }
inline double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline V3 cross(const V3& a, const V3& b) {
  return mk(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }

struct M3 {
  double m[3][3];
};
inline M3 m3_identity() { return M3{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline M3 m3_mul(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return c;
}
inline V3 m3_mulv(const M3& a, const V3& x) {
  return mk(a.m[0][0] * x[0] + a.m[0][1] * x[1] + a.m[0][2] * x[2],
            a.m[1][0] * x[0] + a.m[1][1] * x[1] + a.m[1][2] * x[2],
            a.m[2][0] * x[0] + a.m[2][1] * x[1] + a.m[2][2] * x[2]);
}
inline M3 skew(const V3& t) { return M3{{{0, -t[2], t[1]}, {t[2], 0, -t[0]}, {-t[1], t[0], 0}}}; }

// ----------------------------------------------------------------------------------------------
// Eigen::Quaterniond semantics used by g2o::SE3Quat
// ----------------------------------------------------------------------------------------------
struct Quat {
  double x, y, z, w;
};
inline Quat q_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline void q_normalize(Quat& q) {
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n;
  q.y /= n;
  q.z /= n;
  q.w /= n;
}
inline Quat q_conj(const Quat& q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
// Eigen QuaternionBase::_transformVector
inline V3 q_rot(const Quat& q, const V3& v) {
  V3 qv = mk(q.x, q.y, q.z);
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}
// Eigen QuaternionBase::toRotationMatrix
inline M3 q_to_R(const Quat& q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 R;
  R.m[0][0] = 1 - (tyy + tzz);
  R.m[0][1] = txy - twz;
  R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;
  R.m[1][1] = 1 - (txx + tzz);
  R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;
  R.m[2][1] = tyz + twx;
  R.m[2][2] = 1 - (txx + tyy);
  return R;
}
// Eigen quaternion-from-rotation-matrix (internal::quaternionbase_assign_impl<.,3,3>)
inline Quat q_from_R(const M3& R) {
  Quat q;
  double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R.m[2][1] - R.m[1][2]) * t;
    q.y = (R.m[0][2] - R.m[2][0]) * t;
    q.z = (R.m[1][0] - R.m[0][1]) * t;
  } else {
    int i = 0;
    if (R.m[1][1] > R.m[0][0]) i = 1;
    if (R.m[2][2] > R.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
    double qv[3];
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (R.m[k][j] - R.m[j][k]) * t;
    qv[j] = (R.m[j][i] + R.m[i][j]) * t;
    qv[k] = (R.m[k][i] + R.m[i][k]) * t;
    q.x = qv[0];
    q.y = qv[1];
    q.z = qv[2];
  }
  return q;
}

// ----------------------------------------------------------------------------------------------
// g2o::SE3Quat
// ----------------------------------------------------------------------------------------------
struct SE3 {
  Quat r;
  V3 t;
};
inline void se3_normalize_rotation(SE3& T) {
  if (T.r.w < 0) {
    T.r.x *= -1;
    T.r.y *= -1;
    T.r.z *= -1;
    T.r.w *= -1;
  }
  q_normalize(T.r);
}
inline SE3 se3_make(const Quat& q, const V3& t) {
  SE3 T{q, t};
  se3_normalize_rotation(T);
  return T;
}
inline SE3 se3_inverse(const SE3& T) {
  SE3 r;
  r.r = q_conj(T.r);
  r.t = q_rot(r.r, (-1.0) * T.t);
  return r;
}
inline SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r = a;
  r.t = r.t + q_rot(a.r, b.t);
  r.r = q_mul(a.r, b.r);
  se3_normalize_rotation(r);
  return r;
}
inline V3 se3_map(const SE3& T, const V3& x) { return q_rot(T.r, x) + T.t; }
inline SE3 se3_exp(const double* u) {
  V3 omega = mk(u[0], u[1], u[2]);
  V3 upsilon = mk(u[3], u[4], u[5]);
  double theta = norm(omega);
  M3 Omega = skew(omega);
  M3 Omega2 = m3_mul(Omega, Omega);
  M3 R, V;
  M3 I = m3_identity();
  if (theta < 0.00001) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        R.m[i][j] = I.m[i][j] + Omega.m[i][j] + 0.5 * Omega2.m[i][j];
        V.m[i][j] = I.m[i][j] + 0.5 * Omega.m[i][j] + (1.0 / 6.0) * Omega2.m[i][j];
      }
  } else {
    const double a = std::sin(theta) / theta;
    const double b = (1 - std::cos(theta)) / (theta * theta);
    const double c = (theta - std::sin(theta)) / std::pow(theta, 3);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        R.m[i][j] = I.m[i][j] + a * Omega.m[i][j] + b * Omega2.m[i][j];
        V.m[i][j] = I.m[i][j] + b * Omega.m[i][j] + c * Omega2.m[i][j];
      }
  }
  return se3_make(q_from_R(R), m3_mulv(V, upsilon));
}

// ----------------------------------------------------------------------------------------------
// g2o::Line3D (Pluecker [w, d])
// ----------------------------------------------------------------------------------------------
struct Line {
  double l[6];
  V3 w() const { return mk(l[0], l[1], l[2]); }
  V3 d() const { return mk(l[3], l[4], l[5]); }
};
inline void line_normalize(Line& L) {
  double n = 1.0 / norm(L.d());
  for (int i = 0; i < 6; ++i) L.l[i] *= n;
}
inline Line line_from_cartesian(const double* c) {
  V3 p = mk(c[0], c[1], c[2]);
  V3 v = mk(c[3], c[4], c[5]);
  V3 d = (1.0 / norm(v)) * v;
  p = p - dot(d, p) * d;
  V3 w = cross(p, p + d);
  return Line{{w[0], w[1], w[2], d[0], d[1], d[2]}};
}
// Isometry3 * Line3D: A = [[R, [t]x R], [0, R]], v' = A v, column-sequential accumulation
inline Line line_transform(const M3& R, const V3& t, const Line& L) {
  M3 S = m3_mul(skew(t), R);
  Line o;
  for (int i = 0; i < 3; ++i) {
    double acc = R.m[i][0] * L.l[0];
    acc += R.m[i][1] * L.l[1];
    acc += R.m[i][2] * L.l[2];
    acc += S.m[i][0] * L.l[3];
    acc += S.m[i][1] * L.l[4];
    acc += S.m[i][2] * L.l[5];
    o.l[i] = acc;
    double dd = R.m[i][0] * L.l[3];
    dd += R.m[i][1] * L.l[4];
    dd += R.m[i][2] * L.l[5];
    o.l[3 + i] = dd;
  }
  return o;
}
inline Line line_oplus(const Line& L, const double* v) {
  // toOrthonormal
  const V3 w = L.w(), d = L.d();
  const double mx = norm(d), my = norm(w);
  const double wn = 1.0 / std::sqrt(mx * mx + my * my);
  double W[2][2] = {{my * wn, -mx * wn}, {mx * wn, my * wn}};
  const double mn = 1.0 / my, dn = 1.0 / mx;
  V3 mdc = cross(w, d);
  const double mdcn = 1.0 / norm(mdc);
  M3 U;
  for (int i = 0; i < 3; ++i) {
    U.m[i][0] = w[i] * mn;
    U.m[i][1] = d[i] * dn;
    U.m[i][2] = mdc[i] * mdcn;
  }
  // update
  const double c = std::cos(v[3]), s = std::sin(v[3]);
  double Wu[2][2] = {{c, -s}, {s, c}};
  Quat q{v[0], v[1], v[2], std::sqrt(1 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))};
  q_normalize(q);
  M3 Uu = q_to_R(q);
  M3 Un = m3_mul(U, Uu);
  double Wn[2][2];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) Wn[i][j] = W[i][0] * Wu[0][j] + W[i][1] * Wu[1][j];
  // fromOrthonormal (normalises) + normalize again (Line3D::oplus)
  Line o;
  for (int i = 0; i < 3; ++i) {
    o.l[i] = Un.m[i][0] * Wn[0][0];
    o.l[3 + i] = Un.m[i][1] * Wn[1][0];
  }
  line_normalize(o);
  line_normalize(o);
  return o;
}

// ----------------------------------------------------------------------------------------------
// graph
// ----------------------------------------------------------------------------------------------
enum VKind { V_POSE = 0, V_POINT = 1, V_LINE = 2 };
enum EType { E_MONO_PT = 0, E_STEREO_PT = 1, E_MONO_LN = 2, E_STEREO_LN = 3, E_MONO_POSE = 4, E_STEREO_POSE = 5 };

struct Vertex {
  int kind;
  int id;
  bool fixed = false;
  bool marginalized = false;
  SE3 pose;      // V_POSE
  double est[6]; // V_POINT (3) / V_LINE (6)
  int dim() const { return kind == V_POSE ? 6 : (kind == V_POINT ? 3 : 4); }
  // optimisation state
  int hidx = -1;  // hessian index (position in index mapping)
  int col = -1;   // scalar column inside its (pose | landmark) partition
  bool active = false;
  double A[36];   // hessian diagonal block dim x dim row-major
  double b[6];
  // backup stack (LM push/pop + numeric Jacobian push/pop)
  std::vector<SE3> bk_pose;
  std::vector<std::array<double, 6>> bk_est;
  void push() {
    if (kind == V_POSE) bk_pose.push_back(pose);
    else {
      std::array<double, 6> a;
      std::memcpy(a.data(), est, sizeof(est));
      bk_est.push_back(a);
    }
  }
  void pop() {
    if (kind == V_POSE) {
      pose = bk_pose.back();
      bk_pose.pop_back();
    } else {
      std::memcpy(est, bk_est.back().data(), sizeof(est));
      bk_est.pop_back();
    }
  }
  void discard_top() {
    if (kind == V_POSE) bk_pose.pop_back();
    else bk_est.pop_back();
  }
  void oplus(const double* u) {
    if (kind == V_POSE) {
      pose = se3_mul(se3_exp(u), pose); // VertexSE3Expmap::oplusImpl
    } else if (kind == V_POINT) {
      for (int i = 0; i < 3; ++i) est[i] += u[i]; // VertexPointXYZ
    } else {
      Line L;
      std::memcpy(L.l, est, sizeof(L.l));
      L = line_oplus(L, u); // VertexLine3D::oplusImpl (vertex_line3d.h:26-29)
      std::memcpy(est, L.l, sizeof(L.l));
    }
  }
};

struct Edge {
  int type;
  int v_lm = -1;  // vertex index of landmark (binary edges; g2o vertex 0)
  int v_pose = -1;
  double meas[8];
  double Xw[3]; // pose-only edges
  double fx, fy, cx, cy, bf;
  double b;     // stereo line: bf / fx (g2o_optimization.cc:165)
  double Kv[3]; // line edges (:143,:166)
  double info;  // information = info * I
  int dim;
  double err[4];
  int level = 0;
  bool robust = true;
  double delta = 0; // Huber delta (float-rounded sqrt(thr), :77-78, :125-126)
  double Jl[16];    // dim x ld
  double Jp[24];    // dim x 6
  double chi2() const {
    double s = 0;
    for (int i = 0; i < dim; ++i) s += err[i] * info * err[i];
    return s;
  }
};

inline void huber(double e, double delta, double* rho) { // RobustKernelHuber::robustify
  const double dsqr = delta * delta;
  if (e <= dsqr) {
    rho[0] = e;
    rho[1] = 1.;
    rho[2] = 0.;
  } else {
    const double sqrte = std::sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

struct HplBlock {
  int pose_hidx;
  double B[24]; // 6 x ld row-major: pose rows, landmark cols
};

struct Graph {
  std::vector<Vertex> V;
  std::vector<Edge> E;
  bool bf_float = true;
  double num_delta = 1e-9;
  OrcStats* stats = nullptr;
  int cur_pass = 0;

  // active sets
  std::vector<int> act_edges;
  std::vector<int> ivmap; // hessian index -> vertex
  int n_pose_act = 0, n_lm_act = 0, size_poses = 0, size_lms = 0;
  bool do_schur = false;
  // solver storage
  std::vector<double> x, bvec;
  std::vector<std::vector<HplBlock>> hpl; // per active landmark (index hidx - n_pose_act)
  std::vector<double> Hschur, bschur, coeff;
  // LM state
  double lambda = 0, ni = 2;

  // -------------------------------------------------------------------------------------------
  // edges
  // -------------------------------------------------------------------------------------------
  void line_residual(const Edge& e, const SE3& T, const Line& L, double* err) const {
    // edge_project_line.cc:21-42 / edge_project_stereo_line.cc:22-51
    M3 R = q_to_R(T.r);
    Line Lc = line_transform(R, T.t, L);
    V3 w = Lc.w();
    double l0 = e.fy * w[0], l1 = e.fx * w[1], l2 = e.Kv[0] * w[0] + e.Kv[1] * w[1] + e.Kv[2] * w[2];
    double n = std::sqrt(l0 * l0 + l1 * l1);
    if (e.type == E_MONO_LN) {
      double e0 = e.meas[0] * l0 + e.meas[1] * l1 + l2;
      double e1 = e.meas[2] * l0 + e.meas[3] * l1 + l2;
      err[0] = e0 / n;
      err[1] = e1 / n;
      return;
    }
    err[0] = (e.meas[0] * l0 + e.meas[1] * l1 + l2) / n;
    err[1] = (e.meas[2] * l0 + e.meas[3] * l1 + l2) / n;
    V3 tr = T.t;
    tr[0] -= e.b; // T_right(0,3) -= b
    Line Lr = line_transform(R, tr, L);
    V3 wr = Lr.w();
    double r0 = e.fy * wr[0], r1 = e.fx * wr[1], r2 = e.Kv[0] * wr[0] + e.Kv[1] * wr[1] + e.Kv[2] * wr[2];
    double nr = std::sqrt(r0 * r0 + r1 * r1);
    err[2] = (e.meas[4] * r0 + e.meas[5] * r1 + r2) / nr;
    err[3] = (e.meas[6] * r0 + e.meas[7] * r1 + r2) / nr;
  }

  void compute_error(Edge& e) const {
    const Vertex& vp = V[e.v_pose];
    switch (e.type) {
      case E_MONO_PT:
      case E_MONO_POSE: {
        V3 X = (e.type == E_MONO_PT) ? mk(V[e.v_lm].est[0], V[e.v_lm].est[1], V[e.v_lm].est[2]) : mk(e.Xw[0], e.Xw[1], e.Xw[2]);
        V3 c = se3_map(vp.pose, X);
        // project2d + intrinsics (EdgeSE3ProjectXYZ::cam_project)
        double px = c[0] / c[2], py = c[1] / c[2];
        e.err[0] = e.meas[0] - (px * e.fx + e.cx);
        e.err[1] = e.meas[1] - (py * e.fy + e.cy);
        break;
      }
      case E_STEREO_PT:
      case E_STEREO_POSE: {
        V3 X = (e.type == E_STEREO_PT) ? mk(V[e.v_lm].est[0], V[e.v_lm].est[1], V[e.v_lm].est[2]) : mk(e.Xw[0], e.Xw[1], e.Xw[2]);
        V3 c = se3_map(vp.pose, X);
        const double invz = 1.0 / c[2];
        // EdgeStereoSE3ProjectXYZ::cam_project(xyz, const float& bf): bf rounded to float (binary edge only)
        double bf = e.bf;
        if (e.type == E_STEREO_PT && bf_float) bf = (double)(float)e.bf;
        double u = c[0] * invz * e.fx + e.cx;
        double v = c[1] * invz * e.fy + e.cy;
        double ur = u - bf * invz;
        e.err[0] = e.meas[0] - u;
        e.err[1] = e.meas[1] - v;
        e.err[2] = e.meas[2] - ur;
        break;
      }
      default: {
        Line L;
        std::memcpy(L.l, V[e.v_lm].est, sizeof(L.l));
        line_residual(e, vp.pose, L, e.err);
      }
    }
  }

  bool depth_positive(const Edge& e) const { // isDepthPositive
    const Vertex& vl = V[e.v_lm];
    V3 c = se3_map(V[e.v_pose].pose, mk(vl.est[0], vl.est[1], vl.est[2]));
    return c[2] > 0.0;
  }

  void linearize(Edge& e) {
    Vertex& vp = V[e.v_pose];
    if (e.type == E_MONO_PT || e.type == E_STEREO_PT) {
      // EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ::linearizeOplus
      const Vertex& vl = V[e.v_lm];
      V3 c = se3_map(vp.pose, mk(vl.est[0], vl.est[1], vl.est[2]));
      M3 R = q_to_R(vp.pose.r);
      const double x = c[0], y = c[1], z = c[2], z_2 = z * z;
      const double fx = e.fx, fy = e.fy, bf = e.bf;
      double* Ji = e.Jl; // dim x 3
      double* Jj = e.Jp; // dim x 6
      if (e.type == E_STEREO_PT) {
        for (int k = 0; k < 3; ++k) {
          Ji[0 * 3 + k] = -fx * R.m[0][k] / z + fx * x * R.m[2][k] / z_2;
          Ji[1 * 3 + k] = -fy * R.m[1][k] / z + fy * y * R.m[2][k] / z_2;
          Ji[2 * 3 + k] = Ji[0 * 3 + k] - bf * R.m[2][k] / z_2;
        }
      } else {
        // -1/z * tmp * R, tmp = [[fx,0,-x/z*fx],[0,fy,-y/z*fy]]
        double tmp[2][3] = {{fx, 0, -x / z * fx}, {0, fy, -y / z * fy}};
        for (int r = 0; r < 2; ++r)
          for (int k = 0; k < 3; ++k) {
            double s = (-1. / z * tmp[r][0]) * R.m[0][k] + (-1. / z * tmp[r][1]) * R.m[1][k] + (-1. / z * tmp[r][2]) * R.m[2][k];
            Ji[r * 3 + k] = s;
          }
      }
      Jj[0] = x * y / z_2 * fx;
      Jj[1] = -(1 + (x * x / z_2)) * fx;
      Jj[2] = y / z * fx;
      Jj[3] = -1. / z * fx;
      Jj[4] = 0;
      Jj[5] = x / z_2 * fx;
      Jj[6] = (1 + y * y / z_2) * fy;
      Jj[7] = -x * y / z_2 * fy;
      Jj[8] = -x / z * fy;
      Jj[9] = 0;
      Jj[10] = -1. / z * fy;
      Jj[11] = y / z_2 * fy;
      if (e.type == E_STEREO_PT) {
        Jj[12] = Jj[0] - bf * y / z_2;
        Jj[13] = Jj[1] + bf * x / z_2;
        Jj[14] = Jj[2];
        Jj[15] = Jj[3];
        Jj[16] = 0;
        Jj[17] = Jj[5] - bf / z_2;
      }
      return;
    }
    if (e.type == E_MONO_POSE || e.type == E_STEREO_POSE) {
      // Edge(Stereo)SE3ProjectXYZOnlyPose::linearizeOplus
      V3 c = se3_map(vp.pose, mk(e.Xw[0], e.Xw[1], e.Xw[2]));
      const double x = c[0], y = c[1];
      const double invz = 1.0 / c[2], invz_2 = invz * invz;
      const double fx = e.fx, fy = e.fy, bf = e.bf;
      double* J = e.Jp;
      J[0] = x * y * invz_2 * fx;
      J[1] = -(1 + (x * x * invz_2)) * fx;
      J[2] = y * invz * fx;
      J[3] = -invz * fx;
      J[4] = 0;
      J[5] = x * invz_2 * fx;
      J[6] = (1 + y * y * invz_2) * fy;
      J[7] = -x * y * invz_2 * fy;
      J[8] = -x * invz * fy;
      J[9] = 0;
      J[10] = -invz * fy;
      J[11] = y * invz_2 * fy;
      if (e.type == E_STEREO_POSE) {
        J[12] = J[0] - bf * y * invz_2;
        J[13] = J[1] + bf * x * invz_2;
        J[14] = J[2];
        J[15] = J[3];
        J[16] = 0;
        J[17] = J[5] - bf * invz_2;
      }
      return;
    }
    // line edges: BaseBinaryEdge numeric linearizeOplus, central differences delta = 1e-9 (§9.8)
    Vertex& vl = V[e.v_lm];
    const double delta = num_delta, scalar = 1 / (2 * delta);
    double err_before[4];
    std::memcpy(err_before, e.err, sizeof(err_before));
    double ep[4];
    if (!vl.fixed) {
      double add[4] = {0, 0, 0, 0};
      for (int d = 0; d < 4; ++d) {
        vl.push();
        add[d] = delta;
        vl.oplus(add);
        compute_error(e);
        std::memcpy(ep, e.err, sizeof(ep));
        vl.pop();
        vl.push();
        add[d] = -delta;
        vl.oplus(add);
        compute_error(e);
        for (int r = 0; r < e.dim; ++r) ep[r] -= e.err[r];
        vl.pop();
        add[d] = 0.0;
        for (int r = 0; r < e.dim; ++r) e.Jl[r * 4 + d] = scalar * ep[r];
      }
    }
    if (!vp.fixed) {
      double add[6] = {0, 0, 0, 0, 0, 0};
      for (int d = 0; d < 6; ++d) {
        vp.push();
        add[d] = delta;
        vp.oplus(add);
        compute_error(e);
        std::memcpy(ep, e.err, sizeof(ep));
        vp.pop();
        vp.push();
        add[d] = -delta;
        vp.oplus(add);
        compute_error(e);
        for (int r = 0; r < e.dim; ++r) ep[r] -= e.err[r];
        vp.pop();
        add[d] = 0.0;
        for (int r = 0; r < e.dim; ++r) e.Jp[r * 6 + d] = scalar * ep[r];
      }
    }
    std::memcpy(e.err, err_before, sizeof(err_before));
  }

  // BaseBinaryEdge / BaseUnaryEdge::constructQuadraticForm (§9.7)
  void construct_quadratic_form(Edge& e) {
    Vertex& vp = V[e.v_pose];
    const int D = e.dim;
    double w = 1.0;
    if (e.robust) {
      double rho[3];
      huber(e.chi2(), e.delta, rho);
      w = rho[1];
    }
    // omega_r = -(omega * err) [* rho1]; weightedOmega = rho1 * omega
    double omega_r[4];
    for (int r = 0; r < D; ++r) {
      omega_r[r] = -(e.info * e.err[r]);
      if (e.robust) omega_r[r] *= w;
    }
    const double wo = e.robust ? w * e.info : e.info;
    const bool pose_free = !vp.fixed;
    if (e.v_lm < 0) { // unary pose-only edge
      if (pose_free) {
        for (int i = 0; i < 6; ++i) {
          double s = 0;
          for (int r = 0; r < D; ++r) s += e.Jp[r * 6 + i] * omega_r[r];
          vp.b[i] += s;
          for (int j = 0; j < 6; ++j) {
            double a = 0;
            for (int r = 0; r < D; ++r) a += e.Jp[r * 6 + i] * wo * e.Jp[r * 6 + j];
            vp.A[i * 6 + j] += a;
          }
        }
      }
      return;
    }
    Vertex& vl = V[e.v_lm];
    const int ld = vl.dim();
    // landmark ("from", never fixed in the reference: Position3d::fixed is ignored, :53-59)
    if (!vl.fixed) {
      for (int i = 0; i < ld; ++i) {
        double s = 0;
        for (int r = 0; r < D; ++r) s += e.Jl[r * ld + i] * omega_r[r];
        vl.b[i] += s;
        for (int j = 0; j < ld; ++j) {
          double a = 0;
          for (int r = 0; r < D; ++r) a += e.Jl[r * ld + i] * wo * e.Jl[r * ld + j];
          vl.A[i * ld + j] += a;
        }
      }
      if (pose_free) {
        // Hpl block for (pose, landmark): B^T * wOmega * A
        std::vector<HplBlock>& col = hpl[vl.hidx - n_pose_act];
        HplBlock* blk = nullptr;
        for (auto& hb : col)
          if (hb.pose_hidx == vp.hidx) blk = &hb;
        if (!blk) {
          HplBlock nb;
          nb.pose_hidx = vp.hidx;
          std::memset(nb.B, 0, sizeof(nb.B));
          auto it = col.begin();
          while (it != col.end() && it->pose_hidx < vp.hidx) ++it;
          it = col.insert(it, nb);
          blk = &*it;
        }
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < ld; ++j) {
            double a = 0;
            for (int r = 0; r < D; ++r) a += e.Jp[r * 6 + i] * wo * e.Jl[r * ld + j];
            blk->B[i * ld + j] += a;
          }
      }
    }
    if (pose_free) {
      for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int r = 0; r < D; ++r) s += e.Jp[r * 6 + i] * omega_r[r];
        vp.b[i] += s;
        for (int j = 0; j < 6; ++j) {
          double a = 0;
          for (int r = 0; r < D; ++r) a += e.Jp[r * 6 + i] * wo * e.Jp[r * 6 + j];
          vp.A[i * 6 + j] += a;
        }
      }
    }
  }

  // -------------------------------------------------------------------------------------------
  // SparseOptimizer::initializeOptimization(level) + buildIndexMapping (§9.12)
  // -------------------------------------------------------------------------------------------
  bool initialize_optimization(int level) {
    act_edges.clear();
    for (auto& v : V) {
      v.active = false;
      v.hidx = -1;
    }
    for (size_t k = 0; k < E.size(); ++k) {
      Edge& e = E[k];
      if (e.level != level) continue;
      bool all_fixed = V[e.v_pose].fixed && (e.v_lm < 0 || V[e.v_lm].fixed);
      if (all_fixed) continue;
      act_edges.push_back((int)k);
      V[e.v_pose].active = true;
      if (e.v_lm >= 0) V[e.v_lm].active = true;
    }
    ivmap.clear();
    n_pose_act = n_lm_act = size_poses = size_lms = 0;
    // V is stored in ascending g2o-id order (poses, points, lines; g2o_optimization.cc:39-70)
    for (int k = 0; k < 2; ++k)
      for (size_t i = 0; i < V.size(); ++i) {
        Vertex& v = V[i];
        if (!v.active || v.fixed) continue;
        if ((int)v.marginalized != k) continue;
        v.hidx = (int)ivmap.size();
        ivmap.push_back((int)i);
        if (k == 0) {
          v.col = size_poses;
          n_pose_act++;
          size_poses += v.dim();
        } else {
          v.col = size_lms;
          n_lm_act++;
          size_lms += v.dim();
        }
      }
    return !ivmap.empty();
  }

  void compute_active_errors() {
    for (int k : act_edges) compute_error(E[k]);
    if (stats) stats->edges_evaluated += (int64_t)act_edges.size();
  }
  double active_robust_chi2() const {
    double chi = 0, rho[3];
    for (int k : act_edges) {
      const Edge& e = E[k];
      if (e.robust) {
        huber(e.chi2(), e.delta, rho);
        chi += rho[0];
      } else
        chi += e.chi2();
    }
    return chi;
  }

  void build_system() {
    for (int vi : ivmap) {
      std::memset(V[vi].A, 0, sizeof(V[vi].A));
      std::memset(V[vi].b, 0, sizeof(V[vi].b));
    }
    hpl.assign(n_lm_act, {});
    for (int k : act_edges) {
      linearize(E[k]);
      construct_quadratic_form(E[k]);
    }
    if (stats) stats->edges_linearized += (int64_t)act_edges.size();
    for (int vi : ivmap) {
      const Vertex& v = V[vi];
      int base = v.col + (v.marginalized ? size_poses : 0);
      for (int i = 0; i < v.dim(); ++i) bvec[base + i] = v.b[i];
    }
  }

  // general inverse of a small matrix: Eigen dynamic-size inverse() = PartialPivLU (BlockSolverX)
  static void small_inverse(const double* A, int n, double* inv) {
    double M[4][8];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        M[i][j] = A[i * n + j];
        M[i][n + j] = (i == j) ? 1.0 : 0.0;
      }
    for (int c = 0; c < n; ++c) {
      int p = c;
      double best = std::fabs(M[c][c]);
      for (int r = c + 1; r < n; ++r)
        if (std::fabs(M[r][c]) > best) {
          best = std::fabs(M[r][c]);
          p = r;
        }
      if (p != c)
        for (int j = 0; j < 2 * n; ++j) std::swap(M[c][j], M[p][j]);
      for (int r = c + 1; r < n; ++r) {
        double f = M[r][c] / M[c][c];
        for (int j = c; j < 2 * n; ++j) M[r][j] -= f * M[c][j];
      }
    }
    for (int c = n - 1; c >= 0; --c) {
      for (int j = n; j < 2 * n; ++j) {
        double s = M[c][j];
        for (int k = c + 1; k < n; ++k) s -= M[c][k] * M[k][j];
        M[c][j] = s / M[c][c];
      }
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) inv[i * n + j] = M[i][n + j];
  }

  // dense upper LLT solve (stands in for Eigen SimplicialLLT<.,Upper>; fails iff a pivot <= 0, §9.11)
  static bool llt_solve_upper(std::vector<double>& A, int n, const double* b, double* xo) {
    // A = U^T U using the upper triangle of A (row-major); overwrite upper with U
    for (int k = 0; k < n; ++k) {
      double d = A[k * n + k];
      for (int p = 0; p < k; ++p) d -= A[p * n + k] * A[p * n + k];
      if (d <= 0.0) return false;
      double ukk = std::sqrt(d);
      A[k * n + k] = ukk;
      for (int j = k + 1; j < n; ++j) {
        double s = A[k * n + j];
        for (int p = 0; p < k; ++p) s -= A[p * n + k] * A[p * n + j];
        A[k * n + j] = s / ukk;
      }
    }
    static thread_local std::vector<double> y;
    y.resize(n);
    for (int i = 0; i < n; ++i) { // U^T y = b
      double s = b[i];
      for (int p = 0; p < i; ++p) s -= A[p * n + i] * y[p];
      y[i] = s / A[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) { // U x = y
      double s = y[i];
      for (int p = i + 1; p < n; ++p) s -= A[i * n + p] * xo[p];
      xo[i] = s / A[i * n + i];
    }
    return true;
  }

  // BlockSolver::solve with lambda already added to the diagonals (§9.10)
  bool solver_solve() {
    const int n = size_poses;
    if (!do_schur) {
      std::vector<double>& H = Hschur;
      H.assign((size_t)n * n, 0.0);
      for (int k = 0; k < n_pose_act; ++k) {
        const Vertex& v = V[ivmap[k]];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) H[(size_t)(v.col + i) * n + v.col + j] = v.A[i * 6 + j];
      }
      return llt_solve_upper(H, n, bvec.data(), x.data());
    }
    Hschur.assign((size_t)n * n, 0.0);
    for (int k = 0; k < n_pose_act; ++k) {
      const Vertex& v = V[ivmap[k]];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) Hschur[(size_t)(v.col + i) * n + v.col + j] = v.A[i * 6 + j];
    }
    coeff.assign(n + size_lms, 0.0);
    std::vector<double> dinv_all((size_t)n_lm_act * 16);
    for (int li = 0; li < n_lm_act; ++li) {
      const Vertex& vl = V[ivmap[n_pose_act + li]];
      const int ld = vl.dim();
      double* Dinv = &dinv_all[(size_t)li * 16];
      small_inverse(vl.A, ld, Dinv);
      double db[4];
      for (int i = 0; i < ld; ++i) {
        double s = 0;
        for (int j = 0; j < ld; ++j) s += Dinv[i * ld + j] * bvec[n + vl.col + j];
        db[i] = s;
      }
      const std::vector<HplBlock>& col = hpl[li];
      for (size_t a = 0; a < col.size(); ++a) {
        const HplBlock& Bi = col[a];
        const int ci = V[ivmap[Bi.pose_hidx]].col;
        double BDinv[24];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < ld; ++j) {
            double s = 0;
            for (int k = 0; k < ld; ++k) s += Bi.B[i * ld + k] * Dinv[k * ld + j];
            BDinv[i * ld + j] = s;
          }
        for (int i = 0; i < 6; ++i) {
          double s = 0;
          for (int k = 0; k < ld; ++k) s += Bi.B[i * ld + k] * db[k];
          coeff[ci + i] += s;
        }
        for (size_t c = a; c < col.size(); ++c) { // upper triangle: i2 >= i1
          const HplBlock& Bj = col[c];
          const int cj = V[ivmap[Bj.pose_hidx]].col;
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
              double s = 0;
              for (int k = 0; k < ld; ++k) s += BDinv[i * ld + k] * Bj.B[j * ld + k];
              Hschur[(size_t)(ci + i) * n + cj + j] -= s;
            }
        }
      }
    }
    bschur.resize(n);
    for (int i = 0; i < n; ++i) bschur[i] = bvec[i] - coeff[i];
    bool ok = llt_solve_upper(Hschur, n, bschur.data(), x.data());
    if (!ok) return false;
    // landmarks: xl = Dinv * (bl - Hpl^T xp)
    for (int li = 0; li < n_lm_act; ++li) {
      const Vertex& vl = V[ivmap[n_pose_act + li]];
      const int ld = vl.dim();
      double cl[4];
      for (int j = 0; j < ld; ++j) cl[j] = bvec[n + vl.col + j];
      for (const HplBlock& Bi : hpl[li]) {
        const int ci = V[ivmap[Bi.pose_hidx]].col;
        for (int j = 0; j < ld; ++j) {
          double s = 0;
          for (int i = 0; i < 6; ++i) s += Bi.B[i * ld + j] * (-x[ci + i]);
          cl[j] += s;
        }
      }
      const double* Dinv = &dinv_all[(size_t)li * 16];
      for (int i = 0; i < ld; ++i) {
        double s = 0;
        for (int j = 0; j < ld; ++j) s += Dinv[i * ld + j] * cl[j];
        x[n + vl.col + i] = s;
      }
    }
    return true;
  }

  void set_lambda(double lam, std::vector<double>& diag_backup) {
    diag_backup.clear();
    for (int vi : ivmap) {
      Vertex& v = V[vi];
      const int d = v.dim();
      for (int i = 0; i < d; ++i) {
        diag_backup.push_back(v.A[i * d + i]);
        v.A[i * d + i] += lam;
      }
    }
  }
  void restore_diagonal(const std::vector<double>& diag_backup) {
    size_t k = 0;
    for (int vi : ivmap) {
      Vertex& v = V[vi];
      const int d = v.dim();
      for (int i = 0; i < d; ++i) v.A[i * d + i] = diag_backup[k++];
    }
  }

  void trace(int iter, int trial, int accepted, double c0, double c1, double lam, double rho) {
    if (!stats || !stats->trace || stats->n_trace >= stats->trace_cap) return;
    OrcTraceRow& r = stats->trace[stats->n_trace++];
    r.pass = cur_pass;
    r.iter = iter;
    r.trial = trial;
    r.accepted = accepted;
    r.chi_before = c0;
    r.chi_after = c1;
    r.lambda = lam;
    r.rho = rho;
  }

  // OptimizationAlgorithmLevenberg::solve (§9.9). Returns true for OK, false for Terminate.
  bool lm_solve(int iteration, double& final_chi) {
    compute_active_errors();
    double currentChi = active_robust_chi2();
    double tempChi = currentChi;
    build_system();
    if (iteration == 0) {
      double maxDiagonal = 0;
      for (int vi : ivmap) {
        const Vertex& v = V[vi];
        const int d = v.dim();
        for (int j = 0; j < d; ++j) maxDiagonal = std::max(std::fabs(v.A[j * d + j]), maxDiagonal);
      }
      lambda = 1e-5 * maxDiagonal;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    std::vector<double> diag_backup;
    const int nx = size_poses + size_lms;
    do {
      for (int vi : ivmap) V[vi].push();
      // g2o pushes all *active* vertices; fixed ones never change so skipping them is equivalent
      const double lam_used = lambda;
      set_lambda(lambda, diag_backup);
      bool ok2 = solver_solve();
      {
        const double* u = x.data();
        for (int vi : ivmap) {
          V[vi].oplus(u);
          u += V[vi].dim();
        }
      }
      restore_diagonal(diag_backup);
      compute_active_errors();
      tempChi = active_robust_chi2();
      if (!ok2) tempChi = DBL_MAX;
      rho = (currentChi - tempChi);
      double scale = 0;
      for (int j = 0; j < nx; ++j) scale += x[j] * (lambda * x[j] + bvec[j]);
      scale += 1e-3;
      rho /= scale;
      int accepted = 0;
      const double chi_before = currentChi;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        double scaleFactor = std::max(1. / 3., alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        for (int vi : ivmap) V[vi].discard_top();
        accepted = 1;
      } else {
        lambda *= ni;
        ni *= 2;
        for (int vi : ivmap) V[vi].pop();
        if (!std::isfinite(lambda)) {
          trace(iteration, qmax, 0, chi_before, tempChi, lam_used, rho);
          if (stats) stats->trials[cur_pass & 3]++;
          break;
        }
      }
      trace(iteration, qmax, accepted, chi_before, tempChi, lam_used, rho);
      if (stats) stats->trials[cur_pass & 3]++;
      qmax++;
    } while (rho < 0 && qmax < 10);
    final_chi = currentChi;
    if (qmax == 10 || rho == 0 || !std::isfinite(lambda)) return false;
    return true;
  }

  // SparseOptimizer::optimize(iterations)
  int optimize(int iterations, double* final_chi = nullptr) {
    if (ivmap.empty()) return -1;
    // OptimizationAlgorithmWithHessian::init: Schur iff an active vertex is marginalized
    do_schur = false;
    for (int vi : ivmap)
      if (V[vi].marginalized) do_schur = true;
    x.assign(size_poses + size_lms, 0.0);
    bvec.assign(size_poses + size_lms, 0.0);
    int done = 0;
    bool ok = true;
    double chi = 0;
    for (int i = 0; i < iterations && ok; ++i) {
      ok = lm_solve(i, chi);
      ++done;
      if (stats) stats->iters[cur_pass & 3]++;
    }
    if (final_chi) *final_chi = chi;
    return done;
  }
};

void stats_reset(OrcStats* s) {
  if (!s) return;
  std::memset(s->iters, 0, sizeof(s->iters));
  std::memset(s->trials, 0, sizeof(s->trials));
  s->edges_linearized = 0;
  s->edges_evaluated = 0;
  s->final_chi2 = 0;
  s->n_trace = 0;
}

void set_cam(Edge& e, const double* c) {
  e.fx = c[0];
  e.fy = c[1];
  e.cx = c[2];
  e.cy = c[3];
  e.bf = c[4];
}

} // namespace

// ================================================================================================
// LocalmapOptimization — g2o_optimization.cc:21-252
// ================================================================================================
// One reusable graph per worker thread: vectors keep their capacity across problems, so a batch
// does not hammer mmap/munmap (large allocations serialise threads on the process mm lock).
static Graph& tls_graph() {
  static thread_local Graph g;
  g.V.clear();
  g.E.clear();
  g.act_edges.clear();
  g.ivmap.clear();
  g.stats = nullptr;
  g.cur_pass = 0;
  g.lambda = 0;
  g.ni = 2;
  g.num_delta = 1e-9;
  return g;
}

extern "C" int orc_local_ba(OrcLocalProblem* P, const OrcConfig* cfg, OrcStats* stats) {
  stats_reset(stats);
  Graph& G = tls_graph();
  G.stats = stats;
  G.bf_float = cfg->stereo_bf_float != 0;
  G.num_delta = cfg->numeric_delta > 0 ? cfg->numeric_delta : 1e-9;
  std::map<int, int> pose_of, point_of, line_of;
  // frame vertices (:39-48); ids must be ascending (std::map order)
  for (int i = 0; i < P->n_poses; ++i) {
    if (i && P->pose_id[i] <= P->pose_id[i - 1]) return -1;
    Vertex v;
    v.kind = V_POSE;
    v.id = P->pose_id[i];
    v.fixed = P->pose_fixed[i] != 0;
    const double* p = &P->pose_p[3 * i];
    const double* q = &P->pose_q[4 * i];
    v.pose = se3_inverse(se3_make(Quat{q[0], q[1], q[2], q[3]}, mk(p[0], p[1], p[2]))); // :42
    pose_of[v.id] = (int)G.V.size();
    G.V.push_back(v);
  }
  for (int i = 0; i < P->n_points; ++i) { // :51-61
    if (i && P->point_id[i] <= P->point_id[i - 1]) return -1;
    Vertex v;
    v.kind = V_POINT;
    v.id = P->point_id[i];
    v.marginalized = true;
    std::memset(v.est, 0, sizeof(v.est));
    for (int k = 0; k < 3; ++k) v.est[k] = P->point_p[3 * i + k];
    point_of[v.id] = (int)G.V.size();
    G.V.push_back(v);
  }
  for (int i = 0; i < P->n_lines; ++i) { // :64-70
    if (i && P->line_id[i] <= P->line_id[i - 1]) return -1;
    Vertex v;
    v.kind = V_LINE;
    v.id = P->line_id[i];
    v.marginalized = true;
    for (int k = 0; k < 6; ++k) v.est[k] = P->line_L[6 * i + k];
    line_of[v.id] = (int)G.V.size();
    G.V.push_back(v);
  }
  const float thHuberMonoPoint = std::sqrt(cfg->mono_point);     // :77
  const float thHuberStereoPoint = std::sqrt(cfg->stereo_point); // :78
  const float thHuberMonoLine = std::sqrt(cfg->mono_line);       // :125
  const float thHuberStereoLine = std::sqrt(cfg->stereo_line);   // :126
  const size_t e_mono0 = G.E.size();
  for (int i = 0; i < P->n_mono_pt; ++i) { // :81-97
    Edge e;
    e.type = E_MONO_PT;
    auto pl = point_of.find(P->mp_id_point[i]);
    auto pp = pose_of.find(P->mp_id_pose[i]);
    if (pl == point_of.end() || pp == pose_of.end() || P->mp_id_cam[i] < 0 || P->mp_id_cam[i] >= P->n_cams) return -2;
    e.v_lm = pl->second;
    e.v_pose = pp->second;
    e.meas[0] = P->mp_kp[2 * i];
    e.meas[1] = P->mp_kp[2 * i + 1];
    e.info = 1.0;
    e.dim = 2;
    e.delta = thHuberMonoPoint;
    set_cam(e, &P->cams[5 * P->mp_id_cam[i]]);
    G.E.push_back(e);
  }
  const size_t e_stereo0 = G.E.size();
  for (int i = 0; i < P->n_stereo_pt; ++i) { // :100-118
    Edge e;
    e.type = E_STEREO_PT;
    auto pl = point_of.find(P->sp_id_point[i]);
    auto pp = pose_of.find(P->sp_id_pose[i]);
    if (pl == point_of.end() || pp == pose_of.end() || P->sp_id_cam[i] < 0 || P->sp_id_cam[i] >= P->n_cams) return -2;
    e.v_lm = pl->second;
    e.v_pose = pp->second;
    for (int k = 0; k < 3; ++k) e.meas[k] = P->sp_kp[3 * i + k];
    e.info = 1.0;
    e.dim = 3;
    e.delta = thHuberStereoPoint;
    set_cam(e, &P->cams[5 * P->sp_id_cam[i]]);
    G.E.push_back(e);
  }
  const size_t e_mline0 = G.E.size();
  for (int i = 0; i < P->n_mono_ln; ++i) { // :128-146
    Edge e;
    e.type = E_MONO_LN;
    auto pl = line_of.find(P->ml_id_line[i]);
    auto pp = pose_of.find(P->ml_id_pose[i]);
    if (pl == line_of.end() || pp == pose_of.end() || P->ml_id_cam[i] < 0 || P->ml_id_cam[i] >= P->n_cams) return -2;
    e.v_lm = pl->second;
    e.v_pose = pp->second;
    for (int k = 0; k < 4; ++k) e.meas[k] = P->ml_l2d[4 * i + k];
    e.info = 0.1;
    e.dim = 2;
    e.delta = thHuberMonoLine;
    set_cam(e, &P->cams[5 * P->ml_id_cam[i]]);
    e.Kv[0] = -e.fy * e.cx;
    e.Kv[1] = -e.fx * e.cy;
    e.Kv[2] = e.fx * e.fy;
    e.b = 0;
    G.E.push_back(e);
  }
  const size_t e_sline0 = G.E.size();
  for (int i = 0; i < P->n_stereo_ln; ++i) { // :149-169
    Edge e;
    e.type = E_STEREO_LN;
    auto pl = line_of.find(P->sl_id_line[i]);
    auto pp = pose_of.find(P->sl_id_pose[i]);
    if (pl == line_of.end() || pp == pose_of.end() || P->sl_id_cam[i] < 0 || P->sl_id_cam[i] >= P->n_cams) return -2;
    e.v_lm = pl->second;
    e.v_pose = pp->second;
    for (int k = 0; k < 8; ++k) e.meas[k] = P->sl_l2d[8 * i + k];
    e.info = 0.1;
    e.dim = 4;
    e.delta = thHuberStereoLine;
    set_cam(e, &P->cams[5 * P->sl_id_cam[i]]);
    e.b = e.bf / e.fx;
    e.Kv[0] = -e.fy * e.cx;
    e.Kv[1] = -e.fx * e.cy;
    e.Kv[2] = e.fx * e.fy;
    G.E.push_back(e);
  }
  // errors default to zero like a freshly constructed g2o edge
  for (auto& e : G.E) std::memset(e.err, 0, sizeof(e.err));

  // solve (:172-173)
  G.cur_pass = 0;
  double chi = 0;
  G.initialize_optimization(0);
  G.optimize(cfg->iters_pass1, &chi);

  // check inlier observations (:176-206)
  const double thr[4] = {cfg->mono_point, cfg->stereo_point, cfg->mono_line, cfg->stereo_line};
  for (auto& e : G.E) {
    const bool is_point = (e.type == E_MONO_PT || e.type == E_STEREO_PT);
    if (e.chi2() > thr[e.type] || (is_point && !G.depth_positive(e))) e.level = 1;
    e.robust = false;
  }
  // optimize again without the outliers (:209-210)
  G.cur_pass = 1;
  if (G.initialize_optimization(0)) G.optimize(cfg->iters_pass2, &chi);
  if (stats) stats->final_chi2 = chi;

  // final flags (:213-231)
  for (int i = 0; i < P->n_mono_pt; ++i) {
    const Edge& e = G.E[e_mono0 + i];
    P->mp_inlier[i] = (e.chi2() <= cfg->mono_point && G.depth_positive(e)) ? 1 : 0;
  }
  for (int i = 0; i < P->n_stereo_pt; ++i) {
    const Edge& e = G.E[e_stereo0 + i];
    P->sp_inlier[i] = (e.chi2() <= cfg->stereo_point && G.depth_positive(e)) ? 1 : 0;
  }
  for (int i = 0; i < P->n_mono_ln; ++i) P->ml_inlier[i] = (G.E[e_mline0 + i].chi2() <= cfg->mono_line) ? 1 : 0;
  for (int i = 0; i < P->n_stereo_ln; ++i) P->sl_inlier[i] = (G.E[e_sline0 + i].chi2() <= cfg->stereo_line) ? 1 : 0;

  // recover optimized data (:235-251)
  for (int i = 0; i < P->n_poses; ++i) {
    SE3 Twc = se3_inverse(G.V[pose_of[P->pose_id[i]]].pose);
    for (int k = 0; k < 3; ++k) P->pose_p[3 * i + k] = Twc.t[k];
    P->pose_q[4 * i + 0] = Twc.r.x;
    P->pose_q[4 * i + 1] = Twc.r.y;
    P->pose_q[4 * i + 2] = Twc.r.z;
    P->pose_q[4 * i + 3] = Twc.r.w;
  }
  for (int i = 0; i < P->n_points; ++i)
    for (int k = 0; k < 3; ++k) P->point_p[3 * i + k] = G.V[point_of[P->point_id[i]]].est[k];
  for (int i = 0; i < P->n_lines; ++i)
    for (int k = 0; k < 6; ++k) P->line_L[6 * i + k] = G.V[line_of[P->line_id[i]]].est[k];
  return 0;
}

// ================================================================================================
// FrameOptimization — g2o_optimization.cc:256-397
// ================================================================================================
extern "C" int orc_frame_opt(OrcFrameProblem* P, const OrcConfig* cfg, OrcStats* stats) {
  stats_reset(stats);
  Graph& G = tls_graph();
  G.stats = stats;
  G.bf_float = cfg->stereo_bf_float != 0;
  G.num_delta = cfg->numeric_delta > 0 ? cfg->numeric_delta : 1e-9;
  std::map<int, int> point_of;
  for (int i = 0; i < P->n_points; ++i) point_of[P->point_id[i]] = i;
  Vertex v;
  v.kind = V_POSE;
  v.id = 0;
  const SE3 T_init = se3_inverse(se3_make(Quat{P->pose_q[0], P->pose_q[1], P->pose_q[2], P->pose_q[3]},
                                          mk(P->pose_p[0], P->pose_p[1], P->pose_p[2]))); // :271
  v.pose = T_init;
  G.V.push_back(v);
  const float deltaMonoPoint = std::sqrt(cfg->mono_point);     // :282
  const float deltaStereoPoint = std::sqrt(cfg->stereo_point); // :283
  for (int i = 0; i < P->n_mono_pt; ++i) { // :288-309
    Edge e;
    e.type = E_MONO_POSE;
    e.v_pose = 0;
    auto it = point_of.find(P->mp_id_point[i]);
    if (it == point_of.end() || P->mp_id_cam[i] < 0 || P->mp_id_cam[i] >= P->n_cams) return -2;
    for (int k = 0; k < 3; ++k) e.Xw[k] = P->point_p[3 * it->second + k];
    e.meas[0] = P->mp_kp[2 * i];
    e.meas[1] = P->mp_kp[2 * i + 1];
    e.info = 1.0;
    e.dim = 2;
    e.delta = deltaMonoPoint;
    set_cam(e, &P->cams[5 * P->mp_id_cam[i]]);
    std::memset(e.err, 0, sizeof(e.err));
    G.E.push_back(e);
  }
  const size_t e_stereo0 = G.E.size();
  for (int i = 0; i < P->n_stereo_pt; ++i) { // :313-333
    Edge e;
    e.type = E_STEREO_POSE;
    e.v_pose = 0;
    auto it = point_of.find(P->sp_id_point[i]);
    if (it == point_of.end() || P->sp_id_cam[i] < 0 || P->sp_id_cam[i] >= P->n_cams) return -2;
    for (int k = 0; k < 3; ++k) e.Xw[k] = P->point_p[3 * it->second + k];
    for (int k = 0; k < 3; ++k) e.meas[k] = P->sp_kp[3 * i + k];
    e.info = 1.0;
    e.dim = 3;
    e.delta = deltaStereoPoint;
    set_cam(e, &P->cams[5 * P->sp_id_cam[i]]);
    std::memset(e.err, 0, sizeof(e.err));
    G.E.push_back(e);
  }
  // ---- extension (see oracle.h): edges to fixed line vertices, set up like :128-169
  std::map<int, int> line_of;
  for (int i = 0; i < P->n_lines; ++i) {
    Vertex vl;
    vl.kind = V_LINE;
    vl.id = 1 + i;
    vl.fixed = true;
    vl.marginalized = true;
    for (int k = 0; k < 6; ++k) vl.est[k] = P->line_L[6 * i + k];
    line_of[P->line_id[i]] = (int)G.V.size();
    G.V.push_back(vl);
  }
  const float deltaMonoLine = std::sqrt(cfg->mono_line);     // :284
  const float deltaStereoLine = std::sqrt(cfg->stereo_line); // :285
  const size_t e_mline0 = G.E.size();
  for (int i = 0; i < P->n_mono_ln; ++i) {
    Edge e;
    e.type = E_MONO_LN;
    auto pl = line_of.find(P->ml_id_line[i]);
    if (pl == line_of.end() || P->ml_id_cam[i] < 0 || P->ml_id_cam[i] >= P->n_cams) return -2;
    e.v_lm = pl->second;
    e.v_pose = 0;
    for (int k = 0; k < 4; ++k) e.meas[k] = P->ml_l2d[4 * i + k];
    e.info = 0.1;
    e.dim = 2;
    e.delta = deltaMonoLine;
    set_cam(e, &P->cams[5 * P->ml_id_cam[i]]);
    e.Kv[0] = -e.fy * e.cx;
    e.Kv[1] = -e.fx * e.cy;
    e.Kv[2] = e.fx * e.fy;
    e.b = 0;
    std::memset(e.err, 0, sizeof(e.err));
    G.E.push_back(e);
  }
  const size_t e_sline0 = G.E.size();
  for (int i = 0; i < P->n_stereo_ln; ++i) {
    Edge e;
    e.type = E_STEREO_LN;
    auto pl = line_of.find(P->sl_id_line[i]);
    if (pl == line_of.end() || P->sl_id_cam[i] < 0 || P->sl_id_cam[i] >= P->n_cams) return -2;
    e.v_lm = pl->second;
    e.v_pose = 0;
    for (int k = 0; k < 8; ++k) e.meas[k] = P->sl_l2d[8 * i + k];
    e.info = 0.1;
    e.dim = 4;
    e.delta = deltaStereoLine;
    set_cam(e, &P->cams[5 * P->sl_id_cam[i]]);
    e.Kv[0] = -e.fy * e.cx;
    e.Kv[1] = -e.fx * e.cy;
    e.Kv[2] = e.fx * e.fy;
    e.b = e.bf / e.fx;
    std::memset(e.err, 0, sizeof(e.err));
    G.E.push_back(e);
  }
  int num_outlier = 0;
  double chi = 0;
  for (int iter = 0; iter < cfg->rounds; ++iter) { // :339
    G.cur_pass = iter;
    G.V[0].pose = T_init; // :340
    if (G.initialize_optimization(0)) G.optimize(cfg->iters_round, &chi);
    num_outlier = 0;
    for (int i = 0; i < P->n_mono_pt; ++i) { // :345-365
      Edge& e = G.E[i];
      if (!P->mp_inlier[i]) G.compute_error(e);
      const float chi2 = (float)e.chi2();
      if (chi2 > cfg->mono_point) {
        P->mp_inlier[i] = 0;
        e.level = 1;
        num_outlier++;
      } else {
        P->mp_inlier[i] = 1;
        e.level = 0;
      }
      if (iter == 2) e.robust = false;
    }
    for (int i = 0; i < P->n_stereo_pt; ++i) { // :368-385
      Edge& e = G.E[e_stereo0 + i];
      if (!P->sp_inlier[i]) G.compute_error(e);
      const float chi2 = (float)e.chi2();
      if (chi2 > cfg->stereo_point) {
        P->sp_inlier[i] = 0;
        e.level = 1;
        num_outlier++;
      } else {
        P->sp_inlier[i] = 1;
        e.level = 0;
      }
      if (iter == 2) e.robust = false;
    }
    for (int i = 0; i < P->n_mono_ln; ++i) { // extension: same classification for the line edges
      Edge& e = G.E[e_mline0 + i];
      if (!P->ml_inlier[i]) G.compute_error(e);
      const float chi2 = (float)e.chi2();
      if (chi2 > cfg->mono_line) {
        P->ml_inlier[i] = 0;
        e.level = 1;
        num_outlier++;
      } else {
        P->ml_inlier[i] = 1;
        e.level = 0;
      }
      if (iter == 2) e.robust = false;
    }
    for (int i = 0; i < P->n_stereo_ln; ++i) {
      Edge& e = G.E[e_sline0 + i];
      if (!P->sl_inlier[i]) G.compute_error(e);
      const float chi2 = (float)e.chi2();
      if (chi2 > cfg->stereo_line) {
        P->sl_inlier[i] = 0;
        e.level = 1;
        num_outlier++;
      } else {
        P->sl_inlier[i] = 1;
        e.level = 0;
      }
      if (iter == 2) e.robust = false;
    }
    if (G.E.size() < 10) break; // :387 (total, not active, edges)
  }
  if (stats) stats->final_chi2 = chi;
  SE3 Twc = se3_inverse(G.V[0].pose); // :391-393
  for (int k = 0; k < 3; ++k) P->pose_p[k] = Twc.t[k];
  P->pose_q[0] = Twc.r.x;
  P->pose_q[1] = Twc.r.y;
  P->pose_q[2] = Twc.r.z;
  P->pose_q[3] = Twc.r.w;
  return P->n_mono_pt + P->n_stereo_pt + P->n_mono_ln + P->n_stereo_ln - num_outlier; // :396 (+ extension edges)
}

extern "C" int orc_max_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

// One problem per worker thread, dynamic scheduling over an atomic counter (what an OpenMP
// "parallel for schedule(dynamic,1)" over independent g2o optimizers would do).
template <class F>
static void parallel_for(int n, int n_threads, F f) {
  if (n_threads <= 0) n_threads = orc_max_threads();
  n_threads = std::min(n_threads, std::max(n, 1));
  if (n_threads <= 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t)
    pool.emplace_back([&]() {
      for (;;) {
        int i = next.fetch_add(1);
        if (i >= n) break;
        f(i);
      }
    });
  for (auto& th : pool) th.join();
}

extern "C" int orc_frame_opt_batch(int32_t n, OrcFrameProblem* probs, const OrcConfig* cfg, OrcStats* stats,
                                   int32_t* ret, int32_t n_threads) {
  parallel_for(n, n_threads, [&](int i) {
    int r = orc_frame_opt(&probs[i], cfg, stats ? &stats[i] : nullptr);
    if (ret) ret[i] = r;
  });
  return 0;
}

extern "C" int orc_local_ba_batch(int32_t n, OrcLocalProblem* probs, const OrcConfig* cfg, OrcStats* stats,
                                  int32_t n_threads) {
  std::atomic<int> bad{0};
  parallel_for(n, n_threads, [&](int i) {
    int r = orc_local_ba(&probs[i], cfg, stats ? &stats[i] : nullptr);
    if (r != 0) bad++;
  });
  return bad ? -1 : 0;
}

// ================================================================================================
// unit-level entry points
// ================================================================================================
static SE3 pose7_to_se3(const double* p) { return SE3{Quat{p[0], p[1], p[2], p[3]}, mk(p[4], p[5], p[6])}; }
static void se3_to_pose7(const SE3& T, double* p) {
  p[0] = T.r.x;
  p[1] = T.r.y;
  p[2] = T.r.z;
  p[3] = T.r.w;
  p[4] = T.t[0];
  p[5] = T.t[1];
  p[6] = T.t[2];
}
extern "C" void orc_pose_from_twc(const double* p3, const double* q4, double* pose7) {
  se3_to_pose7(se3_inverse(se3_make(Quat{q4[0], q4[1], q4[2], q4[3]}, mk(p3[0], p3[1], p3[2]))), pose7);
}
extern "C" void orc_pose_to_twc(const double* pose7, double* p3, double* q4) {
  SE3 T = se3_inverse(pose7_to_se3(pose7));
  for (int k = 0; k < 3; ++k) p3[k] = T.t[k];
  q4[0] = T.r.x;
  q4[1] = T.r.y;
  q4[2] = T.r.z;
  q4[3] = T.r.w;
}
extern "C" void orc_se3_exp(const double* u6, double* pose7) { se3_to_pose7(se3_exp(u6), pose7); }
extern "C" void orc_pose_oplus(const double* pose7, const double* u6, double* out7) {
  se3_to_pose7(se3_mul(se3_exp(u6), pose7_to_se3(pose7)), out7);
}
extern "C" void orc_line_oplus(const double* L6, const double* v4, double* out6) {
  Line L;
  std::memcpy(L.l, L6, sizeof(L.l));
  L = line_oplus(L, v4);
  std::memcpy(out6, L.l, sizeof(L.l));
}
extern "C" void orc_line_from_cartesian(const double* pv6, double* out6) {
  Line L = line_from_cartesian(pv6);
  std::memcpy(out6, L.l, sizeof(L.l));
}
extern "C" void orc_line_transform(const double* pose7, const double* L6, double* out6) {
  SE3 T = pose7_to_se3(pose7);
  Line L;
  std::memcpy(L.l, L6, sizeof(L.l));
  L = line_transform(q_to_R(T.r), T.t, L);
  std::memcpy(out6, L.l, sizeof(L.l));
}
extern "C" int orc_edge_eval(int edge_type, const double* pose7, const double* lm, const double* meas,
                             const double* cam5, int stereo_bf_float, double* err, double* Jl, double* Jp) {
  Graph G;
  G.bf_float = stereo_bf_float != 0;
  Vertex vp;
  vp.kind = V_POSE;
  vp.id = 0;
  vp.pose = pose7_to_se3(pose7);
  G.V.push_back(vp);
  Edge e;
  e.type = edge_type;
  e.v_pose = 0;
  set_cam(e, cam5);
  e.b = e.bf / e.fx;
  e.Kv[0] = -e.fy * e.cx;
  e.Kv[1] = -e.fx * e.cy;
  e.Kv[2] = e.fx * e.fy;
  std::memset(e.err, 0, sizeof(e.err));
  std::memset(e.Jl, 0, sizeof(e.Jl));
  std::memset(e.Jp, 0, sizeof(e.Jp));
  static const int dims[6] = {2, 3, 2, 4, 2, 3};
  static const int nmeas[6] = {2, 3, 4, 8, 2, 3};
  if (edge_type < 0 || edge_type > 5) return -1;
  e.dim = dims[edge_type];
  e.info = (edge_type == E_MONO_LN || edge_type == E_STEREO_LN) ? 0.1 : 1.0;
  for (int k = 0; k < nmeas[edge_type]; ++k) e.meas[k] = meas[k];
  int ld = 0;
  if (edge_type <= E_STEREO_LN) {
    Vertex vl;
    const bool is_line = edge_type >= E_MONO_LN;
    vl.kind = is_line ? V_LINE : V_POINT;
    vl.id = 1;
    vl.marginalized = true;
    std::memset(vl.est, 0, sizeof(vl.est));
    for (int k = 0; k < (is_line ? 6 : 3); ++k) vl.est[k] = lm[k];
    G.V.push_back(vl);
    e.v_lm = 1;
    ld = is_line ? 4 : 3;
  } else {
    for (int k = 0; k < 3; ++k) e.Xw[k] = lm[k];
  }
  G.compute_error(e);
  G.linearize(e);
  for (int r = 0; r < e.dim; ++r) err[r] = e.err[r];
  if (Jl)
    for (int k = 0; k < e.dim * ld; ++k) Jl[k] = e.Jl[k];
  if (Jp)
    for (int k = 0; k < e.dim * 6; ++k) Jp[k] = e.Jp[k];
  return e.dim;
}
extern "C" void orc_huber(double chi2, double thr, double* rho3) {
  const float d = std::sqrt(thr);
  huber(chi2, (double)d, rho3);
}

// ------------------------------------------------------------------------------------------------
// Triangulation of new map points: Map::TriangulateMappoint (src/map.cc:292-339). The least-squares
// point closest to the observation rays: A = N I - sum b b^T / |b|^2, rhs = sum c - sum b (b . c) / |b|^2
// (:320-326), solved with Eigen::ColPivHouseholderQR<Matrix3d> under setThreshold(1e-5) (:328-334).
// Eigen is not vendored by the reference; the QR below restates its algorithm (Householder reflections with
// column pivoting on the largest remaining column norm; rank = pivots with |R_ii| > threshold * max |R_jj|).
// ------------------------------------------------------------------------------------------------
namespace {
struct Qr3 {
  double R[3][3]; // upper triangle after the reflections, Householder vectors below
  double tau[3];
  int perm[3];
  int rank;
};
void qr3_colpiv(const double A[3][3], double threshold, Qr3& q) {
  double M[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i][j] = A[i][j];
  for (int j = 0; j < 3; ++j) q.perm[j] = j;
  double maxpivot = 0.0;
  for (int k = 0; k < 3; ++k) {
    int best = k;
    double best_n2 = -1.0;
    for (int j = k; j < 3; ++j) {
      double n2 = 0.0;
      for (int i = k; i < 3; ++i) n2 += M[i][j] * M[i][j];
      if (n2 > best_n2) {
        best_n2 = n2;
        best = j;
      }
    }
    if (best != k) {
      for (int i = 0; i < 3; ++i) std::swap(M[i][k], M[i][best]);
      std::swap(q.perm[k], q.perm[best]);
    }
    // makeHouseholder on M[k..2][k]
    double tail2 = 0.0;
    for (int i = k + 1; i < 3; ++i) tail2 += M[i][k] * M[i][k];
    const double c0 = M[k][k];
    double beta, tau;
    if (tail2 <= std::numeric_limits<double>::min()) {
      tau = 0.0;
      beta = c0;
      for (int i = k + 1; i < 3; ++i) M[i][k] = 0.0;
    } else {
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      for (int i = k + 1; i < 3; ++i) M[i][k] /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    M[k][k] = beta;
    q.tau[k] = tau;
    if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
    // apply H = I - tau v v^T (v = [1, essential]) to the remaining columns
    for (int j = k + 1; j < 3; ++j) {
      double dot = M[k][j];
      for (int i = k + 1; i < 3; ++i) dot += M[i][k] * M[i][j];
      dot *= tau;
      M[k][j] -= dot;
      for (int i = k + 1; i < 3; ++i) M[i][j] -= dot * M[i][k];
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) q.R[i][j] = M[i][j];
  q.rank = 0;
  for (int i = 0; i < 3; ++i)
    if (std::fabs(M[i][i]) > threshold * maxpivot) ++q.rank;
}
void qr3_solve(const Qr3& q, const double b[3], double x[3]) {
  double c[3] = {b[0], b[1], b[2]};
  for (int k = 0; k < 3; ++k) { // c = Q^T b
    double dot = c[k];
    for (int i = k + 1; i < 3; ++i) dot += q.R[i][k] * c[i];
    dot *= q.tau[k];
    c[k] -= dot;
    for (int i = k + 1; i < 3; ++i) c[i] -= dot * q.R[i][k];
  }
  double y[3];
  for (int i = 2; i >= 0; --i) {
    double v = c[i];
    for (int j = i + 1; j < 3; ++j) v -= q.R[i][j] * y[j];
    y[i] = v / q.R[i][i];
  }
  for (int j = 0; j < 3; ++j) x[q.perm[j]] = y[j];
}
} // namespace

extern "C" int orc_triangulate_points(int32_t n_points, const int32_t* obs_begin, const int32_t* obs_frame, const double* obs_uv,
                                      int32_t n_obs, const double* frame_twc, int32_t n_frames, const double* cam5,
                                      double* out_xyz, uint8_t* out_ok) {
  const double fx_inv = 1.0 / cam5[0], fy_inv = 1.0 / cam5[1], cx = cam5[2], cy = cam5[3];
  int n_done = 0;
  for (int i = 0; i < n_points; ++i) {
    out_ok[i] = 0;
    const int o0 = obs_begin[i], o1 = obs_begin[i + 1];
    const int N = o1 - o0;
    if (N < 2) continue; // :317
    double BBt[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, csum[3] = {0, 0, 0}, bbc[3] = {0, 0, 0};
    for (int o = o0; o < o1; ++o) {
      const int f = obs_frame[o];
      const M3 R3 = q_to_R(Quat{frame_twc[3 * n_frames + f], frame_twc[4 * n_frames + f], frame_twc[5 * n_frames + f],
                                frame_twc[6 * n_frames + f]});
      const double c[3] = {frame_twc[f], frame_twc[n_frames + f], frame_twc[2 * n_frames + f]};
      const double bp[3] = {(obs_uv[o] - cx) * fx_inv, (obs_uv[n_obs + o] - cy) * fy_inv, 1.0};
      double b[3];
      for (int r = 0; r < 3; ++r) b[r] = R3.m[r][0] * bp[0] + R3.m[r][1] * bp[1] + R3.m[r][2] * bp[2];
      const double inv_n2 = 1.0 / (b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
      const double bc = b[0] * c[0] + b[1] * c[1] + b[2] * c[2];
      for (int r = 0; r < 3; ++r) {
        for (int s2 = 0; s2 < 3; ++s2) BBt[r][s2] += b[r] * inv_n2 * b[s2];
        csum[r] += c[r];
        bbc[r] += b[r] * inv_n2 * bc;
      }
    }
    double A[3][3], rhs[3];
    for (int r = 0; r < 3; ++r) {
      for (int s2 = 0; s2 < 3; ++s2) A[r][s2] = (r == s2 ? (double)N : 0.0) - BBt[r][s2];
      rhs[r] = csum[r] - bbc[r];
    }
    Qr3 qr;
    qr3_colpiv(A, 1e-5, qr);
    if (qr.rank < 3) continue; // :332
    double x[3];
    qr3_solve(qr, rhs, x);
    out_xyz[i] = x[0];
    out_xyz[n_points + i] = x[1];
    out_xyz[2 * n_points + i] = x[2];
    out_ok[i] = 1;
    ++n_done;
  }
  return n_done;
}

// ---- line endpoint refresh after the local BA (Map::UppdateMapline, map.cc:121-177) ----
// g2o::Line3D::toCartesian (g2o types/slam3d_addons/line3d.cpp; g2o is not in the tree, see oracle.h): direction
// d / |d|, anchor = (W^T W + 1e-9 I).ldlt().solve(W^T w) with W = -skew(d). Eigen's LDLT<Matrix3d> is restated as its
// published unblocked algorithm: pivot on the largest |diagonal| of the not yet updated trailing part, left-looking
// column update, unit-lower forward solve in axpy order, pseudo-inverse of D, unit-upper backward solve in dot order.
static void ldlt3_solve(double A[3][3], const double b[3], double x[3]) {
  int tr[3];
  for (int k = 0; k < 3; ++k) {
    int big = k;
    for (int i = k + 1; i < 3; ++i)
      if (std::fabs(A[i][i]) > std::fabs(A[big][big])) big = i; // maxCoeff: first maximum
    tr[k] = big;
    if (big != k) { // symmetric transposition on the lower triangle
      for (int j = 0; j < k; ++j) std::swap(A[k][j], A[big][j]);
      for (int i = big + 1; i < 3; ++i) std::swap(A[i][k], A[i][big]);
      std::swap(A[k][k], A[big][big]);
      for (int i = k + 1; i < big; ++i) std::swap(A[i][k], A[big][i]);
    }
    if (k > 0) {
      double temp[2];
      for (int j = 0; j < k; ++j) temp[j] = A[j][j] * A[k][j];
      double acc = A[k][0] * temp[0];
      for (int j = 1; j < k; ++j) acc += A[k][j] * temp[j];
      A[k][k] -= acc;
      for (int i = k + 1; i < 3; ++i) {
        double a2 = A[i][0] * temp[0];
        for (int j = 1; j < k; ++j) a2 += A[i][j] * temp[j];
        A[i][k] -= a2;
      }
    }
    const double akk = A[k][k];
    if (std::fabs(akk) > 0.0)
      for (int i = k + 1; i < 3; ++i) A[i][k] /= akk;
  }
  double y[3] = {b[0], b[1], b[2]};
  for (int k = 0; k < 3; ++k) std::swap(y[k], y[tr[k]]);
  y[1] -= y[0] * A[1][0]; // L y' = y, column by column
  y[2] -= y[0] * A[2][0];
  y[2] -= y[1] * A[2][1];
  const double tol = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < 3; ++i) y[i] = std::fabs(A[i][i]) > tol ? y[i] / A[i][i] : 0.0;
  y[1] -= A[2][1] * y[2]; // L^T x = y, row by row
  y[0] -= A[1][0] * y[1] + A[2][0] * y[2];
  for (int k = 2; k >= 0; --k) std::swap(y[k], y[tr[k]]);
  x[0] = y[0];
  x[1] = y[1];
  x[2] = y[2];
}

static void line_to_cartesian(const double* wd, double* cart /* anchor(3), direction(3) */) {
  const double w0 = wd[0], w1 = wd[1], w2 = wd[2], dx = wd[3], dy = wd[4], dz = wd[5];
  const double nrm = std::sqrt(dx * dx + dy * dy + dz * dz);
  cart[3] = dx / nrm;
  cart[4] = dy / nrm;
  cart[5] = dz / nrm;
  // W = [[0, dz, -dy], [-dz, 0, dx], [dy, -dx, 0]]; A = W^T W + 1e-9 I, entries summed over k = 0, 1, 2
  double A[3][3];
  A[0][0] = (dz * dz + dy * dy) + 1e-9;
  A[1][1] = (dz * dz + dx * dx) + 1e-9;
  A[2][2] = (dy * dy + dx * dx) + 1e-9;
  A[1][0] = A[0][1] = -(dx * dy);
  A[2][0] = A[0][2] = -(dx * dz);
  A[2][1] = A[1][2] = -(dy * dz);
  const double b[3] = {-(dz * w1) + dy * w2, dz * w0 + -(dx * w2), -(dy * w0) + dx * w1};
  ldlt3_solve(A, b, cart);
}

extern "C" void orc_line_to_cartesian(const double* wd6, double* cart6) { line_to_cartesian(wd6, cart6); }

extern "C" int orc_update_maplines(int32_t n_lines, const double* line_wd, const int32_t* pt_begin, const int32_t* pt_index,
                                   const double* point_xyz, int32_t n_points, double* endpoints, uint8_t* out_ok) {
  int done = 0;
  for (int l = 0; l < n_lines; ++l) {
    out_ok[l] = 0;
    double wd[6], cart[6];
    for (int k = 0; k < 6; ++k) wd[k] = line_wd[(size_t)k * n_lines + l];
    if (pt_begin[l + 1] == pt_begin[l]) continue; // map.cc:127 (no observers) and :166 (nothing found)
    line_to_cartesian(wd, cart);
    const double* lp = cart;
    const double* v = cart + 3;
    int md = 0; // array().abs().maxCoeff: first maximum
    for (int k = 1; k < 3; ++k)
      if (std::fabs(v[k]) > std::fabs(v[md])) md = k;
    // (sic) DBL_MIN is the smallest POSITIVE double: a line whose nearby points all have a non-positive main
    // coordinate finds no maximum and keeps its old endpoints (map.cc:153-166)
    double max_d = DBL_MIN, min_d = DBL_MAX;
    bool find_max = false, find_min = false;
    for (int o = pt_begin[l]; o < pt_begin[l + 1]; ++o) {
      const int pi = pt_index[o];
      const double X[3] = {point_xyz[pi], point_xyz[(size_t)n_points + pi], point_xyz[2 * (size_t)n_points + pi]};
      const double dp[3] = {X[0] - lp[0], X[1] - lp[1], X[2] - lp[2]};
      const double c0 = v[1] * dp[2] - v[2] * dp[1], c1 = v[2] * dp[0] - v[0] * dp[2], c2 = v[0] * dp[1] - v[1] * dp[0];
      const double dist = std::sqrt(c0 * c0 + c1 * c1 + c2 * c2); // line_processor.cc:77-90
      if (dist > 0.2) continue;
      const double di = X[md];
      if (di > max_d) {
        max_d = di;
        find_max = true;
      }
      if (di < min_d) {
        min_d = di;
        find_min = true;
      }
    }
    if (!find_max || !find_min) continue;
    const double r1 = (max_d - lp[md]) / v[md], r2 = (min_d - lp[md]) / v[md];
    for (int k = 0; k < 3; ++k) {
      endpoints[(size_t)k * n_lines + l] = lp[k] + r1 * v[k];
      endpoints[(size_t)(3 + k) * n_lines + l] = lp[k] + r2 * v[k];
    }
    out_ok[l] = 1;
    ++done;
  }
  return done;
}
