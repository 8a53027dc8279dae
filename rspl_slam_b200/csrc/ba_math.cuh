// ba_math.cuh — device-side geometry of the point+line BA path (sm_100a, fp64).
//
// What each function replaces in the reference stack (g2o is un-vendored by the reference; the
// behaviour is the one SURVEY.md §9 fixes):
//   SE3 state / exp / oplus           g2o::SE3Quat, VertexSE3Expmap::oplusImpl           (§9.2)
//   point residuals + Jacobians       g2o::Edge(Stereo)SE3ProjectXYZ(+OnlyPose)          (§9.3)
//   line residuals                    /root/reference/src/g2o_optimization/edge_project_line.cc:21-42,
//                                     edge_project_stereo_line.cc:22-51
//   line manifold                     g2o::Line3D::oplus via VertexLine3D (vertex_line3d.h:26-29) (§9.6)
//   Huber                             g2o::RobustKernelHuber, delta = (float)sqrt(thr)    (§9.5)
// The line Jacobians are analytic here (the reference gets them from g2o's numeric central
// differences, §9.8); tests/test_edges_gpu.py gates them against the oracle's numeric ones.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#define BA_DEV __device__ __forceinline__

namespace ba {

struct Cam {
  double fx, fy, cx, cy, bf;
};

// optimiser pose Tcw: unit quaternion (x,y,z,w) + translation, as g2o::SE3Quat stores it
struct Pose {
  double q[4];
  double t[3];
};

// Checked build (-DRSPL_BA_CHECKED, tests/test_checked_build.py): bounds / invariant asserts in the kernels. The pool's
// compute-sanitizer is closed, so this is the substitute for memcheck: a violated assert prints its location and traps
// (the launch then fails with an error the tests see). Compiled out of the shipped library.
#ifdef RSPL_BA_CHECKED
#include <stdio.h>
#define BA_CHECK(cond)                                                                                      \
  do {                                                                                                      \
    if (!(cond)) {                                                                                          \
      printf("BA_CHECK failed: %s (%s:%d) block (%d,%d,%d) thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
             (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x);                                           \
      __trap();                                                                                             \
    }                                                                                                       \
  } while (0)
#else
#define BA_CHECK(cond) ((void)0)
#endif

BA_DEV void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// ------------------------------------------------------------------------------------------------
// Reciprocal / reciprocal square root / square root without the library's range test (all hot loops). The CUDA library versions test the exponent range
// and call an out-of-line slow path; that conditional call ends the basic block, so the compiler cannot interleave
// the evaluation of several edges. These are the same MUFU seed + FMA refinement without the range test: for normal
// arguments (|x| in ~[1e-280, 1e280]) the reciprocal came out correctly rounded on every one of 200 k samples over 560
// decades and rsqrt / sqrt within 2 ulp (tests/test_edges_gpu.py); 0, denormals and infinities
// give NaN instead of +-inf / 0 (a depth of exactly 0 or a chi2 of 0 under the square root never reach them:
// see the callers).
// ------------------------------------------------------------------------------------------------
BA_DEV double rcp_nr(double z) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(z)); // MUFU.RCP64H: ~20 bits
  double e = fma(-z, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y); // y (1 + e + e^2): ~60 bits
  e = fma(-z, y, 1.0);
  return fma(y, e, y); // final correction
}
BA_DEV double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); // MUFU.RSQ64H: ~20 bits
  double e = fma(-x, y * y, 1.0);
  double p = fma(e, 0.375, 0.5);
  y = fma(p, e * y, y); // y (1 + e/2 + 3 e^2/8): ~60 bits
  e = fma(-x, y * y, 1.0);
  return fma(0.5 * e, y, y); // final correction
}
BA_DEV double sqrt_nr(double x) { // x * rsqrt(x), exact at 0
  const double r = x * rsqrt_nr(x);
  return x == 0.0 ? 0.0 : r;
}


BA_DEV void quat_to_R(const double* q, double* R) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1 - (txx + tyy);
}

// rotation matrix -> quaternion with Eigen's branch structure
BA_DEV void R_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    q[0] = v[0];
    q[1] = v[1];
    q[2] = v[2];
  }
}

BA_DEV void quat_rot(const double* q, const double* v, double* o) {
  double uv[3], c[3];
  cross3(q, v, uv);
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  cross3(q, uv, c);
  o[0] = v[0] + q[3] * uv[0] + c[0];
  o[1] = v[1] + q[3] * uv[1] + c[1];
  o[2] = v[2] + q[3] * uv[2] + c[2];
}

BA_DEV void pose_normalize(Pose& T) {
  if (T.q[3] < 0) {
    T.q[0] = -T.q[0];
    T.q[1] = -T.q[1];
    T.q[2] = -T.q[2];
    T.q[3] = -T.q[3];
  }
  const double inv = rsqrt(T.q[0] * T.q[0] + T.q[1] * T.q[1] + T.q[2] * T.q[2] + T.q[3] * T.q[3]);
  T.q[0] *= inv;
  T.q[1] *= inv;
  T.q[2] *= inv;
  T.q[3] *= inv;
}

BA_DEV Pose pose_inverse(const Pose& T) {
  Pose r;
  r.q[0] = -T.q[0];
  r.q[1] = -T.q[1];
  r.q[2] = -T.q[2];
  r.q[3] = T.q[3];
  const double mt[3] = {-T.t[0], -T.t[1], -T.t[2]};
  quat_rot(r.q, mt, r.t);
  return r;
}

// caller's Twc (p, q) -> optimiser Tcw, g2o_optimization.cc:42 / :271
BA_DEV Pose pose_from_twc(const double* p, const double* q) {
  Pose T;
  T.q[0] = q[0];
  T.q[1] = q[1];
  T.q[2] = q[2];
  T.q[3] = q[3];
  T.t[0] = p[0];
  T.t[1] = p[1];
  T.t[2] = p[2];
  pose_normalize(T);
  return pose_inverse(T);
}

BA_DEV Pose pose_mul(const Pose& a, const Pose& b) {
  Pose r;
  double rt[3];
  quat_rot(a.q, b.t, rt);
  r.t[0] = a.t[0] + rt[0];
  r.t[1] = a.t[1] + rt[1];
  r.t[2] = a.t[2] + rt[2];
  r.q[3] = a.q[3] * b.q[3] - a.q[0] * b.q[0] - a.q[1] * b.q[1] - a.q[2] * b.q[2];
  r.q[0] = a.q[3] * b.q[0] + a.q[0] * b.q[3] + a.q[1] * b.q[2] - a.q[2] * b.q[1];
  r.q[1] = a.q[3] * b.q[1] + a.q[1] * b.q[3] + a.q[2] * b.q[0] - a.q[0] * b.q[2];
  r.q[2] = a.q[3] * b.q[2] + a.q[2] * b.q[3] + a.q[0] * b.q[1] - a.q[1] * b.q[0];
  pose_normalize(r);
  return r;
}

// g2o::SE3Quat::exp, update = [omega(3), upsilon(3)]
BA_DEV Pose pose_exp(const double* u) {
  const double wx = u[0], wy = u[1], wz = u[2];
  const double theta = sqrt(wx * wx + wy * wy + wz * wz);
  // Omega = skew(omega), Omega2 = Omega*Omega
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  const double O2[9] = {-(wy * wy + wz * wz), wx * wy, wx * wz, wx * wy, -(wx * wx + wz * wz), wy * wz,
                        wx * wz, wy * wz, -(wx * wx + wy * wy)};
  double a, b, c;
  if (theta < 0.00001) {
    a = 1.0;
    b = 0.5;
    c = 1.0 / 6.0;
  } else {
    double s, co;
    sincos(theta, &s, &co);
    a = s / theta;
    b = (1 - co) / (theta * theta);
    c = (theta - s) / (theta * theta * theta);
  }
  double R[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = id + a * O[i] + b * O2[i];
    V[i] = id + b * O[i] + c * O2[i];
  }
  Pose T;
  R_to_quat(R, T.q);
  T.t[0] = V[0] * u[3] + V[1] * u[4] + V[2] * u[5];
  T.t[1] = V[3] * u[3] + V[4] * u[4] + V[5] * u[5];
  T.t[2] = V[6] * u[3] + V[7] * u[4] + V[8] * u[5];
  pose_normalize(T);
  return T;
}

// VertexSE3Expmap::oplusImpl: T <- exp(u) * T
BA_DEV Pose pose_oplus(const Pose& T, const double* u) { return pose_mul(pose_exp(u), T); }

// The same update on a rotation-matrix pose, for inner loops that never need the quaternion:
// R <- dR R, t <- dR t + V upsilon with (dR, V) of SE3Quat::exp. The quaternion round trip and its
// two normalisations (which keep g2o's estimate on the manifold to rounding level) are skipped; the
// caller converts back once at the end. One reciprocal replaces the three divisions of the closed form.
struct PoseRt {
  double R[9];
  double t[3];
};
BA_DEV void poseRt_oplus(const PoseRt& T, const double* u, PoseRt& out) {
  const double wx = u[0], wy = u[1], wz = u[2];
  const double th2 = wx * wx + wy * wy + wz * wz;
  double a, b, c;
  if (th2 < 1e-10) { // theta < 1e-5
    a = 1.0;
    b = 0.5;
    c = 1.0 / 6.0;
  } else {
    const double theta = sqrt(th2), it = 1.0 / theta;
    double s, co;
    sincos(theta, &s, &co);
    a = s * it;
    b = (1 - co) * it * it;
    c = (theta - s) * it * it * it;
  }
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  const double O2[9] = {-(wy * wy + wz * wz), wx * wy, wx * wz, wx * wy, -(wx * wx + wz * wz), wy * wz,
                        wx * wz, wy * wz, -(wx * wx + wy * wy)};
  double dR[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    dR[i] = id + a * O[i] + b * O2[i];
    V[i] = id + b * O[i] + c * O2[i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      out.R[3 * i + j] = dR[3 * i] * T.R[j] + dR[3 * i + 1] * T.R[3 + j] + dR[3 * i + 2] * T.R[6 + j];
    out.t[i] = dR[3 * i] * T.t[0] + dR[3 * i + 1] * T.t[1] + dR[3 * i + 2] * T.t[2] + V[3 * i] * u[3] +
               V[3 * i + 1] * u[4] + V[3 * i + 2] * u[5];
  }
}

// ------------------------------------------------------------------------------------------------
// cp.async (global -> shared memory without staging registers); groups are per thread
// ------------------------------------------------------------------------------------------------
BA_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
BA_DEV void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
BA_DEV void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
BA_DEV void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
BA_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
BA_DEV void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Huber (RobustKernelHuber::robustify); returns rho0, writes the weight rho1
// ------------------------------------------------------------------------------------------------
BA_DEV double huber(double e, double delta, double& w) {
  const double dsqr = delta * delta;
  if (e <= dsqr) {
    w = 1.0;
    return e;
  }
  // one rsqrt (without range test) instead of sqrt + divide (the pair costs ~50 fp64 instructions on the device):
  // rho1 = delta / sqrt(e), rho0 = 2 delta sqrt(e) - delta^2 with sqrt(e) = e * rsqrt(e)
  const double rs = rsqrt_nr(e);
  w = delta * rs;
  return 2 * (e * rs) * delta - dsqr;
}

// ------------------------------------------------------------------------------------------------
// point edges. Xc = R X + t (camera frame). STEREO: 3 rows (u, v, u_right); else 2 rows.
// residual = meas - projection. bf_res is the bf used in the residual (float-rounded for the
// binary stereo edge when stereo_bf_float, §9.3), bf the one used in the Jacobians.
// ------------------------------------------------------------------------------------------------
BA_DEV void transform_point(const double* R, const double* t, const double* X, double* Xc) {
  Xc[0] = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  Xc[1] = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  Xc[2] = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
}

// 1 / z given (rcp_nr: within an ulp of the reference's division; its mono edge divides x / z, here x * (1 / z))
template <bool STEREO>
BA_DEV void point_residual_iz(const Cam& cam, double bf_res, const double* Xc, double invz, const double* m, double* r) {
  const double u = Xc[0] * invz * cam.fx + cam.cx;
  const double v = Xc[1] * invz * cam.fy + cam.cy;
  r[0] = m[0] - u;
  r[1] = m[1] - v;
  if (STEREO) r[2] = m[2] - (u - bf_res * invz);
}

template <bool STEREO>
BA_DEV void point_residual(const Cam& cam, double bf_res, const double* Xc, const double* m, double* r) {
  point_residual_iz<STEREO>(cam, bf_res, Xc, rcp_nr(Xc[2]), m, r);
}

// d r / d xi (rows x 6, omega first), row-major Jp[row*6+col]
template <bool STEREO>
BA_DEV void point_jac_pose_iz(const Cam& cam, const double* Xc, double invz, double* Jp);
template <bool STEREO>
BA_DEV void point_jac_pose(const Cam& cam, const double* Xc, double* Jp) {
  point_jac_pose_iz<STEREO>(cam, Xc, rcp_nr(Xc[2]), Jp);
}
template <bool STEREO>
BA_DEV void point_jac_pose_iz(const Cam& cam, const double* Xc, double invz, double* Jp) {
  const double x = Xc[0], y = Xc[1];
  const double invz2 = invz * invz;
  Jp[0] = x * y * invz2 * cam.fx;
  Jp[1] = -(1 + x * x * invz2) * cam.fx;
  Jp[2] = y * invz * cam.fx;
  Jp[3] = -invz * cam.fx;
  Jp[4] = 0;
  Jp[5] = x * invz2 * cam.fx;
  Jp[6] = (1 + y * y * invz2) * cam.fy;
  Jp[7] = -x * y * invz2 * cam.fy;
  Jp[8] = -x * invz * cam.fy;
  Jp[9] = 0;
  Jp[10] = -invz * cam.fy;
  Jp[11] = y * invz2 * cam.fy;
  if (STEREO) {
    Jp[12] = Jp[0] - cam.bf * y * invz2;
    Jp[13] = Jp[1] + cam.bf * x * invz2;
    Jp[14] = Jp[2];
    Jp[15] = Jp[3];
    Jp[16] = 0;
    Jp[17] = Jp[5] - cam.bf * invz2;
  }
}

// d r / d X (rows x 3), row-major Jl[row*3+col]
template <bool STEREO>
BA_DEV void point_jac_point(const Cam& cam, const double* R, const double* Xc, double* Jl) {
  const double x = Xc[0], y = Xc[1];
  const double invz = rcp_nr(Xc[2]), invz2 = invz * invz;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    Jl[k] = -cam.fx * R[k] * invz + cam.fx * x * R[6 + k] * invz2;
    Jl[3 + k] = -cam.fy * R[3 + k] * invz + cam.fy * y * R[6 + k] * invz2;
    if (STEREO) Jl[6 + k] = Jl[k] - cam.bf * R[6 + k] * invz2;
  }
}

// ------------------------------------------------------------------------------------------------
// line edges. L = [w(3), d(3)] world Pluecker line. meas: left x1,y1,x2,y2 [, right x1,y1,x2,y2].
// ------------------------------------------------------------------------------------------------
struct LineCam { // camera-frame quantities shared by residual and Jacobians
  double wc[3]; // left moment
  double dc[3]; // direction
};

BA_DEV void line_to_camera(const double* R, const double* t, const double* L, LineCam& lc) {
  double Rw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Rw[i] = R[3 * i] * L[0] + R[3 * i + 1] * L[1] + R[3 * i + 2] * L[2];
    lc.dc[i] = R[3 * i] * L[3] + R[3 * i + 1] * L[4] + R[3 * i + 2] * L[5];
  }
  double c[3];
  cross3(t, lc.dc, c);
  lc.wc[0] = Rw[0] + c[0];
  lc.wc[1] = Rw[1] + c[1];
  lc.wc[2] = Rw[2] + c[2];
}

// residuals of one image (2 rows) from the camera-frame moment wv; also returns the 2x3
// derivative g = d e / d wv when WITH_G.
template <bool WITH_G>
BA_DEV void line_image_residual(const Cam& cam, const double* wv, const double* m, double* e, double* g) {
  const double kv0 = -cam.fy * cam.cx, kv1 = -cam.fx * cam.cy, kv2 = cam.fx * cam.fy;
  const double l0 = cam.fy * wv[0], l1 = cam.fx * wv[1];
  const double l2 = kv0 * wv[0] + kv1 * wv[1] + kv2 * wv[2];
  // 1 / |l_xy| with one rsqrt instead of a square root and three divisions (differs from the divisions of
  // edge_project_line.cc:32-33 by rounding only)
  const double inv = rsqrt_nr(l0 * l0 + l1 * l1);
  e[0] = (m[0] * l0 + m[1] * l1 + l2) * inv;
  e[1] = (m[2] * l0 + m[3] * l1 + l2) * inv;
  if (WITH_G) {
    const double n0 = l0 * inv, n1 = l1 * inv;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      // d e / d l = (p~ - e * (l0, l1, 0)/n) / n ;  d l / d wv = [[fy,0,0],[0,fx,0],[kv0,kv1,kv2]]
      const double a0 = (m[2 * k] - e[k] * n0) * inv;
      const double a1 = (m[2 * k + 1] - e[k] * n1) * inv;
      const double a2 = inv;
      g[3 * k + 0] = a0 * cam.fy + a2 * kv0;
      g[3 * k + 1] = a1 * cam.fx + a2 * kv1;
      g[3 * k + 2] = a2 * kv2;
    }
  }
}

// right-camera moment: T_right = T with t.x -= b (edge_project_stereo_line.cc:34-35)
BA_DEV void line_right_moment(const LineCam& lc, double b, double* wr) {
  wr[0] = lc.wc[0];
  wr[1] = lc.wc[1] + b * lc.dc[2];
  wr[2] = lc.wc[2] - b * lc.dc[1];
}

template <bool STEREO>
BA_DEV void line_residual(const Cam& cam, const double* R, const double* t, const double* L, const double* m,
                          double* r) {
  LineCam lc;
  line_to_camera(R, t, L, lc);
  line_image_residual<false>(cam, lc.wc, m, r, nullptr);
  if (STEREO) {
    double wr[3];
    line_right_moment(lc, cam.bf * rcp_nr(cam.fx), wr);
    line_image_residual<false>(cam, wr, m + 4, r + 2, nullptr);
  }
}

// residual + analytic Jacobians. Jp[row*6+col] (omega, upsilon), Jl[row*4+col] (Line3D::oplus tangent).
template <bool STEREO>
BA_DEV void line_linearize(const Cam& cam, const double* R, const double* t, const double* L, const double* m,
                           double* r, double* Jp, double* Jl) {
  LineCam lc;
  line_to_camera(R, t, L, lc);
  double g[12]; // up to 4 rows x 3: d e / d (camera-frame moment of that image)
  line_image_residual<true>(cam, lc.wc, m, r, g);
  const double b = cam.bf * rcp_nr(cam.fx);
  if (STEREO) {
    double wr[3];
    line_right_moment(lc, b, wr);
    line_image_residual<true>(cam, wr, m + 4, r + 2, g + 6);
  }
  constexpr int ROWS = STEREO ? 4 : 2;
  // ---- pose: d wc = omega x wc + upsilon x dc ; right image adds -b e_x x (omega x dc)
#pragma unroll
  for (int k = 0; k < ROWS; ++k) {
    const double* gk = g + 3 * k;
    // gk . (omega x wc) = omega . (wc x gk)
    double a[3], c[3];
    cross3(lc.wc, gk, a);
    cross3(lc.dc, gk, c);
    if (STEREO && k >= 2) {
      // gk . (-b e_x x (omega x dc)) = omega . (dc x (-b * (gk x e_x)))
      const double h[3] = {0.0, -b * gk[2], b * gk[1]}; // -b * (gk x e_x)
      double e2[3];
      cross3(lc.dc, h, e2);
      a[0] += e2[0];
      a[1] += e2[1];
      a[2] += e2[2];
    }
    Jp[6 * k + 0] = a[0];
    Jp[6 * k + 1] = a[1];
    Jp[6 * k + 2] = a[2];
    Jp[6 * k + 3] = c[0];
    Jp[6 * k + 4] = c[1];
    Jp[6 * k + 5] = c[2];
  }
  // ---- line: orthonormal representation U = [w/|w|, d/|d|, (w x d)/|w x d|], c = |w|/|d|
  const double dd = L[3] * L[3] + L[4] * L[4] + L[5] * L[5], ww = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
  const double ind = rsqrt_nr(dd), inw = rsqrt_nr(ww);
  const double nd = dd * ind, cw = (ww * inw) * ind;
  double u0[3], u1[3], u2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    u0[i] = L[i] * inw;
    u1[i] = L[3 + i] * ind;
  }
  cross3(u0, u1, u2);
  {
    const double in2 = rsqrt_nr(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    u2[0] *= in2;
    u2[1] *= in2;
    u2[2] *= in2;
  }
  // world-frame tangent directions of the normalised line (w~, d~):
  //  a0: dw = 0,            dd =  2 u2
  //  a1: dw = -2 c u2,      dd =  0
  //  a2: dw =  2 c u1,      dd = -2 u0
  //  th: dw = -(1+c^2) u0,  dd =  0
  // the residual is homogeneous of degree 0 in L, so derivatives are taken on the normalised line
  // and divided by nothing further; the camera-frame moment scales with 1/|d| though:
  double dW[4][3], dD[4][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dW[0][i] = 0.0;
    dD[0][i] = 2.0 * u2[i];
    dW[1][i] = -2.0 * cw * u2[i];
    dD[1][i] = 0.0;
    dW[2][i] = 2.0 * cw * u1[i];
    dD[2][i] = -2.0 * u0[i];
    dW[3][i] = -(1.0 + cw * cw) * u0[i];
    dD[3][i] = 0.0;
  }
  // g was evaluated at wc built from L (scale |d|); e is scale free, so d e / d w~ = |d| * g.
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double Rw[3], Rd[3], c3[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Rw[i] = R[3 * i] * dW[j][0] + R[3 * i + 1] * dW[j][1] + R[3 * i + 2] * dW[j][2];
      Rd[i] = R[3 * i] * dD[j][0] + R[3 * i + 1] * dD[j][1] + R[3 * i + 2] * dD[j][2];
    }
    cross3(t, Rd, c3);
    const double dwc[3] = {Rw[0] + c3[0], Rw[1] + c3[1], Rw[2] + c3[2]};
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const double* gk = g + 3 * k;
      double v = gk[0] * dwc[0] + gk[1] * dwc[1] + gk[2] * dwc[2];
      if (STEREO && k >= 2) v += gk[1] * (b * Rd[2]) - gk[2] * (b * Rd[1]); // right moment shift
      Jl[4 * k + j] = nd * v;
    }
  }
}

// Line3D::oplus (vertex_line3d.h:26-29): orthonormal representation update + normalisation
BA_DEV void line_oplus(const double* L, const double* v, double* o) {
  const double w[3] = {L[0], L[1], L[2]}, d[3] = {L[3], L[4], L[5]};
  const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], ww = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double dn = rsqrt_nr(dd), mn = rsqrt_nr(ww); // 1 / |d|, 1 / |w|
  const double mx = dd * dn, my = ww * mn;
  const double wn = rsqrt_nr(mx * mx + my * my);
  const double W00 = my * wn, W01 = -mx * wn, W10 = mx * wn, W11 = my * wn;
  double mdc[3];
  cross3(w, d, mdc);
  const double mdcn = rsqrt_nr(mdc[0] * mdc[0] + mdc[1] * mdc[1] + mdc[2] * mdc[2]);
  double U[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    U[3 * i + 0] = w[i] * mn;
    U[3 * i + 1] = d[i] * dn;
    U[3 * i + 2] = mdc[i] * mdcn;
  }
  double s, c;
  sincos(v[3], &s, &c);
  double q[4] = {v[0], v[1], v[2], sqrt_nr(1 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))};
  const double qn = rsqrt_nr(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= qn;
  q[1] *= qn;
  q[2] *= qn;
  q[3] *= qn;
  double Ru[9];
  quat_to_R(q, Ru);
  // Un = U * Ru (only columns 0 and 1 are used); Wn = W * Rot2(v3) (only column 0 is used)
  const double Wn00 = W00 * c + W01 * s;
  const double Wn10 = W10 * c + W11 * s;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double un0 = U[3 * i] * Ru[0] + U[3 * i + 1] * Ru[3] + U[3 * i + 2] * Ru[6];
    const double un1 = U[3 * i] * Ru[1] + U[3 * i + 1] * Ru[4] + U[3 * i + 2] * Ru[7];
    o[i] = un0 * Wn00;
    o[3 + i] = un1 * Wn10;
  }
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) { // fromOrthonormal normalises, oplus normalises again
    const double n = rsqrt_nr(o[3] * o[3] + o[4] * o[4] + o[5] * o[5]);
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] *= n;
  }
}

} // namespace ba
