// triangulate.cuh — batched triangulation of new map points, SURVEY 8(f) rank 4: the step that produces the landmark
// initial values the bundle adjustment consumes. Replaces the per-point body of Map::TriangulateMappoint
// (/root/reference/src/map.cc:292-339): the point closest (least squares) to its observation rays,
//   A = N I - sum_i b_i b_i^T / |b_i|^2,   rhs = sum_i c_i - sum_i b_i (b_i . c_i) / |b_i|^2        (:320-326)
// with b_i = R_i ((u - cx) / fx, (v - cy) / fy, 1) (camera.cc:150-155), c_i the camera centre, solved with a
// column-pivoting Householder QR of the 3 x 3 matrix and the reference's rank test (threshold 1e-5 on
// |R_ii| / max |R_jj|, :328-332). One thread per point: the observations of a point are a CSR segment, the keyframe
// poses (7 doubles each) are few and stay in L1/L2; HBM-bound at 20 bytes per observation + 25 bytes per point.
#pragma once

#include "ba_math.cuh"

namespace ba {

struct TriDev {
  int n_points, n_obs, n_frames;
  const int* obs_begin;    // [n_points + 1]
  const int* obs_frame;    // [n_obs]
  const double* obs_uv;    // [2][n_obs]
  const double* frame_twc; // [7][n_frames]: p, q (x, y, z, w)
  double fx_inv, fy_inv, cx, cy;
  double* out_xyz;         // [3][n_points]
  uint8_t* out_ok;         // [n_points]
  int* n_done;             // points triangulated
};

// column-pivoting Householder QR of a 3 x 3 matrix (what Eigen::ColPivHouseholderQR<Matrix3d> computes), then
// x = A^-1 b; returns false when fewer than 3 pivots pass |R_ii| > threshold * max |R_jj|. Fully unrolled: every
// index is static.
BA_DEV bool qr3_solve(double (&M)[3][3], const double (&b)[3], double threshold, double (&x)[3]) {
  int perm[3] = {0, 1, 2};
  double c[3] = {b[0], b[1], b[2]};
  double maxpivot = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int best = k;
    double best_n2 = -1.0;
#pragma unroll
    for (int j = k; j < 3; ++j) {
      double n2 = 0.0;
#pragma unroll
      for (int i = k; i < 3; ++i) n2 += M[i][j] * M[i][j];
      if (n2 > best_n2) {
        best_n2 = n2;
        best = j;
      }
    }
#pragma unroll
    for (int j = k + 1; j < 3; ++j) { // swap columns k <-> best (static indices)
      if (best == j) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double t = M[i][k];
          M[i][k] = M[i][j];
          M[i][j] = t;
        }
        const int tp = perm[k];
        perm[k] = perm[j];
        perm[j] = tp;
      }
    }
    double tail2 = 0.0;
#pragma unroll
    for (int i = k + 1; i < 3; ++i) tail2 += M[i][k] * M[i][k];
    const double c0 = M[k][k];
    double beta, tk;
    if (tail2 <= 2.2250738585072014e-308) {
      tk = 0.0;
      beta = c0;
#pragma unroll
      for (int i = k + 1; i < 3; ++i) M[i][k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
#pragma unroll
      for (int i = k + 1; i < 3; ++i) M[i][k] /= (c0 - beta);
      tk = (beta - c0) / beta;
    }
    M[k][k] = beta;
    maxpivot = fmax(maxpivot, fabs(beta));
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      double dot = M[k][j];
#pragma unroll
      for (int i = k + 1; i < 3; ++i) dot += M[i][k] * M[i][j];
      dot *= tk;
      M[k][j] -= dot;
#pragma unroll
      for (int i = k + 1; i < 3; ++i) M[i][j] -= dot * M[i][k];
    }
    { // the right-hand side rides along: c = H_k c
      double dot = c[k];
#pragma unroll
      for (int i = k + 1; i < 3; ++i) dot += M[i][k] * c[i];
      dot *= tk;
      c[k] -= dot;
#pragma unroll
      for (int i = k + 1; i < 3; ++i) c[i] -= dot * M[i][k];
    }
  }
  int rank = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) rank += fabs(M[i][i]) > threshold * maxpivot ? 1 : 0;
  if (rank < 3) return false;
  double y[3];
  y[2] = c[2] / M[2][2];
  y[1] = (c[1] - M[1][2] * y[2]) / M[1][1];
  y[0] = (c[0] - M[0][1] * y[1] - M[0][2] * y[2]) / M[0][0];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (perm[j] == q) x[q] = y[j];
  }
  return true;
}

__global__ void __launch_bounds__(128) triangulate_points_kernel(const __grid_constant__ TriDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_points) return;
  const int o0 = d.obs_begin[i], o1 = d.obs_begin[i + 1], N = o1 - o0;
  bool ok = false;
  double x[3] = {0, 0, 0};
  if (N >= 2) { // map.cc:317
    double BBt[6] = {0, 0, 0, 0, 0, 0}, csum[3] = {0, 0, 0}, bbc[3] = {0, 0, 0}; // BBt: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
    for (int o = o0; o < o1; ++o) {
      const int f = d.obs_frame[o];
      const double q[4] = {d.frame_twc[3 * (size_t)d.n_frames + f], d.frame_twc[4 * (size_t)d.n_frames + f],
                           d.frame_twc[5 * (size_t)d.n_frames + f], d.frame_twc[6 * (size_t)d.n_frames + f]};
      double R[9];
      quat_to_R(q, R);
      const double c[3] = {d.frame_twc[f], d.frame_twc[(size_t)d.n_frames + f], d.frame_twc[2 * (size_t)d.n_frames + f]};
      const double bp[3] = {(d.obs_uv[o] - d.cx) * d.fx_inv, (d.obs_uv[(size_t)d.n_obs + o] - d.cy) * d.fy_inv, 1.0};
      double b[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) b[r] = R[3 * r] * bp[0] + R[3 * r + 1] * bp[1] + R[3 * r + 2] * bp[2];
      const double inv_n2 = 1.0 / (b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
      const double bc = b[0] * c[0] + b[1] * c[1] + b[2] * c[2];
      const double bn[3] = {b[0] * inv_n2, b[1] * inv_n2, b[2] * inv_n2};
      BBt[0] += bn[0] * b[0];
      BBt[1] += bn[0] * b[1];
      BBt[2] += bn[0] * b[2];
      BBt[3] += bn[1] * b[1];
      BBt[4] += bn[1] * b[2];
      BBt[5] += bn[2] * b[2];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        csum[r] += c[r];
        bbc[r] += bn[r] * bc;
      }
    }
    const double n = (double)N;
    double A[3][3] = {{n - BBt[0], -BBt[1], -BBt[2]}, {-BBt[1], n - BBt[3], -BBt[4]}, {-BBt[2], -BBt[4], n - BBt[5]}};
    const double rhs[3] = {csum[0] - bbc[0], csum[1] - bbc[1], csum[2] - bbc[2]};
    ok = qr3_solve(A, rhs, 1e-5, x);
  }
  d.out_ok[i] = ok ? 1 : 0;
  if (ok) {
    d.out_xyz[i] = x[0];
    d.out_xyz[(size_t)d.n_points + i] = x[1];
    d.out_xyz[2 * (size_t)d.n_points + i] = x[2];
  }
  const unsigned m = __ballot_sync(__activemask(), ok);
  if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1) && m) atomicAdd(d.n_done, __popc(m));
}

} // namespace ba
