// dense_chol.cuh — hand-written dense Cholesky factorisation + solve in HBM for reduced camera systems that are
// neither small (<= ~26 free poses: kb_solve, shared memory) nor banded chains (bcr_solver.cuh): local windows of
// 27+ keyframes and global problems with loop closures. Replaces the cuSOLVER potrf / potrs and cuBLAS trsm / syrk
// calls of round 1 (reference counterpart: g2o's LinearSolverEigen, set up at
// /root/reference/src/g2o_optimization/g2o_optimization.cc:26-36; pivot rule of SURVEY 9.11).
//
// Storage: the row-major upper triangle kb_assemble_dense writes == the column-major LOWER triangle, leading
// dimension n; element (r, c), r >= c, lives at A[c * n + r]. n is a multiple of 6.
// Right-looking blocked algorithm, panel width DC_NB = 48 columns, two launches per panel:
//   dc_panel : every CTA factorises the 48 x 48 diagonal block in shared memory (redundantly: it is the serial
//              part, and recomputing it saves a launch and a round trip through HBM); CTA 0 keeps the factor in a
//              side buffer, the others solve X L_kk^T = A_panel for 256 rows each (one thread per row, the row in
//              registers, coalesced column-major traffic).
//   dc_update: A_ij -= L_i L_j^T on 96 x 96 tiles of the trailing lower triangle, 6 x 6 register tiles over two
//              [48][96] shared-memory slabs (12 shared loads per 36 FMA; conflict-free 16-byte reads); one extra
//              CTA copies the diagonal factor into place.
//   dc_solve : forward and backward substitution by one CTA of 1024 threads with the right-hand side in shared
//              memory (block diagonal solves by one warp; panel products: a thread per row forward, a warp per
//              column backward, both coalesced, fixed summation order).
// Every sum has a fixed owner and order: results are bitwise reproducible.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bcr_solver.cuh"

namespace ba {

constexpr int DC_NB = 48;      // panel width (multiple of 6)
constexpr int DC_TILE = 96;    // trailing-update tile
constexpr int DC_THREADS = 256;
constexpr int DC_SOLVE_THREADS = 1024;

struct DenseChol {
  double* A;     // [n][n] lower triangle, column-major
  double* Ld;    // [n][DC_NB] diagonal factors of the panels (row r of panel k at Ld[(k0 + r) * DC_NB + c])
  double* dinv;  // [n] reciprocal diagonal of L
  double* rhs;   // [n] right-hand side in, solution out
  int* info;     // != 0: a pivot <= 0 (the factorisation failed; the solution is garbage)
  int n;
};

__host__ __device__ constexpr size_t dc_panel_smem(int kb) { return sizeof(double) * ((size_t)kb * (kb + 1) + kb) + 16; }
constexpr size_t DC_UPDATE_SMEM = sizeof(double) * 2 * DC_NB * DC_TILE;

// grid 1 + ceil((n - k0 - kb) / DC_THREADS)
__global__ void __launch_bounds__(DC_THREADS) dc_panel(const __grid_constant__ DenseChol s, int k0) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  const int n = s.n, kb = n - k0 < DC_NB ? n - k0 : DC_NB, ld = kb + 1, tid = threadIdx.x;
  double* Ls = reinterpret_cast<double*>(dc_smem);
  double* dinv = Ls + (size_t)kb * ld;
  int* s_fail = reinterpret_cast<int*>(dinv + kb);
  if (tid == 0) *s_fail = 0;
  for (int idx = tid; idx < kb * kb; idx += DC_THREADS) { // column c, row r (r fastest: coalesced)
    const int c = idx / kb, r = idx - c * kb;
    if (r >= c) Ls[r * ld + c] = s.A[(size_t)(k0 + c) * n + k0 + r];
  }
  __syncthreads();
  bcr_cta_cholesky(Ls, ld, kb, dinv, s_fail);
  if (*s_fail) { // (uniform) the solve is rejected as a whole; the later panels run on garbage that nobody uses
    if (tid == 0 && blockIdx.x == 0) atomicOr(s.info, 1);
  }
  if (blockIdx.x == 0) {
    for (int idx = tid; idx < kb * kb; idx += DC_THREADS) {
      const int r = idx / kb, c = idx - r * kb;
      s.Ld[(size_t)(k0 + r) * DC_NB + c] = c <= r ? Ls[r * ld + c] : 0.0;
    }
    for (int r = tid; r < kb; r += DC_THREADS) s.dinv[k0 + r] = dinv[r];
    return;
  }
  const int i = k0 + kb + (blockIdx.x - 1) * DC_THREADS + tid;
  if (i >= n) return;
  double x[DC_NB];
#pragma unroll
  for (int c = 0; c < DC_NB; ++c) {
    if (c < kb) {
      double v = s.A[(size_t)(k0 + c) * n + i];
#pragma unroll
      for (int q = 0; q < c; ++q) v -= x[q] * Ls[c * ld + q];
      x[c] = v * dinv[c];
      s.A[(size_t)(k0 + c) * n + i] = x[c];
    }
  }
}

// tile (ti, tj), tj <= ti, of the lower triangle from the linear index t
__device__ __forceinline__ void dc_tile_decode(int t, int& ti, int& tj) {
  int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  ti = r;
  tj = t - r * (r + 1) / 2;
}

// grid T (T + 1) / 2 + 1, T = ceil((n - k0 - kb) / DC_TILE); the last CTA moves the diagonal factor into place
__global__ void __launch_bounds__(DC_THREADS) dc_update(const __grid_constant__ DenseChol s, int k0, int n_tiles) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  const int n = s.n, kb = n - k0 < DC_NB ? n - k0 : DC_NB, tid = threadIdx.x;
  if ((int)blockIdx.x == n_tiles) {
    for (int idx = tid; idx < kb * kb; idx += DC_THREADS) {
      const int c = idx / kb, r = idx - c * kb;
      if (r >= c) s.A[(size_t)(k0 + c) * n + k0 + r] = s.Ld[(size_t)(k0 + r) * DC_NB + c];
    }
    return;
  }
  int ti, tj;
  dc_tile_decode(blockIdx.x, ti, tj);
  const int base = k0 + kb, r0 = base + ti * DC_TILE, c0 = base + tj * DC_TILE;
  double* Pi = reinterpret_cast<double*>(dc_smem); // [DC_NB][DC_TILE]: panel rows r0.. (k-major)
  double* Pj = Pi + DC_NB * DC_TILE;               // panel rows c0..
  for (int idx = tid; idx < DC_NB * DC_TILE; idx += DC_THREADS) {
    const int kk = idx / DC_TILE, r = idx - kk * DC_TILE;
    const bool in_k = kk < kb;
    Pi[idx] = (in_k && r0 + r < n) ? s.A[(size_t)(k0 + kk) * n + r0 + r] : 0.0;
    Pj[idx] = (in_k && c0 + r < n) ? s.A[(size_t)(k0 + kk) * n + c0 + r] : 0.0;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  double acc[6][6];
#pragma unroll
  for (int u = 0; u < 6; ++u)
#pragma unroll
    for (int v = 0; v < 6; ++v) acc[u][v] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < DC_NB; ++kk) {
    const double2* pa = reinterpret_cast<const double2*>(Pi + kk * DC_TILE + ty * 6);
    const double2* pb = reinterpret_cast<const double2*>(Pj + kk * DC_TILE + tx * 6);
    const double2 a0 = pa[0], a1 = pa[1], a2 = pa[2], b0 = pb[0], b1 = pb[1], b2 = pb[2];
    const double a[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y}, b[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
    for (int u = 0; u < 6; ++u)
#pragma unroll
      for (int v = 0; v < 6; ++v) acc[u][v] += a[u] * b[v];
  }
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    const int c = c0 + tx * 6 + v;
    if (c >= n) continue;
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int r = r0 + ty * 6 + u;
      if (r < n && r >= c) s.A[(size_t)c * n + r] -= acc[u][v];
    }
  }
}

// L y = b, L^T x = y; one CTA. xs: the right-hand side in shared memory when it fits (use_smem), else s.rhs itself.
__global__ void __launch_bounds__(DC_SOLVE_THREADS) dc_solve(const __grid_constant__ DenseChol s, int use_smem) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  const int n = s.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* x = use_smem ? reinterpret_cast<double*>(dc_smem) : s.rhs;
  if (use_smem) {
    for (int i = tid; i < n; i += DC_SOLVE_THREADS) x[i] = s.rhs[i];
  }
  __syncthreads();
  // ---- forward, right-looking: y_k = L_kk^-1 b_k (warp 0, column-oriented), then b_r -= L_rk y_k (everyone)
  for (int k0 = 0; k0 < n; k0 += DC_NB) {
    const int kb = n - k0 < DC_NB ? n - k0 : DC_NB;
    if (warp == 0) {
      for (int i = 0; i < kb; ++i) {
        const double yi = x[k0 + i] * s.dinv[k0 + i];
        __syncwarp();
        if (lane == 0) x[k0 + i] = yi;
        for (int j = i + 1 + lane; j < kb; j += 32) x[k0 + j] -= s.A[(size_t)(k0 + i) * n + k0 + j] * yi;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int r = k0 + kb + tid; r < n; r += DC_SOLVE_THREADS) {
      double v = x[r];
      for (int c = 0; c < kb; ++c) v -= s.A[(size_t)(k0 + c) * n + r] * x[k0 + c];
      x[r] = v;
    }
    __syncthreads();
  }
  // ---- backward, left-looking: y_k -= L_(r>k),k^T x_r (everyone, fixed-order reduction), then x_k = L_kk^-T y_k
  const int nblk = (n + DC_NB - 1) / DC_NB;
  for (int kbi = nblk - 1; kbi >= 0; --kbi) {
    const int k0 = kbi * DC_NB, kb = n - k0 < DC_NB ? n - k0 : DC_NB;
    if (k0 + kb < n) {
      // a warp per column of the panel: lanes stride the rows below the block (coalesced), fixed butterfly
      for (int c = warp; c < kb; c += DC_SOLVE_THREADS / 32) {
        const double* col = s.A + (size_t)(k0 + c) * n;
        double p = 0.0;
        for (int r = k0 + kb + lane; r < n; r += 32) p += col[r] * x[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == 0) x[k0 + c] -= p;
      }
      __syncthreads();
    }
    if (warp == 0) {
      for (int i = kb - 1; i >= 0; --i) {
        const double xi = x[k0 + i] * s.dinv[k0 + i];
        __syncwarp();
        if (lane == 0) x[k0 + i] = xi;
        for (int j = lane; j < i; j += 32) x[k0 + j] -= s.A[(size_t)(k0 + j) * n + k0 + i] * xi;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (use_smem) {
    for (int i = tid; i < n; i += DC_SOLVE_THREADS) s.rhs[i] = x[i];
  }
}

} // namespace ba
