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
//              registers, coalesced column-major traffic). The right-hand side rides along (forward substitution).
//   dc_update: A_ij -= L_i L_j^T on 96 x 96 tiles of the trailing lower triangle, 6 x 6 register tiles (strided by
//              16 so that the tile's read-modify-write is coalesced) over two [48][96] shared-memory slabs filled by
//              cp.async; one extra CTA copies the diagonal factor into place.
//   dc_solve : backward substitution by one CTA of 1024 threads with the right-hand side in shared memory (diagonal
//              blocks staged in shared memory and solved by one warp, panel products a thread per column).
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
  double* y;     // [n] L^-1 b, written panel by panel (not into rhs: the other CTAs of the panel still read b_k there)
  int* info;     // != 0: a pivot <= 0 (the factorisation failed; the solution is garbage)
  int n;
};

__host__ __device__ constexpr size_t dc_panel_smem(int kb) { return sizeof(double) * ((size_t)kb * (kb + 1) + 2 * kb) + 16; }
constexpr size_t DC_UPDATE_SMEM = sizeof(double) * 2 * DC_NB * DC_TILE;
constexpr size_t DC_SOLVE_STATIC_SMEM = sizeof(double) * (DC_NB * (DC_NB + 1) + DC_NB) + 1024; // dc_solve: Lb, xb (+ slack)

// grid 1 + ceil((n - k0 - kb) / DC_THREADS). The right-hand side rides along (forward substitution): every CTA also
// solves y_k = L_kk^-1 b_k, and a row thread finishes with b_i -= L_i,k y_k.
__global__ void __launch_bounds__(DC_THREADS) dc_panel(const __grid_constant__ DenseChol s, int k0) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  const int n = s.n, kb = n - k0 < DC_NB ? n - k0 : DC_NB, ld = kb + 1, tid = threadIdx.x;
  double* Ls = reinterpret_cast<double*>(dc_smem);
  double* dinv = Ls + (size_t)kb * ld;
  double* yk = dinv + kb;
  int* s_fail = reinterpret_cast<int*>(yk + kb);
  if (tid == 0) *s_fail = 0;
  for (int idx = tid; idx < kb * kb; idx += DC_THREADS) { // column c, row r (r fastest: coalesced)
    const int c = idx / kb, r = idx - c * kb;
    if (r >= c) Ls[r * ld + c] = s.A[(size_t)(k0 + c) * n + k0 + r];
  }
  for (int r = tid; r < kb; r += DC_THREADS) yk[r] = s.rhs[k0 + r];
  __syncthreads();
  bcr_cta_cholesky(Ls, ld, kb, dinv, s_fail);
  if (*s_fail) { // (uniform) the solve is rejected as a whole; the later panels run on garbage that nobody uses
    if (tid == 0 && blockIdx.x == 0) atomicOr(s.info, 1);
  }
  if (tid < 32) { // y_k = L_kk^-1 b_k, column-oriented, one warp
    for (int i = 0; i < kb; ++i) {
      const double yi = yk[i] * dinv[i];
      __syncwarp();
      if (tid == 0) yk[i] = yi;
      for (int j = i + 1 + tid; j < kb; j += 32) yk[j] -= Ls[j * ld + i] * yi;
      __syncwarp();
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int idx = tid; idx < kb * kb; idx += DC_THREADS) {
      const int r = idx / kb, c = idx - r * kb;
      s.Ld[(size_t)(k0 + r) * DC_NB + c] = c <= r ? Ls[r * ld + c] : 0.0;
    }
    for (int r = tid; r < kb; r += DC_THREADS) {
      s.dinv[k0 + r] = dinv[r];
      s.y[k0 + r] = yk[r];
    }
    return;
  }
  const int i = k0 + kb + (blockIdx.x - 1) * DC_THREADS + tid;
  if (i >= n) return;
  double x[DC_NB];
#pragma unroll
  for (int c = 0; c < DC_NB; ++c) x[c] = c < kb ? s.A[(size_t)(k0 + c) * n + i] : 0.0;
  double bi = s.rhs[i];
#pragma unroll
  for (int c = 0; c < DC_NB; ++c) {
    if (c < kb) { // right-looking over the row held in registers
      const double xc = x[c] * dinv[c];
      x[c] = xc;
#pragma unroll
      for (int q = c + 1; q < DC_NB; ++q)
        if (q < kb) x[q] -= xc * Ls[q * ld + c];
      bi -= xc * yk[c];
    }
  }
#pragma unroll
  for (int c = 0; c < DC_NB; ++c)
    if (c < kb) s.A[(size_t)(k0 + c) * n + i] = x[c];
  s.rhs[i] = bi;
}

// tile (ti, tj), tj <= ti, of the lower triangle from the linear index t
__device__ __forceinline__ void dc_tile_decode(int t, int& ti, int& tj) {
  int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  ti = r;
  tj = t - r * (r + 1) / 2;
}

// grid T (T + 1) / 2 + 1, T = ceil((n - k0 - kb) / DC_TILE); the last CTA moves the diagonal factor into place.
// Thread (tx, ty) of the 16 x 16 layout owns rows tx + 16 u and columns ty + 16 v (u, v < 6): a half-warp touches
// 16 consecutive rows of one column, so the read-modify-write of the tile is coalesced in the column-major matrix, and
// the shared-memory reads are conflict-free (rows) or broadcasts (columns).
__global__ void __launch_bounds__(DC_THREADS, 2) dc_update(const __grid_constant__ DenseChol s, int k0, int n_tiles) {
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
  // slabs: 16-byte cp.async where the two doubles exist (n, r0, c0 are multiples of 6, so pairs never straddle n)
  for (int idx = tid; idx < DC_NB * (DC_TILE / 2); idx += DC_THREADS) {
    const int kk = idx / (DC_TILE / 2), r = 2 * (idx - kk * (DC_TILE / 2));
    const bool in_k = kk < kb;
    double* di = Pi + kk * DC_TILE + r;
    double* dj = Pj + kk * DC_TILE + r;
    if (in_k && r0 + r < n) cp_async16(di, s.A + (size_t)(k0 + kk) * n + r0 + r);
    else di[0] = di[1] = 0.0;
    if (in_k && c0 + r < n) cp_async16(dj, s.A + (size_t)(k0 + kk) * n + c0 + r);
    else dj[0] = dj[1] = 0.0;
  }
  cp_async_commit();
  const int tx = tid & 15, ty = tid >> 4;
  // the tile itself, requested while the slabs are in flight
  double acc[6][6];
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    const int c = c0 + ty + 16 * v;
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int r = r0 + tx + 16 * u;
      acc[u][v] = (c < n && r < n && r >= c) ? s.A[(size_t)c * n + r] : 0.0;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll 4
  for (int kk = 0; kk < DC_NB; ++kk) {
    double a[6], b[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) a[u] = Pi[kk * DC_TILE + tx + 16 * u];
#pragma unroll
    for (int v = 0; v < 6; ++v) b[v] = Pj[kk * DC_TILE + ty + 16 * v];
#pragma unroll
    for (int u = 0; u < 6; ++u)
#pragma unroll
      for (int v = 0; v < 6; ++v) acc[u][v] -= a[u] * b[v];
  }
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    const int c = c0 + ty + 16 * v;
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int r = r0 + tx + 16 * u;
      if (c < n && r < n && r >= c) s.A[(size_t)c * n + r] = acc[u][v];
    }
  }
}

// L^T x = y (the forward substitution rode along with the factorisation); one CTA, right-looking from the last block:
// x_k = L_kk^-T y_k with the diagonal block staged in shared memory (one warp, 48 dependent steps), then
// y_j -= sum_{r in block} L[r][j] x[r] for every column j to the left (a thread per column: 48 contiguous rows).
// x: the right-hand side in shared memory when it fits (use_smem), else s.rhs itself.
__global__ void __launch_bounds__(DC_SOLVE_THREADS) dc_solve(const __grid_constant__ DenseChol s, int use_smem) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  __shared__ double Lb[DC_NB][DC_NB + 1];
  __shared__ double xb[DC_NB];
  const int n = s.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* x = use_smem ? reinterpret_cast<double*>(dc_smem) : s.rhs;
  for (int i = tid; i < n; i += DC_SOLVE_THREADS) x[i] = s.y[i];
  const int nblk = (n + DC_NB - 1) / DC_NB;
  for (int kbi = nblk - 1; kbi >= 0; --kbi) {
    const int k0 = kbi * DC_NB, kb = n - k0 < DC_NB ? n - k0 : DC_NB;
    for (int idx = tid; idx < kb * kb; idx += DC_SOLVE_THREADS) {
      const int c = idx / kb, r = idx - c * kb;
      if (r >= c) Lb[r][c] = s.A[(size_t)(k0 + c) * n + k0 + r];
    }
    __syncthreads(); // (also orders the column updates of the previous block before the reads of x below)
    if (warp == 0) {
      for (int i = kb - 1; i >= 0; --i) {
        const double xi = x[k0 + i] * s.dinv[k0 + i];
        __syncwarp();
        if (lane == 0) {
          x[k0 + i] = xi;
          xb[i] = xi;
        }
        for (int j = lane; j < i; j += 32) x[k0 + j] -= Lb[i][j] * xi;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int j = tid; j < k0; j += DC_SOLVE_THREADS) {
      const double* col = s.A + (size_t)j * n + k0;
      double sum = 0.0;
#pragma unroll 8
      for (int r = 0; r < kb; ++r) sum += col[r] * xb[r];
      x[j] -= sum;
    }
  }
  __syncthreads();
  if (use_smem) {
    for (int i = tid; i < n; i += DC_SOLVE_THREADS) s.rhs[i] = x[i];
  }
}

} // namespace ba
