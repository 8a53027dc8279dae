// band_solver.cuh — Cholesky factorisation + solve of a banded reduced camera system in ONE launch.
//
// A visual-odometry chain without loop closures couples only keyframes within `band` positions of
// each other, so the reduced camera system S (n = 6 x poses) is banded. Cut into tiles of t >= 6 * band
// unknowns it is block-tridiagonal, S = U^T U with U block-bidiagonal and no fill outside the tiles.
// The chain of tiles is inherently sequential, and through library calls (potrf / trsm / syrk / gemv /
// trsv per tile: ~220 launches of tiny kernels per factorisation) it is pure launch latency: 14 ms for
// the 11 994-unknown system of the C5 benchmark. Here one CTA walks the chain with the current tile
// row [ D_k | E_k | b_k ] in shared memory:
//   forward  : right-looking block Cholesky with 6 x 6 pivots on the augmented tile row, which yields
//              U_kk, U_k,k+1 = U_kk^-T E_k and y_k = U_kk^-T b_k at once; then the Schur update
//              D_k+1 -= U_k,k+1^T U_k,k+1, b_k+1 -= U_k,k+1^T y_k directly in HBM (L2-resident band);
//   backward : x_k = U_kk^-1 (y_k - U_k,k+1 x_k+1).
// Status: correct (tests/test_global_gpu.py) but NOT the default -- measured 16.8 ms per factorisation + solve on
// the C5 system against 14.2 ms for the library-call chain: a single CTA pays ~126 us of dependent
// shared-memory / FP64 latency per tile. Selected with RSPL_BA_BAND_FUSED=1.
// Storage is the dense path's: A row-major n x n, upper band valid (the strictly lower part is scratch).
// Fails (info = 1) iff a pivot <= 0, the rule of g2o's LinearSolverEigen.
#pragma once

#include <cuda_runtime.h>

namespace ba {

constexpr int BAND_T_MAX = 96;                     // unknowns per tile (16 poses)
constexpr int BAND_LDM = 2 * BAND_T_MAX + 1;       // augmented row: D (t) | E (t) | rhs (1); odd => conflict-free columns
constexpr int BAND_THREADS = 256;
constexpr size_t BAND_SMEM = sizeof(double) * ((size_t)BAND_T_MAX * BAND_LDM + BAND_T_MAX + 8) + 16;

__global__ void __launch_bounds__(BAND_THREADS) k_band_chol_solve(double* __restrict__ A, double* __restrict__ rhs, int n,
                                                                  int t, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char band_smem_raw[];
  double* M = reinterpret_cast<double*>(band_smem_raw);  // [t][BAND_LDM]
  double* xn = M + (size_t)BAND_T_MAX * BAND_LDM;        // [t] x of the next tile (backward sweep)
  int* s_fail = reinterpret_cast<int*>(xn + BAND_T_MAX);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (n + t - 1) / t;
  if (tid == 0) *s_fail = 0;
  __syncthreads();

  // ---------------------------------------------------------------- forward sweep
  for (int k = 0; k < T; ++k) {
    const int o = k * t;
    const int tk = (o + t <= n) ? t : n - o;                       // rows of this tile (multiple of 6)
    const int tn = (k + 1 < T) ? ((o + 2 * t <= n) ? t : n - o - t) : 0; // columns of E_k
    const int mt = tk + tn + 1;                                    // augmented width
    for (int r = warp; r < tk; r += BAND_THREADS / 32) {
      const double* src = A + (size_t)(o + r) * n + o;
#pragma unroll 6
      for (int c = lane; c < tk + tn; c += 32) M[r * BAND_LDM + c] = src[c];
    }
    for (int r = tid; r < tk; r += BAND_THREADS) M[r * BAND_LDM + tk + tn] = rhs[o + r];
    __syncthreads();
    for (int r0 = 0; r0 < tk; r0 += 6) {
      // 6 x 6 upper Cholesky of the pivot block, computed redundantly by every thread in registers (a single
      // thread working through shared memory costs ~3 us per block: sqrt / divide latency), reciprocal pivots via rsqrt
      double U[6][6], iu[6];
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int c = p; c < 6; ++c) U[p][c] = M[(r0 + p) * BAND_LDM + r0 + c];
      bool bad = false;
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        double dgl = U[p][p];
#pragma unroll
        for (int q = 0; q < p; ++q) dgl -= U[q][p] * U[q][p];
        if (!(dgl > 0.0)) bad = true;
        const double r = rsqrt(dgl);
        iu[p] = r;
        U[p][p] = dgl * r;
#pragma unroll
        for (int c = p + 1; c < 6; ++c) {
          double v = U[p][c];
#pragma unroll
          for (int q = 0; q < p; ++q) v -= U[q][p] * U[q][c];
          U[p][c] = v * r;
        }
      }
      // row panel: columns right of the pivot block, X <- U_jj^-T X (one thread per column)
      const int c_first = r0 + 6;
      for (int c = c_first + tid; c < mt; c += BAND_THREADS) {
        double z[6];
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          double v = M[(r0 + p) * BAND_LDM + c];
#pragma unroll
          for (int q = 0; q < p; ++q) v -= U[q][p] * z[q];
          z[p] = v * iu[p];
        }
#pragma unroll
        for (int p = 0; p < 6; ++p) M[(r0 + p) * BAND_LDM + c] = z[p];
      }
      __syncthreads();
      if (tid == 0) { // write the factor of the pivot block back (after the barrier: everyone has read the block)
        if (bad) *s_fail = 1;
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
          for (int c = p; c < 6; ++c) M[(r0 + p) * BAND_LDM + r0 + c] = U[p][c];
      }
      // trailing update of the rows below: M[i][c] -= sum_p X[p][i] X[p][c]; a thread keeps its 4 columns of
      // the panel in registers and walks down the rows (the part left of the diagonal is scratch)
      const int ng = (mt - c_first + 3) >> 2; // column groups of 4 (<= 47)
      if (c_first < tk) {
        const int nrs = BAND_THREADS / ng; // row slots: thread = (column group, row slot)
        const int g = tid % ng, rs = tid / ng;
        if (rs < nrs) {
          const int c0 = c_first + 4 * g;
          double xc[6][4];
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) xc[p][q] = (c0 + q < mt) ? M[(r0 + p) * BAND_LDM + c0 + q] : 0.0;
          for (int i = c_first + rs; i < tk; i += nrs) {
            double xi[6];
#pragma unroll
            for (int p = 0; p < 6; ++p) xi[p] = M[(r0 + p) * BAND_LDM + i];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (c0 + q < mt) {
                double v = M[i * BAND_LDM + c0 + q];
#pragma unroll
                for (int p = 0; p < 6; ++p) v -= xi[p] * xc[p][q];
                M[i * BAND_LDM + c0 + q] = v;
              }
            }
          }
        }
      }
      __syncthreads();
    }
    // M now holds U_kk (upper of the D part), U_k,k+1 (E part), y_k (last column): store for the backward sweep
    for (int r = warp; r < tk; r += BAND_THREADS / 32) {
      double* dst = A + (size_t)(o + r) * n + o;
#pragma unroll 6
      for (int c = lane; c < tk + tn; c += 32) dst[c] = M[r * BAND_LDM + c];
    }
    for (int r = tid; r < tk; r += BAND_THREADS) rhs[o + r] = M[r * BAND_LDM + tk + tn];
    // Schur update of the next tile in HBM: D_k+1 -= E^T E (upper 6 x 6 blocks), b_k+1 -= E^T y
    if (tn > 0) {
      const int nb = tn / 6;
      const int nblk = nb * (nb + 1) / 2;
      for (int bi = tid; bi < nblk; bi += BAND_THREADS) {
        int br = 0, rem = bi; // (br, bc), br <= bc
        while (rem >= nb - br) {
          rem -= nb - br;
          ++br;
        }
        const int bc = br + rem;
        double acc[6][6];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b2 = 0; b2 < 6; ++b2) acc[a][b2] = 0.0;
        for (int p = 0; p < tk; ++p) {
          const double* row = M + p * BAND_LDM + tk;
          double ea[6], eb[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            ea[a] = row[6 * br + a];
            eb[a] = row[6 * bc + a];
          }
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 6; ++b2) acc[a][b2] += ea[a] * eb[b2];
        }
        double* dst = A + (size_t)(o + t + 6 * br) * n + o + t + 6 * bc;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b2 = 0; b2 < 6; ++b2) dst[(size_t)a * n + b2] -= acc[a][b2];
      }
      for (int c = tid; c < tn; c += BAND_THREADS) {
        double v = 0.0;
        for (int p = 0; p < tk; ++p) v += M[p * BAND_LDM + tk + c] * M[p * BAND_LDM + tk + tn];
        rhs[o + t + c] -= v;
      }
    }
    __syncthreads(); // (global writes of this CTA are visible to its own later reads after the barrier)
  }

  // ---------------------------------------------------------------- backward sweep
  for (int k = T - 1; k >= 0; --k) {
    const int o = k * t;
    const int tk = (o + t <= n) ? t : n - o;
    const int tn = (k + 1 < T) ? ((o + 2 * t <= n) ? t : n - o - t) : 0;
    for (int r = warp; r < tk; r += BAND_THREADS / 32) {
      const double* src = A + (size_t)(o + r) * n + o;
#pragma unroll 6
      for (int c = lane; c < tk + tn; c += 32) M[r * BAND_LDM + c] = src[c];
    }
    __syncthreads();
    // v = y_k - U_k,k+1 x_k+1 (one thread per row), kept in the last column
    for (int r = tid; r < tk; r += BAND_THREADS) {
      double v = rhs[o + r];
      for (int c = 0; c < tn; ++c) v -= M[r * BAND_LDM + tk + c] * xn[c];
      M[r * BAND_LDM + tk + tn] = v;
    }
    __syncthreads();
    if (warp == 0) {
      // U_kk x = v from the bottom, column-oriented: a lane owns rows lane, lane + 32, lane + 64; once x[r] is
      // known every lane removes column r from its rows (no reduction, one shuffle per row)
      double v[3], idg[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int r = lane + 32 * q;
        v[q] = r < tk ? M[r * BAND_LDM + tk + tn] : 0.0;
        idg[q] = r < tk ? 1.0 / M[r * BAND_LDM + r] : 0.0;
      }
      for (int r = tk - 1; r >= 0; --r) {
        const int q_own = r >> 5, l_own = r & 31;
        double xr = 0.0;
#pragma unroll
        for (int q = 0; q < 3; ++q)
          if (q == q_own) xr = v[q] * idg[q];
        xr = __shfl_sync(0xffffffffu, xr, l_own);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int rr = lane + 32 * q;
          if (rr == r) v[q] = xr;
          else if (rr < r) v[q] -= M[rr * BAND_LDM + r] * xr;
        }
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int r = lane + 32 * q;
        if (r < tk) M[r * BAND_LDM + tk + tn] = v[q];
      }
    }
    __syncthreads();
    for (int r = tid; r < tk; r += BAND_THREADS) {
      const double x = M[r * BAND_LDM + tk + tn];
      rhs[o + r] = x;
      xn[r] = x;
    }
    __syncthreads();
  }
  if (tid == 0) *info = *s_fail;
}

} // namespace ba
