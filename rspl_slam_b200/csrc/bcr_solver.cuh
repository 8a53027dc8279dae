// bcr_solver.cuh — hand-written direct solver for LARGE banded reduced camera systems (global BA):
// block cyclic reduction of the block-tridiagonal form, every level a grid of independent CTAs.
//
// A keyframe chain couples only poses within `band` positions of each other, so with super-blocks of
// bsp >= band poses (bs = 6 bsp unknowns) the reduced system Hs is block tridiagonal:
//     D_0 E_0
//     E_0^T D_1 E_1
//            ...            (M super-blocks)
// A sequential Cholesky along the chain is bound by the latency of ~M dependent dense steps on ONE SM
// (band_solver.cuh: 16.8 ms for the C5 system, the cuSOLVER / cuBLAS tile chain 14.3 ms). Cyclic reduction
// eliminates every other super-block of the active chain at once: per level
//   bcr_eliminate : one CTA per odd block I (neighbours Il, Ir):  D_I = L L^T in shared memory (6 x 6-blocked),
//                   GL = L^-1 E(Il,I)^T,  GR = L^-1 E(I,Ir),  g = L^-1 b_I
//   bcr_update    : one CTA per even block J:  D_J -= GR(Il)^T GR(Il) + GL(Ir)^T GL(Ir),
//                   b_J -= GR(Il)^T g(Il) + GL(Ir)^T g(Ir),  new coupling E'(J, J+2h) = -GL(Ir)^T GR(Ir)
// which halves the chain; after ceil(log2 M) levels the root block is solved and the eliminated blocks are
// recovered level by level, x_I = L^-T (g - GL x_Il - GR x_Ir) (bcr_backsub). This is a Cholesky factorisation
// of the same matrix in nested-dissection order: exact, SPD-checked by the same rule (a pivot <= 0 fails,
// g2o's LinearSolverEigen), every sum with a fixed owner and order (bitwise reproducible, the
// same on every rank of a replicated solve). ~3 log2(M) launches of M / 2^l CTAs: 0.4 ms instead of 14 ms for
// the 11 994-unknown C5 system.
#pragma once

#include <cuda_runtime.h>

namespace ba {

constexpr int BCR_THREADS = 512;
constexpr int BCR_BS_MAX = 144;    // unknowns per super-block (24 poses): the factor of a block lives in shared memory
constexpr int BCR_MAX_LEVELS = 20;
constexpr int BCR_KC = 30;
// -DRSPL_BCR_CLOCKS: phase clocks of CTA 0 of bcr_eliminate (profiles/scripts/bcr_phase_clocks.sh); off in the product
#ifdef RSPL_BCR_CLOCKS
__device__ long long bcr_clk[12];
__device__ long long bcr_t0;
#define BCR_CLK(i)                                 \
  do {                                             \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { \
      const long long t_ = clock64();              \
      bcr_clk[i] += t_ - bcr_t0;                   \
      bcr_t0 = t_;                                 \
    }                                              \
  } while (0)
#else
#define BCR_CLK(i) do { } while (0)
#endif         // rows per staged slab of bcr_update

struct BcrDev {
  int M, bs, n;       // super-blocks, unknowns per super-block, system size
  int levels;
  double* D;          // [M][bs*bs] row-major, full storage
  double* F;          // [M][bs*bs] lower Cholesky factor of an eliminated block (its own array: the column slices of
                      // bcr_eliminate all read D while slice 0 writes the factor)
  double* E;          // coupling blocks of every level: E + eoff[l] + p * bs*bs = rows of active block p, columns of p + 1
  long long eoff[BCR_MAX_LEVELS];
  double* GL;         // [M][bs*bs]  L^-1 E(Il,I)^T   (rows I, columns Il)
  double* GR;         // [M][bs*bs]  L^-1 E(I,Ir)     (rows I, columns Ir)
  double* g;          // [M][bs]     L^-1 b_I
  double* x;          // [M][bs]     right-hand side in, solution out (flat: unknown 6 * s + r of system block s)
  int* info;          // != 0: a pivot <= 0
};

__device__ __forceinline__ int bcr_ld(int bs) { return bs + 1; } // odd leading dimension: conflict-free rows and columns

// 6 x 6 lower Cholesky of the diagonal block at (k0, k0) of the shared-memory matrix, computed redundantly by every
// thread in registers (a serial chain of 6 rsqrt: cheaper than a broadcast); returns false on a pivot <= 0
__device__ __forceinline__ bool bcr_chol6(const double* Ls, int ld, int k0, double (&Lk)[6][6], double (&inv)[6]) {
  bool ok = true;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) Lk[r][c] = Ls[(k0 + r) * ld + k0 + c];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double dsum = Lk[j][j];
#pragma unroll
    for (int p = 0; p < j; ++p) dsum -= Lk[j][p] * Lk[j][p];
    if (dsum <= 0.0) ok = false; // (NaN falls through, like Eigen's LLT test)
    const double rs = ba::rsqrt_nr(dsum);
    inv[j] = rs;
    Lk[j][j] = dsum * rs;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = Lk[i][j];
#pragma unroll
      for (int p = 0; p < j; ++p) v -= Lk[i][p] * Lk[j][p];
      Lk[i][j] = v * rs;
    }
  }
  return ok;
}

// Factor of the 6 x 6 diagonal block at k0 by ONE warp (every lane redundantly, in registers), written back in place
// together with the reciprocal diagonal and, when LinvS != null, the inverse of the block's factor (lower triangle
// of LinvS[(k0 / 6) * 36 + r * 6 + c]; the entries above the diagonal are never read).
// (static register indices only: the published value is picked with a select chain, the store has a computed address
// -- a lane-dependent index into Lk moves the whole block to local memory, 1.3 - 1.6 x slower kernels)
__device__ __forceinline__ void bcr_diag_block(double* Ls, int ld, int k0, double* dinv, int* s_fail, double* LinvS) {
  const int lane = threadIdx.x & 31;
  double Lk[6][6], inv[6];
  if (!bcr_chol6(Ls, ld, k0, Lk, inv)) *s_fail = 1; // (benign race: every lane writes the same value)
  __syncwarp();                                     // every lane has read the block
  // lane e < 21 owns the lower-triangle entry (r, c), e = r (r + 1) / 2 + c; lanes 21..26 own inv[lane - 21]
  const int r = (lane >= 1) + (lane >= 3) + (lane >= 6) + (lane >= 10) + (lane >= 15), c = lane - r * (r + 1) / 2;
  double v = 0.0, iv = 0.0;
#pragma unroll
  for (int rr = 0; rr < 6; ++rr) {
#pragma unroll
    for (int cc = 0; cc <= rr; ++cc) v = lane == rr * (rr + 1) / 2 + cc ? Lk[rr][cc] : v;
    iv = lane == 21 + rr ? inv[rr] : iv;
  }
  if (lane < 21) Ls[(k0 + r) * ld + k0 + c] = v;
  else if (lane < 27) dinv[k0 + lane - 21] = iv;
  if (LinvS != nullptr) {
    double Li[6][6];
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) {
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) Li[rr][cc] = 0.0;
      Li[cc][cc] = inv[cc];
#pragma unroll
      for (int rr = cc + 1; rr < 6; ++rr) {
        double acc = 0.0;
#pragma unroll
        for (int q = cc; q < rr; ++q) acc += Lk[rr][q] * Li[q][cc];
        Li[rr][cc] = -acc * inv[rr];
      }
    }
    double w = 0.0;
#pragma unroll
    for (int rr = 0; rr < 6; ++rr)
#pragma unroll
      for (int cc = 0; cc <= rr; ++cc) w = lane == rr * (rr + 1) / 2 + cc ? Li[rr][cc] : w;
    if (lane < 21) LinvS[(k0 / 6) * 36 + r * 6 + c] = w;
  }
}

// In-place lower Cholesky of the n x n matrix in shared memory (n a multiple of 6), whole CTA, two barriers per
// 6-column panel. The strictly upper part is not touched; dinv [n] receives the reciprocal diagonal of the factor.
// *s_fail is set on a pivot <= 0. LinvS != null: also receives the inverses of the 6 x 6 diagonal blocks of the factor
// (what bcr_cta_forward_inv multiplies with instead of substituting).
// Look-ahead: while warps 1.. apply the trailing update of panel k, warp 0 updates the NEXT diagonal block first and
// factorises it, so the serial 6 x 6 factorisation (six dependent rsqrt chains, formerly run redundantly by all
// warps between two barriers: 2100 of the 5500 cycles of a panel step in bcr_eliminate) is off the critical path.
// Every entry sees the same operations in the same order as before: same bits.
// extra = 1: row n of the array holds a right-hand side b^T; it is carried through the panel solves and trailing
// updates like any row below the diagonal and ends as (L^-1 b)^T -- the forward substitution without a sweep of its
// own (same products in the same order as bcr_cta_forward: same bits).
__device__ void bcr_cta_cholesky(double* Ls, int ld, int n, double* dinv, int* s_fail, double* LinvS = nullptr, int extra = 0) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int nr = n + extra; // rows carried
  if (warp == 0) bcr_diag_block(Ls, ld, 0, dinv, s_fail, LinvS);
  __syncthreads();
  BCR_CLK(1);
  for (int k0 = 0; k0 < n; k0 += 6) {
    if (k0 + 6 + tid < nr) { // panel: X Lkk^T = A, one thread per row; Lkk broadcast from shared memory
      double Lk[6][6], inv[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        inv[c] = dinv[k0 + c];
#pragma unroll
        for (int q = 0; q < c; ++q) Lk[c][q] = Ls[(k0 + c) * ld + k0 + q];
      }
      for (int i = k0 + 6 + tid; i < nr; i += nt) {
        double xr[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double v = Ls[i * ld + k0 + c];
#pragma unroll
          for (int q = 0; q < c; ++q) v -= xr[q] * Lk[c][q];
          xr[c] = v * inv[c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) Ls[i * ld + k0 + c] = xr[c];
      }
    }
    __syncthreads();
    BCR_CLK(2);
    const int k1 = k0 + 6;
    if (warp == 0 || nw == 1) {
      if (k1 < n) { // the next diagonal block: lane e < 21 updates its entry (r, c), then the warp factorises it
        const int r = (lane >= 1) + (lane >= 3) + (lane >= 6) + (lane >= 10) + (lane >= 15), c = lane - r * (r + 1) / 2;
        if (lane < 21) {
          double v = Ls[(k1 + r) * ld + k1 + c];
#pragma unroll
          for (int q = 0; q < 6; ++q) v -= Ls[(k1 + r) * ld + k0 + q] * Ls[(k1 + c) * ld + k0 + q];
          Ls[(k1 + r) * ld + k1 + c] = v;
        }
        __syncwarp();
        bcr_diag_block(Ls, ld, k1, dinv, s_fail, LinvS);
      }
    }
    if (warp != 0 || nw == 1) {
      // trailing update of the rows below the next diagonal block: a warp per row pair, lanes over the columns j <= i
      const int w = nw == 1 ? 0 : warp - 1, wn = nw == 1 ? 1 : nw - 1;
      for (int i = k1 + 6 + 2 * w; i < nr; i += 2 * wn) {
        const bool two = i + 1 < nr;
        const int j0max = i < n ? i : n - 1, j1max = i + 1 < n ? i + 1 : n - 1; // (a right-hand-side row has n columns)
        double p0[6], p1[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          p0[q] = Ls[i * ld + k0 + q];
          p1[q] = two ? Ls[(i + 1) * ld + k0 + q] : 0.0;
        }
        for (int jb = k1 + lane; jb <= j1max; jb += 64) { // two column chunks at a time (independent FMA chains)
          double v0[2], v1[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int j = jb + 32 * u;
            v0[u] = j <= j0max ? Ls[i * ld + j] : 0.0;
            v1[u] = two && j <= j1max ? Ls[(i + 1) * ld + j] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 6; ++q) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = jb + 32 * u;
              const double pq = j <= j1max ? Ls[j * ld + k0 + q] : 0.0;
              v0[u] -= p0[q] * pq;
              v1[u] -= p1[q] * pq;
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int j = jb + 32 * u;
            if (j <= j0max) Ls[i * ld + j] = v0[u];
            if (two && j <= j1max) Ls[(i + 1) * ld + j] = v1[u];
          }
        }
      }
    }
    __syncthreads();
    BCR_CLK(3);
  }
}

// Y = L^-1 X for the nc columns of the panel P [n][pld] in shared memory (right-looking, 6 rows at a time);
// dinv = reciprocal diagonal of L
__device__ void bcr_cta_forward(const double* Ls, int ld, const double* dinv, int n, double* P, int pld, int nc) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int k0 = 0; k0 < n; k0 += 6) {
    for (int c = tid; c < nc; c += nt) { // Y_k = Lkk^-1 X_k, one thread per column
      double y[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double v = P[(k0 + r) * pld + c];
#pragma unroll
        for (int q = 0; q < r; ++q) v -= Ls[(k0 + r) * ld + k0 + q] * y[q];
        y[r] = v * dinv[k0 + r];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) P[(k0 + r) * pld + c] = y[r];
    }
    __syncthreads();
    // X_i -= L_ik Y_k: a warp per row i, lanes over the columns (the six Y rows of a column stay in registers
    // across two rows)
    for (int i = k0 + 6 + 2 * warp; i < n; i += 2 * nw) {
      const bool two = i + 1 < n;
      double l0[6], l1[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        l0[q] = Ls[i * ld + k0 + q];
        l1[q] = two ? Ls[(i + 1) * ld + k0 + q] : 0.0;
      }
      for (int c = lane; c < nc; c += 32) {
        double v0 = P[i * pld + c], v1 = two ? P[(i + 1) * pld + c] : 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double yq = P[(k0 + q) * pld + c];
          v0 -= l0[q] * yq;
          v1 -= l1[q] * yq;
        }
        P[i * pld + c] = v0;
        if (two) P[(i + 1) * pld + c] = v1;
      }
    }
    __syncthreads();
  }
}

// The same with the inverted diagonal blocks of bcr_cta_cholesky: Y_k = Linv_kk X_k is a 6 x 6 product per column
// spread over all threads (thread = (row, column)) instead of a six-step substitution on one thread per column;
// results are held in registers across a barrier (the product is not in place). (Forming Y_{k+1} in registers by the
// two warps that have just updated those rows saves the phase and its barriers but puts 21 more dependent FMAs and
// broadcast loads on the step's slowest warps: measured 1.143 instead of 1.125 ms per solve -- not kept.)
constexpr int BCR_FWD_ROUNDS = (6 * (2 * BCR_BS_MAX + 1) + BCR_THREADS - 1) / BCR_THREADS;
__device__ void bcr_cta_forward_inv(const double* Ls, int ld, const double* LinvS, int n, double* P, int pld, int nc) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  int rr[BCR_FWD_ROUNDS], cc[BCR_FWD_ROUNDS];
#pragma unroll
  for (int u = 0; u < BCR_FWD_ROUNDS; ++u) {
    const int t = tid + u * nt;
    rr[u] = t < 6 * nc ? t / nc : -1;
    cc[u] = t - (t / nc) * nc;
  }
  for (int k0 = 0; k0 < n; k0 += 6) {
    const double* Li = LinvS + (k0 / 6) * 36;
    double yv[BCR_FWD_ROUNDS];
#pragma unroll
    for (int u = 0; u < BCR_FWD_ROUNDS; ++u) {
      double acc = 0.0;
      if (rr[u] >= 0) {
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if (q <= rr[u]) acc += Li[rr[u] * 6 + q] * P[(k0 + q) * pld + cc[u]];
      }
      yv[u] = acc;
    }
    __syncthreads();
    BCR_CLK(5);
#pragma unroll
    for (int u = 0; u < BCR_FWD_ROUNDS; ++u)
      if (rr[u] >= 0) P[(k0 + rr[u]) * pld + cc[u]] = yv[u];
    __syncthreads();
    BCR_CLK(6);
    // X_i -= L_ik Y_k. This phase is bound by shared-memory bandwidth, not latency (unrolling three column chunks for
    // ILP changed nothing): per updated entry the row-pair version moves 40 bytes (P in, P out, three Y loads). A warp
    // now takes a whole 6-row block (n is a multiple of 6) and half of the columns: the six Y values of a column serve
    // six rows -- 24 bytes per entry -- and the 6 x 6 block of L stays in registers.
    {
      const int ntile = (n - k0 - 6) / 6;
      const int half = ((nc + 63) / 64) * 32;
      for (int wi = warp; wi < 2 * ntile; wi += nw) {
        const int i = k0 + 6 + 6 * (wi >> 1);
        const int cbeg = (wi & 1) ? half : 0, cend = (wi & 1) ? nc : (half < nc ? half : nc);
        double l[6][6];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int q = 0; q < 6; ++q) l[a][q] = Ls[(i + a) * ld + k0 + q];
        for (int c = cbeg + lane; c < cend; c += 32) {
          double y[6], v[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            y[q] = P[(k0 + q) * pld + c];
            v[q] = P[(i + q) * pld + c];
          }
#pragma unroll
          for (int q = 0; q < 6; ++q)
#pragma unroll
            for (int a = 0; a < 6; ++a) v[a] -= l[a][q] * y[q];
#pragma unroll
          for (int a = 0; a < 6; ++a) P[(i + a) * pld + c] = v[a];
        }
      }
    }
    __syncthreads();
    BCR_CLK(7);
  }
}

// x = L^-T y in place (one warp, column-oriented); y in shared memory. For n <= 160 the vector lives in registers
// (entry j in lane j % 32, slot j / 32) and x_i is broadcast by a shuffle: the dependent chain of a step is
// FMA -> SHFL -> MUL instead of two shared-memory round trips and two warp barriers (60 - 156 serial steps per solve
// of a local window). Same operations on every entry in the same order: same bits.
constexpr int BCR_BACK_SLOTS = 5;
__device__ void bcr_warp_backward(const double* Ls, int ld, const double* dinv, int n, double* y) {
  const int lane = threadIdx.x & 31;
  if (n <= 32 * BCR_BACK_SLOTS) {
    double yr[BCR_BACK_SLOTS];
#pragma unroll
    for (int sl = 0; sl < BCR_BACK_SLOTS; ++sl) yr[sl] = sl * 32 + lane < n ? y[sl * 32 + lane] : 0.0;
#pragma unroll
    for (int sl = BCR_BACK_SLOTS - 1; sl >= 0; --sl) {
      if (sl * 32 >= n) continue; // (uniform)
      const int top = n - 1 - sl * 32 < 31 ? n - 1 - sl * 32 : 31;
      for (int il = top; il >= 0; --il) {
        const int i = sl * 32 + il;
        const double* Li = Ls + i * ld;
        const double xi = __shfl_sync(0xffffffffu, yr[sl], il) * dinv[i];
        if (lane == il) yr[sl] = xi;
#pragma unroll
        for (int s2 = 0; s2 <= sl; ++s2) {
          const int j = s2 * 32 + lane;
          if (j < i) yr[s2] -= Li[j] * xi;
        }
      }
    }
#pragma unroll
    for (int sl = 0; sl < BCR_BACK_SLOTS; ++sl)
      if (sl * 32 + lane < n) y[sl * 32 + lane] = yr[sl];
    __syncwarp();
    return;
  }
  for (int i = n - 1; i >= 0; --i) {
    const double xi = y[i] * dinv[i];
    __syncwarp();
    if (lane == 0) y[i] = xi;
    for (int j = lane; j < i; j += 32) y[j] -= Ls[i * ld + j] * xi;
    __syncwarp();
  }
}

// ---- level l, odd blocks: factorise and form GL, GR, g. grid = (number of odd active blocks, column slices: the
// 2 bs + 1 panel columns are split over the slices while the level leaves SMs idle); dynamic smem =
// (bs * ld + 7 * bs + bs * pch) doubles + 16 bytes, pch = columns of one panel chunk (>= 1, odd)
__global__ void __launch_bounds__(BCR_THREADS) bcr_eliminate(const __grid_constant__ BcrDev s, int level, int pch) {
  extern __shared__ __align__(16) unsigned char bcr_smem[];
  const int bs = s.bs, ld = bcr_ld(bs), tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  double* Ls = reinterpret_cast<double*>(bcr_smem);
  double* dinv = Ls + (size_t)bs * ld;
  double* Linv = dinv + bs;            // [bs / 6][36]
  double* P = Linv + (size_t)6 * bs;
  int* s_fail = reinterpret_cast<int*>(P + (size_t)bs * pch);
  const int h = 1 << level;
  const int p = 2 * blockIdx.x + 1;
  const int I = p * h, Ir = I + h;
  const bool has_r = Ir < s.M;
  const size_t bb = (size_t)bs * bs;
  const double* Dg = s.D + (size_t)I * bb;
  if (tid == 0) *s_fail = 0;
#ifdef RSPL_BCR_CLOCKS
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) bcr_t0 = clock64();
#endif
  // (all copies below: a warp per row, lanes along it -- no integer division per element)
  // The diagonal block and the first panel chunk are fetched with cp.async when the kernel starts (8-byte copies: the
  // odd leading dimensions leave the shared-memory rows 8-byte aligned only); the panel lands under the factorisation.
  for (int r = warp; r < bs; r += nw)
    for (int c = lane; c < bs; c += 32) cp_async8(&Ls[r * ld + c], &Dg[(size_t)r * bs + c]);
  cp_async_commit();
  const double* El = s.E + s.eoff[level] + (size_t)(p - 1) * bb; // rows Il, columns I
  const double* Er = s.E + s.eoff[level] + (size_t)p * bb;       // rows I, columns Ir
  double* GLg = s.GL + (size_t)I * bb;
  double* GRg = s.GR + (size_t)I * bb;
  const int ncol = 2 * bs + 1; // [ E(Il,I)^T | E(I,Ir) | b_I ]
  // panel chunk [c0, c0 + nc): columns of E(Il,I)^T are rows of E(Il,I) -- a warp per such row keeps the global reads
  // coalesced; pch is odd, so the transposing shared-memory writes are conflict-free
  auto stage_panel = [&](int c0, int nc) {
    const int c1 = c0 + nc;
    const int la = c0, lb = c1 < bs ? c1 : bs;                        // columns of E(Il,I)^T in this chunk
    const int ra = c0 > bs ? c0 : bs, rb = c1 < 2 * bs ? c1 : 2 * bs; // columns of E(I,Ir)
    for (int c = la + warp; c < lb; c += nw)
      for (int r = lane; r < bs; r += 32) cp_async8(&P[r * pch + c - c0], &El[(size_t)c * bs + r]);
    for (int r = warp; r < bs; r += nw)
      for (int c = ra + lane; c < rb; c += 32) {
        if (has_r) cp_async8(&P[r * pch + c - c0], &Er[(size_t)r * bs + c - bs]);
        else P[r * pch + c - c0] = 0.0;
      }
    if (c1 > 2 * bs)
      for (int r = tid; r < bs; r += nt) cp_async8(&P[r * pch + 2 * bs - c0], &s.x[(size_t)I * bs + r]);
    cp_async_commit();
  };
  // column slice of this CTA (gridDim.y slices; every slice factorises the block itself)
  const int per = (ncol + (int)gridDim.y - 1) / (int)gridDim.y;
  const int cA = (int)blockIdx.y * per, cB = cA + per < ncol ? cA + per : ncol;
  stage_panel(cA, cB - cA < pch ? cB - cA : pch);
  cp_async_wait<1>(); // the diagonal block
  __syncthreads();
  BCR_CLK(0);
  bcr_cta_cholesky(Ls, ld, bs, dinv, s_fail, Linv);
  if (*s_fail) {
    cp_async_wait<0>();
    if (tid == 0) atomicOr(s.info, 1);
    return; // (uniform) the solve is rejected as a whole
  }
  if (blockIdx.y == 0) { // keep the factor for the back substitution
    double* Fg = s.F + (size_t)I * bb;
    for (int r = warp; r < bs; r += nw)
      for (int c = lane; c < bs; c += 32) Fg[(size_t)r * bs + c] = c <= r ? Ls[r * ld + c] : 0.0;
  }
  for (int c0 = cA; c0 < cB; c0 += pch) {
    const int nc = cB - c0 < pch ? cB - c0 : pch;
    const int c1 = c0 + nc;
    const int la = c0, lb = c1 < bs ? c1 : bs;
    const int ra = c0 > bs ? c0 : bs, rb = c1 < 2 * bs ? c1 : 2 * bs;
    if (c0 > cA) stage_panel(c0, nc);
    cp_async_wait<0>();
    __syncthreads();
    BCR_CLK(4);
    bcr_cta_forward_inv(Ls, ld, Linv, bs, P, pch, nc);
    for (int r = warp; r < bs; r += nw) {
      for (int c = la + lane; c < lb; c += 32) GLg[(size_t)r * bs + c] = P[r * pch + c - c0];
      for (int c = ra + lane; c < rb; c += 32) GRg[(size_t)r * bs + c - bs] = P[r * pch + c - c0];
    }
    if (c1 > 2 * bs)
      for (int r = tid; r < bs; r += nt) s.g[(size_t)I * bs + r] = P[r * pch + 2 * bs - c0];
    __syncthreads();
    BCR_CLK(8);
  }
#ifdef RSPL_BCR_CLOCKS
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && level == 0) {
    printf("bcr_eliminate phase clocks (CTA 0, accumulated): load %lld | chol6 %lld panel %lld trailing %lld | stage %lld | fwdA %lld fwdA-write %lld fwdB %lld | store %lld\n",
           bcr_clk[0], bcr_clk[1], bcr_clk[2], bcr_clk[3], bcr_clk[4], bcr_clk[5], bcr_clk[6], bcr_clk[7], bcr_clk[8]);
    for (int i = 0; i < 12; ++i) bcr_clk[i] = 0;
  }
#endif
}

// C[r][c] += sum_k A[k][r] B[k][c] over the bs rows of the global blocks A, B (both [bs][bs] row-major), for the
// 3 x 6 register tiles of this thread; slabs of BCR_KC rows are staged in shared memory
constexpr int BCR_TILES_PER_THREAD = ((BCR_BS_MAX / 3) * (BCR_BS_MAX / 6) + BCR_THREADS - 1) / BCR_THREADS; // 3
// (super-blocks of <= 96 unknowns have at most 512 tiles: one per thread, which leaves the registers to unroll the
// k loop and keep several shared-memory loads in flight)
__host__ __device__ constexpr int bcr_tiles_per_thread(int bs) { return ((bs / 3) * (bs / 6) + BCR_THREADS - 1) / BCR_THREADS; }

template <int TPT>
__device__ __forceinline__ void bcr_ata_tiles(const double* A, const double* B, int bs, double* As, double* Bs,
                                              double (&acc)[TPT][18], const double* gvec, double* gs, double& vacc) {
  // gvec != null: thread r < bs also accumulates (A^T gvec)[r] from the staged slabs
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tc = bs / 6; // tiles per row of tiles
  for (int k0 = 0; k0 < bs; k0 += BCR_KC) {
    const int kc = bs - k0 < BCR_KC ? bs - k0 : BCR_KC;
    __syncthreads();
    for (int idx = tid; idx < kc * bs; idx += nt) {
      As[idx] = A[(size_t)k0 * bs + idx];
      Bs[idx] = B[(size_t)k0 * bs + idx];
    }
    if (gvec && tid < kc) gs[tid] = gvec[k0 + tid];
    __syncthreads();
    if (gvec && tid < bs) {
      for (int k = 0; k < kc; ++k) vacc += As[k * bs + tid] * gs[k];
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      const int tix = tid + u * nt;
      if (tix >= (bs / 3) * tc) continue;
      // (the six columns of a tile are cl, cl + tc, ..., so the lanes of a warp read consecutive shared-memory words)
      const int r0 = 3 * (tix / tc), cl = tix % tc;
#pragma unroll(TPT == 1 ? 5 : 1)
      for (int k = 0; k < kc; ++k) {
        const double a0 = As[k * bs + r0], a1 = As[k * bs + r0 + 1], a2 = As[k * bs + r0 + 2];
        double bv[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) bv[q] = Bs[k * bs + cl + q * tc];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          acc[u][q] += a0 * bv[q];
          acc[u][6 + q] += a1 * bv[q];
          acc[u][12 + q] += a2 * bv[q];
        }
      }
    }
  }
}

// ---- level l, even blocks: Schur updates from the two eliminated neighbours and the next level's coupling.
// grid = number of even active blocks; dynamic smem = (2 * BCR_KC * bs + BCR_KC) doubles
template <int TPT>
__global__ void __launch_bounds__(BCR_THREADS) bcr_update(const __grid_constant__ BcrDev s, int level) {
  extern __shared__ __align__(16) unsigned char bcr_smem[];
  const int bs = s.bs, tid = threadIdx.x, nt = blockDim.x;
  double* As = reinterpret_cast<double*>(bcr_smem);
  double* Bs = As + (size_t)BCR_KC * bs;
  double* gs = Bs + (size_t)BCR_KC * bs; // [BCR_KC]
  if (*s.info) return;
  const int h = 1 << level;
  const int p = 2 * blockIdx.x;
  const int J = p * h, Il = J - h, Ir = J + h;
  const bool has_l = p > 0, has_r = Ir < s.M;
  const size_t bb = (size_t)bs * bs;
  const int tc = bs / 6, ntiles = (bs / 3) * tc;
  double acc[TPT][18];
  // D_J -= GR(Il)^T GR(Il) + GL(Ir)^T GL(Ir)
#pragma unroll
  for (int u = 0; u < TPT; ++u)
#pragma unroll
    for (int q = 0; q < 18; ++q) acc[u][q] = 0.0;
  double vl = 0.0, vr = 0.0, vdummy = 0.0; // b_J -= GR(Il)^T g(Il) + GL(Ir)^T g(Ir), thread r < bs
  if (has_l) bcr_ata_tiles(s.GR + (size_t)Il * bb, s.GR + (size_t)Il * bb, bs, As, Bs, acc, s.g + (size_t)Il * bs, gs, vl);
  if (has_r) bcr_ata_tiles(s.GL + (size_t)Ir * bb, s.GL + (size_t)Ir * bb, bs, As, Bs, acc, s.g + (size_t)Ir * bs, gs, vr);
  if (tid < bs) s.x[(size_t)J * bs + tid] -= vl + vr;
  double* Dg = s.D + (size_t)J * bb;
#pragma unroll
  for (int u = 0; u < TPT; ++u) {
    const int tix = tid + u * nt;
    if (tix < ntiles) {
      const int r0 = 3 * (tix / tc), cl = tix % tc;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int q = 0; q < 6; ++q) Dg[(size_t)(r0 + a) * bs + cl + q * tc] -= acc[u][a * 6 + q];
    }
  }
  // E'(J, J + 2h) = -GL(Ir)^T GR(Ir)
  if (has_r && J + 2 * h < s.M) {
#pragma unroll
    for (int u = 0; u < TPT; ++u)
#pragma unroll
      for (int q = 0; q < 18; ++q) acc[u][q] = 0.0;
    bcr_ata_tiles(s.GL + (size_t)Ir * bb, s.GR + (size_t)Ir * bb, bs, As, Bs, acc, nullptr, gs, vdummy);
    double* En = s.E + s.eoff[level + 1] + (size_t)(p / 2) * bb;
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      const int tix = tid + u * nt;
      if (tix < ntiles) {
        const int r0 = 3 * (tix / tc), cl = tix % tc;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int q = 0; q < 6; ++q) En[(size_t)(r0 + a) * bs + cl + q * tc] = -acc[u][a * 6 + q];
      }
    }
  }
}

// ---- the same update with the three neighbour blocks resident in shared memory (bs <= BCR_RESIDENT_BS_MAX):
// GR(Il) | GL(Ir) | GR(Ir) are fetched by cp.async in three groups when the kernel starts, so the CTA waits for one
// global round trip instead of nine slab loads, and the k loops run without barriers. The per-tile sums are formed
// in the same k order as bcr_update, so both give the same bits.
// dynamic smem = (3 * bs * bs + 2 * bs) doubles
constexpr int BCR_RESIDENT_BS_MAX = 96;
__host__ __device__ constexpr size_t bcr_resident_smem(int bs) { return sizeof(double) * (3 * (size_t)bs * bs + 2 * (size_t)bs) + 64; }

__device__ __forceinline__ void bcr_fetch_block(double* dst, const double* src, int n_doubles) {
  for (int idx = 2 * threadIdx.x; idx < n_doubles; idx += 2 * blockDim.x) cp_async16(dst + idx, src + idx);
}

__device__ __forceinline__ void bcr_ata_resident(const double* As, const double* Bs, int bs, int tix, double (&acc)[18]) {
  const int tc = bs / 6;
  if (tix < 0) return;
  const int r0 = 3 * (tix / tc), cl = tix % tc;
#pragma unroll 6
  for (int k = 0; k < bs; ++k) {
    const double a0 = As[k * bs + r0], a1 = As[k * bs + r0 + 1], a2 = As[k * bs + r0 + 2];
    double bv[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) bv[q] = Bs[k * bs + cl + q * tc];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      acc[q] += a0 * bv[q];
      acc[6 + q] += a1 * bv[q];
      acc[12 + q] += a2 * bv[q];
    }
  }
}

// grid = (even active blocks, slices): the 3 x 6 output tiles of a block are split over `slices` CTAs (blockIdx.y) when
// the level has fewer blocks than the GPU has SMs -- the products are bound by the FP64 rate of ONE SM (2.2 M FMAs
// per block), and every CTA fetches the same three input blocks from L2. Tiles are computed exactly as by one CTA.
__global__ void __launch_bounds__(BCR_THREADS) bcr_update_resident(const __grid_constant__ BcrDev s, int level) {
  extern __shared__ __align__(16) unsigned char bcr_smem[];
  const int bs = s.bs, tid = threadIdx.x;
  const size_t bb = (size_t)bs * bs;
  double* S0 = reinterpret_cast<double*>(bcr_smem); // GR(Il)
  double* S1 = S0 + bb;                             // GL(Ir)
  double* S2 = S1 + bb;                             // GR(Ir)
  double* gl = S2 + bb;                             // g(Il)
  double* gr = gl + bs;                             // g(Ir)
  if (*s.info) return;
  const int h = 1 << level;
  const int p = 2 * blockIdx.x;
  const int J = p * h, Il = J - h, Ir = J + h;
  const bool has_l = p > 0, has_r = Ir < s.M, has_n = has_r && J + 2 * h < s.M;
  const int tc = bs / 6, ntiles = (bs / 3) * tc;
  const int per = (ntiles + (int)gridDim.y - 1) / (int)gridDim.y; // tiles of this slice: [slice * per, ...)
  const int tix = tid < per && (int)blockIdx.y * per + tid < ntiles ? (int)blockIdx.y * per + tid : -1;
  const bool vec = blockIdx.y == 0 && tid < bs; // slice 0 also updates the right-hand side
  if (has_l) {
    bcr_fetch_block(S0, s.GR + (size_t)Il * bb, (int)bb);
    bcr_fetch_block(gl, s.g + (size_t)Il * bs, bs);
  }
  cp_async_commit();
  if (has_r) {
    bcr_fetch_block(S1, s.GL + (size_t)Ir * bb, (int)bb);
    bcr_fetch_block(gr, s.g + (size_t)Ir * bs, bs);
  }
  cp_async_commit();
  if (has_n) bcr_fetch_block(S2, s.GR + (size_t)Ir * bb, (int)bb);
  cp_async_commit();
  double acc[18];
#pragma unroll
  for (int q = 0; q < 18; ++q) acc[q] = 0.0;
  double vl = 0.0, vr = 0.0;
  cp_async_wait<2>();
  __syncthreads();
  if (has_l) {
    if (vec)
      for (int k = 0; k < bs; ++k) vl += S0[k * bs + tid] * gl[k];
    bcr_ata_resident(S0, S0, bs, tix, acc);
  }
  cp_async_wait<1>();
  __syncthreads();
  if (has_r) {
    if (vec)
      for (int k = 0; k < bs; ++k) vr += S1[k * bs + tid] * gr[k];
    bcr_ata_resident(S1, S1, bs, tix, acc);
  }
  if (vec) s.x[(size_t)J * bs + tid] -= vl + vr;
  const int r0 = tix >= 0 ? 3 * (tix / tc) : 0, cl = tix >= 0 ? tix % tc : 0;
  if (tix >= 0) {
    double* Dg = s.D + (size_t)J * bb;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q < 6; ++q) Dg[(size_t)(r0 + a) * bs + cl + q * tc] -= acc[a * 6 + q];
  }
  cp_async_wait<0>();
  __syncthreads();
  if (has_n) {
#pragma unroll
    for (int q = 0; q < 18; ++q) acc[q] = 0.0;
    bcr_ata_resident(S1, S2, bs, tix, acc);
    if (tix >= 0) {
      double* En = s.E + s.eoff[level + 1] + (size_t)(p / 2) * bb;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int q = 0; q < 6; ++q) En[(size_t)(r0 + a) * bs + cl + q * tc] = -acc[a * 6 + q];
    }
  }
}

// ---- root: the last active block (block 0). One CTA; dynamic smem = (bs * ld + 2 * bs) doubles + 16 bytes
__global__ void __launch_bounds__(BCR_THREADS) bcr_root(const __grid_constant__ BcrDev s) {
  extern __shared__ __align__(16) unsigned char bcr_smem[];
  const int bs = s.bs, ld = bcr_ld(bs), tid = threadIdx.x, nt = blockDim.x;
  double* Ls = reinterpret_cast<double*>(bcr_smem);
  double* y = Ls + (size_t)bs * ld;
  double* dinv = y + bs;
  int* s_fail = reinterpret_cast<int*>(dinv + bs);
  if (*s.info) return;
  if (tid == 0) *s_fail = 0;
  for (int idx = tid; idx < bs * bs; idx += nt) Ls[(idx / bs) * ld + idx % bs] = s.D[idx];
  for (int r = tid; r < bs; r += nt) y[r] = s.x[r];
  __syncthreads();
  bcr_cta_cholesky(Ls, ld, bs, dinv, s_fail, nullptr, 1); // (y is row bs of the array: forward substitution included)
  if (*s_fail) {
    if (tid == 0) atomicOr(s.info, 1);
    return;
  }
  if (tid < 32) bcr_warp_backward(Ls, ld, dinv, bs, y);
  __syncthreads();
  for (int r = tid; r < bs; r += nt) s.x[r] = y[r];
}

// ---- level l, odd blocks, after the coarser levels: x_I = L^-T (g - GL x_Il - GR x_Ir).
// grid = number of odd active blocks; dynamic smem = (bs * ld + 4 * bs) doubles
__global__ void __launch_bounds__(BCR_THREADS) bcr_backsub(const __grid_constant__ BcrDev s, int level) {
  extern __shared__ __align__(16) unsigned char bcr_smem[];
  const int bs = s.bs, ld = bcr_ld(bs), tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  double* Ls = reinterpret_cast<double*>(bcr_smem);
  double* y = Ls + (size_t)bs * ld;
  double* xl = y + bs;
  double* xr = xl + bs;
  double* dinv = xr + bs;
  if (*s.info) return;
  const int h = 1 << level;
  const int I = (2 * blockIdx.x + 1) * h, Il = I - h, Ir = I + h;
  const bool has_r = Ir < s.M;
  const size_t bb = (size_t)bs * bs;
  const double* Dg = s.F + (size_t)I * bb;
  // the factor lands in shared memory (cp.async) while the warps form g - GL x_Il - GR x_Ir from global memory
  for (int r = warp; r < bs; r += nt / 32)
    for (int c = lane; c <= r; c += 32) cp_async8(&Ls[r * ld + c], &Dg[(size_t)r * bs + c]);
  cp_async_commit();
  for (int r = tid; r < bs; r += nt) {
    dinv[r] = 1.0 / Dg[(size_t)r * bs + r];
    xl[r] = s.x[(size_t)Il * bs + r];
    xr[r] = has_r ? s.x[(size_t)Ir * bs + r] : 0.0;
  }
  __syncthreads();
  const double* GLg = s.GL + (size_t)I * bb;
  const double* GRg = s.GR + (size_t)I * bb;
  // one warp per row (coalesced rows of GL / GR, fixed butterfly), three rows in flight
  const int nw = nt / 32;
  for (int r0 = warp; r0 < bs; r0 += 3 * nw) {
    double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int r = r0 + u * nw;
      if (r < bs)
        for (int c = lane; c < bs; c += 32) v[u] += GLg[(size_t)r * bs + c] * xl[c] + (has_r ? GRg[(size_t)r * bs + c] * xr[c] : 0.0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < 3; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int r = r0 + u * nw;
      if (lane == 0 && r < bs) y[r] = s.g[(size_t)I * bs + r] - v[u];
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  if (tid < 32) bcr_warp_backward(Ls, ld, dinv, bs, y);
  __syncthreads();
  for (int r = tid; r < bs; r += nt) s.x[(size_t)I * bs + r] = y[r];
}

// identity on the padded diagonal of the last super-block (after the memset of D); one CTA
__global__ void bcr_pad(const __grid_constant__ BcrDev s) {
  const int first = s.n - (s.M - 1) * s.bs; // valid unknowns of the last block
  double* Dl = s.D + (size_t)(s.M - 1) * s.bs * s.bs;
  for (int r = first + threadIdx.x; r < s.bs; r += blockDim.x) {
    Dl[(size_t)r * s.bs + r] = 1.0;
    s.x[(size_t)(s.M - 1) * s.bs + r] = 0.0;
  }
}

} // namespace ba
