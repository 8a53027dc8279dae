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
constexpr int BCR_KC = 30;         // rows per staged slab of bcr_update

struct BcrDev {
  int M, bs, n;       // super-blocks, unknowns per super-block, system size
  int levels;
  double* D;          // [M][bs*bs] row-major, full storage; overwritten by the lower Cholesky factor when eliminated
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

// In-place lower Cholesky of the n x n matrix in shared memory (n a multiple of 6), whole CTA, three barriers per
// 6-column panel. The strictly upper part is not touched; dinv [n] receives the reciprocal diagonal of the factor.
// *s_fail is set on a pivot <= 0.
__device__ void bcr_cta_cholesky(double* Ls, int ld, int n, double* dinv, int* s_fail) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int k0 = 0; k0 < n; k0 += 6) {
    double Lk[6][6], inv[6];
    if (!bcr_chol6(Ls, ld, k0, Lk, inv)) *s_fail = 1; // (benign race: every thread writes the same value)
    __syncthreads();                                  // everyone has read the diagonal block
    // (static indices, one writer: a thread-dependent index into Lk moves the whole block to local memory -- 336 bytes
    // of stack traffic per thread and panel; it made kb_solve, bcr_eliminate and bcr_root 1.3 - 1.6 x slower)
    {
      // thread r * 6 + c publishes L[r][c]: the value is picked with a select chain over static indices, the store
      // itself has a computed address
      double v = 0.0, iv = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
#pragma unroll
        for (int c = 0; c <= r; ++c) v = tid == r * 6 + c ? Lk[r][c] : v;
        iv = tid == r * 7 ? inv[r] : iv;
      }
      const int r = tid / 6, c = tid - 6 * r;
      if (tid < 36 && c <= r) {
        Ls[(k0 + r) * ld + k0 + c] = v;
        if (c == r) dinv[k0 + r] = iv;
      }
    }
    for (int i = k0 + 6 + tid; i < n; i += nt) { // panel: X Lkk^T = A, one thread per row
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
    __syncthreads();
    // trailing update of the lower triangle: a warp per row i, lanes over the columns j <= i
    for (int i = k0 + 6 + warp; i < n; i += nw) {
      double pi[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) pi[q] = Ls[i * ld + k0 + q];
      for (int j = k0 + 6 + lane; j <= i; j += 32) {
        const double* pj = Ls + j * ld + k0;
        double v = Ls[i * ld + j];
#pragma unroll
        for (int q = 0; q < 6; ++q) v -= pi[q] * pj[q];
        Ls[i * ld + j] = v;
      }
    }
    __syncthreads();
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

// x = L^-T y in place (one warp, column-oriented); y in shared memory
__device__ void bcr_warp_backward(const double* Ls, int ld, const double* dinv, int n, double* y) {
  const int lane = threadIdx.x & 31;
  for (int i = n - 1; i >= 0; --i) {
    const double xi = y[i] * dinv[i];
    __syncwarp();
    if (lane == 0) y[i] = xi;
    for (int j = lane; j < i; j += 32) y[j] -= Ls[i * ld + j] * xi;
    __syncwarp();
  }
}

// ---- level l, odd blocks: factorise and form GL, GR, g. grid = number of odd active blocks; dynamic smem =
// (bs * ld + bs + bs * pch) doubles + 16 bytes, pch = columns of one panel chunk (>= 1, odd)
__global__ void __launch_bounds__(BCR_THREADS) bcr_eliminate(const __grid_constant__ BcrDev s, int level, int pch) {
  extern __shared__ __align__(16) unsigned char bcr_smem[];
  const int bs = s.bs, ld = bcr_ld(bs), tid = threadIdx.x, nt = blockDim.x;
  double* Ls = reinterpret_cast<double*>(bcr_smem);
  double* dinv = Ls + (size_t)bs * ld;
  double* P = dinv + bs;
  int* s_fail = reinterpret_cast<int*>(P + (size_t)bs * pch);
  const int h = 1 << level;
  const int p = 2 * blockIdx.x + 1;
  const int I = p * h, Il = I - h, Ir = I + h;
  const bool has_r = Ir < s.M;
  const size_t bb = (size_t)bs * bs;
  double* Dg = s.D + (size_t)I * bb;
  if (tid == 0) *s_fail = 0;
  for (int idx = tid; idx < bs * bs; idx += nt) Ls[(idx / bs) * ld + idx % bs] = Dg[idx];
  __syncthreads();
  bcr_cta_cholesky(Ls, ld, bs, dinv, s_fail);
  if (*s_fail) {
    if (tid == 0) atomicOr(s.info, 1);
    return; // (uniform) the solve is rejected as a whole
  }
  for (int idx = tid; idx < bs * bs; idx += nt) { // keep the factor for the back substitution
    const int r = idx / bs, c = idx % bs;
    Dg[idx] = c <= r ? Ls[r * ld + c] : 0.0;
  }
  const double* El = s.E + s.eoff[level] + (size_t)(p - 1) * bb; // rows Il, columns I
  const double* Er = s.E + s.eoff[level] + (size_t)p * bb;       // rows I, columns Ir
  double* GLg = s.GL + (size_t)I * bb;
  double* GRg = s.GR + (size_t)I * bb;
  const int ncol = 2 * bs + 1; // [ E(Il,I)^T | E(I,Ir) | b_I ]
  for (int c0 = 0; c0 < ncol; c0 += pch) {
    const int nc = ncol - c0 < pch ? ncol - c0 : pch;
    // (columns of E(Il,I)^T are rows of E(Il,I): r fastest keeps those global reads coalesced; pch is odd, so the
    // transposing shared-memory writes are conflict-free)
    for (int idx = tid; idx < bs * nc; idx += nt) {
      const int cl = idx / bs, r = idx - cl * bs, c = c0 + cl;
      if (c < bs) P[r * pch + cl] = El[(size_t)c * bs + r];
    }
    for (int idx = tid; idx < bs * nc; idx += nt) {
      const int r = idx / nc, cl = idx - r * nc, c = c0 + cl;
      if (c >= 2 * bs) P[r * pch + cl] = s.x[(size_t)I * bs + r];
      else if (c >= bs) P[r * pch + cl] = has_r ? Er[(size_t)r * bs + c - bs] : 0.0;
    }
    __syncthreads();
    bcr_cta_forward(Ls, ld, dinv, bs, P, pch, nc);
    for (int idx = tid; idx < bs * nc; idx += nt) {
      const int r = idx / nc, c = c0 + idx % nc;
      const double v = P[r * pch + idx % nc];
      if (c < bs) GLg[(size_t)r * bs + c] = v;
      else if (c < 2 * bs) GRg[(size_t)r * bs + c - bs] = v;
      else s.g[(size_t)I * bs + r] = v;
    }
    __syncthreads();
  }
  (void)Il;
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

__device__ __forceinline__ void bcr_ata_resident(const double* As, const double* Bs, int bs, double (&acc)[18]) {
  const int tc = bs / 6, tix = threadIdx.x;
  if (tix >= (bs / 3) * tc) return;
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
    if (tid < bs)
      for (int k = 0; k < bs; ++k) vl += S0[k * bs + tid] * gl[k];
    bcr_ata_resident(S0, S0, bs, acc);
  }
  cp_async_wait<1>();
  __syncthreads();
  if (has_r) {
    if (tid < bs)
      for (int k = 0; k < bs; ++k) vr += S1[k * bs + tid] * gr[k];
    bcr_ata_resident(S1, S1, bs, acc);
  }
  if (tid < bs) s.x[(size_t)J * bs + tid] -= vl + vr;
  const int r0 = 3 * (tid / tc), cl = tid % tc;
  if (tid < ntiles) {
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
    bcr_ata_resident(S1, S2, bs, acc);
    if (tid < ntiles) {
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
  bcr_cta_cholesky(Ls, ld, bs, dinv, s_fail);
  if (*s_fail) {
    if (tid == 0) atomicOr(s.info, 1);
    return;
  }
  bcr_cta_forward(Ls, ld, dinv, bs, y, 1, 1);
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
  const double* Dg = s.D + (size_t)I * bb;
  for (int idx = tid; idx < bs * bs; idx += nt) Ls[(idx / bs) * ld + idx % bs] = Dg[idx];
  for (int r = tid; r < bs; r += nt) {
    dinv[r] = 1.0 / Dg[(size_t)r * bs + r];
    xl[r] = s.x[(size_t)Il * bs + r];
    xr[r] = has_r ? s.x[(size_t)Ir * bs + r] : 0.0;
  }
  __syncthreads();
  const double* GLg = s.GL + (size_t)I * bb;
  const double* GRg = s.GR + (size_t)I * bb;
  for (int r = warp; r < bs; r += nt / 32) { // one warp per row: coalesced rows of GL / GR, fixed butterfly
    double v = 0.0;
    for (int c = lane; c < bs; c += 32) v += GLg[(size_t)r * bs + c] * xl[c] + (has_r ? GRg[(size_t)r * bs + c] * xr[c] : 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) y[r] = s.g[(size_t)I * bs + r] - v;
  }
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
