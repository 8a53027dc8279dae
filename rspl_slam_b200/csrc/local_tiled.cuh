// local_tiled.cuh — Schur elimination of the landmarks WITHOUT a round trip of Z = W L^-T through HBM.
//
// The batched path of local_batched.cuh wrote one 144/192-byte Z block per edge (kb_schur_prep), gathered every
// block ~5 times from L2 (kb_schur_reduce: one warp per pose pair) and once more in kb_backsub: L2-request bound
// at 0.14 of the HBM contract roofline. Here the landmarks of a window are cut into TILES sized by shared memory,
// and one CTA per tile does both halves of g2o's Schur step (BlockSolver::solve):
//   phase 1a  thread per landmark : L = chol(Hll + lambda), y = L^-1 bl                  -> shared memory
//   phase 1b  thread per EDGE     : W = Jp^T (rho1 Omega) Jl from the 40-byte edge record, Z = W L^-T -> shared memory
//   phase 2   8 lanes per pose pair: sum over the tile's entries of that pair's list of Z_i Z_j^T (and Z_i y on
//             the diagonal); a lane owns one landmark-coordinate column of one entry per iteration (a rank-1
//             update of 36 + 6 private accumulators), a fixed 3-step transpose reduction sums the 8 lanes
// and writes the tile's partial reduced system hs_tile[window][tile][pair][42]. kb_solve adds the tiles in tile
// order. Back-substitution (kb_backsub_rc) no longer reads Z either: it recomputes W^T xp from the edge record at
// the pre-update state, xl = (Hll + lambda)^-1 (bl - sum W^T xp). Every sum has a fixed owner and order: results
// stay bitwise reproducible and independent of the batch composition.
//
// Tiles are defined by quantiles of the cumulative shared-memory cost A * (landmarks before l) + B * (edges
// before l) along the window's internal landmark order, so a tile never exceeds Q + A + B * max degree bytes.
#pragma once

#include "local_batched.cuh"

namespace ba {

template <int KIND>
struct TileCost {
  // per landmark: L (packed lower) + 1/diag + y + the landmark state; per edge: the Z block + a 16-bit tile-local landmark index (+ at
  // run time 4 bytes per staged pair entry: a landmark of degree k has at most k(k+1)/2 entries, (k+1)/2 per edge)
  static constexpr int LN = KT<KIND>::LD * (KT<KIND>::LD + 1) / 2 + 2 * KT<KIND>::LD + KT<KIND>::SD; // doubles: 15 (points), 24 (lines)
  static constexpr int A = LN * 8;
  static constexpr int B = ZBlk<KIND>::N * 8 + 2;
};
BA_DEV int tile_cost_a(int kind) { return kind ? TileCost<1>::A : TileCost<0>::A; }

constexpr int TILE_THREADS = 256;
constexpr int TILE_GROUPS = TILE_THREADS / 8;


// ---- setup -------------------------------------------------------------------------------------
// tile boundaries by binary search on the (strictly increasing) cumulative cost; grid (ceil((Tcap+1)/128), W, 2)
__global__ void __launch_bounds__(128) kt_tiles_lm(const __grid_constant__ LocalDev d, const __grid_constant__ TileDev td) {
  const int w = blockIdx.y, kind = blockIdx.z;
  const int t = blockIdx.x * 128 + threadIdx.x;
  const KindDev& k = d.k[kind];
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  const int e0 = edge_base(k, w), ne = edge_base(k, w + 1) - e0;
  const long long A = tile_cost_a(kind), B = td.cost_b[kind];
  const long long total = A * nl + B * ne;
  const int nt = (int)((total + td.Q - 1) / td.Q);
  if (t == 0) td.ntile[w * 2 + kind] = nt;
  if (t > nt || t > td.Tcap) return;
  // first landmark i in [0, nl] with cost(i) >= t * Q  (cost(nl) = total)
  const long long target = (long long)t * td.Q;
  int lo = 0, hi = nl;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const long long c = A * mid + B * (k.ebeg[l0 + mid] - e0);
    if (c >= target) hi = mid;
    else lo = mid + 1;
  }
  td.tile_lm[(size_t)(w * 2 + kind) * (td.Tcap + 1) + t] = l0 + (t == nt ? nl : lo);
}

// per (window, kind, tile, compact pair): first entry of the pair's list whose first edge lies in the tile or later.
// grid (ceil(Pmax * (Tcap+1) / 256), W, 2)
__global__ void __launch_bounds__(256) kt_tiles_pairs(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                      const __grid_constant__ TileDev td) {
  const int w = blockIdx.y, kind = blockIdx.z;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int li = idx % b.Pmax, t = idx / b.Pmax;
  const int nt = td.ntile[w * 2 + kind];
  if (t > nt || t > td.Tcap || li >= b.n_ne[w]) return;
  const KindDev& k = d.k[kind];
  const int p = b.ne_list[(size_t)w * b.Pmax + li];
  const int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1) + 2 * p + kind;
  int lo = pb[0], hi = pb[1];
  if (t == nt) {
    lo = hi;
  } else {
    const int l = td.tile_lm[(size_t)(w * 2 + kind) * (td.Tcap + 1) + t];
    const int efirst = k.ebeg[l];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (b.pairs[mid].x >= efirst) hi = mid;
      else lo = mid + 1;
    }
  }
  td.tpb[((size_t)(w * 2 + kind) * (td.Tcap + 1) + t) * b.Pmax + li] = lo;
}

// offsets of the pairs inside a tile's entry block (exclusive scan over the compact pairs); one warp per
// (tile, window, kind); the total is parked in tent_base for kt_tiles_base. grid (Tcap, W, 2), 32 threads
__global__ void __launch_bounds__(32) kt_tiles_scan(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                    const __grid_constant__ TileDev td) {
  const int t = blockIdx.x, w = blockIdx.y, kind = blockIdx.z, lane = threadIdx.x;
  const size_t wk = (size_t)(w * 2 + kind);
  if (t >= td.ntile[wk]) {
    if (lane == 0) td.tent_base[wk * td.Tcap + t] = 0;
    return;
  }
  const int n = b.n_ne[w];
  const int* tp0 = td.tpb + (wk * (td.Tcap + 1) + t) * b.Pmax;
  const int* tp1 = tp0 + b.Pmax;
  int* so = td.tso + (wk * td.Tcap + t) * (b.Pmax + 1);
  int run = 0;
  for (int base = 0; base < n; base += 32) {
    const int li = base + lane;
    const int v = li < n ? tp1[li] - tp0[li] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (li < n) so[li] = run + x - v;
    run += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) {
    so[n] = run;
    td.tent_base[wk * td.Tcap + t] = run;
  }
}

// totals -> positions of the tiles' entry blocks inside the window's region of tent; one thread per window
__global__ void __launch_bounds__(128) kt_tiles_base(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                     const __grid_constant__ TileDev td) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= d.n_windows) return;
  // (every block starts on a multiple of 4 entries = 16 bytes: the region of window w is shifted by the slack of the
  // windows before it)
  int run = (int)b.pair_base[w] + 8 * td.Tcap * w;
  for (int kind = 0; kind < 2; ++kind)
    for (int t = 0; t < td.Tcap; ++t) {
      int* p = td.tent_base + (size_t)(w * 2 + kind) * td.Tcap + t;
      const int c = *p;
      run = (run + 3) & ~3;
      *p = run;
      run += c;
    }
}

// tile-major entry blocks: thread per (pair, tile, window, kind) copies its slice with tile-local edge indices;
// grid (ceil(Pmax * Tcap / 256), W, 2)
__global__ void __launch_bounds__(256) kt_tiles_fill(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                     const __grid_constant__ TileDev td) {
  const int w = blockIdx.y, kind = blockIdx.z;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int li = idx % b.Pmax, t = idx / b.Pmax;
  const size_t wk = (size_t)(w * 2 + kind);
  if (t >= td.ntile[wk] || li >= b.n_ne[w]) return;
  const KindDev& k = d.k[kind];
  const int ea = k.ebeg[td.tile_lm[wk * (td.Tcap + 1) + t]];
  const int* tp0 = td.tpb + (wk * (td.Tcap + 1) + t) * b.Pmax;
  const int beg = tp0[li], end = tp0[b.Pmax + li];
  ushort2* dst = td.tent + td.tent_base[wk * td.Tcap + t] + td.tso[(wk * td.Tcap + t) * (b.Pmax + 1) + li];
  for (int i = beg; i < end; ++i) {
    const int2 e = b.pairs[i];
    BA_CHECK(e.x >= ea && e.y >= ea && e.x - ea < 65536 && e.y - ea < 65536);
    dst[i - beg] = make_ushort2((unsigned short)(e.x - ea), (unsigned short)(e.y - ea));
  }
}

// one descriptor per (tile, window, kind); grid (ceil(Tcap / 128), W, 2)
__global__ void __launch_bounds__(128) kt_tiles_desc(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                     const __grid_constant__ TileDev td) {
  const int w = blockIdx.y, kind = blockIdx.z, t = blockIdx.x * 128 + threadIdx.x;
  if (t >= td.Tcap) return;
  const size_t wk = (size_t)(w * 2 + kind);
  const size_t tile_id = wk * td.Tcap + t;
  int4 a = make_int4(0, 0, 0, 0), c = make_int4(0, 0, 0, 0);
  if (t < td.ntile[wk]) {
    const KindDev& k = d.k[kind];
    const int* tl = td.tile_lm + wk * (td.Tcap + 1) + t;
    const int n_ne = b.n_ne[w];
    a = make_int4(tl[0], tl[1], k.ebeg[tl[0]], k.ebeg[tl[1]]);
    c = make_int4(td.tso[tile_id * (b.Pmax + 1) + n_ne], td.tent_base[tile_id], n_ne, 1);
  }
  td.desc[2 * tile_id] = a;
  td.desc[2 * tile_id + 1] = c;
}

// processing order of the compact pairs of a window: longest list first (stable rank sort), so that the four
// 8-lane groups of a warp work on lists of similar length; one warp per window, grid ceil(W / 4), 128 threads
__global__ void __launch_bounds__(128) kt_order(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                const __grid_constant__ TileDev td) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= d.n_windows) return;
  const int n = b.n_ne[w];
  const int* list = b.ne_list + (size_t)w * b.Pmax;
  const int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1);
  for (int i = lane; i < n; i += 32) {
    const int len = pb[2 * list[i] + 2] - pb[2 * list[i]];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const int lj = pb[2 * list[j] + 2] - pb[2 * list[j]];
      rank += (lj > len || (lj == len && j < i)) ? 1 : 0;
    }
    int fi, fj;
    pair_decode(list[i], b.ws[w].nf, fi, fj);
    td.order[(size_t)w * b.Pmax + rank] = i | (fi == fj ? 1 << 30 : 0);
  }
}

// ---- the fused Schur tile kernel ---------------------------------------------------------------
// sums 48 per-lane values over the 8 lanes of a group (xor 4, 2, 1): lane8 ends with elements 6*lane8 .. 6*lane8+5
BA_DEV void group8_transpose_reduce48(double* v, int lane8) {
#pragma unroll
  for (int off = 4, n = 48; off >= 1; off >>= 1, n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane8 & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = upper ? v[i] : v[i + half];
      const double keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

#ifndef TILE_MIN_CTAS
#define TILE_MIN_CTAS 2
#endif

// shared-memory bytes of a tile that do not depend on its landmarks: per-pair offsets and processing order, the
// window's pose table (R, t, in-system flag) and the cameras
__host__ __device__ inline size_t tile_fixed_bytes(int Pmax, int max_poses, int n_cameras) {
  return (size_t)(2 * Pmax + 2) * 4 + (size_t)max_poses * 13 * 8 + (size_t)n_cameras * 5 * 8 + 64;
}

// The kernel is written so that every global load depends on nothing but the block and thread indices: the
// window's small tables, the tile's pair entries, the landmark blocks and the first batch of edge records are all
// requested up front (one exposed memory latency), the next batch of edge records is requested before the
// current one is processed, and the three compute phases read shared memory only.
template <int KIND>
__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_CTAS)
    kt_schur_tile(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b, const __grid_constant__ LocalOpt o,
                  const __grid_constant__ TileDev td) {
  using T = KT<KIND>;
  constexpr int LD = T::LD, SD = T::SD, MD = T::MD, ZN = ZBlk<KIND>::N, NTRI = LD * (LD + 1) / 2;
  constexpr int LN = TileCost<KIND>::LN; // L packed, 1/diag, y, X
  constexpr int OFF_INV = NTRI, OFF_Y = NTRI + LD, OFF_X = NTRI + 2 * LD;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const int w = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
  WinState& s = b.ws[w];
  const size_t tile_id = (size_t)(w * 2 + KIND) * td.Tcap + t;
  // (the window state and the tile descriptor are independent reads: one memory latency, not a chain of five)
  const int4 d0 = td.desc[2 * tile_id], d1 = td.desc[2 * tile_id + 1];
  if (s.stage != STAGE_NEED_TRIAL) return;
  if (!d1.w) return;
  const KindDev& k = d.k[KIND];
  const int la = d0.x, lb = d0.y, ea = d0.z, eb = d0.w;
  const int n_ent = d1.x, n_ne = d1.z;
  const int* tso_g = td.tso + tile_id * (b.Pmax + 1);
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0, f0 = b.nf_begin[w], l0 = k.lm_begin[w];
  const int nl = lb - la, ne = eb - ea;
  const int n_ent4 = (n_ent + 3) >> 2;
  // ---- shared-memory layout
  double* Zs = reinterpret_cast<double*>(tile_smem);                       // [ne][ZN], column-major blocks (ZCOL)
  double* Ls = Zs + (size_t)ne * ZN;                                        // [nl][LN]
  double* Ps = Ls + (size_t)nl * LN;                                        // [np][13]: R, t, in-system flag
  double* Cs = Ps + (size_t)np * 13;                                        // [n_cameras][5]
  uint4* ent4 = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(Cs + (size_t)d.n_cameras * 5) + 15) & ~(uintptr_t)15);
  const ushort2* ent = reinterpret_cast<const ushort2*>(ent4);             // [n_ent] staged pair entries
  int* tso = reinterpret_cast<int*>(ent4 + n_ent4);                         // [n_ne + 1]
  int* ords = tso + n_ne + 1;                                               // [n_ne] processing order | diag << 30
  unsigned short* elm = reinterpret_cast<unsigned short*>(ords + n_ne);     // [ne] tile-local landmark of an edge
#ifdef RSPL_BA_CHECKED
  {
    unsigned dyn_smem;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
    BA_CHECK(reinterpret_cast<unsigned char*>(elm + ne) <= tile_smem + dyn_smem);
    BA_CHECK(nl > 0 && ne >= 0 && la >= l0 && lb <= k.lm_begin[w + 1] && eb <= k.n_edge);
  }
#endif
  const double lambda = s.lambda;
  const bool robust = s.robust;

  // ---- first batch of edge records (registers), requested before anything else is waited for
  struct Rec {
    int info, lm;
    unsigned char lvl;
    double m[MD];
  };
  auto load_rec = [&](int j, Rec& r) {
    const int e = ea + j;
    r.info = k.info[e];
    r.lm = k.lm[e];
    r.lvl = k.lvl[e];
#pragma unroll
    for (int q = 0; q < MD; ++q) r.m[q] = k.meas[(size_t)q * k.n_edge + e];
  };
  Rec cur;
  if (tid < ne) load_rec(tid, cur);

  // ---- the first landmark of this thread (phase 1a), requested now as well
  double Hup0[T::HD], bl0[LD], X0[SD];
  bool act0 = false;
  if (tid < nl) {
    const int l = la + tid;
    act0 = k.act[l];
#pragma unroll
    for (int q = 0; q < T::HD; ++q) Hup0[q] = k.H[(size_t)q * k.n_lm + l];
#pragma unroll
    for (int q = 0; q < LD; ++q) bl0[q] = k.b[(size_t)q * k.n_lm + l];
#pragma unroll
    for (int q = 0; q < SD; ++q) X0[q] = k.x[(size_t)q * k.n_lm + l];
  }

  // ---- stage the window tables and the tile's pair entries: cp.async, so that every copy is in flight at once and
  // none holds registers (they are waited for at the barrier that ends phase 1a, which does not read them)
  {
    const uint4* src = reinterpret_cast<const uint4*>(td.tent + d1.y);
    for (int i = tid; i < n_ent4; i += TILE_THREADS) cp_async16(ent4 + i, src + i);
    for (int i = tid; i <= n_ne; i += TILE_THREADS) cp_async4(tso + i, tso_g + i);
    for (int i = tid; i < n_ne; i += TILE_THREADS) cp_async4(ords + i, td.order + (size_t)w * b.Pmax + i);
    for (int i = tid; i < np * 12; i += TILE_THREADS) {
      const int p = i / 12, q = i - p * 12;
      cp_async8(Ps + (size_t)p * 13 + q, q < 9 ? b.P_R + 9 * (size_t)(p0 + p) + q : b.P_t + 3 * (size_t)(p0 + p) + q - 9);
    }
    for (int i = tid; i < d.n_cameras * 5; i += TILE_THREADS) cp_async8(Cs + i, d.cameras + i);
    cp_async_commit();
    for (int p = tid; p < np; p += TILE_THREADS) {
      const int fi = b.free_idx[p0 + p];
      Ps[(size_t)p * 13 + 12] = (fi >= 0 && b.sys_idx[f0 + fi] >= 0) ? 1.0 : 0.0;
    }
  }

  // ---- phase 1a: landmark factors (and the landmark state, for phase 1b)
  int fail = 0;
  for (int i = tid; i < nl; i += TILE_THREADS) {
    const int l = la + i;
    double* Lm = Ls + (size_t)i * LN;
    double Hup[T::HD], bl[LD], X[SD];
    bool act;
    if (i == tid) { // requested in the prologue
      act = act0;
#pragma unroll
      for (int q = 0; q < T::HD; ++q) Hup[q] = Hup0[q];
#pragma unroll
      for (int q = 0; q < LD; ++q) bl[q] = bl0[q];
#pragma unroll
      for (int q = 0; q < SD; ++q) X[q] = X0[q];
    } else {
      act = k.act[l];
#pragma unroll
      for (int q = 0; q < T::HD; ++q) Hup[q] = k.H[(size_t)q * k.n_lm + l];
#pragma unroll
      for (int q = 0; q < LD; ++q) bl[q] = k.b[(size_t)q * k.n_lm + l];
#pragma unroll
      for (int q = 0; q < SD; ++q) X[q] = k.x[(size_t)q * k.n_lm + l];
    }
#pragma unroll
    for (int q = 0; q < SD; ++q) Lm[OFF_X + q] = X[q];
    if (!act) { // every edge of an inactive landmark is excluded (level 1): its Z blocks are zero
#pragma unroll
      for (int q = 0; q < OFF_X; ++q) Lm[q] = 0.0;
      continue;
    }
    double Lf[NTRI], inv[LD];
    if (!small_chol<LD>(Hup, lambda, Lf, inv)) fail = 1;
    double y[LD];
#pragma unroll
    for (int a = 0; a < LD; ++a) {
      double v = bl[a];
#pragma unroll
      for (int p = 0; p < a; ++p) v -= Lf[a * (a + 1) / 2 + p] * y[p];
      y[a] = v * inv[a];
    }
#pragma unroll
    for (int q = 0; q < NTRI; ++q) Lm[q] = Lf[q];
#pragma unroll
    for (int q = 0; q < LD; ++q) {
      Lm[OFF_INV + q] = inv[q];
      Lm[OFF_Y + q] = y[q];
    }
  }
  if (fail) atomicOr(&s.prep_fail, 1);
  cp_async_wait<0>(); // the staged tables
  __syncthreads();

  // ---- phase 1b: Z blocks, one thread per edge; the next record is in flight while this one is processed
  for (int j = tid; j < ne; j += TILE_THREADS) {
    Rec nxt;
    if (j + TILE_THREADS < ne) load_rec(j + TILE_THREADS, nxt);
    const int p = cur.info & 0xffff;
    BA_CHECK(p < np && cur.lm >= la - l0 && cur.lm < lb - l0);
    const int li = l0 + cur.lm - la; // tile-local landmark
    elm[j] = (unsigned short)li;
    double2* Zj = reinterpret_cast<double2*>(Zs + (size_t)j * ZN);
    const double* Pp = Ps + (size_t)p * 13;
    if (Pp[12] == 0.0 || cur.lvl) { // fixed pose / pose outside the system / excluded edge (level 1): no contribution
#pragma unroll
      for (int q = 0; q < ZN / 2; ++q) Zj[q] = make_double2(0.0, 0.0);
    } else {
      const bool stereo = (cur.info >> 30) & 1;
      Cam cam;
      load_cam(Cs, (cur.info >> 16) & 0xff, cam);
      const double* Lm = Ls + (size_t)li * LN;
      double X[SD], R[9], tt[3], r[4], Jp[24], Jl[16];
#pragma unroll
      for (int q = 0; q < SD; ++q) X[q] = Lm[OFF_X + q];
#pragma unroll
      for (int q = 0; q < 9; ++q) R[q] = Pp[q];
#pragma unroll
      for (int q = 0; q < 3; ++q) tt[q] = Pp[9 + q];
      eval_edge<KIND, true>(cam, o.bf_float, stereo, R, tt, X, cur.m, r, Jp, Jl);
      double wgt = 1.0;
      if (robust) huber(edge_chi2<KIND>(r), o.delta[2 * KIND + (stereo ? 1 : 0)], wgt);
      const double wo = (KIND == 0 ? 1.0 : 0.1) * wgt;
      double Lf[NTRI], inv[LD];
#pragma unroll
      for (int q = 0; q < NTRI; ++q) Lf[q] = Lm[q];
#pragma unroll
      for (int q = 0; q < LD; ++q) inv[q] = Lm[OFF_INV + q];
      double z[6][LD];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int c = 0; c < LD; ++c) {
          double h = 0;
#pragma unroll
          for (int rr = 0; rr < T::ROWS; ++rr) h += Jp[rr * 6 + a] * Jl[rr * LD + c];
          double v = wo * h;
#pragma unroll
          for (int pp = 0; pp < c; ++pp) v -= Lf[c * (c + 1) / 2 + pp] * z[a][pp];
          z[a][c] = v * inv[c];
        }
      }
#pragma unroll
      for (int c = 0; c < LD; ++c) {
        Zj[c * 3 + 0] = make_double2(z[0][c], z[1][c]);
        Zj[c * 3 + 1] = make_double2(z[2][c], z[3][c]);
        Zj[c * 3 + 2] = make_double2(z[4][c], z[5][c]);
      }
    }
    cur = nxt;
  }
  __shared__ int s_next_quad;
  if (tid == 0) s_next_quad = TILE_THREADS / 32;
  __syncthreads();

  // ---- phase 2: pair products over the tile's entries, 8 lanes per pose pair, four pairs ("quad") per warp at a
  // time. The pairs are ordered by list length (kt_order), so the four lists of a quad have similar lengths; warps
  // take the next quad from a shared counter, which evens out the long (diagonal) and short lists. Which warp
  // computes a pair does not change its value.
  const int lane = tid & 31, lane8 = tid & 7, g4 = (tid >> 3) & 3;
  const int nquads = (n_ne + 3) >> 2;
  const int tt2 = KIND ? td.Tp + t : t;
  double* out_base = td.hs_tile + ((size_t)w * (td.Tp + td.Tl) + tt2) * b.Pmax * 42;
  for (int quad = tid >> 5; quad < nquads;) {
    const int oi = quad * 4 + g4;
    const bool have = oi < n_ne;
    int li = 0, beg = 0, ncol = 0;
    bool diag = false;
    if (have) {
      const int ov = ords[oi];
      li = ov & 0x3fffffff;
      diag = (ov >> 30) & 1;
      beg = tso[li];
      ncol = (tso[li + 1] - beg) * LD;
    }
    double acc[48];
#pragma unroll
    for (int q = 0; q < 48; ++q) acc[q] = 0.0;
    if (diag) {
      // Z_i Z_i^T is symmetric: only the upper triangle is accumulated (the reduced-system Cholesky reads nothing else)
      for (int c = lane8; c < ncol; c += 8) {
        const ushort2 cur2 = ent[beg + c / LD];
        const int q = c % LD;
        const int ji = cur2.x;
        const double2* a2 = reinterpret_cast<const double2*>(Zs + (size_t)ji * ZN + q * ZCOL);
        const double yq = Ls[(size_t)elm[ji] * LN + OFF_Y + q];
        double za[6];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const double2 v = a2[u];
          za[2 * u] = v.x;
          za[2 * u + 1] = v.y;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
          for (int cc = r; cc < 6; ++cc) acc[r * 6 + cc] += za[r] * za[cc];
          acc[36 + r] += za[r] * yq;
        }
      }
    } else {
      for (int c = lane8; c < ncol; c += 8) {
        const ushort2 cur2 = ent[beg + c / LD];
        const int q = c % LD;
        const double2* a2 = reinterpret_cast<const double2*>(Zs + (size_t)cur2.x * ZN + q * ZCOL);
        const double2* b2 = reinterpret_cast<const double2*>(Zs + (size_t)cur2.y * ZN + q * ZCOL);
        double za[6], zb[6];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const double2 v = a2[u];
          za[2 * u] = v.x;
          za[2 * u + 1] = v.y;
          const double2 v2 = b2[u];
          zb[2 * u] = v2.x;
          zb[2 * u + 1] = v2.y;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) acc[r * 6 + cc] += za[r] * zb[cc];
      }
    }
    __syncwarp();
    group8_transpose_reduce48(acc, lane8);
    if (have && lane8 < 7) {
      double2* out = reinterpret_cast<double2*>(out_base + (size_t)li * 42 + 6 * lane8);
      out[0] = make_double2(acc[0], acc[1]);
      out[1] = make_double2(acc[2], acc[3]);
      out[2] = make_double2(acc[4], acc[5]);
    }
    int nq = 0;
    if (lane == 0) nq = atomicAdd(&s_next_quad, 1);
    quad = __shfl_sync(0xffffffffu, nq, 0);
  }
}

// Sum of the per-tile partial reduced systems of every window in tile order (points, then lines) -> hs_part, the
// array kb_solve assembles from; one thread per (pair, element), eight independent loads in flight, the additions
// stay in tile order; grid (ceil(Pmax * 42 / 256), W)
constexpr int TILE_SUM_LANES = 1;
__global__ void __launch_bounds__(256) kt_tile_sum(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                   const __grid_constant__ TileDev td) {
  const int w = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (b.ws[w].stage != STAGE_NEED_TRIAL || idx >= b.n_ne[w] * 42) return;
  const int ntp = td.ntile[w * 2], ntl = td.ntile[w * 2 + 1];
  const size_t stride = (size_t)b.Pmax * 42;
  const double* base = td.hs_tile + (size_t)w * (td.Tp + td.Tl) * stride + idx;
  double v = 0.0;
  int t = 0;
  for (; t + 8 <= ntp; t += 8) {
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = base[(size_t)(t + u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) v += a[u];
  }
  for (; t < ntp; ++t) v += base[(size_t)t * stride];
  double vl = 0.0;
  for (t = 0; t < ntl; ++t) vl += base[(size_t)(td.Tp + t) * stride];
  b.hs_part[(size_t)w * stride + idx] = v + vl;
}

// Loop condition of the whole-schedule CUDA graph (conditional WHILE node): non-zero while any window is still
// iterating. `count` = 1 inside the loop body: one more super-step done (read back at download for the statistics).
__global__ void __launch_bounds__(256) kt_cond(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                               cudaGraphConditionalHandle h, int count) {
  __shared__ int s_any;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  int any = 0;
  for (int w = threadIdx.x; w < d.n_windows; w += 256) any |= b.ws[w].stage != STAGE_DONE ? 1 : 0;
  if (any) s_any = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    cudaGraphSetConditional(h, s_any ? 1u : 0u);
    if (count) ++*b.n_active;
  }
}

// ---- back-substitution without Z: xl = (Hll + lambda)^-1 (bl - sum_e W_e^T xp), W_e^T xp = wo Jl^T (Jp xp) recomputed
// from the edge record at the pre-update state (pose backups P_bR / P_bt; the landmark itself is updated here).
template <int KIND>
BA_DEV void backsub_rc_one(const LocalDev& d, const BatchDev& b, const TileDev& td, const LocalOpt& o, const KindDev& k, int w,
                           int l, double lambda, bool robust, double& chi_part, double& scale_part) {
  using T = KT<KIND>;
  constexpr int LD = T::LD;
  const int p0 = d.pose_begin[w], f0 = b.nf_begin[w];
  double X[T::SD];
  load_lm<KIND>(k, l, X);
  double u[LD];
#pragma unroll
  for (int a = 0; a < LD; ++a) u[a] = 0.0;
  const int ea = k.ebeg[l], eb = k.ebeg[l + 1];
  // (the header of the next edge and the slot of its pose are requested before this edge is evaluated)
  int info_n = 0, fi_n = -1;
  unsigned char lvl_n = 1;
  if (ea < eb) {
    lvl_n = k.lvl[ea];
    info_n = k.info[ea];
    fi_n = b.free_idx[p0 + (info_n & 0xffff)];
  }
  for (int e = ea; e < eb; ++e) {
    const int info = info_n, fi = fi_n;
    const bool skip = lvl_n != 0;
    if (e + 1 < eb) {
      lvl_n = k.lvl[e + 1];
      info_n = k.info[e + 1];
      fi_n = b.free_idx[p0 + (info_n & 0xffff)];
    }
    if (skip) continue;
    const int p = info & 0xffff;
    BA_CHECK(p < d.pose_begin[w + 1] - d.pose_begin[w]);
    if (fi < 0 || b.sys_idx[f0 + fi] < 0) continue;
    const bool stereo = (info >> 30) & 1;
    Cam cam;
    load_cam(d.cameras, (info >> 16) & 0xff, cam);
    double m[T::MD], r[4], Jp[24], Jl[16];
    load_edge<KIND>(k, e, m);
    eval_edge<KIND, true>(cam, o.bf_float, stereo, td.P_bR + 9 * (size_t)(p0 + p), b.P_bt + 3 * (size_t)(p0 + p), X, m, r, Jp, Jl);
    double wgt = 1.0;
    if (robust) huber(edge_chi2<KIND>(r), o.delta[2 * KIND + (stereo ? 1 : 0)], wgt);
    const double wo = (KIND == 0 ? 1.0 : 0.1) * wgt;
    const double* xv = b.xp + (size_t)(f0 + fi) * 6;
    double g[T::ROWS];
#pragma unroll
    for (int rr = 0; rr < T::ROWS; ++rr) {
      double acc = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) acc += Jp[rr * 6 + a] * xv[a];
      g[rr] = wo * acc;
    }
#pragma unroll
    for (int a = 0; a < LD; ++a) {
      double acc = 0;
#pragma unroll
      for (int rr = 0; rr < T::ROWS; ++rr) acc += Jl[rr * LD + a] * g[rr];
      u[a] += acc;
    }
  }
  double Hup[T::HD], Lf[LD * (LD + 1) / 2], inv[LD], v[LD], xl[LD];
#pragma unroll
  for (int q = 0; q < T::HD; ++q) Hup[q] = k.H[(size_t)q * k.n_lm + l];
  small_chol<LD>(Hup, lambda, Lf, inv);
  double bl[LD];
#pragma unroll
  for (int a = 0; a < LD; ++a) { // L v = bl - u
    bl[a] = k.b[(size_t)a * k.n_lm + l];
    double t2 = bl[a] - u[a];
#pragma unroll
    for (int p = 0; p < a; ++p) t2 -= Lf[a * (a + 1) / 2 + p] * v[p];
    v[a] = t2 * inv[a];
  }
#pragma unroll
  for (int a = LD - 1; a >= 0; --a) { // L^T xl = v
    double t2 = v[a];
#pragma unroll
    for (int p = a + 1; p < LD; ++p) t2 -= Lf[p * (p + 1) / 2 + a] * xl[p];
    xl[a] = t2 * inv[a];
  }
#pragma unroll
  for (int a = 0; a < LD; ++a) scale_part += xl[a] * (lambda * xl[a] + bl[a]);
  double Xn[T::SD];
#pragma unroll
  for (int q = 0; q < T::SD; ++q) k.xb[(size_t)q * k.n_lm + l] = X[q];
  if (KIND == 0) {
#pragma unroll
    for (int q = 0; q < 3; ++q) Xn[q] = X[q] + xl[q];
  } else {
    line_oplus(X, xl, Xn);
  }
#pragma unroll
  for (int q = 0; q < T::SD; ++q) k.x[(size_t)q * k.n_lm + l] = Xn[q];
  for (int e = ea; e < eb; ++e) {
    if (k.lvl[e]) continue;
    const int info = k.info[e];
    const int p = info & 0xffff;
    BA_CHECK(p < d.pose_begin[w + 1] - d.pose_begin[w]);
    const bool stereo = (info >> 30) & 1;
    Cam cam;
    load_cam(d.cameras, (info >> 16) & 0xff, cam);
    double m[T::MD], r[4];
    load_edge<KIND>(k, e, m);
    eval_edge<KIND, false>(cam, o.bf_float, stereo, b.P_R + 9 * (size_t)(p0 + p), b.P_t + 3 * (size_t)(p0 + p), Xn, m, r,
                           nullptr, nullptr);
    const double c2 = edge_chi2<KIND>(r);
    k.chi2[e] = c2;
    double wgt;
    chi_part += robust ? huber(c2, o.delta[2 * KIND + (stereo ? 1 : 0)], wgt) : c2;
  }
}

template <int KIND>
__global__ void __launch_bounds__(BT) kt_backsub_rc(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                    const __grid_constant__ LocalOpt o, const __grid_constant__ TileDev td) {
  __shared__ double red[BW * 2];
  const int w = blockIdx.y;
  const int c = blockIdx.x + (KIND ? b.Cp : 0);
  const WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_TRIAL || !s.solve_ok) return;
  double v[2] = {0, 0}; // chi1, scale
  {
    const KindDev& k = d.k[KIND];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i < nl && k.act[l0 + i]) backsub_rc_one<KIND>(d, b, td, o, k, w, l0 + i, s.lambda, s.robust, v[0], v[1]);
  }
  cta_reduce<2>(v, red);
  if (threadIdx.x == 0) {
    double* pp = b.part + ((size_t)w * b.C + c) * 4;
    pp[0] = cta_reduce_get<2>(red, 0);
    pp[1] = cta_reduce_get<2>(red, 1);
  }
}

} // namespace ba
