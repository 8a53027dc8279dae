// local_kernel.cuh — LocalmapOptimization on the device
// (/root/reference/src/g2o_optimization/g2o_optimization.cc:21-252): one CTA per local window runs the
// whole two-pass Levenberg-Marquardt schedule (LM(10) with Huber -> chi2 / depth flagging -> LM(5)
// without kernel -> inlier flags -> write-back) without returning to the host.
//
// Kernels
//   setup_* kernels      build, per window, the landmark-major edge order (mono edges of a landmark
//                        first, then its stereo edges, each in the caller's order = g2o's active-edge
//                        order restricted to the landmark), the CSR offsets, the pose-major edge lists
//                        and the landmark x pose slot table; convert Twc -> Tcw (:42). Grids over
//                        (chunk, window, kind), shared by both solve paths.
//   local_solve_kernel   per LM iteration:
//     K1 linearize_landmarks   residual, Jacobians, Huber, chi2; Hll, bl per landmark (private fp64
//                              accumulation in g2o's edge order); W = Jp^T (rho1 Omega) Jl per edge
//     K1' accumulate_poses     Hpp, bp per free pose over its pose-major list (fixed-order butterfly)
//     K3 schur_prep            per landmark Cholesky of Hll + lambda, Z = W L^-T, y = L^-1 bl
//     K3' schur_reduce         Hs_ij = Hpp_ij + lambda - sum Z_i Z_j^T, bs_i = bp_i - sum Z_i y; output-
//                              stationary (one warp per pose pair, fixed order => bitwise deterministic)
//     K4 cholesky_solve        in-shared-memory LLT of the reduced camera system by warp 0
//     K5/K6 backsub_update_eval xl = L^-T (y - sum Z^T xp), manifold updates, new chi2 per edge
//   then the flagging of :176-206 / :213-231 with g2o's stale-error semantics (§9.12).
// No floating-point atomics anywhere: results do not depend on scheduling.
#pragma once

#include <float.h>
#include <stdint.h>

#include "ba_math.cuh"
#include "frame_kernel.cuh"

namespace ba {

constexpr int LOCAL_THREADS = 256;
constexpr int LOCAL_WARPS = LOCAL_THREADS / 32;
constexpr int SLOT_NONE = 255;

enum { LOCAL_ERR_DUP_EDGE = 1, LOCAL_ERR_DEGREE = 2 };

template <int KIND>
struct KT;
template <>
struct KT<0> { // points
  static constexpr int LD = 3, SD = 3, MD = 3, HD = 6, WD = 18, ROWS = 3;
};
template <>
struct KT<1> { // lines
  static constexpr int LD = 4, SD = 6, MD = 8, HD = 10, WD = 24, ROWS = 4;
};

// one landmark kind (0: points, 1: lines); class 0 = mono, class 1 = stereo constraints
struct KindDev {
  int n_lm, n_edge;
  const int* lm_begin;     // [W+1]
  const int* cls_begin[2]; // [W+1]
  const int* cls_pose[2];
  const int* cls_lm[2];
  const int* cls_cam[2]; // may be null
  const double* cls_meas[2];
  int cls_n[2];
  const double* lm_in; // [SD][n_lm]
  // sorted (landmark-major) edges
  double* meas;  // [MD][n_edge]
  int* info;     // pose | cam << 16 | stereo << 30
  int* lm;       // window-local landmark index
  int* src;      // class << 30 | index in the class array (batch-global)
  double* chi2;  // last evaluated chi2 of the edge (stale-error semantics)
  uint8_t* lvl;  // g2o level: 0 active, 1 excluded
  double* W;     // [n_edge][WD]
  double* Z;     // [n_edge][WD]
  int* ebeg;     // [n_lm + 1] offsets into the sorted edge arrays
  int* cursor;   // [n_lm] scratch
  // Internal landmark numbering: inside a window the landmarks are ordered by DESCENDING degree (stable), so the
  // 32 landmarks a warp of the thread-per-landmark kernels works on have (almost) the same number of edges and
  // its lanes run the same trip count. newidx / orig map the caller's window-local index to the internal one and back.
  int* newidx;   // [n_lm] caller index -> internal index (window-local), stored at lm_begin[w] + caller index
  int* orig;     // [n_lm] internal index -> caller index
  double* x;     // [SD][n_lm] state
  double* xb;    // [SD][n_lm] LM backup
  double* H;     // [HD][n_lm] upper triangle of Hll
  double* b;     // [LD][n_lm]
  double* y;     // [LD][n_lm] L^-1 bl
  uint8_t* act;  // [n_lm] landmark active in the current pass
  uint8_t* slot; // [n_lm][slot_stride] free-pose index -> offset of the edge inside the landmark's segment
  int* plist;    // [n_edge] pose-major edge lists (positions in the sorted arrays)
  int* pbeg;     // [n_poses + 1]
  uint8_t* out_inl[2];
  double* lm_out; // [SD][n_lm]
};

struct LocalDev {
  int n_windows, n_cameras, n_poses;
  const double* cameras;
  const int* pose_begin;
  const double* pose_twc;
  const uint8_t* pose_fixed;
  double* pose_tcw; // [7][n_poses] optimiser poses built by the setup kernel
  double* pose_out; // [7][n_poses]
  int slot_stride;  // >= max free poses per window
  int use_slots;    // the landmark x free-pose slot table is in use (windows of <= 64 free poses): degrees <= 254
  int* setup_free_idx; // [n_poses] window-local free index or -1 (scratch of the setup kernel)
  KindDev k[2];
  void* stats;
  int* err; // device error flag
  int* maxdeg; // [2] largest landmark degree of the batch per kind (setup_scan)
  long long* phase; // [n_windows][8] cycles per phase (diagnostics), may be null
};

struct LocalOpt {
  double thr[4];   // mono pt, stereo pt, mono ln, stereo ln
  double delta[4]; // (float)sqrt(thr)
  int iters[2];
  int bf_float;
  int max_poses, max_free; // smem sizing
};

BA_DEV int edge_base(const KindDev& k, int w) { return k.cls_begin[0][w] + k.cls_begin[1][w]; }

// ------------------------------------------------------------------------------------------------
// setup: landmark-major sorted edge arrays, slot table, initial states. A handful of small kernels
// over (chunk, window, kind) grids so that one huge window (global BA) is as parallel as 1024 small
// ones. Order inside a landmark = (class, caller index) = g2o's edge order for that vertex; the
// scatter uses integer atomics and is followed by a per-landmark sort, so the result is deterministic.
// ------------------------------------------------------------------------------------------------
// poses: Twc -> Tcw (g2o_optimization.cc:42) and the window-local free index; grid = windows
__global__ void __launch_bounds__(LOCAL_THREADS) setup_poses(const __grid_constant__ LocalDev d) {
  const int w = blockIdx.x, tid = threadIdx.x;
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0;
  if (tid == 0) {
    int nf = 0;
    for (int p = 0; p < np; ++p) d.setup_free_idx[p0 + p] = d.pose_fixed[p0 + p] ? -1 : nf++;
  }
  for (int p = tid; p < np; p += LOCAL_THREADS) {
    const double pp[3] = {d.pose_twc[p0 + p], d.pose_twc[d.n_poses + p0 + p], d.pose_twc[2 * d.n_poses + p0 + p]};
    const double qq[4] = {d.pose_twc[3 * d.n_poses + p0 + p], d.pose_twc[4 * d.n_poses + p0 + p],
                          d.pose_twc[5 * d.n_poses + p0 + p], d.pose_twc[6 * d.n_poses + p0 + p]};
    const Pose T = pose_from_twc(pp, qq);
#pragma unroll
    for (int q = 0; q < 4; ++q) d.pose_tcw[(size_t)q * d.n_poses + p0 + p] = T.q[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) d.pose_tcw[(size_t)(4 + q) * d.n_poses + p0 + p] = T.t[q];
  }
}

// MODE 0: degree per landmark (cursor zeroed by the host); MODE 1: scatter the sort keys behind the
// per-landmark cursor (zeroed again by setup_scan). grid (edge chunks, windows, kinds)
template <int MODE>
__global__ void __launch_bounds__(LOCAL_THREADS) setup_edges(const __grid_constant__ LocalDev d) {
  const int w = blockIdx.y;
  const KindDev& k = d.k[blockIdx.z];
  const int l0 = k.lm_begin[w];
  const int idx = blockIdx.x * LOCAL_THREADS + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int i = k.cls_begin[c][w] + idx;
    if (i >= k.cls_begin[c][w + 1]) continue;
    if (MODE == 0) {
      atomicAdd(&k.cursor[l0 + k.cls_lm[c][i]], 1); // by caller index
    } else {
      const int l = l0 + k.newidx[l0 + k.cls_lm[c][i]]; // internal index
      const int pos = k.ebeg[l] + atomicAdd(&k.cursor[l], 1);
      k.src[pos] = (c << 30) | i;
    }
  }
}

// internal landmark order of one (window, kind): stable counting sort by descending degree. Degrees come in
// `cursor` (caller index); they leave in `ebeg` (internal index) for setup_scan. grid (windows, kinds), 256 threads
__global__ void __launch_bounds__(LOCAL_THREADS) setup_order(const __grid_constant__ LocalDev d) {
  __shared__ int s_start[256]; // first internal index of a degree bucket, advanced chunk by chunk
  __shared__ int s_deg[LOCAL_THREADS];
  const int w = blockIdx.x, tid = threadIdx.x;
  const KindDev& k = d.k[blockIdx.y];
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  s_start[tid] = 0;
  __syncthreads();
  for (int i = tid; i < nl; i += LOCAL_THREADS) {
    const int dg = k.cursor[l0 + i];
    atomicAdd(&s_start[dg < 255 ? dg : 255], 1); // histogram (integer: deterministic)
  }
  __syncthreads();
  if (tid == 0) { // bucket starts, largest degree first
    int run = 0;
    for (int b = 255; b >= 0; --b) {
      const int cnt = s_start[b];
      s_start[b] = run;
      run += cnt;
    }
  }
  __syncthreads();
  for (int base = 0; base < nl; base += LOCAL_THREADS) {
    const int i = base + tid;
    const int dg = i < nl ? k.cursor[l0 + i] : -1;
    const int bk = dg < 255 ? dg : 255;
    s_deg[tid] = dg < 0 ? -1 : bk;
    __syncthreads();
    int rank = 0, total = 0;
    if (dg >= 0) {
      for (int t2 = 0; t2 < LOCAL_THREADS; ++t2) {
        const int same = s_deg[t2] == bk ? 1 : 0;
        total += same;
        rank += t2 < tid ? same : 0;
      }
      const int pos = s_start[bk] + rank;
      k.newidx[l0 + i] = pos;
      k.orig[l0 + pos] = i;
      k.ebeg[l0 + pos] = dg;
    }
    __syncthreads();
    if (dg >= 0 && rank == total - 1) s_start[bk] += total; // the last landmark of the bucket in this chunk advances it
    __syncthreads();
  }
}

// exclusive scan of the degrees of one (window, kind) -> ebeg; grid (windows, kinds), 1024 threads
__global__ void __launch_bounds__(1024) setup_scan(const __grid_constant__ LocalDev d) {
  __shared__ int s_warp[32];
  __shared__ int s_run;
  const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const KindDev& k = d.k[blockIdx.y];
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  if (tid == 0) s_run = edge_base(k, w);
  __syncthreads();
  for (int base = 0; base < nl; base += 1024) {
    const int i = base + tid;
    const int v = i < nl ? k.ebeg[l0 + i] : 0; // degree by internal index (setup_order), scanned in place
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      s_warp[lane] = t;
    }
    __syncthreads();
    const int run = s_run;
    if (i < nl) {
      k.ebeg[l0 + i] = run + (warp ? s_warp[warp - 1] : 0) + x - v;
      k.cursor[l0 + i] = 0;
      if (v > 254 && d.use_slots) atomicOr(d.err, LOCAL_ERR_DEGREE);
    }
    {
      const int mv = __reduce_max_sync(0xffffffffu, v);
      if (lane == 0 && mv > 0) atomicMax(&d.maxdeg[blockIdx.y], mv);
    }
    __syncthreads();
    if (tid == 0) s_run = run + s_warp[31];
    __syncthreads();
  }
  if (tid == 0 && w == d.n_windows - 1) k.ebeg[k.n_lm] = k.n_edge;
}

// per landmark: order its keys (class-major, caller order), slot table (cleared by a memset), state copy
template <int KIND>
BA_DEV void setup_landmark(const LocalDev& d, const KindDev& k, int w, int l) {
  using T = KT<KIND>;
  const int* free_idx = d.setup_free_idx + d.pose_begin[w];
  const int a = k.ebeg[l], n = k.cursor[l];
  for (int u = 1; u < n; ++u) {
    const int key = k.src[a + u];
    int v = u - 1;
    while (v >= 0 && k.src[a + v] > key) {
      k.src[a + v + 1] = k.src[a + v];
      --v;
    }
    k.src[a + v + 1] = key;
  }
  if (d.use_slots) {
    uint8_t* sl = k.slot + (size_t)l * d.slot_stride;
    for (int u = 0; u < n && u < 255; ++u) {
      const int key = k.src[a + u];
      const int c = key >> 30, idx = key & 0x3fffffff;
      const int fi = free_idx[k.cls_pose[c][idx]];
      if (fi >= 0) {
        if (sl[fi] != SLOT_NONE) atomicOr(d.err, LOCAL_ERR_DUP_EDGE);
        sl[fi] = (uint8_t)u;
      }
    }
  } else { // large windows build their pair lists from the landmarks (no slot table): look for a repeated pose directly
    for (int u = 1; u < n; ++u) {
      const int key = k.src[a + u];
      const int pu = k.cls_pose[key >> 30][key & 0x3fffffff];
      for (int v = 0; v < u; ++v) {
        const int kv = k.src[a + v];
        if (k.cls_pose[kv >> 30][kv & 0x3fffffff] == pu) atomicOr(d.err, LOCAL_ERR_DUP_EDGE);
      }
    }
  }
  const int lo = k.lm_begin[w] + k.orig[l]; // the caller's landmark
#pragma unroll
  for (int q = 0; q < T::SD; ++q) k.x[(size_t)q * k.n_lm + l] = k.lm_in[(size_t)q * k.n_lm + lo];
}

// grid (landmark chunks, windows, kinds)
__global__ void __launch_bounds__(LOCAL_THREADS) setup_landmarks(const __grid_constant__ LocalDev d) {
  const int w = blockIdx.y;
  const KindDev& k = d.k[blockIdx.z];
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  const int i = blockIdx.x * LOCAL_THREADS + threadIdx.x;
  if (i >= nl) return;
  if (blockIdx.z == 0) setup_landmark<0>(d, k, w, l0 + i);
  else setup_landmark<1>(d, k, w, l0 + i);
}

// gather the edge records into landmark-major planes; grid (edge chunks, windows, kinds)
template <int KIND>
BA_DEV void setup_gather_one(const KindDev& k, int l0, int e) {
  using T = KT<KIND>;
  const int key = k.src[e];
  const int c = key >> 30, idx = key & 0x3fffffff;
  const int cam = k.cls_cam[c] ? k.cls_cam[c][idx] : 0;
  k.info[e] = k.cls_pose[c][idx] | (cam << 16) | (c << 30);
  k.lm[e] = k.newidx[l0 + k.cls_lm[c][idx]];
  const int nm = c ? T::MD : (KIND == 0 ? 2 : 4); // mono: 2 of 3 (points), 4 of 8 (lines)
#pragma unroll
  for (int q = 0; q < T::MD; ++q)
    k.meas[(size_t)q * k.n_edge + e] = q < nm ? k.cls_meas[c][(size_t)q * k.cls_n[c] + idx] : 0.0;
  k.chi2[e] = 0.0;
  k.lvl[e] = 0;
}

__global__ void __launch_bounds__(LOCAL_THREADS) setup_gather(const __grid_constant__ LocalDev d) {
  const int w = blockIdx.y;
  const KindDev& k = d.k[blockIdx.z];
  const int e0 = edge_base(k, w), ne = edge_base(k, w + 1) - e0;
  const int i = blockIdx.x * LOCAL_THREADS + threadIdx.x;
  if (i >= ne) return;
  const int l0 = k.lm_begin[w];
  if (blockIdx.z == 0) setup_gather_one<0>(k, l0, e0 + i);
  else setup_gather_one<1>(k, l0, e0 + i);
}

// 6. pose-major edge lists, one CTA per (pose, window, kind): the CTA scans the window's sorted edge records
// (each warp a contiguous slice, ballot compaction => ascending edge order, deterministic). Its own
// launches because the work is O(poses x edges): a 2000-keyframe window would keep the single setup
// CTA busy for seconds. MODE 0 counts into pbeg, MODE 1 fills plist (pbeg scanned in between).
template <int MODE>
__global__ void __launch_bounds__(LOCAL_THREADS) setup_pose_lists(const __grid_constant__ LocalDev d) {
  __shared__ int s_cnt[LOCAL_WARPS];
  const int p = blockIdx.x, w = blockIdx.y;
  const KindDev& k = d.k[blockIdx.z];
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0;
  if (p >= np) return;
  const int e0 = edge_base(k, w), ne = edge_base(k, w + 1) - e0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (((ne + LOCAL_WARPS - 1) / LOCAL_WARPS) + 31) & ~31; // slice per warp, multiple of 32
  const int a = warp * per, b = a + per < ne ? a + per : ne;
  int cnt = 0;
  for (int base = a; base < b; base += 32) {
    const bool hit = (base + lane < b) && ((k.info[e0 + base + lane] & 0xffff) == p);
    cnt += __popc(__ballot_sync(0xffffffffu, hit));
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (MODE == 0) {
    if (tid == 0) {
      int tot = 0;
#pragma unroll
      for (int q = 0; q < LOCAL_WARPS; ++q) tot += s_cnt[q];
      k.pbeg[p0 + p] = tot;
    }
    return;
  }
  int out = k.pbeg[p0 + p];
  for (int q = 0; q < warp; ++q) out += s_cnt[q];
  for (int base = a; base < b; base += 32) {
    const int e = e0 + base + lane;
    const bool hit = (base + lane < b) && ((k.info[e] & 0xffff) == p);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) k.plist[out + __popc(m & ((1u << lane) - 1))] = e;
    out += __popc(m);
  }
}

// counts -> offsets, one thread per (window, kind)
__global__ void __launch_bounds__(128) setup_pose_scan(const __grid_constant__ LocalDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * d.n_windows) return;
  const int w = i >> 1;
  const KindDev& k = d.k[i & 1];
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0;
  int run = edge_base(k, w);
  for (int p = 0; p < np; ++p) {
    const int cnt = k.pbeg[p0 + p];
    k.pbeg[p0 + p] = run;
    run += cnt;
  }
  if (w == d.n_windows - 1) k.pbeg[d.n_poses] = k.n_edge;
}

// ------------------------------------------------------------------------------------------------
// solve: shared-memory window state
// ------------------------------------------------------------------------------------------------
struct WinScalars {
  long long ph[8];
  long long t_last;
  double lambda, ni, chi_cur, chi_tmp, scale, maxdiag;
  int verdict, n_sys, solve_ok, n_active_edges;
  DevStats st;
};

struct WinSmem {
  double* q;    // [NP][4]
  double* t;    // [NP][3]
  double* R;    // [NP][9]
  double* bq;   // [NP][4] backup
  double* bt;   // [NP][3]
  double* Hpp;  // [NF][21]
  double* bp;   // [NF][6]
  double* Hs;   // [n][n]
  double* bs;   // [n]
  double* xp;   // [n]
  double* red;  // [LOCAL_WARPS][4]
  int* free_idx; // [NP] pose -> free index or -1
  int* sys_idx;  // [NF] free index -> block index in the current system or -1
  int* pose_of;  // [NF] free index -> pose
  int* pact;     // [NP] #active edges of the pose in this pass
  WinScalars* sc;
};

BA_DEV size_t local_smem_bytes_dev(int NP, int NF) {
  const int n = 6 * NF;
  return sizeof(double) * ((size_t)NP * 23 + (size_t)NF * 27 + (size_t)n * n + 2 * n + LOCAL_WARPS * 4) +
         sizeof(int) * ((size_t)2 * NP + 2 * NF) + sizeof(WinScalars) + 64;
}
inline size_t local_smem_bytes(int NP, int NF) {
  const int n = 6 * NF;
  return sizeof(double) * ((size_t)NP * 23 + (size_t)NF * 27 + (size_t)n * n + 2 * n + LOCAL_WARPS * 4) +
         sizeof(int) * ((size_t)2 * NP + 2 * NF) + sizeof(WinScalars) + 64;
}

BA_DEV WinSmem carve(unsigned char* base, int NP, int NF) {
  WinSmem s;
  const int n = 6 * NF;
  double* p = reinterpret_cast<double*>(base);
  s.q = p;
  p += NP * 4;
  s.t = p;
  p += NP * 3;
  s.R = p;
  p += NP * 9;
  s.bq = p;
  p += NP * 4;
  s.bt = p;
  p += NP * 3;
  s.Hpp = p;
  p += NF * 21;
  s.bp = p;
  p += NF * 6;
  s.Hs = p;
  p += (size_t)n * n;
  s.bs = p;
  p += n;
  s.xp = p;
  p += n;
  s.red = p;
  p += LOCAL_WARPS * 4;
  s.sc = reinterpret_cast<WinScalars*>(p);
  int* ip = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(p) + ((sizeof(WinScalars) + 7) & ~size_t(7)));
  s.free_idx = ip;
  ip += NP;
  s.pact = ip;
  ip += NP;
  s.sys_idx = ip;
  ip += NF;
  s.pose_of = ip;
  return s;
}

// phase timer (thread 0 only; diagnostics)
BA_DEV void tick(const WinSmem& s, int phase) {
  if (threadIdx.x == 0) {
    const long long now = clock64();
    s.sc->ph[phase] += now - s.sc->t_last;
    s.sc->t_last = now;
  }
}

// deterministic block sum of one double per thread; result returned to every thread
BA_DEV double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_allreduce(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = red[0];
#pragma unroll
  for (int i = 1; i < LOCAL_WARPS; ++i) s += red[i];
  return s;
}
BA_DEV double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = red[0];
#pragma unroll
  for (int i = 1; i < LOCAL_WARPS; ++i) s = fmax(s, red[i]);
  return s;
}

// Evaluates one edge (mono edges are stereo edges whose extra rows are zero: same H/b/chi2).
template <int KIND, bool WITH_J>
BA_DEV void eval_edge(const Cam& cam, int bf_float, bool stereo, const double* R, const double* t, const double* X,
                      const double* m, double* r, double* Jp, double* Jl) {
  if (KIND == 0) {
    double Xc[3];
    transform_point(R, t, X, Xc);
    const double bf_res = bf_float ? (double)(float)cam.bf : cam.bf;
    if (stereo) {
      point_residual<true>(cam, bf_res, Xc, m, r);
    } else {
      point_residual<false>(cam, cam.bf, Xc, m, r);
      r[2] = 0.0;
    }
    if (WITH_J) {
      point_jac_pose<true>(cam, Xc, Jp);
      point_jac_point<true>(cam, R, Xc, Jl);
      if (!stereo) {
#pragma unroll
        for (int c = 0; c < 6; ++c) Jp[12 + c] = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) Jl[6 + c] = 0.0;
      }
    }
  } else {
    if (WITH_J) {
      line_linearize<true>(cam, R, t, X, m, r, Jp, Jl);
      if (!stereo) {
        r[2] = r[3] = 0.0;
#pragma unroll
        for (int c = 0; c < 12; ++c) Jp[12 + c] = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) Jl[8 + c] = 0.0;
      }
    } else {
      if (stereo) {
        line_residual<true>(cam, R, t, X, m, r);
      } else {
        line_residual<false>(cam, R, t, X, m, r);
        r[2] = r[3] = 0.0;
      }
    }
  }
}

template <int KIND>
BA_DEV double edge_chi2(const double* r) {
  double c = 0;
#pragma unroll
  for (int i = 0; i < KT<KIND>::ROWS; ++i) c += r[i] * r[i];
  return KIND == 0 ? c : 0.1 * c; // information I (points) / 0.1 I (lines), §9.4
}

template <int KIND>
BA_DEV void load_edge(const KindDev& k, int e, double* m) {
#pragma unroll
  for (int q = 0; q < KT<KIND>::MD; ++q) m[q] = k.meas[(size_t)q * k.n_edge + e];
}
template <int KIND>
BA_DEV void load_lm(const KindDev& k, int l, double* X) {
#pragma unroll
  for (int q = 0; q < KT<KIND>::SD; ++q) X[q] = k.x[(size_t)q * k.n_lm + l];
}

// Cholesky factor of a small SPD matrix given as packed upper triangle + lambda on the diagonal.
// Lf packed lower (row-major: L00, L10, L11, L20, ...). Returns false on a non-positive pivot.
template <int N>
BA_DEV bool small_chol(const double* Hup, double lambda, double* Lf, double* inv) {
  double A[N][N];
  int q = 0;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      A[i][j] = Hup[q++] + (i == j ? lambda : 0.0);
      A[j][i] = A[i][j];
    }
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double dsum = A[j][j];
#pragma unroll
    for (int p = 0; p < j; ++p) dsum -= Lf[j * (j + 1) / 2 + p] * Lf[j * (j + 1) / 2 + p];
    if (!(dsum > 0.0)) ok = false;
    const double ljj = sqrt(dsum);
    Lf[j * (j + 1) / 2 + j] = ljj;
    inv[j] = 1.0 / ljj;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s = A[i][j];
#pragma unroll
      for (int p = 0; p < j; ++p) s -= Lf[i * (i + 1) / 2 + p] * Lf[j * (j + 1) / 2 + p];
      Lf[i * (i + 1) / 2 + j] = s * inv[j];
    }
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// K1: linearise, landmark-major (one thread per landmark; edges of the landmark in g2o order)
// ------------------------------------------------------------------------------------------------
template <int KIND>
BA_DEV void linearize_landmarks(const LocalDev& d, const LocalOpt& o, const KindDev& k, int w, const WinSmem& s,
                                bool robust, double& chi_part, double& maxdiag_part, int& nact_part) {
  using T = KT<KIND>;
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  for (int i = threadIdx.x; i < nl; i += LOCAL_THREADS) {
    const int l = l0 + i;
    if (!k.act[l]) continue;
    double X[T::SD];
    load_lm<KIND>(k, l, X);
    double H[T::HD], b[T::LD];
#pragma unroll
    for (int q = 0; q < T::HD; ++q) H[q] = 0;
#pragma unroll
    for (int q = 0; q < T::LD; ++q) b[q] = 0;
    const int ea = k.ebeg[l], eb = k.ebeg[l + 1];
    for (int e = ea; e < eb; ++e) {
      double* We = k.W + (size_t)e * T::WD;
      const int info = k.info[e];
      const int p = info & 0xffff;
      const bool stereo = (info >> 30) & 1;
      const int fi = s.free_idx[p];
      const bool in_sys = fi >= 0 && s.sys_idx[fi] >= 0;
      if (k.lvl[e]) {
        if (in_sys) {
#pragma unroll
          for (int q = 0; q < T::WD; q += 2) *reinterpret_cast<double2*>(We + q) = make_double2(0.0, 0.0);
        }
        continue;
      }
      Cam cam;
      load_cam(d.cameras, (info >> 16) & 0xff, cam);
      double m[T::MD], r[4], Jp[24], Jl[16];
      load_edge<KIND>(k, e, m);
      eval_edge<KIND, true>(cam, o.bf_float, stereo, s.R + 9 * p, s.t + 3 * p, X, m, r, Jp, Jl);
      const double c2 = edge_chi2<KIND>(r);
      k.chi2[e] = c2;
      double wgt = 1.0;
      const double rho0 = robust ? huber(c2, o.delta[2 * KIND + (stereo ? 1 : 0)], wgt) : c2;
      chi_part += rho0;
      nact_part += 1;
      const double wo = (KIND == 0 ? 1.0 : 0.1) * wgt; // rho1 * Omega
      // Hll += Jl^T wo Jl ; bl -= Jl^T wo r
      int q = 0;
#pragma unroll
      for (int a = 0; a < T::LD; ++a) {
        double g = 0;
#pragma unroll
        for (int rr = 0; rr < T::ROWS; ++rr) g += Jl[rr * T::LD + a] * r[rr];
        b[a] -= wo * g;
#pragma unroll
        for (int c = a; c < T::LD; ++c) {
          double h = 0;
#pragma unroll
          for (int rr = 0; rr < T::ROWS; ++rr) h += Jl[rr * T::LD + a] * Jl[rr * T::LD + c];
          H[q++] += wo * h;
        }
      }
      // W = Jp^T wo Jl (6 x LD), only for poses that are in the reduced system
      if (in_sys) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double row[T::LD];
#pragma unroll
          for (int c = 0; c < T::LD; ++c) {
            double h = 0;
#pragma unroll
            for (int rr = 0; rr < T::ROWS; ++rr) h += Jp[rr * 6 + a] * Jl[rr * T::LD + c];
            row[c] = wo * h;
          }
          if (T::LD == 3) {
            We[a * 3 + 0] = row[0];
            We[a * 3 + 1] = row[1];
            We[a * 3 + 2] = row[2];
          } else {
            *reinterpret_cast<double2*>(We + a * 4) = make_double2(row[0], row[1]);
            *reinterpret_cast<double2*>(We + a * 4 + 2) = make_double2(row[2], row[3]);
          }
        }
      }
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < T::LD; ++a)
#pragma unroll
      for (int c = a; c < T::LD; ++c) {
        if (c == a) maxdiag_part = fmax(maxdiag_part, fabs(H[q]));
        k.H[(size_t)q * k.n_lm + l] = H[q];
        ++q;
      }
#pragma unroll
    for (int a = 0; a < T::LD; ++a) k.b[(size_t)a * k.n_lm + l] = b[a];
  }
}

// K1': Hpp, bp of the free poses in the system, pose-major; one warp per pose, lanes stride the list
template <int KIND>
BA_DEV void accumulate_pose_kind(const LocalDev& d, const LocalOpt& o, const KindDev& k, int w, const WinSmem& s, int p,
                                 bool robust, int lane, double* acc) {
  using T = KT<KIND>;
  const int l0 = k.lm_begin[w];
  const int p0 = d.pose_begin[w];
  const int a = k.pbeg[p0 + p], b = k.pbeg[p0 + p + 1];
  for (int it = a + lane; it < b; it += 32) {
    const int e = k.plist[it];
    if (k.lvl[e]) continue;
    const int info = k.info[e];
    const bool stereo = (info >> 30) & 1;
    Cam cam;
    load_cam(d.cameras, (info >> 16) & 0xff, cam);
    double X[T::SD], m[T::MD], r[4], Jp[24], Jl[16];
    load_lm<KIND>(k, l0 + k.lm[e], X);
    load_edge<KIND>(k, e, m);
    eval_edge<KIND, true>(cam, o.bf_float, stereo, s.R + 9 * p, s.t + 3 * p, X, m, r, Jp, Jl);
    double wgt = 1.0;
    if (robust) huber(edge_chi2<KIND>(r), o.delta[2 * KIND + (stereo ? 1 : 0)], wgt);
    const double wo = (KIND == 0 ? 1.0 : 0.1) * wgt;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double g = 0;
#pragma unroll
      for (int rr = 0; rr < T::ROWS; ++rr) g += Jp[rr * 6 + i] * r[rr];
      acc[21 + i] -= wo * g;
#pragma unroll
      for (int j = i; j < 6; ++j) {
        double h = 0;
#pragma unroll
        for (int rr = 0; rr < T::ROWS; ++rr) h += Jp[rr * 6 + i] * Jp[rr * 6 + j];
        acc[up6(i, j)] += wo * h;
      }
    }
  }
}

BA_DEV void accumulate_poses(const LocalDev& d, const LocalOpt& o, int w, const WinSmem& s, int nf, bool robust) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int fi = warp; fi < nf; fi += LOCAL_WARPS) {
    if (s.sys_idx[fi] < 0) continue;
    const int p = s.pose_of[fi];
    double acc[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = 0;
    accumulate_pose_kind<0>(d, o, d.k[0], w, s, p, robust, lane, acc);
    accumulate_pose_kind<1>(d, o, d.k[1], w, s, p, robust, lane, acc);
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = warp_allreduce(acc[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 21; ++q) s.Hpp[fi * 21 + q] = acc[q];
#pragma unroll
      for (int q = 0; q < 6; ++q) s.bp[fi * 6 + q] = acc[21 + q];
    }
  }
}

// K3: per landmark L = chol(Hll + lambda), y = L^-1 bl, Z_e = W_e L^-T
template <int KIND>
BA_DEV void schur_prep(const LocalDev& d, const KindDev& k, int w, const WinSmem& s, double lambda, int& fail) {
  using T = KT<KIND>;
  constexpr int LD = T::LD;
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  for (int i = threadIdx.x; i < nl; i += LOCAL_THREADS) {
    const int l = l0 + i;
    if (!k.act[l]) continue;
    double Hup[T::HD], Lf[LD * (LD + 1) / 2], inv[LD];
#pragma unroll
    for (int q = 0; q < T::HD; ++q) Hup[q] = k.H[(size_t)q * k.n_lm + l];
    if (!small_chol<LD>(Hup, lambda, Lf, inv)) fail = 1;
    double y[LD];
#pragma unroll
    for (int a = 0; a < LD; ++a) {
      double v = k.b[(size_t)a * k.n_lm + l];
#pragma unroll
      for (int p = 0; p < a; ++p) v -= Lf[a * (a + 1) / 2 + p] * y[p];
      y[a] = v * inv[a];
      k.y[(size_t)a * k.n_lm + l] = y[a];
    }
    const int ea = k.ebeg[l], eb = k.ebeg[l + 1];
    for (int e = ea; e < eb; ++e) {
      const int fi = s.free_idx[k.info[e] & 0xffff];
      if (fi < 0 || s.sys_idx[fi] < 0) continue;
      const double* We = k.W + (size_t)e * T::WD;
      double* Ze = k.Z + (size_t)e * T::WD;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double z[LD];
        // z L^T = w  <=>  L z^T = w^T (forward substitution)
#pragma unroll
        for (int c = 0; c < LD; ++c) {
          double v = We[a * LD + c];
#pragma unroll
          for (int p = 0; p < c; ++p) v -= Lf[c * (c + 1) / 2 + p] * z[p];
          z[c] = v * inv[c];
        }
#pragma unroll
        for (int c = 0; c < LD; ++c) Ze[a * LD + c] = z[c];
      }
    }
  }
}

// K3': one warp per pose pair (i <= j) of the current system, output stationary
template <int KIND>
BA_DEV void schur_pair_kind(const LocalDev& d, const KindDev& k, int w, int pi_pose, int fj, bool diag, int lane,
                            double* acc) {
  using T = KT<KIND>;
  constexpr int LD = T::LD;
  const int l0 = k.lm_begin[w];
  const int p0 = d.pose_begin[w];
  const int a = k.pbeg[p0 + pi_pose], b = k.pbeg[p0 + pi_pose + 1];
  for (int it = a + lane; it < b; it += 32) {
    const int e = k.plist[it];
    const int l = l0 + k.lm[e];
    if (!k.act[l]) continue;
    int e2 = e;
    if (!diag) {
      const int sl = k.slot[(size_t)l * d.slot_stride + fj];
      if (sl == SLOT_NONE) continue;
      e2 = k.ebeg[l] + sl;
    }
    double Zi[T::WD], Zj[T::WD];
    const double* zi = k.Z + (size_t)e * T::WD;
    const double* zj = k.Z + (size_t)e2 * T::WD;
#pragma unroll
    for (int q = 0; q < T::WD; q += 2) {
      const double2 v = *reinterpret_cast<const double2*>(zi + q);
      Zi[q] = v.x;
      Zi[q + 1] = v.y;
    }
    if (diag) {
#pragma unroll
      for (int q = 0; q < T::WD; ++q) Zj[q] = Zi[q];
    } else {
#pragma unroll
      for (int q = 0; q < T::WD; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(zj + q);
        Zj[q] = v.x;
        Zj[q + 1] = v.y;
      }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < LD; ++q) v += Zi[r * LD + q] * Zj[c * LD + q];
        acc[r * 6 + c] += v;
      }
    if (diag) {
      double y[LD];
#pragma unroll
      for (int q = 0; q < LD; ++q) y[q] = k.y[(size_t)q * k.n_lm + l];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < LD; ++q) v += Zi[r * LD + q] * y[q];
        acc[36 + r] += v;
      }
    }
  }
}

BA_DEV void schur_reduce(const LocalDev& d, int w, const WinSmem& s, int nf, int n, double lambda) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // enumerate pairs (fi <= fj) of free poses that are in the system, in a fixed order
  int pair = 0;
  for (int fi = 0; fi < nf; ++fi) {
    if (s.sys_idx[fi] < 0) continue;
    for (int fj = fi; fj < nf; ++fj) {
      if (s.sys_idx[fj] < 0) continue;
      if ((pair++ % LOCAL_WARPS) != warp) continue;
      const bool diag = fi == fj;
      double acc[42];
#pragma unroll
      for (int q = 0; q < 42; ++q) acc[q] = 0;
      schur_pair_kind<0>(d, d.k[0], w, s.pose_of[fi], fj, diag, lane, acc);
      schur_pair_kind<1>(d, d.k[1], w, s.pose_of[fi], fj, diag, lane, acc);
#pragma unroll
      for (int q = 0; q < 36; ++q) acc[q] = warp_allreduce(acc[q]);
      if (diag) {
#pragma unroll
        for (int q = 36; q < 42; ++q) acc[q] = warp_allreduce(acc[q]);
      }
      const int bi = 6 * s.sys_idx[fi], bj = 6 * s.sys_idx[fj];
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double v = -acc[r * 6 + c];
            if (diag) {
              const int rr = r < c ? r : c, cc = r < c ? c : r;
              v += s.Hpp[fi * 21 + up6(rr, cc)] + (r == c ? lambda : 0.0);
            }
            s.Hs[(size_t)(bi + r) * n + bj + c] = v;
          }
        if (diag) {
#pragma unroll
          for (int r = 0; r < 6; ++r) s.bs[bi + r] = s.bp[fi * 6 + r] - acc[36 + r];
        }
      }
    }
  }
}

// K4: upper Cholesky (U^T U = Hs, upper triangle) + two triangular sweeps, warp 0, shared memory.
// Fails iff a pivot <= 0 (LinearSolverEigen, §9.11).
BA_DEV void cholesky_solve(const WinSmem& s, int n) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  double* A = s.Hs;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    // column sweep: U[k][j] = (A[k][j] - sum_{p<k} U[p][k] U[p][j]) / U[k][k]
    double dk = 0;
    for (int j = k + lane; j < n; j += 32) {
      double v = A[(size_t)k * n + j];
      {
        // four independent partial sums hide the shared-memory load latency of the dependent chain
        double v1 = 0, v2 = 0, v3 = 0;
        int p = 0;
        for (; p + 3 < k; p += 4) {
          v -= A[(size_t)p * n + k] * A[(size_t)p * n + j];
          v1 -= A[(size_t)(p + 1) * n + k] * A[(size_t)(p + 1) * n + j];
          v2 -= A[(size_t)(p + 2) * n + k] * A[(size_t)(p + 2) * n + j];
          v3 -= A[(size_t)(p + 3) * n + k] * A[(size_t)(p + 3) * n + j];
        }
        for (; p < k; ++p) v -= A[(size_t)p * n + k] * A[(size_t)p * n + j];
        v += (v1 + v2) + v3;
      }
      A[(size_t)k * n + j] = v;
      if (j == k) dk = v;
    }
    dk = __shfl_sync(0xffffffffu, dk, 0);
    if (dk <= 0.0) ok = false;
    const double ukk = sqrt(dk), inv = 1.0 / ukk;
    __syncwarp();
    for (int j = k + lane; j < n; j += 32) A[(size_t)k * n + j] = (j == k) ? ukk : A[(size_t)k * n + j] * inv;
    __syncwarp();
  }
  if (lane == 0) s.sc->solve_ok = ok ? 1 : 0;
  if (!ok) return;
  // U^T y = bs (forward, column sweep), then U x = y (backward)
  for (int i = lane; i < n; i += 32) s.xp[i] = s.bs[i];
  __syncwarp();
  for (int i = 0; i < n; ++i) {
    const double yi = s.xp[i] / A[(size_t)i * n + i];
    __syncwarp();
    if (lane == 0) s.xp[i] = yi;
    for (int j = i + 1 + lane; j < n; j += 32) s.xp[j] -= A[(size_t)i * n + j] * yi;
    __syncwarp();
  }
  for (int i = n - 1; i >= 0; --i) {
    const double xi = s.xp[i] / A[(size_t)i * n + i];
    __syncwarp();
    if (lane == 0) s.xp[i] = xi;
    for (int j = lane; j < i; j += 32) s.xp[j] -= A[(size_t)j * n + i] * xi;
    __syncwarp();
  }
}

// K5/K6: back-substitution, manifold update and re-evaluation, landmark-major
template <int KIND>
BA_DEV void backsub_update_eval(const LocalDev& d, const LocalOpt& o, const KindDev& k, int w, const WinSmem& s,
                                double lambda, bool robust, double& chi_part, double& scale_part) {
  using T = KT<KIND>;
  constexpr int LD = T::LD;
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  for (int i = threadIdx.x; i < nl; i += LOCAL_THREADS) {
    const int l = l0 + i;
    if (!k.act[l]) continue;
    double v[LD];
#pragma unroll
    for (int a = 0; a < LD; ++a) v[a] = k.y[(size_t)a * k.n_lm + l];
    const int ea = k.ebeg[l], eb = k.ebeg[l + 1];
    for (int e = ea; e < eb; ++e) {
      const int fi = s.free_idx[k.info[e] & 0xffff];
      if (fi < 0) continue;
      const int si = s.sys_idx[fi];
      if (si < 0) continue;
      const double* Ze = k.Z + (size_t)e * T::WD;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const double xr = s.xp[6 * si + r];
#pragma unroll
        for (int a = 0; a < LD; ++a) v[a] -= Ze[r * LD + a] * xr;
      }
    }
    // xl = L^-T v
    double Hup[T::HD], Lf[LD * (LD + 1) / 2], inv[LD], xl[LD];
#pragma unroll
    for (int q = 0; q < T::HD; ++q) Hup[q] = k.H[(size_t)q * k.n_lm + l];
    small_chol<LD>(Hup, lambda, Lf, inv);
#pragma unroll
    for (int a = LD - 1; a >= 0; --a) {
      double t2 = v[a];
#pragma unroll
      for (int p = a + 1; p < LD; ++p) t2 -= Lf[p * (p + 1) / 2 + a] * xl[p];
      xl[a] = t2 * inv[a];
    }
#pragma unroll
    for (int a = 0; a < LD; ++a) scale_part += xl[a] * (lambda * xl[a] + k.b[(size_t)a * k.n_lm + l]);
    // backup + oplus
    double X[T::SD], Xn[T::SD];
    load_lm<KIND>(k, l, X);
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
    // computeActiveErrors at the new state
    for (int e = ea; e < eb; ++e) {
      if (k.lvl[e]) continue;
      const int info = k.info[e];
      const int p = info & 0xffff;
      const bool stereo = (info >> 30) & 1;
      Cam cam;
      load_cam(d.cameras, (info >> 16) & 0xff, cam);
      double m[T::MD], r[4];
      load_edge<KIND>(k, e, m);
      eval_edge<KIND, false>(cam, o.bf_float, stereo, s.R + 9 * p, s.t + 3 * p, Xn, m, r, nullptr, nullptr);
      const double c2 = edge_chi2<KIND>(r);
      k.chi2[e] = c2;
      double wgt;
      chi_part += robust ? huber(c2, o.delta[2 * KIND + (stereo ? 1 : 0)], wgt) : c2;
    }
  }
}

template <int KIND>
BA_DEV void restore_landmarks(const KindDev& k, int w) {
  using T = KT<KIND>;
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  for (int i = threadIdx.x; i < nl; i += LOCAL_THREADS) {
    const int l = l0 + i;
    if (!k.act[l]) continue;
#pragma unroll
    for (int q = 0; q < T::SD; ++q) k.x[(size_t)q * k.n_lm + l] = k.xb[(size_t)q * k.n_lm + l];
  }
}

// active sets of a pass (§9.12): landmark active <=> it has a level-0 edge; pose counts in s.pact
template <int KIND>
BA_DEV void mark_active(const LocalDev& d, const KindDev& k, int w, const WinSmem& s) {
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  for (int i = threadIdx.x; i < nl; i += LOCAL_THREADS) {
    const int l = l0 + i;
    int any = 0;
    for (int e = k.ebeg[l]; e < k.ebeg[l + 1]; ++e) {
      if (k.lvl[e]) continue;
      any = 1;
      atomicAdd(&s.pact[k.info[e] & 0xffff], 1);
    }
    k.act[l] = (uint8_t)any;
  }
}

// flagging after pass 1 (:176-206): level 1 iff chi2 > thr or (points) depth <= 0
template <int KIND>
BA_DEV void flag_pass1(const LocalDev& d, const LocalOpt& o, const KindDev& k, int w, const WinSmem& s) {
  using T = KT<KIND>;
  const int l0 = k.lm_begin[w];
  const int e0 = edge_base(k, w), e1 = edge_base(k, w + 1);
  for (int e = e0 + threadIdx.x; e < e1; e += LOCAL_THREADS) {
    const int info = k.info[e];
    const bool stereo = (info >> 30) & 1;
    bool out = k.chi2[e] > o.thr[2 * KIND + (stereo ? 1 : 0)];
    if (KIND == 0) {
      const int p = info & 0xffff;
      double X[3], Xc[3];
      load_lm<0>(k, l0 + k.lm[e], X);
      transform_point(s.R + 9 * p, s.t + 3 * p, X, Xc);
      if (!(Xc[2] > 0.0)) out = true;
    }
    k.lvl[e] = out ? 1 : 0;
  }
  (void)sizeof(T);
}

// final flags (:213-231) scattered back to the caller's constraint order
template <int KIND>
BA_DEV void final_flags(const LocalDev& d, const LocalOpt& o, const KindDev& k, int w, const WinSmem& s) {
  const int l0 = k.lm_begin[w];
  const int e0 = edge_base(k, w), e1 = edge_base(k, w + 1);
  for (int e = e0 + threadIdx.x; e < e1; e += LOCAL_THREADS) {
    const int info = k.info[e];
    const bool stereo = (info >> 30) & 1;
    bool inl = k.chi2[e] <= o.thr[2 * KIND + (stereo ? 1 : 0)];
    if (KIND == 0) {
      const int p = info & 0xffff;
      double X[3], Xc[3];
      load_lm<0>(k, l0 + k.lm[e], X);
      transform_point(s.R + 9 * p, s.t + 3 * p, X, Xc);
      inl = inl && (Xc[2] > 0.0);
    }
    const int key = k.src[e];
    k.out_inl[key >> 30][key & 0x3fffffff] = inl ? 1 : 0;
  }
}

template <int KIND>
BA_DEV void write_landmarks(const KindDev& k, int w) {
  using T = KT<KIND>;
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  for (int i = threadIdx.x; i < nl; i += LOCAL_THREADS)
#pragma unroll
    for (int q = 0; q < T::SD; ++q) k.lm_out[(size_t)q * k.n_lm + l0 + k.orig[l0 + i]] = k.x[(size_t)q * k.n_lm + l0 + i];
}

// One LM pass = SparseOptimizer::initializeOptimization(0) + optimize(iters) (§9.9, §9.12)
BA_DEV void lm_pass(const LocalDev& d, const LocalOpt& o, int w, const WinSmem& s, int np, int nf, int pass,
                    bool robust) {
  const int tid = threadIdx.x;
  WinScalars& sc = *s.sc;
  // ---- active sets
  for (int p = tid; p < np; p += LOCAL_THREADS) s.pact[p] = 0;
  __syncthreads();
  mark_active<0>(d, d.k[0], w, s);
  mark_active<1>(d, d.k[1], w, s);
  __syncthreads();
  if (tid == 0) {
    int nsys = 0;
    for (int fi = 0; fi < nf; ++fi) s.sys_idx[fi] = s.pact[s.pose_of[fi]] > 0 ? nsys++ : -1;
    sc.n_sys = nsys;
    sc.lambda = 0;
    sc.ni = 2;
  }
  __syncthreads();
  const int n = 6 * sc.n_sys;
  const int iters = o.iters[pass];
  for (int it = 0; it < iters; ++it) {
    // ---- computeActiveErrors + activeRobustChi2 + buildSystem
    double chi_part = 0, maxd = 0;
    int nact = 0;
    tick(s, 7);
    linearize_landmarks<0>(d, o, d.k[0], w, s, robust, chi_part, maxd, nact);
    linearize_landmarks<1>(d, o, d.k[1], w, s, robust, chi_part, maxd, nact);
    __syncthreads();
    tick(s, 0);
    accumulate_poses(d, o, w, s, nf, robust);
    __syncthreads();
    tick(s, 1);
    const double chi0 = block_sum(chi_part, s.red);
    const double nact_all = block_sum((double)nact, s.red);
    if (nact_all == 0.0) break; // no active edge: g2o's optimize() returns without iterating
    if (it == 0) {
      double m2 = block_max(maxd, s.red);
      __syncthreads();
      if (tid == 0) {
        for (int fi = 0; fi < nf; ++fi) {
          if (s.sys_idx[fi] < 0) continue;
          for (int i = 0; i < 6; ++i) m2 = fmax(m2, fabs(s.Hpp[fi * 21 + up6(i, i)]));
        }
        sc.lambda = 1e-5 * m2; // computeLambdaInit: tau * max |H_jj| over all active free vertices
        sc.ni = 2;
      }
    }
    if (tid == 0) {
      sc.chi_cur = chi0;
      sc.st.edges_linearized += (long long)nact_all;
      sc.st.edges_evaluated += (long long)nact_all;
    }
    __syncthreads();
    int qmax = 0;
    int verdict;
    do {
      const double lambda = sc.lambda;
      int fail = 0;
      schur_prep<0>(d, d.k[0], w, s, lambda, fail);
      schur_prep<1>(d, d.k[1], w, s, lambda, fail);
      const int any_fail = __syncthreads_or(fail);
      tick(s, 2);
      if (n > 0) {
        schur_reduce(d, w, s, nf, n, lambda);
        __syncthreads();
        tick(s, 3);
        cholesky_solve(s, n);
        __syncthreads();
        tick(s, 4);
      } else if (tid == 0) {
        sc.solve_ok = 1;
      }
      __syncthreads();
      const bool ok = sc.solve_ok && !any_fail;
      double chi_part2 = 0, scale_part = 0;
      if (ok) {
        // poses: backup, oplus (VertexSE3Expmap: exp(x) * T), refresh R
        for (int fi = tid; fi < nf; fi += LOCAL_THREADS) {
          const int si = s.sys_idx[fi];
          if (si < 0) continue;
          const int p = s.pose_of[fi];
          Pose T;
#pragma unroll
          for (int q = 0; q < 4; ++q) T.q[q] = s.bq[4 * p + q] = s.q[4 * p + q];
#pragma unroll
          for (int q = 0; q < 3; ++q) T.t[q] = s.bt[3 * p + q] = s.t[3 * p + q];
          const Pose Tn = pose_oplus(T, s.xp + 6 * si);
          double Rn[9];
          quat_to_R(Tn.q, Rn);
#pragma unroll
          for (int q = 0; q < 4; ++q) s.q[4 * p + q] = Tn.q[q];
#pragma unroll
          for (int q = 0; q < 3; ++q) s.t[3 * p + q] = Tn.t[q];
#pragma unroll
          for (int q = 0; q < 9; ++q) s.R[9 * p + q] = Rn[q];
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const double xv = s.xp[6 * si + q];
            scale_part += xv * (lambda * xv + s.bp[fi * 6 + q]);
          }
        }
        __syncthreads();
        backsub_update_eval<0>(d, o, d.k[0], w, s, lambda, robust, chi_part2, scale_part);
        backsub_update_eval<1>(d, o, d.k[1], w, s, lambda, robust, chi_part2, scale_part);
      }
      const double chi1 = block_sum(chi_part2, s.red);
      const double scale = block_sum(scale_part, s.red);
      tick(s, 5);
      if (tid == 0) {
        // a failed factorisation is a rejected step (tempChi = DBL_MAX, §9.9); the states were not touched
        const double tempChi = ok ? chi1 : DBL_MAX;
        double rho = sc.chi_cur - tempChi;
        rho /= (ok ? scale : 0.0) + 1e-3;
        bool stop_lambda = false;
        int accepted = 0;
        if (rho > 0 && isfinite(tempChi)) {
          const double c = 2 * rho - 1;
          double alpha = 1. - c * c * c;
          alpha = fmin(alpha, 2. / 3.);
          sc.lambda *= fmax(1. / 3., alpha);
          sc.ni = 2;
          sc.chi_cur = tempChi;
          accepted = 1;
        } else {
          sc.lambda *= sc.ni;
          sc.ni *= 2;
          if (!isfinite(sc.lambda)) stop_lambda = true;
        }
        const int q1 = stop_lambda ? qmax : qmax + 1;
        int v;
        if (!stop_lambda && rho < 0 && q1 < 10) v = 1;
        else if (q1 == 10 || rho == 0 || !isfinite(sc.lambda)) v = 2;
        else v = 0;
        sc.verdict = v | (accepted ? 0 : 4) | (ok ? 8 : 0);
        sc.st.trials[pass]++;
        if (ok) sc.st.edges_evaluated += (long long)nact_all;
      }
      __syncthreads();
      const int vv = sc.verdict;
      verdict = vv & 3;
      if ((vv & 4) && (vv & 8)) { // rejected after an applied update: pop()
        for (int fi = tid; fi < nf; fi += LOCAL_THREADS) {
          if (s.sys_idx[fi] < 0) continue;
          const int p = s.pose_of[fi];
#pragma unroll
          for (int q = 0; q < 4; ++q) s.q[4 * p + q] = s.bq[4 * p + q];
#pragma unroll
          for (int q = 0; q < 3; ++q) s.t[3 * p + q] = s.bt[3 * p + q];
          double Rn[9];
          quat_to_R(s.q + 4 * p, Rn);
#pragma unroll
          for (int q = 0; q < 9; ++q) s.R[9 * p + q] = Rn[q];
        }
        restore_landmarks<0>(d.k[0], w);
        restore_landmarks<1>(d.k[1], w);
      }
      __syncthreads();
      tick(s, 6);
      qmax++;
    } while (verdict == 1);
    if (tid == 0) sc.st.iters[pass]++;
    if (verdict == 2) break;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LOCAL_THREADS) local_solve_kernel(const __grid_constant__ LocalDev d,
                                                                     const __grid_constant__ LocalOpt o) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = blockIdx.x;
  const int tid = threadIdx.x;
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0;
  WinSmem s = carve(smem_raw, o.max_poses, o.max_free);
  __shared__ int s_nf;
  if (tid == 0) {
    int nf = 0;
    for (int p = 0; p < np; ++p) {
      if (d.pose_fixed[p0 + p]) {
        s.free_idx[p] = -1;
      } else {
        s.free_idx[p] = nf;
        s.pose_of[nf] = p;
        ++nf;
      }
    }
    s_nf = nf;
    WinScalars& sc = *s.sc;
#pragma unroll
    for (int i = 0; i < 4; ++i) sc.st.iters[i] = sc.st.trials[i] = 0;
    sc.st.edges_linearized = sc.st.edges_evaluated = 0;
    sc.st.final_chi2 = 0;
    sc.st.final_lambda = 0;
    sc.chi_cur = 0;
    sc.lambda = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sc.ph[i] = 0;
    sc.t_last = clock64();
  }
  for (int p = tid; p < np; p += LOCAL_THREADS) {
    double q[4], R[9];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = s.q[4 * p + i] = d.pose_tcw[(size_t)i * d.n_poses + p0 + p];
#pragma unroll
    for (int i = 0; i < 3; ++i) s.t[3 * p + i] = d.pose_tcw[(size_t)(4 + i) * d.n_poses + p0 + p];
    quat_to_R(q, R);
#pragma unroll
    for (int i = 0; i < 9; ++i) s.R[9 * p + i] = R[i];
  }
  __syncthreads();
  const int nf = s_nf;

  // pass 1: optimizer.initializeOptimization(); optimize(10) with Huber (:172-173)
  lm_pass(d, o, w, s, np, nf, 0, true);
  // check inlier observations, strip kernels (:176-206)
  flag_pass1<0>(d, o, d.k[0], w, s);
  flag_pass1<1>(d, o, d.k[1], w, s);
  __syncthreads();
  // pass 2: initializeOptimization(0); optimize(5) (:209-210)
  lm_pass(d, o, w, s, np, nf, 1, false);
  // final flags + write-back (:213-251)
  final_flags<0>(d, o, d.k[0], w, s);
  final_flags<1>(d, o, d.k[1], w, s);
  write_landmarks<0>(d.k[0], w);
  write_landmarks<1>(d.k[1], w);
  for (int p = tid; p < np; p += LOCAL_THREADS) {
    Pose T;
#pragma unroll
    for (int i = 0; i < 4; ++i) T.q[i] = s.q[4 * p + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) T.t[i] = s.t[3 * p + i];
    const Pose Twc = pose_inverse(T); // :237-239
#pragma unroll
    for (int i = 0; i < 3; ++i) d.pose_out[(size_t)i * d.n_poses + p0 + p] = Twc.t[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) d.pose_out[(size_t)(3 + i) * d.n_poses + p0 + p] = Twc.q[i];
  }
  __syncthreads();
  tick(s, 7);
  if (tid == 0 && d.phase) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d.phase[(size_t)w * 8 + i] = s.sc->ph[i];
  }
  if (tid == 0 && d.stats) {
    s.sc->st.final_chi2 = s.sc->chi_cur;
    s.sc->st.final_lambda = s.sc->lambda;
    reinterpret_cast<DevStats*>(d.stats)[w] = s.sc->st;
  }
}

} // namespace ba
