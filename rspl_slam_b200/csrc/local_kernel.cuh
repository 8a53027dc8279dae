// local_kernel.cuh — LocalmapOptimization on the device, part 1: data layout and setup
// (/root/reference/src/g2o_optimization/g2o_optimization.cc:21-252).
//
//   setup_* kernels      build, per window, the landmark-major edge order (mono edges of a landmark
//                        first, then its stereo edges, each in the caller's order = g2o's active-edge
//                        order restricted to the landmark), the CSR offsets, the pose-major edge lists
//                        and the landmark x pose slot table; convert Twc -> Tcw (:42). Grids over
//                        (chunk, window, kind).
//   device helpers       per-edge evaluation (residual, Jacobians), small Cholesky, write-back.
// The Levenberg-Marquardt schedule itself lives in local_batched.cuh (one kernel per LM phase over all windows)
// and local_tiled.cuh (Schur elimination per shared-memory tile).
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
  // scratch of the large-window setup variants (null when the batch has no large window)
  int* order_hist;     // [(w*2+kind)][order_chunks][256] degree histograms per 1024-landmark chunk -> bucket offsets
  int order_chunks;
  int* pose_cursor[2]; // [n_poses] per kind: fill cursors of the pose-major lists
  int* plist_tmp[2];   // [n_edge] per kind: scratch of the list sort
  KindDev k[2];
  void* stats;
  int* err; // device error flag
  int* maxdeg; // [2] largest landmark degree of the batch per kind (setup_scan)
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

// ---- the same stable counting sort for LARGE windows (global BA: 10^6 landmarks in one window), spread over
// 1024-landmark chunks: histogram per chunk -> per-bucket offsets of every chunk -> placement per chunk. The result is
// identical to setup_order's.
constexpr int ORDER_CHUNK = 1024;
// grid (chunks, windows, kinds), 256 threads
__global__ void __launch_bounds__(LOCAL_THREADS) setup_order_hist(const __grid_constant__ LocalDev d) {
  __shared__ int s_h[256];
  const int c = blockIdx.x, w = blockIdx.y, kind = blockIdx.z, tid = threadIdx.x;
  const KindDev& k = d.k[kind];
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  s_h[tid] = 0;
  __syncthreads();
  for (int i = c * ORDER_CHUNK + tid; i < nl && i < (c + 1) * ORDER_CHUNK; i += LOCAL_THREADS) {
    const int dg = k.cursor[l0 + i];
    atomicAdd(&s_h[dg < 255 ? dg : 255], 1);
  }
  __syncthreads();
  d.order_hist[((size_t)(w * 2 + kind) * d.order_chunks + c) * 256 + tid] = s_h[tid];
}
// histograms -> first internal index of every (chunk, bucket), largest degree first; grid (windows, kinds), 256 threads
__global__ void __launch_bounds__(256) setup_order_offsets(const __grid_constant__ LocalDev d) {
  __shared__ int s_start[256];
  const int w = blockIdx.x, kind = blockIdx.y, b = threadIdx.x;
  const KindDev& k = d.k[kind];
  const int nl = k.lm_begin[w + 1] - k.lm_begin[w];
  const int nc = (nl + ORDER_CHUNK - 1) / ORDER_CHUNK;
  int* h = d.order_hist + (size_t)(w * 2 + kind) * d.order_chunks * 256;
  int tot = 0;
  for (int c = 0; c < nc; ++c) tot += h[(size_t)c * 256 + b];
  s_start[b] = tot;
  __syncthreads();
  if (b == 0) {
    int run = 0;
    for (int q = 255; q >= 0; --q) {
      const int cnt = s_start[q];
      s_start[q] = run;
      run += cnt;
    }
  }
  __syncthreads();
  int run = s_start[b];
  for (int c = 0; c < nc; ++c) {
    const int cnt = h[(size_t)c * 256 + b];
    h[(size_t)c * 256 + b] = run;
    run += cnt;
  }
}
// placement of one chunk (four rounds of 256 landmarks, stable inside the chunk); grid (chunks, windows, kinds)
__global__ void __launch_bounds__(LOCAL_THREADS) setup_order_place(const __grid_constant__ LocalDev d) {
  __shared__ int s_start[256];
  __shared__ int s_deg[LOCAL_THREADS];
  const int c = blockIdx.x, w = blockIdx.y, kind = blockIdx.z, tid = threadIdx.x;
  const KindDev& k = d.k[kind];
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  if (c * ORDER_CHUNK >= nl) return;
  s_start[tid] = d.order_hist[((size_t)(w * 2 + kind) * d.order_chunks + c) * 256 + tid];
  __syncthreads();
  const int end = nl < (c + 1) * ORDER_CHUNK ? nl : (c + 1) * ORDER_CHUNK;
  for (int base = c * ORDER_CHUNK; base < end; base += LOCAL_THREADS) {
    const int i = base + tid;
    const int dg = i < end ? k.cursor[l0 + i] : -1;
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
    if (dg >= 0 && rank == total - 1) s_start[bk] += total;
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
  BA_CHECK(k.lm[e] >= 0 && l0 + k.lm[e] < k.n_lm);
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

// ---- pose-major lists of LARGE windows (hundreds of poses): setup_pose_lists costs O(poses x edges); here the edges
// are counted and scattered per pose with integer atomics (arbitrary order) and every list is then sorted ascending,
// which makes the result identical to setup_pose_lists' and deterministic.
// MODE 0 counts into pbeg (zeroed by the host), MODE 1 scatters behind pose_cursor (zeroed by the host);
// grid (edge chunks, windows, kinds)
template <int MODE>
__global__ void __launch_bounds__(LOCAL_THREADS) setup_pose_scatter(const __grid_constant__ LocalDev d) {
  const int w = blockIdx.y, kind = blockIdx.z;
  const KindDev& k = d.k[kind];
  const int e0 = edge_base(k, w), ne = edge_base(k, w + 1) - e0;
  const int i = blockIdx.x * LOCAL_THREADS + threadIdx.x;
  if (i >= ne) return;
  const int p = d.pose_begin[w] + (k.info[e0 + i] & 0xffff);
  if (MODE == 0) {
    atomicAdd(&k.pbeg[p], 1);
  } else {
    const int pos = atomicAdd(&d.pose_cursor[kind][p], 1);
    k.plist[k.pbeg[p] + pos] = e0 + i;
  }
}
// ascending sort of one pose's list: bitonic in shared memory up to POSE_SORT_CAP entries, rank sort through
// plist_tmp beyond; grid (poses, windows, kinds), 256 threads
constexpr int POSE_SORT_CAP = 8192;
__global__ void __launch_bounds__(LOCAL_THREADS) setup_pose_sort(const __grid_constant__ LocalDev d) {
  __shared__ int s_v[POSE_SORT_CAP];
  const int p = blockIdx.x, w = blockIdx.y, kind = blockIdx.z, tid = threadIdx.x;
  const KindDev& k = d.k[kind];
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0;
  if (p >= np) return;
  const int a = k.pbeg[p0 + p], n = k.pbeg[p0 + p + 1] - a;
  if (n < 2) return;
  int* seg = k.plist + a;
  if (n <= POSE_SORT_CAP) {
    int N = 1;
    while (N < n) N <<= 1;
    for (int i = tid; i < N; i += LOCAL_THREADS) s_v[i] = i < n ? seg[i] : 0x7fffffff;
    __syncthreads();
    for (int size = 2; size <= N; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < N / 2; i += LOCAL_THREADS) {
          const int lo = 2 * i - (i & (stride - 1)); // index with bit `stride` cleared
          const int hi = lo + stride;
          const bool up = (lo & size) == 0;
          const int x = s_v[lo], y = s_v[hi];
          if ((x > y) == up) {
            s_v[lo] = y;
            s_v[hi] = x;
          }
        }
        __syncthreads();
      }
    for (int i = tid; i < n; i += LOCAL_THREADS) seg[i] = s_v[i];
  } else {
    int* tmp = d.plist_tmp[kind] + a;
    for (int i = tid; i < n; i += LOCAL_THREADS) {
      const int v = seg[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += seg[j] < v ? 1 : 0;
      tmp[rank] = v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += LOCAL_THREADS) seg[i] = tmp[i];
  }
}

// ------------------------------------------------------------------------------------------------
// device helpers shared by the solve kernels (local_batched.cuh, local_tiled.cuh)
// ------------------------------------------------------------------------------------------------
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
    const double rs = rsqrt_nr(dsum); // 1 / l_jj without sqrt + divide (and their range tests)
    Lf[j * (j + 1) / 2 + j] = dsum * rs;
    inv[j] = rs;
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

template <int KIND>
BA_DEV void write_landmarks(const KindDev& k, int w, int chunk, int n_chunks) {
  using T = KT<KIND>;
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  for (int i = chunk * blockDim.x + threadIdx.x; i < nl; i += n_chunks * blockDim.x)
#pragma unroll
    for (int q = 0; q < T::SD; ++q) k.lm_out[(size_t)q * k.n_lm + l0 + k.orig[l0 + i]] = k.x[(size_t)q * k.n_lm + l0 + i];
}

} // namespace ba
