// local_batched.cuh — batched LocalmapOptimization: one kernel per LM phase over ALL windows.
//
// Same algorithm, data layout and device functions as local_kernel.cuh (the one-CTA-per-window
// persistent kernel), reorganised for throughput on large batches (BASELINE config C4: 1024
// windows): every phase is a grid over (window, chunk) work items with a small register
// footprint, so tens of thousands of CTAs keep HBM busy instead of 8 warps per SM chasing
// dependent loads. Each window carries its own Levenberg-Marquardt state machine in HBM
// (WinState); all accept/reject decisions are taken on the device (k_decide). The host only
// replays the fixed kernel sequence of a "super-step" and polls one integer (number of windows
// still iterating) every few super-steps.
//
//   super-step = k_linearize -> k_pose_blocks -> k_begin_trial -> k_schur_prep -> k_schur_reduce
//                -> k_solve -> k_backsub -> k_decide -> k_restore
// A window in stage NEED_LIN executes all of them (a new LM iteration), a window in stage
// NEED_TRIAL (rejected step, new lambda) skips the first two, a DONE window skips everything.
// All reductions use fixed work assignments and fixed-order trees: results are bitwise
// reproducible and independent of how many windows share the batch (or the GPU count).
#pragma once

#include "local_kernel.cuh"
#include "bcr_solver.cuh"

namespace ba {

constexpr int BT = 128; // threads per CTA of the per-landmark / per-pair kernels
constexpr int PAIRS_LM_MIN_NF = 64; // windows with more free poses build their pair lists from the landmarks
constexpr int BW = BT / 32;

enum { STAGE_NEED_LIN = 0, STAGE_NEED_TRIAL = 1, STAGE_DONE = 2 };

// Layout of a Z block in the batched path: COLUMN-major, Z[e][q][0..5] with q the landmark coordinate:
// kb_schur_reduce splits a pair entry by column (Z_i Z_j^T = sum_q z_i,q z_j,q^T).
constexpr int ZCOL = 6;
template <int KIND>
struct ZBlk {
  static constexpr int N = KT<KIND>::LD * ZCOL; // doubles per block: 18 (points), 24 (lines)
};

struct WinState {
  int stage, pass, it, qmax, n_sys, solve_ok, prep_fail, restore;
  int robust, iters, nf, pad;
  double lambda, ni, chi_cur, scale_pose, nact;
  DevStats st;
};

struct BatchDev { // extra state of the batched path (all in HBM)
  WinState* ws;   // [W]
  double* P_q;    // [NP][4]
  double* P_t;    // [NP][3]
  double* P_R;    // [NP][9]
  double* P_bq;   // [NP][4]
  double* P_bt;   // [NP][3]
  int* free_idx;  // [NP] window-local free index or -1
  int* pact;      // [NP] #active edges per pose in the current pass
  int* nf_begin;  // [W+1] offsets into the free-pose arrays
  int* pose_of;   // [NF] window-local pose of a free index
  int* sys_idx;   // [NF] block index in the reduced system or -1
  double* Hpp;    // [NF][21]
  double* bp;     // [NF][6]
  // Global BA (one problem, landmarks partitioned over ranks): kernels write their
  // rank-local partial sums to the *_w buffers, the host all-reduces them into the buffers every rank
  // reads. Without a communicator the *_w pointers alias the read buffers.
  int global;       // 1: this batch is one shard of a distributed problem (exactly one window)
  double* Hpp_w;    // [NF][21] followed by bp_w: one contiguous block of NF * 27 doubles
  double* bp_w;     // [NF][6]
  double* hs_part_w; // [W][Pmax][42] + 8 (tail[0] = landmark-Cholesky failure flag of this rank)
  int* pact_w;      // [NP]
  double* gs_w;     // [8] rank-local scalars: chi, nact, max diag | - | chi1, landmark scale
  double* gs;       // [8] reduced over ranks
  double* xp;     // [NF][6] pose increments by free index
  double* hs_part; // [W][Pmax][42]
  double* part;   // [W][C][4] per-chunk partial sums
  int* pair_beg;  // [W][2*Pmax+1] entry offsets: point entries, line entries per pair
  int2* pairs;    // pair entries (e_i, e_j) into the sorted edge arrays
  // compact list of the pose pairs that can be non-zero: every diagonal pair and every pair with entries
  // (on any rank). kb_schur_reduce / kb_solve / kb_assemble_dense run over this list, and hs_part is indexed
  // by list position -- a 2000-keyframe window has 2.0 M pairs of which 3 % are non-empty.
  int* ne_list;    // [W][Pmax] pair index, ascending
  int* n_ne;       // [W]
  int* ne_flag;    // [W][Pmax] scratch: 1 = keep
  int* diag_pos;   // [NF] list position of the diagonal pair of a free pose
  int2* blk;       // [W][Pmax] per compact pair and pass: x = si | sj << 16 (block coordinates in the reduced system) or -1,
                   //           y = free-pose offset f0 + fi for a diagonal pair, else -1 (kb_pair_blocks)
  int2* pairs_tmp; // scratch of the same size (landmark-driven builder of large windows), may be null
  int* pair_cursor; // [W][2*Pmax] scratch of that builder
  const long long* pair_base; // [W+1] region of each window inside `pairs`
  int* n_active;  // device counter for the host poll
  // dense-solve path (reduced systems too large for shared memory): per window a row-major n x n
  // matrix with the upper triangle filled (= column-major lower for dense_chol.cuh), rhs / solution, factorisation info
  double* dense_H;
  const long long* dense_off; // [W] offset of the window's matrix in dense_H
  double* dense_b;            // [6 * NF] by free-pose offset
  int* dense_info;            // [W]
  int Cp, Cl, C;  // landmark chunks per window: points, lines, total
  int Pmax, NFmax;
};

BA_DEV int pair_index(int fi, int fj, int nf) { return fi * nf - fi * (fi - 1) / 2 + (fj - fi); }
// inverse of pair_index: closed form + fix-up (windows of the dense-solve path have ~10^6 pairs)
BA_DEV void pair_decode(int p, int nf, int& fi, int& fj) {
  const double t = 2.0 * nf + 1.0;
  int row = (int)((t - sqrt(t * t - 8.0 * (double)p)) * 0.5);
  if (row < 0) row = 0;
  if (row > nf - 1) row = nf - 1;
  while (row > 0 && pair_index(row, row, nf) > p) --row;
  while (row + 1 < nf && pair_index(row + 1, row + 1, nf) <= p) ++row;
  fi = row;
  fj = row + (p - pair_index(row, row, nf));
}

// deterministic CTA reduction of N doubles per thread (BT threads); thread q < N ends with sum q
template <int N>
BA_DEV void cta_reduce(double* v, double* smem /* [BW][N] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < N; ++q) v[q] = warp_allreduce(v[q]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < N; ++q) smem[warp * N + q] = v[q];
  }
  __syncthreads();
}
template <int N>
BA_DEV double cta_reduce_get(const double* smem, int q) {
  double s = smem[q];
#pragma unroll
  for (int w = 1; w < BW; ++w) s += smem[w * N + q];
  return s;
}

// ------------------------------------------------------------------------------------------------
// setup: free-pose tables, optimiser poses, pair lists
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT) kb_init(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                              const __grid_constant__ LocalOpt o) {
  const int w = blockIdx.x, tid = threadIdx.x;
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0;
  const int f0 = b.nf_begin[w];
  if (tid == 0) {
    int nf = 0;
    for (int p = 0; p < np; ++p) {
      if (d.pose_fixed[p0 + p]) {
        b.free_idx[p0 + p] = -1;
      } else {
        b.free_idx[p0 + p] = nf;
        b.pose_of[f0 + nf] = p;
        ++nf;
      }
    }
    WinState& s = b.ws[w];
    s.stage = STAGE_DONE;
    s.pass = 0;
    s.it = s.qmax = s.n_sys = 0;
    s.solve_ok = 1;
    s.prep_fail = 0;
    s.restore = 0;
    s.robust = 1;
    s.iters = 0;
    s.nf = nf;
    s.lambda = 0;
    s.ni = 2;
    s.chi_cur = 0;
    s.scale_pose = 0;
    s.nact = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s.st.iters[i] = s.st.trials[i] = 0;
    s.st.edges_linearized = s.st.edges_evaluated = 0;
    s.st.final_chi2 = s.st.final_lambda = 0;
  }
  for (int p = tid; p < np; p += BT) {
    double q[4], R[9];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = b.P_q[4 * (size_t)(p0 + p) + i] = d.pose_tcw[(size_t)i * d.n_poses + p0 + p];
#pragma unroll
    for (int i = 0; i < 3; ++i) b.P_t[3 * (size_t)(p0 + p) + i] = d.pose_tcw[(size_t)(4 + i) * d.n_poses + p0 + p];
    quat_to_R(q, R);
#pragma unroll
    for (int i = 0; i < 9; ++i) b.P_R[9 * (size_t)(p0 + p) + i] = R[i];
  }
}

// pair lists: one warp per (window, pair); MODE 0 counts, MODE 1 fills (ballot compaction: deterministic)
template <int KIND, int MODE>
BA_DEV int pair_scan_kind(const LocalDev& d, const BatchDev& b, const KindDev& k, int w, int pose_i, int fj, bool diag,
                          int lane, int2* out) {
  const int l0 = k.lm_begin[w];
  const int p0 = d.pose_begin[w];
  const int a = k.pbeg[p0 + pose_i], e_end = k.pbeg[p0 + pose_i + 1];
  int n = 0;
  for (int base = a; base < e_end; base += 32) {
    const int it = base + lane;
    bool hit = false;
    int e = 0, e2 = 0;
    if (it < e_end) {
      e = k.plist[it];
      if (diag) {
        hit = true;
        e2 = e;
      } else {
        const int l = l0 + k.lm[e];
        const int sl = k.slot[(size_t)l * d.slot_stride + fj];
        if (sl != SLOT_NONE) {
          hit = true;
          e2 = k.ebeg[l] + sl;
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (MODE == 1 && hit) out[n + __popc(m & ((1u << lane) - 1))] = make_int2(e, e2);
    n += __popc(m);
  }
  return n;
}

// MODE 0 counts, MODE 1 fills; grid (ceil(P / BW), windows): one warp per (window, pair), so a single
// large window (C3: 210 pairs) is spread over many CTAs instead of one
template <int MODE>
__global__ void __launch_bounds__(BT) kb_pairs(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * BW + (threadIdx.x >> 5);
  const int nf = b.ws[w].nf;
  if (p >= nf * (nf + 1) / 2) return;
  if (MODE == 1 && (*d.err & LOCAL_ERR_DUP_EDGE)) return;
  int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1);
  int fi, fj;
  pair_decode(p, nf, fi, fj);
  const int pose_i = b.pose_of[b.nf_begin[w] + fi];
  const bool diag = fi == fj;
  if (MODE == 0) {
    const int n0 = pair_scan_kind<0, 0>(d, b, d.k[0], w, pose_i, fj, diag, lane, nullptr);
    const int n1 = pair_scan_kind<1, 0>(d, b, d.k[1], w, pose_i, fj, diag, lane, nullptr);
    if (lane == 0) {
      pb[2 * p] = n0;
      pb[2 * p + 1] = n1;
    }
  } else {
    BA_CHECK(pb[2 * p] >= b.pair_base[w] && pb[2 * p + 1] >= pb[2 * p] && pb[2 * p + 1] <= b.pair_base[w + 1]);
    pair_scan_kind<0, 1>(d, b, d.k[0], w, pose_i, fj, diag, lane, b.pairs + pb[2 * p]);
    pair_scan_kind<1, 1>(d, b, d.k[1], w, pose_i, fj, diag, lane, b.pairs + pb[2 * p + 1]);
  }
}

// exclusive scan of the per-pair counts of every window (one thread per window)
__global__ void __launch_bounds__(128) kb_pairs_scan(const __grid_constant__ LocalDev d,
                                                     const __grid_constant__ BatchDev b) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= d.n_windows) return;
  const int nf = b.ws[w].nf;
  const int np_pairs = nf * (nf + 1) / 2;
  int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1);
  long long run = b.pair_base[w];
  const long long cap = b.pair_base[w + 1];
  for (int q = 0; q < 2 * np_pairs; ++q) {
    const int c = pb[q];
    pb[q] = (int)run;
    run += c;
  }
  pb[2 * np_pairs] = (int)run;
  if (run > cap) { // only possible with duplicate (pose, landmark) edges, which are rejected anyway
    atomicOr(d.err, LOCAL_ERR_DUP_EDGE);
    for (int q = 0; q <= 2 * np_pairs; ++q) pb[q] = (int)b.pair_base[w];
  }
}

// ---- pair lists of LARGE windows (hundreds of poses, ~10^6 pairs, most of them empty) ----------------
// kb_pairs costs O(pairs x edges per pose); here the lists are built from the landmarks instead: a
// landmark with k free observers contributes k(k+1)/2 entries. MODE 0 counts them per pair, MODE 1 writes
// them behind an atomic cursor (arbitrary order), kb_pairs_sort then orders every list by its first edge
// (= landmark order), which makes the result identical to kb_pairs' and deterministic.
template <int KIND, int MODE>
BA_DEV void pairs_lm_kind(const LocalDev& d, const BatchDev& b, const KindDev& k, int w, int l) {
  const int p0 = d.pose_begin[w];
  const int nf = b.ws[w].nf;
  int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1);
  int* cur = b.pair_cursor + (size_t)w * 2 * b.Pmax;
  const int ea = k.ebeg[l], eb = k.ebeg[l + 1];
  for (int a = ea; a < eb; ++a) {
    const int fa = b.free_idx[p0 + (k.info[a] & 0xffff)];
    if (fa < 0) continue;
    for (int c = a; c < eb; ++c) {
      const int fc = c == a ? fa : b.free_idx[p0 + (k.info[c] & 0xffff)];
      if (fc < 0) continue;
      const bool swap = fc < fa;
      const int q = 2 * pair_index(swap ? fc : fa, swap ? fa : fc, nf) + KIND;
      if (MODE == 0) {
        atomicAdd(&pb[q], 1);
      } else {
        const int pos = atomicAdd(&cur[q], 1);
        b.pairs[pb[q] + pos] = make_int2(swap ? c : a, swap ? a : c);
      }
    }
  }
}

// grid (Cp + Cl, windows), one thread per landmark
template <int MODE>
__global__ void __launch_bounds__(BT) kb_pairs_lm(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y, c = blockIdx.x;
  if (MODE == 1 && (*d.err & LOCAL_ERR_DUP_EDGE)) return;
  if (c < b.Cp) {
    const KindDev& k = d.k[0];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = c * BT + threadIdx.x;
    if (i < nl) pairs_lm_kind<0, MODE>(d, b, k, w, l0 + i);
  } else {
    const KindDev& k = d.k[1];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = (c - b.Cp) * BT + threadIdx.x;
    if (i < nl) pairs_lm_kind<1, MODE>(d, b, k, w, l0 + i);
  }
}

// Exclusive scan of a chunk of SCAN_ITEMS * 1024 values by a CTA of 1024 threads: thread t owns SCAN_ITEMS consecutive
// values v[] (replaced by their exclusive prefixes within the chunk); returns the chunk total to every thread.
// s_warp: 32 ints of shared memory. Two barriers per chunk.
constexpr int SCAN_ITEMS = 8;
BA_DEV int cta_scan_chunk(int (&v)[SCAN_ITEMS], int* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int sum = 0;
#pragma unroll
  for (int u = 0; u < SCAN_ITEMS; ++u) {
    const int x = v[u];
    v[u] = sum;
    sum += x;
  }
  int x = sum; // inclusive warp scan of the thread sums
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads(); // s_warp of the previous chunk has been read
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  int t = s_warp[lane]; // every warp scans the 32 warp totals itself (no third barrier)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, t, o);
    if (lane >= o) t += y;
  }
  const int total = __shfl_sync(0xffffffffu, t, 31);
  const int warp_before = __shfl_sync(0xffffffffu, t, warp > 0 ? warp - 1 : 0);
  const int before = (warp ? warp_before : 0) + x - sum;
#pragma unroll
  for (int u = 0; u < SCAN_ITEMS; ++u) v[u] += before;
  return total;
}

// ---- device-wide versions of the two scans for LARGE windows (millions of pairs: the one-CTA loops below took
// 2.7 + 0.9 ms on C5): chunk sums -> scan of the chunk sums -> rescan of every chunk with its offset. WHICH 0: the
// 2P pair counts (pair_beg), 1: the keep flags (-> compact list). scratch: [W][n_chunks] ints. grid (chunks, windows).
constexpr int SCAN_CHUNK = 1024 * SCAN_ITEMS;
template <int WHICH>
BA_DEV int big_scan_len(const BatchDev& b, int w) {
  const int nf = b.ws[w].nf;
  return WHICH == 0 ? nf * (nf + 1) : nf * (nf + 1) / 2;
}
template <int WHICH>
BA_DEV int big_scan_value(const BatchDev& b, int w, int i) {
  return WHICH == 0 ? b.pair_beg[(size_t)w * (2 * b.Pmax + 1) + i] : (b.ne_flag[(size_t)w * b.Pmax + i] ? 1 : 0);
}
template <int WHICH>
__global__ void __launch_bounds__(1024) kb_big_scan_sums(const __grid_constant__ BatchDev b, int* scratch, int n_chunks) {
  __shared__ int s_warp[32];
  const int w = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = big_scan_len<WHICH>(b, w);
  const int i0 = chunk * SCAN_CHUNK + tid * SCAN_ITEMS;
  int sum = 0;
#pragma unroll
  for (int u = 0; u < SCAN_ITEMS; ++u) sum += i0 + u < n ? big_scan_value<WHICH>(b, w, i0 + u) : 0;
  sum = __reduce_add_sync(0xffffffffu, sum);
  if (lane == 0) s_warp[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    const int t = __reduce_add_sync(0xffffffffu, s_warp[lane]);
    if (lane == 0) scratch[(size_t)w * n_chunks + chunk] = t;
  }
}
// grid = windows, 1024 threads: exclusive scan of the chunk sums in place (n_chunks <= SCAN_CHUNK), totals
template <int WHICH>
__global__ void __launch_bounds__(1024) kb_big_scan_offsets(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                            int* scratch, int n_chunks) {
  __shared__ int s_warp[32];
  const int w = blockIdx.x, tid = threadIdx.x;
  int* part = scratch + (size_t)w * n_chunks;
  const int i0 = tid * SCAN_ITEMS;
  int v[SCAN_ITEMS];
#pragma unroll
  for (int u = 0; u < SCAN_ITEMS; ++u) v[u] = i0 + u < n_chunks ? part[i0 + u] : 0;
  const int total = cta_scan_chunk(v, s_warp);
#pragma unroll
  for (int u = 0; u < SCAN_ITEMS; ++u)
    if (i0 + u < n_chunks) part[i0 + u] = v[u];
  if (tid == 0) {
    if (WHICH == 0) {
      const long long run = b.pair_base[w] + total;
      b.pair_beg[(size_t)w * (2 * b.Pmax + 1) + big_scan_len<0>(b, w)] = (int)run;
      if (run > b.pair_base[w + 1]) atomicOr(d.err, LOCAL_ERR_DUP_EDGE); // capacity: only with duplicate edges
    } else {
      b.n_ne[w] = total;
    }
  }
}
template <int WHICH>
__global__ void __launch_bounds__(1024) kb_big_scan_apply(const __grid_constant__ BatchDev b, const int* scratch, int n_chunks) {
  __shared__ int s_warp[32];
  const int w = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int n = big_scan_len<WHICH>(b, w);
  if (chunk * SCAN_CHUNK >= n) return;
  const int i0 = chunk * SCAN_CHUNK + tid * SCAN_ITEMS;
  int v[SCAN_ITEMS], in[SCAN_ITEMS];
#pragma unroll
  for (int u = 0; u < SCAN_ITEMS; ++u) in[u] = v[u] = i0 + u < n ? big_scan_value<WHICH>(b, w, i0 + u) : 0;
  cta_scan_chunk(v, s_warp);
  const int off = scratch[(size_t)w * n_chunks + chunk];
  if (WHICH == 0) {
    int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1);
    const long long base = b.pair_base[w] + off;
#pragma unroll
    for (int u = 0; u < SCAN_ITEMS; ++u)
      if (i0 + u < n) pb[i0 + u] = (int)(base + v[u]);
  } else {
    const int nf = b.ws[w].nf, f0 = b.nf_begin[w];
    int* list = b.ne_list + (size_t)w * b.Pmax;
#pragma unroll
    for (int u = 0; u < SCAN_ITEMS; ++u) {
      if (!in[u]) continue;
      const int p = i0 + u, pos = off + v[u];
      list[pos] = p;
      int fi, fj;
      pair_decode(p, nf, fi, fj);
      if (fi == fj) b.diag_pos[f0 + fi] = pos;
    }
  }
}

// exclusive scan of one window's 2P counts by a whole CTA (1024 threads x 8 values per chunk, running total in a
// register of every thread); grid = windows
__global__ void __launch_bounds__(1024) kb_pairs_scan_cta(const __grid_constant__ LocalDev d,
                                                          const __grid_constant__ BatchDev b) {
  __shared__ int s_warp[32];
  const int w = blockIdx.x, tid = threadIdx.x;
  const int nf = b.ws[w].nf;
  const int n = nf * (nf + 1); // 2 * pairs
  int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1);
  long long run = b.pair_base[w];
  for (int base = 0; base < n; base += 1024 * SCAN_ITEMS) {
    const int i0 = base + tid * SCAN_ITEMS;
    int v[SCAN_ITEMS];
#pragma unroll
    for (int u = 0; u < SCAN_ITEMS; ++u) v[u] = i0 + u < n ? pb[i0 + u] : 0;
    const int total = cta_scan_chunk(v, s_warp);
#pragma unroll
    for (int u = 0; u < SCAN_ITEMS; ++u)
      if (i0 + u < n) pb[i0 + u] = (int)(run + v[u]);
    run += total;
  }
  if (tid == 0) {
    pb[n] = (int)run;
    if (run > b.pair_base[w + 1]) atomicOr(d.err, LOCAL_ERR_DUP_EDGE); // capacity: only with duplicate edges
  }
}

// one warp per (window, pair): order the pair's point and line lists by their first edge (rank sort
// through pairs_tmp: the keys of a list are distinct)
// grid (ceil(max compact pairs / BW), windows): only the pairs of the compact list are visited (3 % of the 2 M pairs
// of a 2000-keyframe chain). Lists of up to PAIRS_SORT_SHORT entries: rank sort by one warp; longer ones (the diagonal
// pair of a pose lists every edge of the pose: ~2000 entries on C5, where the quadratic rank sort cost 4.6 ms) are
// left to kb_pairs_sort_long.
constexpr int PAIRS_SORT_SHORT = 128;
constexpr int PAIRS_SORT_CAP = 4096; // entries of the shared-memory bitonic sort (32 KB); beyond: CTA-wide rank sort
__global__ void __launch_bounds__(BT) kb_pairs_sort(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y, lane = threadIdx.x & 31;
  const int li = blockIdx.x * BW + (threadIdx.x >> 5);
  if (li >= b.n_ne[w] || (*d.err & LOCAL_ERR_DUP_EDGE)) return;
  const int p = b.ne_list[(size_t)w * b.Pmax + li];
  const int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1) + 2 * p;
  for (int kind = 0; kind < 2; ++kind) {
    const int beg = pb[kind], n = pb[kind + 1] - beg;
    if (n < 2 || n > PAIRS_SORT_SHORT) continue;
    int2* seg = b.pairs + beg;
    int2* tmp = b.pairs_tmp + beg;
    for (int i = lane; i < n; i += 32) {
      const int2 v = seg[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += seg[j].x < v.x ? 1 : 0;
      tmp[rank] = v;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) seg[i] = tmp[i];
    __syncwarp();
  }
}

// one CTA per pair of the compact list: lists longer than PAIRS_SORT_SHORT by a bitonic sort in shared memory (keys
// are distinct, so every sort gives the same order); grid (max compact pairs, windows), 256 threads
__global__ void __launch_bounds__(256) kb_pairs_sort_long(const __grid_constant__ LocalDev d,
                                                          const __grid_constant__ BatchDev b) {
  __shared__ int2 buf[PAIRS_SORT_CAP];
  const int w = blockIdx.y, li = blockIdx.x, tid = threadIdx.x;
  if (li >= b.n_ne[w] || (*d.err & LOCAL_ERR_DUP_EDGE)) return;
  const int p = b.ne_list[(size_t)w * b.Pmax + li];
  const int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1) + 2 * p;
  for (int kind = 0; kind < 2; ++kind) {
    const int beg = pb[kind], n = pb[kind + 1] - beg;
    if (n <= PAIRS_SORT_SHORT) continue; // (uniform over the CTA)
    int2* seg = b.pairs + beg;
    if (n > PAIRS_SORT_CAP) { // rank sort through pairs_tmp, whole CTA
      int2* tmp = b.pairs_tmp + beg;
      for (int i = tid; i < n; i += 256) {
        const int2 v = seg[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += seg[j].x < v.x ? 1 : 0;
        tmp[rank] = v;
      }
      __syncthreads();
      for (int i = tid; i < n; i += 256) seg[i] = tmp[i];
      __syncthreads();
      continue;
    }
    int m = 256;
    while (m < n) m <<= 1;
    for (int i = tid; i < m; i += 256) buf[i] = i < n ? seg[i] : make_int2(0x7fffffff, 0);
    __syncthreads();
    for (int k2 = 2; k2 <= m; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < m; i += 256) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int2 a = buf[i], c = buf[ixj];
            const bool up = (i & k2) == 0;
            if ((a.x > c.x) == up) {
              buf[i] = c;
              buf[ixj] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < n; i += 256) seg[i] = buf[i];
    __syncthreads();
  }
}

// keep-flag of every pair of a window: diagonal, or with entries; grid (ceil(Pmax / 256), windows)
__global__ void __launch_bounds__(256) kb_pairs_flag(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y, p = blockIdx.x * 256 + threadIdx.x;
  if (p >= b.Pmax) return;
  const int nf = b.ws[w].nf;
  int keep = 0;
  if (p < nf * (nf + 1) / 2) {
    const int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1) + 2 * p;
    int fi, fj;
    pair_decode(p, nf, fi, fj);
    keep = (fi == fj || pb[2] - pb[0] > 0) ? 1 : 0;
  }
  b.ne_flag[(size_t)w * b.Pmax + p] = keep;
}

// flags -> ascending compact list (CTA scan, 1024 threads x 8 flags per chunk); grid = windows
__global__ void __launch_bounds__(1024) kb_pairs_compact(const __grid_constant__ LocalDev d,
                                                         const __grid_constant__ BatchDev b) {
  __shared__ int s_warp[32];
  const int w = blockIdx.x, tid = threadIdx.x;
  const int nf = b.ws[w].nf, f0 = b.nf_begin[w];
  const int n = nf * (nf + 1) / 2;
  const int* flag = b.ne_flag + (size_t)w * b.Pmax;
  int* list = b.ne_list + (size_t)w * b.Pmax;
  int run = 0;
  for (int base = 0; base < n; base += 1024 * SCAN_ITEMS) {
    const int p0 = base + tid * SCAN_ITEMS;
    int v[SCAN_ITEMS], keep[SCAN_ITEMS];
#pragma unroll
    for (int u = 0; u < SCAN_ITEMS; ++u) keep[u] = v[u] = p0 + u < n ? (flag[p0 + u] ? 1 : 0) : 0;
    const int total = cta_scan_chunk(v, s_warp);
#pragma unroll
    for (int u = 0; u < SCAN_ITEMS; ++u) {
      if (!keep[u]) continue;
      const int p = p0 + u, pos = run + v[u];
      list[pos] = p;
      int fi, fj;
      pair_decode(p, nf, fi, fj);
      if (fi == fj) b.diag_pos[f0 + fi] = pos;
    }
    run += total;
  }
  if (tid == 0) b.n_ne[w] = run;
}

// largest free-pose index distance of a pose pair that shares a landmark (dense path: is the reduced
// system banded?); grid (ceil(Pmax / 256), windows)
__global__ void __launch_bounds__(256) kb_pair_band(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                    int* band) {
  const int w = blockIdx.y, p = blockIdx.x * 256 + threadIdx.x;
  const int nf = b.ws[w].nf;
  if (p >= nf * (nf + 1) / 2) return;
  const int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1) + 2 * p;
  if (pb[2] - pb[0] <= 0) return;
  int fi, fj;
  pair_decode(p, nf, fi, fj);
  atomicMax(&band[w], fj - fi);
}

// ------------------------------------------------------------------------------------------------
// pass begin: active sets (§9.12), reduced-system indices
// ------------------------------------------------------------------------------------------------
// Active edges per pose are counted into pact_w (zeroed by the host), one thread per landmark;
// grid (Cp + Cl, windows). Inactive landmarks get their Z blocks and y zeroed: kb_schur_reduce and
// kb_backsub read them for every listed edge without testing `act`.
template <int KIND>
BA_DEV void mark_active_one(const LocalDev& d, const BatchDev& b, const KindDev& k, int w, int l) {
  const int p0 = d.pose_begin[w];
  int any = 0;
  for (int e = k.ebeg[l]; e < k.ebeg[l + 1]; ++e) {
    if (k.lvl[e]) continue;
    any = 1;
    atomicAdd(&b.pact_w[p0 + (k.info[e] & 0xffff)], 1);
  }
  k.act[l] = (uint8_t)any;
  if (!any) {
    for (int e = k.ebeg[l]; e < k.ebeg[l + 1]; ++e)
#pragma unroll
      for (int q = 0; q < ZBlk<KIND>::N; ++q) k.Z[(size_t)e * ZBlk<KIND>::N + q] = 0.0;
#pragma unroll
    for (int q = 0; q < KT<KIND>::LD; ++q) k.y[(size_t)q * k.n_lm + l] = 0.0;
  }
}

__global__ void __launch_bounds__(BT) kb_mark_active(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y, c = blockIdx.x;
  if (c < b.Cp) {
    const KindDev& k = d.k[0];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = c * BT + threadIdx.x;
    if (i < nl) mark_active_one<0>(d, b, k, w, l0 + i);
  } else {
    const KindDev& k = d.k[1];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = (c - b.Cp) * BT + threadIdx.x;
    if (i < nl) mark_active_one<1>(d, b, k, w, l0 + i);
  }
}

// reduced-system indices and LM state of the pass from the (all-reduced, in global mode) counts; one thread per window
__global__ void __launch_bounds__(128) kb_begin_pass(const __grid_constant__ LocalDev d,
                                                     const __grid_constant__ BatchDev b,
                                                     const __grid_constant__ LocalOpt o, int pass) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= d.n_windows) return;
  const int p0 = d.pose_begin[w];
  WinState& s = b.ws[w];
  const int f0 = b.nf_begin[w];
  int nsys = 0;
  for (int fi = 0; fi < s.nf; ++fi) b.sys_idx[f0 + fi] = b.pact[p0 + b.pose_of[f0 + fi]] > 0 ? nsys++ : -1;
  s.n_sys = nsys;
  s.pass = pass;
  s.it = 0;
  s.qmax = 0;
  s.lambda = 0;
  s.ni = 2;
  s.robust = pass == 0 ? 1 : 0;
  s.iters = o.iters[pass];
  s.restore = 0;
  s.stage = s.iters > 0 ? STAGE_NEED_LIN : STAGE_DONE;
}

// block coordinates of the compact pairs in this pass's reduced system; grid (ceil(Pmax / 256), windows)
__global__ void __launch_bounds__(256) kb_pair_blocks(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y, li = blockIdx.x * 256 + threadIdx.x;
  if (li >= b.n_ne[w]) return;
  const int f0 = b.nf_begin[w];
  int fi, fj;
  pair_decode(b.ne_list[(size_t)w * b.Pmax + li], b.ws[w].nf, fi, fj);
  const int si = b.sys_idx[f0 + fi], sj = b.sys_idx[f0 + fj];
  b.blk[(size_t)w * b.Pmax + li] = make_int2((si < 0 || sj < 0) ? -1 : (si | (sj << 16)), fi == fj ? f0 + fi : -1);
}

// ------------------------------------------------------------------------------------------------
// K1: linearise, one thread per landmark (edges of a landmark in g2o order)
// ------------------------------------------------------------------------------------------------
template <int KIND>
BA_DEV void linearize_one(const LocalDev& d, const BatchDev& b, const LocalOpt& o, const KindDev& k, int w, int l,
                          bool robust, double& chi_part, double& maxdiag_part, double& nact_part) {
  // The pose x landmark block W = Jp^T (rho1 Omega) Jl is NOT materialised here: kb_schur_prep
  // recomputes it from the edge record (~200 flop) instead of a 144/192 B round trip through HBM.
  using T = KT<KIND>;
  const int p0 = d.pose_begin[w];
  double X[T::SD];
  load_lm<KIND>(k, l, X);
  double H[T::HD], bb[T::LD];
#pragma unroll
  for (int q = 0; q < T::HD; ++q) H[q] = 0;
#pragma unroll
  for (int q = 0; q < T::LD; ++q) bb[q] = 0;
  const int ea = k.ebeg[l], eb = k.ebeg[l + 1];
  // the record of the next edge is requested before this one is evaluated (the landmarks of a warp have the same
  // degree, so the loop is latency- rather than divergence-bound)
  int info_n = 0;
  unsigned char lvl_n = 1;
  double m_n[T::MD];
  if (ea < eb) {
    lvl_n = k.lvl[ea];
    info_n = k.info[ea];
    load_edge<KIND>(k, ea, m_n);
  }
  for (int e = ea; e < eb; ++e) {
    const int info = info_n;
    const bool skip = lvl_n != 0;
    double m[T::MD];
#pragma unroll
    for (int q = 0; q < T::MD; ++q) m[q] = m_n[q];
    if (e + 1 < eb) {
      lvl_n = k.lvl[e + 1];
      info_n = k.info[e + 1];
      load_edge<KIND>(k, e + 1, m_n);
    }
    if (skip) continue;
    const int p = info & 0xffff;
    BA_CHECK(p < d.pose_begin[w + 1] - d.pose_begin[w]);
    const bool stereo = (info >> 30) & 1;
    Cam cam;
    load_cam(d.cameras, (info >> 16) & 0xff, cam);
    double r[4], Jp[24], Jl[16];
    eval_edge<KIND, true>(cam, o.bf_float, stereo, b.P_R + 9 * (size_t)(p0 + p), b.P_t + 3 * (size_t)(p0 + p), X, m, r,
                          Jp, Jl);
    const double c2 = edge_chi2<KIND>(r);
    k.chi2[e] = c2;
    double wgt = 1.0;
    const double rho0 = robust ? huber(c2, o.delta[2 * KIND + (stereo ? 1 : 0)], wgt) : c2;
    chi_part += rho0;
    nact_part += 1.0;
    const double wo = (KIND == 0 ? 1.0 : 0.1) * wgt;
    int q = 0;
#pragma unroll
    for (int a = 0; a < T::LD; ++a) {
      double g = 0;
#pragma unroll
      for (int rr = 0; rr < T::ROWS; ++rr) g += Jl[rr * T::LD + a] * r[rr];
      bb[a] -= wo * g;
#pragma unroll
      for (int c = a; c < T::LD; ++c) {
        double h = 0;
#pragma unroll
        for (int rr = 0; rr < T::ROWS; ++rr) h += Jl[rr * T::LD + a] * Jl[rr * T::LD + c];
        H[q++] += wo * h;
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
  for (int a = 0; a < T::LD; ++a) k.b[(size_t)a * k.n_lm + l] = bb[a];
}

// grid (chunks of this kind, windows): the chunk index is the fastest-varying block index so the
// CTAs of one window run together and share its poses / edges in L1 / L2
template <int KIND>
__global__ void __launch_bounds__(BT, KIND == 0 ? 5 : 2) kb_linearize(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                   const __grid_constant__ LocalOpt o) {
  __shared__ double red[BW * 3];
  const int w = blockIdx.y;
  const int c = blockIdx.x + (KIND ? b.Cp : 0); // slot in the per-window partial-sum table
  const WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_LIN) return;
  const bool robust = s.robust;
  double v[3] = {0, 0, 0}; // chi, nact, maxdiag
  double mx = 0;
  {
    const KindDev& k = d.k[KIND];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i < nl && k.act[l0 + i]) linearize_one<KIND>(d, b, o, k, w, l0 + i, robust, v[0], mx, v[1]);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  v[2] = 0;
  cta_reduce<3>(v, red);
  __shared__ double redmx[BW];
  if ((threadIdx.x & 31) == 0) redmx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m2 = redmx[0];
#pragma unroll
    for (int q = 1; q < BW; ++q) m2 = fmax(m2, redmx[q]);
    double* pp = b.part + ((size_t)w * b.C + c) * 4;
    pp[0] = cta_reduce_get<3>(red, 0);
    pp[1] = cta_reduce_get<3>(red, 1);
    pp[2] = m2;
  }
}

// K1': Hpp, bp per free pose; one CTA per (window, free pose)
template <int KIND>
BA_DEV void pose_block_kind(const LocalDev& d, const BatchDev& b, const LocalOpt& o, const KindDev& k, int w, int p,
                            bool robust, double* acc) {
  using T = KT<KIND>;
  const int l0 = k.lm_begin[w];
  const int p0 = d.pose_begin[w];
  const int a = k.pbeg[p0 + p], e_end = k.pbeg[p0 + p + 1];
  const double* R = b.P_R + 9 * (size_t)(p0 + p);
  const double* t = b.P_t + 3 * (size_t)(p0 + p);
  // The pose-major list is an indirection (plist -> edge -> landmark -> state): three dependent loads per edge. The
  // edge index is requested two iterations ahead and the edge header (level, info, landmark) one ahead, so that only
  // the landmark state / measurement gather of the current edge is waited for.
  int e_n = 0, e_nn = 0, info_n = 0, lm_n = 0;
  unsigned char lvl_n = 1;
  {
    const int it0 = a + threadIdx.x;
    if (it0 < e_end) {
      e_n = k.plist[it0];
      lvl_n = k.lvl[e_n];
      info_n = k.info[e_n];
      lm_n = k.lm[e_n];
    }
    if (it0 + BT < e_end) e_nn = k.plist[it0 + BT];
  }
  for (int it = a + threadIdx.x; it < e_end; it += BT) {
    const int e = e_n, info = info_n, lmi = lm_n;
    const bool skip = lvl_n != 0;
    if (it + BT < e_end) {
      e_n = e_nn;
      lvl_n = k.lvl[e_n];
      info_n = k.info[e_n];
      lm_n = k.lm[e_n];
    }
    if (it + 2 * BT < e_end) e_nn = k.plist[it + 2 * BT];
    if (skip) continue;
    const bool stereo = (info >> 30) & 1;
    Cam cam;
    load_cam(d.cameras, (info >> 16) & 0xff, cam);
    double X[T::SD], m[T::MD], r[4], Jp[24], Jl[16];
    load_lm<KIND>(k, l0 + lmi, X);
    load_edge<KIND>(k, e, m);
    eval_edge<KIND, true>(cam, o.bf_float, stereo, R, t, X, m, r, Jp, Jl);
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

__global__ void __launch_bounds__(BT, 4) kb_pose_blocks(const __grid_constant__ LocalDev d,
                                                     const __grid_constant__ BatchDev b,
                                                     const __grid_constant__ LocalOpt o) {
  __shared__ double red[BW * 27];
  const int w = blockIdx.y, fi = blockIdx.x;
  const WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_LIN || fi >= s.nf) return;
  const int f0 = b.nf_begin[w];
  if (b.sys_idx[f0 + fi] < 0) return;
  const int p = b.pose_of[f0 + fi];
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0;
  pose_block_kind<0>(d, b, o, d.k[0], w, p, s.robust, acc);
  pose_block_kind<1>(d, b, o, d.k[1], w, p, s.robust, acc);
  cta_reduce<27>(acc, red);
  if (threadIdx.x < 27) {
    const double v = cta_reduce_get<27>(red, threadIdx.x);
    if (threadIdx.x < 21) b.Hpp_w[(size_t)(f0 + fi) * 21 + threadIdx.x] = v;
    else b.bp_w[(size_t)(f0 + fi) * 6 + threadIdx.x - 21] = v;
  }
}

// one WARP per window: chi0, lambda init (computeLambdaInit), stage NEED_LIN -> NEED_TRIAL. Lanes stride the chunk
// partials / the free poses and a fixed butterfly adds them (deterministic); grid ceil(W / 4), 128 threads
__global__ void __launch_bounds__(128) kb_begin_trial(const __grid_constant__ LocalDev d,
                                                      const __grid_constant__ BatchDev b) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= d.n_windows) return;
  WinState& s = b.ws[w];
  const int stage = s.stage, it = s.it, nf = s.nf;
  __syncwarp();
  if (lane == 0) {
    s.restore = 0;
    s.prep_fail = 0;
  }
  if (stage != STAGE_NEED_LIN) return;
  double chi = 0, nact = 0, mx = 0;
  if (b.global) { // sums over every rank's landmarks (kb_global_sums + all-reduce)
    chi = b.gs[0];
    nact = b.gs[1];
    mx = b.gs[2];
  } else {
    // (point and line chunks are summed separately: the slot of the first line chunk, Cp, depends on the largest
    // window of the batch, and the result of a window must not)
    double chi_l = 0, nact_l = 0;
    for (int c = lane; c < b.Cp; c += 32) {
      const double* pp = b.part + ((size_t)w * b.C + c) * 4;
      chi += pp[0];
      nact += pp[1];
      mx = fmax(mx, pp[2]);
    }
    for (int c = b.Cp + lane; c < b.C; c += 32) {
      const double* pp = b.part + ((size_t)w * b.C + c) * 4;
      chi_l += pp[0];
      nact_l += pp[1];
      mx = fmax(mx, pp[2]);
    }
    chi = warp_allreduce(chi) + warp_allreduce(chi_l);
    nact = warp_allreduce(nact) + warp_allreduce(nact_l);
  }
  if (it == 0) {
    const int f0 = b.nf_begin[w];
    for (int idx = lane; idx < nf * 6; idx += 32) {
      const int fi = idx / 6, i = idx - 6 * fi;
      if (b.sys_idx[f0 + fi] < 0) continue;
      mx = fmax(mx, fabs(b.Hpp[(size_t)(f0 + fi) * 21 + up6(i, i)]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane != 0) return;
  if (nact == 0.0) { // no active edge: optimize() returns without iterating
    s.stage = STAGE_DONE;
    return;
  }
  if (it == 0) {
    s.lambda = 1e-5 * mx;
    s.ni = 2;
  }
  s.chi_cur = chi;
  s.nact = nact;
  s.qmax = 0;
  s.st.edges_linearized += (long long)nact;
  s.st.edges_evaluated += (long long)nact;
  s.stage = STAGE_NEED_TRIAL;
}

// K3: per landmark L = chol(Hll + lambda), y = L^-1 bl, Z_e = W_e L^-T
template <int KIND>
BA_DEV void schur_prep_one(const LocalDev& d, const BatchDev& b, const LocalOpt& o, const KindDev& k, int w, int l,
                           double lambda, bool robust, int& fail) {
  using T = KT<KIND>;
  constexpr int LD = T::LD;
  const int p0 = d.pose_begin[w], f0 = b.nf_begin[w];
  double X[T::SD];
  load_lm<KIND>(k, l, X);
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
    const int info = k.info[e];
    const int p = info & 0xffff;
    BA_CHECK(p < d.pose_begin[w + 1] - d.pose_begin[w]);
    const int fi = b.free_idx[p0 + p];
    if (fi < 0 || b.sys_idx[f0 + fi] < 0) continue;
    double* Ze = k.Z + (size_t)e * ZBlk<KIND>::N;
    double wv[T::WD];
    if (k.lvl[e]) { // excluded edge (level 1): no contribution
#pragma unroll
      for (int q = 0; q < ZBlk<KIND>::N; q += 2) *reinterpret_cast<double2*>(Ze + q) = make_double2(0.0, 0.0);
      continue;
    }
    { // W = Jp^T (rho1 Omega) Jl, recomputed from the edge record
      const bool stereo = (info >> 30) & 1;
      Cam cam;
      load_cam(d.cameras, (info >> 16) & 0xff, cam);
      double m[T::MD], r[4], Jp[24], Jl[16];
      load_edge<KIND>(k, e, m);
      eval_edge<KIND, true>(cam, o.bf_float, stereo, b.P_R + 9 * (size_t)(p0 + p), b.P_t + 3 * (size_t)(p0 + p), X, m, r,
                            Jp, Jl);
      double wgt = 1.0;
      if (robust) huber(edge_chi2<KIND>(r), o.delta[2 * KIND + (stereo ? 1 : 0)], wgt);
      const double wo = (KIND == 0 ? 1.0 : 0.1) * wgt;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = 0; c < LD; ++c) {
          double h = 0;
#pragma unroll
          for (int rr = 0; rr < T::ROWS; ++rr) h += Jp[rr * 6 + a] * Jl[rr * LD + c];
          wv[a * LD + c] = wo * h;
        }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double z[LD];
#pragma unroll
      for (int c = 0; c < LD; ++c) {
        double v = wv[a * LD + c];
#pragma unroll
        for (int p = 0; p < c; ++p) v -= Lf[c * (c + 1) / 2 + p] * z[p];
        z[c] = v * inv[c];
      }
#pragma unroll
      for (int c = 0; c < LD; ++c) wv[a * LD + c] = z[c];
    }
    // column-major store (see ZCOL)
#pragma unroll
    for (int c = 0; c < LD; ++c) {
      double2* col = reinterpret_cast<double2*>(Ze + c * ZCOL);
      col[0] = make_double2(wv[0 * LD + c], wv[1 * LD + c]);
      col[1] = make_double2(wv[2 * LD + c], wv[3 * LD + c]);
      col[2] = make_double2(wv[4 * LD + c], wv[5 * LD + c]);
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(BT) kb_schur_prep(const __grid_constant__ LocalDev d,
                                                    const __grid_constant__ BatchDev b,
                                                    const __grid_constant__ LocalOpt o) {
  const int w = blockIdx.y;
  WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_TRIAL) return;
  int fail = 0;
  const KindDev& k = d.k[KIND];
  const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
  const int i = blockIdx.x * BT + threadIdx.x;
  if (i < nl && k.act[l0 + i]) schur_prep_one<KIND>(d, b, o, k, w, l0 + i, s.lambda, s.robust, fail);
  if (fail) atomicOr(&s.prep_fail, 1);
}

// K3': one WARP per (window, pose pair): sum_e Z_i Z_j^T (and Z_i y on the diagonal) over the pair list.
// A lane owns one COLUMN q of one pair entry per iteration and adds the rank-1 update z_i,q z_j,q^T to its
// private accumulators (all lanes' accumulators are summed at the end anyway). The loop is a gather of
// 64-byte columns at data-dependent addresses and its first versions were latency-bound: 84 accumulator
// registers per lane leave room for ~12 warps per SM, too few to cover the L2 round trip with one or two
// loads in flight per lane. So the columns travel global -> shared memory with cp.async into a
// lane-private ring of RED_STAGES stages (the 227 KB of shared memory hold the in-flight data instead of
// registers), entry records are fetched one further iteration ahead, and a lane only ever reads what it
// copied itself, so no barrier is needed.
#ifndef RED_STAGES_N
#define RED_STAGES_N 4
#endif
constexpr int RED_STAGES = RED_STAGES_N;
constexpr int RED_RING = RED_STAGES * 6 * 32; // double2 per warp: [stage][piece][lane], conflict-free both ways

template <int KIND>
BA_DEV void schur_pair_entries(const KindDev& k, int w, const int2* ent, int n, bool diag, int lane, double* acc,
                               double2* ring) {
  constexpr int LD = KT<KIND>::LD, ZB = ZBlk<KIND>::N, S = RED_STAGES;
  const int l0 = k.lm_begin[w];
  const int nsub = n * LD;
  const int nit = (nsub + 31) >> 5;
  if (nit == 0) return;
  // register pipeline in front of the copies: entry record (two iterations before its copies are
  // issued), then -- diagonal pairs only -- the landmark index that locates y (one iteration before)
  auto fetch = [&](int i) -> int2 {
    const int sub = i * 32 + lane;
    return sub < nsub ? ent[sub / LD] : make_int2(-1, -1);
  };
  auto fetch_lm = [&](const int2& ee) -> int { return (diag && ee.x >= 0) ? l0 + k.lm[ee.x] : 0; };
  auto issue = [&](int i, const int2& ee, int l) {
    if (ee.x >= 0) {
      const int q = (i * 32 + lane) % LD;
      double2* dst = ring + (i % S) * (6 * 32) + lane;
      const double* a = k.Z + (size_t)ee.x * ZB + q * ZCOL;
#pragma unroll
      for (int j = 0; j < 3; ++j) cp_async16(dst + j * 32, a + 2 * j);
      if (diag) {
        cp_async8(dst + 3 * 32, k.y + (size_t)q * k.n_lm + l);
      } else {
        const double* bb = k.Z + (size_t)ee.y * ZB + q * ZCOL;
#pragma unroll
        for (int j = 0; j < 3; ++j) cp_async16(dst + (3 + j) * 32, bb + 2 * j);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < S; ++st) {
    const int2 ee = fetch(st);
    issue(st, ee, fetch_lm(ee));
  }
  int2 e1 = fetch(S), e2 = fetch(S + 1);
  int l1 = fetch_lm(e1);
  for (int i = 0; i < nit; ++i) {
    cp_async_wait<S - 1>();
    if (i * 32 + lane < nsub) {
      const double2* src = ring + (i % S) * (6 * 32) + lane;
      double za[6];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double2 v = src[j * 32];
        za[2 * j] = v.x;
        za[2 * j + 1] = v.y;
      }
      if (diag) {
        const double yq = src[3 * 32].x;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[r * 6 + c] += za[r] * za[c];
          acc[36 + r] += za[r] * yq;
        }
      } else {
        double zb[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double2 v = src[(3 + j) * 32];
          zb[2 * j] = v.x;
          zb[2 * j + 1] = v.y;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[r * 6 + c] += za[r] * zb[c];
      }
    }
    const int2 e_now = e1;
    const int l_now = l1;
    e1 = e2;
    l1 = fetch_lm(e1);
    e2 = fetch(i + S + 2);
    issue(i + S, e_now, l_now);
  }
  cp_async_wait<0>();
}

// Reduces 64 per-lane values across the warp with 62 shuffles instead of 64 x 5: at every step a
// lane keeps one half of its values and trades the other half with its partner. Fixed pattern =>
// bitwise deterministic. On return lane L holds the totals of elements 2L (v[0]) and 2L+1 (v[1]).
BA_DEV void warp_transpose_reduce64(double* v, int lane) {
#pragma unroll
  for (int off = 16, n = 64; off >= 1; off >>= 1, n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = upper ? v[i] : v[i + half];
      const double keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// One warp per CTA: the pairs of a window have very different list lengths (a diagonal pair lists every
// edge of its pose, a far pair a few dozen), and warps sharing a CTA would hold its registers until the
// longest one finishes (measured: 7 resident warps per SM instead of 12).
__global__ void __launch_bounds__(32) kb_schur_reduce(const __grid_constant__ LocalDev d,
                                                      const __grid_constant__ BatchDev b) {
  // pair index fastest: the warps of one window run together, so the window's Z blocks (each read by
  // every pair that contains its pose) are fetched from HBM once and then hit in L2
  __shared__ __align__(16) double2 ring[RED_RING];
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.y, li = blockIdx.x; // li: position in the compact pair list
  const WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_TRIAL) return;
  const int n_ne = b.n_ne[w];
  if (b.global && li == 0 && lane == 0) b.hs_part_w[(size_t)n_ne * 42] = s.prep_fail ? 1.0 : 0.0;
  if (li >= n_ne) return;
  const int p = b.ne_list[(size_t)w * b.Pmax + li];
  const int nf = s.nf;
  int fi, fj;
  pair_decode(p, nf, fi, fj);
  const int f0 = b.nf_begin[w];
  if (b.sys_idx[f0 + fi] < 0 || b.sys_idx[f0 + fj] < 0) return;
  const int* pb = b.pair_beg + (size_t)w * (2 * b.Pmax + 1) + 2 * p;
  const bool diag = fi == fj;
  double acc[64];
#pragma unroll
  for (int q = 0; q < 64; ++q) acc[q] = 0;
  schur_pair_entries<0>(d.k[0], w, b.pairs + pb[0], pb[1] - pb[0], diag, lane, acc, ring);
  schur_pair_entries<1>(d.k[1], w, b.pairs + pb[1], pb[2] - pb[1], diag, lane, acc, ring);
  warp_transpose_reduce64(acc, lane);
  if (2 * lane < 42) {
    double* out = b.hs_part_w + ((size_t)w * b.Pmax + li) * 42 + 2 * lane;
    out[0] = acc[0];
    out[1] = acc[1];
  }
}

// Tables of the tiled Schur path (local_tiled.cuh): landmark tiles sized by shared memory, their slices of the
// pair lists, the per-tile partial reduced systems.
struct TileDev {
  int Q;           // tile quantile in bytes of shared memory
  int Tcap;        // tiles per (window, kind) <= Tcap
  int Tp, Tl;      // grid widths: max tiles of points / lines over the windows
  int* tile_lm;    // [(w*2+kind)*(Tcap+1) + t] first landmark (batch-global index) of tile t; entry ntile = end
  int* ntile;      // [w*2+kind]
  int cost_b[2];   // shared-memory bytes per edge (Z block, landmark index, its share of the staged pair entries)
  int* tpb;        // [((w*2+kind)*(Tcap+1) + t)*Pmax + li] first entry of tile t in the kind-list of compact pair li
  int* tso;        // [((w*2+kind)*Tcap + t)*(Pmax+1) + li] offset of pair li inside the tile's entry block; [n_ne] = total
  int* tent_base;  // [(w*2+kind)*Tcap + t] position of the tile's entry block in tent
  ushort2* tent;   // tile-major copy of the pair entries, edge indices relative to the tile's first edge
  int* order;      // [w*Pmax + o] compact pair position, longest list first, | 1 << 30 for a diagonal pair
  int4* desc;      // [((w*2+kind)*Tcap + t)*2]: {la, lb, ea, eb}, {n_ent, tent_base, n_ne, valid}: everything a tile CTA
                   // needs to find its data, in one 32-byte read (kt_tiles_desc)
  double* hs_tile; // [((w*(Tp+Tl) + tt)*Pmax + li)*42], tt = t (points) or Tp + t (lines)
  double* P_bR;    // [NP][9] rotation of the pose backup (pre-update state of the current trial)
};

// K4 + pose update: one CTA per window; reduced system assembled and factorised in shared memory.
// TILED: the tiled Schur path (hs_part summed from the tiles by kt_tile_sum) also keeps the rotation of the pose backup.
template <bool TILED>
__global__ void __launch_bounds__(256) kb_solve(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                const __grid_constant__ TileDev td) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = blockIdx.x, tid = threadIdx.x;
  WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_TRIAL) return;
  const int nf = s.nf, f0 = b.nf_begin[w], p0 = d.pose_begin[w];
  const int n = 6 * s.n_sys;
  // lower triangle, row-major with an odd leading dimension (conflict-free rows and columns); the 6 x 6-blocked
  // factorisation and the two sweeps are bcr_solver.cuh's shared-memory routines
  const int ld = n + 1;
  double* Hs = reinterpret_cast<double*>(smem_raw);
  double* xs = Hs + (size_t)n * ld;
  double* dinv = xs + n;      // [n] reciprocal diagonal of the factor
  double* sc_part = dinv + n; // [nf] pose part of the LM scale
  __shared__ int s_ok;
  __shared__ int s_fail;
  const double lambda = s.lambda;
  // assemble: zero background, then the pairs of the compact list (block (si, sj), si <= sj, goes to the lower
  // triangle as its transpose)
  for (int idx = tid; idx < n * ld; idx += blockDim.x) Hs[idx] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  const int n_ne = b.n_ne[w];
  const int2* blk = b.blk + (size_t)w * b.Pmax;
  for (int idx = tid; idx < n_ne * 36; idx += blockDim.x) {
    const int li = idx / 36, rc = idx - 36 * li, r = rc / 6, c = rc - 6 * r;
    const int2 bk = blk[li];
    if (bk.x < 0) continue;
    const int si = bk.x & 0xffff, sj = bk.x >> 16;
    double v = -b.hs_part[((size_t)w * b.Pmax + li) * 42 + rc];
    if (bk.y >= 0) { // diagonal pair: only r <= c is meaningful (and only the lower triangle of Hs is read)
      if (r > c) continue;
      v += b.Hpp[(size_t)bk.y * 21 + up6(r, c)] + (r == c ? lambda : 0.0);
    }
    Hs[(size_t)(6 * sj + c) * ld + 6 * si + r] = v;
  }
  for (int idx = tid; idx < nf * 6; idx += blockDim.x) {
    const int fi = idx / 6, r = idx - 6 * fi;
    const int si = b.sys_idx[f0 + fi];
    if (si < 0) continue;
    const int li = b.diag_pos[f0 + fi];
    xs[6 * si + r] = b.bp[(size_t)(f0 + fi) * 6 + r] - b.hs_part[((size_t)w * b.Pmax + li) * 42 + 36 + r];
  }
  __syncthreads();
  if (n > 0) {
    // fails iff a pivot <= 0, like LinearSolverEigen (§9.11); xs is row n of the array: the right-hand side rides
    // through the factorisation (forward substitution included)
    bcr_cta_cholesky(Hs, ld, n, dinv, &s_fail, nullptr, 1);
    if (!s_fail && tid < 32) bcr_warp_backward(Hs, ld, dinv, n, xs);
    __syncthreads();
    if (tid == 0) s_ok = !s_fail && !s.prep_fail;
  } else if (tid == 0) {
    s_ok = !s.prep_fail;
  }
  __syncthreads();
  const bool ok = s_ok;
  if (ok) {
    for (int fi = tid; fi < nf; fi += blockDim.x) {
      const int si = b.sys_idx[f0 + fi];
      double part = 0;
      if (si >= 0) {
        const size_t gp = (size_t)(p0 + b.pose_of[f0 + fi]);
        Pose T;
        double x6[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          x6[q] = xs[6 * si + q];
          b.xp[(size_t)(f0 + fi) * 6 + q] = x6[q];
          part += x6[q] * (lambda * x6[q] + b.bp[(size_t)(f0 + fi) * 6 + q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) T.q[q] = b.P_bq[4 * gp + q] = b.P_q[4 * gp + q];
#pragma unroll
        for (int q = 0; q < 3; ++q) T.t[q] = b.P_bt[3 * gp + q] = b.P_t[3 * gp + q];
        if (TILED) {
#pragma unroll
          for (int q = 0; q < 9; ++q) td.P_bR[9 * gp + q] = b.P_R[9 * gp + q];
        }
        const Pose Tn = pose_oplus(T, x6);
        double Rn[9];
        quat_to_R(Tn.q, Rn);
#pragma unroll
        for (int q = 0; q < 4; ++q) b.P_q[4 * gp + q] = Tn.q[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) b.P_t[3 * gp + q] = Tn.t[q];
#pragma unroll
        for (int q = 0; q < 9; ++q) b.P_R[9 * gp + q] = Rn[q];
      }
      sc_part[fi] = part;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double sp = 0;
    if (ok)
      for (int fi = 0; fi < nf; ++fi) sp += sc_part[fi];
    s.scale_pose = sp;
    s.solve_ok = ok ? 1 : 0;
  }
}

// Dense-solve path, K4 split in three: assemble -> dense Cholesky / cyclic reduction (host-enqueued) -> pose update.
// grid (longest compact pair list, W), 64 threads: one CTA per pose pair block.
__global__ void __launch_bounds__(64) kb_assemble_dense(const __grid_constant__ LocalDev d,
                                                        const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y, li = blockIdx.x, tid = threadIdx.x; // li: position in the compact pair list
  const WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_TRIAL || li >= b.n_ne[w]) return;
  const int nf = s.nf;
  const int p = b.ne_list[(size_t)w * b.Pmax + li];
  int fi, fj;
  pair_decode(p, nf, fi, fj);
  const int f0 = b.nf_begin[w];
  const int si = b.sys_idx[f0 + fi], sj = b.sys_idx[f0 + fj];
  if (si < 0 || sj < 0) return;
  const int n = 6 * s.n_sys;
  double* Hs = b.dense_H + b.dense_off[w]; // (zeroed by the host before this launch: pairs off the list are zero blocks)
  const double* part = b.hs_part + ((size_t)w * b.Pmax + li) * 42;
  if (tid < 36) {
    const int r = tid / 6, c = tid % 6;
    double v = -part[tid];
    if (fi == fj) {
      const int rr = r < c ? r : c, cc = r < c ? c : r;
      v += b.Hpp[(size_t)(f0 + fi) * 21 + up6(rr, cc)] + (r == c ? s.lambda : 0.0);
    }
    Hs[(size_t)(6 * si + r) * n + 6 * sj + c] = v;
  } else if (tid < 42 && fi == fj) {
    const int r = tid - 36;
    b.dense_b[(size_t)6 * f0 + 6 * si + r] = b.bp[(size_t)(f0 + fi) * 6 + r] - part[36 + r];
  }
}

// Block-tridiagonal assembly for the cyclic-reduction solver (bcr_solver.cuh): pose pair (si, sj) of window 0 goes
// to the diagonal super-block D[si / bsp] (both triangles) or to the level-0 coupling E[si / bsp]; D and E are
// zeroed by the host before this launch. grid (longest compact pair list), 64 threads
__global__ void __launch_bounds__(64) kb_assemble_bcr(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                      const __grid_constant__ BcrDev s, int bsp) {
  const int w = 0, li = blockIdx.x, tid = threadIdx.x;
  const WinState& st = b.ws[w];
  if (st.stage != STAGE_NEED_TRIAL || li >= b.n_ne[w]) return;
  const int nf = st.nf;
  const int p = b.ne_list[(size_t)w * b.Pmax + li];
  int fi, fj;
  pair_decode(p, nf, fi, fj);
  const int f0 = b.nf_begin[w];
  const int si = b.sys_idx[f0 + fi], sj = b.sys_idx[f0 + fj];
  if (si < 0 || sj < 0) return;
  const int bs = s.bs;
  const size_t bb = (size_t)bs * bs;
  const int I = si / bsp, J = sj / bsp;
  const int r0 = 6 * (si - I * bsp), c0 = 6 * (sj - J * bsp);
  const double* part = b.hs_part + ((size_t)w * b.Pmax + li) * 42;
  if (tid < 36) {
    const int r = tid / 6, c = tid % 6;
    double v = -part[tid];
    if (fi == fj) { // symmetric diagonal block from its upper triangle
      const int rr = r < c ? r : c, cc = r < c ? c : r;
      v = -part[rr * 6 + cc] + b.Hpp[(size_t)(f0 + fi) * 21 + up6(rr, cc)] + (r == c ? st.lambda : 0.0);
    }
    if (I == J) {
      s.D[(size_t)I * bb + (size_t)(r0 + r) * bs + c0 + c] = v;
      if (si != sj) s.D[(size_t)I * bb + (size_t)(c0 + c) * bs + r0 + r] = v;
    } else if (J == I + 1) {
      s.E[s.eoff[0] + (size_t)I * bb + (size_t)(r0 + r) * bs + c0 + c] = v;
    } else {
      atomicOr(s.info, 2); // outside the band the solver was set up for (cannot happen: checked at setup)
    }
  } else if (tid < 42 && fi == fj) {
    const int r = tid - 36;
    s.x[(size_t)6 * si + r] = b.bp[(size_t)(f0 + fi) * 6 + r] - part[36 + r];
  }
}

// pose update after the dense solve (the tail of kb_solve); one CTA per window
__global__ void __launch_bounds__(256) kb_post_solve(const __grid_constant__ LocalDev d,
                                                     const __grid_constant__ BatchDev b) {
  __shared__ double red[8];
  const int w = blockIdx.x, tid = threadIdx.x;
  WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_TRIAL) return;
  const int nf = s.nf, f0 = b.nf_begin[w], p0 = d.pose_begin[w];
  const bool prep_fail = b.global ? b.hs_part[(size_t)b.n_ne[w] * 42] != 0.0 : s.prep_fail != 0; // any rank's landmarks
  const bool ok = (s.n_sys == 0 || b.dense_info[w] == 0) && !prep_fail; // potrf info > 0 <=> a pivot <= 0 (§9.11)
  const double lambda = s.lambda;
  const double* xs = b.dense_b + (size_t)6 * f0;
  double part = 0;
  if (ok) {
    for (int fi = tid; fi < nf; fi += blockDim.x) {
      const int si = b.sys_idx[f0 + fi];
      if (si < 0) continue;
      const size_t gp = (size_t)(p0 + b.pose_of[f0 + fi]);
      Pose T;
      double x6[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        x6[q] = xs[6 * si + q];
        b.xp[(size_t)(f0 + fi) * 6 + q] = x6[q];
        part += x6[q] * (lambda * x6[q] + b.bp[(size_t)(f0 + fi) * 6 + q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) T.q[q] = b.P_bq[4 * gp + q] = b.P_q[4 * gp + q];
#pragma unroll
      for (int q = 0; q < 3; ++q) T.t[q] = b.P_bt[3 * gp + q] = b.P_t[3 * gp + q];
      const Pose Tn = pose_oplus(T, x6);
      double Rn[9];
      quat_to_R(Tn.q, Rn);
#pragma unroll
      for (int q = 0; q < 4; ++q) b.P_q[4 * gp + q] = Tn.q[q];
#pragma unroll
      for (int q = 0; q < 3; ++q) b.P_t[3 * gp + q] = Tn.t[q];
#pragma unroll
      for (int q = 0; q < 9; ++q) b.P_R[9 * gp + q] = Rn[q];
    }
  }
  // fixed-order block sum of the pose part of the LM scale
  part = warp_allreduce(part);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    double sp = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) sp += red[q];
    s.scale_pose = ok ? sp : 0.0;
    s.solve_ok = ok ? 1 : 0;
  }
}

// K5/K6: back-substitution, manifold update, re-evaluation; one thread per landmark
template <int KIND>
BA_DEV void backsub_one(const LocalDev& d, const BatchDev& b, const LocalOpt& o, const KindDev& k, int w, int l,
                        double lambda, bool robust, double& chi_part, double& scale_part) {
  using T = KT<KIND>;
  constexpr int LD = T::LD;
  const int p0 = d.pose_begin[w], f0 = b.nf_begin[w];
  double v[LD];
#pragma unroll
  for (int a = 0; a < LD; ++a) v[a] = k.y[(size_t)a * k.n_lm + l];
  const int ea = k.ebeg[l], eb = k.ebeg[l + 1];
  for (int e = ea; e < eb; ++e) {
    const int fi = b.free_idx[p0 + (k.info[e] & 0xffff)];
    if (fi < 0 || b.sys_idx[f0 + fi] < 0) continue;
    // (the block is read with 16-byte loads: this kernel is bound by L1 wavefronts, one per line an instruction touches)
    const double2* Ze2 = reinterpret_cast<const double2*>(k.Z + (size_t)e * ZBlk<KIND>::N);
    const double* xv = b.xp + (size_t)(f0 + fi) * 6;
    double xr[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) xr[r] = xv[r];
#pragma unroll
    for (int a = 0; a < LD; ++a) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double2 z = Ze2[a * 3 + j];
        v[a] -= z.x * xr[2 * j] + z.y * xr[2 * j + 1];
      }
    }
  }
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
__global__ void __launch_bounds__(BT) kb_backsub(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                 const __grid_constant__ LocalOpt o) {
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
    if (i < nl && k.act[l0 + i]) backsub_one<KIND>(d, b, o, k, w, l0 + i, s.lambda, s.robust, v[0], v[1]);
  }
  cta_reduce<2>(v, red);
  if (threadIdx.x == 0) {
    double* pp = b.part + ((size_t)w * b.C + c) * 4;
    pp[0] = cta_reduce_get<2>(red, 0);
    pp[1] = cta_reduce_get<2>(red, 1);
  }
}

// Global BA: fixed-order sum of this rank's per-chunk partials (window 0) for the all-reduce.
// which = 0 after linearisation (chi, nact, max diagonal), 1 after back-substitution (chi1, scale).
__global__ void __launch_bounds__(256) kb_global_sums(const __grid_constant__ BatchDev b, int which) {
  __shared__ double red[3][8];
  const int tid = threadIdx.x;
  const WinState& s = b.ws[0];
  if (which == 0 ? s.stage != STAGE_NEED_LIN : (s.stage != STAGE_NEED_TRIAL || !s.solve_ok)) return;
  double v0 = 0, v1 = 0, mx = 0;
  for (int c = tid; c < b.C; c += 256) {
    const double* pp = b.part + (size_t)c * 4;
    v0 += pp[0];
    v1 += pp[1];
    if (which == 0) mx = fmax(mx, pp[2]);
  }
  v0 = warp_allreduce(v0);
  v1 = warp_allreduce(v1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = v0;
    red[1][tid >> 5] = v1;
    red[2][tid >> 5] = mx;
  }
  __syncthreads();
  if (tid == 0) {
    double a0 = 0, a1 = 0, am = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      a0 += red[0][q];
      a1 += red[1][q];
      am = fmax(am, red[2][q]);
    }
    double* o = b.gs_w + (which ? 4 : 0);
    o[0] = a0;
    o[1] = a1;
    if (which == 0) o[2] = am;
  }
}

// one WARP per window: the Levenberg accept / reject logic (§9.9) and the window's next stage (every lane takes
// the same decision from the same warp-reduced sums; lane 0 writes the state, the lanes share the pose restore);
// grid ceil(W / 4), 128 threads
BA_DEV void decide_window(const LocalDev& d, const BatchDev& b, int w, int lane) {
  WinState& s = b.ws[w];
  if (s.stage != STAGE_NEED_TRIAL) return;
  const bool ok = s.solve_ok;
  double chi1 = 0, scale = 0;
  if (ok && b.global) {
    chi1 = b.gs[4];
    scale = b.gs[5];
  } else if (ok) {
    double chi_l = 0, scale_l = 0; // (points and lines separately, see kb_begin_trial)
    for (int c = lane; c < b.Cp; c += 32) {
      const double* pp = b.part + ((size_t)w * b.C + c) * 4;
      chi1 += pp[0];
      scale += pp[1];
    }
    for (int c = b.Cp + lane; c < b.C; c += 32) {
      const double* pp = b.part + ((size_t)w * b.C + c) * 4;
      chi_l += pp[0];
      scale_l += pp[1];
    }
    chi1 = warp_allreduce(chi1) + warp_allreduce(chi_l);
    scale = warp_allreduce(scale) + warp_allreduce(scale_l);
  }
  scale += s.scale_pose;
  const double tempChi = ok ? chi1 : DBL_MAX; // a failed factorisation is a rejected step
  double rho = s.chi_cur - tempChi;
  rho /= (ok ? scale : 0.0) + 1e-3;
  double lambda = s.lambda, ni = s.ni, chi_cur = s.chi_cur;
  bool stop_lambda = false, accepted = false;
  if (rho > 0 && isfinite(tempChi)) {
    const double c = 2 * rho - 1;
    double alpha = 1. - c * c * c;
    alpha = fmin(alpha, 2. / 3.);
    lambda *= fmax(1. / 3., alpha);
    ni = 2;
    chi_cur = tempChi;
    accepted = true;
  } else {
    lambda *= ni;
    ni *= 2;
    if (!isfinite(lambda)) stop_lambda = true;
  }
  const int q1 = stop_lambda ? s.qmax : s.qmax + 1;
  const int nf = s.nf, pass = s.pass, it = s.it, iters = s.iters;
  const double nact = s.nact;
  __syncwarp(); // every lane has read the state before lane 0 rewrites it
  if (!accepted && ok) { // pop(): poses here, landmarks in kb_restore
    const int p0 = d.pose_begin[w], f0 = b.nf_begin[w];
    for (int fi = lane; fi < nf; fi += 32) {
      if (b.sys_idx[f0 + fi] < 0) continue;
      const size_t gp = (size_t)(p0 + b.pose_of[f0 + fi]);
      double q[4], R[9];
      for (int i = 0; i < 4; ++i) q[i] = b.P_q[4 * gp + i] = b.P_bq[4 * gp + i];
      for (int i = 0; i < 3; ++i) b.P_t[3 * gp + i] = b.P_bt[3 * gp + i];
      quat_to_R(q, R);
      for (int i = 0; i < 9; ++i) b.P_R[9 * gp + i] = R[i];
    }
  }
  if (lane != 0) return;
  s.lambda = lambda;
  s.ni = ni;
  s.chi_cur = chi_cur;
  s.qmax = q1;
  s.st.trials[pass]++;
  if (ok) s.st.edges_evaluated += (long long)nact;
  if (!accepted && ok) s.restore = 1;
  if (!stop_lambda && rho < 0 && q1 < 10) {
    s.stage = STAGE_NEED_TRIAL; // retry with the larger lambda
    return;
  }
  s.st.iters[pass]++;
  s.it = it + 1;
  const bool terminate = (q1 == 10 || rho == 0 || !isfinite(lambda));
  s.stage = (terminate || it + 1 >= iters) ? STAGE_DONE : STAGE_NEED_LIN;
}

// COND (the whole-schedule graph): the last warp of the grid to finish also evaluates the loop condition of the WHILE
// node (any window still iterating?) and counts the super-step, which saves the separate one-CTA kernel on the
// critical path of every super-step. ticket: d.err[3], zeroed with the error flags by the setup.
template <bool COND>
__global__ void __launch_bounds__(128) kb_decide(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                                 cudaGraphConditionalHandle h) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w < d.n_windows) decide_window(d, b, w, lane);
  if (!COND) return;
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence();
    const int n_warps = (int)(gridDim.x * (blockDim.x >> 5));
    last = atomicAdd(d.err + 3, 1) == n_warps - 1 ? 1 : 0;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  int any = 0;
  for (int i = lane; i < d.n_windows; i += 32) any |= *reinterpret_cast<volatile const int*>(&b.ws[i].stage) != STAGE_DONE ? 1 : 0;
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) {
    d.err[3] = 0;
    cudaGraphSetConditional(h, any ? 1u : 0u);
    ++*b.n_active;
  }
}

// grid (Cp + Cl, windows): chunk < Cp restores points, else lines
__global__ void __launch_bounds__(BT) kb_restore(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b) {
  const int w = blockIdx.y, c = blockIdx.x;
  if (!b.ws[w].restore) return;
  if (c < b.Cp) {
    const KindDev& k = d.k[0];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = c * BT + threadIdx.x;
    if (i < nl && k.act[l0 + i]) {
#pragma unroll
      for (int q = 0; q < 3; ++q) k.x[(size_t)q * k.n_lm + l0 + i] = k.xb[(size_t)q * k.n_lm + l0 + i];
    }
  } else {
    const KindDev& k = d.k[1];
    const int l0 = k.lm_begin[w], nl = k.lm_begin[w + 1] - l0;
    const int i = (c - b.Cp) * BT + threadIdx.x;
    if (i < nl && k.act[l0 + i]) {
#pragma unroll
      for (int q = 0; q < 6; ++q) k.x[(size_t)q * k.n_lm + l0 + i] = k.xb[(size_t)q * k.n_lm + l0 + i];
    }
  }
}

__global__ void __launch_bounds__(256) kb_count_active(const __grid_constant__ LocalDev d,
                                                       const __grid_constant__ BatchDev b) {
  int n = 0;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < d.n_windows; w += gridDim.x * blockDim.x)
    n += b.ws[w].stage != STAGE_DONE;
  n = __reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(b.n_active, n);
}

// flagging / write-back, edge-parallel
template <int KIND, bool FINAL>
BA_DEV void flag_edges(const LocalDev& d, const BatchDev& b, const LocalOpt& o, const KindDev& k, int w, int chunk,
                       int nchunks) {
  const int l0 = k.lm_begin[w];
  const int p0 = d.pose_begin[w];
  const int e0 = edge_base(k, w), e1 = edge_base(k, w + 1);
  for (int e = e0 + chunk * blockDim.x + threadIdx.x; e < e1; e += nchunks * blockDim.x) {
    const int info = k.info[e];
    const bool stereo = (info >> 30) & 1;
    const double thr = o.thr[2 * KIND + (stereo ? 1 : 0)];
    bool depth_ok = true;
    if (KIND == 0) {
      const int p = info & 0xffff;
      BA_CHECK(p < d.pose_begin[w + 1] - d.pose_begin[w]);
    BA_CHECK(p < d.pose_begin[w + 1] - d.pose_begin[w]);
      double X[3], Xc[3];
      load_lm<0>(k, l0 + k.lm[e], X);
      transform_point(b.P_R + 9 * (size_t)(p0 + p), b.P_t + 3 * (size_t)(p0 + p), X, Xc);
      depth_ok = Xc[2] > 0.0;
    }
    if (FINAL) {
      const int key = k.src[e];
      k.out_inl[key >> 30][key & 0x3fffffff] = (k.chi2[e] <= thr && depth_ok) ? 1 : 0; // :213-231
    } else {
      k.lvl[e] = (k.chi2[e] > thr || !depth_ok) ? 1 : 0; // :176-206
    }
  }
}

template <bool FINAL>
__global__ void __launch_bounds__(256) kb_flag(const __grid_constant__ LocalDev d, const __grid_constant__ BatchDev b,
                                               const __grid_constant__ LocalOpt o) {
  const int w = blockIdx.x;
  flag_edges<0, FINAL>(d, b, o, d.k[0], w, blockIdx.y, gridDim.y);
  flag_edges<1, FINAL>(d, b, o, d.k[1], w, blockIdx.y, gridDim.y);
}

__global__ void __launch_bounds__(256) kb_writeback(const __grid_constant__ LocalDev d,
                                                    const __grid_constant__ BatchDev b) {
  // grid (windows, chunks): the chunks share a window's landmarks (one huge window is as parallel as many small ones)
  const int w = blockIdx.x, tid = threadIdx.x;
  write_landmarks<0>(d.k[0], w, blockIdx.y, gridDim.y);
  write_landmarks<1>(d.k[1], w, blockIdx.y, gridDim.y);
  if (blockIdx.y) return;
  const int p0 = d.pose_begin[w], np = d.pose_begin[w + 1] - p0;
  for (int p = tid; p < np; p += blockDim.x) {
    Pose T;
#pragma unroll
    for (int i = 0; i < 4; ++i) T.q[i] = b.P_q[4 * (size_t)(p0 + p) + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) T.t[i] = b.P_t[3 * (size_t)(p0 + p) + i];
    const Pose Twc = pose_inverse(T);
#pragma unroll
    for (int i = 0; i < 3; ++i) d.pose_out[(size_t)i * d.n_poses + p0 + p] = Twc.t[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) d.pose_out[(size_t)(3 + i) * d.n_poses + p0 + p] = Twc.q[i];
  }
  if (tid == 0 && d.stats) {
    WinState& s = b.ws[w];
    s.st.final_chi2 = s.chi_cur;
    s.st.final_lambda = s.lambda;
    reinterpret_cast<DevStats*>(d.stats)[w] = s.st;
  }
}

} // namespace ba
