// frame_kernel.cuh — K7: batched pose-only optimisation, the whole FrameOptimization
// (/root/reference/src/g2o_optimization/g2o_optimization.cc:256-397) of one frame per WARP.
//
// One launch runs, for every frame of the batch, the reference's 4 rounds x LM(10) with Huber
// reweighting, chi2 re-classification between rounds (the (float)chi2 comparison of :351-352,
// the pose reset of :340, the kernel drop after round index 2, :364) and the g2o Levenberg
// accept/reject logic entirely on the device. Edges are read from
// structure-of-arrays planes (coalesced); the 6x6 system is reduced with a fixed-order
// shuffle butterfly (bitwise deterministic), solved redundantly by every lane with an in-register
// Cholesky (no broadcast, no block barrier: a warp never waits for another frame), and the only
// HBM writes are the final flags, pose and stats.
#pragma once

#include <float.h>
#include <stdint.h>

#include "ba_math.cuh"

namespace ba {

struct FrameDev {
  int n_frames;
  int n_cameras;
  const double* cameras;      // [n_cameras][5]
  const double* pose_twc;     // [7][F]
  const int* mono_begin;      // [F+1]
  const int* stereo_begin;    // [F+1]
  int n_mono, n_stereo;       // plane strides
  const double* mono_meas;    // [2][n_mono]
  const double* mono_xw;      // [3][n_mono]
  const int* mono_cam;        // may be null
  const uint8_t* mono_inl_in; // may be null
  const double* stereo_meas;  // [3][n_stereo]
  const double* stereo_xw;    // [3][n_stereo]
  const int* stereo_cam;
  const uint8_t* stereo_inl_in;
  // line extension (constraints on fixed lines; absent in the reference): n_mline + n_sline == 0 when unused
  const int* mline_begin;     // [F+1]
  const int* sline_begin;     // [F+1]
  int n_mline, n_sline;
  const double* mline_lw;     // [6][n_mline]
  const double* mline_meas;   // [4][n_mline]
  const int* mline_cam;       // may be null
  const uint8_t* mline_inl_in;
  const double* sline_lw;     // [6][n_sline]
  const double* sline_meas;   // [8][n_sline]
  const int* sline_cam;
  const uint8_t* sline_inl_in;
  uint8_t* mline_inl;
  uint8_t* sline_inl;
  uint8_t* mline_lvl;
  uint8_t* sline_lvl;
  // outputs / work
  double* out_pose_twc;       // [7][F]
  uint8_t* mono_inl;          // [n_mono]   current ->inlier flag (level = !inlier after round 0)
  uint8_t* stereo_inl;        // [n_stereo]
  uint8_t* mono_lvl;          // [n_mono]   edge level (0 active / 1 excluded)
  uint8_t* stereo_lvl;        // [n_stereo]
  int* num_inliers;           // [F]
  void* stats;                // RsplBaStats[F]
};

struct FrameOpt {
  double thr_mono, thr_stereo;
  double delta_mono, delta_stereo; // (float)sqrt(thr) widened
  double thr_mline, thr_sline, delta_mline, delta_sline; // line extension
  int rounds, iters;
  Cam cam0;                        // camera 0, read straight from the constant bank when SINGLE_CAM
  double b0;                       // cam0.bf / cam0.fx (stereo line edges), divided once on the host
  int frame0, frame1;              // frame range of this launch (chunks of a pipelined batch)
};

struct DevStats { // layout == RsplBaStats
  int iters[4];
  int trials[4];
  long long edges_linearized;
  long long edges_evaluated;
  double final_chi2;
  double final_lambda;
};

#ifndef FRAME_MIN_BLOCKS
#define FRAME_MIN_BLOCKS 4 // x 4 warps per SM = 128 registers: best of {3, 4, 5} measured on B200 (profiles/README.md)
#endif
#ifndef FRAME_LINES_MIN_BLOCKS
#define FRAME_LINES_MIN_BLOCKS 4 // instantiation with the line extension: 3 / 4 / 5 CTAs per SM measured 4.03 / 3.60 / 5.34 ms (C2 with 60 lines)
#endif
#ifndef FRAME_WARPS_PER_CTA
#define FRAME_WARPS_PER_CTA 1 // frames per CTA in throughput mode (one warp each); 1 / 2 / 4: c2p 1.76 / 1.81 / 1.80 ms (finer tail)
#endif
constexpr int FRAME_WARPS = FRAME_WARPS_PER_CTA;
constexpr int FRAME_THREADS = 32 * FRAME_WARPS;
constexpr int NACC = 28; // 21 (H upper) + 6 (b) + 1 (robust chi2)
constexpr int FRAME_CTA_WARPS = 8;  // latency variant: one CTA of 8 warps per frame

// Measured and dropped in round 2 (gpurun_out logs summarised in profiles/README.md): software prefetch of the next
// record into registers (no gain); a cp.async ring through shared memory (c2p 1.92 -> 2.36 ms: the 8-byte LDGSTS and
// their address arithmetic cost more issue slots than the hidden latency was worth); evaluating 2 / 4 / 8 edges of a
// lane side by side in one branch-free block (slower by 2 - 40 %: in throughput mode the kernel is bound by issue slots
// and the FP64 pipe together with latency, and every variant that adds instructions loses). What helped: levels in a
// register bit mask, one warp per CTA (finer tail), reciprocal / rsqrt without the library's range test and
// out-of-line slow path (halves the single-frame latency: the serial 6x6 solve is a chain of six rsqrt).
constexpr int LINE_LVL_CAP = 256;     // line levels of a frame kept in shared memory (beyond: global)

// per-warp (= per-frame) state that is touched once per trial: kept in shared memory so the
// edge loops keep their registers
struct WarpState {
  PoseRt T0, Tbackup, Te; // initial pose; LM backup; pose of the last error evaluation (stale errors, §9.12)
  double H[21], b[6], x[6];
  DevStats st;
};

BA_DEV void load_cam(const double* cams, int idx, Cam& c) {
  const double* p = cams + 5 * idx;
  c.fx = p[0];
  c.fy = p[1];
  c.cx = p[2];
  c.cy = p[3];
  c.bf = p[4];
}

// index of (i,j), i<=j, in the packed upper triangle of a 6x6
__host__ __device__ constexpr int up6(int i, int j) { return i * 6 - i * (i - 1) / 2 + (j - i); }

// H += w J^T J ; b -= w J^T r (information = I, §9.4, §9.7), using the structure of the pose Jacobian of a
// point edge (§9.3): row0 = [a0 a1 a2 a3 0 a5], row1 = [b0 b1 b2 0 b4 b5], row2 = row0 + [c0 c1 0 0 0 c5]
// — the zero entries are skipped explicitly (the compiler may not drop 0 * x in IEEE arithmetic).
template <int ROWS>
BA_DEV void accumulate_pose_only(const double* J, const double* r, double w, double* acc) {
  const double* a = J;
  const double* b = J + 6;
  const double wr0 = w * r[0], wr1 = w * r[1];
  // gradient
  acc[21 + 0] -= a[0] * wr0 + b[0] * wr1;
  acc[21 + 1] -= a[1] * wr0 + b[1] * wr1;
  acc[21 + 2] -= a[2] * wr0 + b[2] * wr1;
  acc[21 + 3] -= a[3] * wr0;
  acc[21 + 4] -= b[4] * wr1;
  acc[21 + 5] -= a[5] * wr0 + b[5] * wr1;
  // row 0 and row 1 outer products (entries with column 4 of row 0 / column 3 of row 1 vanish)
  const double wa[6] = {w * a[0], w * a[1], w * a[2], w * a[3], 0.0, w * a[5]};
  const double wb[6] = {w * b[0], w * b[1], w * b[2], 0.0, w * b[4], w * b[5]};
  acc[up6(0, 0)] += wa[0] * a[0] + wb[0] * b[0];
  acc[up6(0, 1)] += wa[0] * a[1] + wb[0] * b[1];
  acc[up6(0, 2)] += wa[0] * a[2] + wb[0] * b[2];
  acc[up6(0, 3)] += wa[0] * a[3];
  acc[up6(0, 4)] += wb[0] * b[4];
  acc[up6(0, 5)] += wa[0] * a[5] + wb[0] * b[5];
  acc[up6(1, 1)] += wa[1] * a[1] + wb[1] * b[1];
  acc[up6(1, 2)] += wa[1] * a[2] + wb[1] * b[2];
  acc[up6(1, 3)] += wa[1] * a[3];
  acc[up6(1, 4)] += wb[1] * b[4];
  acc[up6(1, 5)] += wa[1] * a[5] + wb[1] * b[5];
  acc[up6(2, 2)] += wa[2] * a[2] + wb[2] * b[2];
  acc[up6(2, 3)] += wa[2] * a[3];
  acc[up6(2, 4)] += wb[2] * b[4];
  acc[up6(2, 5)] += wa[2] * a[5] + wb[2] * b[5];
  acc[up6(3, 3)] += wa[3] * a[3];
  acc[up6(3, 5)] += wa[3] * a[5];
  acc[up6(4, 4)] += wb[4] * b[4];
  acc[up6(4, 5)] += wb[4] * b[5];
  acc[up6(5, 5)] += wa[5] * a[5] + wb[5] * b[5];
  if (ROWS == 3) {
    const double* c = J + 12; // c[4] == 0
    const double wr2 = w * r[2];
    acc[21 + 0] -= c[0] * wr2;
    acc[21 + 1] -= c[1] * wr2;
    acc[21 + 2] -= c[2] * wr2;
    acc[21 + 3] -= c[3] * wr2;
    acc[21 + 5] -= c[5] * wr2;
    const double wc[6] = {w * c[0], w * c[1], w * c[2], w * c[3], 0.0, w * c[5]};
    acc[up6(0, 0)] += wc[0] * c[0];
    acc[up6(0, 1)] += wc[0] * c[1];
    acc[up6(0, 2)] += wc[0] * c[2];
    acc[up6(0, 3)] += wc[0] * c[3];
    acc[up6(0, 5)] += wc[0] * c[5];
    acc[up6(1, 1)] += wc[1] * c[1];
    acc[up6(1, 2)] += wc[1] * c[2];
    acc[up6(1, 3)] += wc[1] * c[3];
    acc[up6(1, 5)] += wc[1] * c[5];
    acc[up6(2, 2)] += wc[2] * c[2];
    acc[up6(2, 3)] += wc[2] * c[3];
    acc[up6(2, 5)] += wc[2] * c[5];
    acc[up6(3, 3)] += wc[3] * c[3];
    acc[up6(3, 5)] += wc[3] * c[5];
    acc[up6(5, 5)] += wc[5] * c[5];
  }
}

// H += wo J^T J ; b -= wo J^T r for a dense ROWS x 6 Jacobian (line edges)
template <int ROWS>
BA_DEV void accumulate_pose_dense(const double* J, const double* r, double wo, double* acc) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double g = 0;
#pragma unroll
    for (int q = 0; q < ROWS; ++q) g += J[q * 6 + i] * r[q];
    acc[21 + i] -= wo * g;
#pragma unroll
    for (int j = i; j < 6; ++j) {
      double h = 0;
#pragma unroll
      for (int q = 0; q < ROWS; ++q) h += J[q * 6 + i] * J[q * 6 + j];
      acc[up6(i, j)] += wo * h;
    }
  }
}

// line edges of a frame (extension): information 0.1 I (g2o_optimization.cc:133,154), line vertex fixed
// Image-line quantities of one camera from the camera-frame moment (w0, w1, w2), with explicit fma so that the row
// mapping (linearising pass) and the edge mapping (residual-only pass) produce the same bits: the Levenberg loop
// compares their chi2 sums (rho == 0 terminates).
BA_DEV void line_to_camera_fma(const double* R, const double* t, const double* L, LineCam& lc) {
  double Rw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Rw[i] = fma(R[3 * i + 2], L[2], fma(R[3 * i + 1], L[1], R[3 * i] * L[0]));
    lc.dc[i] = fma(R[3 * i + 2], L[5], fma(R[3 * i + 1], L[4], R[3 * i] * L[3]));
  }
  lc.wc[0] = Rw[0] + fma(t[1], lc.dc[2], -(t[2] * lc.dc[1]));
  lc.wc[1] = Rw[1] + fma(t[2], lc.dc[0], -(t[0] * lc.dc[2]));
  lc.wc[2] = Rw[2] + fma(t[0], lc.dc[1], -(t[1] * lc.dc[0]));
}
struct LineImg {
  double l0, l1, l2, inv;
};
BA_DEV LineImg line_image(const Cam& cam, double w0, double w1, double w2) {
  const double kv0 = -cam.fy * cam.cx, kv1 = -cam.fx * cam.cy, kv2 = cam.fx * cam.fy;
  LineImg im;
  im.l0 = cam.fy * w0;
  im.l1 = cam.fx * w1;
  im.l2 = fma(kv2, w2, fma(kv1, w1, kv0 * w0));
  im.inv = rsqrt_nr(fma(im.l1, im.l1, im.l0 * im.l0));
  return im;
}
BA_DEV double line_row(const LineImg& im, double mx, double my) { return fma(mx, im.l0, fma(my, im.l1, im.l2)) * im.inv; }
// right camera: T with t.x -= b (edge_project_stereo_line.cc:34-35)
BA_DEV void line_right(const LineCam& lc, double b, double& w1, double& w2) {
  w1 = fma(b, lc.dc[2], lc.wc[1]);
  w2 = fma(-b, lc.dc[1], lc.wc[2]);
}

// The levels of the first LINE_LVL_CAP line edges of the frame (mono first, then stereo) live in shared memory (llvl).
BA_DEV bool line_excluded(const FrameDev& d, const uint8_t* llvl, int ei, bool st, int e) {
  return ei < LINE_LVL_CAP ? llvl[ei] != 0 : (st ? d.sline_lvl[e] : d.mline_lvl[e]) != 0;
}

// Linearising pass over the line edges, one lane per residual ROW: a frame has few line edges (60 in config C2) of very different cost (2 or 4
// rows), so a lane per edge leaves most of the warp idle (one pass over 20 mono + two over 40 stereo edges). Here
// the item space is 4 rows x (mono + stereo edges) -- rows 2, 3 of a mono edge are empty -- so 32 lanes take 8 whole
// edges per step; the chi2 of an edge (needed by its Huber weight) is a 2-step shuffle sum over its 4 lanes. The
// camera-frame line is recomputed by each of the 4 lanes (27 FMAs) instead of shared.
template <bool SINGLE_CAM, int STRIDE>
BA_DEV void line_pass_rows(const FrameDev& d, const FrameOpt& o, int ml0, int ml1, int sl0, int sl1, const uint8_t* llvl,
                           const double* R, const double* t, bool robust, int lane, double* acc) {
  const int nm = ml1 - ml0, n_items = 4 * (nm + (sl1 - sl0));
  for (int base = 0; base < n_items; base += STRIDE) { // uniform trip count: the shuffles below need the whole warp
    const int item = base + lane;
    const int ei = item >> 2, row = item & 3;
    const bool st = ei >= nm;
    const int e = st ? sl0 + (ei - nm) : ml0 + ei;
    const bool act = item < n_items && (st || row < 2) && !line_excluded(d, llvl, ei, st, e);
    double L[6] = {0, 0, 0, 0, 0, 0}, mx = 0, my = 0;
    if (act) {
      const size_t stride = st ? d.n_sline : d.n_mline;
      const double* lw = (st ? d.sline_lw : d.mline_lw) + e;
      const int mq = 4 * (row >> 1) + 2 * (row & 1); // endpoint (x, y) of this row: left 1, left 2, right 1, right 2
      const double* ms = (st ? d.sline_meas : d.mline_meas) + mq * stride + e;
#pragma unroll
      for (int q = 0; q < 6; ++q) L[q] = lw[q * stride];
      mx = ms[0];
      my = ms[stride];
    }
    double r = 0.0, J[6] = {0, 0, 0, 0, 0, 0};
    if (act) {
      Cam camv;
      if (!SINGLE_CAM) load_cam(d.cameras, st ? d.sline_cam[e] : d.mline_cam[e], camv);
      const Cam& cam = SINGLE_CAM ? o.cam0 : camv;
      LineCam lc;
      line_to_camera_fma(R, t, L, lc);
      const double b = SINGLE_CAM ? o.b0 : cam.bf / cam.fx;
      const bool right = row >= 2;
      double w1 = lc.wc[1], w2 = lc.wc[2];
      if (right) line_right(lc, b, w1, w2);
      const LineImg im = line_image(cam, lc.wc[0], w1, w2);
      const double kv0 = -cam.fy * cam.cx, kv1 = -cam.fx * cam.cy, kv2 = cam.fx * cam.fy;
      const double l0 = im.l0, l1 = im.l1, inv = im.inv;
      r = line_row(im, mx, my);
      {
        const double n0 = l0 * inv, n1 = l1 * inv;
        const double a0 = (mx - r * n0) * inv, a1 = (my - r * n1) * inv, a2 = inv;
        const double g[3] = {a0 * cam.fy + a2 * kv0, a1 * cam.fx + a2 * kv1, a2 * kv2};
        double a[3], c[3];
        cross3(lc.wc, g, a);
        cross3(lc.dc, g, c);
        if (right) {
          const double h[3] = {0.0, -b * g[2], b * g[1]};
          double e2[3];
          cross3(lc.dc, h, e2);
          a[0] += e2[0];
          a[1] += e2[1];
          a[2] += e2[2];
        }
        J[0] = a[0];
        J[1] = a[1];
        J[2] = a[2];
        J[3] = c[0];
        J[4] = c[1];
        J[5] = c[2];
      }
    }
    double c2 = __dmul_rn(r, r);
    c2 = __dadd_rn(c2, __shfl_xor_sync(0xffffffffu, c2, 1));
    c2 = __dadd_rn(c2, __shfl_xor_sync(0xffffffffu, c2, 2));
    c2 *= 0.1;
    double w = 1.0;
    const double rho0 = robust ? huber(c2, st ? o.delta_sline : o.delta_mline, w) : c2;
    if (act) {
      if (row == 0) acc[NACC - 1] += rho0;
      {
        const double wo = 0.1 * w;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          acc[21 + i] -= wo * J[i] * r;
#pragma unroll
          for (int j = i; j < 6; ++j) acc[up6(i, j)] += wo * J[i] * J[j];
        }
      }
    }
  }
}

// Residual-only pass over the line edges, one lane per EDGE: the camera-frame line and the two normalisations are
// computed once per edge instead of once per row (the row mapping above pays them four times, which only the
// linearising pass amortises over its 27 accumulations per row). Mono edges first, then stereo.
template <bool SINGLE_CAM, int STRIDE>
BA_DEV void line_pass_edges(const FrameDev& d, const FrameOpt& o, int ml0, int ml1, int sl0, int sl1, const uint8_t* llvl,
                            const double* R, const double* t, bool robust, int lane, double* acc) {
  const int nm = ml1 - ml0, ne = nm + (sl1 - sl0);
  for (int ei = lane; ei < ne; ei += STRIDE) {
    const bool st = ei >= nm;
    const int e = st ? sl0 + (ei - nm) : ml0 + ei;
    if (line_excluded(d, llvl, ei, st, e)) continue;
    Cam camv;
    if (!SINGLE_CAM) load_cam(d.cameras, st ? d.sline_cam[e] : d.mline_cam[e], camv);
    const Cam& cam = SINGLE_CAM ? o.cam0 : camv;
    const size_t stride = st ? d.n_sline : d.n_mline;
    const double* lw = (st ? d.sline_lw : d.mline_lw) + e;
    const double* ms = (st ? d.sline_meas : d.mline_meas) + e;
    double L[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) L[q] = lw[q * stride];
    LineCam lc;
    line_to_camera_fma(R, t, L, lc);
    double c2;
    {
      const LineImg im = line_image(cam, lc.wc[0], lc.wc[1], lc.wc[2]);
      const double r0 = line_row(im, ms[0], ms[stride]), r1 = line_row(im, ms[2 * stride], ms[3 * stride]);
      c2 = __dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1));
    }
    if (st) {
      double w1, w2;
      line_right(lc, SINGLE_CAM ? o.b0 : cam.bf / cam.fx, w1, w2);
      const LineImg im = line_image(cam, lc.wc[0], w1, w2);
      const double r2 = line_row(im, ms[4 * stride], ms[5 * stride]), r3 = line_row(im, ms[6 * stride], ms[7 * stride]);
      // same association as the row mapping: (r0^2 + r1^2) + (r2^2 + r3^2)
      c2 = __dadd_rn(c2, __dadd_rn(__dmul_rn(r2, r2), __dmul_rn(r3, r3)));
    }
    c2 *= 0.1;
    double w = 1.0;
    acc[NACC - 1] += robust ? huber(c2, st ? o.delta_sline : o.delta_mline, w) : c2;
  }
}

// fixed-order butterfly: every lane ends with the same bitwise-deterministic sum
BA_DEV double warp_allreduce(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 6x6 Cholesky solve of (H + lambda I) x = b, H packed upper. Returns false iff a pivot <= 0
// (LinearSolverEigen / SimplicialLLT failure rule, §9.11); x is left untouched then.
BA_DEV bool solve6(const double* Hp, const double* b, double lambda, double* x) {
  double U[6][6], inv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) U[i][j] = Hp[up6(i, j)] + (i == j ? lambda : 0.0);
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double d = U[k][k];
#pragma unroll
    for (int p = 0; p < k; ++p) d -= U[p][k] * U[p][k];
    if (d <= 0.0) ok = false; // NaN falls through like Eigen's test
    const double r = rsqrt_nr(d);
    inv[k] = r;
    U[k][k] = d * r;
#pragma unroll
    for (int j = k + 1; j < 6; ++j) {
      double sacc = U[k][j];
#pragma unroll
      for (int p = 0; p < k; ++p) sacc -= U[p][k] * U[p][j];
      U[k][j] = sacc * r;
    }
  }
  if (!ok) return false;
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double sacc = b[i];
#pragma unroll
    for (int p = 0; p < i; ++p) sacc -= U[p][i] * y[p];
    y[i] = sacc * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double sacc = y[i];
#pragma unroll
    for (int p = i + 1; p < 6; ++p) sacc -= U[i][p] * x[p];
    x[i] = sacc * inv[i];
  }
  return true;
}

// level of the k-th edge of this thread in a class: the first 32 live in a register bit mask, the rest in HBM
BA_DEV bool edge_excluded(uint32_t mask, const uint8_t* lvl, int k, int e) {
  return k < 32 ? ((mask >> k) & 1u) != 0 : lvl[e] != 0;
}

// One pass of a warp over the active point edges of one class (mono | stereo) of its frame at the pose (R,t).
//  LINEARIZE: accumulate H, b and the robust chi2 (= computeActiveErrors + activeRobustChi2 + buildSystem)
//  else      : robust chi2 only (= computeActiveErrors + activeRobustChi2)
// The thread's k-th edge is e0 + lane + k * STRIDE.
template <bool LINEARIZE, bool STEREO, bool SINGLE_CAM, int STRIDE>
BA_DEV void point_pass(const FrameDev& d, const FrameOpt& o, int e0, int e1, uint32_t lvlmask, const double* R,
                       const double* t, bool robust, int lane, double* acc) {
  const double* xw = STEREO ? d.stereo_xw : d.mono_xw;
  const double* ms = STEREO ? d.stereo_meas : d.mono_meas;
  const uint8_t* lvl = STEREO ? d.stereo_lvl : d.mono_lvl;
  const size_t n = STEREO ? d.n_stereo : d.n_mono;
  const double delta = STEREO ? o.delta_stereo : o.delta_mono;
  for (int e = e0 + lane, k = 0; e < e1; e += STRIDE, ++k) {
    if (edge_excluded(lvlmask, lvl, k, e)) continue;
    Cam camv;
    if (!SINGLE_CAM) load_cam(d.cameras, (STEREO ? d.stereo_cam : d.mono_cam)[e], camv);
    const Cam& cam = SINGLE_CAM ? o.cam0 : camv;
    const double X[3] = {xw[e], xw[n + e], xw[2 * n + e]};
    const double m[3] = {ms[e], ms[n + e], STEREO ? ms[2 * n + e] : 0.0};
    double Xc[3], r[3];
    transform_point(R, t, X, Xc);
    const double invz = rcp_nr(Xc[2]);
    point_residual_iz<STEREO>(cam, cam.bf, Xc, invz, m, r);
    const double chi2 = STEREO ? r[0] * r[0] + r[1] * r[1] + r[2] * r[2] : r[0] * r[0] + r[1] * r[1];
    double w = 1.0;
    const double rho0 = robust ? huber(chi2, delta, w) : chi2;
    acc[NACC - 1] += rho0;
    if (LINEARIZE) {
      double J[STEREO ? 18 : 12];
      point_jac_pose_iz<STEREO>(cam, Xc, invz, J);
      accumulate_pose_only<STEREO ? 3 : 2>(J, r, w, acc);
    }
  }
}

template <bool LINEARIZE, bool SINGLE_CAM, int STRIDE>
BA_DEV void edge_pass(const FrameDev& d, const FrameOpt& o, int m0, int m1, int s0, int s1, uint32_t mlvl, uint32_t slvl,
                      const double* R, const double* t, bool robust, int lane, double* acc) {
  point_pass<LINEARIZE, false, SINGLE_CAM, STRIDE>(d, o, m0, m1, mlvl, R, t, robust, lane, acc);
  point_pass<LINEARIZE, true, SINGLE_CAM, STRIDE>(d, o, s0, s1, slvl, R, t, robust, lane, acc);
}

BA_DEV void pose_to_Rt(const Pose& T, double* R, double* t) {
  quat_to_R(T.q, R);
  t[0] = T.t[0];
  t[1] = T.t[1];
  t[2] = T.t[2];
}

// One warp per frame: every lane carries the same pose / LM scalars (computed redundantly, so no
// broadcast and no block barrier is ever needed); lanes stride over the frame's edges.
// HAS_LINES: the batch carries the line extension (the point-only instantiation is the reference's path and
// keeps its register budget).
// WPF = warps per frame. 1 (default, throughput): four frames per CTA, one warp each, no block barrier. FRAME_CTA_WARPS
// (latency: the reference calls FrameOptimization with ONE frame, map_builder.cc:583-584): the whole CTA works on one
// frame, its threads stride the edges, and the sums go through a fixed-order cross-warp reduction in shared memory
// (deterministic, but a different association than the one-warp variant: the two modes agree to rounding, not bitwise).
template <int WPF>
BA_DEV void frame_sync() {
  if (WPF == 1) __syncwarp();
  else __syncthreads();
}
// every thread of the frame ends with the same totals; red: [WPF][N] shared scratch
// The N <= 32 sums of a warp at once: a transposing butterfly (lane L ends with the complete sum of element L after
// 16 + 8 + 4 + 2 + 1 exchanges) followed by N broadcasts, instead of five exchanges per element: 59 instead of 140
// 64-bit shuffles for the 28 sums of a linearisation. One owner and one fixed tree per sum: deterministic, and every
// lane receives the same bits.
template <int N>
BA_DEV void warp_allreduce_many(double* v) {
  const int lane = threadIdx.x & 31;
  double t[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) t[k] = k < N ? v[k] : 0.0;
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = upper ? t[i] : t[i + half];
      const double keep = upper ? t[i + half] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = __shfl_sync(0xffffffffu, t[0], k);
}

template <int WPF, int N>
BA_DEV void frame_allreduce(double* v, double* red) {
  if (N >= 8) {
    warp_allreduce_many<N>(v);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_allreduce(v[k]);
  }
  if (WPF > 1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) red[warp * N + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double sum = red[k];
#pragma unroll
      for (int w2 = 1; w2 < WPF; ++w2) sum += red[w2 * N + k];
      v[k] = sum;
    }
  }
}

template <bool SINGLE_CAM, bool HAS_LINES, int WPF>
__global__ void __launch_bounds__(32 * (WPF == 1 ? FRAME_WARPS : WPF),
                                  WPF == 1 ? (HAS_LINES ? FRAME_LINES_MIN_BLOCKS : FRAME_MIN_BLOCKS) * 4 / FRAME_WARPS : 2)
    frame_opt_kernel(const __grid_constant__ FrameDev d, const __grid_constant__ FrameOpt o) {
  constexpr int STRIDE = 32 * WPF;
  __shared__ uint8_t line_lvl[HAS_LINES ? (WPF == 1 ? FRAME_WARPS : 1) * LINE_LVL_CAP : 1];
  __shared__ WarpState wstate[WPF == 1 ? FRAME_WARPS : 1];
  __shared__ double red[WPF == 1 ? 1 : WPF * NACC];
  __shared__ int red_i[WPF == 1 ? 1 : WPF];
  const int warp = threadIdx.x >> 5;
  const int lane = WPF == 1 ? (threadIdx.x & 31) : threadIdx.x; // index of this thread within its frame
  const int f = WPF == 1 ? o.frame0 + blockIdx.x * FRAME_WARPS + warp : o.frame0 + blockIdx.x;
  if (f >= o.frame1) return;
  WarpState& ws = wstate[WPF == 1 ? warp : 0];
  uint8_t* llvl = line_lvl + (HAS_LINES && WPF == 1 ? warp : 0) * LINE_LVL_CAP;
  uint32_t mlvl = 0, slvl = 0; // levels of this thread's first 32 mono / stereo edges (bit k: edge begin + lane + k * STRIDE)
  const int m0 = d.mono_begin[f], m1 = d.mono_begin[f + 1];
  const int s0 = d.stereo_begin[f], s1 = d.stereo_begin[f + 1];
  const int ml0 = HAS_LINES ? d.mline_begin[f] : 0, ml1 = HAS_LINES ? d.mline_begin[f + 1] : 0;
  const int sl0 = HAS_LINES ? d.sline_begin[f] : 0, sl1 = HAS_LINES ? d.sline_begin[f + 1] : 0;
  const int n_edges = (m1 - m0) + (s1 - s0) + (ml1 - ml0) + (sl1 - sl0);
  BA_CHECK(0 <= m0 && m0 <= m1 && m1 <= d.n_mono && 0 <= s0 && s0 <= s1 && s1 <= d.n_stereo);
  BA_CHECK(!HAS_LINES || (0 <= ml0 && ml0 <= ml1 && ml1 <= d.n_mline && 0 <= sl0 && sl0 <= sl1 && sl1 <= d.n_sline));

  // ---- setup: pose (g2o_optimization.cc:271), flags, levels
  PoseRt T; // optimiser pose Tcw as rotation matrix + translation (quaternion only at the boundary)
  {
    const double p[3] = {d.pose_twc[f], d.pose_twc[d.n_frames + f], d.pose_twc[2 * d.n_frames + f]};
    const double q[4] = {d.pose_twc[3 * d.n_frames + f], d.pose_twc[4 * d.n_frames + f],
                         d.pose_twc[5 * d.n_frames + f], d.pose_twc[6 * d.n_frames + f]};
    const Pose Tq = pose_from_twc(p, q);
    quat_to_R(Tq.q, T.R);
    T.t[0] = Tq.t[0];
    T.t[1] = Tq.t[1];
    T.t[2] = Tq.t[2];
  }
  if (lane == 0) {
    ws.T0 = T;
    ws.Te = T;
#pragma unroll
    for (int i = 0; i < 6; ++i) ws.x[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) ws.st.iters[i] = ws.st.trials[i] = 0;
    ws.st.edges_linearized = ws.st.edges_evaluated = 0;
  }
  for (int e = m0 + lane; e < m1; e += STRIDE) {
    d.mono_lvl[e] = 0;
    d.mono_inl[e] = d.mono_inl_in ? d.mono_inl_in[e] : 1;
  }
  for (int e = s0 + lane; e < s1; e += STRIDE) {
    d.stereo_lvl[e] = 0;
    d.stereo_inl[e] = d.stereo_inl_in ? d.stereo_inl_in[e] : 1;
  }
  if (HAS_LINES) {
    for (int e = ml0 + lane; e < ml1; e += STRIDE) {
      d.mline_lvl[e] = 0;
      d.mline_inl[e] = d.mline_inl_in ? d.mline_inl_in[e] : 1;
    }
    for (int e = sl0 + lane; e < sl1; e += STRIDE) {
      d.sline_lvl[e] = 0;
      d.sline_inl[e] = d.sline_inl_in ? d.sline_inl_in[e] : 1;
    }
    for (int i = lane; i < LINE_LVL_CAP; i += STRIDE) llvl[i] = 0;
  }
  frame_sync<WPF>();

  bool robust = true;
  int n_active = n_edges; // every edge starts at level 0
  int num_outlier = 0;
  double lambda = 0, ni = 2, chi_cur = 0;

  for (int round = 0; round < o.rounds; ++round) {
    const int sr = round < 4 ? round : 3;
    T = ws.T0; // frame_vertex->setEstimate(initial) (:340)
    if (n_active > 0) {
      // ---------------- optimizer.optimize(iters) ----------------
      for (int it = 0; it < o.iters; ++it) {
        {
          double acc[NACC];
#pragma unroll
          for (int k = 0; k < NACC; ++k) acc[k] = 0;
          edge_pass<true, SINGLE_CAM, STRIDE>(d, o, m0, m1, s0, s1, mlvl, slvl, T.R, T.t, robust, lane, acc);
          if (HAS_LINES) {
            line_pass_rows<SINGLE_CAM, STRIDE>(d, o, ml0, ml1, sl0, sl1, llvl, T.R, T.t, robust, lane, acc);
          }
          frame_allreduce<WPF, NACC>(acc, red);
          if (it == 0) { // computeLambdaInit: tau * max diag, ni = 2
            double mx = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) mx = fmax(fabs(acc[up6(i, i)]), mx);
            lambda = 1e-5 * mx;
            ni = 2;
          }
          chi_cur = acc[NACC - 1];
          frame_sync<WPF>();
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 21; ++k) ws.H[k] = acc[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) ws.b[k] = acc[21 + k];
            ws.Te = T;
          }
          frame_sync<WPF>();
        }
        if (lane == 0) {
          ws.st.edges_linearized += n_active;
          ws.st.edges_evaluated += n_active;
        }
        int qmax = 0;
        double rho = 0;
        bool retry;
        do {
          // ---- trial: backup, solve, update, evaluate
          double x[6], b[6];
          bool ok;
          {
            double H[21];
#pragma unroll
            for (int k = 0; k < 21; ++k) H[k] = ws.H[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              b[k] = ws.b[k];
              x[k] = ws.x[k]; // stale x survives a failed factorisation (LinearSolverEigen returns early)
            }
            ok = solve6(H, b, lambda, x);
          }
          PoseRt Tn;
          poseRt_oplus(T, x, Tn);
          frame_sync<WPF>();
          if (lane == 0) {
            ws.Tbackup = T;
            ws.Te = Tn;
#pragma unroll
            for (int k = 0; k < 6; ++k) ws.x[k] = x[k];
          }
          frame_sync<WPF>();
          T = Tn;
          double tempChi;
          {
            double a2[NACC];
            a2[NACC - 1] = 0;
            edge_pass<false, SINGLE_CAM, STRIDE>(d, o, m0, m1, s0, s1, mlvl, slvl, T.R, T.t, robust, lane, a2);
            if (HAS_LINES) {
              line_pass_edges<SINGLE_CAM, STRIDE>(d, o, ml0, ml1, sl0, sl1, llvl, T.R, T.t, robust, lane, a2);
            }
            frame_allreduce<WPF, 1>(a2 + NACC - 1, red);
            tempChi = a2[NACC - 1];
          }
          if (lane == 0) {
            ws.st.edges_evaluated += n_active;
            ws.st.trials[sr]++;
          }
          if (!ok) tempChi = DBL_MAX;
          rho = chi_cur - tempChi;
          double scale = 0;
#pragma unroll
          for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
          scale += 1e-3;
          rho /= scale;
          bool stop_lambda = false;
          if (rho > 0 && isfinite(tempChi)) {
            const double c = 2 * rho - 1;
            double alpha = 1. - c * c * c;
            alpha = fmin(alpha, 2. / 3.);
            const double scaleFactor = fmax(1. / 3., alpha);
            lambda *= scaleFactor;
            ni = 2;
            chi_cur = tempChi;
          } else {
            lambda *= ni;
            ni *= 2;
            T = ws.Tbackup;
            if (!isfinite(lambda)) stop_lambda = true;
          }
          if (!stop_lambda) qmax++;
          retry = !stop_lambda && rho < 0 && qmax < 10;
        } while (retry);
        if (lane == 0) ws.st.iters[sr]++;
        if (qmax == 10 || rho == 0 || !isfinite(lambda)) break; // Terminate
      }
    }
    // ---------------- classification (:344-385) ----------------
    const PoseRt Tev = ws.Te;
    const double* Re = Tev.R;
    const double* te = Tev.t;
    int my_out = 0;
    for (int e = m0 + lane, k = 0; e < m1; e += STRIDE, ++k) {
      Cam camv;
      if (!SINGLE_CAM) load_cam(d.cameras, d.mono_cam[e], camv);
      const Cam& cam = SINGLE_CAM ? o.cam0 : camv;
      const double X[3] = {d.mono_xw[e], d.mono_xw[d.n_mono + e], d.mono_xw[2 * d.n_mono + e]};
      const double m[2] = {d.mono_meas[e], d.mono_meas[d.n_mono + e]};
      // active edges keep the error of their last evaluation; the reference recomputes only
      // edges whose ->inlier is false (:347-349), at the current estimate
      const bool recompute = !d.mono_inl[e];
      const bool was_active = !edge_excluded(mlvl, d.mono_lvl, k, e) && n_active > 0;
      double Xc[3], r[2];
      if (recompute || !was_active) transform_point(T.R, T.t, X, Xc);
      else transform_point(Re, te, X, Xc);
      point_residual<false>(cam, cam.bf, Xc, m, r);
      const float chi2 = (float)(r[0] * r[0] + r[1] * r[1]);
      const bool out = (double)chi2 > o.thr_mono;
      d.mono_inl[e] = out ? 0 : 1;
      if (k < 32) mlvl = (mlvl & ~(1u << k)) | ((uint32_t)out << k);
      else d.mono_lvl[e] = out ? 1 : 0;
      my_out += out;
    }
    for (int e = s0 + lane, k = 0; e < s1; e += STRIDE, ++k) {
      Cam camv;
      if (!SINGLE_CAM) load_cam(d.cameras, d.stereo_cam[e], camv);
      const Cam& cam = SINGLE_CAM ? o.cam0 : camv;
      const double X[3] = {d.stereo_xw[e], d.stereo_xw[d.n_stereo + e], d.stereo_xw[2 * d.n_stereo + e]};
      const double m[3] = {d.stereo_meas[e], d.stereo_meas[d.n_stereo + e], d.stereo_meas[2 * d.n_stereo + e]};
      const bool recompute = !d.stereo_inl[e];
      const bool was_active = !edge_excluded(slvl, d.stereo_lvl, k, e) && n_active > 0;
      double Xc[3], r[3];
      if (recompute || !was_active) transform_point(T.R, T.t, X, Xc);
      else transform_point(Re, te, X, Xc);
      point_residual<true>(cam, cam.bf, Xc, m, r);
      const float chi2 = (float)(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      const bool out = (double)chi2 > o.thr_stereo;
      d.stereo_inl[e] = out ? 0 : 1;
      if (k < 32) slvl = (slvl & ~(1u << k)) | ((uint32_t)out << k);
      else d.stereo_lvl[e] = out ? 1 : 0;
      my_out += out;
    }
    if (HAS_LINES) { // same classification for the line edges (extension)
      for (int e = ml0 + lane; e < ml1; e += STRIDE) {
        Cam camv;
        if (!SINGLE_CAM) load_cam(d.cameras, d.mline_cam[e], camv);
        const Cam& cam = SINGLE_CAM ? o.cam0 : camv;
        double L[6], m[4], r[2];
#pragma unroll
        for (int q = 0; q < 6; ++q) L[q] = d.mline_lw[(size_t)q * d.n_mline + e];
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = d.mline_meas[(size_t)q * d.n_mline + e];
        const bool recompute = !d.mline_inl[e];
        const int ei = e - ml0;
        const bool was_active = !line_excluded(d, llvl, ei, false, e) && n_active > 0;
        if (recompute || !was_active) line_residual<false>(cam, T.R, T.t, L, m, r);
        else line_residual<false>(cam, Re, te, L, m, r);
        const float chi2 = (float)(0.1 * (r[0] * r[0] + r[1] * r[1]));
        const bool out = (double)chi2 > o.thr_mline;
        d.mline_inl[e] = out ? 0 : 1;
        if (ei < LINE_LVL_CAP) llvl[ei] = out ? 1 : 0;
        else d.mline_lvl[e] = out ? 1 : 0;
        my_out += out;
      }
      for (int e = sl0 + lane; e < sl1; e += STRIDE) {
        Cam camv;
        if (!SINGLE_CAM) load_cam(d.cameras, d.sline_cam[e], camv);
        const Cam& cam = SINGLE_CAM ? o.cam0 : camv;
        double L[6], m[8], r[4];
#pragma unroll
        for (int q = 0; q < 6; ++q) L[q] = d.sline_lw[(size_t)q * d.n_sline + e];
#pragma unroll
        for (int q = 0; q < 8; ++q) m[q] = d.sline_meas[(size_t)q * d.n_sline + e];
        const bool recompute = !d.sline_inl[e];
        const int ei = (ml1 - ml0) + (e - sl0);
        const bool was_active = !line_excluded(d, llvl, ei, true, e) && n_active > 0;
        if (recompute || !was_active) line_residual<true>(cam, T.R, T.t, L, m, r);
        else line_residual<true>(cam, Re, te, L, m, r);
        const float chi2 = (float)(0.1 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]));
        const bool out = (double)chi2 > o.thr_sline;
        d.sline_inl[e] = out ? 0 : 1;
        if (ei < LINE_LVL_CAP) llvl[ei] = out ? 1 : 0;
        else d.sline_lvl[e] = out ? 1 : 0;
        my_out += out;
      }
    }
    num_outlier = __reduce_add_sync(0xffffffffu, my_out);
    if (WPF > 1) {
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red_i[warp] = num_outlier;
      __syncthreads();
      num_outlier = 0;
#pragma unroll
      for (int w2 = 0; w2 < (WPF > 1 ? WPF : 1); ++w2) num_outlier += red_i[w2];
    }
    n_active = n_edges - num_outlier;
    if (round == 2) robust = false; // e->setRobustKernel(0) (:364,:384)
    frame_sync<WPF>(); // the line levels in shared memory are read by other lanes (row mapping of line_pass_rows)
    if (n_edges < 10) break; // optimizer.edges().size() < 10 (:387)
  }

  // ---- recover optimized data (:391-396)
  if (lane == 0) {
    Pose Tq;
    R_to_quat(T.R, Tq.q);
    Tq.t[0] = T.t[0];
    Tq.t[1] = T.t[1];
    Tq.t[2] = T.t[2];
    pose_normalize(Tq);
    const Pose Twc = pose_inverse(Tq);
    d.out_pose_twc[f] = Twc.t[0];
    d.out_pose_twc[d.n_frames + f] = Twc.t[1];
    d.out_pose_twc[2 * d.n_frames + f] = Twc.t[2];
    d.out_pose_twc[3 * d.n_frames + f] = Twc.q[0];
    d.out_pose_twc[4 * d.n_frames + f] = Twc.q[1];
    d.out_pose_twc[5 * d.n_frames + f] = Twc.q[2];
    d.out_pose_twc[6 * d.n_frames + f] = Twc.q[3];
    d.num_inliers[f] = n_edges - num_outlier;
    ws.st.final_chi2 = chi_cur;
    ws.st.final_lambda = lambda;
    if (d.stats) reinterpret_cast<DevStats*>(d.stats)[f] = ws.st;
  }
}

} // namespace ba
