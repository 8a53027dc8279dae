// line_endpoints.cuh — batched endpoint refresh of map lines, SURVEY 8(f) rank 2: the step that follows the local
// bundle adjustment and consumes the optimised Line3Ds (Map::LocalMapOptimization calls Map::UppdateMapline for every
// optimised line, /root/reference/src/map.cc:790-797). Replaces the body of Map::UppdateMapline (map.cc:121-177):
//   (anchor, direction) = Line3D::toCartesian()                                                   (:143)
//   among the line's map points within 0.2 of the line (EigenPointLineDistance3D, line_processor.cc:77-90) take the
//   largest and the smallest coordinate along the direction's main axis md                         (:148-164)
//   endpoints = anchor + (coordinate - anchor[md]) / direction[md] * direction                     (:168-172)
// including its quirk: the running maximum starts at DBL_MIN (the smallest positive double), so a line whose nearby
// points all have a non-positive main coordinate is not refreshed.
// LINE_EP_LANES lanes per line stride over the line's CSR segment of point indices (coalesced index reads, the
// gathers of a line in flight together); minimum / maximum are order-independent, so the result does not depend on the
// lane count. The conversion to (anchor, direction) runs lane-per-line (see the kernel). All arithmetic is written
// with the non-contracting intrinsics in the reference's operation order (the reference is built without FMA),
// toCartesian included: Eigen's LDLT<Matrix3d> pivots on the raw diagonal, so the permutation is known up front and
// the factorisation runs on the permuted matrix with static indices.
// HBM: 101 bytes per line + 4 bytes per point reference; the point gathers (24 bytes each) hit L2 (points are shared
// between lines and the whole point array of a local map is a few MB).
#pragma once

#include <cfloat>

#include "ba_math.cuh"

namespace ba {

constexpr int LINE_EP_LANES = 8;
constexpr int LINE_EP_THREADS = 128;
// (register cap A/B with -DLINE_EP_MIN_BLOCKS=8 / 10, profiles/scripts/ends_regcap_ab.sh: 80 registers 72 us; capped to
// 64 / 48 registers -- 8 / 10 CTAs per SM, 40 / 96 bytes of spills -- 77 us both: not kept)
#ifdef LINE_EP_MIN_BLOCKS
#define LINE_EP_BOUNDS __launch_bounds__(LINE_EP_THREADS, LINE_EP_MIN_BLOCKS)
#else
#define LINE_EP_BOUNDS __launch_bounds__(LINE_EP_THREADS)
#endif

struct LineEndpointsDev {
  int n_lines, n_points;
  size_t line_stride, point_stride; // plane strides of line_wd and point_xyz (>= n_lines, >= n_points)
  const double* line_wd;   // [6][line_stride] g2o::Line3D [w, d]
  const int* pt_begin;     // [n_lines + 1]
  const int* pt_index;     // [pt_begin[n_lines]] into point_xyz
  const double* point_xyz; // [3][point_stride]
  double* endpoints;       // [6][n_lines], written where ok
  uint8_t* out_ok;         // [n_lines]
  int* n_done;
};

BA_DEV double sel3(double v0, double v1, double v2, int i) { return i == 0 ? v0 : (i == 1 ? v1 : v2); }

// g2o::Line3D::toCartesian (g2o types/slam3d_addons/line3d.cpp): direction d / |d|, anchor =
// (W^T W + 1e-9 I).ldlt().solve(W^T w), W = -skew(d)
BA_DEV void line_to_cartesian(const double (&w)[3], const double (&d)[3], double (&anchor)[3], double (&dir)[3]) {
  const double dx = d[0], dy = d[1], dz = d[2];
  const double xx = __dmul_rn(dx, dx), yy = __dmul_rn(dy, dy), zz = __dmul_rn(dz, dz);
  const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(xx, yy), zz));
  dir[0] = __ddiv_rn(dx, nrm);
  dir[1] = __ddiv_rn(dy, nrm);
  dir[2] = __ddiv_rn(dz, nrm);
  const double a00 = __dadd_rn(__dadd_rn(zz, yy), 1e-9), a11 = __dadd_rn(__dadd_rn(zz, xx), 1e-9), a22 = __dadd_rn(__dadd_rn(yy, xx), 1e-9);
  const double a10 = -__dmul_rn(dx, dy), a20 = -__dmul_rn(dx, dz), a21 = -__dmul_rn(dy, dz);
  const double r0 = __dadd_rn(-__dmul_rn(dz, w[1]), __dmul_rn(dy, w[2]));
  const double r1 = __dadd_rn(__dmul_rn(dz, w[0]), -__dmul_rn(dx, w[2]));
  const double r2 = __dadd_rn(-__dmul_rn(dy, w[0]), __dmul_rn(dx, w[1]));
  // pivot order from the raw diagonal (first maximum on ties), as two transpositions
  int p0 = 0, p1 = 1, p2 = 2;
  {
    int big = 0;
    if (fabs(a11) > fabs(a00)) big = 1;
    if (fabs(a22) > fabs(sel3(a00, a11, a22, big))) big = 2;
    if (big == 1) { p0 = 1; p1 = 0; }
    if (big == 2) { p0 = 2; p2 = 0; }
    if (fabs(sel3(a00, a11, a22, p2)) > fabs(sel3(a00, a11, a22, p1))) { const int t = p1; p1 = p2; p2 = t; }
  }
  // B = P A P^T (lower triangle); the off-diagonal entry of the index pair {p, q} is a10, a20, a21 for p + q = 1, 2, 3
  const double b00 = sel3(a00, a11, a22, p0);
  double b11 = sel3(a00, a11, a22, p1), b22 = sel3(a00, a11, a22, p2);
  double l10 = sel3(a10, a20, a21, p1 + p0 - 1), l20 = sel3(a10, a20, a21, p2 + p0 - 1), l21 = sel3(a10, a20, a21, p2 + p1 - 1);
  if (fabs(b00) > 0.0) {
    l10 = __ddiv_rn(l10, b00);
    l20 = __ddiv_rn(l20, b00);
  }
  {
    const double t0 = __dmul_rn(b00, l10);
    b11 = __dsub_rn(b11, __dmul_rn(l10, t0));
    l21 = __dsub_rn(l21, __dmul_rn(l20, t0));
    if (fabs(b11) > 0.0) l21 = __ddiv_rn(l21, b11);
  }
  {
    const double t0 = __dmul_rn(b00, l20), t1 = __dmul_rn(b11, l21);
    b22 = __dsub_rn(b22, __dadd_rn(__dmul_rn(l20, t0), __dmul_rn(l21, t1)));
  }
  double y0 = sel3(r0, r1, r2, p0), y1 = sel3(r0, r1, r2, p1), y2 = sel3(r0, r1, r2, p2);
  y1 = __dsub_rn(y1, __dmul_rn(y0, l10));
  y2 = __dsub_rn(y2, __dmul_rn(y0, l20));
  y2 = __dsub_rn(y2, __dmul_rn(y1, l21));
  const double tol = 1.0 / DBL_MAX;
  y0 = fabs(b00) > tol ? __ddiv_rn(y0, b00) : 0.0;
  y1 = fabs(b11) > tol ? __ddiv_rn(y1, b11) : 0.0;
  y2 = fabs(b22) > tol ? __ddiv_rn(y2, b22) : 0.0;
  y1 = __dsub_rn(y1, __dmul_rn(l21, y2));
  y0 = __dsub_rn(y0, __dadd_rn(__dmul_rn(l10, y1), __dmul_rn(l20, y2)));
#pragma unroll
  for (int j = 0; j < 3; ++j) anchor[j] = p0 == j ? y0 : (p1 == j ? y1 : y2);
}

// One warp per 32 lines. Lane j owns line L0 + j: it reads the Line3D and the CSR bounds (coalesced), converts to
// (anchor, direction) -- the expensive part, a dozen fp64 divisions, done once per line and not once per lane group --
// and later writes the endpoints (coalesced). The point segments are walked in 8 rounds of 4 lines, LINE_EP_LANES lanes
// per line, with the owner's anchor / direction broadcast by shuffles; the first point index of every round is
// fetched up front so that only the gathers of a round wait for each other.
__global__ void LINE_EP_BOUNDS line_endpoints_kernel(const __grid_constant__ LineEndpointsDev d) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int GROUPS = 32 / LINE_EP_LANES, ROUNDS = 32 / GROUPS;
  const int lane = threadIdx.x & 31, sub = lane % LINE_EP_LANES, grp = lane / LINE_EP_LANES;
  const int l = blockIdx.x * LINE_EP_THREADS + threadIdx.x;
  const bool live = l < d.n_lines;
  int begin = 0, end = 0;
  double anchor[3] = {0.0, 0.0, 0.0}, dir[3] = {0.0, 0.0, 1.0};
  if (live) {
    begin = d.pt_begin[l];
    end = d.pt_begin[l + 1];
    BA_CHECK(begin >= 0 && end >= begin);
  }
  // first point of this lane in every round (-1: none)
  int first[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int src = r * GROUPS + grp;
    const int o = __shfl_sync(FULL, begin, src) + sub;
    first[r] = o < __shfl_sync(FULL, end, src) ? d.pt_index[o] : -1;
  }
  int md = 0;
  if (end > begin) {
    double w[3], dd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      w[k] = d.line_wd[(size_t)k * d.line_stride + l];
      dd[k] = d.line_wd[(size_t)(3 + k) * d.line_stride + l];
    }
    line_to_cartesian(w, dd, anchor, dir);
    if (fabs(dir[1]) > fabs(dir[0])) md = 1;
    if (fabs(dir[2]) > fabs(sel3(dir[0], dir[1], dir[2], md))) md = 2;
  }
  double my_max = DBL_MIN, my_min = DBL_MAX;
  // (gathering the first point of round r + 1 while round r computes was measured: 94 registers instead of 80, one
  // resident CTA fewer per SM, 84 us instead of 81 us at 307 k lines -- not kept)
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int src = r * GROUPS + grp;
    const double a0 = __shfl_sync(FULL, anchor[0], src), a1 = __shfl_sync(FULL, anchor[1], src), a2 = __shfl_sync(FULL, anchor[2], src);
    const double v0 = __shfl_sync(FULL, dir[0], src), v1 = __shfl_sync(FULL, dir[1], src), v2 = __shfl_sync(FULL, dir[2], src);
    const int smd = __shfl_sync(FULL, md, src);
    const int s_end = __shfl_sync(FULL, end, src);
    int o = __shfl_sync(FULL, begin, src) + sub;
    int pi = first[r];
    double max_d = DBL_MIN, min_d = DBL_MAX;
    while (pi >= 0) {
      BA_CHECK(pi < d.n_points);
      const double X0 = d.point_xyz[pi], X1 = d.point_xyz[d.point_stride + pi], X2 = d.point_xyz[2 * d.point_stride + pi];
      o += LINE_EP_LANES;
      pi = o < s_end ? d.pt_index[o] : -1;
      const double q0 = __dsub_rn(X0, a0), q1 = __dsub_rn(X1, a1), q2 = __dsub_rn(X2, a2);
      const double c0 = __dsub_rn(__dmul_rn(v1, q2), __dmul_rn(v2, q1));
      const double c1 = __dsub_rn(__dmul_rn(v2, q0), __dmul_rn(v0, q2));
      const double c2 = __dsub_rn(__dmul_rn(v0, q1), __dmul_rn(v1, q0));
      const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, c0), __dmul_rn(c1, c1)), __dmul_rn(c2, c2)));
      const double di = sel3(X0, X1, X2, smd);
      const bool near = !(dist > 0.2); // `if (dist[i] > 0.2) continue;`
      if (near && di > max_d) max_d = di;
      if (near && di < min_d) min_d = di;
    }
    // (max_d > DBL_MIN <=> some coordinate passed `di > max_point_d`; likewise for the minimum)
#pragma unroll
    for (int s = LINE_EP_LANES / 2; s > 0; s >>= 1) {
      const double om = __shfl_xor_sync(FULL, max_d, s), on = __shfl_xor_sync(FULL, min_d, s);
      max_d = om > max_d ? om : max_d;
      min_d = on < min_d ? on : min_d;
    }
    // hand the result of line r * GROUPS + g (held by every lane of group g) to its owner lane
    const double tmax = __shfl_sync(FULL, max_d, (lane % GROUPS) * LINE_EP_LANES), tmin = __shfl_sync(FULL, min_d, (lane % GROUPS) * LINE_EP_LANES);
    if (lane / GROUPS == r) {
      my_max = tmax;
      my_min = tmin;
    }
  }
  const bool ok = live && my_max > DBL_MIN && my_min < DBL_MAX;
  if (live) {
    d.out_ok[l] = ok ? 1 : 0;
    if (ok) {
      const double lm = sel3(anchor[0], anchor[1], anchor[2], md), vm = sel3(dir[0], dir[1], dir[2], md);
      const double r1 = __ddiv_rn(__dsub_rn(my_max, lm), vm), r2 = __ddiv_rn(__dsub_rn(my_min, lm), vm);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        d.endpoints[(size_t)k * d.n_lines + l] = __dadd_rn(anchor[k], __dmul_rn(r1, dir[k]));
        d.endpoints[(size_t)(3 + k) * d.n_lines + l] = __dadd_rn(anchor[k], __dmul_rn(r2, dir[k]));
      }
    }
  }
  const unsigned done = __ballot_sync(FULL, ok);
  if (lane == 0 && done) atomicAdd(d.n_done, __popc(done));
}

} // namespace ba
