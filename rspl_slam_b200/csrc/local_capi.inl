// local_capi.inl — C-ABI entry points of the LocalmapOptimization batch (included by capi.cu).
// Reference boundary: LocalmapOptimization(...) /root/reference/include/g2o_optimization/g2o_optimization.h:15-18.

namespace {

constexpr size_t LOCAL_STAGE_MAX = 4 << 20; // input / output range of a batch that goes through the pinned mirror

struct KindHost { // host view of one landmark kind of an RsplLocalBatch
  const int32_t* lm_begin;
  const int32_t* cls_begin[2];
  const int32_t* cls_pose[2];
  const int32_t* cls_lm[2];
  const int32_t* cls_cam[2];
  const double* cls_meas[2];
  const double* lm_in;
  int md[2]; // measurement planes per class
};

KindHost kind_host(const RsplLocalBatch* in, int kind) {
  KindHost h;
  if (kind == 0) {
    h.lm_begin = in->point_begin;
    h.cls_begin[0] = in->mono_pt_begin;
    h.cls_begin[1] = in->stereo_pt_begin;
    h.cls_pose[0] = in->mp_pose;
    h.cls_pose[1] = in->sp_pose;
    h.cls_lm[0] = in->mp_point;
    h.cls_lm[1] = in->sp_point;
    h.cls_cam[0] = in->mp_cam;
    h.cls_cam[1] = in->sp_cam;
    h.cls_meas[0] = in->mp_meas;
    h.cls_meas[1] = in->sp_meas;
    h.lm_in = in->point_xyz;
    h.md[0] = 2;
    h.md[1] = 3;
  } else {
    h.lm_begin = in->line_begin;
    h.cls_begin[0] = in->mono_ln_begin;
    h.cls_begin[1] = in->stereo_ln_begin;
    h.cls_pose[0] = in->ml_pose;
    h.cls_pose[1] = in->sl_pose;
    h.cls_lm[0] = in->ml_line;
    h.cls_lm[1] = in->sl_line;
    h.cls_cam[0] = in->ml_cam;
    h.cls_cam[1] = in->sl_cam;
    h.cls_meas[0] = in->ml_meas;
    h.cls_meas[1] = in->sl_meas;
    h.lm_in = in->line_wd;
    h.md[0] = 4;
    h.md[1] = 8;
  }
  return h;
}

} // namespace

extern "C" int rspl_ba_local_batch_upload(RsplBaContext* c, const RsplLocalBatch* in) {
  if (!c || !in) return RSPL_BA_ERR_INVALID;
  c->local_uploaded = c->local_solved = false;
  const int W = in->n_windows;
  if (W < 0 || in->n_cameras < 1 || in->n_cameras > 255 || !in->cameras)
    return fail(c, RSPL_BA_ERR_INVALID, "local batch: bad header");
  if (W == 0) {
    c->l_n_windows = 0;
    c->local_uploaded = true;
    return RSPL_BA_OK;
  }
  if (!offsets_ok(in->pose_begin, W) || !offsets_ok(in->point_begin, W) || !offsets_ok(in->line_begin, W) ||
      !offsets_ok(in->mono_pt_begin, W) || !offsets_ok(in->stereo_pt_begin, W) || !offsets_ok(in->mono_ln_begin, W) ||
      !offsets_ok(in->stereo_ln_begin, W))
    return fail(c, RSPL_BA_ERR_INVALID, "local batch: bad offsets");
  const int NP = in->pose_begin[W];
  if (NP <= 0 || !in->pose_twc || !in->pose_fixed) return fail(c, RSPL_BA_ERR_INVALID, "local batch: no poses");
  KindHost kh[2] = {kind_host(in, 0), kind_host(in, 1)};
  int n_lm[2], n_cls[2][2];
  for (int k = 0; k < 2; ++k) {
    n_lm[k] = kh[k].lm_begin[W];
    if (n_lm[k] && !kh[k].lm_in) return fail(c, RSPL_BA_ERR_INVALID, "local batch: null landmark array");
    for (int cl = 0; cl < 2; ++cl) {
      n_cls[k][cl] = kh[k].cls_begin[cl][W];
      if (n_cls[k][cl] && (!kh[k].cls_pose[cl] || !kh[k].cls_lm[cl] || !kh[k].cls_meas[cl]))
        return fail(c, RSPL_BA_ERR_INVALID, "local batch: null edge array");
      if (n_cls[k][cl] && in->n_cameras > 1 && !kh[k].cls_cam[cl])
        return fail(c, RSPL_BA_ERR_INVALID, "local batch: several cameras but no per-edge camera index");
      // (the O(edges) index checks run below, while the copies are in flight)
    }
  }
  // per-window limits of the shared-memory path
  int max_poses = 1, max_free = 1;
  c->batch_ready = false;
  c->l_nf_begin.assign(W + 1, 0);
  c->l_pair_base.assign(W + 1, 0);
  c->l_max_pts = c->l_max_lns = c->l_max_edges = 0;
  c->l_cost[0].assign(W, 0);
  c->l_cost[1].assign(W, 0);
  c->l_cost_nl[0].assign(W, 0);
  c->l_cost_nl[1].assign(W, 0);
  for (int w = 0; w < W; ++w) {
    const int a = in->pose_begin[w], b = in->pose_begin[w + 1];
    int nf = 0;
    for (int p = a; p < b; ++p) nf += in->pose_fixed[p] ? 0 : 1;
    if (b - a > max_poses) max_poses = b - a;
    if (nf > max_free) max_free = nf;
    c->l_nf_begin[w + 1] = c->l_nf_begin[w] + nf;
    const int npt = in->point_begin[w + 1] - in->point_begin[w], nln = in->line_begin[w + 1] - in->line_begin[w];
    const long long ne = (long long)(in->mono_pt_begin[w + 1] - in->mono_pt_begin[w]) +
                         (in->stereo_pt_begin[w + 1] - in->stereo_pt_begin[w]) +
                         (in->mono_ln_begin[w + 1] - in->mono_ln_begin[w]) +
                         (in->stereo_ln_begin[w + 1] - in->stereo_ln_begin[w]);
    c->l_cost[0][w] = (long long)(in->mono_pt_begin[w + 1] - in->mono_pt_begin[w]) + (in->stereo_pt_begin[w + 1] - in->stereo_pt_begin[w]);
    c->l_cost[1][w] = (long long)(in->mono_ln_begin[w + 1] - in->mono_ln_begin[w]) + (in->stereo_ln_begin[w + 1] - in->stereo_ln_begin[w]);
    c->l_cost_nl[0][w] = npt;
    c->l_cost_nl[1][w] = nln;
    if (npt > c->l_max_pts) c->l_max_pts = npt;
    if (nln > c->l_max_lns) c->l_max_lns = nln;
    if (ne > c->l_max_edges) c->l_max_edges = (int)ne;
    // a landmark with k free observers yields k(k+1)/2 <= k(nf+1)/2 pair entries; for large windows
    // that bound is far too loose, so count sum_l deg(l)(deg(l)+1)/2 over the window's landmarks
    long long cap = (ne * (nf + 1) + 1) / 2 + 2;
    if (cap > (1LL << 24)) {
      long long exact = 2;
      std::vector<int> deg;
      for (int k = 0; k < 2; ++k) {
        const int l0 = kh[k].lm_begin[w], nl = kh[k].lm_begin[w + 1] - l0;
        deg.assign(nl > 0 ? nl : 1, 0);
        for (int cl = 0; cl < 2; ++cl)
          for (int i = kh[k].cls_begin[cl][w]; i < kh[k].cls_begin[cl][w + 1]; ++i) {
            const int l = kh[k].cls_lm[cl][i]; // (range-checked here: the index validation runs later)
            if ((unsigned)l < (unsigned)nl) deg[l]++;
          }
        for (int l = 0; l < nl; ++l) exact += (long long)deg[l] * (deg[l] + 1) / 2;
      }
      if (exact < cap) cap = exact;
    }
    c->l_pair_base[w + 1] = c->l_pair_base[w] + cap;
  }
  if (max_poses > 65535) // the packed edge record keeps the pose in 16 bits
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: %d poses in one window (limit 65535)", max_poses);

  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  // Arena: every input first, then every output, then the work arrays — so that a small batch (the reference's call:
  // one window) can go up in ONE copy from a pinned mirror of the input range and come back in ONE copy of the
  // output range (local_stage_*), instead of ~30 + ~10 copies of a few KB each.
  struct KOff {
    size_t lm_begin, cls_begin[2], cls_pose[2], cls_lm[2], cls_cam[2], cls_meas[2], lm_in;
    size_t meas, info, lm, src, chi2, lvl, Zb, ebeg, cursor, newidx, orig, x, xb, H, b, y, act, slot, plist, pbeg, out_inl[2], lm_out;
  } ko[2];
  Arena a;
  const size_t o_cam = a.take(sizeof(double) * 5 * in->n_cameras);
  const size_t o_pb = a.take(sizeof(int) * (W + 1));
  const size_t o_ptw = a.take(sizeof(double) * 7 * NP);
  const size_t o_pfx = a.take(NP);
  for (int k = 0; k < 2; ++k) {
    const int SD = k ? 6 : 3;
    KOff& o = ko[k];
    o.lm_begin = a.take(sizeof(int) * (W + 1));
    for (int cl = 0; cl < 2; ++cl) {
      o.cls_begin[cl] = a.take(sizeof(int) * (W + 1));
      o.cls_pose[cl] = a.take(sizeof(int) * n_cls[k][cl]);
      o.cls_lm[cl] = a.take(sizeof(int) * n_cls[k][cl]);
      o.cls_cam[cl] = a.take(sizeof(int) * n_cls[k][cl]);
      o.cls_meas[cl] = a.take(sizeof(double) * kh[k].md[cl] * n_cls[k][cl]);
    }
    o.lm_in = a.take(sizeof(double) * SD * n_lm[k]);
  }
  const size_t o_in_end = a.off;
  // outputs
  const size_t o_err = a.take(sizeof(int) * 4); // error flag, max degree of points / lines, decide ticket
  const size_t o_pout = a.take(sizeof(double) * 7 * NP);
  const size_t o_stats = a.take(sizeof(ba::DevStats) * W);
  for (int k = 0; k < 2; ++k) {
    const int SD = k ? 6 : 3;
    for (int cl = 0; cl < 2; ++cl) ko[k].out_inl[cl] = a.take(n_cls[k][cl]);
    ko[k].lm_out = a.take(sizeof(double) * SD * n_lm[k]);
  }
  const size_t o_out_end = a.off;
  // work
  const size_t o_tcw = a.take(sizeof(double) * 7 * NP);
  const size_t o_sfi = a.take(sizeof(int) * NP);
  const int slot_stride = (max_free + 3) & ~3;
  // large-window setup variants (local_kernel.cuh): multi-CTA landmark ordering, pose lists by scatter + sort
  const int max_lm_w = c->l_max_pts > c->l_max_lns ? c->l_max_pts : c->l_max_lns;
  const bool big_order = max_lm_w > 16384, big_pose = max_poses > 64;
  const int order_chunks = big_order ? (max_lm_w + ba::ORDER_CHUNK - 1) / ba::ORDER_CHUNK : 0;
  const size_t o_ohist = big_order ? a.take(sizeof(int) * (size_t)W * 2 * order_chunks * 256) : 0;
  size_t o_pcur[2] = {0, 0}, o_ptmp[2] = {0, 0};
  for (int k = 0; k < 2 && big_pose; ++k) {
    o_pcur[k] = a.take(sizeof(int) * (size_t)NP);
    o_ptmp[k] = a.take(sizeof(int) * ((size_t)n_cls[k][0] + n_cls[k][1] + 1));
  }
  for (int k = 0; k < 2; ++k) {
    const int LD = k ? 4 : 3, SD = k ? 6 : 3, MD = k ? 8 : 3, HD = k ? 10 : 6, WD = 6 * LD;
    const size_t ne = (size_t)n_cls[k][0] + n_cls[k][1], nl = n_lm[k];
    KOff& o = ko[k];
    o.meas = a.take(sizeof(double) * MD * ne);
    o.info = a.take(sizeof(int) * ne);
    o.lm = a.take(sizeof(int) * ne);
    o.src = a.take(sizeof(int) * ne);
    o.chi2 = a.take(sizeof(double) * ne);
    o.lvl = a.take(ne);
    o.Zb = a.take(sizeof(double) * WD * ne);
    o.ebeg = a.take(sizeof(int) * (nl + 1));
    o.cursor = a.take(sizeof(int) * nl);
    o.newidx = a.take(sizeof(int) * nl);
    o.orig = a.take(sizeof(int) * nl);
    o.x = a.take(sizeof(double) * SD * nl);
    o.xb = a.take(sizeof(double) * SD * nl);
    o.H = a.take(sizeof(double) * HD * nl);
    o.b = a.take(sizeof(double) * LD * nl);
    o.y = a.take(sizeof(double) * LD * nl);
    o.act = a.take(nl);
    o.slot = a.take(nl * (size_t)slot_stride);
    o.plist = a.take(sizeof(int) * ne);
    o.pbeg = a.take(sizeof(int) * (NP + 1));
  }
  c->l_in_end = o_in_end;
  c->l_out_begin = o_in_end;
  c->l_out_end = o_out_end;
  // small batches: stage the inputs in a pinned mirror of [0, o_in_end) and send it with one copy
  const bool staged = o_out_end <= LOCAL_STAGE_MAX; // (inputs [0, o_in_end) and outputs [o_in_end, o_out_end) both inside the mirror)
  if (staged && c->l_stage_cap < LOCAL_STAGE_MAX) {
    if (c->l_stage) cudaFreeHost(c->l_stage);
    c->l_stage = nullptr;
    c->l_stage_cap = 0;
    if (cudaHostAlloc(&c->l_stage, LOCAL_STAGE_MAX, cudaHostAllocDefault) == cudaSuccess) c->l_stage_cap = LOCAL_STAGE_MAX;
    else cudaGetLastError();
  }
  c->l_staged = staged && c->l_stage_cap >= LOCAL_STAGE_MAX;
  CU_TRY(c, c->local_buf.reserve(a.off));
  char* base = c->local_buf.as<char>();
  cudaStream_t s = c->stream;
  char* stage = c->l_staged ? (char*)c->l_stage : nullptr;
#define H2D(off, src, bytes)                                                                              \
  do {                                                                                                    \
    if ((bytes) > 0) {                                                                                    \
      if (stage) memcpy(stage + (off), (src), (bytes));                                                   \
      else CU_TRY(c, cudaMemcpyAsync(base + (off), (src), (bytes), cudaMemcpyHostToDevice, s));           \
    }                                                                                                     \
  } while (0)
  // component planes [comps][n]: one copy when the caller's planes are contiguous, one per component when the batch is
  // a window range of a larger one (stride = the plane length of the whole batch)
#define H2D_PLANES(off, src, comps, n, stride)                                                                       \
  do {                                                                                                               \
    if ((stride) == 0 || (stride) == (size_t)(n)) H2D((off), (src), sizeof(double) * (size_t)(comps) * (size_t)(n));  \
    else                                                                                                             \
      for (int q_ = 0; q_ < (comps); ++q_)                                                                           \
        H2D((off) + sizeof(double) * (size_t)q_ * (size_t)(n), (src) + (size_t)q_ * (stride), sizeof(double) * (size_t)(n)); \
  } while (0)
  H2D(o_cam, in->cameras, sizeof(double) * 5 * in->n_cameras);
  H2D(o_pb, in->pose_begin, sizeof(int) * (W + 1));
  H2D_PLANES(o_ptw, in->pose_twc, 7, NP, c->l_stride.pose);
  H2D(o_pfx, in->pose_fixed, (size_t)NP);
  for (int k = 0; k < 2; ++k) {
    const int SD = k ? 6 : 3;
    H2D(ko[k].lm_begin, kh[k].lm_begin, sizeof(int) * (W + 1));
    H2D_PLANES(ko[k].lm_in, kh[k].lm_in, SD, n_lm[k], c->l_stride.lm[k]);
    for (int cl = 0; cl < 2; ++cl) {
      H2D(ko[k].cls_begin[cl], kh[k].cls_begin[cl], sizeof(int) * (W + 1));
      H2D(ko[k].cls_pose[cl], kh[k].cls_pose[cl], sizeof(int) * n_cls[k][cl]);
      H2D(ko[k].cls_lm[cl], kh[k].cls_lm[cl], sizeof(int) * n_cls[k][cl]);
      // (one camera: the indices are validated on the host below and never read on the device, so they stay there)
      if (kh[k].cls_cam[cl] && in->n_cameras > 1) H2D(ko[k].cls_cam[cl], kh[k].cls_cam[cl], sizeof(int) * n_cls[k][cl]);
      H2D_PLANES(ko[k].cls_meas[cl], kh[k].cls_meas[cl], kh[k].md[cl], n_cls[k][cl], c->l_stride.cls[k][cl]);
    }
  }
#undef H2D_PLANES
#undef H2D
  if (stage) CU_TRY(c, cudaMemcpyAsync(base, stage, o_in_end, cudaMemcpyHostToDevice, s));
  // index validation on the host while the DMA engine works (a rejected batch has been copied for nothing,
  // but is never solved: local_uploaded stays false)
  const char* bad = nullptr;
  for (int k = 0; k < 2 && !bad; ++k)
    for (int cl = 0; cl < 2 && !bad; ++cl) {
      if (!n_cls[k][cl]) continue;
      if (!indices_ok(kh[k].cls_pose[cl], kh[k].cls_begin[cl], in->pose_begin, W) ||
          !indices_ok(kh[k].cls_lm[cl], kh[k].cls_begin[cl], kh[k].lm_begin, W))
        bad = "local batch: edge references a vertex outside its window";
      else if (!cams_ok(kh[k].cls_cam[cl], n_cls[k][cl], in->n_cameras))
        bad = "local batch: id_camera out of range";
    }
  CU_TRY(c, cudaStreamSynchronize(s));
  if (bad) return fail(c, RSPL_BA_ERR_INVALID, "%s", bad);

  ba::LocalDev& d = c->ld;
  d.n_windows = W;
  d.n_cameras = in->n_cameras;
  d.n_poses = NP;
  d.cameras = (const double*)(base + o_cam);
  d.pose_begin = (const int*)(base + o_pb);
  d.pose_twc = (const double*)(base + o_ptw);
  d.pose_fixed = (const uint8_t*)(base + o_pfx);
  d.pose_tcw = (double*)(base + o_tcw);
  d.pose_out = (double*)(base + o_pout);
  d.slot_stride = slot_stride;
  d.use_slots = max_free <= ba::PAIRS_LM_MIN_NF ? 1 : 0;
  d.stats = (void*)(base + o_stats);
  d.err = (int*)(base + o_err);
  d.maxdeg = d.err + 1;
  d.setup_free_idx = (int*)(base + o_sfi);
  d.order_hist = big_order ? (int*)(base + o_ohist) : nullptr;
  d.order_chunks = order_chunks;
  for (int k = 0; k < 2; ++k) {
    d.pose_cursor[k] = big_pose ? (int*)(base + o_pcur[k]) : nullptr;
    d.plist_tmp[k] = big_pose ? (int*)(base + o_ptmp[k]) : nullptr;
  }
  for (int k = 0; k < 2; ++k) {
    ba::KindDev& kd = d.k[k];
    const KOff& o = ko[k];
    kd.n_lm = n_lm[k];
    kd.n_edge = n_cls[k][0] + n_cls[k][1];
    kd.lm_begin = (const int*)(base + o.lm_begin);
    for (int cl = 0; cl < 2; ++cl) {
      kd.cls_begin[cl] = (const int*)(base + o.cls_begin[cl]);
      kd.cls_pose[cl] = (const int*)(base + o.cls_pose[cl]);
      kd.cls_lm[cl] = (const int*)(base + o.cls_lm[cl]);
      kd.cls_cam[cl] = (kh[k].cls_cam[cl] && in->n_cameras > 1) ? (const int*)(base + o.cls_cam[cl]) : nullptr;
      kd.cls_meas[cl] = (const double*)(base + o.cls_meas[cl]);
      kd.cls_n[cl] = n_cls[k][cl];
      kd.out_inl[cl] = (uint8_t*)(base + o.out_inl[cl]);
    }
    kd.lm_in = (const double*)(base + o.lm_in);
    kd.meas = (double*)(base + o.meas);
    kd.info = (int*)(base + o.info);
    kd.lm = (int*)(base + o.lm);
    kd.src = (int*)(base + o.src);
    kd.chi2 = (double*)(base + o.chi2);
    kd.lvl = (uint8_t*)(base + o.lvl);
    kd.Z = (double*)(base + o.Zb);
    kd.ebeg = (int*)(base + o.ebeg);
    kd.cursor = (int*)(base + o.cursor);
    kd.newidx = (int*)(base + o.newidx);
    kd.orig = (int*)(base + o.orig);
    kd.x = (double*)(base + o.x);
    kd.xb = (double*)(base + o.xb);
    kd.H = (double*)(base + o.H);
    kd.b = (double*)(base + o.b);
    kd.y = (double*)(base + o.y);
    kd.act = (uint8_t*)(base + o.act);
    kd.slot = (uint8_t*)(base + o.slot);
    kd.plist = (int*)(base + o.plist);
    kd.pbeg = (int*)(base + o.pbeg);
    kd.lm_out = (double*)(base + o.lm_out);
  }
  c->l_n_windows = W;
  c->l_np = NP;
  c->l_npt = n_lm[0];
  c->l_nln = n_lm[1];
  c->l_n[0] = n_cls[0][0];
  c->l_n[1] = n_cls[0][1];
  c->l_n[2] = n_cls[1][0];
  c->l_n[3] = n_cls[1][1];
  c->l_max_poses = max_poses;
  c->l_max_free_poses = max_free;
  c->local_uploaded = true;
  return RSPL_BA_OK;
}

namespace {

// Allocates (grow-only) and wires the HBM state of the batched path for the uploaded batch.
int batched_prepare(RsplBaContext* c) {
  if (c->batch_ready && c->batch_global == c->global_mode) return RSPL_BA_OK;
  c->batch_global = c->global_mode;
  const int W = c->l_n_windows, NP = c->l_np;
  const int NF = c->l_nf_begin[W];
  const int NFmax = c->l_max_free_poses;
  const int Pmax = NFmax * (NFmax + 1) / 2;
  const int Cp = (c->l_max_pts + ba::BT - 1) / ba::BT, Cl = (c->l_max_lns + ba::BT - 1) / ba::BT;
  const int C = Cp + Cl > 0 ? Cp + Cl : 1;
  const long long n_pairs = c->l_pair_base[W];
  if (n_pairs > 0x7fffffffLL) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: pair lists exceed 2^31 entries");
  if ((long long)NFmax * (NFmax + 1) / 2 > 0x3fffffffLL)
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: window too large for the batched path");
  Arena a;
  const size_t o_ws = a.take(sizeof(ba::WinState) * W);
  const size_t o_q = a.take(sizeof(double) * 4 * NP), o_t = a.take(sizeof(double) * 3 * NP);
  const size_t o_R = a.take(sizeof(double) * 9 * NP);
  const size_t o_bq = a.take(sizeof(double) * 4 * NP), o_bt = a.take(sizeof(double) * 3 * NP);
  const size_t o_fi = a.take(sizeof(int) * NP), o_pact = a.take(sizeof(int) * NP);
  const size_t o_nfb = a.take(sizeof(int) * (W + 1));
  const size_t o_pof = a.take(sizeof(int) * (NF + 1)), o_sys = a.take(sizeof(int) * (NF + 1));
  // Hpp and bp are one contiguous block of NF * 27 doubles (a single all-reduce in global mode)
  const size_t o_hpp = a.take(sizeof(double) * 27 * (NF + 1));
  const size_t o_xp = a.take(sizeof(double) * 6 * (NF + 1));
  const size_t hs_count = 42 * (size_t)W * (Pmax > 0 ? Pmax : 1) + 8;
  const size_t o_hsp = a.take(sizeof(double) * hs_count);
  const bool global = c->global_mode;
  // rank-local write buffers of the global mode (alias the read buffers otherwise)
  const size_t o_hpp_w = global ? a.take(sizeof(double) * 27 * (NF + 1)) : o_hpp;
  const size_t o_hsp_w = global ? a.take(sizeof(double) * hs_count) : o_hsp;
  const size_t o_pact_w = global ? a.take(sizeof(int) * NP) : o_pact;
  const size_t o_gs = a.take(sizeof(double) * 16);
  const size_t o_part = a.take(sizeof(double) * 4 * (size_t)W * C);
  const size_t o_pbeg = a.take(sizeof(int) * (size_t)W * (2 * Pmax + 1));
  const size_t o_pairs = a.take(sizeof(int2) * (size_t)(n_pairs + 1));
  // landmark-driven pair-list builder of large windows: a second entry buffer and per-pair cursors
  const bool big_pairs = NFmax > ba::PAIRS_LM_MIN_NF;
  const size_t o_ne_list = a.take(sizeof(int) * (size_t)W * (Pmax > 0 ? Pmax : 1));
  const size_t o_ne_flag = a.take(sizeof(int) * (size_t)W * (Pmax > 0 ? Pmax : 1));
  const size_t o_n_ne = a.take(sizeof(int) * W);
  const size_t o_diag_pos = a.take(sizeof(int) * (NF + 1));
  const size_t o_blk = a.take(sizeof(int2) * (size_t)W * (Pmax > 0 ? Pmax : 1));
  const size_t o_pairs_tmp = big_pairs ? a.take(sizeof(int2) * (size_t)(n_pairs + 1)) : 0;
  const size_t o_cursor = big_pairs ? a.take(sizeof(int) * (size_t)W * 2 * (Pmax > 0 ? Pmax : 1)) : 0;
  const size_t o_pbase = a.take(sizeof(long long) * (W + 1));
  const size_t o_nact = a.take(sizeof(int));
  CU_TRY(c, c->batch_buf.reserve(a.off));
  char* base = c->batch_buf.as<char>();
  CU_TRY(c, cudaMemcpyAsync(base + o_nfb, c->l_nf_begin.data(), sizeof(int) * (W + 1), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(base + o_pbase, c->l_pair_base.data(), sizeof(long long) * (W + 1), cudaMemcpyHostToDevice,
                            c->stream));
  CU_TRY(c, cudaMemsetAsync(base + o_part, 0, sizeof(double) * 4 * (size_t)W * C, c->stream));
  CU_TRY(c, cudaMemsetAsync(base + o_gs, 0, sizeof(double) * 16, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  ba::BatchDev& b = c->bd;
  b.ws = (ba::WinState*)(base + o_ws);
  b.P_q = (double*)(base + o_q);
  b.P_t = (double*)(base + o_t);
  b.P_R = (double*)(base + o_R);
  b.P_bq = (double*)(base + o_bq);
  b.P_bt = (double*)(base + o_bt);
  b.free_idx = (int*)(base + o_fi);
  b.pact = (int*)(base + o_pact);
  b.nf_begin = (int*)(base + o_nfb);
  b.pose_of = (int*)(base + o_pof);
  b.sys_idx = (int*)(base + o_sys);
  b.Hpp = (double*)(base + o_hpp);
  b.bp = b.Hpp + (size_t)21 * NF;
  b.global = global ? 1 : 0;
  b.Hpp_w = (double*)(base + o_hpp_w);
  b.bp_w = b.Hpp_w + (size_t)21 * NF;
  b.hs_part_w = (double*)(base + o_hsp_w);
  b.pact_w = (int*)(base + o_pact_w);
  b.gs_w = (double*)(base + o_gs);
  b.gs = b.gs_w + 8;
  b.xp = (double*)(base + o_xp);
  b.hs_part = (double*)(base + o_hsp);
  b.part = (double*)(base + o_part);
  b.pair_beg = (int*)(base + o_pbeg);
  b.pairs = (int2*)(base + o_pairs);
  b.ne_list = (int*)(base + o_ne_list);
  b.ne_flag = (int*)(base + o_ne_flag);
  b.n_ne = (int*)(base + o_n_ne);
  b.diag_pos = (int*)(base + o_diag_pos);
  b.blk = (int2*)(base + o_blk);
  b.pairs_tmp = big_pairs ? (int2*)(base + o_pairs_tmp) : nullptr;
  b.pair_cursor = big_pairs ? (int*)(base + o_cursor) : nullptr;
  b.pair_base = (const long long*)(base + o_pbase);
  b.n_active = (int*)(base + o_nact);
  b.Cp = Cp;
  b.Cl = Cl;
  b.C = C;
  b.Pmax = Pmax > 0 ? Pmax : 1;
  b.NFmax = NFmax;
  c->batch_ready = true;
  return RSPL_BA_OK;
}

// Tables of the tiled Schur path (local_tiled.cuh): tile size, tile counts, device arrays. maxdeg = largest landmark
// degree per kind (read back from the setup kernels); smem_tile = dynamic shared memory of kt_schur_tile per kind.
int tiles_prepare(RsplBaContext* c, const int maxdeg[2], size_t smem_tile[2]) {
  const int W = c->l_n_windows;
  // shared-memory cost of a window's landmarks of one kind: A per landmark, B per edge; the staged pair entries add
  // 2 (maxdeg + 1) bytes per edge to the compile-time per-edge cost
  const int cost_a[2] = {ba::TileCost<0>::A, ba::TileCost<1>::A};
  const int cost_b[2] = {ba::TileCost<0>::B + 2 * (maxdeg[0] + 1), ba::TileCost<1>::B + 2 * (maxdeg[1] + 1)};
  const int Pmax = c->bd.Pmax;
  const long long fixed = (long long)ba::tile_fixed_bytes(Pmax, c->l_max_poses, c->ld.n_cameras);
  auto cost = [&](int k, int w) { return c->l_cost_nl[k][w] * cost_a[k] + c->l_cost[k][w] * cost_b[k]; };
  long long total = 0;
  for (int k = 0; k < 2; ++k)
    for (int w = 0; w < W; ++w) total += cost(k, w);
  // Two CTAs per SM. The tile size is the same for every batch: the tiles of a window (and with them the order of
  // its sums) must not depend on what else is in the batch.
  long long Q = 100 << 10;
  if (const char* e = getenv("RSPL_BA_TILE_Q")) Q = atoll(e) > 4096 ? atoll(e) : Q;
  (void)total;
  for (int k = 0; k < 2; ++k) {
    const long long room = (long long)c->smem_optin - cost_a[k] - (long long)cost_b[k] * maxdeg[k] - fixed;
    if (room < 4096) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: landmark degree %d too large for the tiled Schur path", maxdeg[k]);
    if (Q > room) Q = room;
  }
  int T[2] = {0, 0};
  for (int k = 0; k < 2; ++k)
    for (int w = 0; w < W; ++w) {
      const int nt = (int)((cost(k, w) + Q - 1) / Q);
      if (nt > T[k]) T[k] = nt;
    }
  const int Tcap = T[0] > T[1] ? (T[0] > 0 ? T[0] : 1) : (T[1] > 0 ? T[1] : 1);
  Arena a;
  const size_t o_tlm = a.take(sizeof(int) * (size_t)W * 2 * (Tcap + 1));
  const size_t o_tso = a.take(sizeof(int) * (size_t)W * 2 * Tcap * (Pmax + 1));
  const size_t o_tb = a.take(sizeof(int) * (size_t)W * 2 * Tcap);
  const size_t o_desc = a.take(sizeof(int4) * (size_t)W * 2 * Tcap * 2);
  const size_t o_tent = a.take(sizeof(ushort2) * (size_t)(c->l_pair_base[W] + 8LL * Tcap * W + 8));
  const size_t o_nt = a.take(sizeof(int) * (size_t)W * 2);
  const size_t o_tpb = a.take(sizeof(int) * (size_t)W * 2 * (Tcap + 1) * Pmax);
  const size_t o_ord = a.take(sizeof(int) * (size_t)W * Pmax);
  const size_t o_hs = a.take(sizeof(double) * (size_t)W * (T[0] + T[1] > 0 ? T[0] + T[1] : 1) * Pmax * 42);
  const size_t o_bR = a.take(sizeof(double) * 9 * (size_t)c->l_np);
  CU_TRY(c, c->tile_buf.reserve(a.off));
  char* base = c->tile_buf.as<char>();
  ba::TileDev& td = c->td;
  td.Q = (int)Q;
  td.Tcap = Tcap;
  td.Tp = T[0];
  td.Tl = T[1];
  td.tile_lm = (int*)(base + o_tlm);
  td.ntile = (int*)(base + o_nt);
  td.tpb = (int*)(base + o_tpb);
  td.tso = (int*)(base + o_tso);
  td.tent_base = (int*)(base + o_tb);
  td.desc = (int4*)(base + o_desc);
  td.tent = (ushort2*)(base + o_tent);
  td.cost_b[0] = cost_b[0];
  td.cost_b[1] = cost_b[1];
  td.order = (int*)(base + o_ord);
  td.hs_tile = (double*)(base + o_hs);
  td.P_bR = (double*)(base + o_bR);
  for (int k = 0; k < 2; ++k) smem_tile[k] = (size_t)(Q + cost_a[k] + (long long)cost_b[k] * maxdeg[k] + fixed);
  return RSPL_BA_OK;
}

// grids of the setup kernels (shared by the graph path and the host-driven path)
void enqueue_local_setup(RsplBaContext* c, cudaStream_t s) {
  const int W = c->l_n_windows;
  cudaMemsetAsync(c->ld.err, 0, sizeof(int) * 4, s);
  for (int k = 0; k < 2; ++k)
    if (c->ld.k[k].n_lm > 0) {
      cudaMemsetAsync(c->ld.k[k].slot, ba::SLOT_NONE, (size_t)c->ld.k[k].n_lm * c->ld.slot_stride, s);
      cudaMemsetAsync(c->ld.k[k].cursor, 0, sizeof(int) * (size_t)c->ld.k[k].n_lm, s);
    }
  const int T = ba::LOCAL_THREADS;
  const int max_lm = c->l_max_pts > c->l_max_lns ? c->l_max_pts : c->l_max_lns;
  const dim3 g_e((c->l_max_edges + T - 1) / T > 0 ? (c->l_max_edges + T - 1) / T : 1, W, 2);
  const dim3 g_l((max_lm + T - 1) / T > 0 ? (max_lm + T - 1) / T : 1, W, 2);
  ba::setup_poses<<<W, T, 0, s>>>(c->ld);
  ba::setup_edges<0><<<g_e, T, 0, s>>>(c->ld);
  if (c->ld.order_hist) { // large windows: the counting sort spread over 1024-landmark chunks (same result)
    const dim3 g_c(c->ld.order_chunks, W, 2);
    ba::setup_order_hist<<<g_c, T, 0, s>>>(c->ld);
    ba::setup_order_offsets<<<dim3(W, 2), 256, 0, s>>>(c->ld);
    ba::setup_order_place<<<g_c, T, 0, s>>>(c->ld);
    c->launches += 2;
  } else {
    ba::setup_order<<<dim3(W, 2), T, 0, s>>>(c->ld);
  }
  ba::setup_scan<<<dim3(W, 2), 1024, 0, s>>>(c->ld);
  ba::setup_edges<1><<<g_e, T, 0, s>>>(c->ld);
  ba::setup_landmarks<<<g_l, T, 0, s>>>(c->ld);
  ba::setup_gather<<<g_e, T, 0, s>>>(c->ld);
  const dim3 g_pl(c->l_max_poses, W, 2);
  if (c->ld.pose_cursor[0]) { // large windows: count / scatter per pose with integer atomics, then sort every list (same result)
    for (int k = 0; k < 2; ++k) {
      cudaMemsetAsync(c->ld.k[k].pbeg, 0, sizeof(int) * ((size_t)c->l_np + 1), s);
      cudaMemsetAsync(c->ld.pose_cursor[k], 0, sizeof(int) * (size_t)c->l_np, s);
    }
    ba::setup_pose_scatter<0><<<g_e, T, 0, s>>>(c->ld);
    ba::setup_pose_scan<<<(2 * W + 127) / 128, 128, 0, s>>>(c->ld);
    ba::setup_pose_scatter<1><<<g_e, T, 0, s>>>(c->ld);
    ba::setup_pose_sort<<<g_pl, T, 0, s>>>(c->ld);
    c->launches += 1;
  } else {
    ba::setup_pose_lists<0><<<g_pl, T, 0, s>>>(c->ld);
    ba::setup_pose_scan<<<(2 * W + 127) / 128, 128, 0, s>>>(c->ld);
    ba::setup_pose_lists<1><<<g_pl, T, 0, s>>>(c->ld);
  }
  c->launches += 10;
}

// ---- the whole LocalmapOptimization schedule as ONE CUDA graph ----------------------------------
// setup kernels -> pair lists -> tile tables -> [pass 1: WHILE(any window iterating) { super-step }] -> flagging ->
// [pass 2: WHILE { super-step }] -> final flags -> write-back. The loop conditions are set on the device (kt_cond,
// cudaGraphSetConditional), so the host neither polls nor synchronises; the instantiated graph is cached in the
// context under the byte image of every kernel argument block, so a call with the same shapes and buffers is a
// single cudaGraphLaunch. Used whenever the reduced systems fit shared memory (no HBM-resident solve, no NCCL in the loop)
// and profiling is off.
constexpr int POSE_BRANCH_MAX_W = 32; // batches up to this many windows run kb_pose_blocks on its own graph branch

struct LocalGraphKey {
  ba::LocalDev d;
  ba::BatchDev b;
  ba::TileDev td;
  ba::LocalOpt lo;
  int W, max_pts, max_lns, max_edges, max_poses;
  unsigned long long smem_solve, smem_tile[2];
};

int capture_local_graph(RsplBaContext* c, const ba::LocalOpt& lo, size_t smem_solve, const size_t smem_tile[2],
                        RsplBaContext::LocalGraph& out) {
  const ba::LocalDev& d = c->ld;
  const ba::BatchDev& b = c->bd;
  const ba::TileDev& td = c->td;
  const int W = c->l_n_windows;
  cudaStream_t s = c->stream;
  if (!c->s_body) CU_TRY(c, cudaStreamCreateWithFlags(&c->s_body, cudaStreamNonBlocking));
  if (!c->s_aux) {
    CU_TRY(c, cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
    for (int i = 0; i < 6; ++i) CU_TRY(c, cudaEventCreateWithFlags(&c->fork_ev[i], cudaEventDisableTiming));
  }
  if (!c->s_aux2) {
    CU_TRY(c, cudaStreamCreateWithFlags(&c->s_aux2, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) CU_TRY(c, cudaEventCreateWithFlags(&c->fork_ev2[i], cudaEventDisableTiming));
  }
  const dim3 g_pt(b.Cp, W), g_ln(b.Cl, W), g_lm(b.C, W), g_pose(b.NFmax > 0 ? b.NFmax : 1, W), g_pair((b.Pmax + ba::BW - 1) / ba::BW, W),
      g_win((W + 127) / 128), g_winw((W + 3) / 4), g_tp(td.Tp > 0 ? td.Tp : 1, W), g_tl(td.Tl > 0 ? td.Tl : 1, W);
  const int edge_chunks = (c->l_max_edges + 256 * 8 - 1) / (256 * 8) > 0 ? (c->l_max_edges + 256 * 8 - 1) / (256 * 8) : 1;
  const int lm_chunks = ((c->l_max_pts > c->l_max_lns ? c->l_max_pts : c->l_max_lns) + 256 * 8 - 1) / (256 * 8) > 0
                            ? ((c->l_max_pts > c->l_max_lns ? c->l_max_pts : c->l_max_lns) + 256 * 8 - 1) / (256 * 8) : 1;
  const bool fork_lines = b.Cp && b.Cl;
  int n_launch = 0;
#define GK(stream, kern, grid, block, shm, ...)      \
  do {                                               \
    kern<<<grid, block, shm, stream>>>(__VA_ARGS__); \
    ++n_launch;                                      \
  } while (0)
  // one super-step on (sm, sl): the line kernels of a phase run beside the point kernels (fork / join events
  // become graph dependencies)
  auto super_step = [&](cudaStream_t sm, cudaStream_t sl, cudaGraphConditionalHandle cond) {
    auto fork = [&](int k) {
      if (!fork_lines) return;
      cudaEventRecord(c->fork_ev[2 * k], sm);
      cudaStreamWaitEvent(sl, c->fork_ev[2 * k], 0);
    };
    auto join = [&](int k) {
      if (!fork_lines) return;
      cudaEventRecord(c->fork_ev[2 * k + 1], sl);
      cudaStreamWaitEvent(sm, c->fork_ev[2 * k + 1], 0);
    };
    cudaStream_t sline = fork_lines ? sl : sm;
    fork(0);
    // (the pose blocks are independent of both linearisations; small batches are latency-bound, so they get a third
    // branch of the graph there instead of following the line kernel)
    const bool fork_pose = fork_lines && W <= POSE_BRANCH_MAX_W;
    if (fork_pose) {
      cudaEventRecord(c->fork_ev2[0], sm);
      cudaStreamWaitEvent(c->s_aux2, c->fork_ev2[0], 0);
    }
    if (b.Cp) GK(sm, ba::kb_linearize<0>, g_pt, ba::BT, 0, d, b, lo);
    if (b.Cl) GK(sline, ba::kb_linearize<1>, g_ln, ba::BT, 0, d, b, lo);
    GK(fork_pose ? c->s_aux2 : sline, ba::kb_pose_blocks, g_pose, ba::BT, 0, d, b, lo);
    if (fork_pose) {
      cudaEventRecord(c->fork_ev2[1], c->s_aux2);
      cudaStreamWaitEvent(sm, c->fork_ev2[1], 0);
    }
    join(0);
    GK(sm, ba::kb_begin_trial, g_winw, 128, 0, d, b);
    fork(1);
    if (b.Cp && td.Tp) GK(sm, ba::kt_schur_tile<0>, g_tp, ba::TILE_THREADS, smem_tile[0], d, b, lo, td);
    if (b.Cl && td.Tl) GK(sline, ba::kt_schur_tile<1>, g_tl, ba::TILE_THREADS, smem_tile[1], d, b, lo, td);
    join(1);
    GK(sm, ba::kt_tile_sum, dim3((b.Pmax * 42 * ba::TILE_SUM_LANES + 255) / 256, W), 256, 0, d, b, td);
    GK(sm, ba::kb_solve<true>, W, 256, smem_solve, d, b, td);
    fork(2);
    if (b.Cp) GK(sm, ba::kt_backsub_rc<0>, g_pt, ba::BT, 0, d, b, lo, td);
    if (b.Cl) GK(sline, ba::kt_backsub_rc<1>, g_ln, ba::BT, 0, d, b, lo, td);
    join(2);
    GK(sm, ba::kb_decide<true>, g_winw, 128, 0, d, b, cond);
    if (b.C) GK(sm, ba::kb_restore, g_lm, ba::BT, 0, d, b);
  };
  cudaGraph_t graph = nullptr;
  bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  int launches_step = 0;
  if (ok) {
    const int64_t l0 = c->launches;
    enqueue_local_setup(c, s);
    n_launch += (int)(c->launches - l0);
    c->launches = l0;
    cudaMemsetAsync(b.n_active, 0, sizeof(int), s);
    GK(s, ba::kb_init, W, ba::BT, 0, d, b, lo);
    GK(s, ba::kb_pairs<0>, g_pair, ba::BT, 0, d, b);
    GK(s, ba::kb_pairs_scan, g_win, 128, 0, d, b);
    GK(s, ba::kb_pairs<1>, g_pair, ba::BT, 0, d, b);
    GK(s, ba::kb_pairs_flag, dim3((b.Pmax + 255) / 256, W), 256, 0, d, b);
    GK(s, ba::kb_pairs_compact, W, 1024, 0, d, b);
    GK(s, ba::kt_tiles_lm, dim3((td.Tcap + 1 + 127) / 128, W, 2), 128, 0, d, td);
    GK(s, ba::kt_tiles_pairs, dim3(((size_t)b.Pmax * (td.Tcap + 1) + 255) / 256, W, 2), 256, 0, d, b, td);
    GK(s, ba::kt_tiles_scan, dim3(td.Tcap, W, 2), 32, 0, d, b, td);
    GK(s, ba::kt_tiles_base, (W + 127) / 128, 128, 0, d, b, td);
    GK(s, ba::kt_tiles_fill, dim3(((size_t)b.Pmax * td.Tcap + 255) / 256, W, 2), 256, 0, d, b, td);
    GK(s, ba::kt_tiles_desc, dim3((td.Tcap + 127) / 128, W, 2), 128, 0, d, b, td);
    GK(s, ba::kt_order, (W + 3) / 4, 128, 0, d, b, td);
    for (int pass = 0; pass < 2 && ok; ++pass) {
      cudaMemsetAsync(b.pact_w, 0, sizeof(int) * (size_t)c->l_np, s);
      if (b.C) GK(s, ba::kb_mark_active, g_lm, ba::BT, 0, d, b);
      GK(s, ba::kb_begin_pass, g_win, 128, 0, d, b, lo, pass);
      GK(s, ba::kb_pair_blocks, dim3((b.Pmax + 255) / 256, W), 256, 0, d, b);
      // WHILE node: created by hand behind the nodes captured so far, its body captured from s_body
      cudaStreamCaptureStatus st;
      cudaGraph_t g = nullptr;
      const cudaGraphNode_t* deps = nullptr;
      size_t ndeps = 0;
      cudaGraphConditionalHandle h;
      ok = cudaStreamGetCaptureInfo(s, &st, nullptr, &g, &deps, &ndeps) == cudaSuccess && g &&
           cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault) == cudaSuccess;
      if (!ok) break;
      GK(s, ba::kt_cond, 1, 256, 0, d, b, h, 0);
      ok = cudaStreamGetCaptureInfo(s, &st, nullptr, &g, &deps, &ndeps) == cudaSuccess;
      if (!ok) break;
      cudaGraphNodeParams np = {};
      np.type = cudaGraphNodeTypeConditional;
      np.conditional.handle = h;
      np.conditional.type = cudaGraphCondTypeWhile;
      np.conditional.size = 1;
      cudaGraphNode_t node;
      ok = cudaGraphAddNode(&node, g, deps, ndeps, &np) == cudaSuccess;
      if (!ok) break;
      cudaGraph_t body = np.conditional.phGraph_out[0];
      ok = cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies) == cudaSuccess &&
           cudaStreamBeginCaptureToGraph(c->s_body, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (!ok) break;
      const int before = n_launch;
      super_step(c->s_body, c->s_aux, h); // (kb_decide sets the loop condition and counts the super-step)
      launches_step = n_launch - before;
      n_launch = before;
      ok = cudaStreamEndCapture(c->s_body, nullptr) == cudaSuccess;
      if (!ok) break;
      if (pass == 0) GK(s, ba::kb_flag<false>, dim3(W, edge_chunks), 256, 0, d, b, lo);
    }
    if (ok) {
      GK(s, ba::kb_flag<true>, dim3(W, edge_chunks), 256, 0, d, b, lo);
      GK(s, ba::kb_writeback, dim3(W, lm_chunks), 256, 0, d, b);
    }
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    ok = ok && e == cudaSuccess && graph;
  }
#undef GK
  if (!ok) {
    if (graph) cudaGraphDestroy(graph);
    // leave no stream in capture mode behind
    cudaStreamCaptureStatus st;
    if (cudaStreamIsCapturing(c->s_body, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
      cudaGraph_t dump = nullptr;
      cudaStreamEndCapture(c->s_body, &dump);
    }
    if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
      cudaGraph_t dump = nullptr;
      cudaStreamEndCapture(s, &dump);
      if (dump) cudaGraphDestroy(dump);
    }
    cudaGetLastError();
    return RSPL_BA_ERR_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess || !exec) {
    cudaGetLastError();
    return RSPL_BA_ERR_CUDA;
  }
  out.exec = exec;
  out.launches_fixed = n_launch;
  out.launches_step = launches_step;
  return RSPL_BA_OK;
}

// returns RSPL_BA_OK when the solve was enqueued as a graph, RSPL_BA_ERR_STATE when this batch does not qualify
// (the caller then takes the host-driven path), another code on failure
int local_solve_graph(RsplBaContext* c, const ba::LocalOpt& lo) {
  if (c->prof || c->global_mode) return RSPL_BA_ERR_STATE;
  if (const char* env = getenv("RSPL_BA_GRAPH"))
    if (!strcmp(env, "off")) return RSPL_BA_ERR_STATE;
  if (const char* env = getenv("RSPL_BA_SCHUR"))
    if (!strcmp(env, "legacy")) return RSPL_BA_ERR_STATE;
  int rc = batched_prepare(c);
  if (rc != RSPL_BA_OK) return rc;
  const ba::BatchDev& b = c->bd;
  const int W = c->l_n_windows;
  if (W > 65535) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: more than 65535 windows in one call");
  const int n_max = 6 * b.NFmax;
  const size_t smem_solve = sizeof(double) * ((size_t)n_max * n_max + 3 * n_max + b.NFmax + 8);
  if (smem_solve > c->smem_optin || b.pairs_tmp) return RSPL_BA_ERR_STATE; // dense reduced system: host-driven path
  // one constraint per (pose, landmark) pair => a landmark's degree is bounded by the poses of its window
  const int deg_bound = c->l_max_poses < 254 ? c->l_max_poses : 254;
  const int maxdeg[2] = {deg_bound, deg_bound};
  size_t smem_tile[2];
  rc = tiles_prepare(c, maxdeg, smem_tile);
  if (rc != RSPL_BA_OK) return rc;
  CU_TRY(c, cudaFuncSetAttribute(ba::kb_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
  CU_TRY(c, cudaFuncSetAttribute(ba::kt_schur_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile[0]));
  CU_TRY(c, cudaFuncSetAttribute(ba::kt_schur_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile[1]));
  LocalGraphKey key;
  memset(&key, 0, sizeof(key));
  key.d = c->ld;
  key.b = c->bd;
  key.td = c->td;
  key.lo = lo;
  key.W = W;
  key.max_pts = c->l_max_pts;
  key.max_lns = c->l_max_lns;
  key.max_edges = c->l_max_edges;
  key.max_poses = c->l_max_poses;
  key.smem_solve = smem_solve;
  key.smem_tile[0] = smem_tile[0];
  key.smem_tile[1] = smem_tile[1];
  RsplBaContext::LocalGraph* hit = nullptr;
  for (auto& g : c->graph_cache)
    if (g.exec && g.key.size() == sizeof(key) && !memcmp(g.key.data(), &key, sizeof(key))) hit = &g;
  if (!hit) {
    RsplBaContext::LocalGraph fresh;
    rc = capture_local_graph(c, lo, smem_solve, smem_tile, fresh);
    if (rc != RSPL_BA_OK) return fail(c, rc, "local batch: CUDA graph capture of the LM schedule failed");
    fresh.key.assign((const unsigned char*)&key, (const unsigned char*)&key + sizeof(key));
    if (c->graph_cache.size() < 8) {
      c->graph_cache.push_back(fresh);
      hit = &c->graph_cache.back();
    } else { // replace the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < c->graph_cache.size(); ++i)
        if (c->graph_cache[i].stamp < c->graph_cache[lru].stamp) lru = i;
      if (c->graph_cache[lru].exec) cudaGraphExecDestroy(c->graph_cache[lru].exec);
      c->graph_cache[lru] = fresh;
      hit = &c->graph_cache[lru];
    }
  }
  hit->stamp = ++c->graph_stamp;
  CU_TRY(c, cudaGraphLaunch(hit->exec, c->stream));
  c->launches += hit->launches_fixed;
  c->l_graph_launches_step = hit->launches_step; // (x super-steps: added at download, from the device counter)
  c->l_last_path = 4;
  c->l_super_steps = 0;
  return RSPL_BA_OK;
}

// The batched LocalmapOptimization: fixed kernel sequence per super-step, per-window LM state
// machines on the device, host polls the number of still-iterating windows.
int local_solve_batched(RsplBaContext* c, const ba::LocalOpt& lo) {
  int rc = batched_prepare(c);
  if (rc != RSPL_BA_OK) return rc;
  c->l_last_path = 2;
  c->l_super_steps = 0;
  const ba::LocalDev& d = c->ld;
  const ba::BatchDev& b = c->bd;
  const int W = c->l_n_windows;
  cudaStream_t s = c->stream;
  if (W > 65535) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: more than 65535 windows in one call");
  // (chunk | pose | pair) index fastest, window on grid.y: the CTAs of one window run together
  const dim3 g_pt(b.Cp, W), g_ln(b.Cl, W), g_lm(b.C, W), g_pose(b.NFmax > 0 ? b.NFmax : 1, W), g_pair((b.Pmax + ba::BW - 1) / ba::BW, W),
      g_pair1(b.Pmax > 0 ? b.Pmax : 1, W), g_win((W + 127) / 128), g_winw((W + 3) / 4);
  const int n_max = 6 * b.NFmax;
  const size_t smem_solve = sizeof(double) * ((size_t)n_max * n_max + 3 * n_max + b.NFmax + 8);
  // reduced systems beyond shared memory: cyclic reduction or dense Cholesky in HBM (dense_solver.inl)
  // (global BA always takes this route: its collectives sit between the kernels of a super-step)
  const bool global = c->global_mode;
  if (global && W != 1) return fail(c, RSPL_BA_ERR_INVALID, "global BA: upload exactly one window (this rank's shard)");
  const bool dense = global || smem_solve > c->smem_optin;
  DenseLayout dl;
  std::vector<int> n_sys_host(W, 0);
  if (dense) {
    c->l_last_path = 3; // (buffers are set up after the pair lists: their band decides the factorisation)
  } else {
    CU_TRY(c, cudaFuncSetAttribute(ba::kb_solve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
    CU_TRY(c, cudaFuncSetAttribute(ba::kb_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
  }
#define LAUNCH(cls, kern, grid, block, shm, ...)       \
  do {                                                 \
    ProfScope ps_(c, cls);                             \
    kern<<<grid, block, shm, s>>>(__VA_ARGS__);        \
    c->launches++;                                     \
  } while (0)
  LAUNCH(PC_PAIRS, ba::kb_init, W, ba::BT, 0, d, b, lo);
  if (b.pairs_tmp) { // large windows: lists built from the landmarks, then ordered
    const size_t n_cnt = (size_t)W * (2 * b.Pmax + 1);
    CU_TRY(c, cudaMemsetAsync(b.pair_beg, 0, sizeof(int) * n_cnt, s));
    CU_TRY(c, cudaMemsetAsync(b.pair_cursor, 0, sizeof(int) * (size_t)W * 2 * b.Pmax, s));
    LAUNCH(PC_PAIRS, ba::kb_pairs_lm<0>, g_lm, ba::BT, 0, d, b);
    {
      // device-wide scan of the 2P counts (pairs_tmp is free until the sort: scratch for the chunk sums)
      const int nch = (int)((2 * (long long)b.Pmax + ba::SCAN_CHUNK - 1) / ba::SCAN_CHUNK);
      if (nch <= ba::SCAN_CHUNK && (size_t)W * nch * sizeof(int) <= sizeof(int2) * (size_t)(c->l_pair_base[W] + 1)) {
        int* scratch = (int*)b.pairs_tmp;
        LAUNCH(PC_PAIRS, ba::kb_big_scan_sums<0>, dim3(nch, W), 1024, 0, b, scratch, nch);
        LAUNCH(PC_PAIRS, ba::kb_big_scan_offsets<0>, W, 1024, 0, d, b, scratch, nch);
        LAUNCH(PC_PAIRS, ba::kb_big_scan_apply<0>, dim3(nch, W), 1024, 0, b, scratch, nch);
      } else {
        LAUNCH(PC_PAIRS, ba::kb_pairs_scan_cta, W, 1024, 0, d, b);
      }
    }
    LAUNCH(PC_PAIRS, ba::kb_pairs_lm<1>, g_lm, ba::BT, 0, d, b); // (lists in arbitrary order: sorted below, once the compact list exists)
  } else {
    LAUNCH(PC_PAIRS, ba::kb_pairs<0>, g_pair, ba::BT, 0, d, b);
    LAUNCH(PC_PAIRS, ba::kb_pairs_scan, g_win, 128, 0, d, b);
    LAUNCH(PC_PAIRS, ba::kb_pairs<1>, g_pair, ba::BT, 0, d, b);
  }
  CU_TRY(c, cudaGetLastError());
  // compact list of the pairs that can be non-zero (union over the ranks in global mode)
  LAUNCH(PC_PAIRS, ba::kb_pairs_flag, dim3((b.Pmax + 255) / 256, W), 256, 0, d, b);
  if (global) {
    rc = comm_all_reduce(c, b.ne_flag, b.ne_flag, (size_t)b.Pmax, kNcclInt32, kNcclMax);
    if (rc != RSPL_BA_OK) return rc;
  }
  {
    const int nch = (int)(((long long)b.Pmax + ba::SCAN_CHUNK - 1) / ba::SCAN_CHUNK);
    if (b.pairs_tmp && nch <= ba::SCAN_CHUNK && (size_t)W * nch * sizeof(int) <= sizeof(int2) * (size_t)(c->l_pair_base[W] + 1)) {
      int* scratch = (int*)b.pairs_tmp; // (free again: the sort below is the next user)
      LAUNCH(PC_PAIRS, ba::kb_big_scan_sums<1>, dim3(nch, W), 1024, 0, b, scratch, nch);
      LAUNCH(PC_PAIRS, ba::kb_big_scan_offsets<1>, W, 1024, 0, d, b, scratch, nch);
      LAUNCH(PC_PAIRS, ba::kb_big_scan_apply<1>, dim3(nch, W), 1024, 0, b, scratch, nch);
    } else {
      LAUNCH(PC_PAIRS, ba::kb_pairs_compact, W, 1024, 0, d, b);
    }
  }
  std::vector<int> n_ne_host(W, 0);
  int setup_flags[4] = {0, 0, 0, 0}; // error flag, max degree of points / lines
  CU_TRY(c, cudaMemcpyAsync(n_ne_host.data(), b.n_ne, sizeof(int) * W, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaMemcpyAsync(setup_flags, d.err, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaStreamSynchronize(s));
  // g2o accepts what the slot table cannot hold; say so before iterating on a truncated table
  if (setup_flags[0] & ba::LOCAL_ERR_DUP_EDGE)
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: two constraints join the same (pose, landmark) pair");
  if (setup_flags[0] & ba::LOCAL_ERR_DEGREE)
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: a landmark has more than 254 observations");
  int n_ne_max = 1;
  for (int w = 0; w < W; ++w) n_ne_max = n_ne_host[w] > n_ne_max ? n_ne_host[w] : n_ne_max;
  if (b.pairs_tmp) {
    LAUNCH(PC_PAIRS, ba::kb_pairs_sort, dim3((n_ne_max + ba::BW - 1) / ba::BW, W), ba::BT, 0, d, b);
    LAUNCH(PC_PAIRS, ba::kb_pairs_sort_long, dim3(n_ne_max, W), 256, 0, d, b);
  }
  const dim3 g_ne(n_ne_max, W);
  if (dense) {
    // which pose pairs share landmarks: a banded pattern allows the block-tridiagonal factorisation
    std::vector<int> band(W, 0);
    int* d_band = (int*)b.part; // scratch: `part` is rewritten by the first linearisation
    CU_TRY(c, cudaMemsetAsync(d_band, 0, sizeof(int) * W, s));
    LAUNCH(PC_PAIRS, ba::kb_pair_band, dim3((b.Pmax + 255) / 256, W), 256, 0, d, b, d_band);
    if (global) {
      rc = comm_all_reduce(c, d_band, d_band, (size_t)W, kNcclInt32, kNcclMax);
      if (rc != RSPL_BA_OK) return rc;
    }
    CU_TRY(c, cudaMemcpyAsync(band.data(), d_band, sizeof(int) * W, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaMemsetAsync(d_band, 0, sizeof(int) * W, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    rc = dense_prepare(c, dl, band);
    if (rc != RSPL_BA_OK) return rc;
    c->bd.dense_H = dl.H;
    c->bd.dense_b = dl.b;
    c->bd.dense_info = dl.info;
    c->bd.dense_off = dl.d_off;
  }
  // Schur elimination: per-tile fused kernel (local_tiled.cuh) unless the reduced system goes to HBM (dense / global)
  bool tiled = !dense;
  if (const char* env = getenv("RSPL_BA_SCHUR"))
    if (!strcmp(env, "legacy")) tiled = false;
  const ba::TileDev& td = c->td;
  size_t smem_tile[2] = {0, 0};
  if (tiled) {
    rc = tiles_prepare(c, setup_flags + 1, smem_tile);
    if (rc != RSPL_BA_OK) return rc;
    c->l_last_path = 4;
    LAUNCH(PC_PAIRS, ba::kt_tiles_lm, dim3((td.Tcap + 1 + 127) / 128, W, 2), 128, 0, d, td);
    LAUNCH(PC_PAIRS, ba::kt_tiles_pairs, dim3(((size_t)b.Pmax * (td.Tcap + 1) + 255) / 256, W, 2), 256, 0, d, b, td);
    LAUNCH(PC_PAIRS, ba::kt_tiles_scan, dim3(td.Tcap, W, 2), 32, 0, d, b, td);
    LAUNCH(PC_PAIRS, ba::kt_tiles_base, (W + 127) / 128, 128, 0, d, b, td);
    LAUNCH(PC_PAIRS, ba::kt_tiles_fill, dim3(((size_t)b.Pmax * td.Tcap + 255) / 256, W, 2), 256, 0, d, b, td);
    LAUNCH(PC_PAIRS, ba::kt_tiles_desc, dim3((td.Tcap + 127) / 128, W, 2), 128, 0, d, b, td);
    LAUNCH(PC_PAIRS, ba::kt_order, (W + 3) / 4, 128, 0, d, b, td);
    CU_TRY(c, cudaFuncSetAttribute(ba::kt_schur_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile[0]));
    CU_TRY(c, cudaFuncSetAttribute(ba::kt_schur_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile[1]));
  }
  const dim3 g_tp(td.Tp > 0 ? td.Tp : 1, W), g_tl(td.Tl > 0 ? td.Tl : 1, W);
  const int edge_chunks = (c->l_max_edges + 256 * 8 - 1) / (256 * 8) > 0 ? (c->l_max_edges + 256 * 8 - 1) / (256 * 8) : 1;
  const int lm_chunks = ((c->l_max_pts > c->l_max_lns ? c->l_max_pts : c->l_max_lns) + 256 * 8 - 1) / (256 * 8) > 0
                            ? ((c->l_max_pts > c->l_max_lns ? c->l_max_pts : c->l_max_lns) + 256 * 8 - 1) / (256 * 8) : 1;
  // one super-step = a fixed sequence of launches with constant arguments
  int dense_rc = RSPL_BA_OK, coll_rc = RSPL_BA_OK;
  const int NF_all = c->l_nf_begin[W];
  // The line kernels (few landmarks, 190-200 registers, ~10 % resident warps) run on a second stream beside the
  // point kernels of the same phase; fork / join events become graph dependencies under capture. Not when
  // profiling (event pairs are recorded on the main stream) and not on the dense path.
  const bool fork_lines = !c->prof && !dense && b.Cp && b.Cl;
  if (fork_lines && !c->s_aux) {
    CU_TRY(c, cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
    for (int i = 0; i < 6; ++i) CU_TRY(c, cudaEventCreateWithFlags(&c->fork_ev[i], cudaEventDisableTiming));
  }
  cudaStream_t s_ln = fork_lines ? c->s_aux : s;
  auto fork = [&](int k) {
    if (!fork_lines) return;
    cudaEventRecord(c->fork_ev[2 * k], s);
    cudaStreamWaitEvent(s_ln, c->fork_ev[2 * k], 0);
  };
  auto join = [&](int k) {
    if (!fork_lines) return;
    cudaEventRecord(c->fork_ev[2 * k + 1], s_ln);
    cudaStreamWaitEvent(s, c->fork_ev[2 * k + 1], 0);
  };
#define LAUNCH_LN(cls, kern, grid, block, shm, ...)    \
  do {                                                 \
    ProfScope ps_(c, cls);                             \
    kern<<<grid, block, shm, s_ln>>>(__VA_ARGS__);     \
    c->launches++;                                     \
  } while (0)
  auto super_step = [&]() {
    fork(0);
    if (b.Cp) LAUNCH(PC_LINEARIZE, ba::kb_linearize<0>, g_pt, ba::BT, 0, d, b, lo);
    if (b.Cl) LAUNCH_LN(PC_LINEARIZE, ba::kb_linearize<1>, g_ln, ba::BT, 0, d, b, lo);
    LAUNCH(PC_POSE_BLOCKS, ba::kb_pose_blocks, g_pose, ba::BT, 0, d, b, lo);
    join(0);
    if (global) { // pose blocks and the linearisation scalars of every rank's landmarks
      LAUNCH(PC_CONTROL, ba::kb_global_sums, 1, 256, 0, b, 0);
      ProfScope ps_(c, PC_COLLECTIVE);
      if (coll_rc == RSPL_BA_OK) coll_rc = comm_all_reduce(c, b.Hpp_w, b.Hpp, (size_t)27 * NF_all, kNcclFloat64, kNcclSum);
      if (coll_rc == RSPL_BA_OK) coll_rc = comm_all_reduce(c, b.gs_w, b.gs, 2, kNcclFloat64, kNcclSum);
      if (coll_rc == RSPL_BA_OK) coll_rc = comm_all_reduce(c, b.gs_w + 2, b.gs + 2, 1, kNcclFloat64, kNcclMax);
    }
    LAUNCH(PC_CONTROL, ba::kb_begin_trial, g_winw, 128, 0, d, b);
    fork(1);
    if (tiled) {
      if (b.Cp && td.Tp) LAUNCH(PC_SCHUR_TILE, ba::kt_schur_tile<0>, g_tp, ba::TILE_THREADS, smem_tile[0], d, b, lo, td);
      if (b.Cl && td.Tl) LAUNCH_LN(PC_SCHUR_TILE, ba::kt_schur_tile<1>, g_tl, ba::TILE_THREADS, smem_tile[1], d, b, lo, td);
    } else {
      if (b.Cp) LAUNCH(PC_SCHUR_PREP, ba::kb_schur_prep<0>, g_pt, ba::BT, 0, d, b, lo);
      if (b.Cl) LAUNCH_LN(PC_SCHUR_PREP, ba::kb_schur_prep<1>, g_ln, ba::BT, 0, d, b, lo);
    }
    join(1);
    if (!tiled) LAUNCH(PC_SCHUR_REDUCE, ba::kb_schur_reduce, g_ne, 32, 0, d, b);
    if (tiled) {
      LAUNCH(PC_SOLVE, ba::kt_tile_sum, dim3((b.Pmax * 42 * ba::TILE_SUM_LANES + 255) / 256, W), 256, 0, d, b, td);
      LAUNCH(PC_SOLVE, ba::kb_solve<true>, W, 256, smem_solve, d, b, td);
    } else if (!dense) {
      LAUNCH(PC_SOLVE, ba::kb_solve<false>, W, 256, smem_solve, d, b, td);
    } else {
      if (global) { // rank-local Schur complement pieces -> sum over ranks (+ the Cholesky-failure flag in the tail)
        ProfScope ps_(c, PC_COLLECTIVE);
        if (coll_rc == RSPL_BA_OK)
          coll_rc = comm_all_reduce(c, b.hs_part_w, b.hs_part, (size_t)42 * n_ne_host[0] + 1, kNcclFloat64, kNcclSum);
      }
      if (dl.bcr_bsp) { // banded single window: hand-written cyclic reduction (bcr_solver.cuh)
        if (dense_rc == RSPL_BA_OK) dense_rc = bcr_assemble_solve(c, dl, n_sys_host[0], n_ne_host[0]);
      } else {
        {
          ProfScope ps_(c, PC_ASSEMBLE);
          if (cudaMemsetAsync(dl.H, 0, sizeof(double) * (size_t)dl.total, s) != cudaSuccess) dense_rc = RSPL_BA_ERR_CUDA;
        }
        LAUNCH(PC_ASSEMBLE, ba::kb_assemble_dense, g_ne, 64, 0, d, b);
        {
          ProfScope ps_(c, PC_SOLVE);
          dense_rc = dense_factor_solve(c, dl, n_sys_host);
        }
      }
      LAUNCH(PC_ASSEMBLE, ba::kb_post_solve, W, 256, 0, d, b);
    }
    fork(2);
    if (tiled) {
      if (b.Cp) LAUNCH(PC_BACKSUB, ba::kt_backsub_rc<0>, g_pt, ba::BT, 0, d, b, lo, td);
      if (b.Cl) LAUNCH_LN(PC_BACKSUB, ba::kt_backsub_rc<1>, g_ln, ba::BT, 0, d, b, lo, td);
    } else {
      if (b.Cp) LAUNCH(PC_BACKSUB, ba::kb_backsub<0>, g_pt, ba::BT, 0, d, b, lo);
      if (b.Cl) LAUNCH_LN(PC_BACKSUB, ba::kb_backsub<1>, g_ln, ba::BT, 0, d, b, lo);
    }
    join(2);
    if (global) {
      LAUNCH(PC_CONTROL, ba::kb_global_sums, 1, 256, 0, b, 1);
      ProfScope ps_(c, PC_COLLECTIVE);
      if (coll_rc == RSPL_BA_OK) coll_rc = comm_all_reduce(c, b.gs_w + 4, b.gs + 4, 2, kNcclFloat64, kNcclSum);
    }
    LAUNCH(PC_CONTROL, ba::kb_decide<false>, g_winw, 128, 0, d, b, cudaGraphConditionalHandle{});
    if (b.C) LAUNCH(PC_CONTROL, ba::kb_restore, g_lm, ba::BT, 0, d, b);
  };
  // ... captured once into a CUDA graph and replayed (launch-bound for small batches: a C1 window
  // spends ~32 super-steps x 12 tiny kernels). Profiling needs event pairs around every launch,
  // which a captured graph cannot carry, so it falls back to plain launches.
  cudaGraphExec_t gexec = nullptr;
  int64_t launches_per_step = 0;
  if (!c->prof && !dense) { // library calls of the dense path are enqueued directly
    cudaGraph_t graph = nullptr;
    const int64_t before = c->launches;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      super_step();
      if (cudaStreamEndCapture(s, &graph) == cudaSuccess && graph) {
        if (cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) gexec = nullptr;
        cudaGraphDestroy(graph);
      }
    }
    launches_per_step = c->launches - before;
    c->launches = before;
    cudaGetLastError();
  }
  struct GraphGuard {
    cudaGraphExec_t& g;
    ~GraphGuard() {
      if (g) cudaGraphExecDestroy(g);
    }
  } graph_guard{gexec};
  for (int pass = 0; pass < 2; ++pass) {
    // active-edge counts per pose; in global mode a pose is in the reduced system if ANY rank holds an active edge of it
    CU_TRY(c, cudaMemsetAsync(b.pact_w, 0, sizeof(int) * (size_t)c->l_np, s));
    if (b.C) LAUNCH(PC_CONTROL, ba::kb_mark_active, g_lm, ba::BT, 0, d, b);
    if (global) {
      {
        ProfScope ps_(c, PC_COLLECTIVE);
        coll_rc = comm_all_reduce(c, b.pact_w, b.pact, (size_t)c->l_np, kNcclInt32, kNcclSum);
      }
      if (coll_rc != RSPL_BA_OK) return coll_rc;
    }
    LAUNCH(PC_CONTROL, ba::kb_begin_pass, g_win, 128, 0, d, b, lo, pass);
    LAUNCH(PC_CONTROL, ba::kb_pair_blocks, dim3((b.Pmax + 255) / 256, W), 256, 0, d, b);
    if (dense) { // the host needs the system sizes of this pass for the library calls
      std::vector<ba::WinState> ws(W);
      CU_TRY(c, cudaMemcpyAsync(ws.data(), b.ws, sizeof(ba::WinState) * W, cudaMemcpyDeviceToHost, s));
      CU_TRY(c, cudaStreamSynchronize(s));
      for (int w = 0; w < W; ++w) n_sys_host[w] = ws[w].n_sys;
    }
    const int worst = lo.iters[pass] * 10 + 1; // <= 10 trials per LM iteration (§9.9)
    int done_steps = 0;
    while (done_steps < worst) {
      // (a super-step of the dense path costs milliseconds: poll after every one instead of wasting factorisations)
      const int burst = dense ? 1 : (done_steps == 0 ? (lo.iters[pass] < 4 ? lo.iters[pass] : 4) : 4);
      for (int k = 0; k < burst && done_steps < worst; ++k, ++done_steps) {
        if (gexec) {
          CU_TRY(c, cudaGraphLaunch(gexec, s));
          c->launches += launches_per_step;
        } else {
          super_step();
        }
        c->l_super_steps++;
      }
      CU_TRY(c, cudaGetLastError());
      if (dense_rc != RSPL_BA_OK) return dense_rc;
      if (coll_rc != RSPL_BA_OK) return coll_rc;
      if (lo.iters[pass] == 0) break;
      // poll: how many windows are still iterating?
      int n_active = 0;
      CU_TRY(c, cudaMemsetAsync(b.n_active, 0, sizeof(int), s));
      LAUNCH(PC_CONTROL, ba::kb_count_active, 32, 256, 0, d, b);
      CU_TRY(c, cudaMemcpyAsync(&n_active, b.n_active, sizeof(int), cudaMemcpyDeviceToHost, s));
      CU_TRY(c, cudaStreamSynchronize(s));
      if (n_active == 0) break;
    }
    if (pass == 0) LAUNCH(PC_FLAG_WRITEBACK, ba::kb_flag<false>, dim3(W, edge_chunks), 256, 0, d, b, lo);
  }
  LAUNCH(PC_FLAG_WRITEBACK, ba::kb_flag<true>, dim3(W, edge_chunks), 256, 0, d, b, lo);
  LAUNCH(PC_FLAG_WRITEBACK, ba::kb_writeback, dim3(W, lm_chunks), 256, 0, d, b);
#undef LAUNCH
#undef LAUNCH_LN
  CU_TRY(c, cudaGetLastError());
  return RSPL_BA_OK;
}

} // namespace

extern "C" int rspl_ba_local_batch_solve(RsplBaContext* c, const RsplBaOptions* opt) {
  if (!c || !opt) return RSPL_BA_ERR_INVALID;
  if (!c->local_uploaded) return fail(c, RSPL_BA_ERR_STATE, "local_batch_solve before upload");
  if (opt->local_iters_pass1 < 0 || opt->local_iters_pass2 < 0)
    return fail(c, RSPL_BA_ERR_INVALID, "negative iteration count");
  c->local_solved = true;
  if (c->l_n_windows == 0) return RSPL_BA_OK;
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  ba::LocalOpt lo{};
  lo.thr[0] = opt->thr_mono_point;
  lo.thr[1] = opt->thr_stereo_point;
  lo.thr[2] = opt->thr_mono_line;
  lo.thr[3] = opt->thr_stereo_line;
  for (int i = 0; i < 4; ++i) lo.delta[i] = (double)(float)sqrt(lo.thr[i]); // const float thHuber* = sqrt(cfg.*) (:77-78,:125-126)
  lo.iters[0] = opt->local_iters_pass1;
  lo.iters[1] = opt->local_iters_pass2;
  lo.bf_float = opt->stereo_bf_float;
  lo.max_poses = c->l_max_poses;
  lo.max_free = c->l_max_free_poses;
  if (c->l_n_windows > 65535) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: more than 65535 windows in one call");
  c->l_graph_launches_step = 0;
  {
    const int grc = local_solve_graph(c, lo);
    if (grc != RSPL_BA_ERR_STATE) return grc;
  }
  {
    ProfScope ps(c, PC_LOCAL_SETUP);
    enqueue_local_setup(c, c->stream);
  }
  CU_TRY(c, cudaGetLastError());
  // host-driven path: profiling, dense reduced systems (large windows, global BA)
  return local_solve_batched(c, lo);
}

extern "C" int rspl_ba_local_batch_download(RsplBaContext* c, RsplLocalBatchResult* out) {
  if (!c || !out) return RSPL_BA_ERR_INVALID;
  if (!c->local_solved) return fail(c, RSPL_BA_ERR_STATE, "local_batch_download before solve");
  if (c->l_n_windows == 0) return RSPL_BA_OK;
  const ba::LocalDev& d = c->ld;
  if (!out->pose_twc || (c->l_npt && !out->point_xyz) || (c->l_nln && !out->line_wd) ||
      (c->l_n[0] && !out->mp_inlier) || (c->l_n[1] && !out->sp_inlier) || (c->l_n[2] && !out->ml_inlier) ||
      (c->l_n[3] && !out->sl_inlier))
    return fail(c, RSPL_BA_ERR_INVALID, "local result: null output arrays");
  SetDevice guard(c->device);
  cudaStream_t s = c->stream;
  int err = 0, steps = 0;
#define D2H(dst, src, bytes)                                                                        \
  do {                                                                                              \
    if ((bytes) > 0) CU_TRY(c, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, s)); \
  } while (0)
  if (c->l_staged) {
    // one copy of the output range into the pinned mirror, then host copies into the caller's arrays
    char* base = c->local_buf.as<char>();
    char* stage = (char*)c->l_stage;
    CU_TRY(c, cudaMemcpyAsync(stage + c->l_out_begin, base + c->l_out_begin, c->l_out_end - c->l_out_begin, cudaMemcpyDeviceToHost, s));
    if (c->l_graph_launches_step) CU_TRY(c, cudaMemcpyAsync(&steps, c->bd.n_active, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    auto hp = [&](const void* dev) { return stage + ((const char*)dev - base); };
    auto planes = [&](double* dst, const double* dev, int comps, size_t n, size_t stride) {
      if (!n) return;
      const double* src = (const double*)hp(dev);
      if (stride == 0 || stride == n) memcpy(dst, src, sizeof(double) * comps * n);
      else
        for (int q = 0; q < comps; ++q) memcpy(dst + (size_t)q * stride, src + (size_t)q * n, sizeof(double) * n);
    };
    err = *(const int*)hp(d.err);
    planes(out->pose_twc, d.pose_out, 7, (size_t)c->l_np, c->l_stride.pose);
    planes(out->point_xyz, d.k[0].lm_out, 3, (size_t)c->l_npt, c->l_stride.lm[0]);
    planes(out->line_wd, d.k[1].lm_out, 6, (size_t)c->l_nln, c->l_stride.lm[1]);
    if (c->l_n[0]) memcpy(out->mp_inlier, hp(d.k[0].out_inl[0]), (size_t)c->l_n[0]);
    if (c->l_n[1]) memcpy(out->sp_inlier, hp(d.k[0].out_inl[1]), (size_t)c->l_n[1]);
    if (c->l_n[2]) memcpy(out->ml_inlier, hp(d.k[1].out_inl[0]), (size_t)c->l_n[2]);
    if (c->l_n[3]) memcpy(out->sl_inlier, hp(d.k[1].out_inl[1]), (size_t)c->l_n[3]);
    if (out->stats) memcpy(out->stats, hp(d.stats), sizeof(RsplBaStats) * c->l_n_windows);
  } else {
  D2H(&err, d.err, sizeof(int));
  if (c->l_graph_launches_step) D2H(&steps, c->bd.n_active, sizeof(int));
#define D2H_PLANES(dst, src, comps, n, stride)                                                                       \
  do {                                                                                                               \
    if ((stride) == 0 || (stride) == (size_t)(n)) D2H((dst), (src), sizeof(double) * (size_t)(comps) * (size_t)(n));  \
    else                                                                                                             \
      for (int q_ = 0; q_ < (comps); ++q_) D2H((dst) + (size_t)q_ * (stride), (src) + (size_t)q_ * (size_t)(n), sizeof(double) * (size_t)(n)); \
  } while (0)
  D2H_PLANES(out->pose_twc, d.pose_out, 7, c->l_np, c->l_stride.pose);
  D2H_PLANES(out->point_xyz, d.k[0].lm_out, 3, c->l_npt, c->l_stride.lm[0]);
  D2H_PLANES(out->line_wd, d.k[1].lm_out, 6, c->l_nln, c->l_stride.lm[1]);
#undef D2H_PLANES
  D2H(out->mp_inlier, d.k[0].out_inl[0], (size_t)c->l_n[0]);
  D2H(out->sp_inlier, d.k[0].out_inl[1], (size_t)c->l_n[1]);
  D2H(out->ml_inlier, d.k[1].out_inl[0], (size_t)c->l_n[2]);
  D2H(out->sl_inlier, d.k[1].out_inl[1], (size_t)c->l_n[3]);
  if (out->stats) D2H(out->stats, d.stats, sizeof(RsplBaStats) * c->l_n_windows);
  CU_TRY(c, cudaStreamSynchronize(s));
  }
#undef D2H
  if (c->l_graph_launches_step) { // the graph's super-steps ran without the host counting them
    c->launches += (int64_t)steps * c->l_graph_launches_step;
    c->l_super_steps = steps;
    c->l_graph_launches_step = 0;
  }
  if (err & ba::LOCAL_ERR_DUP_EDGE)
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: two constraints join the same (pose, landmark) pair");
  if (err & ba::LOCAL_ERR_DEGREE)
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local batch: a landmark has more than 254 observations");
  return RSPL_BA_OK;
}

namespace {

// Window range [w0, w1) of a batch as a batch of its own: rebased offset arrays, shifted array pointers; the component
// planes keep the stride of the whole batch (RsplBaContext::l_stride tells upload / download).
struct LocalChunk {
  std::vector<int32_t> beg[7]; // pose, point, line, mono pt, stereo pt, mono ln, stereo ln
  RsplLocalBatch in;
  RsplLocalBatchResult out;
  RsplBaContext::LocalStrides stride;
};

void make_local_chunk(const RsplLocalBatch* in, const RsplLocalBatchResult* out, int w0, int w1, LocalChunk& ch) {
  const int W = in->n_windows, n = w1 - w0;
  const int32_t* src[7] = {in->pose_begin, in->point_begin, in->line_begin, in->mono_pt_begin, in->stereo_pt_begin,
                           in->mono_ln_begin, in->stereo_ln_begin};
  size_t b0[7], tot[7];
  for (int a = 0; a < 7; ++a) {
    b0[a] = (size_t)src[a][w0];
    tot[a] = (size_t)src[a][W];
    ch.beg[a].resize(n + 1);
    for (int w = 0; w <= n; ++w) ch.beg[a][w] = src[a][w0 + w] - src[a][w0];
  }
  ch.in = *in;
  ch.in.n_windows = n;
  ch.in.pose_begin = ch.beg[0].data();
  ch.in.point_begin = ch.beg[1].data();
  ch.in.line_begin = ch.beg[2].data();
  ch.in.mono_pt_begin = ch.beg[3].data();
  ch.in.stereo_pt_begin = ch.beg[4].data();
  ch.in.mono_ln_begin = ch.beg[5].data();
  ch.in.stereo_ln_begin = ch.beg[6].data();
  auto sh = [](auto* p, size_t off) { return p ? p + off : p; };
  ch.in.pose_twc = sh(in->pose_twc, b0[0]);
  ch.in.pose_fixed = sh(in->pose_fixed, b0[0]);
  ch.in.point_xyz = sh(in->point_xyz, b0[1]);
  ch.in.line_wd = sh(in->line_wd, b0[2]);
  ch.in.mp_pose = sh(in->mp_pose, b0[3]);
  ch.in.mp_point = sh(in->mp_point, b0[3]);
  ch.in.mp_cam = sh(in->mp_cam, b0[3]);
  ch.in.mp_meas = sh(in->mp_meas, b0[3]);
  ch.in.sp_pose = sh(in->sp_pose, b0[4]);
  ch.in.sp_point = sh(in->sp_point, b0[4]);
  ch.in.sp_cam = sh(in->sp_cam, b0[4]);
  ch.in.sp_meas = sh(in->sp_meas, b0[4]);
  ch.in.ml_pose = sh(in->ml_pose, b0[5]);
  ch.in.ml_line = sh(in->ml_line, b0[5]);
  ch.in.ml_cam = sh(in->ml_cam, b0[5]);
  ch.in.ml_meas = sh(in->ml_meas, b0[5]);
  ch.in.sl_pose = sh(in->sl_pose, b0[6]);
  ch.in.sl_line = sh(in->sl_line, b0[6]);
  ch.in.sl_cam = sh(in->sl_cam, b0[6]);
  ch.in.sl_meas = sh(in->sl_meas, b0[6]);
  ch.out = *out;
  ch.out.pose_twc = sh(out->pose_twc, b0[0]);
  ch.out.point_xyz = sh(out->point_xyz, b0[1]);
  ch.out.line_wd = sh(out->line_wd, b0[2]);
  ch.out.mp_inlier = sh(out->mp_inlier, b0[3]);
  ch.out.sp_inlier = sh(out->sp_inlier, b0[4]);
  ch.out.ml_inlier = sh(out->ml_inlier, b0[5]);
  ch.out.sl_inlier = sh(out->sl_inlier, b0[6]);
  ch.out.stats = sh(out->stats, (size_t)w0);
  ch.stride.pose = tot[0];
  ch.stride.lm[0] = tot[1];
  ch.stride.lm[1] = tot[2];
  ch.stride.cls[0][0] = tot[3];
  ch.stride.cls[0][1] = tot[4];
  ch.stride.cls[1][0] = tot[5];
  ch.stride.cls[1][1] = tot[6];
}

// Large batches of independent windows: the windows are cut into chunks (balanced by constraint count) that go through
// child contexts with their own streams and workspaces, so that the upload of chunk k + 1 overlaps the solve of chunk k
// (the whole-schedule graph launch is asynchronous) and the download of chunk k the solve of chunk k + 1. Windows are
// independent and every window's result is independent of the batch it is in, so the bits do not change.
int local_batch_chunked(RsplBaContext* c, const RsplLocalBatch* in, const RsplBaOptions* opt, RsplLocalBatchResult* out,
                        int n_chunks) {
  const int W = in->n_windows;
  while ((int)c->kids.size() < n_chunks) {
    RsplBaContext* k = nullptr;
    const int rc = rspl_ba_create(c->device, nullptr, &k);
    if (rc != RSPL_BA_OK || !k) return fail(c, rc != RSPL_BA_OK ? rc : RSPL_BA_ERR_CUDA, "local batch: child context");
    c->kids.push_back(k);
  }
  // boundaries by cumulative constraint count
  std::vector<int> bounds(n_chunks + 1, W);
  bounds[0] = 0;
  {
    auto edges_upto = [&](int w) {
      return (long long)in->mono_pt_begin[w] + in->stereo_pt_begin[w] + in->mono_ln_begin[w] + in->stereo_ln_begin[w];
    };
    const long long total = edges_upto(W);
    int w = 0;
    for (int k = 1; k < n_chunks; ++k) {
      while (w < W && edges_upto(w) * n_chunks < total * k) ++w;
      bounds[k] = w;
    }
  }
  std::vector<LocalChunk> chunks(n_chunks);
  std::vector<int64_t> launches0(n_chunks);
  int rc = RSPL_BA_OK, bad = -1, n_started = 0;
  for (int k = 0; k < n_chunks && rc == RSPL_BA_OK; ++k) {
    if (bounds[k + 1] <= bounds[k]) continue;
    RsplBaContext* kc = c->kids[k];
    launches0[k] = kc->launches;
    make_local_chunk(in, out, bounds[k], bounds[k + 1], chunks[k]);
    kc->l_stride = chunks[k].stride;
    rc = rspl_ba_local_batch_upload(kc, &chunks[k].in);
    if (rc == RSPL_BA_OK) rc = rspl_ba_local_batch_solve(kc, opt);
    if (rc != RSPL_BA_OK) bad = k;
    n_started = k + 1;
  }
  for (int k = 0; k < n_started; ++k) {
    if (bounds[k + 1] <= bounds[k]) continue;
    RsplBaContext* kc = c->kids[k];
    if (rc == RSPL_BA_OK) {
      rc = rspl_ba_local_batch_download(kc, &chunks[k].out);
      if (rc != RSPL_BA_OK) bad = k;
    } else {
      cudaStreamSynchronize(kc->stream); // a failed call returns with nothing in flight
    }
    kc->l_stride = RsplBaContext::LocalStrides();
    c->launches += kc->launches - launches0[k];
  }
  // (the batch of an earlier rspl_ba_local_batch_upload stays resident in this context: the chunks lived in the children)
  if (rc != RSPL_BA_OK) return fail(c, rc, "%s", bad >= 0 ? c->kids[bad]->err : "local batch: chunk failed");
  return RSPL_BA_OK;
}

} // namespace

extern "C" int rspl_ba_local_batch(RsplBaContext* c, const RsplLocalBatch* in, const RsplBaOptions* opt,
                                   RsplLocalBatchResult* out) {
  if (c && in && opt && out && in->n_windows >= 128 && !c->prof && !c->global_mode && c->comm_ranks == 1 &&
      in->pose_begin && in->mono_pt_begin && in->stereo_pt_begin && in->mono_ln_begin && in->stereo_ln_begin &&
      in->point_begin && in->line_begin) {
    int n_chunks = 4;
    if (const char* e = getenv("RSPL_BA_LOCAL_CHUNKS")) n_chunks = atoi(e) > 0 ? atoi(e) : n_chunks;
    if (n_chunks > 16) n_chunks = 16;
    const int W = in->n_windows;
    // (the offset arrays are only trusted after this cheap check; everything else is validated per chunk)
    const bool offs = offsets_ok(in->pose_begin, W) && offsets_ok(in->point_begin, W) && offsets_ok(in->line_begin, W) &&
                      offsets_ok(in->mono_pt_begin, W) && offsets_ok(in->stereo_pt_begin, W) &&
                      offsets_ok(in->mono_ln_begin, W) && offsets_ok(in->stereo_ln_begin, W);
    const long long edges = offs ? (long long)in->mono_pt_begin[W] + in->stereo_pt_begin[W] + in->mono_ln_begin[W] + in->stereo_ln_begin[W] : 0;
    long long min_mb = 64; // below that the upload is too short to be worth overlapping
    if (const char* e = getenv("RSPL_BA_LOCAL_CHUNK_MIN_MB")) min_mb = atoll(e) >= 0 ? atoll(e) : min_mb;
    if (offs && n_chunks > 1 && edges * 36 >= (min_mb << 20)) return local_batch_chunked(c, in, opt, out, n_chunks);
  }
  int rc = rspl_ba_local_batch_upload(c, in);
  if (rc != RSPL_BA_OK) return rc;
  rc = rspl_ba_local_batch_solve(c, opt);
  if (rc != RSPL_BA_OK) return rc;
  return rspl_ba_local_batch_download(c, out);
}

// ---- SURVEY 8(f) rank 2 on the resident result: the endpoint refresh of Map::UppdateMapline (map.cc:121-177) for
// all lines of the batch just solved, reading the optimised Line3Ds and map points where the solve left them in HBM.
// Only the CSR lists travel up (4 bytes per point reference), endpoints and flags come back.
extern "C" int rspl_ba_local_batch_update_maplines(RsplBaContext* c, const int32_t* pt_begin, const int32_t* pt_index,
                                                   double* endpoints, uint8_t* out_ok, int32_t* n_done) {
  if (!c) return RSPL_BA_ERR_INVALID;
  if (n_done) *n_done = 0;
  if (!c->local_solved) return fail(c, RSPL_BA_ERR_STATE, "local_batch_update_maplines before solve");
  const int n_lines = c->l_nln, n_points = c->l_npt;
  if (n_lines == 0) return RSPL_BA_OK;
  if (!offsets_ok(pt_begin, n_lines) || !endpoints || !out_ok) return fail(c, RSPL_BA_ERR_INVALID, "update_maplines: bad offsets or null arrays");
  const int n_ref = pt_begin[n_lines];
  if (n_ref > 0 && !pt_index) return fail(c, RSPL_BA_ERR_INVALID, "update_maplines: null point indices");
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  Arena a;
  const size_t o_beg = a.take(sizeof(int) * ((size_t)n_lines + 1)), o_idx = a.take(sizeof(int) * (size_t)n_ref);
  const size_t o_end = a.take(sizeof(double) * 6 * (size_t)n_lines), o_ok = a.take((size_t)n_lines), o_cnt = a.take(sizeof(int));
  CU_TRY(c, c->unit_buf.reserve(a.off));
  char* base = c->unit_buf.as<char>();
  cudaStream_t s = c->stream;
  CU_TRY(c, cudaMemcpyAsync(base + o_beg, pt_begin, sizeof(int) * ((size_t)n_lines + 1), cudaMemcpyHostToDevice, s));
  if (n_ref > 0) CU_TRY(c, cudaMemcpyAsync(base + o_idx, pt_index, sizeof(int) * (size_t)n_ref, cudaMemcpyHostToDevice, s));
  CU_TRY(c, cudaMemsetAsync(base + o_end, 0, a.off - o_end, s)); // endpoints of lines that are not refreshed read 0
  if (!cams_ok(pt_index, n_ref, n_points)) { // (threaded range check, behind the uploads already queued)
    cudaStreamSynchronize(s);
    return fail(c, RSPL_BA_ERR_INVALID, "update_maplines: point index out of range");
  }
  ba::LineEndpointsDev d;
  d.n_lines = n_lines;
  d.n_points = n_points;
  d.line_stride = (size_t)n_lines; // the result planes of the solve are dense
  d.point_stride = (size_t)n_points;
  d.line_wd = c->ld.k[1].lm_out;
  d.pt_begin = (const int*)(base + o_beg);
  d.pt_index = (const int*)(base + o_idx);
  d.point_xyz = c->ld.k[0].lm_out;
  d.endpoints = (double*)(base + o_end);
  d.out_ok = (uint8_t*)(base + o_ok);
  d.n_done = (int*)(base + o_cnt);
  {
    ProfScope ps(c, PC_FRAME);
    ba::line_endpoints_kernel<<<(n_lines + ba::LINE_EP_THREADS - 1) / ba::LINE_EP_THREADS, ba::LINE_EP_THREADS, 0, s>>>(d);
  }
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  CU_TRY(c, cudaMemcpyAsync(endpoints, base + o_end, sizeof(double) * 6 * (size_t)n_lines, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaMemcpyAsync(out_ok, base + o_ok, (size_t)n_lines, cudaMemcpyDeviceToHost, s));
  int cnt = 0;
  CU_TRY(c, cudaMemcpyAsync(&cnt, base + o_cnt, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaStreamSynchronize(s));
  if (n_done) *n_done = cnt;
  return RSPL_BA_OK;
}

// ---- global BA: every rank uploads ONE window holding all poses (identical on every
// rank) and its own share of the landmarks with all their constraints; the solve is collective over the
// context's communicator (rspl_ba_comm_init). Upload and download are the local-batch calls.
extern "C" int rspl_ba_global_upload(RsplBaContext* c, const RsplLocalBatch* in) {
  if (in && in->n_windows != 1) return c ? fail(c, RSPL_BA_ERR_INVALID, "global BA: exactly one window per rank") : RSPL_BA_ERR_INVALID;
  return rspl_ba_local_batch_upload(c, in);
}
extern "C" int rspl_ba_global_solve(RsplBaContext* c, const RsplBaOptions* opt) {
  if (!c) return RSPL_BA_ERR_INVALID;
  c->global_mode = true;
  const int rc = rspl_ba_local_batch_solve(c, opt);
  c->global_mode = false;
  return rc;
}
extern "C" int rspl_ba_global_download(RsplBaContext* c, RsplLocalBatchResult* out) {
  return rspl_ba_local_batch_download(c, out);
}
