// dense_solver.inl — reduced camera systems that do not fit shared memory (part of capi.cu).
//
// The batched local-BA path factorises the reduced system of a window inside one CTA as long as
// n = 6 * (free poses) fits the 227 KB of shared memory (n <= ~160). Larger windows keep the system in HBM and solve
// it with this library's own kernels (no cuSOLVER / cuBLAS anywhere):
//  * one window whose pose pairs share landmarks only within <= 24 poses of each other and that is at least four
//    super-blocks long (a keyframe chain: every problem of the C5 generator) -> block cyclic reduction
//    (bcr_solver.cuh), no dense matrix at all;
//  * everything else (windows of 27+ keyframes, systems with loop closures) -> kb_assemble_dense + the blocked
//    right-looking dense Cholesky of dense_chol.cuh.
// Everything around the solve (linearisation, Schur complement, back-substitution, LM control) is unchanged.

namespace {

struct DenseLayout {
  // Hand-written cyclic-reduction solver (bcr_solver.cuh): one window whose reduced system is banded within
  // bcr_bsp <= 24 poses and at least 4 super-blocks long. 0: not used.
  int bcr_bsp = 0;
  ba::BcrDev bcr{};
  std::vector<long long> off; // [W] offset of each window's matrix (doubles)
  long long total = 0;        // doubles
  double* H = nullptr;
  double* b = nullptr;
  double* Ld = nullptr;       // [n_max][DC_NB] panel diagonal factors (dense_chol.cuh)
  double* dinv = nullptr;     // [n_max]
  double* ywork = nullptr;    // [n_max] L^-1 b
  int* info = nullptr;
  long long* d_off = nullptr;
};

// Allocates the dense systems of the uploaded batch (grow-only) and the solver workspaces.
// the resident update keeps three super-blocks in shared memory; RSPL_BA_BCR_SLABS=1 forces the slab-staged kernel
static bool bcr_update_is_resident(const RsplBaContext* c, int bs) {
  const bool slabs = getenv("RSPL_BA_BCR_SLABS") != nullptr;
  return !slabs && bs <= ba::BCR_RESIDENT_BS_MAX && ba::bcr_tiles_per_thread(bs) == 1 && ba::bcr_resident_smem(bs) + 1024 <= (size_t)c->smem_optin;
}

int dense_prepare(RsplBaContext* c, DenseLayout& L, const std::vector<int>& band) {
  const int W = c->l_n_windows;
  L.off.assign(W, 0);
  L.total = 0;
  int n_max = 0;
  for (int w = 0; w < W; ++w) {
    const long long n = 6LL * (c->l_nf_begin[w + 1] - c->l_nf_begin[w]);
    L.off[w] = L.total;
    L.total += (n * n + 31) & ~31LL;
    if (n > n_max) n_max = (int)n;
  }
  // banded single window (global BA, long chains): cyclic reduction needs no dense n x n matrix
  int bcr_bsp = 0;
  if (W == 1 && !getenv("RSPL_BA_DENSE_FULL")) {
    const int nf = c->l_nf_begin[1] - c->l_nf_begin[0];
    int bsp = band[0] > 6 ? band[0] : 6;
    if (const char* e = getenv("RSPL_BA_BCR_POSES")) bsp = atoi(e) > bsp ? atoi(e) : bsp;
    if (6 * bsp <= ba::BCR_BS_MAX && nf >= 4 * bsp) bcr_bsp = bsp;
  }
  if (bcr_bsp) {
    L.total = 32;
    n_max = 0;
  }
  if ((size_t)L.total * sizeof(double) > ((size_t)64 << 30))
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "dense reduced systems of this batch need more than 64 GB");
  Arena a;
  const size_t o_H = a.take(sizeof(double) * (size_t)(L.total + 1));
  const size_t o_b = a.take(sizeof(double) * (size_t)(6 * c->l_nf_begin[W] + ba::BCR_BS_MAX + 8)); // (+ padding of the last super-block)
  const size_t o_info = a.take(sizeof(int) * 2 * W);
  const size_t o_off = a.take(sizeof(long long) * W);
  const size_t o_Ld = a.take(sizeof(double) * ((size_t)n_max * ba::DC_NB + 8));
  const size_t o_dinv = a.take(sizeof(double) * ((size_t)n_max + 8));
  const size_t o_yw = a.take(sizeof(double) * ((size_t)n_max + 8));
  L.bcr_bsp = bcr_bsp;
  size_t o_bD = 0, o_bF = 0, o_bE = 0, o_bGL = 0, o_bGR = 0, o_bg = 0;
  int bcr_M = 0;
  if (bcr_bsp) {
    const int nf = c->l_nf_begin[1] - c->l_nf_begin[0];
    bcr_M = (nf + bcr_bsp - 1) / bcr_bsp;
    const size_t bb = (size_t)36 * bcr_bsp * bcr_bsp;
    o_bD = a.take(sizeof(double) * bb * bcr_M);
    o_bF = a.take(sizeof(double) * bb * bcr_M);
    o_bE = a.take(sizeof(double) * bb * (2 * (size_t)bcr_M + ba::BCR_MAX_LEVELS));
    o_bGL = a.take(sizeof(double) * bb * bcr_M);
    o_bGR = a.take(sizeof(double) * bb * bcr_M);
    o_bg = a.take(sizeof(double) * 6 * bcr_bsp * bcr_M);
  }
  CU_TRY(c, c->dense_buf.reserve(a.off));
  char* base = c->dense_buf.as<char>();
  L.H = (double*)(base + o_H);
  L.b = (double*)(base + o_b);
  L.info = (int*)(base + o_info);
  L.d_off = (long long*)(base + o_off);
  L.Ld = (double*)(base + o_Ld);
  L.dinv = (double*)(base + o_dinv);
  L.ywork = (double*)(base + o_yw);
  if (L.bcr_bsp) {
    ba::BcrDev& s = L.bcr;
    s.bs = 6 * L.bcr_bsp;
    s.M = bcr_M;
    s.n = 0;
    s.levels = 0;
    s.D = (double*)(base + o_bD);
    s.F = (double*)(base + o_bF);
    s.E = (double*)(base + o_bE);
    s.GL = (double*)(base + o_bGL);
    s.GR = (double*)(base + o_bGR);
    s.g = (double*)(base + o_bg);
    s.x = L.b;
    s.info = L.info;
    const size_t ld = s.bs + 1;
    const size_t lbytes = sizeof(double) * s.bs * ld;
    CU_TRY(c, cudaFuncSetAttribute(ba::bcr_eliminate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->smem_optin - 1024)));
    CU_TRY(c, cudaFuncSetAttribute(ba::bcr_root, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(lbytes + sizeof(double) * 2 * s.bs + 64)));
    CU_TRY(c, cudaFuncSetAttribute(ba::bcr_backsub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(lbytes + sizeof(double) * 4 * s.bs + 64)));
    CU_TRY(c, cudaFuncSetAttribute(ba::bcr_update<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (2 * ba::BCR_KC * s.bs + ba::BCR_KC) + 64)));
    CU_TRY(c, cudaFuncSetAttribute(ba::bcr_update<ba::BCR_TILES_PER_THREAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (2 * ba::BCR_KC * s.bs + ba::BCR_KC) + 64)));
    if (bcr_update_is_resident(c, s.bs)) CU_TRY(c, cudaFuncSetAttribute(ba::bcr_update_resident, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba::bcr_resident_smem(s.bs)));
  } else {
    CU_TRY(c, cudaFuncSetAttribute(ba::dc_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba::DC_UPDATE_SMEM));
    CU_TRY(c, cudaFuncSetAttribute(ba::dc_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->smem_optin - ba::DC_SOLVE_STATIC_SMEM)));
  }
  CU_TRY(c, cudaMemcpyAsync(L.d_off, L.off.data(), sizeof(long long) * W, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemsetAsync(L.info, 0, sizeof(int) * 2 * W, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return RSPL_BA_OK;
}

// Assembly + factorisation + solve of the banded single window with the cyclic-reduction kernels (bcr_solver.cuh);
// n_sys = pose blocks of the reduced system in the current pass. Enqueues ~3 log2(M) launches, no host sync.
int bcr_assemble_solve(RsplBaContext* c, DenseLayout& L, int n_sys, int n_ne) {
  ba::BcrDev& s = L.bcr;
  cudaStream_t st = c->stream;
  const int bsp = L.bcr_bsp;
  s.n = 6 * n_sys;
  s.M = (n_sys + bsp - 1) / bsp;
  if (s.M < 1) s.M = 1;
  const size_t bb = (size_t)s.bs * s.bs;
  int Ml = s.M, levels = 0;
  long long off = 0;
  while (true) { // active blocks per level: M, ceil(M / 2), ...
    s.eoff[levels] = off;
    off += (long long)Ml * (long long)bb;
    if (Ml <= 1) break;
    Ml = (Ml + 1) / 2;
    ++levels;
    if (levels >= ba::BCR_MAX_LEVELS - 1) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "reduced system too long for the cyclic-reduction solver");
  }
  s.levels = levels;
  {
    ProfScope ps(c, PC_ASSEMBLE);
    CU_TRY(c, cudaMemsetAsync(s.D, 0, sizeof(double) * bb * s.M, st));
    CU_TRY(c, cudaMemsetAsync(s.E, 0, sizeof(double) * bb * s.M, st)); // level 0 (the higher levels are written in full)
    CU_TRY(c, cudaMemsetAsync(s.info, 0, sizeof(int), st));
    ba::bcr_pad<<<1, 64, 0, st>>>(s);
    ba::kb_assemble_bcr<<<n_ne > 0 ? n_ne : 1, 64, 0, st>>>(c->ld, c->bd, s, bsp);
    c->launches += 2;
  }
  ProfScope ps(c, PC_SOLVE);
  const size_t ld = s.bs + 1;
  const size_t lbytes = sizeof(double) * s.bs * ld;
  int pch = (int)(((long long)c->smem_optin - 2048 - (long long)lbytes - (long long)sizeof(double) * 7 * s.bs) / (long long)(sizeof(double) * s.bs));
  if (pch > 2 * s.bs + 1) pch = 2 * s.bs + 1;
  if (!(pch & 1)) --pch; // odd: conflict-free transposed staging
  if (pch < 1) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "cyclic-reduction solver: super-block does not fit shared memory");
  const size_t smem_el = lbytes + sizeof(double) * s.bs * (pch + 7) + 64; // factor, dinv, inverse diagonal blocks, panel
  const size_t smem_up = sizeof(double) * (2 * ba::BCR_KC * s.bs + ba::BCR_KC) + 64;
  Ml = s.M;
  for (int l = 0; l < levels; ++l) {
    const int n_odd = Ml / 2, n_even = (Ml + 1) / 2;
    if (n_odd > 0) {
      int slices = 1; // panel columns of a block over several CTAs while the level leaves SMs idle
      while (slices < 4 && n_odd * slices * 2 <= c->num_sms) slices *= 2;
      ba::bcr_eliminate<<<dim3(n_odd, slices), ba::BCR_THREADS, smem_el, st>>>(s, l, pch);
    }
    if (bcr_update_is_resident(c, s.bs)) {
      int slices = 1; // output tiles of a block over several CTAs while the level leaves SMs idle
      while (slices < 4 && n_even * slices * 2 <= c->num_sms) slices *= 2;
      ba::bcr_update_resident<<<dim3(n_even, slices), ba::BCR_THREADS, ba::bcr_resident_smem(s.bs), st>>>(s, l);
    }
    else if (ba::bcr_tiles_per_thread(s.bs) == 1) ba::bcr_update<1><<<n_even, ba::BCR_THREADS, smem_up, st>>>(s, l);
    else ba::bcr_update<ba::BCR_TILES_PER_THREAD><<<n_even, ba::BCR_THREADS, smem_up, st>>>(s, l);
    c->launches += 2;
    Ml = n_even;
  }
  ba::bcr_root<<<1, ba::BCR_THREADS, lbytes + sizeof(double) * 2 * s.bs + 64, st>>>(s);
  c->launches += 1;
  for (int l = levels - 1; l >= 0; --l) {
    const int Mlev = (s.M + (1 << l) - 1) >> l; // active blocks at level l
    const int n_odd = Mlev / 2;
    if (n_odd > 0) ba::bcr_backsub<<<n_odd, ba::BCR_THREADS, lbytes + sizeof(double) * 4 * s.bs + 64, st>>>(s, l);
    c->launches += 1;
  }
  CU_TRY(c, cudaGetLastError());
  return RSPL_BA_OK;
}

// Cholesky factorisation + solve (dense_chol.cuh) of every window whose system is non-empty; n_sys[w] = blocks in
// the current pass. Two launches per 48-column panel and one for the substitutions, no host synchronisation.
int dense_factor_solve(RsplBaContext* c, const DenseLayout& L, const std::vector<int>& n_sys) {
  cudaStream_t st = c->stream;
  CU_TRY(c, cudaMemsetAsync(L.info, 0, sizeof(int) * c->l_n_windows, st));
  for (int w = 0; w < c->l_n_windows; ++w) {
    const int n = 6 * n_sys[w];
    if (n == 0) continue;
    ba::DenseChol s;
    s.A = L.H + L.off[w];
    s.Ld = L.Ld;
    s.dinv = L.dinv;
    s.y = L.ywork;
    s.rhs = L.b + (size_t)6 * c->l_nf_begin[w];
    s.info = L.info + w; // a pivot <= 0 leaves info != 0; the substitutions then run on garbage, which kb_post_solve
                         // ignores (rejected step, like g2o's LinearSolverEigen returning false)
    s.n = n;
    for (int k0 = 0; k0 < n; k0 += ba::DC_NB) {
      const int kb = n - k0 < ba::DC_NB ? n - k0 : ba::DC_NB;
      const int rows = n - k0 - kb;
      ba::dc_panel<<<1 + (rows + ba::DC_THREADS - 1) / ba::DC_THREADS, ba::DC_THREADS, ba::dc_panel_smem(kb), st>>>(s, k0);
      const int T = (rows + ba::DC_TILE - 1) / ba::DC_TILE, n_tiles = T * (T + 1) / 2;
      ba::dc_update<<<n_tiles + 1, ba::DC_THREADS, ba::DC_UPDATE_SMEM, st>>>(s, k0, n_tiles);
      c->launches += 2;
    }
    const size_t xbytes = sizeof(double) * (size_t)n + 64;
    const bool in_smem = xbytes + ba::DC_SOLVE_STATIC_SMEM <= c->smem_optin;
    ba::dc_solve<<<1, ba::DC_SOLVE_THREADS, in_smem ? xbytes : 0, st>>>(s, in_smem ? 1 : 0);
    c->launches += 1;
  }
  CU_TRY(c, cudaGetLastError());
  return RSPL_BA_OK;
}

} // namespace

static void dense_release(RsplBaContext*) {}
