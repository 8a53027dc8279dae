// dense_solver.inl — reduced camera systems that do not fit shared memory (part of capi.cu).
//
// The batched local-BA path factorises the reduced system of a window inside one CTA as long as
// n = 6 * (free poses) fits the 227 KB of shared memory (n <= ~160). Larger windows -- the
// global-BA end of the scale, SURVEY 8(e) C5: "replicated dense fp64 factorisation first" -- keep
// the system in HBM (kb_assemble_dense) and hand it to cuSOLVER's dense Cholesky, a plain library
// factorisation; everything around it (linearisation, Schur complement, back-substitution, LM
// control) stays in this library's kernels. cuSOLVER is loaded with dlopen on first use, so the
// library has no link-time dependency on it and the small-window paths never touch it.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <dlfcn.h>

namespace {

struct CusolverApi {
  void* lib = nullptr;
  decltype(&cusolverDnCreate) create = nullptr;
  decltype(&cusolverDnDestroy) destroy = nullptr;
  decltype(&cusolverDnSetStream) set_stream = nullptr;
  decltype(&cusolverDnDpotrf_bufferSize) potrf_buffer = nullptr;
  decltype(&cusolverDnDpotrf) potrf = nullptr;
  decltype(&cusolverDnDpotrs) potrs = nullptr;
  bool ok = false;
};

CusolverApi& cusolver_api() {
  static CusolverApi api;
  if (api.lib) return api;
  const char* names[] = {"libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
  api.create = (decltype(api.create))dlsym(api.lib, "cusolverDnCreate");
  api.destroy = (decltype(api.destroy))dlsym(api.lib, "cusolverDnDestroy");
  api.set_stream = (decltype(api.set_stream))dlsym(api.lib, "cusolverDnSetStream");
  api.potrf_buffer = (decltype(api.potrf_buffer))dlsym(api.lib, "cusolverDnDpotrf_bufferSize");
  api.potrf = (decltype(api.potrf))dlsym(api.lib, "cusolverDnDpotrf");
  api.potrs = (decltype(api.potrs))dlsym(api.lib, "cusolverDnDpotrs");
  api.ok = api.create && api.destroy && api.set_stream && api.potrf_buffer && api.potrf && api.potrs;
  return api;
}

// cuBLAS (same lazy loading) for the block-tridiagonal variant below
struct CublasApi {
  void* lib = nullptr;
  decltype(&cublasCreate_v2) create = nullptr;
  decltype(&cublasDestroy_v2) destroy = nullptr;
  decltype(&cublasSetStream_v2) set_stream = nullptr;
  decltype(&cublasDtrsm_v2) trsm = nullptr;
  decltype(&cublasDsyrk_v2) syrk = nullptr;
  decltype(&cublasDgemv_v2) gemv = nullptr;
  decltype(&cublasDtrsv_v2) trsv = nullptr;
  bool ok = false;
};

CublasApi& cublas_api() {
  static CublasApi api;
  if (api.lib) return api;
  const char* names[] = {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
  api.create = (decltype(api.create))dlsym(api.lib, "cublasCreate_v2");
  api.destroy = (decltype(api.destroy))dlsym(api.lib, "cublasDestroy_v2");
  api.set_stream = (decltype(api.set_stream))dlsym(api.lib, "cublasSetStream_v2");
  api.trsm = (decltype(api.trsm))dlsym(api.lib, "cublasDtrsm_v2");
  api.syrk = (decltype(api.syrk))dlsym(api.lib, "cublasDsyrk_v2");
  api.gemv = (decltype(api.gemv))dlsym(api.lib, "cublasDgemv_v2");
  api.trsv = (decltype(api.trsv))dlsym(api.lib, "cublasDtrsv_v2");
  api.ok = api.create && api.destroy && api.set_stream && api.trsm && api.syrk && api.gemv && api.trsv;
  return api;
}

// first non-zero potrf info of the tiles -> the window's info (block-tridiagonal variant)
__global__ void k_merge_tile_info(const int* tile_info, int n_tiles, int* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int v = 0;
    for (int i = 0; i < n_tiles && v == 0; ++i) v = tile_info[i];
    *out = v;
  }
}

struct DenseLayout {
  // Block-tridiagonal variant: when the pose pairs that share landmarks are all within `band` free-pose
  // indices of each other (a visual-odometry chain without loop closures), the reduced system is banded;
  // cut into tiles of >= band poses it is block-tridiagonal and its Cholesky factor has no fill outside
  // the tiles: O(n t^2) instead of O(n^3) with library calls on t x t tiles. tile_poses[w] = 0: full dense.
  std::vector<int> tile_poses;
  // Hand-written cyclic-reduction solver (bcr_solver.cuh): one window whose reduced system is banded within
  // bcr_bsp <= 24 poses and at least 4 super-blocks long. 0: not used.
  int bcr_bsp = 0;
  ba::BcrDev bcr{};
  int* tile_info = nullptr; // [max tiles]
  int max_tiles = 0;
  std::vector<long long> off; // [W] offset of each window's matrix (doubles)
  long long total = 0;        // doubles
  int lwork = 0;
  double* H = nullptr;
  double* b = nullptr;
  double* work = nullptr;
  int* info = nullptr;
  long long* d_off = nullptr;
};

// Allocates the dense systems of the uploaded batch (grow-only) and the cuSOLVER workspace.
int dense_prepare(RsplBaContext* c, DenseLayout& L, const std::vector<int>& band) {
  CusolverApi& api = cusolver_api();
  if (!api.ok)
    return fail(c, RSPL_BA_ERR_UNSUPPORTED,
                "reduced system does not fit shared memory and libcusolver.so.11 could not be loaded for the dense path");
  if (!c->cusolver) {
    cusolverDnHandle_t h = nullptr;
    if (api.create(&h) != CUSOLVER_STATUS_SUCCESS) return fail(c, RSPL_BA_ERR_CUDA, "cusolverDnCreate failed");
    c->cusolver = h;
  }
  cusolverDnHandle_t h = (cusolverDnHandle_t)c->cusolver;
  if (api.set_stream(h, c->stream) != CUSOLVER_STATUS_SUCCESS) return fail(c, RSPL_BA_ERR_CUDA, "cusolverDnSetStream failed");
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
  // banded single window (global BA, long chains): the hand-written cyclic-reduction solver replaces the library
  // calls and needs no dense n x n matrix
  int bcr_bsp = 0;
  if (W == 1 && !getenv("RSPL_BA_DENSE_FULL") && !getenv("RSPL_BA_DENSE_LIB")) {
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
  const size_t o_info = a.take(sizeof(int) * 2 * W); // [W] potrf info, [W] potrs info (parameter errors only)
  const size_t o_off = a.take(sizeof(long long) * W);
  // block-tridiagonal decision per window
  L.tile_poses.assign(W, 0);
  L.max_tiles = 0;
  const bool allow_tri = !getenv("RSPL_BA_DENSE_FULL") && cublas_api().ok;
  for (int w = 0; w < W && allow_tri; ++w) {
    const int nf = c->l_nf_begin[w + 1] - c->l_nf_begin[w];
    int min_tp = 64; // tiles of at least 384 unknowns: fewer, larger library calls (the chain of tiles is sequential)
    if (const char* e = getenv("RSPL_BA_TILE_POSES")) min_tp = atoi(e) > 0 ? atoi(e) : min_tp;
    int tp = band[w] > min_tp ? band[w] : min_tp;
    const int tiles = (nf + tp - 1) / tp;
    if (tiles >= 4) {
      L.tile_poses[w] = tp;
      if (tiles > L.max_tiles) L.max_tiles = tiles;
    }
  }
  L.bcr_bsp = bcr_bsp;
  size_t o_bD = 0, o_bE = 0, o_bGL = 0, o_bGR = 0, o_bg = 0;
  int bcr_M = 0;
  if (bcr_bsp) {
    const int nf = c->l_nf_begin[1] - c->l_nf_begin[0];
    bcr_M = (nf + bcr_bsp - 1) / bcr_bsp;
    const size_t bb = (size_t)36 * bcr_bsp * bcr_bsp;
    o_bD = a.take(sizeof(double) * bb * bcr_M);
    o_bE = a.take(sizeof(double) * bb * (2 * (size_t)bcr_M + ba::BCR_MAX_LEVELS));
    o_bGL = a.take(sizeof(double) * bb * bcr_M);
    o_bGR = a.take(sizeof(double) * bb * bcr_M);
    o_bg = a.take(sizeof(double) * 6 * bcr_bsp * bcr_M);
    L.tile_poses[0] = 0;
    L.max_tiles = 0;
  }
  const size_t o_tinfo = a.take(sizeof(int) * (L.max_tiles + 1));
  CU_TRY(c, c->dense_buf.reserve(a.off));
  char* base = c->dense_buf.as<char>();
  L.H = (double*)(base + o_H);
  L.b = (double*)(base + o_b);
  L.info = (int*)(base + o_info);
  L.d_off = (long long*)(base + o_off);
  L.tile_info = (int*)(base + o_tinfo);
  int lwork = 0;
  if (n_max > 0 && api.potrf_buffer(h, CUBLAS_FILL_MODE_LOWER, n_max, L.H, n_max, &lwork) != CUSOLVER_STATUS_SUCCESS)
    return fail(c, RSPL_BA_ERR_CUDA, "cusolverDnDpotrf_bufferSize failed");
  L.lwork = lwork;
  // workspace appended behind the systems (second reservation keeps the first pointers valid only if
  // nothing moved: reserve everything in one go)
  const size_t o_work = a.take(sizeof(double) * (size_t)(lwork + 1));
  if (a.off > c->dense_buf.cap) {
    CU_TRY(c, c->dense_buf.reserve(a.off));
    base = c->dense_buf.as<char>();
    L.H = (double*)(base + o_H);
    L.b = (double*)(base + o_b);
    L.info = (int*)(base + o_info);
    L.d_off = (long long*)(base + o_off);
    L.tile_info = (int*)(base + o_tinfo);
  }
  L.work = (double*)(base + o_work);
  if (L.bcr_bsp) {
    ba::BcrDev& s = L.bcr;
    s.bs = 6 * L.bcr_bsp;
    s.M = bcr_M;
    s.n = 0;
    s.levels = 0;
    s.D = (double*)(base + o_bD);
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
  }
  if (L.max_tiles > 0) {
    CublasApi& bl = cublas_api();
    if (!c->cublas) {
      cublasHandle_t bh = nullptr;
      if (bl.create(&bh) != CUBLAS_STATUS_SUCCESS) return fail(c, RSPL_BA_ERR_CUDA, "cublasCreate failed");
      c->cublas = bh;
    }
    if (bl.set_stream((cublasHandle_t)c->cublas, c->stream) != CUBLAS_STATUS_SUCCESS)
      return fail(c, RSPL_BA_ERR_CUDA, "cublasSetStream failed");
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
  int pch = (int)(((long long)c->smem_optin - 2048 - (long long)lbytes - (long long)sizeof(double) * s.bs) / (long long)(sizeof(double) * s.bs));
  if (pch > 2 * s.bs + 1) pch = 2 * s.bs + 1;
  if (!(pch & 1)) --pch; // odd: conflict-free transposed staging
  if (pch < 1) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "cyclic-reduction solver: super-block does not fit shared memory");
  const size_t smem_el = lbytes + sizeof(double) * s.bs * (pch + 1) + 64;
  const size_t smem_up = sizeof(double) * (2 * ba::BCR_KC * s.bs + ba::BCR_KC) + 64;
  Ml = s.M;
  for (int l = 0; l < levels; ++l) {
    const int n_odd = Ml / 2, n_even = (Ml + 1) / 2;
    if (n_odd > 0) ba::bcr_eliminate<<<n_odd, ba::BCR_THREADS, smem_el, st>>>(s, l, pch);
    if (ba::bcr_tiles_per_thread(s.bs) == 1) ba::bcr_update<1><<<n_even, ba::BCR_THREADS, smem_up, st>>>(s, l);
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

// potrf + potrs of every window whose system is non-empty; n_sys[w] = blocks in the current pass
int dense_factor_solve(RsplBaContext* c, const DenseLayout& L, const std::vector<int>& n_sys) {
  CusolverApi& api = cusolver_api();
  cusolverDnHandle_t h = (cusolverDnHandle_t)c->cusolver;
  for (int w = 0; w < c->l_n_windows; ++w) {
    const int n = 6 * n_sys[w];
    if (n == 0) continue;
    double* A = L.H + L.off[w];
    double* rhs = L.b + (size_t)6 * c->l_nf_begin[w];
    if (L.tile_poses[w] > 0) {
      // block-tridiagonal Cholesky; column-major lower view M(i, j) = A[j * n + i], i >= j
      CublasApi& bl = cublas_api();
      cublasHandle_t bh = (cublasHandle_t)c->cublas;
      const int t = 6 * L.tile_poses[w];
      const int T = (n + t - 1) / t;
      const double one = 1.0, minus = -1.0;
      auto tk = [&](int k) { return (k + 1) * t <= n ? t : n - k * t; };
      auto M = [&](int i, int j) { return A + (size_t)j * n + i; };
      int calls = 0;
      for (int k = 0; k < T; ++k) {
        const int o = k * t;
        if (api.potrf(h, CUBLAS_FILL_MODE_LOWER, tk(k), M(o, o), n, L.work, L.lwork, L.tile_info + k) != CUSOLVER_STATUS_SUCCESS)
          return fail(c, RSPL_BA_ERR_CUDA, "cusolverDnDpotrf (tile) failed");
        ++calls;
        if (k + 1 < T) {
          const int o1 = o + t;
          // L(k+1,k) = A(k+1,k) L(k,k)^-T ; A(k+1,k+1) -= L(k+1,k) L(k+1,k)^T
          if (bl.trsm(bh, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, tk(k + 1), tk(k), &one,
                      M(o, o), n, M(o1, o), n) != CUBLAS_STATUS_SUCCESS ||
              bl.syrk(bh, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, tk(k + 1), tk(k), &minus, M(o1, o), n, &one, M(o1, o1), n) !=
                  CUBLAS_STATUS_SUCCESS)
            return fail(c, RSPL_BA_ERR_CUDA, "cuBLAS trsm / syrk (tile) failed");
          calls += 2;
        }
      }
      k_merge_tile_info<<<1, 32, 0, c->stream>>>(L.tile_info, T, L.info + w);
      for (int k = 0; k < T; ++k) { // L y = b
        const int o = k * t;
        if (k > 0 && bl.gemv(bh, CUBLAS_OP_N, tk(k), t, &minus, M(o, o - t), n, rhs + o - t, 1, &one, rhs + o, 1) != CUBLAS_STATUS_SUCCESS)
          return fail(c, RSPL_BA_ERR_CUDA, "cuBLAS gemv (tile) failed");
        if (bl.trsv(bh, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, tk(k), M(o, o), n, rhs + o, 1) != CUBLAS_STATUS_SUCCESS)
          return fail(c, RSPL_BA_ERR_CUDA, "cuBLAS trsv (tile) failed");
        calls += 2;
      }
      for (int k = T - 1; k >= 0; --k) { // L^T x = y
        const int o = k * t;
        if (k + 1 < T &&
            bl.gemv(bh, CUBLAS_OP_T, tk(k + 1), t, &minus, M(o + t, o), n, rhs + o + t, 1, &one, rhs + o, 1) != CUBLAS_STATUS_SUCCESS)
          return fail(c, RSPL_BA_ERR_CUDA, "cuBLAS gemv (tile) failed");
        if (bl.trsv(bh, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, tk(k), M(o, o), n, rhs + o, 1) != CUBLAS_STATUS_SUCCESS)
          return fail(c, RSPL_BA_ERR_CUDA, "cuBLAS trsv (tile) failed");
        calls += 2;
      }
      c->launches += calls + 1;
      continue;
    }
    // row-major upper triangle == column-major lower triangle
    if (api.potrf(h, CUBLAS_FILL_MODE_LOWER, n, A, n, L.work, L.lwork, L.info + w) != CUSOLVER_STATUS_SUCCESS)
      return fail(c, RSPL_BA_ERR_CUDA, "cusolverDnDpotrf failed");
    // a failed factorisation leaves info > 0; the triangular solves then run on garbage, which
    // kb_post_solve ignores (rejected step, like g2o's LinearSolverEigen returning false)
    if (api.potrs(h, CUBLAS_FILL_MODE_LOWER, n, 1, A, n, rhs, n, L.info + c->l_n_windows + w) != CUSOLVER_STATUS_SUCCESS)
      return fail(c, RSPL_BA_ERR_CUDA, "cusolverDnDpotrs failed");
    c->launches += 2;
  }
  return RSPL_BA_OK;
}

} // namespace

static void dense_release(RsplBaContext* c) {
  if (c->cusolver && cusolver_api().ok) cusolver_api().destroy((cusolverDnHandle_t)c->cusolver);
  c->cusolver = nullptr;
  if (c->cublas && cublas_api().ok) cublas_api().destroy((cublasHandle_t)c->cublas);
  c->cublas = nullptr;
}
