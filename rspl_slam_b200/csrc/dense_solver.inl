// dense_solver.inl — reduced camera systems that do not fit shared memory (part of capi.cu).
//
// The batched local-BA path factorises the reduced system of a window inside one CTA as long as
// n = 6 * (free poses) fits the 227 KB of shared memory (n <= ~160). Larger windows -- the
// global-BA end of the scale, SURVEY 8(e) C5: "replicated dense fp64 factorisation first" -- keep
// the system in HBM (kb_assemble_dense) and hand it to cuSOLVER's dense Cholesky, a plain library
// factorisation; everything around it (linearisation, Schur complement, back-substitution, LM
// control) stays in this library's kernels. cuSOLVER is loaded with dlopen on first use, so the
// library has no link-time dependency on it and the small-window paths never touch it.
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

struct DenseLayout {
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
int dense_prepare(RsplBaContext* c, DenseLayout& L) {
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
  if ((size_t)L.total * sizeof(double) > ((size_t)64 << 30))
    return fail(c, RSPL_BA_ERR_UNSUPPORTED, "dense reduced systems of this batch need more than 64 GB");
  Arena a;
  const size_t o_H = a.take(sizeof(double) * (size_t)(L.total + 1));
  const size_t o_b = a.take(sizeof(double) * (size_t)(6 * c->l_nf_begin[W] + 1));
  const size_t o_info = a.take(sizeof(int) * 2 * W); // [W] potrf info, [W] potrs info (parameter errors only)
  const size_t o_off = a.take(sizeof(long long) * W);
  CU_TRY(c, c->dense_buf.reserve(a.off));
  char* base = c->dense_buf.as<char>();
  L.H = (double*)(base + o_H);
  L.b = (double*)(base + o_b);
  L.info = (int*)(base + o_info);
  L.d_off = (long long*)(base + o_off);
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
  }
  L.work = (double*)(base + o_work);
  CU_TRY(c, cudaMemcpyAsync(L.d_off, L.off.data(), sizeof(long long) * W, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemsetAsync(L.info, 0, sizeof(int) * 2 * W, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
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
}
