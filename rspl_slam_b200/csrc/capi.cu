// capi.cu — the C-ABI of include/rspl_ba.h: context, device workspaces, validation, launches.
//
// Host side of the drop-in boundary for /root/reference/src/g2o_optimization/g2o_optimization.cc
// (LocalmapOptimization :21-252, FrameOptimization :256-397). Everything numeric happens in the
// kernels of frame_kernel.cuh / local_kernel.cuh; this file only moves bytes and launches.
// There is no CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <thread>
#include <vector>

#include "../../include/rspl_ba.h"
#include "ba_math.cuh"
#include "frame_kernel.cuh"
#include "local_kernel.cuh"
#include "local_batched.cuh"
#include "local_tiled.cuh"
#include "dense_chol.cuh"
#include "triangulate.cuh"
#include "line_endpoints.cuh"

static_assert(sizeof(RsplBaStats) == sizeof(ba::DevStats), "stats layout");

namespace {

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// bump allocator over one DevBuf so a batch is a single allocation (256-byte aligned slices:
// every plane starts 16-byte aligned for 128-bit loads)
struct Arena {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  }
};

} // namespace

struct RsplBaContext {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t launches = 0;
  char err[512] = {0};
  int num_sms = 0;
  size_t smem_optin = 0;

  // ---- frame batch state
  DevBuf frame_buf;
  ba::FrameDev fd{};
  bool frame_uploaded = false;
  bool frame_solved = false;
  int f_n_frames = 0, f_n_mono = 0, f_n_stereo = 0;
  ba::Cam f_cam0{};
  std::vector<int32_t> f_mb, f_sb; // host copies of the frame edge offsets
  std::vector<int32_t> f_mlb, f_slb; // ... and of the line-extension offsets (zeros when absent)
  int f_n_mline = 0, f_n_sline = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr; // copy streams of the pipelined one-shot call
  cudaStream_t s_cmp[4] = {nullptr, nullptr, nullptr, nullptr}; // chunk kernels may overlap each other
  std::vector<cudaEvent_t> pipe_ev;
  size_t f_arena_bytes = 0;                 // size of the frame arena of the current batch
  void* f_stage = nullptr;                  // pinned mirror of the arena for small batches (one H2D, one D2H per call)
  size_t f_stage_cap = 0;

  // ---- local batch state
  // plane strides of the caller's arrays when the uploaded batch is a window range of a larger one (the chunked
  // one-shot call, local_capi.inl); 0 = the planes are contiguous (stride = count)
  struct LocalStrides {
    size_t pose = 0, lm[2] = {0, 0}, cls[2][2] = {{0, 0}, {0, 0}};
  } l_stride;
  void* l_stage = nullptr;           // pinned mirror of the input / output ranges of small local batches
  size_t l_stage_cap = 0, l_in_end = 0, l_out_begin = 0, l_out_end = 0;
  bool l_staged = false;
  std::vector<RsplBaContext*> kids; // child contexts of the chunked one-shot local call (own streams and workspaces)
  DevBuf local_buf;
  ba::LocalDev ld{};
  bool local_uploaded = false;
  bool local_solved = false;
  int l_n_windows = 0, l_np = 0, l_npt = 0, l_nln = 0, l_n[4] = {0, 0, 0, 0};
  int l_max_free_poses = 0, l_max_poses = 0;
  // batched (multi-kernel) path of the local batch
  DevBuf batch_buf;
  ba::BatchDev bd{};
  bool batch_ready = false;
  bool batch_global = false;
  std::vector<int> l_nf_begin;          // [W+1] free poses per window (prefix)
  std::vector<long long> l_pair_base;   // [W+1] capacity prefix of the pair lists
  int l_max_pts = 0, l_max_lns = 0, l_max_edges = 0;
  cudaStream_t s_aux = nullptr;           // the line kernels of a super-step run beside the point kernels
  cudaEvent_t fork_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t s_aux2 = nullptr;          // ... and the pose blocks beside both (graph path)
  cudaEvent_t fork_ev2[2] = {nullptr, nullptr};
  // tiled Schur path (local_tiled.cuh)
  DevBuf tile_buf;
  ba::TileDev td{};
  std::vector<long long> l_cost[2];     // [W] edges of a window's points / lines (tile sizing)
  std::vector<long long> l_cost_nl[2];  // [W] landmarks of each kind
  // whole-schedule CUDA graphs of the local solve (setup -> LM passes as conditional WHILE nodes -> flags), keyed by
  // the kernel argument blocks: a call whose shapes and buffers match an earlier one only launches the cached graph
  struct LocalGraph {
    std::vector<unsigned char> key;
    cudaGraphExec_t exec = nullptr;
    uint64_t stamp = 0;
    int launches_fixed = 0, launches_step = 0;
  };
  std::vector<LocalGraph> graph_cache;
  uint64_t graph_stamp = 0;
  cudaStream_t s_body = nullptr;          // origin stream of the loop-body capture
  int l_graph_launches_step = 0;          // kernels per super-step of the last graph launch (0: not a graph launch)
  int l_last_path = 0;                  // 2 host-driven + legacy Schur kernels, 3 host-driven + dense reduced solve, 4 tiled Schur (diagnostics)
  int l_super_steps = 0;
  // dense reduced-system solve (windows whose 6*NF x 6*NF system exceeds shared memory): dense_chol.cuh / bcr_solver.cuh
  DevBuf dense_buf;
  // global BA: NCCL communicator (comm.inl); the uploaded window is then one shard of the problem
  void* comm = nullptr; // ncclComm_t
  int comm_ranks = 1, comm_rank = 0;
  int64_t collectives = 0;
  bool global_mode = false;

  // ---- optional per-kernel-class timing with CUDA events on the context stream (rspl_ba_set_profiling)
  bool prof = false;
  struct ProfEv {
    cudaEvent_t a, b;
    int cls;
  };
  std::vector<ProfEv> prof_pool;
  size_t prof_used = 0;

  // ---- unit-level scratch
  DevBuf unit_buf;
};

namespace {

int fail(RsplBaContext* c, int code, const char* fmt, ...) {
  if (c) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof(c->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU_TRY(ctx, expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return fail(ctx, RSPL_BA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

bool offsets_ok(const int32_t* b, int n) {
  if (!b || b[0] != 0) return false;
  for (int i = 0; i < n; ++i)
    if (b[i + 1] < b[i]) return false;
  return true;
}

// The O(edges) input checks run on a few host threads: at 13.6 M constraints (C4) one thread needs ~15 ms to read
// the index arrays, longer than the 10 ms their upload takes on the link it is hidden behind.
template <class F>
bool parallel_all(long long n_items, long long min_per_thread, F&& ok_range) {
  unsigned hw = std::thread::hardware_concurrency();
  long long nt = hw ? (hw < 8 ? hw : 8) : 1;
  if (n_items / (min_per_thread > 0 ? min_per_thread : 1) < nt) nt = n_items / (min_per_thread > 0 ? min_per_thread : 1);
  if (nt <= 1) return ok_range(0, n_items);
  std::vector<std::thread> th;
  std::vector<char> ok((size_t)nt, 1);
  for (long long t = 0; t < nt; ++t)
    th.emplace_back([&, t]() { ok[(size_t)t] = ok_range(n_items * t / nt, n_items * (t + 1) / nt) ? 1 : 0; });
  bool all = true;
  for (long long t = 0; t < nt; ++t) {
    th[(size_t)t].join();
    all = all && ok[(size_t)t];
  }
  return all;
}

bool indices_ok(const int32_t* idx, const int32_t* ebeg, const int32_t* vbeg, int n_units) {
  // every edge of unit u references a vertex in [0, vbeg[u+1]-vbeg[u])
  // (branch-free inner loop so that the compiler vectorises it: this runs over every edge of a batch)
  const long long n_edges = n_units > 0 ? (long long)ebeg[n_units] : 0;
  return parallel_all(n_units, n_units > 0 ? (1LL << 18) * n_units / (n_edges > 0 ? n_edges : 1) + 1 : 1, [&](long long u0, long long u1) {
    unsigned bad = 0;
    for (long long u = u0; u < u1; ++u) {
      const unsigned nv = (unsigned)(vbeg[u + 1] - vbeg[u]);
      const int a = ebeg[u], b = ebeg[u + 1];
      for (int e = a; e < b; ++e) bad |= (unsigned)((unsigned)idx[e] >= nv);
    }
    return bad == 0;
  });
}

bool cams_ok(const int32_t* cam, int n, int n_cameras) {
  if (!cam) return true;
  return parallel_all(n, 1 << 18, [&](long long i0, long long i1) {
    unsigned bad = 0;
    for (long long i = i0; i < i1; ++i) bad |= (unsigned)((unsigned)cam[i] >= (unsigned)n_cameras);
    return bad == 0;
  });
}

struct SetDevice {
  int prev = -1;
  bool ok = true;
  explicit SetDevice(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~SetDevice() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

ba::FrameOpt make_frame_opt(const RsplBaOptions& o) {
  ba::FrameOpt f;
  f.thr_mono = o.thr_mono_point;
  f.thr_stereo = o.thr_stereo_point;
  f.delta_mono = (double)(float)sqrt(o.thr_mono_point);     // const float deltaMonoPoint = sqrt(cfg.mono_point) (:282)
  f.delta_stereo = (double)(float)sqrt(o.thr_stereo_point); // (:283)
  f.thr_mline = o.thr_mono_line;
  f.thr_sline = o.thr_stereo_line;
  f.delta_mline = (double)(float)sqrt(o.thr_mono_line);   // deltaMonoLine (:284, unused by the reference)
  f.delta_sline = (double)(float)sqrt(o.thr_stereo_line); // deltaStereoLine (:285)
  f.rounds = o.frame_rounds;
  f.iters = o.frame_iters;
  return f;
}

} // namespace

// ================================================================================================
// lifecycle
// ================================================================================================
extern "C" int rspl_ba_version(void) { return RSPL_BA_VERSION; }

extern "C" void rspl_ba_default_options(RsplBaOptions* o) {
  if (!o) return;
  o->thr_mono_point = 50.0; // configs/configs_euroc.yaml:57-60
  o->thr_stereo_point = 75.0;
  o->thr_mono_line = 50.0;
  o->thr_stereo_line = 75.0;
  o->local_iters_pass1 = 10;
  o->local_iters_pass2 = 5;
  o->frame_rounds = 4;
  o->frame_iters = 10;
  o->stereo_bf_float = 1;
  o->frame_latency_mode = 0;
}

extern "C" int rspl_ba_create(int device, void* stream, RsplBaContext** out) {
  if (!out) return RSPL_BA_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return RSPL_BA_ERR_CUDA; // no device: the product has no CPU path
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return RSPL_BA_ERR_CUDA;
  }
  if (device >= count) return RSPL_BA_ERR_INVALID;
  RsplBaContext* c = new (std::nothrow) RsplBaContext();
  if (!c) return RSPL_BA_ERR_CUDA;
  c->device = device;
  SetDevice guard(device);
  if (!guard.ok) {
    delete c;
    return RSPL_BA_ERR_CUDA;
  }
  if (stream) {
    c->stream = reinterpret_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return RSPL_BA_ERR_CUDA;
    }
    c->own_stream = true;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return RSPL_BA_ERR_CUDA;
  }
  c->num_sms = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  *out = c;
  return RSPL_BA_OK;
}

static void dense_release(RsplBaContext* c); // dense_solver.inl
static void comm_release(RsplBaContext* c);  // comm.inl

extern "C" void rspl_ba_destroy(RsplBaContext* c) {
  if (!c) return;
  for (RsplBaContext* k : c->kids) rspl_ba_destroy(k);
  c->kids.clear();
  SetDevice guard(c->device);
  cudaStreamSynchronize(c->stream);
  c->frame_buf.release();
  if (c->f_stage) cudaFreeHost(c->f_stage);
  if (c->l_stage) cudaFreeHost(c->l_stage);
  c->local_buf.release();
  c->batch_buf.release();
  c->tile_buf.release();
  c->dense_buf.release();
  dense_release(c);
  comm_release(c);
  c->unit_buf.release();
  for (cudaEvent_t e : c->pipe_ev) cudaEventDestroy(e);
  for (auto& pe : c->prof_pool) {
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  for (auto& g : c->graph_cache)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (c->s_body) cudaStreamDestroy(c->s_body);
  if (c->s_aux) cudaStreamDestroy(c->s_aux);
  if (c->s_aux2) cudaStreamDestroy(c->s_aux2);
  for (int i = 0; i < 2; ++i)
    if (c->fork_ev2[i]) cudaEventDestroy(c->fork_ev2[i]);
  for (int i = 0; i < 6; ++i)
    if (c->fork_ev[i]) cudaEventDestroy(c->fork_ev[i]);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  for (int i = 0; i < 4; ++i)
    if (c->s_cmp[i]) cudaStreamDestroy(c->s_cmp[i]);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" const char* rspl_ba_last_error(const RsplBaContext* c) { return c ? c->err : "null context"; }
extern "C" void* rspl_ba_stream(const RsplBaContext* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int rspl_ba_device(const RsplBaContext* c) { return c ? c->device : -1; }
extern "C" int64_t rspl_ba_launch_count(const RsplBaContext* c) { return c ? c->launches : 0; }
extern "C" int rspl_ba_sync(RsplBaContext* c) {
  if (!c) return RSPL_BA_ERR_INVALID;
  SetDevice guard(c->device);
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return RSPL_BA_OK;
}

// ---- kernel-class profiling -----------------------------------------------------------------------
namespace {
enum ProfClass {
  PC_FRAME = 0, PC_LOCAL_SETUP, PC_LOCAL_PERSISTENT, PC_PAIRS, PC_LINEARIZE, PC_POSE_BLOCKS, PC_SCHUR_PREP,
  PC_SCHUR_REDUCE, PC_SOLVE, PC_BACKSUB, PC_CONTROL, PC_FLAG_WRITEBACK, PC_COLLECTIVE, PC_ASSEMBLE, PC_SCHUR_TILE, PC_COUNT
};
// records an event pair around one launch when profiling is on
struct ProfScope { // event pair around launches on `stream` (default: the context stream)
  RsplBaContext* c;
  int idx = -1;
  cudaStream_t st;
  ProfScope(RsplBaContext* ctx, int cls, cudaStream_t stream = nullptr) : c(ctx), st(stream ? stream : ctx->stream) {
    if (!c->prof) return;
    if (c->prof_used == c->prof_pool.size()) {
      RsplBaContext::ProfEv e;
      if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
      c->prof_pool.push_back(e);
    }
    idx = (int)c->prof_used++;
    c->prof_pool[idx].cls = cls;
    cudaEventRecord(c->prof_pool[idx].a, st);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(c->prof_pool[idx].b, st);
  }
};
} // namespace

extern "C" int rspl_ba_set_profiling(RsplBaContext* c, int enabled) {
  if (!c) return RSPL_BA_ERR_INVALID;
  c->prof = enabled != 0;
  c->prof_used = 0;
  return RSPL_BA_OK;
}

static_assert(PC_COUNT <= RSPL_BA_PROFILE_CLASSES, "profile classes");
extern "C" int rspl_ba_get_profile(RsplBaContext* c, double* ms12, int64_t* launches12) {
  if (!c || !ms12 || !launches12) return RSPL_BA_ERR_INVALID;
  SetDevice guard(c->device);
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < RSPL_BA_PROFILE_CLASSES; ++i) {
    ms12[i] = 0;
    launches12[i] = 0;
  }
  for (size_t i = 0; i < c->prof_used; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->prof_pool[i].a, c->prof_pool[i].b) == cudaSuccess) {
      ms12[c->prof_pool[i].cls] += ms;
      launches12[c->prof_pool[i].cls] += 1;
    }
  }
  c->prof_used = 0;
  return RSPL_BA_OK;
}

extern "C" void* rspl_ba_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
extern "C" void rspl_ba_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

// ================================================================================================
// FrameOptimization batch
// ================================================================================================
namespace {

struct FrameOffsets {
  size_t cam, pose, mb, sb, mm, mx, mc, mi, sm, sx, sc, si, op, omi, osi, ml, sl, ni, st;
  // line extension
  size_t lmb, lsb, lml, lmm, lmc, lmi, lsl, lsm, lsc, lsi, olmi, olsi, lmlv, lslv;
};

// validates the batch, sizes the device arena and wires the kernel argument block (no copies yet)
int frame_prepare(RsplBaContext* c, const RsplFrameBatch* in, FrameOffsets& o, bool defer_cam_check = false) {
  c->frame_uploaded = c->frame_solved = false;
  const int F = in->n_frames;
  if (F < 0 || in->n_cameras < 1 || !in->cameras) return fail(c, RSPL_BA_ERR_INVALID, "frame batch: bad header");
  if (F == 0) {
    c->f_n_frames = c->f_n_mono = c->f_n_stereo = 0;
    return RSPL_BA_OK;
  }
  if (!in->pose_twc || !offsets_ok(in->mono_begin, F) || !offsets_ok(in->stereo_begin, F))
    return fail(c, RSPL_BA_ERR_INVALID, "frame batch: bad pose pointer or edge offsets");
  const int nm = in->mono_begin[F], ns = in->stereo_begin[F];
  if ((nm && (!in->mono_meas || !in->mono_xw)) || (ns && (!in->stereo_meas || !in->stereo_xw)))
    return fail(c, RSPL_BA_ERR_INVALID, "frame batch: null edge arrays");
  if (!defer_cam_check && (!cams_ok(in->mono_cam, nm, in->n_cameras) || !cams_ok(in->stereo_cam, ns, in->n_cameras)))
    return fail(c, RSPL_BA_ERR_INVALID, "frame batch: id_camera out of range");
  if (in->n_cameras > 1 && ((nm && !in->mono_cam) || (ns && !in->stereo_cam)))
    return fail(c, RSPL_BA_ERR_INVALID, "frame batch: several cameras but no per-edge camera index");
  // line extension: both offset arrays or neither
  int nml = 0, nsl = 0;
  if (in->mono_line_begin || in->stereo_line_begin) {
    if (!in->mono_line_begin || !in->stereo_line_begin || !offsets_ok(in->mono_line_begin, F) ||
        !offsets_ok(in->stereo_line_begin, F))
      return fail(c, RSPL_BA_ERR_INVALID, "frame batch: bad line offsets");
    nml = in->mono_line_begin[F];
    nsl = in->stereo_line_begin[F];
    if ((nml && (!in->mono_line_lw || !in->mono_line_meas)) || (nsl && (!in->stereo_line_lw || !in->stereo_line_meas)))
      return fail(c, RSPL_BA_ERR_INVALID, "frame batch: null line arrays");
    if (!defer_cam_check && (!cams_ok(in->mono_line_cam, nml, in->n_cameras) || !cams_ok(in->stereo_line_cam, nsl, in->n_cameras)))
      return fail(c, RSPL_BA_ERR_INVALID, "frame batch: id_camera out of range");
    if (in->n_cameras > 1 && ((nml && !in->mono_line_cam) || (nsl && !in->stereo_line_cam)))
      return fail(c, RSPL_BA_ERR_INVALID, "frame batch: several cameras but no per-edge camera index");
  }
  const bool has_lines = nml + nsl > 0;
  Arena a;
  o.cam = a.take(sizeof(double) * 5 * in->n_cameras);
  o.pose = a.take(sizeof(double) * 7 * F);
  o.mb = a.take(sizeof(int) * (F + 1));
  o.sb = a.take(sizeof(int) * (F + 1));
  o.mm = a.take(sizeof(double) * 2 * nm);
  o.mx = a.take(sizeof(double) * 3 * nm);
  o.mc = a.take(sizeof(int) * nm);
  o.mi = a.take(nm);
  o.sm = a.take(sizeof(double) * 3 * ns);
  o.sx = a.take(sizeof(double) * 3 * ns);
  o.sc = a.take(sizeof(int) * ns);
  o.si = a.take(ns);
  o.op = a.take(sizeof(double) * 7 * F);
  o.omi = a.take(nm);
  o.osi = a.take(ns);
  o.ml = a.take(nm);
  o.sl = a.take(ns);
  o.ni = a.take(sizeof(int) * F);
  o.st = a.take(sizeof(ba::DevStats) * F);
  if (has_lines) {
    o.lmb = a.take(sizeof(int) * (F + 1));
    o.lsb = a.take(sizeof(int) * (F + 1));
    o.lml = a.take(sizeof(double) * 6 * nml);
    o.lmm = a.take(sizeof(double) * 4 * nml);
    o.lmc = a.take(sizeof(int) * nml);
    o.lmi = a.take(nml);
    o.lsl = a.take(sizeof(double) * 6 * nsl);
    o.lsm = a.take(sizeof(double) * 8 * nsl);
    o.lsc = a.take(sizeof(int) * nsl);
    o.lsi = a.take(nsl);
    o.olmi = a.take(nml);
    o.olsi = a.take(nsl);
    o.lmlv = a.take(nml);
    o.lslv = a.take(nsl);
  }
  CU_TRY(c, c->frame_buf.reserve(a.off));
  c->f_arena_bytes = a.off;
  char* base = c->frame_buf.as<char>();
  ba::FrameDev& d = c->fd;
  d.n_frames = F;
  d.n_cameras = in->n_cameras;
  d.cameras = (const double*)(base + o.cam);
  d.pose_twc = (const double*)(base + o.pose);
  d.mono_begin = (const int*)(base + o.mb);
  d.stereo_begin = (const int*)(base + o.sb);
  d.n_mono = nm;
  d.n_stereo = ns;
  d.mono_meas = (const double*)(base + o.mm);
  d.mono_xw = (const double*)(base + o.mx);
  // (one camera: the per-edge indices are validated on the host and never read by the kernels, so they stay there)
  const bool multi_cam = in->n_cameras > 1;
  d.mono_cam = (multi_cam && in->mono_cam) ? (const int*)(base + o.mc) : nullptr;
  d.mono_inl_in = in->mono_inlier ? (const uint8_t*)(base + o.mi) : nullptr;
  d.stereo_meas = (const double*)(base + o.sm);
  d.stereo_xw = (const double*)(base + o.sx);
  d.stereo_cam = (multi_cam && in->stereo_cam) ? (const int*)(base + o.sc) : nullptr;
  d.stereo_inl_in = in->stereo_inlier ? (const uint8_t*)(base + o.si) : nullptr;
  d.out_pose_twc = (double*)(base + o.op);
  d.mono_inl = (uint8_t*)(base + o.omi);
  d.stereo_inl = (uint8_t*)(base + o.osi);
  d.mono_lvl = (uint8_t*)(base + o.ml);
  d.stereo_lvl = (uint8_t*)(base + o.sl);
  d.num_inliers = (int*)(base + o.ni);
  d.stats = (void*)(base + o.st);
  d.n_mline = nml;
  d.n_sline = nsl;
  if (has_lines) {
    d.mline_begin = (const int*)(base + o.lmb);
    d.sline_begin = (const int*)(base + o.lsb);
    d.mline_lw = (const double*)(base + o.lml);
    d.mline_meas = (const double*)(base + o.lmm);
    d.mline_cam = (multi_cam && in->mono_line_cam) ? (const int*)(base + o.lmc) : nullptr;
    d.mline_inl_in = in->mono_line_inlier ? (const uint8_t*)(base + o.lmi) : nullptr;
    d.sline_lw = (const double*)(base + o.lsl);
    d.sline_meas = (const double*)(base + o.lsm);
    d.sline_cam = (multi_cam && in->stereo_line_cam) ? (const int*)(base + o.lsc) : nullptr;
    d.sline_inl_in = in->stereo_line_inlier ? (const uint8_t*)(base + o.lsi) : nullptr;
    d.mline_inl = (uint8_t*)(base + o.olmi);
    d.sline_inl = (uint8_t*)(base + o.olsi);
    d.mline_lvl = (uint8_t*)(base + o.lmlv);
    d.sline_lvl = (uint8_t*)(base + o.lslv);
  } else {
    d.mline_begin = d.sline_begin = nullptr;
    d.mline_lw = d.mline_meas = d.sline_lw = d.sline_meas = nullptr;
    d.mline_cam = d.sline_cam = nullptr;
    d.mline_inl_in = d.sline_inl_in = nullptr;
    d.mline_inl = d.sline_inl = d.mline_lvl = d.sline_lvl = nullptr;
  }
  c->f_n_mline = nml;
  c->f_n_sline = nsl;
  c->f_cam0 = ba::Cam{in->cameras[0], in->cameras[1], in->cameras[2], in->cameras[3], in->cameras[4]};
  c->f_n_frames = F;
  c->f_n_mono = nm;
  c->f_n_stereo = ns;
  return RSPL_BA_OK;
}

// the per-edge camera indices of a batch (the scan frame_prepare skips when asked to: 6.5 MB of host reads on C2, which
// the pipelined call runs while the first chunk is already on the wire)
bool frame_cams_ok(const RsplBaContext* c, const RsplFrameBatch* in) {
  return cams_ok(in->mono_cam, c->f_n_mono, in->n_cameras) && cams_ok(in->stereo_cam, c->f_n_stereo, in->n_cameras) &&
         cams_ok(in->mono_line_cam, c->f_n_mline, in->n_cameras) && cams_ok(in->stereo_line_cam, c->f_n_sline, in->n_cameras);
}

// H2D of the frames [f0, f1) on stream s: their slice of every plane (+ the small header arrays if `header`)
int frame_copy_in(RsplBaContext* c, const RsplFrameBatch* in, const FrameOffsets& o, int f0, int f1, bool header,
                  cudaStream_t s) {
  char* base = c->frame_buf.as<char>();
  const int F = in->n_frames, nm = c->f_n_mono, ns = c->f_n_stereo;
#define H2D(off, src, bytes)                                                                             \
  do {                                                                                                   \
    if ((bytes) > 0)                                                                                     \
      CU_TRY(c, cudaMemcpyAsync(base + (off), (src), (bytes), cudaMemcpyHostToDevice, s));               \
  } while (0)
  if (header) {
    H2D(o.cam, in->cameras, sizeof(double) * 5 * in->n_cameras);
    H2D(o.pose, in->pose_twc, sizeof(double) * 7 * F);
    H2D(o.mb, in->mono_begin, sizeof(int) * (F + 1));
    H2D(o.sb, in->stereo_begin, sizeof(int) * (F + 1));
  }
  const size_t m0 = in->mono_begin[f0], m1 = in->mono_begin[f1], s0 = in->stereo_begin[f0], s1 = in->stereo_begin[f1];
  for (int k = 0; k < 2; ++k) H2D(o.mm + sizeof(double) * ((size_t)k * nm + m0), in->mono_meas + (size_t)k * nm + m0, sizeof(double) * (m1 - m0));
  for (int k = 0; k < 3; ++k) H2D(o.mx + sizeof(double) * ((size_t)k * nm + m0), in->mono_xw + (size_t)k * nm + m0, sizeof(double) * (m1 - m0));
  const bool multi_cam = in->n_cameras > 1;
  if (multi_cam && in->mono_cam) H2D(o.mc + sizeof(int) * m0, in->mono_cam + m0, sizeof(int) * (m1 - m0));
  if (in->mono_inlier) H2D(o.mi + m0, in->mono_inlier + m0, m1 - m0);
  for (int k = 0; k < 3; ++k) H2D(o.sm + sizeof(double) * ((size_t)k * ns + s0), in->stereo_meas + (size_t)k * ns + s0, sizeof(double) * (s1 - s0));
  for (int k = 0; k < 3; ++k) H2D(o.sx + sizeof(double) * ((size_t)k * ns + s0), in->stereo_xw + (size_t)k * ns + s0, sizeof(double) * (s1 - s0));
  if (multi_cam && in->stereo_cam) H2D(o.sc + sizeof(int) * s0, in->stereo_cam + s0, sizeof(int) * (s1 - s0));
  if (in->stereo_inlier) H2D(o.si + s0, in->stereo_inlier + s0, s1 - s0);
  if (c->f_n_mline + c->f_n_sline > 0) {
    const int nml = c->f_n_mline, nsl = c->f_n_sline;
    if (header) {
      H2D(o.lmb, in->mono_line_begin, sizeof(int) * (F + 1));
      H2D(o.lsb, in->stereo_line_begin, sizeof(int) * (F + 1));
    }
    const size_t a0 = in->mono_line_begin[f0], a1 = in->mono_line_begin[f1];
    const size_t b0 = in->stereo_line_begin[f0], b1 = in->stereo_line_begin[f1];
    for (int k = 0; k < 6; ++k) H2D(o.lml + sizeof(double) * ((size_t)k * nml + a0), in->mono_line_lw + (size_t)k * nml + a0, sizeof(double) * (a1 - a0));
    for (int k = 0; k < 4; ++k) H2D(o.lmm + sizeof(double) * ((size_t)k * nml + a0), in->mono_line_meas + (size_t)k * nml + a0, sizeof(double) * (a1 - a0));
    if (multi_cam && in->mono_line_cam) H2D(o.lmc + sizeof(int) * a0, in->mono_line_cam + a0, sizeof(int) * (a1 - a0));
    if (in->mono_line_inlier) H2D(o.lmi + a0, in->mono_line_inlier + a0, a1 - a0);
    for (int k = 0; k < 6; ++k) H2D(o.lsl + sizeof(double) * ((size_t)k * nsl + b0), in->stereo_line_lw + (size_t)k * nsl + b0, sizeof(double) * (b1 - b0));
    for (int k = 0; k < 8; ++k) H2D(o.lsm + sizeof(double) * ((size_t)k * nsl + b0), in->stereo_line_meas + (size_t)k * nsl + b0, sizeof(double) * (b1 - b0));
    if (multi_cam && in->stereo_line_cam) H2D(o.lsc + sizeof(int) * b0, in->stereo_line_cam + b0, sizeof(int) * (b1 - b0));
    if (in->stereo_line_inlier) H2D(o.lsi + b0, in->stereo_line_inlier + b0, b1 - b0);
  }
#undef H2D
  return RSPL_BA_OK;
}

// one instantiation of K7
template <bool SINGLE_CAM, bool HAS_LINES, int WPF>
cudaError_t frame_kernel_launch(int n_frames, cudaStream_t stream, const ba::FrameDev& fd, const ba::FrameOpt& fo) {
  constexpr int warps = WPF == 1 ? ba::FRAME_WARPS : WPF;
  constexpr int frames_per_cta = WPF == 1 ? ba::FRAME_WARPS : 1;
  const int grid = (n_frames + frames_per_cta - 1) / frames_per_cta;
  ba::frame_opt_kernel<SINGLE_CAM, HAS_LINES, WPF><<<grid, 32 * warps, 0, stream>>>(fd, fo);
  return cudaGetLastError();
}

int frame_launch(RsplBaContext* c, const RsplBaOptions* opt, int f0, int f1, cudaStream_t stream = nullptr) {
  if (!stream) stream = c->stream;
  ba::FrameOpt fo = make_frame_opt(*opt);
  fo.cam0 = c->f_cam0;
  fo.b0 = fo.cam0.bf / fo.cam0.fx;
  fo.frame0 = f0;
  fo.frame1 = f1;
  const bool single_cam = c->fd.n_cameras == 1 || (c->fd.mono_cam == nullptr && c->fd.stereo_cam == nullptr &&
                                                   c->fd.mline_cam == nullptr && c->fd.sline_cam == nullptr);
  cudaError_t e;
  {
    ProfScope ps(c, PC_FRAME, stream);
    const bool lines = c->f_n_mline + c->f_n_sline > 0;
    const int n = f1 - f0;
    if (opt->frame_latency_mode) { // one CTA per frame (single calls of the reference's FrameOptimization)
      constexpr int W8 = ba::FRAME_CTA_WARPS;
      if (single_cam && !lines) e = frame_kernel_launch<true, false, W8>(n, stream, c->fd, fo);
      else if (!lines) e = frame_kernel_launch<false, false, W8>(n, stream, c->fd, fo);
      else if (single_cam) e = frame_kernel_launch<true, true, W8>(n, stream, c->fd, fo);
      else e = frame_kernel_launch<false, true, W8>(n, stream, c->fd, fo);
    } else {
      if (single_cam && !lines) e = frame_kernel_launch<true, false, 1>(n, stream, c->fd, fo);
      else if (!lines) e = frame_kernel_launch<false, false, 1>(n, stream, c->fd, fo);
      else if (single_cam) e = frame_kernel_launch<true, true, 1>(n, stream, c->fd, fo);
      else e = frame_kernel_launch<false, true, 1>(n, stream, c->fd, fo);
    }
  }
  c->launches++;
  CU_TRY(c, e);
  return RSPL_BA_OK;
}

// D2H of the results of frames [f0, f1) on stream s. mb / sb: host copies of the edge offsets.
int frame_copy_out(RsplBaContext* c, RsplFrameBatchResult* out, const int32_t* mb, const int32_t* sb, int f0, int f1,
                   cudaStream_t s) {
  const ba::FrameDev& d = c->fd;
  const int F = c->f_n_frames;
  const size_t m0 = mb[f0], m1 = mb[f1], s0 = sb[f0], s1 = sb[f1];
  for (int k = 0; k < 7; ++k)
    CU_TRY(c, cudaMemcpyAsync(out->pose_twc + (size_t)k * F + f0, d.out_pose_twc + (size_t)k * F + f0,
                              sizeof(double) * (f1 - f0), cudaMemcpyDeviceToHost, s));
  if (m1 > m0) CU_TRY(c, cudaMemcpyAsync(out->mono_inlier + m0, d.mono_inl + m0, m1 - m0, cudaMemcpyDeviceToHost, s));
  if (s1 > s0) CU_TRY(c, cudaMemcpyAsync(out->stereo_inlier + s0, d.stereo_inl + s0, s1 - s0, cudaMemcpyDeviceToHost, s));
  if (c->f_n_mline + c->f_n_sline > 0) {
    const size_t a0 = c->f_mlb[f0], a1 = c->f_mlb[f1], b0 = c->f_slb[f0], b1 = c->f_slb[f1];
    if (a1 > a0) CU_TRY(c, cudaMemcpyAsync(out->mono_line_inlier + a0, d.mline_inl + a0, a1 - a0, cudaMemcpyDeviceToHost, s));
    if (b1 > b0) CU_TRY(c, cudaMemcpyAsync(out->stereo_line_inlier + b0, d.sline_inl + b0, b1 - b0, cudaMemcpyDeviceToHost, s));
  }
  if (out->num_inliers)
    CU_TRY(c, cudaMemcpyAsync(out->num_inliers + f0, d.num_inliers + f0, sizeof(int) * (f1 - f0), cudaMemcpyDeviceToHost, s));
  if (out->stats)
    CU_TRY(c, cudaMemcpyAsync(out->stats + f0, (const RsplBaStats*)d.stats + f0, sizeof(RsplBaStats) * (f1 - f0),
                              cudaMemcpyDeviceToHost, s));
  return RSPL_BA_OK;
}

void frame_keep_line_offsets(RsplBaContext* c, const RsplFrameBatch* in) {
  const int F = c->f_n_frames;
  if (c->f_n_mline + c->f_n_sline > 0) {
    c->f_mlb.assign(in->mono_line_begin, in->mono_line_begin + F + 1);
    c->f_slb.assign(in->stereo_line_begin, in->stereo_line_begin + F + 1);
  } else {
    c->f_mlb.assign(F + 1, 0);
    c->f_slb.assign(F + 1, 0);
  }
}

int ensure_pipeline(RsplBaContext* c, int n_chunks) {
  if (!c->s_in) CU_TRY(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  if (!c->s_out) CU_TRY(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i)
    if (!c->s_cmp[i]) CU_TRY(c, cudaStreamCreateWithFlags(&c->s_cmp[i], cudaStreamNonBlocking));
  while ((int)c->pipe_ev.size() < 2 * n_chunks) {
    cudaEvent_t e;
    CU_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->pipe_ev.push_back(e);
  }
  return RSPL_BA_OK;
}

} // namespace

extern "C" int rspl_ba_frame_batch_upload(RsplBaContext* c, const RsplFrameBatch* in) {
  if (!c || !in) return RSPL_BA_ERR_INVALID;
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  FrameOffsets o;
  int rc = frame_prepare(c, in, o);
  if (rc != RSPL_BA_OK) return rc;
  if (c->f_n_frames > 0) {
    rc = frame_copy_in(c, in, o, 0, c->f_n_frames, true, c->stream);
    if (rc != RSPL_BA_OK) return rc;
    CU_TRY(c, cudaStreamSynchronize(c->stream)); // caller buffers may be reused after return
    c->f_mb.assign(in->mono_begin, in->mono_begin + c->f_n_frames + 1);
    c->f_sb.assign(in->stereo_begin, in->stereo_begin + c->f_n_frames + 1);
    frame_keep_line_offsets(c, in);
  }
  c->frame_uploaded = true;
  return RSPL_BA_OK;
}

extern "C" int rspl_ba_frame_batch_solve(RsplBaContext* c, const RsplBaOptions* opt) {
  if (!c || !opt) return RSPL_BA_ERR_INVALID;
  if (!c->frame_uploaded) return fail(c, RSPL_BA_ERR_STATE, "frame_batch_solve before upload");
  if (opt->frame_rounds < 0 || opt->frame_iters < 1)
    return fail(c, RSPL_BA_ERR_INVALID, "frame_rounds must be >= 0 and frame_iters >= 1");
  c->frame_solved = true;
  if (c->f_n_frames == 0) return RSPL_BA_OK;
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  return frame_launch(c, opt, 0, c->f_n_frames);
}

extern "C" int rspl_ba_frame_batch_download(RsplBaContext* c, RsplFrameBatchResult* out) {
  if (!c || !out) return RSPL_BA_ERR_INVALID;
  if (!c->frame_solved) return fail(c, RSPL_BA_ERR_STATE, "frame_batch_download before solve");
  const int F = c->f_n_frames, nm = c->f_n_mono, ns = c->f_n_stereo;
  if (F == 0) return RSPL_BA_OK;
  if (!out->pose_twc || (nm && !out->mono_inlier) || (ns && !out->stereo_inlier) ||
      (c->f_n_mline && !out->mono_line_inlier) || (c->f_n_sline && !out->stereo_line_inlier))
    return fail(c, RSPL_BA_ERR_INVALID, "frame result: null output arrays");
  SetDevice guard(c->device);
  int rc = frame_copy_out(c, out, c->f_mb.data(), c->f_sb.data(), 0, F, c->stream);
  if (rc != RSPL_BA_OK) return rc;
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return RSPL_BA_OK;
}

// Small batches (the reference's call pattern: ONE frame, map_builder.cc:583-584): the ~30 per-plane copies of the
// general path cost more than the kernel (each cudaMemcpyAsync is 3-5 us of host time plus a DMA round trip). Here the
// inputs are gathered into a pinned mirror of the device arena on the host, the arena goes up in ONE copy, and the
// outputs (contiguous in the arena) come back in one copy per group and are scattered to the caller's arrays.
static constexpr size_t FRAME_STAGE_MAX = 256 << 10;

static int frame_batch_staged(RsplBaContext* c, const RsplFrameBatch* in, const RsplBaOptions* opt, RsplFrameBatchResult* out,
                       const FrameOffsets& o) {
  const int F = c->f_n_frames, nm = c->f_n_mono, ns = c->f_n_stereo, nml = c->f_n_mline, nsl = c->f_n_sline;
  const bool lines = nml + nsl > 0;
  if (c->f_stage_cap < c->f_arena_bytes) {
    if (c->f_stage) cudaFreeHost(c->f_stage);
    c->f_stage = nullptr;
    c->f_stage_cap = 0;
    CU_TRY(c, cudaHostAlloc(&c->f_stage, FRAME_STAGE_MAX, cudaHostAllocDefault));
    c->f_stage_cap = FRAME_STAGE_MAX;
  }
  char* h = (char*)c->f_stage;
  char* base = c->frame_buf.as<char>();
#define PUT(off, src, bytes)                          \
  do {                                                \
    if ((src) && (bytes) > 0) memcpy(h + (off), (src), (bytes)); \
  } while (0)
  PUT(o.cam, in->cameras, sizeof(double) * 5 * in->n_cameras);
  PUT(o.pose, in->pose_twc, sizeof(double) * 7 * F);
  PUT(o.mb, in->mono_begin, sizeof(int) * (F + 1));
  PUT(o.sb, in->stereo_begin, sizeof(int) * (F + 1));
  PUT(o.mm, in->mono_meas, sizeof(double) * 2 * nm);
  PUT(o.mx, in->mono_xw, sizeof(double) * 3 * nm);
  PUT(o.mc, in->n_cameras > 1 ? in->mono_cam : nullptr, sizeof(int) * nm);
  PUT(o.mi, in->mono_inlier, (size_t)nm);
  PUT(o.sm, in->stereo_meas, sizeof(double) * 3 * ns);
  PUT(o.sx, in->stereo_xw, sizeof(double) * 3 * ns);
  PUT(o.sc, in->n_cameras > 1 ? in->stereo_cam : nullptr, sizeof(int) * ns);
  PUT(o.si, in->stereo_inlier, (size_t)ns);
  size_t in_end = o.si + ns; // inputs of the point part end here; the outputs follow
  if (lines) {
    PUT(o.lmb, in->mono_line_begin, sizeof(int) * (F + 1));
    PUT(o.lsb, in->stereo_line_begin, sizeof(int) * (F + 1));
    PUT(o.lml, in->mono_line_lw, sizeof(double) * 6 * nml);
    PUT(o.lmm, in->mono_line_meas, sizeof(double) * 4 * nml);
    PUT(o.lmc, in->n_cameras > 1 ? in->mono_line_cam : nullptr, sizeof(int) * nml);
    PUT(o.lmi, in->mono_line_inlier, (size_t)nml);
    PUT(o.lsl, in->stereo_line_lw, sizeof(double) * 6 * nsl);
    PUT(o.lsm, in->stereo_line_meas, sizeof(double) * 8 * nsl);
    PUT(o.lsc, in->n_cameras > 1 ? in->stereo_line_cam : nullptr, sizeof(int) * nsl);
    PUT(o.lsi, in->stereo_line_inlier, (size_t)nsl);
    in_end = o.lsi + nsl; // one copy over the outputs of the point part in between (their content is don't-care)
  }
#undef PUT
  cudaStream_t s = c->stream;
  CU_TRY(c, cudaMemcpyAsync(base, h, in_end, cudaMemcpyHostToDevice, s));
  int rc = frame_launch(c, opt, 0, F, s);
  if (rc != RSPL_BA_OK) return rc;
  // outputs of the point part: [o.op, o.st + stats) is one contiguous range of the arena
  const size_t out0 = o.op, out1 = o.st + sizeof(ba::DevStats) * F;
  CU_TRY(c, cudaMemcpyAsync(h + out0, base + out0, out1 - out0, cudaMemcpyDeviceToHost, s));
  if (lines) CU_TRY(c, cudaMemcpyAsync(h + o.olmi, base + o.olmi, (o.olsi + nsl) - o.olmi, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaStreamSynchronize(s));
  memcpy(out->pose_twc, h + o.op, sizeof(double) * 7 * F);
  if (nm) memcpy(out->mono_inlier, h + o.omi, nm);
  if (ns) memcpy(out->stereo_inlier, h + o.osi, ns);
  if (nml) memcpy(out->mono_line_inlier, h + o.olmi, nml);
  if (nsl) memcpy(out->stereo_line_inlier, h + o.olsi, nsl);
  if (out->num_inliers) memcpy(out->num_inliers, h + o.ni, sizeof(int) * F);
  if (out->stats) memcpy(out->stats, h + o.st, sizeof(RsplBaStats) * F);
  return RSPL_BA_OK;
}

// One call with host buffers. Frames are independent, so the batch is cut into chunks and the
// three stages run as a pipeline on three streams: H2D of chunk k+1 and D2H of chunk k-1 overlap
// the kernel of chunk k (B200 has separate copy engines per direction). Returns when every result
// is in the caller's buffers.
extern "C" int rspl_ba_frame_batch(RsplBaContext* c, const RsplFrameBatch* in, const RsplBaOptions* opt,
                                   RsplFrameBatchResult* out) {
  if (!c || !in || !opt || !out) return RSPL_BA_ERR_INVALID;
  if (opt->frame_rounds < 0 || opt->frame_iters < 1)
    return fail(c, RSPL_BA_ERR_INVALID, "frame_rounds must be >= 0 and frame_iters >= 1");
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  FrameOffsets o;
  int rc = frame_prepare(c, in, o, true); // camera indices are checked below, under the first copy
  if (rc != RSPL_BA_OK) return rc;
  const int F = c->f_n_frames;
  c->frame_uploaded = c->frame_solved = true;
  if (F == 0) return RSPL_BA_OK;
  if (!out->pose_twc || (c->f_n_mono && !out->mono_inlier) || (c->f_n_stereo && !out->stereo_inlier) ||
      (c->f_n_mline && !out->mono_line_inlier) || (c->f_n_sline && !out->stereo_line_inlier))
    return fail(c, RSPL_BA_ERR_INVALID, "frame result: null output arrays");
  c->f_mb.assign(in->mono_begin, in->mono_begin + F + 1);
  c->f_sb.assign(in->stereo_begin, in->stereo_begin + F + 1);
  frame_keep_line_offsets(c, in);
  if (c->f_arena_bytes <= FRAME_STAGE_MAX) {
    if (!frame_cams_ok(c, in)) return fail(c, RSPL_BA_ERR_INVALID, "frame batch: id_camera out of range");
    return frame_batch_staged(c, in, opt, out, o);
  }
  // Pipeline of up to 4 equal chunks. End to end the call is bound by the host-to-device link (104 MB per C2 batch at
  // the ~26 GB/s this pool's boxes reach = 4.0 ms against 2.8 ms of kernels): chunk counts 4 / 6 / 8 and a small
  // first chunk (1/8 ... 1/32 of the batch) all measured 4.07 - 4.4 ms, equal quarters being the best.
  int n_chunks = F / 512;
  int max_chunks = 4;
  if (const char* e = getenv("RSPL_BA_FRAME_CHUNKS")) max_chunks = atoi(e) > 0 ? atoi(e) : max_chunks;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_chunks) n_chunks = max_chunks;
  std::vector<int> bounds(n_chunks + 1, 0);
  for (int k = 0; k <= n_chunks; ++k) bounds[k] = (int)((long long)F * k / n_chunks);
  rc = ensure_pipeline(c, n_chunks + 1);
  if (rc != RSPL_BA_OK) return rc;
  // everything queued earlier on the context stream completes first
  CU_TRY(c, cudaEventRecord(c->pipe_ev[2 * n_chunks], c->stream));
  CU_TRY(c, cudaStreamWaitEvent(c->s_in, c->pipe_ev[2 * n_chunks], 0));
  // on a failure in the middle of the pipeline every stream is drained before returning: copies into the caller's
  // buffers may still be in flight, and the caller is free to release them once the call is back
  auto drain = [&]() {
    cudaStreamSynchronize(c->s_in);
    for (int i = 0; i < 4; ++i) cudaStreamSynchronize(c->s_cmp[i]);
    cudaStreamSynchronize(c->s_out);
    cudaStreamSynchronize(c->stream);
  };
#define PIPE_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      drain();                                                                                \
      return fail(c, RSPL_BA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));              \
    }                                                                                         \
  } while (0)
  for (int k = 0; k < n_chunks; ++k) { // every upload is queued first ...
    rc = frame_copy_in(c, in, o, bounds[k], bounds[k + 1], k == 0, c->s_in);
    if (rc != RSPL_BA_OK) return drain(), rc;
    PIPE_TRY(cudaEventRecord(c->pipe_ev[2 * k], c->s_in));
  }
  if (!frame_cams_ok(c, in)) { // ... the host scans the camera indices while the copies run ...
    drain();
    return fail(c, RSPL_BA_ERR_INVALID, "frame batch: id_camera out of range");
  }
  for (int k = 0; k < n_chunks; ++k) { // ... and no kernel is launched on a batch that fails the scan
    const int f0 = bounds[k], f1 = bounds[k + 1];
    cudaStream_t cs = n_chunks > 1 ? c->s_cmp[k & 3] : c->stream;
    PIPE_TRY(cudaStreamWaitEvent(cs, c->pipe_ev[2 * k], 0));
    rc = frame_launch(c, opt, f0, f1, cs);
    if (rc != RSPL_BA_OK) return drain(), rc;
    PIPE_TRY(cudaEventRecord(c->pipe_ev[2 * k + 1], cs));
    PIPE_TRY(cudaStreamWaitEvent(c->s_out, c->pipe_ev[2 * k + 1], 0));
    rc = frame_copy_out(c, out, c->f_mb.data(), c->f_sb.data(), f0, f1, c->s_out);
    if (rc != RSPL_BA_OK) return drain(), rc;
  }
#undef PIPE_TRY
  CU_TRY(c, cudaStreamSynchronize(c->s_out));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return RSPL_BA_OK;
}

// ================================================================================================
// unit-level device entry points (parity tests)
// ================================================================================================
namespace ba {

__global__ void eval_edges_kernel(int type, int n, const double* pose7, const double* lm, const double* meas, Cam cam,
                                  int bf_float, double* err, double* Jl, double* Jp, double* chi2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* p = pose7 + 7 * i;
  double R[9];
  quat_to_R(p, R);
  const double t[3] = {p[4], p[5], p[6]};
  const double* X = lm + 6 * i;
  const double* m = meas + 8 * i;
  double r[4] = {0, 0, 0, 0}, jl[16], jp[24];
  for (int k = 0; k < 16; ++k) jl[k] = 0;
  for (int k = 0; k < 24; ++k) jp[k] = 0;
  double info = 1.0;
  int dim = 0;
  if (type == 0 || type == 1 || type == 4 || type == 5) {
    // types 4 / 5: g2o::EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose (g2o_optimization.cc:288-333):
    // the same residual and pose Jacobian with the world point held in the edge (Xw), no landmark Jacobian
    const bool only_pose = type >= 4;
    double Xc[3];
    transform_point(R, t, X, Xc);
    if (type == 1 || type == 5) {
      // (the `const float& bf` of EdgeStereoSE3ProjectXYZ::cam_project; the pose-only edge keeps bf a double)
      const double bf_res = (bf_float && !only_pose) ? (double)(float)cam.bf : cam.bf;
      point_residual<true>(cam, bf_res, Xc, m, r);
      point_jac_pose<true>(cam, Xc, jp);
      double j9[9];
      point_jac_point<true>(cam, R, Xc, j9);
      for (int k = 0; k < 9; ++k) jl[k] = only_pose ? 0.0 : j9[k];
      dim = 3;
    } else {
      point_residual<false>(cam, cam.bf, Xc, m, r);
      point_jac_pose<false>(cam, Xc, jp);
      double j6[9];
      point_jac_point<false>(cam, R, Xc, j6);
      for (int k = 0; k < 6; ++k) jl[k] = only_pose ? 0.0 : j6[k];
      dim = 2;
    }
  } else if (type == 2) {
    line_linearize<false>(cam, R, t, X, m, r, jp, jl);
    dim = 2;
    info = 0.1;
  } else {
    line_linearize<true>(cam, R, t, X, m, r, jp, jl);
    dim = 4;
    info = 0.1;
  }
  double c2 = 0;
  for (int k = 0; k < dim; ++k) c2 += r[k] * info * r[k];
  for (int k = 0; k < 4; ++k) err[4 * i + k] = r[k];
  for (int k = 0; k < 16; ++k) Jl[16 * i + k] = jl[k];
  for (int k = 0; k < 24; ++k) Jp[24 * i + k] = jp[k];
  chi2[i] = c2;
}

__global__ void oplus_kernel(int kind, int n, const double* state, const double* upd, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* s = state + 7 * i;
  const double* u = upd + 6 * i;
  double* o = out + 7 * i;
  for (int k = 0; k < 7; ++k) o[k] = 0;
  if (kind == 0) {
    Pose T;
    for (int k = 0; k < 4; ++k) T.q[k] = s[k];
    for (int k = 0; k < 3; ++k) T.t[k] = s[4 + k];
    Pose Tn = pose_oplus(T, u);
    for (int k = 0; k < 4; ++k) o[k] = Tn.q[k];
    for (int k = 0; k < 3; ++k) o[4 + k] = Tn.t[k];
  } else if (kind == 1) {
    for (int k = 0; k < 3; ++k) o[k] = s[k] + u[k];
  } else {
    line_oplus(s, u, o);
  }
}

} // namespace ba

extern "C" int rspl_ba_eval_edges(RsplBaContext* c, int edge_type, int32_t n, const double* pose7, const double* lm,
                                  const double* meas, const double* cam5, int32_t stereo_bf_float, double* err,
                                  double* Jl, double* Jp, double* chi2) {
  if (!c || n < 0 || edge_type < 0 || edge_type > 5 || !pose7 || !lm || !meas || !cam5 || !err || !Jl || !Jp || !chi2)
    return RSPL_BA_ERR_INVALID;
  if (n == 0) return RSPL_BA_OK;
  SetDevice guard(c->device);
  Arena a;
  const size_t o_p = a.take(sizeof(double) * 7 * n), o_l = a.take(sizeof(double) * 6 * n);
  const size_t o_m = a.take(sizeof(double) * 8 * n), o_e = a.take(sizeof(double) * 4 * n);
  const size_t o_jl = a.take(sizeof(double) * 16 * n), o_jp = a.take(sizeof(double) * 24 * n);
  const size_t o_c = a.take(sizeof(double) * n);
  CU_TRY(c, c->unit_buf.reserve(a.off));
  char* base = c->unit_buf.as<char>();
  cudaStream_t s = c->stream;
  CU_TRY(c, cudaMemcpyAsync(base + o_p, pose7, sizeof(double) * 7 * n, cudaMemcpyHostToDevice, s));
  CU_TRY(c, cudaMemcpyAsync(base + o_l, lm, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, s));
  CU_TRY(c, cudaMemcpyAsync(base + o_m, meas, sizeof(double) * 8 * n, cudaMemcpyHostToDevice, s));
  ba::Cam cam{cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
  ba::eval_edges_kernel<<<(n + 127) / 128, 128, 0, s>>>(edge_type, n, (const double*)(base + o_p),
                                                        (const double*)(base + o_l), (const double*)(base + o_m), cam,
                                                        stereo_bf_float, (double*)(base + o_e), (double*)(base + o_jl),
                                                        (double*)(base + o_jp), (double*)(base + o_c));
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  CU_TRY(c, cudaMemcpyAsync(err, base + o_e, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaMemcpyAsync(Jl, base + o_jl, sizeof(double) * 16 * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaMemcpyAsync(Jp, base + o_jp, sizeof(double) * 24 * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaMemcpyAsync(chi2, base + o_c, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaStreamSynchronize(s));
  return RSPL_BA_OK;
}

extern "C" int rspl_ba_oplus(RsplBaContext* c, int kind, int32_t n, const double* state, const double* upd,
                             double* out) {
  if (!c || n < 0 || kind < 0 || kind > 2 || !state || !upd || !out) return RSPL_BA_ERR_INVALID;
  if (n == 0) return RSPL_BA_OK;
  SetDevice guard(c->device);
  Arena a;
  const size_t o_s = a.take(sizeof(double) * 7 * n), o_u = a.take(sizeof(double) * 6 * n);
  const size_t o_o = a.take(sizeof(double) * 7 * n);
  CU_TRY(c, c->unit_buf.reserve(a.off));
  char* base = c->unit_buf.as<char>();
  cudaStream_t s = c->stream;
  CU_TRY(c, cudaMemcpyAsync(base + o_s, state, sizeof(double) * 7 * n, cudaMemcpyHostToDevice, s));
  CU_TRY(c, cudaMemcpyAsync(base + o_u, upd, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, s));
  ba::oplus_kernel<<<(n + 127) / 128, 128, 0, s>>>(kind, n, (const double*)(base + o_s), (const double*)(base + o_u),
                                                   (double*)(base + o_o));
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  CU_TRY(c, cudaMemcpyAsync(out, base + o_o, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaStreamSynchronize(s));
  return RSPL_BA_OK;
}

// Unit-level check of the range-test-free reciprocal / rsqrt / sqrt of ba_math.cuh: op 0 rcp_nr, 1 rsqrt_nr, 2 sqrt_nr.
namespace ba {
__global__ void unit_math_kernel(int op, int n, const double* in, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = in[i];
  out[i] = op == 0 ? rcp_nr(x) : op == 1 ? rsqrt_nr(x) : sqrt_nr(x);
}
} // namespace ba

extern "C" int rspl_ba_unit_math(RsplBaContext* c, int op, int32_t n, const double* in, double* out) {
  if (!c || op < 0 || op > 2 || n < 0 || (n && (!in || !out))) return RSPL_BA_ERR_INVALID;
  if (n == 0) return RSPL_BA_OK;
  SetDevice guard(c->device);
  Arena a;
  const size_t o_i = a.take(sizeof(double) * n), o_o = a.take(sizeof(double) * n);
  CU_TRY(c, c->unit_buf.reserve(a.off));
  char* base = c->unit_buf.as<char>();
  cudaStream_t s = c->stream;
  CU_TRY(c, cudaMemcpyAsync(base + o_i, in, sizeof(double) * n, cudaMemcpyHostToDevice, s));
  ba::unit_math_kernel<<<(n + 255) / 256, 256, 0, s>>>(op, n, (const double*)(base + o_i), (double*)(base + o_o));
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  CU_TRY(c, cudaMemcpyAsync(out, base + o_o, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaStreamSynchronize(s));
  return RSPL_BA_OK;
}

// Batched Map::TriangulateMappoint (triangulate.cuh). Host arrays in, host arrays out; out_xyz of a point that is not
// triangulated is left untouched. *n_done (optional) receives the number of points triangulated.
extern "C" int rspl_ba_triangulate_points(RsplBaContext* c, int32_t n_points, const int32_t* obs_begin,
                                          const int32_t* obs_frame, const double* obs_uv, int32_t n_frames,
                                          const double* frame_twc, const double* cam5, double* out_xyz, uint8_t* out_ok,
                                          int32_t* n_done) {
  if (!c || n_points < 0 || n_frames < 0 || !cam5) return RSPL_BA_ERR_INVALID;
  if (n_done) *n_done = 0;
  if (n_points == 0) return RSPL_BA_OK;
  if (!offsets_ok(obs_begin, n_points) || !out_xyz || !out_ok) return fail(c, RSPL_BA_ERR_INVALID, "triangulate: bad offsets or null outputs");
  const int n_obs = obs_begin[n_points];
  if (n_obs > 0 && (!obs_frame || !obs_uv || !frame_twc)) return fail(c, RSPL_BA_ERR_INVALID, "triangulate: null observation arrays");
  if (!(cam5[0] != 0.0) || !(cam5[1] != 0.0)) return fail(c, RSPL_BA_ERR_INVALID, "triangulate: zero focal length");
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  Arena a;
  const size_t o_beg = a.take(sizeof(int) * ((size_t)n_points + 1)), o_fr = a.take(sizeof(int) * (size_t)n_obs);
  const size_t o_uv = a.take(sizeof(double) * 2 * (size_t)n_obs), o_tw = a.take(sizeof(double) * 7 * (size_t)n_frames);
  const size_t o_xyz = a.take(sizeof(double) * 3 * (size_t)n_points), o_ok = a.take((size_t)n_points), o_cnt = a.take(sizeof(int));
  CU_TRY(c, c->unit_buf.reserve(a.off));
  char* base = c->unit_buf.as<char>();
  cudaStream_t s = c->stream;
  CU_TRY(c, cudaMemcpyAsync(base + o_beg, obs_begin, sizeof(int) * ((size_t)n_points + 1), cudaMemcpyHostToDevice, s));
  if (n_obs > 0) {
    CU_TRY(c, cudaMemcpyAsync(base + o_fr, obs_frame, sizeof(int) * (size_t)n_obs, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaMemcpyAsync(base + o_uv, obs_uv, sizeof(double) * 2 * (size_t)n_obs, cudaMemcpyHostToDevice, s));
  }
  if (n_frames > 0) CU_TRY(c, cudaMemcpyAsync(base + o_tw, frame_twc, sizeof(double) * 7 * (size_t)n_frames, cudaMemcpyHostToDevice, s));
  CU_TRY(c, cudaMemcpyAsync(base + o_xyz, out_xyz, sizeof(double) * 3 * (size_t)n_points, cudaMemcpyHostToDevice, s)); // untouched where !ok
  CU_TRY(c, cudaMemsetAsync(base + o_cnt, 0, sizeof(int), s));
  if (!cams_ok(obs_frame, n_obs, n_frames)) { // (threaded range check, behind the uploads already queued)
    cudaStreamSynchronize(s);
    return fail(c, RSPL_BA_ERR_INVALID, "triangulate: keyframe index out of range");
  }
  ba::TriDev d;
  d.n_points = n_points;
  d.n_obs = n_obs;
  d.n_frames = n_frames;
  d.obs_begin = (const int*)(base + o_beg);
  d.obs_frame = (const int*)(base + o_fr);
  d.obs_uv = (const double*)(base + o_uv);
  d.frame_twc = (const double*)(base + o_tw);
  d.fx_inv = 1.0 / cam5[0]; // Camera::_fx_inv (camera.cc)
  d.fy_inv = 1.0 / cam5[1];
  d.cx = cam5[2];
  d.cy = cam5[3];
  d.out_xyz = (double*)(base + o_xyz);
  d.out_ok = (uint8_t*)(base + o_ok);
  d.n_done = (int*)(base + o_cnt);
  {
    ProfScope ps(c, PC_FRAME);
    ba::triangulate_points_kernel<<<(n_points + 127) / 128, 128, 0, s>>>(d);
  }
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  CU_TRY(c, cudaMemcpyAsync(out_xyz, base + o_xyz, sizeof(double) * 3 * (size_t)n_points, cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaMemcpyAsync(out_ok, base + o_ok, (size_t)n_points, cudaMemcpyDeviceToHost, s));
  int cnt = 0;
  CU_TRY(c, cudaMemcpyAsync(&cnt, base + o_cnt, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU_TRY(c, cudaStreamSynchronize(s));
  if (n_done) *n_done = cnt;
  return RSPL_BA_OK;
}

extern "C" int rspl_ba_update_maplines(RsplBaContext* c, int32_t n_lines, const double* line_wd, const int32_t* pt_begin,
                                       const int32_t* pt_index, int32_t n_points, const double* point_xyz,
                                       double* endpoints, uint8_t* out_ok, int32_t* n_done) {
  if (!c || n_lines < 0 || n_points < 0) return RSPL_BA_ERR_INVALID;
  if (n_done) *n_done = 0;
  if (n_lines == 0) return RSPL_BA_OK;
  if (!offsets_ok(pt_begin, n_lines) || !line_wd || !endpoints || !out_ok) return fail(c, RSPL_BA_ERR_INVALID, "update_maplines: bad offsets or null arrays");
  const int n_ref = pt_begin[n_lines];
  if (n_ref > 0 && (!pt_index || !point_xyz)) return fail(c, RSPL_BA_ERR_INVALID, "update_maplines: null point arrays");
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  Arena a;
  const size_t o_wd = a.take(sizeof(double) * 6 * (size_t)n_lines), o_beg = a.take(sizeof(int) * ((size_t)n_lines + 1));
  const size_t o_idx = a.take(sizeof(int) * (size_t)n_ref), o_xyz = a.take(sizeof(double) * 3 * (size_t)n_points);
  const size_t o_end = a.take(sizeof(double) * 6 * (size_t)n_lines), o_ok = a.take((size_t)n_lines), o_cnt = a.take(sizeof(int));
  CU_TRY(c, c->unit_buf.reserve(a.off));
  char* base = c->unit_buf.as<char>();
  cudaStream_t s = c->stream;
  CU_TRY(c, cudaMemcpyAsync(base + o_wd, line_wd, sizeof(double) * 6 * (size_t)n_lines, cudaMemcpyHostToDevice, s));
  CU_TRY(c, cudaMemcpyAsync(base + o_beg, pt_begin, sizeof(int) * ((size_t)n_lines + 1), cudaMemcpyHostToDevice, s));
  if (n_ref > 0) {
    CU_TRY(c, cudaMemcpyAsync(base + o_idx, pt_index, sizeof(int) * (size_t)n_ref, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaMemcpyAsync(base + o_xyz, point_xyz, sizeof(double) * 3 * (size_t)n_points, cudaMemcpyHostToDevice, s));
  }
  CU_TRY(c, cudaMemcpyAsync(base + o_end, endpoints, sizeof(double) * 6 * (size_t)n_lines, cudaMemcpyHostToDevice, s)); // untouched where !ok
  CU_TRY(c, cudaMemsetAsync(base + o_cnt, 0, sizeof(int), s));
  if (!cams_ok(pt_index, n_ref, n_points)) { // (threaded range check, behind the uploads already queued)
    cudaStreamSynchronize(s);
    return fail(c, RSPL_BA_ERR_INVALID, "update_maplines: point index out of range");
  }
  ba::LineEndpointsDev d;
  d.n_lines = n_lines;
  d.n_points = n_points;
  d.line_stride = (size_t)n_lines;
  d.point_stride = (size_t)n_points;
  d.line_wd = (const double*)(base + o_wd);
  d.pt_begin = (const int*)(base + o_beg);
  d.pt_index = (const int*)(base + o_idx);
  d.point_xyz = (const double*)(base + o_xyz);
  d.endpoints = (double*)(base + o_end);
  d.out_ok = (uint8_t*)(base + o_ok);
  d.n_done = (int*)(base + o_cnt);
  {
    ProfScope ps(c, PC_FRAME);
    ba::line_endpoints_kernel<<<(n_lines + ba::LINE_EP_THREADS - 1) / ba::LINE_EP_THREADS, ba::LINE_EP_THREADS, 0, s>>>(d); // a thread per line
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

#include "dense_solver.inl"
#include "comm.inl"
#include "local_capi.inl"
