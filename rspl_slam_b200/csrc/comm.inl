// comm.inl — NCCL communicator of a context, for global BA (part of capi.cu).
//
// one problem, landmarks (with all their edges) partitioned over the ranks, poses
// replicated; the only exchange steps are sum all-reduces of the pose blocks, of the rank-local Schur
// complement pieces and of a handful of scalars per LM trial. NCCL is loaded with dlopen (inside a
// PyTorch process this resolves to the NCCL PyTorch already loaded), so single-GPU use of the library
// never needs it. One process per GPU; the unique id travels through whatever the host already has
// (torch.distributed in bench.py and the tests).
#include <dlfcn.h>

namespace {

// The few NCCL declarations used, spelled out so that the build does not depend on nccl.h
// (values from nccl.h 2.27 / 2.28: ncclInt32 = 2, ncclFloat64 = 8, ncclSum = 0, ncclMax = 2).
struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
enum { kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2 };

struct NcclApi {
  void* lib = nullptr;
  int (*get_unique_id)(NcclUniqueId*) = nullptr;
  int (*comm_init_rank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*comm_destroy)(NcclComm) = nullptr;
  int (*all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*get_error_string)(int) = nullptr;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api;
  if (api.lib) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
  api.get_unique_id = (decltype(api.get_unique_id))dlsym(api.lib, "ncclGetUniqueId");
  api.comm_init_rank = (decltype(api.comm_init_rank))dlsym(api.lib, "ncclCommInitRank");
  api.comm_destroy = (decltype(api.comm_destroy))dlsym(api.lib, "ncclCommDestroy");
  api.all_reduce = (decltype(api.all_reduce))dlsym(api.lib, "ncclAllReduce");
  api.get_error_string = (decltype(api.get_error_string))dlsym(api.lib, "ncclGetErrorString");
  api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce && api.get_error_string;
  return api;
}

// sum (or max) all-reduce on the context's stream; a context without a communicator (or with a
// single rank) degenerates to a device copy so the global code path can run on one GPU
int comm_all_reduce(RsplBaContext* c, const void* send, void* recv, size_t count, int dtype, int op) {
  if (count == 0) return RSPL_BA_OK;
  if (!c->comm || c->comm_ranks <= 1) {
    if (send != recv)
      CU_TRY(c, cudaMemcpyAsync(recv, send, count * (dtype == kNcclInt32 ? 4 : 8), cudaMemcpyDeviceToDevice, c->stream));
    return RSPL_BA_OK;
  }
  NcclApi& api = nccl_api();
  const int rc = api.all_reduce(send, recv, count, dtype, op, (NcclComm)c->comm, c->stream);
  if (rc != 0) return fail(c, RSPL_BA_ERR_CUDA, "ncclAllReduce failed: %s", api.get_error_string(rc));
  c->collectives++;
  return RSPL_BA_OK;
}

} // namespace

static void comm_release(RsplBaContext* c) {
  if (c->comm && nccl_api().ok) nccl_api().comm_destroy((NcclComm)c->comm);
  c->comm = nullptr;
  c->comm_ranks = 1;
  c->comm_rank = 0;
}

extern "C" int rspl_ba_comm_unique_id(void* id128) {
  if (!id128) return RSPL_BA_ERR_INVALID;
  NcclApi& api = nccl_api();
  if (!api.ok) return RSPL_BA_ERR_UNSUPPORTED;
  NcclUniqueId id;
  if (api.get_unique_id(&id) != 0) return RSPL_BA_ERR_CUDA;
  memcpy(id128, &id, sizeof(id));
  return RSPL_BA_OK;
}

extern "C" int rspl_ba_comm_init(RsplBaContext* c, int n_ranks, int rank, const void* id128) {
  if (!c || n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id128)) return RSPL_BA_ERR_INVALID;
  comm_release(c);
  c->comm_ranks = n_ranks;
  c->comm_rank = rank;
  if (n_ranks == 1) return RSPL_BA_OK; // collectives degenerate to copies
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(c, RSPL_BA_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  SetDevice guard(c->device);
  if (!guard.ok) return fail(c, RSPL_BA_ERR_CUDA, "cudaSetDevice failed");
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm comm = nullptr;
  const int rc = api.comm_init_rank(&comm, n_ranks, id, rank);
  if (rc != 0) {
    c->comm_ranks = 1;
    c->comm_rank = 0;
    return fail(c, RSPL_BA_ERR_CUDA, "ncclCommInitRank failed: %s", api.get_error_string(rc));
  }
  c->comm = comm;
  return RSPL_BA_OK;
}

extern "C" int rspl_ba_comm_destroy(RsplBaContext* c) {
  if (!c) return RSPL_BA_ERR_INVALID;
  comm_release(c);
  return RSPL_BA_OK;
}

extern "C" int rspl_ba_comm_size(const RsplBaContext* c) { return c ? c->comm_ranks : 0; }
extern "C" int rspl_ba_comm_rank(const RsplBaContext* c) { return c ? c->comm_rank : -1; }
extern "C" int64_t rspl_ba_collective_count(const RsplBaContext* c) { return c ? c->collectives : 0; }
