// libdsmgp.so : the collective of a multi-GPU evaluation, inside the library.
//
// Leaves are independent given theta (fit.jl:88-119, 306-311); the only exchange step of an evaluation is the table of
// per-leaf rows [mll(gp), nabla-mll(gp)...] that the O(L) tree passes (optimize.jl:27-89) need.  Every rank fills the rows of
// its own experts and leaves the others zero, so ONE NCCL SUM all-reduce over NVLink assembles the table on every rank.
// NCCL is bound at run time with dlopen: single-GPU users need no NCCL at all, and a process that already carries a copy
// (e.g. the one bundled with PyTorch) shares it instead of loading a second one.
#include <dlfcn.h>

#include "handle.h"

using namespace dsm;
#define g_create_error (dsm::create_error())

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[DSMGP_COMM_ID_BYTES]; };      // ncclUniqueId: 128 opaque bytes (nccl.h)
constexpr int kNcclFloat64 = 8;                                      // ncclDataType_t ncclFloat64 / ncclDouble
constexpr int kNcclSum = 0;                                          // ncclRedOp_t ncclSum

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
std::mutex g_nccl_mu;
NcclApi g_nccl;

bool nccl_load(std::string& err) {
  std::lock_guard<std::mutex> g(g_nccl_mu);
  if (g_nccl.AllReduce) return true;
  const char* env = getenv("DSMGP_NCCL_LIB");
  void* lib = nullptr;
  if (env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  // a copy that is already mapped into the process (PyTorch bundles one) wins over loading a second NCCL
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!lib && dlsym(RTLD_DEFAULT, "ncclAllReduce")) lib = dlopen(nullptr, RTLD_NOW);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { err = std::string("cannot load NCCL (libnccl.so.2; set DSMGP_NCCL_LIB): ") + (dlerror() ? dlerror() : "not found"); return false; }
  g_nccl.lib = lib;
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    g_nccl.AllReduce = nullptr;
    err = "NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy";
    return false;
  }
  return true;
}

std::string nccl_msg(const char* what, int rc) {
  return std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")";
}

}  // namespace

void dsm::comm_destroy(void* comm) {
  if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(comm));
}

int32_t dsm::comm_allreduce_sum(dsmgp_handle* h, double* dev_buf, size_t count) {
  if (!h->comm || !g_nccl.AllReduce) { h->err = "no communicator (dsmgp_comm_init)"; return DSMGP_ERR_COMM; }
  const int rc = g_nccl.AllReduce(dev_buf, dev_buf, count, kNcclFloat64, kNcclSum, static_cast<ncclComm_t>(h->comm), h->stream);
  if (rc != 0) { h->err = nccl_msg("ncclAllReduce", rc); return DSMGP_ERR_COMM; }
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_comm_unique_id(void* id) {
  if (!id) return DSMGP_ERR_ARG;
  if (!nccl_load(g_create_error)) return DSMGP_ERR_COMM;
  NcclUniqueId u;
  const int rc = g_nccl.GetUniqueId(&u);
  if (rc != 0) { g_create_error = nccl_msg("ncclGetUniqueId", rc); return DSMGP_ERR_COMM; }
  memcpy(id, &u, sizeof(u));
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_comm_init(dsmgp_handle* h, const void* id) {
  if (!h) return DSMGP_ERR_ARG;
  if (!id) { h->err = "comm_init: null id"; return DSMGP_ERR_ARG; }
  if (!nccl_load(h->err)) return DSMGP_ERR_COMM;
  if (h->comm) { comm_destroy(h->comm); h->comm = nullptr; }
  cudaSetDevice(h->device);
  NcclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t c = nullptr;
  const int rc = g_nccl.CommInitRank(&c, h->opts.world, u, h->opts.rank);
  if (rc != 0) { h->err = nccl_msg("ncclCommInitRank", rc); return DSMGP_ERR_COMM; }
  h->comm = c;
  return DSMGP_OK;
}
