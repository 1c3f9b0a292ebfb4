// hb_comm.cu -- NCCL over NVLink for the per-generation chain-state exchange (one process per GPU).
// libnccl.so.2 is opened lazily with dlopen: the single-GPU path has no NCCL dependency, and inside a Python
// process that already imported torch the loader returns torch's bundled NCCL (same SONAME).
#include <dlfcn.h>
#include <string.h>

#include <string>

#include "hb_internal.h"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint32 = 3, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct api {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
api g;

int load(std::string& err) {
  if (g.lib) return HB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g.lib) break;
  }
  if (!g.lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return HB_ERR_NCCL; }
#define HB_SYM(field, name)                                                    \
  *(void**)(&g.field) = dlsym(g.lib, name);                                    \
  if (!g.field) { err = std::string("libnccl lacks ") + name; return HB_ERR_NCCL; }
  HB_SYM(GetUniqueId, "ncclGetUniqueId")
  HB_SYM(CommInitRank, "ncclCommInitRank")
  HB_SYM(CommDestroy, "ncclCommDestroy")
  HB_SYM(AllGather, "ncclAllGather")
  HB_SYM(AllReduce, "ncclAllReduce")
  HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
  return HB_OK;
}

int check(int rc, const char* what, std::string& err) {
  if (rc == ncclSuccess) return HB_OK;
  err = std::string(what) + ": " + (g.GetErrorString ? g.GetErrorString(rc) : "nccl error");
  return HB_ERR_NCCL;
}
}  // namespace

struct hb_nccl { ncclComm_t comm = nullptr; int rank = 0, world = 1; };

int hb_nccl_unique_id(void* id128, std::string& err) {
  if (int rc = load(err)) return rc;
  ncclUniqueId id;
  if (int rc = check(g.GetUniqueId(&id), "ncclGetUniqueId", err)) return rc;
  memcpy(id128, &id, sizeof id);
  return HB_OK;
}

int hb_nccl_init(hb_nccl** out, int rank, int world, const void* id128, std::string& err) {
  if (int rc = load(err)) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  auto* c = new hb_nccl();
  c->rank = rank; c->world = world;
  if (int rc = check(g.CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank", err)) { delete c; return rc; }
  *out = c;
  return HB_OK;
}

void hb_nccl_destroy(hb_nccl* c) {
  if (!c) return;
  if (c->comm && g.CommDestroy) g.CommDestroy(c->comm);
  delete c;
}

int hb_nccl_allgather_f64(hb_nccl* c, const double* send, double* recv, size_t count, cudaStream_t st, std::string& err) {
  return check(g.AllGather(send, recv, count, ncclFloat64, c->comm, st), "ncclAllGather", err);
}
int hb_nccl_allreduce_f64(hb_nccl* c, double* buf, size_t count, cudaStream_t st, std::string& err) {
  return check(g.AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->comm, st), "ncclAllReduce", err);
}
int hb_nccl_allreduce_u64(hb_nccl* c, unsigned long long* buf, size_t count, cudaStream_t st, std::string& err) {
  return check(g.AllReduce(buf, buf, count, ncclUint64, ncclSum, c->comm, st), "ncclAllReduce", err);
}
