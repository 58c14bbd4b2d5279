// comm.cu -- data-parallel exchange of Param gradients: one NCCL sum-allreduce over a flat bucket
// on the library stream (BASELINE.json north_star; no counterpart in the reference, which is
// single-process).  NCCL is resolved with dlopen so the library shares whichever libnccl.so.2 the
// process already holds (torch ships 2.28.9; the system has 2.27.3) and loads without NCCL when
// only one rank is used.
#include "common.cuh"

#include <dlfcn.h>

namespace agb {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0 };

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};
static Nccl g_nccl;

static ag_status nccl_load() {
  if (g_nccl.h) return AG_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  AG_CHECK(g_nccl.h, AG_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                  \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                       \
  AG_CHECK(g_nccl.field, AG_ENCCL, "libnccl is missing symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return AG_OK;
}

#define AG_NCCL(call)                                                                   \
  do {                                                                                  \
    ncclResult_t _r = (call);                                                           \
    if (_r != 0) {                                                                      \
      set_error("NCCL error %d (%s) in %s", _r, g_nccl.GetErrorString(_r), #call);      \
      return AG_ENCCL;                                                                  \
    }                                                                                   \
  } while (0)

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_comm_unique_id(uint8_t* id128) {
  AG_CHECK(id128, AG_EINVAL, "ag_comm_unique_id: null");
  ag_status s = nccl_load();
  if (s != AG_OK) return s;
  ncclUniqueId id;
  AG_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return AG_OK;
}

ag_status ag_comm_init(int32_t nranks, int32_t rank, const uint8_t* id128) {
  AG_REQUIRE_INIT();
  AG_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, AG_EINVAL, "ag_comm_init: rank %d of %d", rank,
           nranks);
  AG_CHECK(!g_nccl.comm, AG_ESTATE, "ag_comm_init: communicator already initialised");
  g_nccl.nranks = nranks;
  g_nccl.rank = rank;
  if (nranks == 1) return AG_OK;
  AG_CHECK(id128, AG_EINVAL, "ag_comm_init: null id");
  ag_status s = nccl_load();
  if (s != AG_OK) return s;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  AG_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank));
  return AG_OK;
}

ag_status ag_comm_destroy(void) {
  if (g_nccl.comm) {
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  g_nccl.nranks = 1;
  g_nccl.rank = 0;
  return AG_OK;
}

ag_status ag_allreduce_sum(int32_t dtype, void* buf, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0 || g_nccl.nranks == 1) return AG_OK;
  AG_CHECK(g_nccl.comm, AG_ESTATE, "ag_allreduce_sum: call ag_comm_init first");
  AG_CHECK(buf && n > 0, AG_EINVAL, "ag_allreduce_sum: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_allreduce_sum: dtype %d", dtype);
  AG_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, dtype == AG_F32 ? ncclFloat32 : ncclFloat64, ncclSum,
                           g_nccl.comm, ctx().stream));
  ctx().launches += 1;
  return AG_OK;
}

}  // extern "C"
