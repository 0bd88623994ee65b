// comm.cu -- the path's one exchange: integer sum of the per-cell histogram arena over the ranks.
//
// Replaces (reference): nothing -- the reference is single-process (platform/backend/main.py:109-118 builds one set of
// objects per connection).  The sweep shards (cell x image block) work items over one process per GPU; at the end every rank
// holds partial int64 histograms and one ncclAllReduce(sum, int64) over NVLink / NVSwitch makes them global.  Integer payload:
// the result does not depend on the reduction order or on the number of ranks.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 that PyTorch already loaded, or of the path in FAV_NCCL_LIB), so the
// library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace fav {

typedef void* nccl_comm_t;
struct nccl_uid { char internal[128]; };
typedef int (*nccl_get_uid_fn)(nccl_uid*);
typedef int (*nccl_init_rank_fn)(nccl_comm_t*, int, nccl_uid, int);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*nccl_destroy_fn)(nccl_comm_t);
typedef const char* (*nccl_errstr_fn)(int);
constexpr int NCCL_INT64 = 4, NCCL_SUM = 0;      // ncclDataType_t ncclInt64, ncclRedOp_t ncclSum (nccl.h, stable ABI values)

struct NcclApi {
  void* lib = nullptr;
  nccl_get_uid_fn get_uid = nullptr;
  nccl_init_rank_fn init_rank = nullptr;
  nccl_allreduce_fn allreduce = nullptr;
  nccl_destroy_fn destroy = nullptr;
  nccl_errstr_fn errstr = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* env = getenv("FAV_NCCL_LIB");
  void* lib = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy PyTorch already mapped, if any
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return nullptr;
  api.get_uid = reinterpret_cast<nccl_get_uid_fn>(dlsym(lib, "ncclGetUniqueId"));
  api.init_rank = reinterpret_cast<nccl_init_rank_fn>(dlsym(lib, "ncclCommInitRank"));
  api.allreduce = reinterpret_cast<nccl_allreduce_fn>(dlsym(lib, "ncclAllReduce"));
  api.destroy = reinterpret_cast<nccl_destroy_fn>(dlsym(lib, "ncclCommDestroy"));
  api.errstr = reinterpret_cast<nccl_errstr_fn>(dlsym(lib, "ncclGetErrorString"));
  if (!api.get_uid || !api.init_rank || !api.allreduce || !api.destroy) return nullptr;
  api.lib = lib;
  return &api;
}

void comm_destroy(Ctx* ctx) {
  NcclApi* api = nccl_api();
  if (ctx->nccl_comm && api) api->destroy(reinterpret_cast<nccl_comm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  ctx->world = 1;
  ctx->rank = 0;
}

}  // namespace fav

using namespace fav;

#define FAV_NCCL_OK(expr)                                                                              \
  do {                                                                                                 \
    const int _r = (expr);                                                                             \
    if (_r != 0) {                                                                                     \
      set_error("%s failed: %s", #expr, api->errstr ? api->errstr(_r) : "NCCL error");               \
      return FAV_E_CUDA;                                                                               \
    }                                                                                                  \
  } while (0)

extern "C" int fav_comm_unique_id(void* out128) {
  FAV_REQUIRE(out128, "fav_comm_unique_id: null pointer");
  NcclApi* api = nccl_api();
  FAV_REQUIRE(api, "fav_comm_unique_id: libnccl.so.2 not found (set FAV_NCCL_LIB)");
  nccl_uid id;
  FAV_NCCL_OK(api->get_uid(&id));
  memcpy(out128, &id, sizeof(id));
  return FAV_OK;
}

extern "C" int fav_comm_init(fav_handle h, const void* id128, int rank, int world_size) {
  FAV_REQUIRE(h && id128, "fav_comm_init: null pointer");
  FAV_DEVICE(h);
  FAV_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "fav_comm_init: bad rank %d of %d", rank, world_size);
  comm_destroy(h);
  if (world_size == 1) return FAV_OK;
  NcclApi* api = nccl_api();
  FAV_REQUIRE(api, "fav_comm_init: libnccl.so.2 not found (set FAV_NCCL_LIB)");
  FAV_CUDA_OK(cudaSetDevice(h->device));
  nccl_uid id;
  memcpy(&id, id128, sizeof(id));
  nccl_comm_t comm = nullptr;
  FAV_NCCL_OK(api->init_rank(&comm, world_size, id, rank));
  h->nccl_comm = comm;
  h->world = world_size;
  h->rank = rank;
  return FAV_OK;
}

extern "C" int fav_allreduce(fav_handle h, int64_t* d_hist, size_t count, void* stream) {
  FAV_REQUIRE(h, "null handle");
  FAV_DEVICE(h);
  if (h->world <= 1 || count == 0) return FAV_OK;                 // single rank: nothing to exchange
  FAV_REQUIRE(d_hist, "fav_allreduce: null pointer");
  FAV_REQUIRE(h->nccl_comm, "fav_allreduce: fav_comm_init has not been called");
  NcclApi* api = nccl_api();
  FAV_REQUIRE(api, "fav_allreduce: libnccl.so.2 not found");
  FAV_NCCL_OK(api->allreduce(d_hist, d_hist, count, NCCL_INT64, NCCL_SUM, reinterpret_cast<nccl_comm_t>(h->nccl_comm),
                             reinterpret_cast<cudaStream_t>(stream)));
  return FAV_OK;
}
