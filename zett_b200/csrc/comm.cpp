// libzett_b200.so -- multi-GPU half of the C ABI (include/zett_b200.h): one process per GPU, the predicted rows of every
// rank assembled on every rank by ncclAllGather over NVLink / NVSwitch.
//
// Replaces the reference's device sharding of an inference batch (zett/utils.py:26 PositionalSharding over the local
// devices; scripts/transfer.py:90-91, 105-111: jax.device_put of the batch with that sharding, results pulled back with
// jax.device_get): rows are independent, so the only exchange of the path is this gather.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring the copy already loaded into the process -- PyTorch
// brings its own), so the library still loads on a machine without NCCL and single-GPU callers never touch it.
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/zett_b200.h"

extern "C" void zett_set_last_error_(const char* msg);

namespace {

struct NcclId { char internal[128]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
using NcclComm = void*;
constexpr int kNcclFloat32 = 7;          // ncclDataType_t::ncclFloat32

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string error;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);   // the copy PyTorch (or the host application) already loaded
      if (api.lib) break;
    }
    for (const char* n : names) {
      if (api.lib) break;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!api.lib) { api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
    auto sym = [&](const char* s) { void* p = dlsym(api.lib, s); if (!p) api.error = std::string("missing NCCL symbol ") + s; return p; };
    api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclId, int)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, void*)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
  });
  return &api;
}

int fail(int code, const std::string& msg) {
  zett_set_last_error_(msg.c_str());
  return code;
}

int nccl_fail(NcclApi* a, const char* what, int rc) {
  return fail(ZETT_ERR_CUDA, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}

}  // namespace

struct zett_comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
};

extern "C" {

int zett_comm_unique_id(void* out_id_128_bytes) {
  if (!out_id_128_bytes) return fail(ZETT_ERR_INVALID, "null argument");
  NcclApi* a = nccl();
  if (!a->error.empty()) return fail(ZETT_ERR_CUDA, a->error);
  NcclId id;
  const int rc = a->GetUniqueId(&id);
  if (rc != 0) return nccl_fail(a, "ncclGetUniqueId", rc);
  memcpy(out_id_128_bytes, id.internal, sizeof id.internal);
  return ZETT_OK;
}

int zett_comm_init(int rank, int world, const void* nccl_unique_id, zett_comm** out) {
  if (!out || world < 1 || rank < 0 || rank >= world) return fail(ZETT_ERR_INVALID, "bad rank / world size");
  auto* c = new zett_comm();
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    if (!nccl_unique_id) { delete c; return fail(ZETT_ERR_INVALID, "nccl_unique_id is required for world > 1"); }
    NcclApi* a = nccl();
    if (!a->error.empty()) { delete c; return fail(ZETT_ERR_CUDA, a->error); }
    NcclId id;
    memcpy(id.internal, nccl_unique_id, sizeof id.internal);
    const int rc = a->CommInitRank(&c->comm, world, id, rank);   // on the CUDA device current in this process
    if (rc != 0) { delete c; return nccl_fail(a, "ncclCommInitRank", rc); }
  }
  *out = c;
  return ZETT_OK;
}

int zett_allgather_rows(zett_comm* c, const float* shard_dev, int64_t rows_per_rank, int64_t row_elems, float* full_dev,
                        void* cuda_stream) {
  if (!c || !shard_dev || !full_dev || rows_per_rank < 0 || row_elems <= 0) return fail(ZETT_ERR_INVALID, "bad argument");
  if (rows_per_rank == 0) return ZETT_OK;
  const size_t count = static_cast<size_t>(rows_per_rank) * static_cast<size_t>(row_elems);
  if (c->world == 1) {
    if (shard_dev == full_dev) return ZETT_OK;
    return fail(ZETT_ERR_INVALID, "world == 1: pass the same buffer as shard and full (nothing to gather)");
  }
  NcclApi* a = nccl();
  const int rc = a->AllGather(shard_dev, full_dev, count, kNcclFloat32, c->comm, cuda_stream);
  if (rc != 0) return nccl_fail(a, "ncclAllGather", rc);
  return ZETT_OK;
}

int zett_comm_info(const zett_comm* c, int* rank, int* world, int* nccl_version) {
  if (!c) return fail(ZETT_ERR_INVALID, "null communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (nccl_version) {
    *nccl_version = 0;
    NcclApi* a = c->world > 1 ? nccl() : nullptr;
    if (a && a->GetVersion) a->GetVersion(nccl_version);
  }
  return ZETT_OK;
}

void zett_comm_destroy(zett_comm* c) {
  if (!c) return;
  if (c->comm) nccl()->CommDestroy(c->comm);
  delete c;
}

}  // extern "C"
