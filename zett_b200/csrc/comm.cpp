// libzett_b200.so -- multi-GPU half of the C ABI (include/zett_b200.h): one process per GPU, the predicted rows of every
// rank assembled on every rank by ncclAllGather over NVLink / NVSwitch.
//
// Replaces the reference's device sharding of an inference batch (zett/utils.py:26 PositionalSharding over the local
// devices; scripts/transfer.py:90-91, 105-111: jax.device_put of the batch with that sharding, results pulled back with
// jax.device_get): rows are independent, so the only exchange of the path is this gather.
//
// Two transports behind zett_allgather_rows:
//   * ncclAllGather (the default until a buffer is registered);
//   * PEER COPIES over NVLink / NVSwitch for a registered full matrix (zett_comm_register): every rank pushes its rows into the
//     same slot of every peer's matrix with cudaMemcpyAsync on the caller's stream -- copy engines, no SM.  The forward's
//     GEMMs are persistent kernels that own all 148 SMs with a static tile assignment; a NCCL kernel that holds a few SMs
//     while one of them starts keeps some CTA pairs of that GEMM waiting for as long as the collective runs, which is what
//     an "overlapped" collective must not do.  Pushes complete locally; zett_comm_barrier (a one-element ncclAllReduce)
//     orders them across ranks.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring the copy already loaded into the process -- PyTorch
// brings its own), so the library still loads on a machine without NCCL and single-GPU callers never touch it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

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
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, void*) = nullptr;
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
    api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclComm, void*)>(sym("ncclAllReduce"));
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
  // registered full matrix (peer copies): this rank's buffer and the same buffer of every peer, mapped through CUDA IPC
  char* reg_base = nullptr;
  size_t reg_bytes = 0;
  std::vector<char*> peer_base;      // [world]; own entry = reg_base
  std::vector<void*> peer_mapping;   // [world]; what cudaIpcOpenMemHandle returned (to close), nullptr for the own entry
  int* barrier_word = nullptr;       // device int for zett_comm_barrier
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
  const char* full = reinterpret_cast<const char*>(full_dev);
  if (c->reg_base && full >= c->reg_base && full + count * c->world * sizeof(float) <= c->reg_base + c->reg_bytes) {
    // peer copies: this rank's rows -> its slot in every peer's copy of the registered matrix (copy engines over NVLink)
    const size_t off = static_cast<size_t>(full - c->reg_base) + static_cast<size_t>(c->rank) * count * sizeof(float);
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    if (reinterpret_cast<const char*>(shard_dev) != c->reg_base + off) {
      cudaError_t e = cudaMemcpyAsync(c->reg_base + off, shard_dev, count * sizeof(float), cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return fail(ZETT_ERR_CUDA, std::string("cudaMemcpyAsync (own slot): ") + cudaGetErrorString(e));
    }
    for (int step = 1; step < c->world; ++step) {
      const int peer = (c->rank + step) % c->world;   // every rank starts with a different peer
      cudaError_t e = cudaMemcpyAsync(c->peer_base[peer] + off, shard_dev, count * sizeof(float), cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return fail(ZETT_ERR_CUDA, std::string("cudaMemcpyAsync (peer): ") + cudaGetErrorString(e));
    }
    return ZETT_OK;
  }
  NcclApi* a = nccl();
  const int rc = a->AllGather(shard_dev, full_dev, count, kNcclFloat32, c->comm, cuda_stream);
  if (rc != 0) return nccl_fail(a, "ncclAllGather", rc);
  return ZETT_OK;
}

int zett_comm_ipc_handle(const void* dev_ptr, void* out_handle_64_bytes, int64_t* out_offset) {
  if (!dev_ptr || !out_handle_64_bytes || !out_offset) return fail(ZETT_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  // the handle names the whole allocation; the pointer may sit inside it (a framework's caching allocator)
  using RangeFn = int (*)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess)
    return fail(ZETT_ERR_CUDA, "cuMemGetAddressRange not available");
  unsigned long long base = 0;
  size_t size = 0;
  if (reinterpret_cast<RangeFn>(fn)(&base, &size, reinterpret_cast<unsigned long long>(dev_ptr)) != 0)
    return fail(ZETT_ERR_CUDA, "cuMemGetAddressRange failed: not a device allocation");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ZETT_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(out_handle_64_bytes, &h, 64);
  *out_offset = static_cast<int64_t>(reinterpret_cast<unsigned long long>(dev_ptr) - base);
  return ZETT_OK;
}

int zett_comm_unregister(zett_comm* c) {
  if (!c) return fail(ZETT_ERR_INVALID, "null communicator");
  for (void* m : c->peer_mapping)
    if (m) cudaIpcCloseMemHandle(m);
  c->peer_mapping.clear();
  c->peer_base.clear();
  c->reg_base = nullptr;
  c->reg_bytes = 0;
  return ZETT_OK;
}

int zett_comm_register(zett_comm* c, void* full_dev, int64_t bytes, const void* all_handles, const int64_t* all_offsets) {
  if (!c || !full_dev || bytes <= 0) return fail(ZETT_ERR_INVALID, "bad argument");
  zett_comm_unregister(c);
  if (c->world == 1) return ZETT_OK;
  if (!all_handles || !all_offsets) return fail(ZETT_ERR_INVALID, "handles of all ranks are required");
  c->peer_base.assign(static_cast<size_t>(c->world), nullptr);
  c->peer_mapping.assign(static_cast<size_t>(c->world), nullptr);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer_base[r] = static_cast<char*>(full_dev); continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(all_handles) + 64 * static_cast<size_t>(r), 64);
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();   // not sticky: clear it, the caller falls back to the NCCL transport
      zett_comm_unregister(c);
      return fail(ZETT_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
    }
    c->peer_mapping[r] = base;
    c->peer_base[r] = static_cast<char*>(base) + all_offsets[r];
  }
  c->reg_base = static_cast<char*>(full_dev);
  c->reg_bytes = static_cast<size_t>(bytes);
  return ZETT_OK;
}

int zett_comm_barrier(zett_comm* c, void* cuda_stream) {
  if (!c) return fail(ZETT_ERR_INVALID, "null communicator");
  if (c->world == 1) return ZETT_OK;
  if (!c->barrier_word) {
    if (cudaMalloc(reinterpret_cast<void**>(&c->barrier_word), 256) != cudaSuccess || cudaMemset(c->barrier_word, 0, 256) != cudaSuccess)
      return fail(ZETT_ERR_CUDA, "cannot allocate the barrier word");
  }
  NcclApi* a = nccl();
  const int rc = a->AllReduce(c->barrier_word, c->barrier_word, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->comm, cuda_stream);
  if (rc != 0) return nccl_fail(a, "ncclAllReduce (barrier)", rc);
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
  zett_comm_unregister(c);
  if (c->barrier_word) cudaFree(c->barrier_word);
  if (c->comm) nccl()->CommDestroy(c->comm);
  delete c;
}

}  // extern "C"
