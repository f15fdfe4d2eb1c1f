// comm.cuh -- the one collective step of the render path: the gather of the stripes'
// rows into the root's frame (C1), plus the one-off scene broadcast (C0).  NCCL over NVLink 5 /
// NVSwitch; nothing else crosses GPUs -- pixels are independent given the global depth order, so
// the frame shards by screen-tile stripes with no data-path exchange.
//
// The reference has no multi-device path at all (its only parallelism is euc's row-group
// threading); this is net-new behind the same C ABI.
//
// libnccl.so.2 is dlopen'ed on first use, so a single-GPU caller never needs NCCL installed, and
// a process that already carries an NCCL (a torch launcher) shares that copy.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; no link-time dependency

#include <cstdint>
#include <cstdio>
#include <string>

namespace splat {

struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string error;
};

// nullptr (and api.error set) if NCCL cannot be loaded
inline NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
    return nullptr;
  }
  bool ok = true;
  auto sym = [&](const char *name) {
    void *p = dlsym(api.handle, name);
    if (!p) { ok = false; api.error = std::string("libnccl lacks ") + name; }
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  if (!ok) { dlclose(api.handle); api.handle = nullptr; return nullptr; }
  return &api;
}

// C1.  Every rank's rows [bounds[2r], bounds[2r+1]) of its W x H frame -> the same rows of root's
// frame, in place: one grouped ncclSend / ncclRecv per non-empty stripe, enqueued on `stream`
// right behind the blend kernel that wrote the rows (no staging copy; K5 writes straight into
// the buffer NCCL reads).  8.3 MB per 1080p frame, 33 MB at 4K, in total.
inline ncclResult_t gather_stripes(NcclApi *N, ncclComm_t comm, int n_ranks, int rank, int root, uint32_t *fb, uint32_t W,
                                   const uint32_t *bounds, cudaStream_t stream) {
  if (n_ranks <= 1) return ncclSuccess;
  ncclResult_t rc = N->GroupStart();
  if (rc != ncclSuccess) return rc;
  if (rank == root) {
    for (int r = 0; r < n_ranks && rc == ncclSuccess; ++r) {
      const uint32_t r0 = bounds[2 * r], r1 = bounds[2 * r + 1];
      if (r == root || r1 <= r0) continue;
      rc = N->Recv(fb + (size_t)r0 * W, (size_t)(r1 - r0) * W, ncclUint32, r, comm, stream);
    }
  } else {
    const uint32_t r0 = bounds[2 * rank], r1 = bounds[2 * rank + 1];
    if (r1 > r0) rc = N->Send(fb + (size_t)r0 * W, (size_t)(r1 - r0) * W, ncclUint32, root, comm, stream);
  }
  const ncclResult_t rc2 = N->GroupEnd();
  return rc != ncclSuccess ? rc : rc2;
}

}  // namespace splat
