// NCCL, bound at run time.  The library is only needed by the multi-GPU path
// (gbp_cuda_init_shard); resolving it with dlopen keeps libgbp_cuda.so free of a
// link-time NCCL dependency and makes it share the NCCL instance the process
// already loaded (e.g. the one torch.distributed brought in) instead of pulling
// a second copy with the same soname.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <string>

namespace gbp {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;

  bool load() {
    if (lib) return true;
    const char* env = std::getenv("GBP_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) {
      error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
      return false;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(lib, s);
      if (!p) error = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    AllGather = (decltype(AllGather))sym("ncclAllGather");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString) {
      lib = nullptr;
      return false;
    }
    return true;
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

}  // namespace gbp
