// nccl_dyn.h -- NCCL resolved at run time (dlopen) so that single-GPU use needs no NCCL at all and a
// multi-rank run shares whatever libnccl.so.2 the host process (e.g. torch.distributed) already loaded.
// Only the handful of entry points the Fourier transpose and the global-mean fixers need.
#pragma once
#include <dlfcn.h>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace isca {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  void load() {
    if (lib) return;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) throw std::runtime_error(std::string("missing NCCL symbol ") + n); return p; };
    GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    Send = (decltype(Send))sym("ncclSend");
    Recv = (decltype(Recv))sym("ncclRecv");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
  }
  void ck(int rc, const char* what) const {
    if (rc != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(rc) : "?"));
  }
};

}  // namespace isca
