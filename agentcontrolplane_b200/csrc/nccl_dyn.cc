#include "nccl_dyn.h"
#include <dlfcn.h>
#include <stdio.h>
#include <mutex>

namespace acp {

const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      fprintf(stderr, "[acp_infer] dlopen(libnccl.so.2) failed: %s\n", dlerror());
      return;
    }
    api.CommInitAll = (int (*)(NcclComm*, int, const int*))dlsym(h, "ncclCommInitAll");
    api.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
    api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    api.ok = api.CommInitAll && api.CommDestroy && api.AllReduce && api.AllGather && api.GroupStart &&
             api.GroupEnd && api.GetErrorString;
    if (!api.ok) fprintf(stderr, "[acp_infer] libnccl.so.2 lacks required symbols\n");
  });
  return api;
}

}  // namespace acp
