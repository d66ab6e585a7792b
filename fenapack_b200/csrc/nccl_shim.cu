// NCCL is bound at run time (dlopen), only when a multi-rank context is created.
// Reason: a host process that also uses PyTorch already carries its own
// libnccl.so.2; linking a second copy by soname at load time would make the two
// shadow each other.  Search order: a libnccl.so.2 already mapped into the
// process, $FNP_NCCL_LIBRARY, then the system library path.
#include <dlfcn.h>

#include <cstdlib>

#include "fnp_internal.cuh"

namespace fnp {

static NcclApi g_api;
static bool g_loaded = false;

template <class F>
static void bind(void *h, F &fn, const char *name) {
  fn = reinterpret_cast<F>(dlsym(h, name));
  FNP_REQUIRE(fn != nullptr, FNP_ERR_NCCL, std::string("NCCL symbol not found: ") + name);
}

const NcclApi &nccl() {
  if (g_loaded) return g_api;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) {
    const char *env = std::getenv("FNP_NCCL_LIBRARY");
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
  }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  FNP_REQUIRE(h != nullptr, FNP_ERR_NCCL, std::string("cannot load NCCL: ") + dlerror());
  bind(h, g_api.GetUniqueId, "ncclGetUniqueId");
  bind(h, g_api.CommInitRank, "ncclCommInitRank");
  bind(h, g_api.CommDestroy, "ncclCommDestroy");
  g_api.CommSplit = reinterpret_cast<decltype(g_api.CommSplit)>(dlsym(h, "ncclCommSplit"));   // optional (NCCL >= 2.18)
  bind(h, g_api.AllReduce, "ncclAllReduce");
  bind(h, g_api.AllGather, "ncclAllGather");
  bind(h, g_api.Send, "ncclSend");
  bind(h, g_api.Recv, "ncclRecv");
  bind(h, g_api.GroupStart, "ncclGroupStart");
  bind(h, g_api.GroupEnd, "ncclGroupEnd");
  bind(h, g_api.GetErrorString, "ncclGetErrorString");
  g_loaded = true;
  return g_api;
}

}  // namespace fnp
