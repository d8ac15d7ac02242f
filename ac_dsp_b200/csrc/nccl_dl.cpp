// nccl_dl.cpp -- see nccl_dl.h.
#include "nccl_dl.h"

#include <dlfcn.h>
#include <mutex>
#include <stdlib.h>
#include <string>

namespace b2d {

static NcclApi g_api;
static bool g_ok = false;
static std::string g_why;
static std::once_flag g_once;

static void load_once() {
  // B2D_NCCL_LIB overrides; otherwise whatever the process / loader path already has.
  const char *names[] = {getenv("B2D_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    g_why = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
    return;
  }
  struct { const char *sym; void **dst; } tab[] = {
      {"ncclGetUniqueId", (void **)&g_api.GetUniqueId},   {"ncclCommInitRank", (void **)&g_api.CommInitRank},
      {"ncclCommDestroy", (void **)&g_api.CommDestroy},   {"ncclBroadcast", (void **)&g_api.Broadcast},
      {"ncclAllReduce", (void **)&g_api.AllReduce},       {"ncclGetErrorString", (void **)&g_api.GetErrorString},
      {"ncclGetVersion", (void **)&g_api.GetVersion},
  };
  for (auto &t : tab) {
    *t.dst = dlsym(lib, t.sym);
    if (!*t.dst) {
      g_why = std::string("libnccl lacks symbol ") + t.sym;
      return;
    }
  }
  g_ok = true;
}

const NcclApi *nccl_api(const char **why) {
  std::call_once(g_once, load_once);
  if (!g_ok) {
    if (why) *why = g_why.c_str();
    return nullptr;
  }
  return &g_api;
}

}  // namespace b2d
